import json,sys
d=json.load(sys.stdin); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {k:v["ms"] for k,v in d["kernels"].items()})
