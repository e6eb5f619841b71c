"""wgrad variants on the bench shapes: default dispatch, column-box forced on/off, block_n 192/256; cross-checked."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv1_2", 240, 64, 64), ("conv2_1", 120, 64, 128), ("conv2_2", 120, 128, 128), ("conv3_1", 60, 128, 256),
          ("conv3_2", 60, 256, 256), ("conv4_1", 30, 256, 512), ("conv4_2", 30, 512, 512)]
CFG = [("default", {}, 0), ("colbox off", {"DBX_COLBOX": "0"}, 0), ("colbox on", {"DBX_COLBOX": "1"}, 0),
       ("bn256 (pairs)", {"DBX_COLBOX": "0"}, 256), ("bn192", {"DBX_COLBOX": "0"}, 192)]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(200):
    a @ a
torch.cuda.synchronize()
for name, H, cin, cout in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    dy = (torch.randn(B, H, H, cout, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    flops = 2.0 * B * H * H * cin * cout * 9
    ref = None
    for label, env, bn in CFG:
        os.environ.pop("DBX_COLBOX", None)
        os.environ.update(env)
        dw = torch.zeros(cout, 9 * cin, device="cuda")
        try:
            ops.conv_wgrad(x, dy, 3, 3, 1, dw, block_n=bn)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("%-8s %-14s FAILED %s" % (name, label, str(e)[:80]), flush=True)
            continue
        if ref is None:
            ref = dw.clone()
        err = ((dw - ref).abs().max() / ref.abs().max()).item()
        t = timeit(lambda: ops.conv_wgrad(x, dy, 3, 3, 1, dw, block_n=bn), n=20)
        print("%-8s %-14s %7.4f ms %7.1f TFLOP/s  rel maxdiff vs default %.2g" % (name, label, t, flops / t * 1e-9, err), flush=True)
