"""A/B of the column-box fprop mode (DBX_COLBOX_FPROP) on the narrow 3x3 layers; checks it against the generic mode."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv1_2", 240, 64, 64), ("conv2_1 fprop", 120, 64, 128), ("conv2_1 dgrad", 120, 128, 64),
          ("conv2_2", 120, 128, 128), ("conv3_1 dgrad", 60, 256, 128), ("conv3_1 fprop", 60, 128, 256),
          ("conv3_2", 60, 256, 256), ("conv4_2", 30, 512, 512)]
CFG = [  # label, env
    ("default", {}),
    ("colbox nbuf2", {"DBX_COLBOX_FPROP": "1", "DBX_EPI_BUFS": "2", "DBX_HALO": "0"}),
    ("colbox nbuf4", {"DBX_COLBOX_FPROP": "1", "DBX_EPI_BUFS": "4", "DBX_HALO": "0"}),
    ("colbox nbuf8", {"DBX_COLBOX_FPROP": "1", "DBX_EPI_BUFS": "8", "DBX_HALO": "0"}),
    ("halo", {"DBX_COLBOX_FPROP": "0", "DBX_HALO": "1"}),
]
KEYS = sorted({k for _, e in CFG for k in e})
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(200):
    a @ a
torch.cuda.synchronize()
for name, H, cin, cout in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, 9 * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    aux = torch.randn(B, H, H, cout, generator=g, device="cuda").to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda")
    flops = 2.0 * B * H * H * cin * cout * 9
    ref = None
    for label, env in CFG:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        for am in (0, 1):
            out.zero_()
            fn = lambda: ops.conv_fprop(x, wk, 3, 3, 1, out, bias=bias, relu=am == 0, aux=aux if am else None, aux_mode=am)
            try:
                fn()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("%-14s %-18s aux=%d FAILED %s" % (name, label, am, e), flush=True)
                continue
            o = out.float().clone()
            if ref is None or am not in ref:
                ref = ref or {}
                ref[am] = o
            err = (o - ref[am]).abs().max().item()
            t = timeit(fn, n=20)
            print("%-14s %-18s aux=%d %7.4f ms %7.1f TFLOP/s  maxdiff vs generic %.3g" % (name, label, am, t, flops / t * 1e-9, err),
                  flush=True)
