"""A/B of the TMA-staged epilogue vs the direct (registers -> global, no block barrier) epilogue on the layers whose
pace is set by the epilogue."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv1_1 pairs", 240, 64, 128, 1, 120), ("conv1_2", 240, 64, 64, 3, 240), ("conv2_1 fprop", 120, 64, 128, 3, 120),
          ("conv2_1 dgrad", 120, 128, 64, 3, 120), ("conv2_2", 120, 128, 128, 3, 120), ("conv3_2", 60, 256, 256, 3, 60),
          ("heads2 dgrad", 60, 64, 1024, 1, 60)]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(200):
    a @ a
torch.cuda.synchronize()
for name, H, cin, cout, R, W in SHAPES:
    x = torch.randn(B, H, W, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    aux = torch.randn(B, H, W, cout, generator=g, device="cuda").to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda")
    flops = 2.0 * B * H * W * cin * cout * R * R
    ref = {}
    for label, env in (("tma-staged", "0"), ("direct", "1")):
        os.environ["DBX_DIRECT_EPI"] = env
        for am in (0, 1):
            fn = lambda: ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, relu=am == 0, aux=aux if am else None, aux_mode=am)
            out.zero_()
            fn()
            torch.cuda.synchronize()
            o = out.float().clone()
            ref.setdefault(am, o)
            err = (o - ref[am]).abs().max().item()
            t = timeit(fn, n=20)
            print("%-14s %-11s aux=%d %7.4f ms %7.1f TFLOP/s  %6.2f TB/s out  maxdiff %.3g" % (
                name, label, am, t, flops / t * 1e-9, out.numel() * 2 / t * 1e-9, err), flush=True)
