"""Throughput of one fprop shape as a function of pipeline depth (DBX_STAGES) — latency-bound or not?"""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv2_2", 120, 128, 128, 3), ("conv2_1 dgrad", 120, 128, 64, 3), ("conv3_2", 60, 256, 256, 3)]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(200):  # ~1 s of dense GEMM: bring clocks / power state to steady state before timing anything
    a @ a
torch.cuda.synchronize()
for name, H, cin, cout, R in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    flops = 2.0 * B * H * H * cin * cout * R * R
    for kps, nbuf in (("1", "4"), ("1", "2"), ("2", "2")):
        os.environ["DBX_KPS"] = kps
        os.environ["DBX_EPI_BUFS"] = nbuf
        for st in ("2", "3", "4", "5", "6", "7", "8"):
            os.environ["DBX_STAGES"] = st
            t = timeit(lambda: ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, relu=True), n=30)
            print("%-14s kps=%s nbuf=%s stages<=%s %7.4f ms %7.1f TFLOP/s" % (name, kps, nbuf, st, t, flops / t * 1e-9), flush=True)
