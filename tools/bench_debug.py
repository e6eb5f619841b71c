"""Decompose one fprop shape: DBX_DEBUG=0 normal, 1 no TMA loads (MMA issue + epilogue only), 2 no MMAs (TMA +
epilogue only), 3 neither (epilogue + barrier round trips only).  Outputs are garbage in modes 1-3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv1_2", 240, 64, 64, 3), ("conv2_2", 120, 128, 128, 3), ("conv2_1 dgrad", 120, 128, 64, 3),
          ("conv2_1 fprop", 120, 64, 128, 3), ("conv3_2", 60, 256, 256, 3), ("conv4_2", 30, 512, 512, 3),
          ("heads1", 60, 768, 1024, 1), ("heads2 dgrad", 60, 64, 1024, 1)]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(200):
    a @ a
torch.cuda.synchronize()
for name, H, cin, cout, R in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    flops = 2.0 * B * H * H * cin * cout * R * R
    for kps in ["-"]:
        for dbg in ("0", "1", "2", "3"):
            os.environ["DBX_DEBUG"] = dbg
            t = timeit(lambda: ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, relu=True), n=30)
            print("%-14s kps=%s debug=%s %7.4f ms %7.1f TFLOP/s-equivalent" % (name, kps, dbg, t, flops / t * 1e-9), flush=True)
