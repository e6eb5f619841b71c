"""A/B of K blocks per pipeline stage (DBX_KPS) on the fprop/dgrad shapes of the bench (B=32, 240x240 patches)."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [  # name, H, cin, cout, R
    ("conv2_1 fprop", 120, 64, 128, 3), ("conv2_1 dgrad", 120, 128, 64, 3), ("conv2_2", 120, 128, 128, 3),
    ("conv3_1 fprop", 60, 128, 256, 3), ("conv3_1 dgrad", 60, 256, 128, 3), ("conv3_2", 60, 256, 256, 3),
    ("conv4_1 dgrad", 30, 512, 256, 3), ("conv4_2", 30, 512, 512, 3), ("heads1 fprop", 60, 768, 1024, 1),
    ("heads1 dgrad", 60, 1024, 768, 1), ("heads2 dgrad", 60, 64, 1024, 1),
]
g = torch.Generator(device="cuda").manual_seed(0)
for name, H, cin, cout, R in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    aux = torch.randn(B, H, H, cout, generator=g, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    flops = 2.0 * B * H * H * cin * cout * R * R
    ref = None
    for kps in sys.argv[1:] or ["1", "2", "4"]:
        os.environ["DBX_KPS"] = kps
        ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, aux=aux, aux_mode=1)
        o = out.float().clone()
        if ref is None:
            ref = o
        err = (o - ref).abs().max().item()
        t = timeit(lambda: ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, aux=aux, aux_mode=1), n=20)
        print("%-14s kps=%s %7.4f ms %7.1f TFLOP/s  maxdiff vs first %.3g" % (name, kps, t, flops / t * 1e-9, err), flush=True)
