"""A/B of CTA pairs (tcgen05.mma.cta_group::2) on the narrow-N fprop/dgrad shapes (B=32, 240x240 patches)."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [  # name, H, cin, cout
    ("conv2_1 fprop", 120, 64, 128), ("conv2_1 dgrad", 120, 128, 64), ("conv2_2", 120, 128, 128),
    ("conv3_1 fprop", 60, 128, 256), ("conv3_1 dgrad", 60, 256, 128), ("conv4_1 dgrad", 30, 512, 256),
    ("conv3_2", 60, 256, 256), ("conv4_2", 30, 512, 512),
]
g = torch.Generator(device="cuda").manual_seed(0)
for name, H, cin, cout in SHAPES:
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(cout, 9 * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    aux = torch.randn(B, H, H, cout, generator=g, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    flops = 2.0 * B * H * H * cin * cout * 9
    ref = None
    for min_n in ("512", "64"):
        os.environ["DBX_CTA2_MIN_N"] = min_n
        ops.conv_fprop(x, wk, 3, 3, 1, out, bias=bias, aux=aux, aux_mode=1)
        o = out.float().clone()
        if ref is None:
            ref = o
        err = (o - ref).abs().max().item()
        t = timeit(lambda: ops.conv_fprop(x, wk, 3, 3, 1, out, bias=bias, aux=aux, aux_mode=1), n=20)
        print("%-14s cta2_min_n=%3s %7.4f ms %7.1f TFLOP/s  maxdiff vs single %.3g" % (name, min_n, t, flops / t * 1e-9, err),
              flush=True)
