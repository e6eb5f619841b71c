"""Wave quantisation on the 30x30 layers: block_n 256 (256 pair-units on 74 CTA pairs = 3.46 waves) vs 128 (6.92)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops
from tools.bench_layers import timeit

B = 32
SHAPES = [("conv4_1 fprop", 30, 256, 512), ("conv4_2", 30, 512, 512), ("conv4_1 dgrad", 30, 512, 256), ("conv3_2", 60, 256, 256)]
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
    for bn in (0, 256, 128, 192):
        for am in (0, 1):
            fn = lambda: ops.conv_fprop(x, wk, 3, 3, 1, out, bias=bias, relu=am == 0, aux=aux if am else None, aux_mode=am, block_n=bn)
            try:
                fn(); torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("%-14s bn=%3d aux=%d FAILED %s" % (name, bn, am, str(e)[:60])); continue
            t = timeit(fn, n=20)
            print("%-14s bn=%3d aux=%d %7.4f ms %7.1f TFLOP/s" % (name, bn, am, t, flops / t * 1e-9), flush=True)
