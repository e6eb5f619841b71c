"""Launch one fprop/dgrad shape a few times — the process `ncu --set full` is pointed at.
usage: python tools/one_layer.py H cin cout [aux_mode] [reps]   (B = 32, 3x3 pad 1; cin == 0: the heads2 dgrad shape)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops

H, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
aux_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
R = int(os.environ.get("DBX_R", "3"))
B = 32
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, H, cin, generator=g, device="cuda").to(torch.bfloat16)
wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
aux = torch.randn(B, H, H, cout, generator=g, device="cuda").to(torch.bfloat16) if aux_mode else None
bias = torch.zeros(cout, device="cuda")
for _ in range(reps):
    ops.conv_fprop(x, wk, R, R, R // 2, out, bias=bias, relu=aux_mode == 0, aux=aux, aux_mode=aux_mode)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
