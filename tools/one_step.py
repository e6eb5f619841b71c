"""One eager training step (no CUDA graph) of the bench workload — the process `ncu` is pointed at.
usage: python tools/one_step.py [variant] [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import densebox_b200
from bench import synth
from oracle import densebox_oracle as O

variant = sys.argv[1] if len(sys.argv) > 1 else "densebox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
vgg = O.seeded_vgg19(0)
torch.manual_seed(1)
net = getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}[variant])(vgg).cuda()
tr = densebox_b200.DenseBoxTrainer(net, B, use_cuda_graph=False)
b = {k: v.cuda() for k, v in synth(variant, B, 0, 1)[0].items()}
for _ in range(steps):
    L = tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
torch.cuda.synchronize()
print("loss", float(L))
