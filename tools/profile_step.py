"""Per-launch CUDA-event timing of one eager training step (engine profiling hooks); prints ms and TFLOP/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import densebox_b200
from bench import synth
from oracle import densebox_oracle as O

variant = sys.argv[1] if len(sys.argv) > 1 else "densebox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
vgg = O.seeded_vgg19(0)
torch.manual_seed(1)
net = getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}[variant])(vgg).cuda()
tr = densebox_b200.DenseBoxTrainer(net, B, use_cuda_graph=False)
b = {k: v.cuda() for k, v in synth(variant, B, 0, 1)[0].items()}
for _ in range(3):
    tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
tr.eng.profile(True)
tr._fwd_loss_bwd()
tr.eng.sgd_step(tr.lr, tr.momentum, tr.weight_decay)
recs = tr.eng.profile_records()
tot = sum(r[2] for r in recs)
print("%-28s %9s %9s %7s" % ("launch", "ms", "TFLOP/s", "share"))
for tag, fl, ms in recs:
    print("%-28s %9.4f %9s %6.1f%%" % (tag, ms, ("%.1f" % (fl / ms * 1e-9)) if fl else "-", 100 * ms / tot))
print("total %.3f ms over %d launches" % (tot, len(recs)))
