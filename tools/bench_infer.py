"""Config 5 of BASELINE.json: inference-only forward at 1024x1024, batch 16, + per-image top-10 decode + NMS 0.4.
Prints images/s (CUDA events, inputs resident in HBM) for the DenseBoxLMLOC heads (871.2 GFLOP / image)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import densebox_b200
from oracle import densebox_oracle as O

variant = sys.argv[1] if len(sys.argv) > 1 else "lmloc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
GF = {"densebox": 41.977, "lm": 44.946, "lmloc": 47.807}[variant] * (S / 240.0) ** 2
vgg = O.seeded_vgg19(0)
torch.manual_seed(1)
net = getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}[variant])(vgg)
net = net.cuda().eval()
x = torch.randn(B, 3, S, S, device="cuda")


def run():
    with torch.no_grad():
        outs = net(x)
    if variant == "lmloc":
        score, rf, loc, lm, lmloc = outs
        return densebox_b200.decode_nms(rf, loc, lmloc, K=10, nms_thresh=0.4)  # test_lmloc uses the refined score
    score, loc = outs[0], outs[1]
    return densebox_b200.decode_nms(score, loc, None, K=10, nms_thresh=0.4)


for _ in range(3):
    dets = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    dets = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print({"config": "inference %dx%d batch %d %s + top-10 decode + NMS" % (S, S, B, variant), "ms_per_batch": round(ms, 3),
       "images_per_s": round(B / ms * 1e3, 1), "tflops": round(GF * B / ms, 1), "dets_image0": int(len(dets[0]))})
