"""Per-launch CUDA-event timing of one eager inference forward (engine profiling hooks); prints ms and TFLOP/s.
usage: profile_infer.py [variant] [batch] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import densebox_b200
from oracle import densebox_oracle as O

variant = sys.argv[1] if len(sys.argv) > 1 else "lmloc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
vgg = O.seeded_vgg19(0)
torch.manual_seed(1)
net = getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}[variant])(vgg)
net = net.cuda().eval()
x = torch.randn(B, 3, S, S, device="cuda")
with torch.no_grad():
    for _ in range(3):
        outs = net(x)
    eng = next(iter(net._engines.values()))
    eng.profile(True)
    outs = net(x)
    recs = eng.profile_records()
    eng.profile(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        outs = net(x)
    e1.record()
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / 5
    score = outs[1] if variant == "lmloc" else outs[0]
    loc = outs[2] if variant == "lmloc" else outs[1]
    lml = outs[4] if variant == "lmloc" else None
    import time
    dets = densebox_b200.decode_nms(score, loc, lml, K=10, nms_thresh=0.4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dets = densebox_b200.decode_nms(score, loc, lml, K=10, nms_thresh=0.4)
    dec_ms = (time.perf_counter() - t0) / 20 * 1e3
tot = sum(r[2] for r in recs)
print("%-28s %9s %9s %7s" % ("launch", "ms", "TFLOP/s", "share"))
for tag, fl, ms in recs:
    print("%-28s %9.4f %9s %6.1f%%" % (tag, ms, ("%.1f" % (fl / ms * 1e-9)) if fl else "-", 100 * ms / tot))
print("total %.3f ms over %d launches; module forward (incl. output shuffles) %.3f ms; decode_nms (kernel + copy + host rows, wall clock) %.3f ms"
      % (tot, len(recs), fwd_ms, dec_ms))
