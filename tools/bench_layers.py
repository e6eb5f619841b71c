"""Per-layer timing of the tcgen05 kernels at the benchmark shapes (B=32, 240x240 patches).
Prints ms and TFLOP/s for fprop and wgrad of every conv on the hot path. Inputs are larger than L2 for the big
layers; CUDA events on the current stream; 3 warm-up + 10 timed launches each."""
import os
os.environ["DBX_ENABLE_AB"] = "1"  # the library honours its A/B switches only when this is set
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densebox_b200 import ops

B = int(os.environ.get("DBX_B", "32"))
LAYERS = [  # name, H, cin, cout, R
    ("conv1_1(im2col)", 240, 64, 64, 1), ("conv1_2", 240, 64, 64, 3), ("conv2_1", 120, 64, 128, 3),
    ("conv2_2", 120, 128, 128, 3), ("conv3_1", 60, 128, 256, 3), ("conv3_2", 60, 256, 256, 3),
    ("conv4_1", 30, 256, 512, 3), ("conv4_2", 30, 512, 512, 3), ("conv5_1x2", 60, 768, 1024, 1),
    ("conv5_2", 60, 1024, 16, 1),
]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    tot_f = tot_w = 0.0
    for name, H, cin, cout, R in LAYERS:
        pad = R // 2
        x = (torch.randn(B, H, H, cin, generator=g, device="cuda")).to(torch.bfloat16)
        dy = (torch.randn(B, H, H, cout, generator=g, device="cuda")).to(torch.bfloat16)
        wk = (torch.randn(cout, R * R * cin, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
        bias = torch.zeros(cout, device="cuda")
        dw = torch.zeros(cout, R * R * cin, device="cuda")
        flops = 2.0 * B * H * H * cin * cout * R * R
        for bn in ([0] if cout <= 128 else [256]):
            for c2 in ("0", "1"):
                os.environ["DBX_CTA2"] = c2
                tf = timeit(lambda: ops.conv_fprop(x, wk, R, R, pad, out, bias=bias, relu=True, block_n=bn))
                print("%-16s fprop bn=%3d cta2=%s %8.3f ms %8.1f TFLOP/s" % (name, bn, c2, tf, flops / tf * 1e-9),
                      flush=True)
        tw = timeit(lambda: ops.conv_wgrad(x, dy, R, R, pad, dw))
        print("%-16s wgrad        %8.3f ms %8.1f TFLOP/s" % (name, tw, flops / tw * 1e-9), flush=True)
        tot_f += tf
        tot_w += tw
    print("sum fprop %.3f ms, sum wgrad %.3f ms (x multiplicities not applied)" % (tot_f, tot_w))


if __name__ == "__main__":
    main()
