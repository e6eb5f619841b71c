"""Where does the end-to-end (host-pinned inputs) step lose time against the device-resident one?  Runs the two timed
loops of bench.py in both orders, with and without an idle gap before each region, and at two lengths; prints ms/step
and the nvidia-smi clocks of each region.  Also: the drop-in module path, wall time vs the sum of its kernel times."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    c = bench.Ctx()
    c.dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    c.rank, c.world, c.dist, c.pk = 0, 1, None, bench.peaks()
    c.barrier = torch.cuda.synchronize
    c.sampler = bench.ClockSampler(0)
    import densebox_b200
    net = bench.make_net("densebox", c.dev)
    tr = densebox_b200.DenseBoxTrainer(net, 32, lr=1e-9, device=c.dev)
    batches = bench.synth("densebox", 32, 0, 4)
    dev_batches = [{k: v.to(c.dev) for k, v in b.items()} for b in batches]
    for bs in (dev_batches, batches):
        for i in range(4):
            b = bs[i % 4]
            tr.step(b["x"], b["bbox"], rand_neg_idx=b["rand"])
    torch.cuda.synchronize()
    for steps in (20, 100):
        for gap in (0.0, 1.0):
            for order in (("dev", "host"), ("host", "dev")):
                res = []
                for which in order:
                    time.sleep(gap)
                    ms, _, clk = bench.timed_steps(c, tr, batches if which == "host" else dev_batches, steps,
                                                   host=which == "host", sampler=c.sampler)
                    res.append("%s %.3f ms (%s MHz, %s W)" % (which, ms, clk["sm_mhz"], clk["power_w_max"]))
                print("steps %3d gap %.1fs: %s" % (steps, gap, " | ".join(res)), flush=True)
    # host-side cost of one e2e iteration without waiting for the GPU
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for i in range(n):
        b, nxt = batches[i % 4], batches[(i + 1) % 4]
        tr.step(b["x"], b["bbox"], rand_neg_idx=b["rand"], async_loss=True)
        tr.prefetch(nxt["x"], nxt["bbox"], rand_neg_idx=nxt["rand"])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host enqueue time per e2e iteration: %.3f ms (then %.3f ms to drain)" % ((t1 - t0) / n * 1e3, (t2 - t1) * 1e3 / n))
    # H2D bandwidth of the pinned batch
    x = batches[0]["x"]
    d = torch.empty_like(x, device=c.dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(x, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D pinned: %.1f GB/s" % (10 * x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
    c.sampler.stop()
    del tr
    torch.cuda.empty_cache()
    # ---- drop-in module path
    from densebox_b200 import densebox_loss
    net = bench.make_net("densebox", c.dev).train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-9, momentum=0.9, weight_decay=5e-8)

    def step(b, sync=False):
        ts = [time.perf_counter()]
        opt.zero_grad()
        score, loc = net(b["x"])
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
        L = densebox_loss(score, loc, b["bbox"], rand_neg_idx=b["rand"])
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
        L.backward()
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
        opt.step()
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
        return ts

    for i in range(3):
        step(dev_batches[i % 4])
    torch.cuda.synchronize()
    ts = step(dev_batches[0], sync=True)
    print("drop-in, synchronised phases (ms): forward %.3f loss %.3f backward %.3f optimizer %.3f" % tuple(
        (b - a) * 1e3 for a, b in zip(ts[:-1], ts[1:])))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        ts = step(dev_batches[i % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("drop-in: host enqueue %.3f ms/step, wall %.3f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
    ts = step(dev_batches[0])
    print("drop-in, host-only phases (ms): forward %.3f loss %.3f backward %.3f optimizer %.3f" % tuple(
        (b - a) * 1e3 for a, b in zip(ts[:-1], ts[1:])))


if __name__ == "__main__":
    main()
