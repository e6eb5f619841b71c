#!/usr/bin/env python
"""bench.py — training patches/s of the DenseBox hot path at 240x240 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variant densebox|lm|lmloc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (N=1): BASELINE.json configs[1] — batch 32 x 240x240, DenseBox (score + bbox heads), bf16 tensor-core math,
fp32 accumulation/master weights, synthetic data, seeded random-init weights.  Weak scaling: 32 patches per GPU.
A step = H2D-free forward + fused loss + backward + (N>1: gradient SUM all-reduce) + SGD on device-resident inputs
(`value`), and the same through the public API with HOST pinned inputs and a D2H read of the loss (`e2e`).
`--impl reference` times the reference's CPU implementation of the same step (the oracle port — the reference is
pure Python/torch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_TRAIN = {"densebox": 125.732, "lm": 134.639, "lmloc": 143.221}  # SURVEY.md §8(d), fwd+dgrad+wgrad per patch
PER_GPU_BATCH = {"densebox": 32, "lm": 32, "lmloc": 32}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def traffic_from_profile(variant, B):
    """DRAM bytes per launch of the dominant kernel family from the committed `ncu --set full` capture
    (profiles/ncu_traffic_r1.json: dram__bytes_read.sum + dram__bytes_write.sum averaged over the 25 fprop/dgrad
    launches of one B=32 DenseBox step); null for any other workload."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    if variant != "densebox" or B != 32 or not os.path.exists(p):
        return None
    d = json.load(open(p))
    return {"bytes": int(d["dram_mb_per_launch"] * 1e6), "launches": d["launches"], "source": d["source"]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up (nvidia-smi needs a few
    hundred ms to produce its first row on an 8-GPU box); only rows stamped inside the timed region (`with` block,
    + one period) are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.t0, self.t1 = gpu_index, [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        sm, mx, pw, reasons = [], 0, [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def synth(variant, B, rank, steps):
    """Synthetic batches of the SURVEY §8(d) recipe, one per step, as host (pinned) tensors."""
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(2 + rank)
    out = []
    rs = np.random.RandomState(100 + rank)
    for s in range(steps):
        x = torch.randn(B, 3, 240, 240, generator=g).pin_memory()
        x0 = rs.randint(40, 101, B); y0 = rs.randint(60, 121, B); w = rs.randint(40, 81, B); h = rs.randint(16, 33, B)
        bbox = torch.tensor(np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float32) / 4.0).pin_memory()
        item = {"x": x, "bbox": bbox,
                "rand": torch.tensor(np.stack([rs.choice(3600, 256, replace=False) for _ in range(B)])).pin_memory()}
        if variant != "densebox":
            c = np.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1) + rs.randint(-2, 3, (B, 8))
            item["vertices"] = torch.tensor(c.astype(np.float32) / 4.0).pin_memory()
            item["lm_rand"] = torch.tensor(rs.randint(0, 3600, (B, 4))).pin_memory()
        out.append(item)
    return out


def cpu_step_time(variant, B, steps, warmup, threads):
    """Reference CPU path (oracle port of DenseBox.py loop body + torch.optim.SGD), fp32, `threads` host threads."""
    import numpy as np
    import torch
    from oracle import densebox_oracle as O
    torch.set_num_threads(threads)
    vgg = O.seeded_vgg19(0)
    P = O.params_from_vgg(vgg, variant, seed_heads=1)
    for v in P.values():
        v.requires_grad_(True)
    opt = torch.optim.SGD([v for k, v in P.items() if not k.startswith("conv3_3")], lr=1e-9, momentum=0.9,
                          weight_decay=5e-8)
    rs = np.random.RandomState(3)
    times = []
    for s in range(warmup + steps):
        x = torch.randn(B, 3, 240, 240, generator=torch.Generator().manual_seed(s))
        lab = O.synth_batch(B, seed=s, with_vertices=variant != "densebox")
        rand = np.stack([rs.choice(3600, 256, replace=False) for _ in range(B)])
        lm_rand = rs.randint(0, 3600, (B, 4))
        g = torch.Generator().manual_seed(50 + s)
        drop = {h: (torch.rand(B, 512, 60, 60, generator=g) < 0.5).float() * 2 for h, _ in O.HEADS[variant]}
        t0 = time.perf_counter()
        opt.zero_grad()
        outs = O.forward(P, x, variant, dropout=drop)
        L, _ = O.loss(outs, variant, lab["bbox"], rand, vertices=lab.get("vertices"), lm_rand_idx=lm_rand)
        L.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return times, float(L.detach())


def run_reference(args, rank):
    if rank != 0:
        return
    variant = args.variant
    threads = os.cpu_count() or 1
    B = 4
    times, loss = cpu_step_time(variant, B, args.steps, args.warmup, threads)
    total = sum(times)
    v = B * len(times) / total
    line = {
        "metric": "training patches/sec at 240x240", "value": v, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "configs[1]: DenseBox (score+bbox) 240x240 training step, fwd+loss+bwd+SGD",
                   "variant": variant, "sample": "batch %d per step on the host CPU" % B},
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of batch %d (oracle port of the reference loop body; the reference is "
                                   "Python/torch and does not travel to the GPU box; label generation vectorised)"
                                   % (len(times), B)},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="densebox", choices=["densebox", "lm", "lmloc"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import densebox_b200
    from oracle import densebox_oracle as O  # cpu_baseline leg + seeded weights recipe only

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    variant = args.variant
    B = args.batch or PER_GPU_BATCH[variant]
    pk = peaks()

    # ---- CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = 4
        times, _ = cpu_step_time(variant, cb, 3, 1, threads)
        cpu_base = {"value": cb / min(times), "unit": "patches/s", "cores": threads, "kind": "port",
                    "sample": "best of 3 steps of batch %d (fwd+loss+bwd+SGD, fp32, oracle port of DenseBox.py "
                              ":2843-2926; vectorised label generation)" % cb}

    vgg = O.seeded_vgg19(0)
    torch.manual_seed(1)
    net = getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}[variant])(vgg)
    net = net.to(dev)
    pg = dist.group.WORLD if world > 1 else None
    tr = densebox_b200.DenseBoxTrainer(net, B, lr=1e-9, momentum=0.9, weight_decay=5e-8, process_group=pg,
                                       use_cuda_graph=not args.no_graph, dropout=True, device=dev)
    nb = 4
    batches = synth(variant, B, rank, nb)
    dev_batches = [{k: v.to(dev) for k, v in b.items()} for b in batches]

    def step(b):
        return tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"],
                       lm_rand_neg_idx=b.get("lm_rand"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local_rank)  # started now: its first rows arrive during the warm-up
    # ---- kernel launches per step (eager step 0 also does the one-time initialisation)
    l0 = tr.eng.launch_count()
    step(dev_batches[0])
    launches_per_step = tr.eng.launch_count() - l0 + (1 if tr.dropout else 0) + (1 if world > 1 else 0)
    for i in range(args.warmup):
        step(dev_batches[i % nb])
    # ---- timed: device-resident inputs
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clk:
        e0.record()
        for i in range(args.steps):
            loss_t = step(dev_batches[i % nb])
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    loss_val = float(loss_t.item())
    # ---- timed: end to end through the public API with HOST (pinned) buffers: every step copies its batch host ->
    # device and reads its loss back.  The usual prefetching-loader pattern: the copy of batch i+1 is started
    # (trainer.prefetch, side stream) while step i computes, so it overlaps instead of serialising.
    def prefetch(b):
        tr.prefetch(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"],
                    lm_rand_neg_idx=b.get("lm_rand"))

    for i in range(2):
        step(batches[i % nb]).item()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    prefetch(batches[0])
    for i in range(args.steps):
        loss_h = step(batches[i % nb])        # consumes the staged copy of this batch
        prefetch(batches[(i + 1) % nb])       # H2D of the next batch, concurrent with this step
        loss_h.item()                         # D2H read of this step's loss
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    h2d = sum(v.numel() * v.element_size() for v in batches[0].values())

    # ---- per-kernel accounting of one eager step (CUDA events around every launch of the engine)
    roof, kern = None, None
    if rank == 0:
        tr.eng.profile(True)
        tr._fwd_loss_bwd()
        tr.eng.sgd_step(tr.lr, tr.momentum, tr.weight_decay)
        recs = tr.eng.profile_records()
        tr.eng.profile(False)
        fam = {}
        for tag, fl, t_ms in recs:
            k = tag.split(":")[0]
            f = fam.setdefault(k, [0.0, 0.0, 0])
            f[0] += fl; f[1] += t_ms; f[2] += 1
        tot_ms = sum(f[1] for f in fam.values())
        kern = {k: {"ms": round(f[1], 4), "share": round(f[1] / tot_ms, 4), "launches": f[2],
                    "tflops": round(f[0] / f[1] * 1e-9, 1) if f[0] > 0 and f[1] > 0 else None}
                for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1])}
        conv = [fam.get("fprop", [0, 0, 0]), fam.get("dgrad", [0, 0, 0])]
        fl = conv[0][0] + conv[1][0]; tm = conv[0][1] + conv[1][1]
        ach = fl / tm * 1e-9
        tp = traffic_from_profile(variant, B)
        roof = {"kernel": "conv_fprop_kernel + conv3x3_halo_kernel (tcgen05 implicit GEMM: %d fprop + %d dgrad launches per step)"
                          % (conv[0][2], conv[1][2]),
                "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4), "traffic": (tp or {}).get("bytes"),
                "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
                "traffic_source": (tp or {}).get("source"), "peak_source": pk["source"],
                "algorithmic_gflop_per_launch_avg": round(fl / max(conv[0][2] + conv[1][2], 1) * 1e-9, 2),
                "avg_launch_ms": round(tm / max(conv[0][2] + conv[1][2], 1), 4),
                "step_tflops": round(GFLOP_TRAIN[variant] * B / ms, 1),
                "step_frac": round(GFLOP_TRAIN[variant] * B / ms / pk["tf_sustained"], 4)}

    if rank == 0:
        value = world * B / (ms * 1e-3)
        line = {
            "metric": "training patches/sec at 240x240", "value": round(value, 1), "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: batch=32 240x240 bf16 training step, score+bbox heads (DenseBox)"
                       if variant == "densebox" else "240x240 bf16 training step, %s" % variant,
                       "variant": variant, "per_gpu_batch": B, "global_batch": B * world,
                       "parallelism": "dp%d" % world, "step": "fwd + fused loss + bwd + grad allreduce(sum) + SGD",
                       "cache": "working set %.1f GB per step >> 126 MB L2 (inputs larger than L2, no flush needed)"
                                % (tr.eng.workspace_bytes / 1e9),
                       "cuda_graph": not args.no_graph, "weights": "seeded vgg19(weights=None) + xavier heads",
                       "optimizer": "SGD lr=1e-9 m=0.9 wd=5e-8 (DenseBox.py:2821-2824)"},
            "e2e": {"value": round(world * B / (ms_e2e * 1e-3), 1), "unit": "patches/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e, 4)},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clk.summary(), "roofline": roof, "kernels": kern, "cpu_baseline": cpu_base,
            "loss": loss_val, "gflop_per_patch": GFLOP_TRAIN[variant],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
