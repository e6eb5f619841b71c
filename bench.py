#!/usr/bin/env python
"""bench.py — training patches/s of the DenseBox hot path at 240x240 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--records all|main]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (every N): BASELINE.json configs[1] — 32 patches of 240x240 per GPU, DenseBox (score + bbox heads),
bf16 tensor-core math, fp32 accumulation / master weights, synthetic data, seeded random-init weights; weak scaling.
A step = forward + fused loss + backward + (N>1: gradient SUM all-reduce) + SGD.
  value : device-resident inputs (the batch is copied device-to-device into the trainer's input slot each step)
  e2e   : the same through the public API with HOST pinned inputs — every step copies its batch host -> device
          (prefetched on a side stream) and reads its loss back (asynchronously, consumed one step later)
Further records on the same JSON line (driver-run evidence for the other BASELINE configs):
  N = 1 : parity (loss of the first bench batch vs the CPU oracle fed the same Philox dropout mask), config3 (B=64
          DenseBoxLM), config5 (1024x1024 B=16 DenseBoxLMLOC inference + top-10 decode + NMS, images/s), sustained
          (>= 300 steps of the headline workload with its own clock record), dropin (the nn.Module path: net(x) ->
          densebox_loss -> backward -> torch.optim.SGD.step), e2e_u8 (the headline step fed uint8 image bytes)
  N > 1 : config4 (DenseBoxLM, 32 per GPU, data parallel: value, the same variant on one GPU of this box, exposed
          communication per step = step time with minus without the all-reduces)
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
GFLOP_INFER_1024 = {"lmloc": 871.2}                                     # BASELINE.md §2, forward per 1024x1024 image
CLS = {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}
METRIC = "training patches/sec at 240x240"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def headline_config(world, graph=True):
    """The `config` object of the headline line; identical in both arms (the reference arm samples it)."""
    return {"workload": "configs[1]: batch=32 240x240 bf16 training step, score+bbox heads (DenseBox)",
            "variant": "densebox", "per_gpu_batch": 32, "global_batch": 32 * world, "parallelism": "dp%d" % world,
            "step": "fwd + fused loss + bwd + grad allreduce(sum) + SGD",
            "cache": "inputs larger than L2: one step streams ~3.6 GB of activations per GPU >> 126 MB L2 (no flush needed)",
            "timing": "CUDA events around K steps between two barriers; W + 40 untimed steps first, then every timed region starts after 1 s of idle + "
                      "3 untimed steps (same power state for value and e2e); steady state under the power cap: record `sustained`",
            "cuda_graph": graph, "weights": "seeded vgg19(weights=None) + xavier heads",
            "optimizer": "SGD lr=1e-9 m=0.9 wd=5e-8 (DenseBox.py:2821-2824)"}


def traffic_from_profile(variant, B):
    """DRAM bytes per launch of the dominant kernel family from the committed `ncu --set full` capture: null for any
    workload other than the one profiled."""
    for name in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if variant == "densebox" and B == 32 and os.path.exists(p):
            d = json.load(open(p))
            return {"bytes": int(d["dram_mb_per_launch"] * 1e6), "launches": d["launches"], "source": d["source"]}
    return None


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

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(timeout=2)
            self.proc = None

    def summary(self, t0=None, t1=None):
        t0 = self.t0 if t0 is None else t0
        t1 = self.t1 if t1 is None else t1
        inside = [r for t, r in self.rows if t0 is not None and t0 <= t <= t1 + 0.06]
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
    variant = "densebox"
    threads = os.cpu_count() or 1
    B = 4
    times, loss = cpu_step_time(variant, B, args.steps, args.warmup, threads)
    total = sum(times)
    v = B * len(times) / total
    line = {
        "metric": METRIC, "value": v, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "impl": "reference", "config": headline_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": threads, "kind": "port",
                         "sample": "%d steps, each a 4-patch sample of the 32-patch batch (fp32, all host threads; oracle "
                                   "PORT of the reference loop body DenseBox.py:2843-2926 + torch.optim.SGD — the "
                                   "reference is Python/torch and does not travel to the GPU box; its label generation "
                                   "is vectorised here, so this is faster than the reference itself)" % len(times)},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Ctx:
    pass


def make_net(variant, dev):
    import torch
    import densebox_b200
    from oracle import densebox_oracle as O  # seeded weights recipe (shared with the CPU arm), outside any timed region
    vgg = O.seeded_vgg19(0)
    torch.manual_seed(1)
    return getattr(densebox_b200, CLS[variant])(vgg).to(dev)


def timed_steps(c, tr, batches, steps, host, sampler=None, gap=1.0):
    """`steps` training steps, CUDA events on the launching stream between two barriers; returns ms per step (this
    rank) and the last loss.  host=True: pinned host batches through prefetch() + asynchronous loss read-back."""
    import torch
    nb = len(batches)

    def kw(b):
        return dict(vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))

    c.barrier()
    if gap:
        # Equal power state for every timed region: a region that starts right behind another one inherits its
        # clocks (tools/e2e_probe.py: the SECOND of two back-to-back 20-step regions runs at ~1 650 MHz instead of
        # 1 965 and reads 6-8 % slower, whichever of value / e2e it is; after 1 s of idle they agree within 0.4 %).
        # The power-capped steady state is reported separately (`sustained`).
        time.sleep(gap)
        c.barrier()
        # ... and three untimed steps behind the idle second, so the region does not start on clocks that are still
        # ramping up (two full runs read value / e2e 3-4 % apart in either direction without them).
        for i in range(3):
            b = batches[i % nb]
            tr.step(b["x"], b["bbox"], **kw(b))
        torch.cuda.synchronize()
        c.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    last = None
    if host:
        tr.prefetch(batches[0]["x"], batches[0]["bbox"], **kw(batches[0]))
        prev = None
        for i in range(steps):
            b, nxt = batches[i % nb], batches[(i + 1) % nb]
            h = tr.step(b["x"], b["bbox"], async_loss=True, **kw(b))   # consumes the staged copy of this batch
            tr.prefetch(nxt["x"], nxt["bbox"], **kw(nxt))               # H2D of the next batch, concurrent with this step
            if prev is not None:
                last = prev.item()                                      # D2H read of the previous step's loss
            prev = h
        last = prev.item()
    else:
        for i in range(steps):
            b = batches[i % nb]
            t = tr.step(b["x"], b["bbox"], **kw(b))
        last = t
    e1.record()
    c.barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / steps
    if not host:
        last = float(last.item())
    clocks = sampler.summary(t0, t1) if sampler is not None else None
    return ms, last, clocks


def max_over_ranks(c, vals):
    import torch
    t = torch.tensor(vals, device=c.dev, dtype=torch.float64)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return [float(v) for v in t]


def loss_parity(c, tr, net_params, variant, batch):
    """The loss of the trainer's FIRST step (eager, Philox dropout from (seed, offset 0)) against the CPU oracle fed the
    same bf16-rounded weights and the same dropout mask (materialised with dbx_dropout_mask).  Returns the record;
    the step counts as the first warm-up step."""
    import ctypes
    import numpy as np
    import torch
    from densebox_b200._lib import check, lib, ptr, stream_ptr
    from oracle import densebox_oracle as O
    B = tr.B
    loss0 = float(tr.step(batch["x"], batch["bbox"], vertices=batch.get("vertices"), rand_neg_idx=batch["rand"],
                          lm_rand_neg_idx=batch.get("lm_rand")).item())
    half, pos = tr.eng.loss_info()
    heads = {"densebox": ["det", "loc"], "lm": ["det", "loc", "landmark"], "lmloc": ["det", "loc", "landmark", "lmloc"]}[variant]
    nh = len(heads)
    mask = torch.empty(B, 60, 60, 512 * nh, dtype=torch.bfloat16, device=c.dev)
    check(lib().dbx_dropout_mask(ptr(mask), ctypes.c_ulonglong(mask.numel()), ctypes.c_ulonglong(tr.seed),
                                 ctypes.c_ulonglong(0), stream_ptr()), "dropout_mask")
    drop = {h: mask[..., 512 * i:512 * (i + 1)].permute(0, 3, 1, 2).float().cpu() for i, h in enumerate(heads)}
    del mask
    P = {k: (v.bfloat16().float() if k.endswith(".weight") else v) for k, v in net_params.items()}
    t0 = time.time()
    with torch.no_grad():
        outs = O.forward(P, batch["x"].bfloat16().float(), variant, dropout=drop)
        L_ref, info = O.loss(outs, variant, batch["bbox"].numpy(), batch["rand"].numpy(),
                             vertices=batch["vertices"].numpy() if "vertices" in batch else None,
                             lm_rand_idx=batch["lm_rand"].numpy() if "lm_rand" in batch else None)
    L_ref = float(L_ref)
    return {"loss": loss0, "oracle_loss": L_ref, "loss_rel_err": abs(loss0 - L_ref) / abs(L_ref), "tolerance": 1e-3,
            "half": half, "oracle_half": info["half"], "pos": pos, "oracle_pos": info["pos"],
            "batch": B, "variant": variant, "dropout": "Philox mask of the step, materialised and fed to the oracle",
            "oracle_seconds": round(time.time() - t0, 1),
            "what": "first step of the bench (train mode) vs oracle.forward+loss on the same bf16-rounded weights/inputs"}


def kernel_profile(tr, variant, B, ms_step, pk):
    """Per-kernel accounting of one eager step (CUDA events around every launch of the engine)."""
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
    nl = max(conv[0][2] + conv[1][2], 1)
    roof = {"kernel": "conv_fprop_kernel + conv3x3_halo_kernel (tcgen05 implicit GEMM: %d fprop + %d dgrad launches per step)"
                      % (conv[0][2], conv[1][2]),
            "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": round(ach / pk["tf_burst"], 4),
            "peak_kind": "burst cuBLAS bf16 (launches timed one by one with CUDA events in a sub-second region)",
            "traffic": (tp or {}).get("bytes"),
            "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
            "traffic_source": (tp or {}).get("source"), "peak_source": pk["source"],
            "algorithmic_gflop_per_launch_avg": round(fl / nl * 1e-9, 2), "avg_launch_ms": round(tm / nl, 4),
            "step_tflops": round(GFLOP_TRAIN[variant] * B / ms_step, 1),
            "step_frac": round(GFLOP_TRAIN[variant] * B / ms_step / pk["tf_burst"], 4)}
    return roof, kern


def run_e2e_u8(c, B=32, steps=20):
    """The headline workload fed as decoded image bytes (uint8 NHWC): ToTensor + Normalize fused into the first kernel,
    a quarter of the host->device bytes."""
    import numpy as np
    import torch
    import densebox_b200
    net = make_net("densebox", c.dev)
    tr = densebox_b200.DenseBoxTrainer(net, B, lr=1e-9, momentum=0.9, weight_decay=5e-8, dropout=True, device=c.dev,
                                       input_u8=True)
    g = torch.Generator().manual_seed(9)
    base = synth("densebox", B, 0, 4)
    batches = [dict(b, x=torch.randint(0, 256, (B, 240, 240, 3), generator=g, dtype=torch.uint8).pin_memory()) for b in base]
    for i in range(5):
        b = batches[i % 4]
        tr.step(b["x"], b["bbox"], rand_neg_idx=b["rand"])
    ms, loss, _ = timed_steps(c, tr, batches, steps, host=True)
    h2d = sum(v.numel() * v.element_size() for v in batches[0].values())
    return {"value": round(B / (ms * 1e-3), 1), "unit": "patches/s", "ms_per_step": round(ms, 4), "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 4, "loss": loss,
            "input": "uint8 [B,240,240,3] pinned host batches; ToTensor + Normalize(ImageNet) inside im2col3x3_c3_u8_kernel"}


def run_training(c, variant, B, steps, warmup, pg, graph=True, parity=False, profile=False, e2e=True, exposed=False):
    """One trainer, one workload: returns the record (rank 0 fills everything, other ranks a subset)."""
    import torch
    import densebox_b200
    from oracle import densebox_oracle as O
    world = c.world if pg is not None else 1
    net = make_net(variant, c.dev)
    net_params = O.params_from_state_dict(net.state_dict(), variant) if parity else None
    tr = densebox_b200.DenseBoxTrainer(net, B, lr=1e-9, momentum=0.9, weight_decay=5e-8, process_group=pg,
                                       use_cuda_graph=graph, dropout=True, device=c.dev)
    nb = 4
    batches = synth(variant, B, c.rank, nb)
    dev_batches = [{k: v.to(c.dev) for k, v in b.items()} for b in batches]
    rec = {"variant": variant, "per_gpu_batch": B, "n_gpus": world, "steps": steps, "warmup": warmup}
    l0 = tr.eng.launch_count()
    if parity and c.rank == 0:
        rec["parity"] = loss_parity(c, tr, net_params, variant, batches[0])
    else:
        b = dev_batches[0]
        tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
    # launches of one step: the engine's own count + the dropout counter update + the loss-ring copy (+ count kernel)
    rec["gpu_launches_per_step"] = int(tr.eng.launch_count() - l0 + (1 if tr.dropout else 0) + 1 + (1 if world > 1 else 0))
    # the W warm-up steps, plus 40 more untimed ones: the CPU oracle check above leaves the GPU idle for ~10 s, and
    # the first timed region read 2-4 % below the second in three full runs out of three without them
    for i in range(max(warmup, 3) + 40):
        b = dev_batches[i % nb]
        tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
    ms, loss, clocks = timed_steps(c, tr, dev_batches, steps, host=False, sampler=c.sampler)
    vals = [ms]
    if e2e:
        for i in range(3):  # warm the host path: staging copies, the graphs of both slots, pinned loss ring
            b = batches[i % nb]
            tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
        ms_e2e, _, _ = timed_steps(c, tr, batches, steps, host=True)
        vals.append(ms_e2e)
    if exposed and world > 1:
        tr._skip_allreduce = True   # measurement only: the same step without its two exchanges
        for i in range(2):
            b = dev_batches[i % nb]
            tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
        ms_nc, _, _ = timed_steps(c, tr, dev_batches, steps, host=False)
        tr._skip_allreduce = False
        vals.append(ms_nc)
    vals = max_over_ranks(c, vals)
    ms = vals[0]
    rec.update({"value": round(world * B / (ms * 1e-3), 1), "unit": "patches/s", "ms_per_step": round(ms, 4),
                "loss": loss, "clocks": clocks, "workspace_gb": round(tr.eng.workspace_bytes / 1e9, 2),
                "step_tflops_per_gpu": round(GFLOP_TRAIN[variant] * B / ms, 1)})
    if e2e:
        h2d = sum(v.numel() * v.element_size() for v in batches[0].values())
        rec["e2e"] = {"value": round(world * B / (vals[1] * 1e-3), 1), "unit": "patches/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": 4, "ms_per_step": round(vals[1], 4),
                      "how": "DenseBoxTrainer.prefetch/step on pinned host tensors; loss read back asynchronously one step later"}
    if exposed and world > 1:
        rec["exposed_comm_us_per_step"] = round((ms - vals[-1]) * 1e3, 1)
        rec["ms_per_step_without_allreduce"] = round(vals[-1], 4)
    if profile and c.rank == 0:
        rec["roofline"], rec["kernels"] = kernel_profile(tr, variant, B, ms, c.pk)
    c.keep = (tr, net, batches, dev_batches)
    return rec


def run_sustained(c, steps):
    """>= 300 steps of the headline workload on the trainer of the main run: the power-capped steady state."""
    tr, net, batches, dev_batches = c.keep
    ms, loss, clocks = timed_steps(c, tr, dev_batches, steps, host=False, sampler=c.sampler, gap=0.0)
    B = tr.B
    tf = GFLOP_TRAIN["densebox"] * B / ms
    return {"steps": steps, "value": round(B / (ms * 1e-3), 1), "unit": "patches/s", "ms_per_step": round(ms, 4),
            "clocks": clocks, "step_tflops": round(tf, 1), "peak": c.pk["tf_sustained"],
            "peak_kind": "sustained cuBLAS bf16 (seconds-long loop under the power cap)",
            "step_frac": round(tf / c.pk["tf_sustained"], 4), "loss": loss}


def run_inference(c, variant="lmloc", N=16, HW=1024, steps=5, warmup=2):
    """configs[4]: forward of a 1024x1024 batch + per-image top-10 decode + NMS 0.4 (test_lmloc DenseBox.py:3565-3647)."""
    import torch
    from densebox_b200 import decode_nms
    net = make_net(variant, c.dev).eval()
    x_host = torch.randn(N, 3, HW, HW, generator=torch.Generator().manual_seed(7)).pin_memory()
    x_dev = x_host.to(c.dev)

    def once(x):
        with torch.no_grad():
            score, rf, loc, lm, lmloc = net(x)
        return decode_nms(rf, loc, lmloc, K=10, nms_thresh=0.4)  # returns host arrays: the detections leave the GPU

    for _ in range(warmup):
        dets = once(x_dev)
    out = {}
    for name, src in (("value", x_dev), ("e2e", x_host)):
        torch.cuda.synchronize()
        time.sleep(1.0)  # same power state for both regions (see timed_steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            dets = once(src.to(c.dev, non_blocking=True) if src is x_host else src)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / steps
    eng = next(iter(net._engines.values()))
    # Inference nets fold conv5_2 o conv5_1 into one 768 -> 17 matrix (no non-linearity between them in the reference,
    # DenseBox.py:158-178): the 768 -> 2048 and 2048 -> 17 products at 256 x 256 pixels are not executed, a 768 -> 17
    # product is.  `frac` is quoted on the FLOPs that run, the reference's count is kept beside it.
    px = (HW // 4) * (HW // 4)
    gf_exec = GFLOP_INFER_1024[variant] - 2.0 * px * (768 * 2048 + 2048 * 17 - 768 * 17) / 1e9
    if os.environ.get("DBX_ENABLE_AB") == "1" and os.environ.get("DBX_HEADS_FOLD") == "0":
        gf_exec = GFLOP_INFER_1024[variant]
    tf = gf_exec * N / out["value"]
    rec = {"workload": "configs[4]: inference forward 1024x1024, batch 16, heads + top-10 decode + NMS 0.4", "variant": variant,
           "batch": N, "value": round(N / (out["value"] * 1e-3), 1), "unit": "images/s", "ms_per_batch": round(out["value"], 3),
           "e2e": {"value": round(N / (out["e2e"] * 1e-3), 1), "unit": "images/s",
                   "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": N * 10 * 14 * 4,
                   "how": "pinned host batch copied in the timed loop (not overlapped), detections copied back"},
           "gflop_per_image": GFLOP_INFER_1024[variant], "gflop_per_image_executed": round(gf_exec, 1),
           "tflops": round(tf, 1), "tflops_at_reference_flops": round(GFLOP_INFER_1024[variant] * N / out["value"], 1),
           "peak": c.pk["tf_burst"],
           "frac": round(tf / c.pk["tf_burst"], 4), "kept_per_image": [int(len(d)) for d in dets][:4],
           "workspace_gb": round(eng.workspace_bytes / 1e9, 2), "steps": steps, "warmup": warmup}
    net._engines.clear()
    return rec


def run_dropin(c, B=32, steps=10, warmup=3):
    """The nn.Module path north_star names: net(x) -> densebox_loss -> .backward() -> torch.optim.SGD.step()."""
    import torch
    from densebox_b200 import densebox_loss
    net = make_net("densebox", c.dev).train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-9, momentum=0.9, weight_decay=5e-8)
    batches = [{k: v.to(c.dev) for k, v in b.items()} for b in synth("densebox", B, 0, 2)]

    def step(b):
        opt.zero_grad()
        score, loc = net(b["x"])
        L = densebox_loss(score, loc, b["bbox"], rand_neg_idx=b["rand"])
        L.backward()
        opt.step()
        return L

    for i in range(warmup):
        step(batches[i % 2])
    torch.cuda.synchronize()
    time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        L = step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rec = {"api": "DenseBox(vgg19)(x) -> densebox_loss(...) -> loss.backward() -> torch.optim.SGD.step() (eager, no CUDA graph)",
           "batch": B, "value": round(B / (ms * 1e-3), 1), "unit": "patches/s", "ms_per_step": round(ms, 4),
           "loss": float(L.item()), "steps": steps, "warmup": warmup}
    net._engines.clear()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="densebox", choices=["densebox", "lm", "lmloc"],
                    help="headline variant (default: configs[1] DenseBox; other values are for ad-hoc measurements)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch of the headline run (default 32)")
    ap.add_argument("--records", default="all", choices=["all", "main"],
                    help="main: only the headline measurement (profiling runs)")
    ap.add_argument("--sustained-steps", type=int, default=300)
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

    torch.cuda.set_device(local_rank)
    c = Ctx()
    c.dev = torch.device("cuda", local_rank)
    c.rank, c.world, c.dist, c.pk = rank, world, dist, peaks()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=c.dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c.barrier = barrier
    variant = args.variant
    B = args.batch or 32
    full = args.records == "all"

    # ---- CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = 4
        times, _ = cpu_step_time(variant, cb, 3, 1, threads)
        cpu_base = {"value": cb / min(times), "unit": "patches/s", "cores": threads, "kind": "port",
                    "sample": "best of 3 steps of batch 4 (fwd+loss+bwd+SGD, fp32, oracle PORT of DenseBox.py "
                              ":2843-2926 with vectorised label generation — faster than the reference itself)"}

    c.sampler = ClockSampler(local_rank) if rank == 0 else None  # started now: its first rows arrive during the warm-up
    pg = dist.group.WORLD if world > 1 else None
    main_rec = run_training(c, variant, B, args.steps, args.warmup, pg, graph=not args.no_graph,
                            parity=full and world == 1, profile=True, e2e=True, exposed=world > 1)
    records = {}
    if full and world == 1:
        records["sustained"] = run_sustained(c, args.sustained_steps)
    c.keep = None
    torch.cuda.empty_cache()
    if full and world == 1:
        records["config3"] = run_training(c, "lm", 64, 10, 3, None, graph=not args.no_graph, parity=True, profile=False)
        records["config3"]["workload"] = "configs[2]: batch=64 multi-task (score+bbox+4-landmark+refine, DenseBoxLM) bf16 training on 1xB200"
        c.keep = None
        torch.cuda.empty_cache()
        records["config5"] = run_inference(c)
        torch.cuda.empty_cache()
        records["dropin"] = run_dropin(c)
        torch.cuda.empty_cache()
        records["e2e_u8"] = run_e2e_u8(c, steps=args.steps)
        torch.cuda.empty_cache()
    if full and world > 1:
        r4 = run_training(c, "lm", 32, args.steps, args.warmup, pg, graph=not args.no_graph, e2e=True, exposed=True)
        r4["workload"] = ("configs[3]: batch=%d data-parallel multi-task (DenseBoxLM) training on %dxB200, NCCL grad "
                          "allreduce(sum) + 1-int positive-count allreduce" % (32 * world, world))
        c.keep = None
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:  # the same variant on ONE GPU of this box (the other ranks wait): base of the LM scaling factor
            r1 = run_training_single(c, "lm", 32, args.steps, args.warmup, not args.no_graph)
            r4["single_gpu_same_box"] = r1
        barrier()
        records["config4"] = r4

    if rank == 0:
        if c.sampler is not None:
            c.sampler.stop()
        line = {
            "metric": METRIC, "value": main_rec["value"], "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": headline_config(world, not args.no_graph) if (variant == "densebox" and B == 32) else
                      {"workload": "240x240 bf16 training step, %s, per-GPU batch %d (ad-hoc)" % (variant, B),
                       "variant": variant, "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world},
            "e2e": main_rec["e2e"],
            "gpu_launches": int(main_rec["gpu_launches_per_step"] * args.steps),
            "gpu_launches_per_step": main_rec["gpu_launches_per_step"],
            "clocks": main_rec["clocks"], "roofline": main_rec.get("roofline"), "kernels": main_rec.get("kernels"),
            "cpu_baseline": cpu_base, "loss": main_rec["loss"], "gflop_per_patch": GFLOP_TRAIN[variant],
            "workspace_gb": main_rec["workspace_gb"],
        }
        if "parity" in main_rec:
            line["parity"] = main_rec["parity"]
            line["loss_rel_err"] = main_rec["parity"]["loss_rel_err"]
        for k in ("exposed_comm_us_per_step", "ms_per_step_without_allreduce"):
            if k in main_rec:
                line[k] = main_rec[k]
        line.update(records)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_training_single(c, variant, B, steps, warmup, graph):
    """Rank 0 alone: the single-GPU number of `variant` on the same box (no process group)."""
    c1 = Ctx()
    c1.dev, c1.rank, c1.world, c1.dist, c1.pk, c1.sampler = c.dev, 0, 1, c.dist, c.pk, None
    import torch
    c1.barrier = torch.cuda.synchronize
    rec = run_training(c1, variant, B, steps, warmup, None, graph=graph, e2e=False)
    c1.keep = None
    torch.cuda.empty_cache()
    return {"value": rec["value"], "unit": "patches/s", "ms_per_step": rec["ms_per_step"]}


if __name__ == "__main__":
    main()
