"""`DenseBoxTrainer` — the reference training-loop body (DenseBox.py:2836-2926 and its LM / LMLOC twins) as one native
step: H2D of the batch, forward, fused loss, hand-written backward, gradient all-reduce (data parallel), SGD.

The reference's loop crosses the host/device boundary 7+ times per step and synchronises twice; here nothing leaves
the GPU between the input copy and the scalar loss.  With `use_cuda_graph=True` forward+loss+backward and the SGD
update are replayed as CUDA graphs.
Data parallel: one process per GPU, batch sharded by rank, gradient SUM all-reduce (the loss is a sum,
DenseBox.py:2917), and the negative quota uses the batch-GLOBAL positive count (:2864-2868) via a 1-int all-reduce.
Both exchanges are hidden behind compute: the count all-reduce runs while the forward graph replays (only the loss
needs it), and the backward pass is cut after the conv4 block so that the all-reduce of the conv4/heads filters
(87 % of the gradient bytes) overlaps the conv3..conv1 half of backward; the small second bucket follows.
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream_ptr
from .engine import NetEngine, unique_param_names

c_int = ctypes.c_int


class DenseBoxTrainer:
    def __init__(self, net, batch_size, lr=1e-9, momentum=0.9, weight_decay=5e-8, lambda_loc=3.0, lambda_det=1.0,
                 lambda_lm=0.5, patch=240, rand_width=256, process_group=None, use_cuda_graph=True, dropout=True,
                 device=None, seed=0):
        self.net = net
        self.variant = net.variant
        self.B = batch_size
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.lambdas = (lambda_loc, lambda_det, lambda_lm)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.dropout = dropout
        self.seed, self.step_no = seed, 0
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.eng = NetEngine(self.variant, batch_size, patch, patch, train=True, device=self.device)
        dev = self.device
        self.x = torch.zeros(batch_size, 3, patch, patch, device=dev)
        self.bbox = torch.zeros(batch_size, 4, device=dev)
        self.vertices = torch.zeros(batch_size, 8, device=dev)
        self.labels = torch.ones(batch_size, device=dev)
        self.rand = torch.zeros(batch_size, rand_width, dtype=torch.int64, device=dev)
        self.lm_rand = torch.zeros(batch_size, 4, dtype=torch.int64, device=dev)
        self.gpos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.use_labels = False
        self.use_graph = use_cuda_graph
        self._graph_fb = self._graph_sgd = None
        self._graphs_dp = None  # (forward, loss + backward stage 0, backward stage 1) when world > 1
        self._graph_lr = None
        self._rng = self.eng.buffer("rng", torch.int64)
        drop_elems = self.eng.buffer("drop", torch.bfloat16).numel()
        self._rng_stride = (drop_elems + 127) // 128  # Philox calls consumed by one step
        # {seed, offset}: the offset advances by one stride per step INSIDE the captured step (device side, no
        # per-step host write); it starts one stride before 0 so that step n draws from offset n * stride
        self._rng[0] = seed
        self._rng[1] = -self._rng_stride
        self._rng_off = self._rng[1:2]
        # prefetch(): host batch i+1 is copied on a side stream into staging buffers while step i computes
        self._pf = None            # {"x": ..., "bbox": ..., ...} staging tensors (lazily allocated)
        self._pf_key = None        # identity of the host tensors staged last
        self._pf_stream = None
        self._pf_ready = None      # event: staging complete (recorded on the copy stream)
        self._pf_free = None       # event: staging buffers consumed (recorded on the compute stream)
        self.load_from_module()

    # ---- parameters
    def load_from_module(self):
        for name in unique_param_names(self.variant):
            w, b = self.net._wb(name)
            self.eng.set_param(name, w.detach().to(self.device), b.detach().to(self.device))
        self.eng.refresh_dgrad()
        self.eng.zero_grad()

    @torch.no_grad()
    def store_to_module(self):
        for name in unique_param_names(self.variant):
            w, b = self.net._wb(name)
            gw, gb = self.eng.get_tensor(name, w, b, grad=False)
            w.copy_(gw)
            b.copy_(gb)

    # ---- one step
    def _stage(self, dst, src):
        if src is None or src is dst:
            return
        src = torch.as_tensor(src)
        dst.copy_(src.reshape(dst.shape) if src.numel() == dst.numel() else src, non_blocking=True)

    @staticmethod
    def _key(*tensors):
        return tuple((t.data_ptr(), tuple(t.shape)) if torch.is_tensor(t) else None for t in tensors)

    def prefetch(self, x, bbox, vertices=None, labels=None, rand_neg_idx=None, lm_rand_neg_idx=None):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a side stream; a following
        `step()` called with the same tensors finds them in device staging buffers and only pays a device-to-device
        copy.  The usual prefetching-loader pattern: `step(batch_i)`, `prefetch(batch_i+1)`, then read the loss.
        The staged copy is matched by tensor identity (data pointer + shape): do not modify the host tensors between
        `prefetch()` and the `step()` that consumes them."""
        args = {"x": x, "bbox": bbox, "vertices": vertices, "labels": labels, "rand": rand_neg_idx,
                "lm_rand": lm_rand_neg_idx}
        if not all(v is None or (torch.is_tensor(v) and not v.is_cuda) for v in args.values()):
            return  # device tensors or non-tensors: nothing to overlap, step() handles them
        if self._pf is None:
            self._pf = {"x": torch.empty_like(self.x), "bbox": torch.empty_like(self.bbox),
                        "vertices": torch.empty_like(self.vertices), "labels": torch.empty_like(self.labels),
                        "rand": torch.empty_like(self.rand), "lm_rand": torch.empty_like(self.lm_rand)}
            self._pf_stream = torch.cuda.Stream(device=self.device)
            self._pf_ready = torch.cuda.Event()
            self._pf_free = torch.cuda.Event()
            self._pf_free.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._pf_stream):
            self._pf_stream.wait_event(self._pf_free)  # the previous staged batch has been consumed
            for k, v in args.items():
                if v is None:
                    continue
                dst = self._pf[k]
                if k == "rand":
                    dst[:, :v.shape[1]].copy_(v[:, :dst.shape[1]], non_blocking=True)
                else:
                    dst.copy_(v.reshape(dst.shape) if v.numel() == dst.numel() else v, non_blocking=True)
            self._pf_ready.record(self._pf_stream)
        self._pf_key = self._key(x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx)

    def _take_prefetched(self, x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx):
        """If exactly these host tensors were staged by prefetch(): substitute the staging buffers (device)."""
        if self._pf_key is None or self._pf_key != self._key(x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx):
            return x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx, False
        self._pf_key = None
        torch.cuda.current_stream(self.device).wait_event(self._pf_ready)
        pf = self._pf
        return (pf["x"], pf["bbox"], pf["vertices"] if vertices is not None else None,
                pf["labels"] if labels is not None else None,
                pf["rand"][:, :rand_neg_idx.shape[1]] if rand_neg_idx is not None else None,
                pf["lm_rand"] if lm_rand_neg_idx is not None else None, True)

    def _fwd_loss_bwd(self):
        e = self.eng
        if self.dropout:
            self._rng_off.add_(self._rng_stride)  # a fresh dropout mask per step (captured with the step)
        e.forward(self.x, dropout_mode=3 if self.dropout else 0)  # Philox in the epilogues, state in the rng region
        self._loss()
        e.backward()

    def _loss(self):
        ll, ld, lm = self.lambdas
        self.eng.loss(self.bbox, vertices=self.vertices if self.variant != "densebox" else None,
                      labels=self.labels if self.use_labels else None, rand_idx=self.rand,
                      lm_rand_idx=self.lm_rand if self.variant != "densebox" else None, lambda_loc=ll, lambda_det=ld,
                      lambda_lm=lm, global_pos_dev=self.gpos if self.world > 1 else None,
                      global_batch=self.B * self.world if self.world > 1 else -1, clamp_lm=self.use_labels)

    def _step_dp(self, graph_ok):
        """forward | loss + backward(heads, conv4) | backward(conv3..conv1) with the two exchanges overlapped."""
        e, dist = self.eng, torch.distributed

        def fwd():
            if self.dropout:
                self._rng_off.add_(self._rng_stride)
            e.forward(self.x, dropout_mode=3 if self.dropout else 0)
            e.join()

        def lb0():
            self._loss()
            e.backward_stage(0)

        parts = (fwd, lb0, lambda: e.backward_stage(1))
        if graph_ok and self._graphs_dp is None:  # capture before any collective of this step is in flight
            self._graphs_dp = []
            for fn in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    fn()
                self._graphs_dp.append(g)
        check(lib().dbx_count_positives(ptr(self.bbox), ptr(self.labels if self.use_labels else None),
                                        c_int(self.B), ptr(self.gpos), stream_ptr()), "count_positives")
        w_count = dist.all_reduce(self.gpos, group=self.pg, async_op=True)
        run = [g.replay for g in self._graphs_dp] if graph_ok else parts
        run[0]()
        w_count.wait()
        run[1]()
        works = [dist.all_reduce(e.grad_bucket(0), group=self.pg, async_op=True)]  # SUM: the loss is a sum (:2917)
        run[2]()
        works += [dist.all_reduce(e.grad_bucket(b), group=self.pg, async_op=True) for b in (1, 2)]
        for w in works:
            w.wait()

    def step(self, x, bbox, vertices=None, labels=None, rand_neg_idx=None, lm_rand_neg_idx=None):
        """x [B,3,240,240] fp32 (host pinned or device), labels in 60-space. Returns the loss as a 0-dim CUDA tensor
        (this rank's shard; call .item() to read it back)."""
        e = self.eng
        x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx, staged = self._take_prefetched(
            x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx)
        self._stage(self.x, x)
        self._stage(self.bbox, bbox)
        self._stage(self.vertices, vertices)
        if labels is not None:
            self._stage(self.labels, labels)
            self.use_labels = True
        if rand_neg_idx is None:
            self.rand.copy_(torch.rand(self.B, 3600, device=self.device).argsort(dim=1)[:, :self.rand.shape[1]])
        else:
            r = torch.as_tensor(rand_neg_idx)
            self.rand[:, :r.shape[1]].copy_(r[:, :self.rand.shape[1]], non_blocking=True)
        if self.variant != "densebox":
            if lm_rand_neg_idx is None:
                self.lm_rand.copy_(torch.randint(0, 3600, (self.B, 4), device=self.device))
            else:
                self._stage(self.lm_rand, lm_rand_neg_idx)
        if staged:  # the staging buffers may be refilled once these device-to-device copies are done
            self._pf_free.record(torch.cuda.current_stream(self.device))
        graph_ok = self.use_graph and self.step_no >= 1  # step 0 runs eagerly (one-time inits, SGD first-step flag)
        if self.world > 1:
            self._step_dp(graph_ok)
        else:
            if graph_ok and self._graph_fb is None:
                self._graph_fb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_fb):
                    self._fwd_loss_bwd()
            if graph_ok:
                self._graph_fb.replay()
            else:
                self._fwd_loss_bwd()
        if graph_ok and (self._graph_sgd is None or self._graph_lr != self.lr):
            self._graph_sgd = torch.cuda.CUDAGraph()
            self._graph_lr = self.lr
            with torch.cuda.graph(self._graph_sgd, capture_error_mode="thread_local"):
                e.sgd_step(self.lr, self.momentum, self.weight_decay)
            # capture does not execute: fall through to replay
        if graph_ok:
            self._graph_sgd.replay()
        else:
            e.sgd_step(self.lr, self.momentum, self.weight_decay)
        self.step_no += 1
        return e.loss_value()

    def kernels_per_step(self):
        """Number of this library's kernel launches in one training step (for bench.py's gpu_launches)."""
        convs = 13 + (3 if self.variant != "densebox" else 0)          # fprop
        fwd = 1 + convs + 3 + 1 + (2 if self.variant != "densebox" else 0)  # im2col, convs, pools, upsample, refine glue
        dgrads = convs - 1
        wgrads = convs * 2                                               # wgrad + bias colsum
        bwd = dgrads + wgrads + 3 + 1 + 1 + (2 if self.variant != "densebox" else 0)
        sgd = 1 + (convs - 1)
        return fwd + 1 + bwd + sgd + (1 if self.dropout else 0) + (1 if self.world > 1 else 0)
