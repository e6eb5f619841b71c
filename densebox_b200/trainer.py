"""`DenseBoxTrainer` — the reference training-loop body (DenseBox.py:2836-2926 and its LM / LMLOC twins) as one native
step: H2D of the batch, forward, fused loss, hand-written backward, gradient all-reduce (data parallel), SGD.

The reference's loop crosses the host/device boundary 7+ times per step and synchronises twice; here nothing leaves
the GPU between the input copy and the scalar loss, and the host never has to wait for the device:
  * the batch lives in one of TWO input slots; `prefetch()` copies the next host batch into the idle slot on a side
    stream while the current step computes, and each slot has its own captured CUDA graph (forward + loss + backward +
    SGD in ONE graph launch) that reads the slot in place — no per-step restaging copy;
  * the loss of every step lands in a ring of device scalars (the tensor `step()` returns stays valid for
    `LOSS_RING` further steps) and, with `async_loss=True`, in a pinned host word behind an event, so a training
    loop can log the loss of step i while step i+1 is already queued (`PendingLoss.item()`).
Data parallel: one process per GPU, batch sharded by rank, gradient SUM all-reduce (the loss is a sum,
DenseBox.py:2917), and the negative quota uses the batch-GLOBAL positive count (:2864-2868) via a 1-int all-reduce.
Both exchanges are hidden behind compute: the count all-reduce runs while the forward graph replays (only the loss
needs it), and the backward pass is cut into three stages (heads + conv4 | conv3 | conv2 + conv1) whose gradient
buckets are contiguous ranges of the flat gradient buffer: the all-reduce of bucket k runs under stage k + 1, only the
last bucket (1 MB: conv1/conv2 filters + biases) is exposed.  The overlapped all-reduces run on a dedicated NCCL
communicator limited to `nccl_max_ctas` CTAs, and the persistent tensor-core kernels of the overlapped stages are
launched on `num_sms - nccl_max_ctas` SMs: a persistent kernel with a static tile schedule must be fully resident,
otherwise the CTAs that wait for NCCL to leave their SM start late and the launch takes up to twice as long (the
round-1 build lost 0.26 ms per step at 8 GPUs to exactly that).
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream_ptr
from .engine import NetEngine, unique_param_names

c_int = ctypes.c_int
LOSS_RING = 256


class PendingLoss:
    """The loss of one step, read back without stalling the launch of the next one: `tensor` is the device scalar,
    `item()` waits for the 4-byte device-to-host copy that was queued right behind the step."""

    __slots__ = ("tensor", "_host", "_event")

    def __init__(self, tensor, host, event):
        self.tensor, self._host, self._event = tensor, host, event

    def item(self):
        self._event.synchronize()
        return float(self._host)

    def __float__(self):
        return self.item()


class DenseBoxTrainer:
    def __init__(self, net, batch_size, lr=1e-9, momentum=0.9, weight_decay=5e-8, lambda_loc=3.0, lambda_det=1.0,
                 lambda_lm=0.5, patch=240, rand_width=256, process_group=None, use_cuda_graph=True, dropout=True,
                 device=None, seed=0, allreduce_loss=False, nccl_max_ctas=8, count_exchange="peer", input_u8=False,
                 mean=None, std=None, reserve_stages=(1, 2), staged=False, sm_reserve=None):
        """input_u8=True: `step()` / `prefetch()` take the batch as decoded image bytes, uint8 [B,patch,patch,3]
        (densebox_b200.data.load_patch_u8), and ToTensor + Normalize(mean, std) (DenseBox.py:766-772) happen inside
        the first kernel — a quarter of the host->device bytes of the fp32 [B,3,patch,patch] form."""
        self.net = net
        self.variant = net.variant
        self.B = batch_size
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.lambdas = (lambda_loc, lambda_det, lambda_lm)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self.allreduce_loss = bool(allreduce_loss) and self.world > 1
        self.pg_overlap, self.sm_reserve, self.reserve_stages = self.pg, 0, tuple(reserve_stages)
        if self.world > 1 and nccl_max_ctas and torch.distributed.get_backend(process_group) == "nccl":
            try:  # a second communicator whose kernels occupy at most `nccl_max_ctas` SMs (ncclConfig_t.maxCTAs)
                opts = torch.distributed.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(nccl_max_ctas)
                opts.config.min_ctas = 1
                ranks = torch.distributed.get_process_group_ranks(process_group)
                self.pg_overlap = torch.distributed.new_group(ranks=ranks, backend="nccl", pg_options=opts)
                self.sm_reserve = (int(nccl_max_ctas) + 1) // 2 * 2
            except Exception:  # older torch / NCCL without per-communicator config: plain overlap, no reservation
                self.pg_overlap, self.sm_reserve = self.pg, 0
        self.staged = bool(staged)   # measurement aid: run the data-parallel step structure on one GPU
        if sm_reserve is not None:
            self.sm_reserve = int(sm_reserve)
        self.dropout = dropout
        # independent dropout streams per rank (the reference's nn.Dropout draws per process): fold the rank in
        self.seed = (seed ^ (self.rank * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF if self.world > 1 else seed
        self.step_no = 0
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.eng = NetEngine(self.variant, batch_size, patch, patch, train=True, device=self.device)
        dev = self.device
        with torch.cuda.device(dev):
            self.slots = [self._new_slot(batch_size, patch, rand_width, dev, input_u8) for _ in range(2)]
            if input_u8:
                self.eng.set_ingest(mean, std)
            self.gpos = torch.zeros(1, dtype=torch.int32, device=dev)
            self._loss_ring = torch.zeros(LOSS_RING, device=dev)
            self._loss_host = torch.zeros(LOSS_RING, pin_memory=True)
            self._loss_events = [None] * LOSS_RING
        # The batch-global positive count (:2864-2868) without a collective on the critical path: every rank stores
        # its count straight into a slot of every peer's buffer (symmetric memory over NVLink, dbx_count_exchange) at
        # the start of the step and the loss kernel sums the slots a forward pass later.  (As a 1-int NCCL all-reduce
        # in front of the forward pass the same exchange cost 0.13 ms per step at 8 GPUs.)
        self._slots, self._peer_ptrs = None, None
        if self.world > 1:
            self._setup_count_slots(count_exchange)
        self.use_graph = use_cuda_graph
        self._graphs = {}          # capture key -> CUDAGraph (single GPU: whole step) / tuple of 3 (data parallel)
        self._graph_sgd = {}       # lr -> CUDAGraph of the SGD update (data parallel: it follows the all-reduce)
        self._skip_allreduce = False   # measurement aid (bench.py: exposed-communication time); never set in training
        self._skip_mask = 0            # measurement aid (tools/dp_probe.py): bit 0 count, 1..3 gradient buckets 0..2
        self._rng = self.eng.buffer("rng", torch.int64)
        drop_elems = self.eng.buffer("drop", torch.bfloat16).numel()
        self._rng_stride = (drop_elems + 127) // 128  # Philox calls consumed by one step
        # {seed, offset}: the offset advances by one stride per step INSIDE the captured step (device side, no
        # per-step host write); it starts one stride before 0 so that step n draws from offset n * stride
        self._rng[0] = self.seed
        self._rng[1] = -self._rng_stride
        self._rng_off = self._rng[1:2]
        # input slots: `_cur` is the slot the next step() uses; prefetch() fills it on the copy stream
        self._cur = 0
        self._staged = None        # (identity key of the host tensors, slot) of the last prefetch()
        self._pf_stream = None
        self._slot_ready = [None, None]   # events: H2D into the slot complete (copy stream)
        self._slot_free = [None, None]    # events: the step that read the slot has finished (compute stream)
        self.load_from_module()

    def _setup_count_slots(self, mode):
        dist, ok, t, ptrs, hdl = torch.distributed, 0, None, None, None
        if mode == "peer" and dist.get_backend(self.pg) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                with torch.cuda.device(self.device):
                    t = symm_mem.empty(2 * self.world + 2, dtype=torch.int64, device=self.device)
                    t.zero_()
                    torch.cuda.synchronize(self.device)
                    hdl = symm_mem.rendezvous(t, group=self.pg)
                    ptrs = [int(p) for p in hdl.buffer_ptrs]
                ok = int(len(ptrs) == self.world and all(ptrs))
            except Exception:
                ok = 0
        flag = torch.tensor([ok], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)  # all ranks take the same path; also orders the zeroing
        torch.cuda.synchronize(self.device)
        if int(flag.item()) == 1:
            self._slots, self._symm_hdl = t, hdl  # the handle keeps the peer mappings alive
            self._peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
            check(lib().dbx_net_set_count_slots(self.eng.h, ptr(t), c_int(self.world)), "net_set_count_slots")

    @staticmethod
    def _new_slot(B, patch, rand_width, dev, input_u8=False):
        x = (torch.zeros(B, patch, patch, 3, dtype=torch.uint8, device=dev) if input_u8
             else torch.zeros(B, 3, patch, patch, device=dev))
        return {"x": x, "bbox": torch.zeros(B, 4, device=dev),
                "vertices": torch.zeros(B, 8, device=dev), "labels": torch.ones(B, device=dev),
                "rand": torch.zeros(B, rand_width, dtype=torch.int64, device=dev),
                "lm_rand": torch.zeros(B, 4, dtype=torch.int64, device=dev), "labels_are_ones": True}

    # ---- parameters
    def load_from_module(self):
        with torch.cuda.device(self.device):
            self.eng.set_params(self.net, device=self.device)
            self.eng.refresh_dgrad()
            self.eng.zero_grad()

    @torch.no_grad()
    def store_to_module(self):
        with torch.cuda.device(self.device):
            self.eng.get_params(self.net)

    # ---- input staging
    @staticmethod
    def _key(*tensors):
        return tuple((t.data_ptr(), tuple(t.shape)) if torch.is_tensor(t) else None for t in tensors)

    def _fill_slot(self, k, x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx):
        """Copy one batch into slot k on the CURRENT stream (host tensors: asynchronous when pinned)."""
        s = self.slots[k]

        def put(dst, src):
            src = torch.as_tensor(src)
            dst.copy_(src.reshape(dst.shape) if src.numel() == dst.numel() else src, non_blocking=True)

        put(s["x"], x)
        put(s["bbox"], bbox)
        if vertices is not None:
            put(s["vertices"], vertices)
        if labels is not None:
            put(s["labels"], labels)
            s["labels_are_ones"] = False
        elif not s["labels_are_ones"]:
            s["labels"].fill_(1.0)
            s["labels_are_ones"] = True
        if rand_neg_idx is None:
            s["rand"].copy_(torch.rand(self.B, 3600, device=self.device).argsort(dim=1)[:, :s["rand"].shape[1]])
        else:
            r = torch.as_tensor(rand_neg_idx)
            w = min(r.shape[1], s["rand"].shape[1])
            if w == s["rand"].shape[1] and r.shape[1] == w:
                s["rand"].copy_(r, non_blocking=True)
            else:
                s["rand"][:, :w].copy_(r[:, :w], non_blocking=True)
        if self.variant != "densebox":
            if lm_rand_neg_idx is None:
                s["lm_rand"].copy_(torch.randint(0, 3600, (self.B, 4), device=self.device))
            else:
                put(s["lm_rand"], lm_rand_neg_idx)

    def prefetch(self, x, bbox, vertices=None, labels=None, rand_neg_idx=None, lm_rand_neg_idx=None):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a side stream, straight into the
        input slot the next `step()` will read; that `step()`, called with the same tensors, copies nothing.  The
        usual prefetching-loader pattern: `step(batch_i)`, `prefetch(batch_i+1)`, then read the loss.  The staged
        copy is matched by tensor identity (data pointer + shape): do not modify the host tensors between
        `prefetch()` and the `step()` that consumes them."""
        args = (x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx)
        if not all(v is None or (torch.is_tensor(v) and not v.is_cuda) for v in args):
            return  # device tensors or non-tensors: nothing to overlap, step() handles them
        if rand_neg_idx is None or (self.variant != "densebox" and lm_rand_neg_idx is None):
            return  # device-side draws happen in step()
        k = self._cur
        with torch.cuda.device(self.device):
            if self._pf_stream is None:
                self._pf_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(self._pf_stream):
                if self._slot_free[k] is not None:
                    self._pf_stream.wait_event(self._slot_free[k])  # the step that last read this slot is done
                self._fill_slot(k, *args)
                if self._slot_ready[k] is None:
                    self._slot_ready[k] = torch.cuda.Event()
                self._slot_ready[k].record(self._pf_stream)
        self._staged = (self._key(*args), k)

    # ---- the step
    def _loss(self, s, clamp_lm):
        ll, ld, lm = self.lambdas
        self.eng.loss(s["bbox"], vertices=s["vertices"] if self.variant != "densebox" else None, labels=s["labels"],
                      rand_idx=s["rand"], lm_rand_idx=s["lm_rand"] if self.variant != "densebox" else None,
                      lambda_loc=ll, lambda_det=ld, lambda_lm=lm,
                      global_pos_dev=self.gpos if (self.world > 1 and self._slots is None) else None,
                      global_batch=self.B * self.world if self.world > 1 else -1, clamp_lm=clamp_lm)

    def _forward(self, s):
        if self.dropout:
            self._rng_off.add_(self._rng_stride)  # a fresh dropout mask per step (captured with the step)
        self.eng.forward(s["x"], dropout_mode=3 if self.dropout else 0)  # Philox in the epilogues, state in `rng`

    def _fwd_loss_bwd(self, s=None, clamp_lm=False):
        s = s if s is not None else self.slots[self._cur]
        self._forward(s)
        self._loss(s, clamp_lm)
        self.eng.backward()

    def _step_single(self, k, clamp_lm, graph_ok):
        s = self.slots[k]
        if not graph_ok:
            self._fwd_loss_bwd(s, clamp_lm)
            self.eng.sgd_step(self.lr, self.momentum, self.weight_decay)
            return
        key = (k, clamp_lm, self.lr)
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._fwd_loss_bwd(s, clamp_lm)
                self.eng.sgd_step(self.lr, self.momentum, self.weight_decay)
            self._graphs[key] = g  # capture does not execute: fall through to the replay
        g.replay()

    def _reserved(self, fn):
        """Run (or capture) `fn` with the tensor kernels confined to num_sms - sm_reserve SMs."""
        if not self.sm_reserve:
            return fn()
        L = lib()
        n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        old = L.dbx_set_tensor_sm_limit(c_int(n_sm - self.sm_reserve))
        try:
            return fn()
        finally:
            L.dbx_set_tensor_sm_limit(c_int(old))

    def _step_dp(self, k, clamp_lm, graph_ok):
        """forward | loss + backward(heads, conv4) | backward(conv3) | backward(conv2, conv1); the all-reduce of each
        gradient bucket runs under the next stage."""
        e, dist, s = self.eng, torch.distributed, self.slots[k]

        def fwd():
            self._forward(s)
            e.join()

        def lb0():
            self._loss(s, clamp_lm)
            e.backward_stage(0)

        def stage(i):
            return (lambda: self._reserved(lambda: e.backward_stage(i))) if i in self.reserve_stages \
                else (lambda: e.backward_stage(i))

        parts = (fwd, lb0, stage(1), stage(2))
        key = (k, clamp_lm)
        if graph_ok and key not in self._graphs:  # capture before any collective of this step is in flight
            gs = []
            for fn in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    fn()
                gs.append(g)
            self._graphs[key] = tuple(gs)
        comm = not self._skip_allreduce and self.world > 1
        sk = self._skip_mask
        if self.world == 1:
            pass  # staged single-GPU measurement: the loss kernel counts its own batch
        elif self._slots is not None:  # peer stores, no collective (runs in measurement modes too: the loss waits for it)
            check(lib().dbx_count_exchange(ptr(s["bbox"]), ptr(s["labels"]), c_int(self.B), self._peer_ptrs,
                                           c_int(self.world), c_int(self.rank), ptr(self._slots), stream_ptr()),
                  "count_exchange")
        else:
            check(lib().dbx_count_positives(ptr(s["bbox"]), ptr(s["labels"]), c_int(self.B), ptr(self.gpos),
                                            stream_ptr()), "count_positives")
            # The fallback exchange completes BEFORE the forward pass starts: overlapped with it, the NCCL kernel takes
            # SMs at some kernel boundary and the next persistent convolution, launched on all SMs, waits for them.
            if comm and not sk & 1:
                dist.all_reduce(self.gpos, group=self.pg_overlap, async_op=True).wait()
        run = [g.replay for g in self._graphs[key]] if graph_ok else parts
        run[0]()
        run[1]()
        works = []
        if comm and not sk & 2:  # SUM: the loss is a sum (:2917)
            works.append(dist.all_reduce(e.grad_bucket(0), group=self.pg_overlap, async_op=True))
        run[2]()
        if comm and not sk & 4:
            works.append(dist.all_reduce(e.grad_bucket(1), group=self.pg_overlap, async_op=True))
        run[3]()
        w_tail = None
        if comm and not sk & 8:
            w_tail = dist.all_reduce(e.grad_bucket(2), group=self.pg, async_op=True)  # exposed: full-speed group
        for w in works:
            w.wait()
        # SGD of buckets 0 + 1 (96 % of the parameters) runs while the last bucket is still being reduced
        e.sgd_step_part(0, self.lr, self.momentum, self.weight_decay)
        if w_tail is not None:
            w_tail.wait()
        e.sgd_step_part(1, self.lr, self.momentum, self.weight_decay)

    def step(self, x, bbox, vertices=None, labels=None, rand_neg_idx=None, lm_rand_neg_idx=None, async_loss=False):
        """x [B,3,240,240] fp32 (host pinned or device), labels in 60-space.  Returns the loss of this rank's shard
        (of the whole batch with `allreduce_loss=True`): a 0-dim CUDA tensor that stays valid for the next
        `LOSS_RING` steps (call .item() to read it back), or with `async_loss=True` a `PendingLoss` whose `.item()`
        waits only for a 4-byte copy queued behind this step — read it after launching the next step and the host
        never stalls the GPU."""
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            args = (x, bbox, vertices, labels, rand_neg_idx, lm_rand_neg_idx)
            if self._staged is not None and self._staged[0] == self._key(*args):
                k = self._staged[1]
                cur.wait_event(self._slot_ready[k])
                if labels is not None:
                    self.slots[k]["labels_are_ones"] = False
            else:
                k = self._cur
                if self._pf_stream is not None and self._slot_ready[k] is not None:
                    cur.wait_event(self._slot_ready[k])  # an unconsumed prefetch into this slot must land first
                self._fill_slot(k, *args)
            self._staged = None  # a prefetch that was not consumed is dropped (its key can never match stale memory)
            clamp_lm = labels is not None  # `_pn` helpers of train_densebox_online (:1899-1907)
            graph_ok = self.use_graph and self.step_no >= 1  # step 0 runs eagerly (one-time inits, SGD first-step flag)
            if self.world > 1 or self.staged:
                self._step_dp(k, clamp_lm, graph_ok)
            else:
                self._step_single(k, clamp_lm, graph_ok)
            j = self.step_no % LOSS_RING
            out = self._loss_ring[j]
            out.copy_(self.eng.loss_value())  # an independent value: the engine's scalar is overwritten every step
            if self.allreduce_loss:
                torch.distributed.all_reduce(out, group=self.pg)
            if self._slot_free[k] is None:
                self._slot_free[k] = torch.cuda.Event()
            self._slot_free[k].record(cur)
            self._cur = 1 - k
            self.step_no += 1
            if not async_loss:
                return out
            self._loss_host[j:j + 1].copy_(out.reshape(1), non_blocking=True)
            ev = self._loss_events[j]
            if ev is None:
                ev = self._loss_events[j] = torch.cuda.Event()
            ev.record(cur)
            return PendingLoss(out, self._loss_host[j], ev)
