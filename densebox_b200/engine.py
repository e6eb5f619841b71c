"""Python handle on the native network engine (dbx_net_* in include/densebox_b200.h).

torch is used for device memory (the workspace is one torch.uint8 tensor), streams and views of the engine's
buffers; every computation is a call into libdensebox_b200.so.
"""
import ctypes

import torch

from ._lib import DbxError, check, lib, ptr, stream_ptr

VARIANTS = {"densebox": 0, "lm": 1, "lmloc": 2}
HEAD_NAMES = {0: ["det", "loc"], 1: ["det", "loc", "landmark"], 2: ["det", "loc", "landmark", "lmloc"]}
BACKBONE = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_4", "conv4_1", "conv4_2",
            "conv4_3", "conv4_4"]

c_int, c_long, c_float, c_ull = ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_ulonglong


def unique_param_names(variant):
    """Names of the (weight, bias) pairs the engine consumes, in the reference's naming (conv3_3 is unused)."""
    v = VARIANTS[variant] if isinstance(variant, str) else variant
    names = list(BACKBONE)
    names += ["conv5_1_" + h for h in HEAD_NAMES[v]] + ["conv5_2_" + h for h in HEAD_NAMES[v]]
    if v >= 1:
        names += ["conv6_1_det", "conv6_2_det", "conv6_3_det"]
    return names


class NetEngine:
    """One replica for a fixed input shape [N,3,H,W] on the current CUDA device."""

    def __init__(self, variant, N, H, W, train=True, device=None):
        if not torch.cuda.is_available():
            raise DbxError("densebox_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.N, self.H, self.W, self.train = N, H, W, bool(train)
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        L = lib()
        nbytes = ctypes.c_size_t(0)
        check(L.dbx_net_workspace_bytes(c_int(self.variant), c_int(N), c_int(H), c_int(W), c_int(int(self.train)),
                                        ctypes.byref(nbytes)), "net_workspace_bytes")
        self.workspace_bytes = nbytes.value
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
            self._ws_off = (-self.ws.data_ptr()) % 256
            h = ctypes.c_void_p(0)
            check(L.dbx_net_create(c_int(self.variant), c_int(N), c_int(H), c_int(W), c_int(int(self.train)),
                                   ctypes.c_void_p(self.ws.data_ptr() + self._ws_off), ctypes.c_size_t(nbytes.value),
                                   stream_ptr(), ctypes.byref(h)), "net_create")
        self.h = h
        self.HC = L.dbx_net_head_channels(self.h)
        self.h4, self.w4 = H // 4, W // 4
        self.generation = 0        # bumped by every forward(): autograd nodes check that their activations are live
        self._param_key = None     # identity + version of the module parameters last packed (modules._sync_params)
        self._dgrad_fresh = False  # the flipped/transposed filter copies match the current weights

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().dbx_net_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- buffers as torch views (no copies)
    def buffer(self, name, dtype, shape=None):
        p, n = ctypes.c_void_p(0), ctypes.c_size_t(0)
        check(lib().dbx_net_buffer(self.h, name.encode(), ctypes.byref(p), ctypes.byref(n)), "net_buffer(%s)" % name)
        off = p.value - self.ws.data_ptr()
        t = self.ws[off:off + n.value].view(dtype)
        return t.view(shape) if shape is not None else t

    def head_out(self):
        return self.buffer("head_out", torch.float32, (self.N, self.h4, self.w4, self.HC))

    def rf_out(self):
        return self.buffer("rf_out", torch.float32, (self.N, self.h4, self.w4, 16))

    def scalars(self):
        return self.buffer("scalars", torch.float32)

    def loss_value(self):
        return self.scalars()[0]

    def loss_info(self):
        """(half, pos) of the last loss launch; raises if fewer random negatives were supplied than the quota needs
        (the reference draws `half` of them per sample, DenseBox.py:2888-2893).  Synchronises."""
        i = self.buffer("scalars", torch.int32)[2:5].tolist()
        if i[2]:
            raise DbxError("densebox loss: the negative quota half=%d exceeds the %s random negatives per sample that "
                           "were supplied" % (i[0], "rand_neg_idx"))
        return i[0], i[1]

    # ---- parameters
    @staticmethod
    def _strides(t):
        s = list(t.stride()) + [0, 0, 0]
        return [c_long(v) for v in s[:4]]

    def set_param(self, name, weight, bias):
        w = weight.detach()
        b = bias.detach()
        assert w.dtype == torch.float32 and w.is_cuda and b.dtype == torch.float32
        L = lib()
        check(L.dbx_net_set_param(self.h, name.encode(), c_int(0), ptr(w), *self._strides(w), stream_ptr()),
              "set_param(%s.weight)" % name)
        check(L.dbx_net_set_param(self.h, name.encode(), c_int(1), ptr(b), *self._strides(b), stream_ptr()),
              "set_param(%s.bias)" % name)

    def get_tensor(self, name, like_w, like_b, grad=False):
        L = lib()
        fn = L.dbx_net_get_grad if grad else L.dbx_net_get_param
        w = torch.empty_like(like_w, dtype=torch.float32, memory_format=torch.contiguous_format)
        b = torch.empty_like(like_b, dtype=torch.float32, memory_format=torch.contiguous_format)
        check(fn(self.h, name.encode(), c_int(0), ptr(w), *self._strides(w), stream_ptr()), "get(%s.weight)" % name)
        check(fn(self.h, name.encode(), c_int(1), ptr(b), *self._strides(b), stream_ptr()), "get(%s.bias)" % name)
        return w, b

    def dropout_stride(self):
        """Philox counters one forward consumes (one call covers 128 elements of the dropped activation)."""
        nh = len(HEAD_NAMES[self.variant])
        return (self.N * self.h4 * self.w4 * 512 * nh + 127) // 128

    def refresh_dgrad(self):
        check(lib().dbx_net_refresh_dgrad(self.h, stream_ptr()), "refresh_dgrad")
        self._dgrad_fresh = True

    # ---- all parameters of a module in one launch (dbx_net_xfer_params)
    def _xfer(self, mode, triples, what):
        n = len(triples)
        if getattr(self, "_xfer_names", None) is None or len(self._xfer_names) != n:
            self._xfer_names = (ctypes.c_char_p * n)(*[t[0].encode() for t in triples])
            self._xfer_wp, self._xfer_bp = (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)()
            self._xfer_ws, self._xfer_bs = (ctypes.c_long * (4 * n))(), (ctypes.c_long * n)()
        wp, bp, ws, bs = self._xfer_wp, self._xfer_bp, self._xfer_ws, self._xfer_bs
        for i, (_, w, b) in enumerate(triples):
            assert w.dtype == torch.float32 and b.dtype == torch.float32 and w.is_cuda and b.is_cuda
            wp[i], bp[i] = w.data_ptr(), b.data_ptr()
            st = w.stride()
            for k in range(4):
                ws[4 * i + k] = st[k] if k < len(st) else 0
            bs[i] = b.stride(0)
        check(lib().dbx_net_xfer_params(self.h, c_int(mode), c_int(n), self._xfer_names, wp, ws, bp, bs, stream_ptr()),
              what)

    def _param_triples(self, module, device=None):
        out = []
        for name in unique_param_names(self.variant):
            w, b = module._wb(name)
            w, b = w.detach(), b.detach()
            if device is not None and w.device != device:
                w, b = w.to(device), b.to(device)
            out.append((name, w, b))
        return out

    def set_params(self, module, device=None):
        """Pack every (weight, bias) the variant uses from a drop-in module (or any object with `_wb(name)`)."""
        self._xfer(0, self._param_triples(module, self.device if device is None else device), "xfer_params(set)")
        self._dgrad_fresh = False

    def get_params(self, module):
        """Write the engine's fp32 master weights back into the module's parameters (in place)."""
        self._xfer(1, self._param_triples(module), "xfer_params(get)")
        if hasattr(module, "_param_epoch"):
            module._param_epoch += 1  # raw in-place writes do not bump torch's version counters

    def get_grads(self, module):
        """Fresh fp32 tensors shaped like the module's parameters holding the engine's gradients: [w0, b0, w1, ...]
        (views of ONE flat allocation: a single allocator call per backward instead of one per tensor)."""
        shapes = []
        for name in unique_param_names(self.variant):
            w, b = module._wb(name)
            shapes += [w.shape, b.shape]
        sizes = [s.numel() for s in shapes]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=self.device)
        out = [t.view(s) for t, s in zip(flat.split(sizes), shapes)]
        names = unique_param_names(self.variant)
        self._xfer(2, [(names[i], out[2 * i], out[2 * i + 1]) for i in range(len(names))], "xfer_params(grads)")
        return out

    def get_outputs(self, want):
        """NCHW fp32 copies of the head maps: want = subset of ('score','loc','lm','lmloc','rf') -> dict."""
        ch = {"score": 1, "loc": 4, "lm": 4, "lmloc": 8, "rf": 1}
        out = {k: torch.empty(self.N, ch[k], self.h4, self.w4, device=self.device) for k in want}
        check(lib().dbx_net_get_outputs(self.h, *[ptr(out.get(k)) for k in ("score", "loc", "lm", "lmloc", "rf")],
                                        stream_ptr()), "net_get_outputs")
        return out

    def set_output_grads(self, grads):
        """grads: dict name -> fp32 NCHW tensor or None (zero); fills the bf16 d_head / d_rf regions."""
        g = {}
        for k in ("score", "loc", "lm", "lmloc", "rf"):
            t = grads.get(k)
            g[k] = t.float().contiguous() if t is not None else None
        check(lib().dbx_net_set_output_grads(self.h, *[ptr(g[k]) for k in ("score", "loc", "lm", "lmloc", "rf")],
                                             stream_ptr()), "net_set_output_grads")

    # ---- the path
    def set_ingest(self, mean=None, std=None):
        """ToTensor + Normalize table for uint8 inputs (densebox_b200.data.ingest_table); default: ImageNet statistics."""
        from .data import IMAGENET_MEAN, IMAGENET_STD, ingest_table
        self._lut = ingest_table(mean or IMAGENET_MEAN, std or IMAGENET_STD).to(self.device)

    def forward(self, x, dropout_mode=0, seed=0, offset=0):
        """x: fp32 [N,3,H,W], normalised (the reference's forward argument) — or uint8 [N,H,W,3], the decoded image
        bytes, normalised inside the first kernel (dbx_net_forward_u8)."""
        assert x.is_cuda and x.is_contiguous()
        if x.dtype == torch.uint8:
            assert tuple(x.shape) == (self.N, self.H, self.W, 3), (tuple(x.shape), (self.N, self.H, self.W, 3))
            if getattr(self, "_lut", None) is None:
                self.set_ingest()
            check(lib().dbx_net_forward_u8(self.h, ptr(x), ptr(self._lut), c_int(dropout_mode), c_ull(seed),
                                           c_ull(offset), stream_ptr()), "net_forward_u8")
        else:
            assert x.dtype == torch.float32
            assert tuple(x.shape) == (self.N, 3, self.H, self.W), (tuple(x.shape), (self.N, 3, self.H, self.W))
            check(lib().dbx_net_forward(self.h, ptr(x), c_int(dropout_mode), c_ull(seed), c_ull(offset), stream_ptr()),
                  "net_forward")
        self.generation += 1

    def loss(self, bbox, vertices=None, labels=None, rand_idx=None, lm_rand_idx=None, lambda_loc=3.0, lambda_det=1.0,
             lambda_lm=0.5, global_pos=-1, global_batch=-1, global_pos_dev=None, clamp_lm=False, d_head_f32=None,
             d_rf_f32=None, mask_out=None, lm_mask_out=None):
        for t, dt in ((bbox, torch.float32), (vertices, torch.float32), (labels, torch.float32),
                      (rand_idx, torch.int64), (lm_rand_idx, torch.int64), (global_pos_dev, torch.int32)):
            assert t is None or (t.is_cuda and t.dtype == dt and t.is_contiguous())
        check(lib().dbx_net_loss(
            self.h, ptr(bbox), ptr(vertices), ptr(labels), ptr(rand_idx),
            c_int(rand_idx.shape[1] if rand_idx is not None else 0), ptr(lm_rand_idx), c_float(lambda_loc),
            c_float(lambda_det), c_float(lambda_lm), c_int(global_pos), c_int(global_batch), ptr(global_pos_dev),
            c_int(int(clamp_lm)), ptr(d_head_f32), ptr(d_rf_f32), ptr(mask_out), ptr(lm_mask_out), stream_ptr()),
            "net_loss")

    def backward(self):
        check(lib().dbx_net_backward(self.h, stream_ptr()), "net_backward")

    def backward_stage(self, stage):
        """A third of backward(): 0 = refine + heads + conv4 block, 1 = conv3 block, 2 = conv2 + conv1 blocks."""
        check(lib().dbx_net_backward_stage(self.h, c_int(stage), stream_ptr()), "net_backward_stage")

    def join(self):
        check(lib().dbx_net_join(self.h, stream_ptr()), "net_join")

    def grad_bucket(self, bucket):
        """View of g32 holding gradient bucket 0 (conv4..heads filters), 1 (conv3 filters) or 2 (conv1/conv2 filters
        and every bias) — contiguous ranges, complete after backward stage 0 / 1 / 2."""
        first, count = ctypes.c_longlong(0), ctypes.c_longlong(0)
        check(lib().dbx_net_grad_bucket(self.h, c_int(bucket), ctypes.byref(first), ctypes.byref(count)),
              "net_grad_bucket")
        return self.flat_grads()[first.value:first.value + count.value]

    def zero_grad(self):
        check(lib().dbx_net_zero_grad(self.h, stream_ptr()), "net_zero_grad")

    def sgd_step(self, lr, momentum=0.9, weight_decay=5e-8):
        check(lib().dbx_net_sgd_step(self.h, c_float(lr), c_float(momentum), c_float(weight_decay), stream_ptr()),
              "net_sgd_step")

    def sgd_step_part(self, part, lr, momentum=0.9, weight_decay=5e-8):
        """part 0 = gradient buckets 0 + 1, part 1 = the last bucket: see dbx_net_sgd_step_part."""
        check(lib().dbx_net_sgd_step_part(self.h, c_int(part), c_float(lr), c_float(momentum), c_float(weight_decay),
                                          stream_ptr()), "net_sgd_step_part")

    def flat_grads(self):
        return self.buffer("g32", torch.float32)

    def flat_params(self):
        return self.buffer("w32", torch.float32)

    # ---- measurement aids
    def profile(self, enable=True):
        check(lib().dbx_net_profile(self.h, c_int(int(enable))), "net_profile")

    def launch_count(self):
        lib().dbx_net_launch_count.restype = ctypes.c_longlong
        return int(lib().dbx_net_launch_count(self.h))

    def profile_records(self):
        """[(tag, algorithmic flops, ms)] of every launch since profile(True); synchronises the device."""
        torch.cuda.synchronize(self.device)
        L = lib()
        out = []
        tag = ctypes.create_string_buffer(96)
        fl, ms = ctypes.c_double(0), ctypes.c_float(0)
        for i in range(L.dbx_net_profile_count(self.h)):
            check(L.dbx_net_profile_get(self.h, c_int(i), tag, c_int(96), ctypes.byref(fl), ctypes.byref(ms)),
                  "net_profile_get")
            out.append((tag.value.decode(), fl.value, ms.value))
        return out
