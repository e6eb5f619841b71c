"""Drop-in `DenseBox`, `DenseBoxLM`, `DenseBoxLMLOC` modules (reference: DenseBox.py:31-228, :232-473, :477-738).

Same constructor (`Net(vgg19)` with any object exposing `.features._modules['0'..'26']`), same attribute names and
therefore the same `state_dict()` keys (64 / 78 / 86 keys, each backbone/head tensor under two names), same forward
return tuples (NCHW fp32, autograd-connected), `.train()/.eval()` toggling the head dropout.  The computation is the
native engine (tcgen05 convolutions, fused element-wise kernels); torch only carries tensors in and out.
"""
import copy

import torch
import torch.nn as nn

from .engine import HEAD_NAMES, NetEngine, unique_param_names

# (block name, index of its Conv2d in vgg19.features) — conv3_3 is constructed (and saved) but never run (:193-195)
_VGG_BLOCKS = [("conv1_1", 0), ("conv1_2", 2), ("conv2_1", 5), ("conv2_2", 7), ("conv3_1", 10), ("conv3_2", 12),
               ("conv3_3", 14), ("conv3_4", 16), ("conv4_1", 19), ("conv4_2", 21), ("conv4_3", 23), ("conv4_4", 25)]
_POOLS = [("pool1", 4), ("pool2", 9), ("pool3", 18)]
_HEAD_CH = {"det": 1, "loc": 4, "landmark": 4, "lmloc": 8}


class _NativeForward(torch.autograd.Function):
    """forward = engine.forward, backward = engine.backward; parameters are listed so autograd routes their grads.

    The engine keeps ONE set of activations per module: a backward pass is only valid for the most recent forward.
    Every forward stamps a generation on the engine; `backward` raises if another forward has overwritten the
    activations in between (e.g. `o1 = net(a); o2 = net(b); (L(o1) + L(o2)).backward()`), instead of silently
    returning the gradients of the wrong batch.  No gradient is produced for the input X (the reference's loops never
    ask for one): an X with requires_grad=True is rejected."""

    @staticmethod
    def forward(ctx, module, eng, x, *params):
        with torch.cuda.device(x.device):
            module._sync_params(eng)
            mode, seed, offset = 0, 0, 0
            if module.training:
                if module.dropout_mask is not None:  # parity tests inject the oracle's {0,2} masks
                    module._inject_dropout(eng)
                    mode = 2
                else:  # Philox inside the conv epilogues: stream = torch's seed + this module, counter = call number
                    mode = 1
                    seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * module._instance) & 0x7FFFFFFFFFFFFFFF
                    offset = module._drop_calls * eng.dropout_stride()
                    module._drop_calls += 1
            eng.forward(x.contiguous(), dropout_mode=mode, seed=seed, offset=offset)
            outs = module._outputs(eng)
        ctx.module, ctx.eng, ctx.generation = module, eng, eng.generation
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        module, eng = ctx.module, ctx.eng
        if ctx.generation != eng.generation:
            raise RuntimeError(
                "densebox_b200: backward() of a forward whose activations were overwritten by a later forward of the "
                "same module (one workspace per module: call backward before the next forward)")
        with torch.cuda.device(eng.device):
            eng.set_output_grads(dict(zip(module._out_keys, gouts)))
            if not eng._dgrad_fresh:
                eng.refresh_dgrad()
            eng.zero_grad()
            eng.backward()
            grads = eng.get_grads(module)
        return (None, None, None) + tuple(grads)


class _DenseBoxBase(nn.Module):
    variant = "densebox"
    _instances = 0

    def __init__(self, vgg19):
        super().__init__()
        _DenseBoxBase._instances += 1
        self._instance = _DenseBoxBase._instances  # separates the dropout streams of two modules under one torch seed
        self._drop_calls = 0
        self._param_epoch = 0                      # bumped by writers that bypass torch's version counters
        feats = vgg19.features._modules
        for name, idx in _VGG_BLOCKS:  # each block twice: `<name>_1`/`<name>_2` and the Sequential `<name>`
            conv = copy.deepcopy(feats[str(idx)])
            relu = copy.deepcopy(feats[str(idx + 1)])
            setattr(self, name + "_1", conv)
            setattr(self, name + "_2", relu)
            setattr(self, name, nn.Sequential(conv, relu))
        for name, idx in _POOLS:
            setattr(self, name, copy.deepcopy(feats[str(idx)]))
        self.dropout_mask = None  # optional {head: [N,512,h,w] 0/2 mask} for train-mode parity
        self._engines = {}

    def _make_head(self, head, wrapper):
        c1 = nn.Conv2d(768, 512, kernel_size=(1, 1))
        c2 = nn.Conv2d(512, _HEAD_CH[head], kernel_size=(1, 1))
        nn.init.xavier_normal_(c1.weight.data)
        nn.init.xavier_normal_(c2.weight.data)
        setattr(self, "conv5_1_" + head, c1)
        setattr(self, "conv5_2_" + head, c2)
        setattr(self, wrapper, nn.Sequential(c1, nn.Dropout(), c2))

    def _make_refine(self):
        self.conv6_1_det = nn.Conv2d(5, 64, kernel_size=(3, 3))
        self.conv6_2_det = nn.Conv2d(64, 64, kernel_size=(5, 5))
        self.conv6_3_det = nn.Conv2d(64, 1, kernel_size=(1, 1))
        for m in (self.conv6_1_det, self.conv6_2_det, self.conv6_3_det):
            nn.init.xavier_normal_(m.weight.data)

    # ---- engine plumbing
    def _wb(self, name):
        m = getattr(self, name + "_1") if name in dict(_VGG_BLOCKS) else getattr(self, name)
        return m.weight, m.bias

    def _engine(self, x):
        if not x.is_cuda:
            raise RuntimeError("densebox_b200: input must live on a CUDA device (no CPU fallback); "
                               "use oracle/ for CPU checks")
        train = bool(self.training or torch.is_grad_enabled())  # eval under no_grad: forward-only workspace
        key = (tuple(x.shape), train, x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            N, C, H, W = x.shape
            if C != 3 or H % 8 or W % 8:
                raise ValueError("DenseBox input must be [N,3,H,W] with H, W multiples of 8, got %s" % (tuple(x.shape),))
            self._engines.clear()  # one live workspace per module
            eng = NetEngine(self.variant, N, H, W, train=train, device=x.device)
            self._engines[key] = eng
        return eng

    def _sync_params(self, eng):
        """Re-pack the weights into the engine only when a parameter changed since the last forward (optimizer steps,
        load_state_dict and .data assignments all change the (data_ptr, version) key): ONE launch for all tensors."""
        key = [self._param_epoch]
        for name in unique_param_names(self.variant):
            w, b = self._wb(name)
            key += [w.data_ptr(), w._version, b.data_ptr(), b._version]
        key = tuple(key)
        if eng._param_key != key:
            eng.set_params(self)
            eng._param_key = key

    def _inject_dropout(self, eng):
        nh = len(HEAD_NAMES[eng.variant])
        drop = eng.buffer("drop", torch.bfloat16, (eng.N, eng.h4, eng.w4, 512 * nh))
        for i, h in enumerate(HEAD_NAMES[eng.variant]):
            m = self.dropout_mask[h].to(drop.device)
            drop[..., 512 * i:512 * (i + 1)] = m.permute(0, 2, 3, 1).to(torch.bfloat16)

    def _outputs(self, eng):
        m = eng.get_outputs(self._out_keys)  # fresh NCHW fp32 tensors, one launch
        return tuple(m[k] for k in self._out_keys)

    def forward(self, X):
        if X.requires_grad:
            raise ValueError("densebox_b200: no gradient is produced for the input X (DenseBox.py's loops never ask "
                             "for one); pass X.detach()")
        params = []
        for name in unique_param_names(self.variant):
            params += list(self._wb(name))
        return _NativeForward.apply(self, self._engine(X), X, *params)


class DenseBox(_DenseBoxBase):
    """Score + bbox heads (DenseBox.py:31-228): forward -> (scores [B,1,h,w], locs [B,4,h,w])."""
    variant = "densebox"
    _out_keys = ("score", "loc")

    def __init__(self, vgg19):
        super().__init__(vgg19)
        self._make_head("det", "output_score")
        self._make_head("loc", "output_loc")


class DenseBoxLM(_DenseBoxBase):
    """+ landmark heat-maps + refine branch (DenseBox.py:232-473): -> (scores, locs, landmarks, refine_scores)."""
    variant = "lm"
    _out_keys = ("score", "loc", "lm", "rf")

    def __init__(self, vgg19):
        super().__init__(vgg19)
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2, padding=0, dilation=1, ceil_mode=False)
        self._make_head("det", "output_score")
        self._make_head("loc", "output_loc")
        self._make_head("landmark", "output_landmark")
        self._make_refine()


class DenseBoxLMLOC(_DenseBoxBase):
    """+ landmark offsets (DenseBox.py:477-738): -> (score, rf_score, bbox_loc, lm_heatmap, lm_loc) (order of :738)."""
    variant = "lmloc"
    _out_keys = ("score", "rf", "loc", "lm", "lmloc")

    def __init__(self, vgg19):
        super().__init__(vgg19)
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2, padding=0, dilation=1, ceil_mode=False)
        self._make_head("det", "output_score")
        self._make_head("loc", "output_bbox_loc")
        self._make_head("lmloc", "output_lmloc")
        self._make_head("landmark", "output_lm_heatmap")
        self._make_refine()
