"""`densebox_loss` — the loss the reference computes inline in its training loops, as one differentiable op.

Reference loop bodies: train_online DenseBox.py:2843-2918 (score+bbox), train_LM_online :2575-2723 (+landmark
heat-maps, refine), train_LMLOC_online :2300-2456 (+landmark offsets), train_densebox_online :2023-2180 (`labels`
given: pos/neg patches).  One CUDA kernel does GT synthesis, hard-negative mining, masks, the masked sums and the
gradients (densebox_b200/csrc/dbx_loss.cu); it reads the NCHW maps returned by forward() in place (dbx_loss_maps) and
writes their gradients in the same layout — no repacking on either side.
"""
import ctypes

import torch

from ._lib import DbxError, check, lib, ptr, stream_ptr

c_int, c_float = ctypes.c_int, ctypes.c_float
_CH = (1, 4, 4, 8, 1)  # score, loc, landmark heat, landmark loc, refine
_scratch = {}          # device index -> zeroed scratch of the loss kernel (its counter resets itself)


def _dev(t, dtype, device):
    if t is None:
        return None
    return torch.as_tensor(t).to(device=device, dtype=dtype).contiguous()


def _scratch_for(dev, B):
    t = _scratch.get(dev.index)
    if t is None or t.numel() < 32 + 4 * B:
        t = _scratch[dev.index] = torch.zeros(32 + 4 * max(B, 256), dtype=torch.uint8, device=dev)
    return t


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, score, loc, lm, rf, lmloc):
        dev = score.device
        B = score.shape[0]
        if tuple(score.shape[1:]) != (1, 60, 60):
            raise ValueError("densebox_loss is defined on 60x60 maps (DenseBox.py:1379), got %s" % (tuple(score.shape),))
        variant = 0 if lm is None else (1 if lmloc is None else 2)
        maps = [score, loc, lm, lmloc, rf]
        for t, c in zip(maps, _CH):
            if t is not None and tuple(t.shape) != (B, c, 60, 60):
                raise ValueError("densebox_loss: expected a [%d,%d,60,60] map, got %s" % (B, c, tuple(t.shape)))
        maps = [t.detach().float().contiguous() if t is not None else None for t in maps]
        grads = [torch.empty_like(t) if t is not None else None for t in maps]
        with torch.cuda.device(dev):
            out = torch.empty(1, device=dev, dtype=torch.float32)
            info = torch.empty(4, device=dev, dtype=torch.int32)
            mask = torch.empty(B, 3600, dtype=torch.uint8, device=dev) if cfg["want_masks"] else None
            lmmask = torch.empty(B, 4, 3600, dtype=torch.uint8, device=dev) if (cfg["want_masks"] and variant) else None
            mp = (ctypes.c_void_p * 5)(*[t.data_ptr() if t is not None else None for t in maps])
            gp = (ctypes.c_void_p * 5)(*[t.data_ptr() if t is not None else None for t in grads])
            st = (ctypes.c_long * 15)()
            for g, c in enumerate(_CH):
                st[3 * g:3 * g + 3] = [c * 3600, 1, 3600]  # (image, pixel, channel) element strides of NCHW
            rand = cfg["rand"]
            check(lib().dbx_loss_maps(
                mp, st, gp, ptr(cfg["bbox"]), ptr(cfg["vertices"]), ptr(cfg["labels"]), ptr(rand),
                c_int(rand.shape[1]), ptr(cfg["lm_rand"]), c_int(variant), c_float(cfg["lambda_loc"]),
                c_float(cfg["lambda_det"]), c_float(cfg["lambda_lm"]), c_int(cfg["global_pos"]),
                c_int(cfg["global_batch"]), ptr(None), c_int(int(cfg["labels"] is not None)), c_int(B),
                ptr(_scratch_for(dev, B)), ptr(out), ptr(info), ptr(mask), ptr(lmmask), stream_ptr()), "loss_maps")
        ctx.grads = grads
        cfg["info"] = info
        cfg["mask"], cfg["lm_mask"] = mask, lmmask
        return out[0]

    @staticmethod
    def backward(ctx, g):
        gs, gl, glm, glmloc, grf = [(t * g if t is not None else None) for t in ctx.grads]
        return None, gs, gl, glm, grf, glmloc


def densebox_loss(score, loc, bbox, *, lm=None, rf=None, lm_loc=None, vertices=None, labels=None, rand_neg_idx=None,
                  lm_rand_neg_idx=None, lambda_loc=3.0, lambda_det=1.0, lambda_lm=0.5, global_pos_count=None,
                  global_batch=None, return_info=False):
    """Multi-task DenseBox loss (a 0-dim fp32 tensor with .backward()).

    score [B,1,60,60], loc [B,4,60,60]; optional lm [B,4,..], rf [B,1,..], lm_loc [B,8,..] select the LM / LMLOC
    loss.  bbox [B,4] / vertices [B,8] are 60-space floats (pixel / 4).  rand_neg_idx [B,>=half] int64 are the
    np.random.choice draws of the reference (None: drawn here with torch, non-parity mode); lm_rand_neg_idx [B,4].
    global_pos_count / global_batch: batch-global values for data-parallel shards.
    return_info=True also returns {half, pos, mask, lm_mask} (synchronises) and raises if rand_neg_idx holds fewer
    columns than the negative quota `half` (the reference draws exactly `half` per sample, DenseBox.py:2888-2893).
    """
    dev = score.device
    if not score.is_cuda:
        raise RuntimeError("densebox_loss runs on CUDA tensors only (no CPU fallback)")
    B = score.shape[0]
    with torch.cuda.device(dev):
        if rand_neg_idx is None:
            rand_neg_idx = torch.rand(B, 3600, device=dev).argsort(dim=1)
        if lm is not None and lm_rand_neg_idx is None:
            lm_rand_neg_idx = torch.randint(0, 3600, (B, 4), device=dev)
    if (lm is None) != (rf is None) or (lm is not None and vertices is None):
        raise ValueError("the LM / LMLOC loss needs lm, rf and vertices together")
    cfg = {
        "bbox": _dev(bbox, torch.float32, dev), "vertices": _dev(vertices, torch.float32, dev),
        "labels": _dev(labels, torch.float32, dev).reshape(-1) if labels is not None else None,
        "rand": _dev(rand_neg_idx, torch.int64, dev), "lm_rand": _dev(lm_rand_neg_idx, torch.int64, dev),
        "lambda_loc": float(lambda_loc), "lambda_det": float(lambda_det), "lambda_lm": float(lambda_lm),
        "global_pos": -1 if global_pos_count is None else int(global_pos_count),
        "global_batch": -1 if global_batch is None else int(global_batch), "want_masks": bool(return_info),
    }
    if tuple(cfg["bbox"].shape) != (B, 4):
        raise ValueError("bbox must be [B,4]")
    out = _Loss.apply(cfg, score, loc, lm, rf, lm_loc)
    if return_info:
        half, pos, short, _ = [int(v) for v in cfg["info"].tolist()]
        if short:
            raise DbxError("densebox_loss: rand_neg_idx has %d columns but the negative quota is half=%d"
                           % (cfg["rand"].shape[1], half))
        return out, {"half": half, "pos": pos, "mask": cfg["mask"], "lm_mask": cfg["lm_mask"]}
    return out
