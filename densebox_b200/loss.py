"""`densebox_loss` — the loss the reference computes inline in its training loops, as one differentiable op.

Reference loop bodies: train_online DenseBox.py:2843-2918 (score+bbox), train_LM_online :2575-2723 (+landmark
heat-maps, refine), train_LMLOC_online :2300-2456 (+landmark offsets), train_densebox_online :2023-2180 (`labels`
given: pos/neg patches).  One CUDA kernel does GT synthesis, hard-negative mining, masks, the masked sums and the
gradients (densebox_b200/csrc/dbx_loss.cu).
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream_ptr

c_int, c_float = ctypes.c_int, ctypes.c_float
_SLICES = {"score": (0, 1), "loc": (1, 5), "lm": (5, 9), "lmloc": (9, 17)}


def _dev(t, dtype, device):
    if t is None:
        return None
    return torch.as_tensor(t).to(device=device, dtype=dtype).contiguous()


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, score, loc, lm, rf, lmloc):
        dev = score.device
        B = score.shape[0]
        if tuple(score.shape[1:]) != (1, 60, 60):
            raise ValueError("densebox_loss is defined on 60x60 maps (DenseBox.py:1379), got %s" % (tuple(score.shape),))
        variant = 0 if lm is None else (1 if lmloc is None else 2)
        HC = 32 if variant == 2 else 16
        head = torch.zeros(B, 60, 60, HC, device=dev, dtype=torch.float32)
        for key, t in (("score", score), ("loc", loc), ("lm", lm), ("lmloc", lmloc)):
            if t is not None:
                a, b = _SLICES[key]
                head[..., a:b] = t.detach().float().permute(0, 2, 3, 1)
        rfp = None
        if variant >= 1:
            rfp = torch.zeros(B, 60, 60, 16, device=dev, dtype=torch.float32)
            rfp[..., 0:1] = rf.detach().float().permute(0, 2, 3, 1)
        scratch = torch.zeros(16 + 4 * B + 16, dtype=torch.uint8, device=dev)
        out = torch.zeros(4, device=dev, dtype=torch.float32)
        info = torch.zeros(2, device=dev, dtype=torch.int32)
        d_head = torch.empty_like(head)
        d_rf = torch.empty_like(rfp) if rfp is not None else None
        mask = torch.empty(B, 3600, dtype=torch.uint8, device=dev) if cfg["want_masks"] else None
        lmmask = torch.empty(B, 4, 3600, dtype=torch.uint8, device=dev) if (cfg["want_masks"] and variant) else None
        rand = cfg["rand"]
        check(lib().dbx_loss_fwd_bwd(
            ptr(head), c_int(HC), ptr(rfp), c_int(16), ptr(cfg["bbox"]), ptr(cfg["vertices"]), ptr(cfg["labels"]),
            ptr(rand), c_int(rand.shape[1]), ptr(cfg["lm_rand"]), c_int(variant), c_float(cfg["lambda_loc"]),
            c_float(cfg["lambda_det"]), c_float(cfg["lambda_lm"]), c_int(cfg["global_pos"]),
            c_int(cfg["global_batch"]), ptr(None), c_int(int(cfg["labels"] is not None)), c_int(B), ptr(scratch),
            ptr(out), ptr(info), ptr(None), ptr(None), ptr(d_head), ptr(d_rf), ptr(mask), ptr(lmmask), stream_ptr()),
            "loss_fwd_bwd")
        ctx.d_head, ctx.d_rf, ctx.variant = d_head, d_rf, variant
        cfg["info"] = info
        cfg["mask"], cfg["lm_mask"] = mask, lmmask
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        dh, dr = ctx.d_head, ctx.d_rf
        pick = lambda key: (dh[..., _SLICES[key][0]:_SLICES[key][1]].permute(0, 3, 1, 2) * g).contiguous()
        gs = pick("score")
        gl = pick("loc")
        glm = pick("lm") if ctx.variant >= 1 else None
        grf = (dr[..., 0:1].permute(0, 3, 1, 2) * g).contiguous() if ctx.variant >= 1 else None
        glmloc = pick("lmloc") if ctx.variant == 2 else None
        return None, gs, gl, glm, grf, glmloc


def densebox_loss(score, loc, bbox, *, lm=None, rf=None, lm_loc=None, vertices=None, labels=None, rand_neg_idx=None,
                  lm_rand_neg_idx=None, lambda_loc=3.0, lambda_det=1.0, lambda_lm=0.5, global_pos_count=None,
                  global_batch=None, return_info=False):
    """Multi-task DenseBox loss (a 0-dim fp32 tensor with .backward()).

    score [B,1,60,60], loc [B,4,60,60]; optional lm [B,4,..], rf [B,1,..], lm_loc [B,8,..] select the LM / LMLOC
    loss.  bbox [B,4] / vertices [B,8] are 60-space floats (pixel / 4).  rand_neg_idx [B,>=half] int64 are the
    np.random.choice draws of the reference (None: drawn here with torch, non-parity mode); lm_rand_neg_idx [B,4].
    global_pos_count / global_batch: batch-global values for data-parallel shards.
    """
    dev = score.device
    if not score.is_cuda:
        raise RuntimeError("densebox_loss runs on CUDA tensors only (no CPU fallback)")
    B = score.shape[0]
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
        half, pos = [int(v) for v in cfg["info"].tolist()]
        return out, {"half": half, "pos": pos, "mask": cfg["mask"], "lm_mask": cfg["lm_mask"]}
    return out
