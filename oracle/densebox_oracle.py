"""ORACLE — CPU restatement of the DenseBox hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module, and only as the checker / the CPU baseline.  The product path (densebox_b200/) never imports it and has no
CPU fallback.

What is restated (file:line into /root/reference/DenseBox.py @ 7340ed0):
  forward()            DenseBox.forward :180-228, DenseBoxLM.forward :412-473, DenseBoxLMLOC.forward :674-738
                       (torch fp32 CPU functional ops; the reference's arithmetic lives in torch ATen — the pinned
                       third-party dependency is this image's torch 2.11.0 / numpy 2.3, the reference pins nothing)
  init_heads()         head construction order / RNG consumption of __init__ :143-178, :350-410, :595-672
  score_box()/gray_box()  init_score_map :1572-1582, mask_gray_zone_cls :1486-1504 (NumPy>=2 scalar semantics)
  gt maps              init_loc_map :1643-1653, init_lm_heatmap :1815-1823, init_lm_locmap :1705-1718, *_pn variants
  loss()               loop bodies train_online :2843-2918, train_LM_online :2575-2723, train_LMLOC_online :2300-2456,
                       train_densebox_online :2023-2180 (mask_by_sel :1368-1400, mask_gray_zone_lm :1435-1462,
                       gen_neg_loss :1917-1933)
  decode()/nms()       parse_out_MN :3114-3217 family, NMS :3398-3443

Pinning: the reference ships no tests or golden vectors.  This restatement is pinned against the
reference ITSELF, imported in the build container: tests/golden/make_golden.py runs the unmodified reference
functions/modules and stores their outputs as fixtures; tests/test_oracle_golden.py checks this file against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

MAP = 60  # training output map side (240 / 4), DenseBox.py:1379
f32 = np.float32

BACKBONE = [  # (name, vgg features index) — conv3_3 is constructed but never used in forward (:193-195)
    ("conv1_1", 0), ("conv1_2", 2), ("conv2_1", 5), ("conv2_2", 7), ("conv3_1", 10), ("conv3_2", 12),
    ("conv3_3", 14), ("conv3_4", 16), ("conv4_1", 19), ("conv4_2", 21), ("conv4_3", 23), ("conv4_4", 25),
]
HEADS = {  # variant -> list of (reference head name, out channels), in construction order
    "densebox": [("det", 1), ("loc", 4)],
    "lm": [("det", 1), ("loc", 4), ("landmark", 4)],
    "lmloc": [("det", 1), ("loc", 4), ("lmloc", 8), ("landmark", 4)],
}


# ------------------------------------------------------------------------------------------------ parameters
def params_from_vgg(vgg19, variant, seed_heads=None):
    """Build the unique-tensor parameter dict {name: tensor} the way the reference __init__ does.

    Backbone tensors are copies of vgg19.features[i]; heads follow the reference's construction order so that the
    same torch RNG state yields the same values (nn.Conv2d default init consumes RNG, then xavier_normal_)."""
    if seed_heads is not None:
        torch.manual_seed(seed_heads)
    P = {}
    feats = vgg19.features
    for name, idx in BACKBONE:
        P[name + ".weight"] = feats[idx].weight.detach().clone()
        P[name + ".bias"] = feats[idx].bias.detach().clone()
    for head, cout in HEADS[variant]:
        c1 = torch.nn.Conv2d(768, 512, 1)
        c2 = torch.nn.Conv2d(512, cout, 1)
        torch.nn.init.xavier_normal_(c1.weight.data)
        torch.nn.init.xavier_normal_(c2.weight.data)
        for nm, m in (("conv5_1_" + head, c1), ("conv5_2_" + head, c2)):
            P[nm + ".weight"] = m.weight.detach().clone()
            P[nm + ".bias"] = m.bias.detach().clone()
    if variant != "densebox":
        c61 = torch.nn.Conv2d(5, 64, 3)
        c62 = torch.nn.Conv2d(64, 64, 5)
        c63 = torch.nn.Conv2d(64, 1, 1)
        for m in (c61, c62, c63):
            torch.nn.init.xavier_normal_(m.weight.data)
        for nm, m in (("conv6_1_det", c61), ("conv6_2_det", c62), ("conv6_3_det", c63)):
            P[nm + ".weight"] = m.weight.detach().clone()
            P[nm + ".bias"] = m.bias.detach().clone()
    return P


def params_from_state_dict(sd, variant):
    """Pick the unique tensors out of a reference state_dict (every backbone/head tensor appears under two names)."""
    P = {}
    for name, _ in BACKBONE:
        P[name + ".weight"] = sd[name + "_1.weight"].detach().cpu().clone().float()
        P[name + ".bias"] = sd[name + "_1.bias"].detach().cpu().clone().float()
    names = ["conv5_1_%s" % h for h, _ in HEADS[variant]] + ["conv5_2_%s" % h for h, _ in HEADS[variant]]
    if variant != "densebox":
        names += ["conv6_1_det", "conv6_2_det", "conv6_3_det"]
    for nm in names:
        P[nm + ".weight"] = sd[nm + ".weight"].detach().cpu().clone().float()
        P[nm + ".bias"] = sd[nm + ".bias"].detach().cpu().clone().float()
    return P


# ------------------------------------------------------------------------------------------------ forward
class _StoreBF16(torch.autograd.Function):
    """Emulation of bf16 STORAGE of a tensor (not part of the reference): forward rounds the value (fwd=True),
    backward rounds the gradient.  Lets the tests separate "the CUDA backward is implemented correctly" (tight match
    against this emulation) from the inherent effect of bf16 activations on ReLU masks / pool arg-maxes."""

    @staticmethod
    def forward(ctx, x, fwd):
        return x.bfloat16().float() if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float(), None


def _cr(x, P, name, st, gb=None, frozen=None):
    z = F.conv2d(x, P[name + ".weight"], P[name + ".bias"], padding=1)
    if frozen is None:
        return st(F.relu(z))
    # frozen decisions: the ReLU mask is the engine's (its stored activation > 0), the value is the engine's, the
    # gradient path is this graph's; the engine stores d(loss)/dz in bf16 -> round the gradient of z
    y = gb(z) * (frozen[name] > 0).to(z.dtype)
    return _subst(y, frozen[name])


def _subst(t, value):
    """Tensor with the VALUE `value` (what the engine stored) and the gradient path of `t`."""
    return value + (t - t.detach())


def _head(fusion, P, head, drop, st, frozen=None, record=None):
    h = F.conv2d(fusion, P["conv5_1_%s.weight" % head], P["conv5_1_%s.bias" % head])
    if drop is not None:  # train mode: nn.Dropout(p=0.5) == multiply by a {0,2} mask (injected for parity)
        h = h * drop
    h = st(h)
    if frozen is not None:
        h = _subst(h, frozen["hd_" + head])
    if record is not None:
        record["hd_" + head] = h.detach().clone()
    return F.conv2d(h, P["conv5_2_%s.weight" % head], P["conv5_2_%s.bias" % head])


def forward(P, X, variant="densebox", dropout=None, return_intermediates=False, emulate_bf16_storage=False,
            frozen=None, record=None):
    """Reference forward.  dropout: None (eval) or {head: mask[B,512,h,w] of 0/2} (train).
    emulate_bf16_storage=True additionally rounds every stored activation (and its gradient) to bf16 at the points
    where the CUDA engine stores bf16 — a test aid, see _StoreBF16; the reference itself is fp32 throughout.
    frozen = {layer: tensor} (a test aid, implies the bf16 emulation): FROZEN-DECISION mode.  Every stored activation
    takes the value the CUDA engine stored (conv*/up/hd_*/score/loc/lm/lmloc/rp/r1/r2/rup/rf, NCHW fp32 copies of the
    engine's buffers) and every ReLU mask is the engine's, so all data-dependent decisions of the backward pass (ReLU
    masks, max-pool arg-maxes — F.max_pool2d on identical values picks identical positions —, mined negatives) are
    the engine's own and the comparison of parameter gradients is free of the bf16 ReLU/arg-max flip noise: what
    remains is accumulation order and bf16 rounding of the stored gradients.
    record = {} collects the detached value of every such stage under the same names (self-test of the frozen mode)."""
    if emulate_bf16_storage or frozen is not None:
        st = lambda t: _StoreBF16.apply(t, True)      # value and gradient stored as bf16
        gb = lambda t: _StoreBF16.apply(t, False)     # only the gradient is stored as bf16
    else:
        st = gb = lambda t: t
    fz = frozen

    def sub(t, name):
        t = _subst(t, fz[name]) if fz is not None else t
        if record is not None:
            record[name] = t.detach().clone()
        return t

    def cr(x, name):
        y = _cr(x, P, name, st, gb, fz)
        if record is not None:
            record[name] = y.detach().clone()
        return y

    if record is not None:
        _rec_heads = record
    else:
        _rec_heads = None
    x = cr(X, "conv1_1"); x = cr(x, "conv1_2"); x = gb(F.max_pool2d(x, 2, 2))
    x = cr(x, "conv2_1"); x = cr(x, "conv2_2"); x = gb(F.max_pool2d(x, 2, 2))
    x = cr(x, "conv3_1"); x = cr(x, "conv3_2")
    c34 = cr(x, "conv3_4")  # conv3_3 skipped (:193-195)
    x = gb(F.max_pool2d(c34, 2, 2))
    x = cr(x, "conv4_1"); x = cr(x, "conv4_2"); x = cr(x, "conv4_3")
    c44 = cr(x, "conv4_4")
    up = sub(st(F.interpolate(c44, size=(c34.shape[2], c34.shape[3]), mode="bilinear", align_corners=True)), "up")
    fusion = gb(torch.cat((up, c34), dim=1))  # upsampled conv4_4 first (:219)
    d = dropout or {}
    score = sub(gb(_head(fusion, P, "det", d.get("det"), st, fz, _rec_heads)), "score")
    loc = sub(gb(_head(fusion, P, "loc", d.get("loc"), st, fz, _rec_heads)), "loc")
    inter = {"conv3_4": c34, "conv4_4": c44, "fusion": fusion}
    if variant == "densebox":
        out = (score, loc)
    else:
        lm = sub(gb(_head(fusion, P, "landmark", d.get("landmark"), st, fz, _rec_heads)), "lm")
        lmloc = sub(gb(_head(fusion, P, "lmloc", d.get("lmloc"), st, fz, _rec_heads)), "lmloc") if variant == "lmloc" else None
        x = torch.cat((lm, score), dim=1)
        x = sub(st(F.max_pool2d(x, 2, 2)), "rp")
        x = sub(st(F.conv2d(x, P["conv6_1_det.weight"], P["conv6_1_det.bias"])), "r1")
        x = sub(st(F.conv2d(x, P["conv6_2_det.weight"], P["conv6_2_det.bias"])), "r2")
        x = sub(st(F.interpolate(x, size=(score.shape[2], score.shape[3]), mode="bilinear", align_corners=True)), "rup")
        rf = sub(gb(F.conv2d(x, P["conv6_3_det.weight"], P["conv6_3_det.bias"])), "rf")
        out = (score, loc, lm, rf) if variant == "lm" else (score, rf, loc, lm, lmloc)  # return orders :473, :738
    return (out, inter) if return_intermediates else out


# ------------------------------------------------------------------------------------------------ label geometry
def _pyslice(a, b, n):
    """Python slice a:b on an axis of length n -> [lo, hi) (empty when hi <= lo)."""
    lo, hi, _ = slice(a, b).indices(n)
    return lo, max(hi, lo)


def score_box(coord, ratio=0.3):
    """init_score_map :1572-1582 with NumPy>=2 scalar semantics: coord is float32, `ratio * w` is a float32 product,
    the rest is Python double.  Returns inclusive-start/exclusive-end python-slice bounds (x0, x1, y0, y1)."""
    c = np.asarray(coord, dtype=f32)
    cx = float(c[0] + c[2]) * 0.5
    cy = float(c[1] + c[3]) * 0.5
    w = f32(c[2] - c[0]); h = f32(c[3] - c[1])
    r = f32(ratio)
    ox = int(cx - float(f32(f32(r * w) * f32(0.5))) + 0.5)
    oy = int(cy - float(f32(f32(r * h) * f32(0.5))) + 0.5)
    ex = int(float(ox) + float(f32(r * w)) + 0.5)
    ey = int(float(oy) + float(f32(r * h)) + 0.5)
    return ox, ex + 1, oy, ey + 1


def gray_box(coord, ratio=0.3, border=2.0):
    """mask_gray_zone_cls :1486-1504: zero [gy:Gy, gx:Gx] then set [gy+2:Gy-2+1, gx+2:Gx-2+1] back to one."""
    c = np.asarray(coord, dtype=f32)
    cx = float(c[0] + c[2]) * 0.5
    cy = float(c[1] + c[3]) * 0.5
    w = f32(c[2] - c[0]); h = f32(c[3] - c[1])
    r = f32(ratio)
    gx = int(cx - float(f32(f32(r * w) * f32(0.5))) - border + 0.5)
    gy = int(cy - float(f32(f32(r * h) * f32(0.5))) - border + 0.5)
    Gx = int(float(gx) + float(f32(r * w)) + border * 2.0 + 0.5)
    Gy = int(float(gy) + float(f32(r * h)) + border * 2.0 + 0.5)
    b = int(border)
    return (gx, Gx, gy, Gy), (gx + b, Gx - b + 1, gy + b, Gy - b + 1)


def lm_points(vert, clamp=False):
    """init_lm_heatmap :1815-1823 (clamp=True: the last init_lm_heatmap_pn definition :1899-1907)."""
    v = np.asarray(vert, dtype=f32)
    pts = []
    for k in range(4):
        x = int(f32(v[2 * k] + f32(0.5)))
        y = int(f32(v[2 * k + 1] + f32(0.5)))
        if clamp:
            x = x if x < MAP else MAP - 1
            y = y if y < MAP else MAP - 1
        pts.append((x, y))
    return pts


def gt_maps(bbox, vertices=None, labels=None):
    """All ground-truth maps as float32 numpy: score [B,1,60,60], loc [B,4,...], lm [B,4,...], lmloc [B,8,...]."""
    bbox = np.asarray(bbox, dtype=f32)
    B = bbox.shape[0]
    lab = np.ones(B, dtype=f32) if labels is None else np.asarray(labels, dtype=f32).reshape(B)
    ys, xs = np.meshgrid(np.arange(MAP, dtype=f32), np.arange(MAP, dtype=f32), indexing="ij")
    score = np.zeros((B, 1, MAP, MAP), f32)
    loc = np.zeros((B, 4, MAP, MAP), f32)
    for b in range(B):
        if lab[b] == 0.0:
            continue
        x0, x1, y0, y1 = score_box(bbox[b])
        ya, yb = _pyslice(y0, y1, MAP); xa, xb = _pyslice(x0, x1, MAP)
        score[b, 0, ya:yb, xa:xb] = 1.0
        loc[b, 0] = xs - bbox[b, 0]; loc[b, 1] = ys - bbox[b, 1]
        loc[b, 2] = xs - bbox[b, 2]; loc[b, 3] = ys - bbox[b, 3]
    out = {"score": score, "loc": loc}
    if vertices is not None:
        vertices = np.asarray(vertices, dtype=f32)
        lm = np.zeros((B, 4, MAP, MAP), f32)
        lmloc = np.zeros((B, 8, MAP, MAP), f32)
        for b in range(B):
            if lab[b] == 0.0:
                continue
            for k, (x, y) in enumerate(lm_points(vertices[b], clamp=labels is not None)):
                lm[b, k, y, x] = 1.0
            for k in range(4):
                lmloc[b, 2 * k] = xs - vertices[b, 2 * k]
                lmloc[b, 2 * k + 1] = ys - vertices[b, 2 * k + 1]
        out["lm"] = lm
        out["lmloc"] = lmloc
    return out


def half_neg_num(score_gt):
    """:2864-2876 — quota from the batch-global positive count."""
    B = score_gt.shape[0]
    pos = int(np.count_nonzero(score_gt))
    neg = int(float(pos) / float(B) + 0.5)
    return int(neg * 0.5 + 0.5), pos


def _set_sel(mask_b, idxs):
    for idx in idxs:  # mask_by_sel :1383-1398
        idx = int(idx)
        if idx < 0 or idx >= MAP * MAP:
            continue
        mask_b[idx // MAP, idx % MAP] = 1.0


def cls_mask(score_np, gt, bbox, rand_idx, half, labels=None):
    """Loss mask of the classification branch: positives + hard negatives (top-k of (s-gt)^2*(1-gt)) + injected random
    negatives, then the gray zone (:2864-2909).  score_np: [B,1,60,60] network output."""
    B = gt.shape[0]
    neg_loss = ((score_np - gt) ** 2 * (1.0 - gt)).reshape(B, -1).astype(f32)
    mask = gt.copy()
    hard = np.zeros((B, half), np.int64)
    for b in range(B):
        # torch.topk: largest k; ties are implementation-defined — fixtures avoid exact ties (index-ascending here)
        order = np.lexsort((np.arange(neg_loss.shape[1]), -neg_loss[b]))
        hard[b] = order[:half]
        _set_sel(mask[b, 0], hard[b])
        _set_sel(mask[b, 0], rand_idx[b][:half])
    lab = np.ones(B, f32) if labels is None else np.asarray(labels, f32).reshape(B)
    for b in range(B):
        if lab[b] == 0.0:
            continue
        (gx, Gx, gy, Gy), (ix, Ix, iy, Iy) = gray_box(bbox[b])
        ya, yb = _pyslice(gy, Gy, MAP); xa, xb = _pyslice(gx, Gx, MAP)
        mask[b, 0, ya:yb, xa:xb] = 0.0
        ya, yb = _pyslice(iy, Iy, MAP); xa, xb = _pyslice(ix, Ix, MAP)
        mask[b, 0, ya:yb, xa:xb] = 1.0
    return mask, hard


def lm_mask(lm_np, lm_gt, lm_rand_idx):
    """Landmark loss mask (:2660-2701): per landmark map, positives + top-1 hard negative + 1 random negative, then a
    5x5 ignore zone around every positive with the centre restored (mask_gray_zone_lm :1453-1462)."""
    B = lm_gt.shape[0]
    mask = lm_gt.copy()
    for k in range(4):
        neg = ((lm_np[:, k] - lm_gt[:, k]) ** 2 * (1.0 - lm_gt[:, k])).reshape(B, -1).astype(f32)
        for b in range(B):
            order = np.lexsort((np.arange(neg.shape[1]), -neg[b]))
            _set_sel(mask[b, k], order[:1])
            _set_sel(mask[b, k], [lm_rand_idx[b][k]])
        for b, y, x in zip(*np.nonzero(lm_gt[:, k])):
            ya, yb = _pyslice(y - 2, y + 3, MAP); xa, xb = _pyslice(x - 2, x + 3, MAP)
            mask[b, k, ya:yb, xa:xb] = 0.0
            mask[b, k, y, x] = 1.0
    return mask


def loss(outs, variant, bbox, rand_idx, vertices=None, lm_rand_idx=None, labels=None, lambda_loc=3.0, lambda_det=1.0,
         lambda_lm=0.5, global_pos=None, global_batch=None):
    """Multi-task loss of the reference loop bodies.  outs: tuple of torch tensors in the variant's forward order.
    Returns (loss tensor with autograd, info dict with masks / quotas).  rand_idx [B,>=half] are the injected
    np.random.choice draws (:2888-2893); lm_rand_idx [B,4] the per-landmark draws (:2676-2683)."""
    if variant == "densebox":
        score, loc = outs; lm = rf = lmloc = None
    elif variant == "lm":
        score, loc, lm, rf = outs; lmloc = None
    else:
        score, rf, loc, lm, lmloc = outs
    gts = gt_maps(bbox, vertices, labels)
    gt = gts["score"]
    B = gt.shape[0]
    if global_pos is None:
        half, pos = half_neg_num(gt)
    else:  # data-parallel shard: quota from the global batch
        pos = global_pos
        half = int(int(float(pos) / float(global_batch) + 0.5) * 0.5 + 0.5)
    s_np = score.detach().cpu().numpy().astype(f32)
    mask, hard = cls_mask(s_np, gt, np.asarray(bbox, f32), np.asarray(rand_idx), half, labels)
    t = lambda a: torch.from_numpy(a)
    m, g = t(mask), t(gt)
    Ls = (score - g) ** 2
    Lloc = (loc - t(gts["loc"])) ** 2
    cls_sum = torch.sum(m * Ls)
    loc_sum = torch.sum(m * g * Lloc)
    info = {"half": half, "pos": pos, "mask": mask, "hard": hard, "cls_sum": float(cls_sum.detach()),
            "loc_sum": float(loc_sum.detach())}
    if variant == "densebox":
        total = cls_sum + torch.sum(lambda_loc * (m * g * Lloc))  # :2917-2918
        return total, info
    mlm = lm_mask(lm.detach().cpu().numpy().astype(f32), gts["lm"], np.asarray(lm_rand_idx))
    info["lm_mask"] = mlm
    det = lambda_det * (cls_sum + lambda_loc * loc_sum)
    lm_loss = lambda_lm * torch.sum(t(mlm) * (lm - t(gts["lm"])) ** 2)
    if variant == "lmloc":
        lm_loss = lm_loss + torch.sum(m * g * (lmloc - t(gts["lmloc"])) ** 2)
    rf_loss = torch.sum(m * (rf - g) ** 2)
    return det + lm_loss + rf_loss, info


# ------------------------------------------------------------------------------------------------ decode + NMS
def decode(score_map, loc_map, lmloc_map=None, K=10, lmheat_map=None):
    """parse_out_MN / parse_DetLMLOC (:3114-3217): top-K of the raw score map, boxes (and landmarks) decoded x4;
    lmheat_map given: parse_DetLM (:3220-3300) — the landmarks of every row are the arg-max positions of the four
    landmark heat-maps x4.  Maps are torch [1,C,h,w]; returns float64 numpy [K, 5 or 13]."""
    h, w = score_map.shape[2], score_map.shape[3]
    heat_lms = []
    if lmheat_map is not None:
        for k in range(4):
            _, li = torch.topk(lmheat_map[0, k].reshape(-1), 1)  # :3285
            li = int(li)
            heat_lms += [float(li % w) * 4.0, float(li // w) * 4.0]
    s = score_map.reshape(-1)
    vals, idx = torch.topk(s, K)
    dets = []
    for v, i in zip(vals.tolist(), idx.tolist()):
        xi, yi = i % w, i // w
        # `xi - map[c, idx]` is a float32 tensor op in the reference; float(...) * 4.0 then runs in Python double
        sub = lambda a, m, c: float(f32(a) - f32(float(m[0, c, yi, xi])))
        row = [sub(xi, loc_map, 0) * 4.0, sub(yi, loc_map, 1) * 4.0, sub(xi, loc_map, 2) * 4.0,
               sub(yi, loc_map, 3) * 4.0, v]
        if lmloc_map is not None:
            for k in range(4):
                row += [sub(xi, lmloc_map, 2 * k) * 4.0, sub(yi, lmloc_map, 2 * k + 1) * 4.0]
        row += heat_lms
        dets.append(row)
    return np.asarray(dets, dtype=np.float64)


def nms(dets, thresh):
    """NMS :3398-3443 — greedy, areas with +1, keep while IoU <= thresh; returns kept indices into dets."""
    x1, y1, x2, y2, sc = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = sc.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        xx1 = np.maximum(x1[i], x1[order[1:]]); yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]]); yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1); h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        inds = np.where(ovr <= thresh)[0]
        order = order[inds + 1]
    return keep


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_batch(B, seed=0, with_vertices=False):
    """SURVEY §8(d) synthetic labels: interior boxes (240-space ints / 4) and corner landmarks with +-2 px jitter."""
    rs = np.random.RandomState(seed)
    x0 = rs.randint(40, 101, B); y0 = rs.randint(60, 121, B)
    w = rs.randint(40, 81, B); h = rs.randint(16, 33, B)
    bbox = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(f32) / f32(4.0)
    out = {"bbox": bbox}
    if with_vertices:
        corners = np.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1).astype(np.int64)
        corners = corners + rs.randint(-2, 3, corners.shape)
        out["vertices"] = corners.astype(f32) / f32(4.0)
    return out


def seeded_vgg19(seed=0):
    import torchvision
    torch.manual_seed(seed)
    return torchvision.models.vgg19(weights=None)
