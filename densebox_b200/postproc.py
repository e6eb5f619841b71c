"""Detection post-processing on the GPU: `decode_nms` = parse_out_MN / parse_DetLMLOC / parse_DetLM + NMS of the
reference (DenseBox.py:3114-3217, :3220-3300, :3398-3443) for a whole batch in one kernel launch (the reference
handles one image at a time on the CPU)."""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

c_int, c_long = ctypes.c_int, ctypes.c_long


def _strides(t):
    """(image, pixel, channel) element strides of an NCHW-shaped tensor view [N,C,h,w] whose (h,w) plane is dense."""
    n, c, h, w = t.stride()
    assert h == t.shape[3] * w, "the (h,w) plane must be addressable with one pixel stride"
    return n, w, c


def decode_nms(score_map, loc_map, lm_loc_map=None, K=10, nms_thresh=0.4, lm_heat_map=None):
    """score_map [N,1,h,w], loc_map [N,4,h,w], lm_loc_map [N,8,h,w] or None (fp32 CUDA tensors, any strides with a
    dense pixel plane, e.g. the tensors returned by forward).  lm_heat_map [N,4,h,w] instead of lm_loc_map selects
    parse_DetLM (:3220-3300): the landmarks are the arg-max positions of the four heat-maps.  Returns a list (one
    entry per image) of float64 numpy arrays [kept, 5 or 13] = rows (xt, yt, xb, yb, score[, x0, y0, ..., x3, y3])
    that survive NMS."""
    if not score_map.is_cuda:
        raise RuntimeError("decode_nms runs on CUDA tensors only (no CPU fallback)")
    if lm_heat_map is not None and lm_loc_map is not None:
        raise ValueError("give either lm_loc_map (parse_DetLMLOC) or lm_heat_map (parse_DetLM)")
    heat = lm_heat_map is not None
    if heat:
        lm_loc_map = lm_heat_map
    N, _, h, w = score_map.shape
    for t in (score_map, loc_map, lm_loc_map):
        assert t is None or (t.dtype == torch.float32 and t.is_cuda)
    dets = torch.empty(N, K, 13, dtype=torch.float32, device=score_map.device)
    keep = torch.empty(N, K, dtype=torch.int32, device=score_map.device)
    s_img, s_pix, _ = _strides(score_map)
    l = _strides(loc_map)
    m = _strides(lm_loc_map) if lm_loc_map is not None else (0, 0, 0)
    fn = lib().dbx_decode_nms_heat if heat else lib().dbx_decode_nms
    check(fn(ptr(score_map), c_long(s_img), c_long(s_pix), ptr(loc_map), c_long(l[0]), c_long(l[1]),
             c_long(l[2]), ptr(lm_loc_map), c_long(m[0]), c_long(m[1]), c_long(m[2]), c_int(N),
             c_int(h), c_int(w), c_int(K), ctypes.c_double(nms_thresh), ptr(dets), ptr(keep),
             stream_ptr()), "decode_nms")
    dets, keep = dets.cpu().numpy().astype(np.float64), keep.cpu().numpy().astype(bool)
    ncol = 13 if lm_loc_map is not None else 5
    return [dets[i][keep[i]][:, :ncol] for i in range(N)]
