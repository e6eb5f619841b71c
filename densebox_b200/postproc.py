"""Detection post-processing on the GPU: `decode_nms` = parse_out_MN / parse_DetLMLOC / parse_DetLM + NMS of the
reference (DenseBox.py:3114-3217, :3220-3300, :3398-3443) for a whole batch in one kernel launch (the reference
handles one image at a time on the CPU)."""
import ctypes
import threading

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
    # one device buffer, one pinned host buffer, ONE device->host copy: rows [N, K, 13] fp32 followed by keep [N, K] int32
    nd = N * K * 13
    out = torch.empty(nd + N * K, dtype=torch.float32, device=score_map.device)
    dets, keep = out[:nd], out[nd:].view(torch.int32)
    s_img, s_pix, _ = _strides(score_map)
    l = _strides(loc_map)
    m = _strides(lm_loc_map) if lm_loc_map is not None else (0, 0, 0)
    fn = lib().dbx_decode_nms_heat if heat else lib().dbx_decode_nms
    with torch.cuda.device(score_map.device):
        check(fn(ptr(score_map), c_long(s_img), c_long(s_pix), ptr(loc_map), c_long(l[0]), c_long(l[1]),
                 c_long(l[2]), ptr(lm_loc_map), c_long(m[0]), c_long(m[1]), c_long(m[2]), c_int(N),
                 c_int(h), c_int(w), c_int(K), ctypes.c_double(nms_thresh), ptr(dets), ptr(keep),
                 stream_ptr()), "decode_nms")
    host = _pinned(out.numel(), score_map.device.index)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(score_map.device).synchronize()
    hv = host.numpy()
    ncol = 13 if lm_loc_map is not None else 5
    rows = hv[:nd].reshape(N, K, 13)[:, :, :ncol].astype(np.float64)
    alive = hv[nd:].view(np.int32).reshape(N, K) != 0
    return [rows[i][alive[i]] for i in range(N)]


_PINNED = {}


def _pinned(n, dev):
    """Page-locked staging buffer for the detections, cached per (thread, device, size); the rows are copied out of it
    before decode_nms returns."""
    key = (threading.get_ident(), dev, n)
    t = _PINNED.get(key)
    if t is None:
        t = _PINNED[key] = torch.empty(n, dtype=torch.float32, pin_memory=True)
    return t


def perspective_matrix(src_pts):
    """The homography of `perspective_transform` (DenseBox.py:3455-3476): the four landmark points (left-up, right-up,
    right-down, left-down) are mapped onto the corners of their axis-aligned bounding rectangle.  Returns (M, Minv) as
    float64 3 x 3 arrays — cv2.getPerspectiveTransform's 8 x 8 linear system and the cofactor inverse cv2 uses."""
    src = np.asarray(src_pts, dtype=np.float32).astype(np.float64)
    if src.shape != (4, 2):
        raise ValueError("src_pts must be 4 points (x, y)")
    lu, ru, rd, ld = src
    min_x, max_x = min(lu[0], ld[0]), max(ru[0], rd[0])
    min_y, max_y = min(lu[1], ru[1]), max(ld[1], rd[1])
    dst = np.array([[min_x, min_y], [max_x, min_y], [max_x, max_y], [min_x, max_y]], dtype=np.float64)
    A, b = np.zeros((8, 8)), np.zeros(8)
    for i in range(4):
        x, y = src[i]
        X, Y = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * X, -y * X]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
        b[i], b[i + 4] = X, Y
    M = np.append(np.linalg.solve(A, b), 1.0).reshape(3, 3)
    a = M
    det = (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
           + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    if det == 0.0:
        raise ValueError("degenerate landmark quadrilateral")
    d = 1.0 / det
    Minv = np.array([
        [(a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) * d, (a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]) * d, (a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]) * d],
        [(a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) * d, (a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]) * d, (a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]) * d],
        [(a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]) * d, (a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]) * d, (a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]) * d]])
    return M, Minv


def perspective_transform(img, src_pts):
    """`perspective_transform(img, src_pts)` of the reference (DenseBox.py:3446-3481) on the GPU: rectify the plate
    spanned by the four decoded landmarks into a canvas 1.5x the image.  img: uint8 [H,W,C] CUDA tensor (C <= 4);
    returns a uint8 [int(1.5 H + .5), int(1.5 W + .5), C] CUDA tensor, bit-identical to cv2.warpPerspective."""
    if not (torch.is_tensor(img) and img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3):
        raise RuntimeError("perspective_transform takes a uint8 [H,W,C] CUDA tensor (no CPU fallback)")
    img = img.contiguous()
    H, W, C = img.shape
    _, minv = perspective_matrix(src_pts)
    dH, dW = int(H * 1.5 + 0.5), int(W * 1.5 + 0.5)
    out = torch.empty(dH, dW, C, dtype=torch.uint8, device=img.device)
    m = (ctypes.c_double * 9)(*minv.reshape(-1).tolist())
    with torch.cuda.device(img.device):
        check(lib().dbx_warp_perspective_u8(ptr(img), c_int(H), c_int(W), c_int(C), m, ptr(out), c_int(dH), c_int(dW),
                                            stream_ptr()), "warp_perspective_u8")
    return out
