"""GPU parity of decode+NMS: bit-exact against the fixture produced by the reference's parse_out_MN /
parse_DetLMLOC / NMS and against the oracle on random batches (score ties do not occur in random floats)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_decode_matches_reference_fixture():
    from densebox_b200 import decode_nms
    d = np.load(os.path.join(G, "decode_nms.npz"))
    t = lambda k: torch.from_numpy(d[k]).cuda()
    for lm, ref in ((None, d["dets_mn"]), (t("lml"), d["dets_lmloc"])):
        got = decode_nms(t("score"), t("loc"), lm, K=10, nms_thresh=1e9)[0]  # threshold off: all 10 rows survive
        assert np.array_equal(got, ref)
        got = decode_nms(t("score"), t("loc"), lm, K=10, nms_thresh=0.4)[0]
        keep = O.nms(ref, 0.4)
        assert np.array_equal(got, ref[sorted(keep)])


@pytest.mark.parametrize("hw", [(60, 60), (256, 256), (16, 24)])
def test_decode_nms_batch_vs_oracle(hw):
    from densebox_b200 import decode_nms
    h, w = hw
    N = 4
    g = torch.Generator().manual_seed(h)
    score = torch.randn(N, 1, h, w, generator=g)
    # clustered boxes so that NMS actually suppresses: loc offsets make neighbouring cells decode to similar boxes
    loc = torch.stack([torch.full((N, h, w), 2.0), torch.full((N, h, w), 2.0), torch.full((N, h, w), -3.0),
                       torch.full((N, h, w), -3.0)], 1) + 0.3 * torch.randn(N, 4, h, w, generator=g)
    bump = torch.zeros(N, 1, h, w)
    bump[:, :, h // 2 - 1:h // 2 + 2, w // 2 - 1:w // 2 + 2] = 6.0  # a 3x3 cluster of top scores
    score = score + bump
    lml = torch.randn(N, 8, h, w, generator=g) * 4
    got = decode_nms(score.cuda(), loc.cuda(), lml.cuda(), K=10, nms_thresh=0.4)
    suppressed = 0
    for i in range(N):
        ref = O.decode(score[i:i + 1], loc[i:i + 1], lml[i:i + 1], K=10)
        keep = sorted(O.nms(ref, 0.4))
        suppressed += 10 - len(keep)
        assert np.array_equal(got[i], ref[keep])
    assert suppressed > 0


def test_decode_on_engine_outputs_strided():
    """Works directly on the NCHW views returned by forward (non-contiguous slices of a larger buffer)."""
    from densebox_b200 import decode_nms
    g = torch.Generator().manual_seed(3)
    buf = torch.randn(2, 20, 20, 16, generator=g).cuda()  # NHWC fp32 like head_out
    score = buf[..., 0:1].permute(0, 3, 1, 2)
    loc = buf[..., 1:5].permute(0, 3, 1, 2)
    got = decode_nms(score, loc, None, K=7, nms_thresh=0.4)
    for i in range(2):
        ref = O.decode(score[i:i + 1].cpu().contiguous(), loc[i:i + 1].cpu().contiguous(), None, K=7)
        assert np.array_equal(got[i], ref[sorted(O.nms(ref, 0.4))])


def test_perspective_transform_matches_reference_fixture():
    """perspective_transform (DenseBox.py:3446-3481) on the GPU against the image the reference produced with
    cv2.warpPerspective: OpenCV's fixed-point bilinear remap restated bit for bit.  The homography comes from a
    different LU solve than cv2's (1e-15 relative), so a source coordinate may land on the other side of a 1/32-pixel
    rounding boundary for a handful of pixels: at most 0.05 % of the values may differ, and then by what a 1/32-pixel
    shift does on a white-noise image (<= 8 grey levels); two of the three cases match bit for bit."""
    from densebox_b200 import perspective_transform
    d = np.load(os.path.join(G, "perspective.npz"))
    img = torch.from_numpy(d["img"]).cuda()
    exact = 0
    for i, pts in enumerate(d["pts"]):
        got = perspective_transform(img, pts).cpu().numpy()
        want = d["out%d" % i]
        assert got.shape == want.shape
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 8 and (diff > 0).mean() <= 5e-4, (i, diff.max(), (diff > 0).mean())
        exact += int(diff.max() == 0)
        assert want.max() > 0
    assert exact >= 2
