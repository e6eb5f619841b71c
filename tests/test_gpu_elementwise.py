"""GPU parity of the HBM-bound kernels against plain torch fp32 ops on the same bf16-rounded inputs (outputs are
rounded once to bf16: tolerance 2^-8 relative + 1e-6)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


def rb(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def close(got, ref, tol=2 ** -7):
    err = (got.float() - ref).abs()
    assert bool((err <= tol * ref.abs() + 1e-6).all()), (err.max().item(), ref.abs().max().item())


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


def test_im2col():
    from densebox_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 3, 16, 24, generator=g, device="cuda")
    out = torch.empty(2, 16, 24, 64, dtype=torch.bfloat16, device="cuda")
    ops.im2col3x3_c3(x, out)
    ref = F.unfold(x, 3, padding=1).view(2, 3, 9, 16, 24).permute(0, 3, 4, 2, 1).reshape(2, 16, 24, 27)
    assert torch.equal(out[..., :27].float(), ref.to(torch.bfloat16).float())
    assert bool((out[..., 27:] == 0).all())


@pytest.mark.parametrize("strided", [False, True])
def test_maxpool_fwd_bwd(strided):
    from densebox_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    N, H, W, C = 2, 12, 20, 64
    cs, coff = (C + 16, 8) if strided else (C, 0)
    ybuf = rb((N, H, W, cs), g).relu()
    ybuf[:, ::4, ::4, :] = 0.5  # exact ties inside windows: gradient must go to the FIRST maximum
    y = ops.View(ybuf, C, coff)
    out = torch.zeros(N, H // 2, W // 2, C, dtype=torch.bfloat16, device="cuda")
    ops.maxpool2x2_fwd(y, out)
    yt = nchw(y.tensor()).requires_grad_(True)
    ref = F.max_pool2d(yt, 2, 2)
    assert torch.equal(nchw(out), ref.detach())
    dp = rb((N, H // 2, W // 2, C), g)
    add = rb((N, H, W, C), g)
    for use_add in (False, True):
        dy = torch.zeros(N, H, W, C, dtype=torch.bfloat16, device="cuda")
        ops.maxpool2x2_bwd(y, dp, dy, add=add if use_add else None)
        gref, = torch.autograd.grad(ref, yt, nchw(dp), retain_graph=True)
        if use_add:
            gref = gref + nchw(add)
        gref = gref * (yt.detach() > 0)
        close(nchw(dy), gref)


@pytest.mark.parametrize("shape", [(30, 30, 60, 60, 512), (24, 24, 60, 60, 64), (7, 9, 16, 20, 64)])
def test_upsample_fwd_bwd(shape):
    from densebox_b200 import ops
    h, w, H, W, C = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    N = 2
    x = rb((N, h, w, C), g).relu()
    out = torch.zeros(N, H, W, C + 64, dtype=torch.bfloat16, device="cuda")
    ops.upsample_bilinear_fwd(x, ops.View(out, C, 32))
    xt = nchw(x).requires_grad_(True)
    ref = F.interpolate(xt, size=(H, W), mode="bilinear", align_corners=True)
    close(nchw(out[..., 32:32 + C]), ref.detach())
    assert bool((out[..., :32] == 0).all()) and bool((out[..., 32 + C:] == 0).all())
    dout = rb((N, H, W, C), g)
    gref, = torch.autograd.grad(ref, xt, nchw(dout))
    for mask in (False, True):
        din = torch.zeros(N, h, w, C, dtype=torch.bfloat16, device="cuda")
        ops.upsample_bilinear_bwd(dout, din, relu_y=x if mask else None)
        r = gref * (xt.detach() > 0) if mask else gref
        err = (nchw(din) - r).abs().max().item()
        assert err <= 2 ** -7 * r.abs().max().item() + 1e-6, err


@pytest.mark.parametrize("C", [16, 64, 768, 1024, 2048])
def test_colsum(C):
    from densebox_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    cs = 64 if C == 16 else C
    buf = rb((2, 30, 30, cs), g)
    db = torch.ones(C, device="cuda")
    ops.colsum(ops.View(buf, C, 0), db)
    ref = 1.0 + buf[..., :C].float().sum(dim=(0, 1, 2))
    assert (db - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_maxpool_bwd_from_argmax_map_matches_full_resolution_backward():
    """maxpool2x2_fwd_idx + maxpool2x2_bwd_idx (pooled map + 2-bit arg-max map) must reproduce maxpool2x2_fwd +
    maxpool2x2_bwd (which re-reads the full-resolution activation) bit for bit, including ties (post-ReLU zeros:
    first maximum in scan order) and the fused bias-gradient column sums."""
    from densebox_b200 import ops
    torch.manual_seed(0)
    N, H, W, C = 3, 24, 16, 128
    y = torch.relu(torch.randn(N, H, W, C, device="cuda")).to(torch.bfloat16)  # ~50 % exact zeros -> many ties
    y[:, ::2, ::2] = y[:, 1::2, 1::2]                                          # and equal non-zero maxima
    dp = torch.randn(N, H // 2, W // 2, C, device="cuda").to(torch.bfloat16)
    out_a = torch.empty(N, H // 2, W // 2, C, dtype=torch.bfloat16, device="cuda")
    out_b = torch.empty_like(out_a)
    idx = torch.empty(N, H // 2, W // 2, C // 8, dtype=torch.int16, device="cuda")
    ops.maxpool2x2_fwd(y, out_a)
    ops.maxpool2x2_fwd_idx(y, out_b, idx)
    assert torch.equal(out_a, out_b)
    dy_a = torch.empty_like(y)
    dy_b = torch.full_like(y, 7.0)
    ops.maxpool2x2_bwd(y, dp, dy_a)
    db = torch.zeros(C, device="cuda")
    ops.maxpool2x2_bwd_idx(out_b, dp, idx, dy_b, db=db)
    assert torch.equal(dy_a, dy_b)
    ref = dy_a.float().sum((0, 1, 2))
    assert (db - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
