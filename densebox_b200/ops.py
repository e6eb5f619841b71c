"""Kernel-level Python wrappers over the C ABI (used by the parity tests and by the module layer).

Activations are NHWC bf16 torch tensors `[N,H,W,cs]`; a *view* is (buffer, C, coff): channels [coff, coff+C).
Nothing here computes on the CPU — each function is one call into libdensebox_b200.so.
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream_ptr

c_int = ctypes.c_int


class View:
    """Channel window of an NHWC bf16/fp32 buffer."""

    def __init__(self, buf, C=None, coff=0):
        assert buf.dim() == 4 and buf.is_contiguous()
        self.buf, self.coff = buf, coff
        self.C = buf.shape[3] - coff if C is None else C
        self.N, self.H, self.W, self.cs = buf.shape

    def tensor(self):
        return self.buf[..., self.coff:self.coff + self.C]


def _v(x):
    return x if isinstance(x, View) else View(x)


def pack_weight_kmajor(w, cin_pad=None, cout_pad=None):
    """torch [Cout,Cin,R,S] -> bf16 [cout_pad][R*S*cin_pad] (tap-major, channel-minor). Test/plumbing helper."""
    co, ci, R, S = w.shape
    cin_pad = cin_pad or ci
    cout_pad = cout_pad or co
    out = torch.zeros(cout_pad, R * S, cin_pad, dtype=torch.bfloat16, device=w.device)
    out[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, R * S, ci).to(torch.bfloat16)
    return out.reshape(cout_pad, R * S * cin_pad).contiguous()


def pack_weight_dgrad(w, cin_pad=None, cout_pad=None):
    """torch [Cout,Cin,R,S] -> bf16 [cin_pad][R*S*cout_pad] with the taps flipped (filter of the data gradient)."""
    return pack_weight_kmajor(w.flip(2, 3).permute(1, 0, 2, 3), cin_pad=cout_pad, cout_pad=cin_pad)


def conv_fprop(x, wk, R, S, pad, out, bias=None, relu=False, aux=None, aux_mode=0, block_n=0):
    x, out = _v(x), _v(out)
    aux_v = _v(aux) if aux is not None else None
    out_fp32 = 1 if out.buf.dtype == torch.float32 else 0
    rc = lib().dbx_conv_fprop(
        ptr(x.buf), c_int(x.N), c_int(x.H), c_int(x.W), c_int(x.C), c_int(x.cs), c_int(x.coff), ptr(wk), c_int(R),
        c_int(S), c_int(pad), c_int(out.C), ptr(bias), c_int(int(relu)), ptr(aux_v.buf if aux_v else None),
        c_int(aux_v.cs if aux_v else 0), c_int(aux_v.coff if aux_v else 0), c_int(aux_mode), ptr(out.buf),
        c_int(out.cs), c_int(out.coff), c_int(out_fp32), c_int(block_n), stream_ptr())
    check(rc, "conv_fprop")
    return out


def conv_wgrad(x, dy, R, S, pad, dw, block_n=0):
    x, dy = _v(x), _v(dy)
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    rc = lib().dbx_conv_wgrad(
        ptr(x.buf), c_int(x.N), c_int(x.H), c_int(x.W), c_int(x.C), c_int(x.cs), c_int(x.coff), ptr(dy.buf),
        c_int(dy.C), c_int(dy.cs), c_int(dy.coff), c_int(R), c_int(S), c_int(pad), ptr(dw), c_int(block_n),
        stream_ptr())
    check(rc, "conv_wgrad")
    return dw


def im2col3x3_c3(x, out):
    N, _, H, W = x.shape
    check(lib().dbx_im2col3x3_c3(ptr(x), ptr(out), c_int(N), c_int(H), c_int(W), stream_ptr()), "im2col3x3_c3")
    return out


def _vargs(v):
    return [ptr(v.buf), c_int(v.cs), c_int(v.coff)]


def maxpool2x2_fwd(y, out):
    y, out = _v(y), _v(out)
    check(lib().dbx_maxpool2x2_fwd(ptr(y.buf), c_int(y.N), c_int(y.H), c_int(y.W), c_int(y.C), c_int(y.cs),
                                   c_int(y.coff), *_vargs(out), stream_ptr()), "maxpool2x2_fwd")
    return out


def maxpool2x2_fwd_idx(y, out, idx):
    """Pooling + compact arg-max map idx (int16 tensor [N, H/2, W/2, C/8])."""
    y, out = _v(y), _v(out)
    assert idx.dtype == torch.int16 and idx.is_contiguous() and idx.numel() == out.N * out.H * out.W * (out.C // 8)
    check(lib().dbx_maxpool2x2_fwd_idx(ptr(y.buf), c_int(y.N), c_int(y.H), c_int(y.W), c_int(y.C), c_int(y.cs),
                                       c_int(y.coff), *_vargs(out), ptr(idx), stream_ptr()), "maxpool2x2_fwd_idx")
    return out


def maxpool2x2_bwd_idx(p, dp, idx, dy, db=None):
    """Backward from the pooled activation + arg-max map (no full-resolution read); db += column sums of dy."""
    p, dp, dy = _v(p), _v(dp), _v(dy)
    check(lib().dbx_maxpool2x2_bwd_idx(ptr(p.buf), c_int(dy.N), c_int(dy.H), c_int(dy.W), c_int(dy.C), c_int(p.cs),
                                       c_int(p.coff), *_vargs(dp), ptr(idx), *_vargs(dy), ptr(db), stream_ptr()),
          "maxpool2x2_bwd_idx")
    return dy


def maxpool2x2_bwd(y, dp, dy, add=None):
    y, dp, dy = _v(y), _v(dp), _v(dy)
    a = _vargs(_v(add)) if add is not None else [ptr(None), c_int(0), c_int(0)]
    check(lib().dbx_maxpool2x2_bwd(ptr(y.buf), c_int(y.N), c_int(y.H), c_int(y.W), c_int(y.C), c_int(y.cs),
                                   c_int(y.coff), *_vargs(dp), *a, *_vargs(dy), stream_ptr()), "maxpool2x2_bwd")
    return dy


def upsample_bilinear_fwd(x, out):
    x, out = _v(x), _v(out)
    check(lib().dbx_upsample_bilinear_fwd(ptr(x.buf), c_int(x.N), c_int(x.H), c_int(x.W), c_int(x.C), c_int(x.cs),
                                          c_int(x.coff), ptr(out.buf), c_int(out.H), c_int(out.W), c_int(out.cs),
                                          c_int(out.coff), stream_ptr()), "upsample_bilinear_fwd")
    return out


def upsample_bilinear_bwd(dout, din, relu_y=None):
    dout, din = _v(dout), _v(din)
    y = _vargs(_v(relu_y)) if relu_y is not None else [ptr(None), c_int(0), c_int(0)]
    check(lib().dbx_upsample_bilinear_bwd(ptr(dout.buf), c_int(dout.N), c_int(dout.H), c_int(dout.W), c_int(dout.C),
                                          c_int(dout.cs), c_int(dout.coff), *y, ptr(din.buf), c_int(din.H),
                                          c_int(din.W), c_int(din.cs), c_int(din.coff), stream_ptr()),
          "upsample_bilinear_bwd")
    return din


def colsum(dy, db):
    dy = _v(dy)
    check(lib().dbx_colsum(ptr(dy.buf), c_int(dy.N), c_int(dy.H), c_int(dy.W), c_int(dy.C), c_int(dy.cs),
                           c_int(dy.coff), ptr(db), stream_ptr()), "colsum")
    return db
