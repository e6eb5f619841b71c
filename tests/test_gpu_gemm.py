"""GPU parity of the two tcgen05 kernels (conv_fprop / conv_wgrad) against torch fp32 convolutions fed the same
bf16-rounded operands.  Tolerance: the kernels accumulate in fp32 and round ONCE to bf16 on store, so
|err| <= 2^-8 * |ref| + accumulation-order noise; we assert max|err| <= 1e-2 * max|ref| (fp32 outputs: 2e-3).

Run as a script (`python tests/test_gpu_gemm.py`) to get a non-stopping diagnostic table.
"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rand_bf16(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


# (name, N, H, W, cin, cout, R, pad, opts)
FPROP_CASES = [
    ("1x1_single_tile", 1, 8, 16, 64, 64, 1, 0, {}),
    ("1x1_k256", 2, 8, 16, 256, 128, 1, 0, {}),
    ("3x3_one_tile", 1, 8, 16, 64, 64, 3, 1, {}),
    ("3x3_16x16", 1, 16, 16, 64, 64, 3, 1, {"bias": True, "relu": True}),
    ("3x3_60_c128_256", 2, 60, 60, 128, 256, 3, 1, {"bias": True, "relu": True}),
    ("3x3_30_c256_512", 4, 30, 30, 256, 512, 3, 1, {"bias": True, "relu": True}),
    ("3x3_120_c64_128", 2, 120, 120, 64, 128, 3, 1, {"bias": True, "relu": True}),
    ("3x3_240_c64_64", 1, 240, 240, 64, 64, 3, 1, {"bias": True, "relu": True}),
    ("3x3_odd_13x27", 3, 13, 27, 64, 48, 3, 1, {"bias": True}),
    ("5x5_pad0", 2, 28, 28, 64, 64, 5, 0, {"bias": True}),
    ("3x3_pad0", 2, 30, 30, 64, 64, 3, 0, {"bias": True}),
    ("3x3_pad2_dgrad_like", 2, 28, 28, 64, 64, 3, 2, {}),
    ("1x1_fp32_n16", 2, 60, 60, 1024, 16, 1, 0, {"bias": True, "fp32": True}),
    ("1x1_heads_768_1024", 1, 60, 60, 768, 1024, 1, 0, {"bias": True, "aux": 2}),
    ("3x3_relu_mask", 2, 30, 30, 128, 128, 3, 1, {"aux": 1}),
    ("3x3_strided_views", 2, 60, 60, 128, 256, 3, 1, {"bias": True, "relu": True, "strided": True}),
    ("3x3_bn128", 2, 30, 30, 256, 512, 3, 1, {"bias": True, "relu": True, "block_n": 128}),
]

WGRAD_CASES = [
    ("w_1x1_small", 1, 8, 16, 64, 64, 1, 0, {}),
    ("w_1x1_c128", 2, 16, 16, 128, 128, 1, 0, {}),
    ("w_3x3_small", 1, 16, 16, 64, 64, 3, 1, {}),
    ("w_3x3_60_c128_256", 2, 60, 60, 128, 256, 3, 1, {}),
    ("w_3x3_30_c256_512", 4, 30, 30, 256, 512, 3, 1, {}),
    ("w_3x3_240_c64_64", 1, 240, 240, 64, 64, 3, 1, {}),
    ("w_3x3_120_c64_128", 2, 120, 120, 64, 128, 3, 1, {}),
    ("w_5x5_pad0", 2, 28, 28, 64, 64, 5, 0, {}),
    ("w_1x1_cout16", 2, 60, 60, 1024, 16, 1, 0, {}),
    ("w_1x1_heads", 1, 60, 60, 768, 1024, 1, 0, {}),
    ("w_3x3_odd", 3, 13, 27, 64, 48, 3, 1, {}),
    ("w_strided", 2, 30, 30, 128, 128, 3, 1, {"strided": True}),
]


def run_fprop(case):
    from densebox_b200 import ops
    _setup()
    name, N, H, W, cin, cout, R, pad, o = case
    g = torch.Generator(device="cuda").manual_seed(zlib.crc32(name.encode()))
    strided = o.get("strided", False)
    x_cs, x_coff = (cin + 64, 32) if strided else (cin, 0)
    xbuf = _rand_bf16((N, H, W, x_cs), g)
    x = ops.View(xbuf, C=cin, coff=x_coff)
    w = _rand_bf16((cout, cin, R, R), g, scale=(cin * R * R) ** -0.5)
    bias = torch.randn(cout, generator=g, device="cuda") if o.get("bias") else None
    OH, OW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    out_dtype = torch.float32 if o.get("fp32") else torch.bfloat16
    o_cs, o_coff = (cout + 32, 16) if strided else (cout, 0)
    obuf = torch.full((N, OH, OW, o_cs), 7.0, dtype=out_dtype, device="cuda")
    out = ops.View(obuf, C=cout, coff=o_coff)
    aux_mode = o.get("aux", 0)
    aux = None
    if aux_mode == 1:
        aux = _rand_bf16((N, OH, OW, cout), g)
    elif aux_mode == 2:
        aux = ((torch.rand((N, OH, OW, cout), generator=g, device="cuda") < 0.5).float() * 2).to(torch.bfloat16)
    wk = ops.pack_weight_kmajor(w)
    ops.conv_fprop(x, wk, R, R, pad, out, bias=bias, relu=o.get("relu", False), aux=aux, aux_mode=aux_mode,
                   block_n=o.get("block_n", 0))
    torch.cuda.synchronize()
    ref = F.conv2d(x.tensor().double().permute(0, 3, 1, 2), w.double(), bias.double() if bias is not None else None,
                   padding=pad).float()
    if o.get("relu"):
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    if aux_mode == 1:
        ref = torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    elif aux_mode == 2:
        ref = ref * aux.float()
    got = out.tensor().float()
    err = (got - ref).abs()
    tol = (2e-3 if o.get("fp32") else 1e-2) * ref.abs().max().item()
    ok = bool(torch.isfinite(got).all()) and err.max().item() <= tol
    if strided:  # bytes outside the view must be untouched
        ok = ok and bool((obuf[..., :o_coff] == 7.0).all()) and bool((obuf[..., o_coff + cout:] == 7.0).all())
    info = dict(name=name, max_err=err.max().item(), tol=tol, ref_max=ref.abs().max().item(), ok=ok)
    if not ok:
        bad = err > tol
        info["bad_frac"] = bad.float().mean().item()
        idx = bad.nonzero()[:6].tolist()
        info["first_bad"] = [(i, got[tuple(i)].item(), ref[tuple(i)].item()) for i in idx]
        info["bad_rows_mod8"] = torch.bincount((bad.any(dim=3).nonzero()[:, 2] % 8), minlength=8).tolist()
        info["bad_ch_mod64"] = torch.bincount((bad.nonzero()[:, 3] % 64), minlength=64).tolist()
    return info


def run_wgrad(case):
    from densebox_b200 import ops
    _setup()
    name, N, H, W, cin, cout, R, pad, o = case
    g = torch.Generator(device="cuda").manual_seed(zlib.crc32(name.encode()))
    strided = o.get("strided", False)
    OH, OW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    x_cs, x_coff = (cin + 64, 32) if strided else (cin, 0)
    d_cs, d_coff = (cout + 32, 16) if strided else (cout, 0)
    xbuf = _rand_bf16((N, H, W, x_cs), g)
    dbuf = _rand_bf16((N, OH, OW, d_cs), g)
    x = ops.View(xbuf, C=cin, coff=x_coff)
    dy = ops.View(dbuf, C=cout, coff=d_coff)
    dw = torch.zeros(cout, R * R * cin, device="cuda")
    ops.conv_wgrad(x, dy, R, R, pad, dw)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.tensor().double().permute(0, 3, 1, 2).contiguous(), (cout, cin, R, R),
                                      dy.tensor().double().permute(0, 3, 1, 2).contiguous(), padding=pad)
    ref = ref.permute(0, 2, 3, 1).reshape(cout, R * R * cin).float()
    err = (dw - ref).abs()
    tol = 2e-4 * ref.abs().max().item()
    ok = bool(torch.isfinite(dw).all()) and err.max().item() <= tol
    info = dict(name=name, max_err=err.max().item(), tol=tol, ref_max=ref.abs().max().item(), ok=ok)
    if not ok:
        bad = err > tol
        info["bad_frac"] = bad.float().mean().item()
        idx = bad.nonzero()[:6].tolist()
        info["first_bad"] = [(i, dw[tuple(i)].item(), ref[tuple(i)].item()) for i in idx]
        info["bad_row_mod32"] = torch.bincount((bad.nonzero()[:, 0] % 32), minlength=32).tolist()
        info["bad_col_div64"] = torch.bincount((bad.nonzero()[:, 1] // 64)).tolist()
    return info


@pytest.mark.parametrize("case", FPROP_CASES, ids=[c[0] for c in FPROP_CASES])
def test_conv_fprop(case):
    info = run_fprop(case)
    assert info["ok"], info


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad(case):
    info = run_wgrad(case)
    assert info["ok"], info


if __name__ == "__main__":
    sel = sys.argv[1:] or ["fprop", "wgrad"]
    for kind, cases, fn in (("fprop", FPROP_CASES, run_fprop), ("wgrad", WGRAD_CASES, run_wgrad)):
        if kind not in sel:
            continue
        for c in cases:
            try:
                print(kind, fn(c), flush=True)
            except Exception as e:  # keep going: one call should report every failure
                print(kind, c[0], "EXCEPTION", repr(e), flush=True)
                if "CUDA error" in repr(e) or "launch failure" in repr(e) or "illegal" in repr(e):
                    print("sticky CUDA error, stopping", flush=True)
                    sys.exit(1)
