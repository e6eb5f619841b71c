"""The bias gradients produced inside the epilogues of the data-gradient / pool-backward / upsample-backward launches
(ConvEpilogue::colsum, maxpool2x2_bwd / upsample_bilinear_bwd `db`) must equal the stand-alone column-sum kernel run
on the same stored gradients: same addends (the bf16 values written to HBM), fp32 accumulation, different order.
Tolerance: 2e-4 relative to the largest entry of each bias gradient (fp32 summation order over <= 1.8 M addends)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_e2e import build, make_inputs  # noqa: E402

pytestmark = pytest.mark.gpu


def _bias_grads(variant, fuse):
    from densebox_b200 import densebox_loss
    os.environ["DBX_FUSE_BIAS"] = "1" if fuse else "0"
    os.environ["DBX_FUSE_BIAS_SHORT"] = "1"  # also the two short-K data gradients the engine leaves un-fused by default
    try:
        _, net = build(variant)
        net = net.cuda().eval()
        x, lab, rand, lm_rand = make_inputs(3, variant)
        outs = net(x.cuda())
        kw = {}
        if variant == "lm":
            score, loc, lm, rf = outs
            kw = dict(lm=lm, rf=rf, vertices=lab["vertices"], lm_rand_neg_idx=lm_rand)
        else:
            score, loc = outs
        L = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand, **kw)
        L.backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().float().cpu() for n, p in net.named_parameters()
                if p.grad is not None and n.endswith(".bias")}, float(L.detach())
    finally:
        os.environ.pop("DBX_FUSE_BIAS", None)
        os.environ.pop("DBX_FUSE_BIAS_SHORT", None)


@pytest.mark.parametrize("variant", ["densebox", "lm"])
def test_fused_bias_gradients_match_colsum(variant):
    ref, L0 = _bias_grads(variant, fuse=False)
    got, L1 = _bias_grads(variant, fuse=True)
    assert L0 == L1
    assert set(ref) == set(got) and len(ref) >= 15
    for n in ref:
        scale = ref[n].abs().max().item()
        assert scale > 0, n
        err = (got[n] - ref[n]).abs().max().item() / scale
        assert err <= 2e-4, (n, err)


def _head_grads(direct):
    """Train-mode step (Philox dropout in place) with the conv5_2 data gradient either as the streaming kernel
    (heads2_dgrad, default) or as the K = 64 tensor-core launch it replaced."""
    from densebox_b200 import densebox_loss
    os.environ["DBX_HEADS2_DGRAD"] = "1" if direct else "0"
    try:
        _, net = build("lm")
        net = net.cuda().train()
        x, lab, rand, lm_rand = make_inputs(2, "lm")
        torch.manual_seed(11)  # dropout stream = (torch seed, module instance, call number): pin all three
        net._instance, net._drop_calls = 1, 0
        score, loc, lm, rf = net(x.cuda())
        L = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand, lm=lm, rf=rf, vertices=lab["vertices"],
                          lm_rand_neg_idx=lm_rand)
        L.backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().float().cpu() for n, p in net.named_parameters() if p.grad is not None}, float(L.detach())
    finally:
        os.environ.pop("DBX_HEADS2_DGRAD", None)


def test_heads2_dgrad_kernel_matches_tensor_core_path():
    """Same dropout bits, same bf16 inputs; fp32 FMA chain vs tensor-core accumulation differ by summation order only:
    every gradient downstream of d_hd (conv5_1 weights/biases, backbone) within 2e-3 of its norm (bf16 rounding of
    d_hd flips a last bit on ~1 % of the elements), the conv5_2 / refine gradients upstream bit-identical."""
    ref, L0 = _head_grads(direct=False)
    got, L1 = _head_grads(direct=True)
    assert L0 == L1
    for n in ref:
        d = (got[n] - ref[n]).norm().item() / (ref[n].norm().item() + 1e-30)
        if n.startswith(("conv5_2", "conv6", "output_")) and ".2." in n or n.startswith("conv5_2"):
            assert d <= 1e-6, (n, d)
        assert d <= 2e-3 or n.startswith(("conv1", "conv2")), (n, d)  # early layers: chaotic ReLU routing
    for n in ("conv5_1_det.weight", "conv5_1_det.bias", "conv5_1_loc.bias", "conv5_1_landmark.weight"):
        assert n in ref


def _all_grads(pairs):
    from densebox_b200 import densebox_loss
    os.environ["DBX_CONV1_PAIRS"] = "1" if pairs else "0"   # read when the engine is created
    try:
        _, net = build("densebox")
        net = net.cuda().eval()
        x, lab, rand, _ = make_inputs(3, "densebox")
        score, loc = net(x.cuda())
        L = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand)
        L.backward()
        torch.cuda.synchronize()
        g = {n: p.grad.detach().float().cpu() for n, p in net.named_parameters() if p.grad is not None}
        return g, float(L.detach()), score.detach().cpu(), {n: p.detach().float().cpu() for n, p in net.named_parameters()}
    finally:
        os.environ.pop("DBX_CONV1_PAIRS", None)


def test_conv1_1_pairs_layout_matches_64_channel_layout():
    """conv1_1 as the 128 x 64 matrix [[W 0] [0 W]] over rows of two pixels (default) vs the plain 64-channel im2col
    layout: the forward pass must be bit-identical (same products, zeros elsewhere), hence every other gradient too;
    the conv1_1 weight / bias gradients (one wgrad + fold, bias from the constant-1 tap) differ by fp32 summation
    order only: 1e-4 of the largest entry."""
    g0, L0, s0, w0 = _all_grads(pairs=False)
    g1, L1, s1, w1 = _all_grads(pairs=True)
    assert torch.equal(s0, s1) and L0 == L1
    for n in w0:
        assert torch.equal(w0[n], w1[n]), n                       # parameters round-trip through both layouts
    assert set(g0) == set(g1)
    for n in g0:
        scale = g0[n].abs().max().item()
        err = (g1[n] - g0[n]).abs().max().item() / (scale + 1e-30)
        if n.startswith(("conv1_1", "conv1_1_1")):
            assert scale > 0 and err <= 1e-4, (n, err)
        else:
            assert err <= 2e-5, (n, err)                          # red.add order in wgrad only


def _fwd_bwd(fuse_pool):
    from densebox_b200 import densebox_loss
    os.environ["DBX_POOL_FUSE"] = "1" if fuse_pool else "0"
    try:
        _, net = build("lm")
        net = net.cuda().eval()
        x, lab, rand, lm_rand = make_inputs(3, "lm")
        outs = net(x.cuda())
        eng = next(iter(net._engines.values()))
        N = 3
        pooled = [eng.buffer("p1", torch.bfloat16, (N, 120, 120, 64)).clone(), eng.buffer("p2", torch.bfloat16, (N, 60, 60, 128)).clone(),
                  eng.buffer("p3", torch.bfloat16, (N, 30, 30, 256)).clone(), eng.buffer("fusion", torch.bfloat16).clone(),
                  eng.buffer("pi1", torch.int16).clone(), eng.buffer("pi2", torch.int16).clone()]
        score, loc, lm, rf = outs
        L = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand, lm=lm, rf=rf, vertices=lab["vertices"],
                          lm_rand_neg_idx=lm_rand)
        L.backward()
        torch.cuda.synchronize()
        g = {n: p.grad.detach().float().cpu() for n, p in net.named_parameters() if p.grad is not None}
        return [o.detach().cpu() for o in outs], pooled, g, float(L.detach())
    finally:
        os.environ.pop("DBX_POOL_FUSE", None)


def test_pool_fused_into_conv_epilogue_matches_separate_pool_kernel():
    """pool1 / pool2 computed inside the epilogues of conv1_2 / conv2_2 (the full-resolution activations never reach
    HBM) against conv + maxpool2x2_fwd_kernel: pooled maps, arg-max maps (first maximum wins ties), head outputs and
    loss bit-identical; gradients equal up to the summation order of the atomics."""
    outs_a, pooled_a, g_a, L_a = _fwd_bwd(False)
    outs_b, pooled_b, g_b, L_b = _fwd_bwd(True)
    for a, b in zip(pooled_a, pooled_b):
        assert torch.equal(a, b)
    for a, b in zip(outs_a, outs_b):
        assert torch.equal(a, b)
    assert L_a == L_b
    for n in g_a:
        assert (g_a[n] - g_b[n]).abs().max().item() <= 1e-4 * g_a[n].abs().max().item() + 1e-30, n
