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
