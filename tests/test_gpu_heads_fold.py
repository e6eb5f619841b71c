"""Inference nets fold conv5_1_* and conv5_2_* into ONE 768 -> HC matrix (DenseBox.py:158-178: Conv -> Dropout -> Conv,
no non-linearity, so without dropout the product of the two weight matrices is the same linear map): heads_fold_kernel
+ one small-N GEMM, the 512-channel hidden maps are never computed.

Compared with the two-GEMM path of the same engine (DBX_HEADS_FOLD=0) and with the oracle.  Tolerances: the two paths
round different intermediates to bf16 (hidden maps vs. the folded matrix); each must stay within the 3e-2 bound of the
other inference tests against the oracle (of the largest entry of the map), hence within 6e-2 of each other."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_e2e import O, build, oracle_params  # noqa: E402

pytestmark = pytest.mark.gpu


def _infer(variant, env, H, W):
    for k, v in env.items():
        os.environ[k] = v
    try:
        _, net = build(variant)
        net = net.cuda().eval()
        g = torch.Generator().manual_seed(11)
        x = torch.randn(2, 3, H, W, generator=g).bfloat16().float()
        with torch.no_grad():
            outs = [o.float().cpu() for o in net(x.cuda())]
        torch.cuda.synchronize()
        return net, x, outs, None
    finally:
        for k in env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("variant,H,W", [("densebox", 240, 240), ("lm", 240, 240), ("lmloc", 136, 200)])
def test_folded_heads_match_two_gemm_path_and_oracle(variant, H, W):
    net, x, plain, _ = _infer(variant, {"DBX_HEADS_FOLD": "0"}, H, W)
    _, _, folded, _ = _infer(variant, {}, H, W)
    _, _, folded2, _ = _infer(variant, {}, H, W)
    with torch.no_grad():
        ref = O.forward(oracle_params(net, variant), x, variant)
    differs = False
    for a, b, b2, r in zip(plain, folded, folded2, ref):
        s = r.abs().max().item()
        ea, eb = (a - r).abs().max().item(), (b - r).abs().max().item()
        print("%s map %s: max |ref| %.4f, two-GEMM err %.5f, folded err %.5f" % (variant, tuple(r.shape), s, ea, eb))
        assert ea <= 3e-2 * s and eb <= 3e-2 * s, (ea, eb, s)
        assert (a - b).abs().max().item() <= 6e-2 * s
        assert torch.equal(b, b2)                       # no atomics on the folded path: bit-reproducible
        differs = differs or not torch.equal(a, b)
    assert differs                                      # the switch really selected a different path


def test_training_net_keeps_the_two_gemm_path():
    """A net that may run backward needs the hidden maps: eval-mode forward with grad enabled must not fold (its head
    outputs equal the DBX_HEADS_FOLD=0 inference outputs bit for bit: same kernels, same weights)."""
    net, x, plain, _ = _infer("densebox", {"DBX_HEADS_FOLD": "0"}, 240, 240)
    _, net2 = build("densebox")
    net2 = net2.cuda().eval()
    outs = [o.detach().float().cpu() for o in net2(x.cuda())]   # grad enabled -> training workspace
    for a, b in zip(plain, outs):
        assert torch.equal(a, b)
