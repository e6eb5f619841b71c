"""CPU check of the algebra behind the inference head fold (densebox_b200/csrc: heads_fold_kernel): the reference puts
only nn.Dropout between conv5_1_* and conv5_2_* (DenseBox.py:158-178), so in eval mode
    conv5_2(conv5_1(x)) == conv1x1(x; W2 @ W1, b2 + W2 @ b1)
for every head.  Checked against the oracle's own eval forward (fp32; the fold in float64) for all three variants: the
folded maps equal the oracle's head maps to 2e-6 of the largest entry — tests/test_gpu_heads_fold.py then only has to
bound the bf16 rounding of the engine."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402

CLS = {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}


@pytest.mark.parametrize("variant", ["densebox", "lm", "lmloc"])
def test_fold_of_the_two_head_convolutions_is_the_same_linear_map(variant):
    import densebox_b200  # only supplies the seeded module (state_dict); nothing here runs on a GPU
    torch.manual_seed(5)
    net = getattr(densebox_b200, CLS[variant])(O.seeded_vgg19(0))
    P = O.params_from_state_dict(net.state_dict(), variant)
    x = torch.randn(1, 3, 64, 88)
    with torch.no_grad():
        out, inter = O.forward(P, x, variant, return_intermediates=True)
    if variant == "densebox":
        ref = {"det": out[0], "loc": out[1]}
    elif variant == "lm":
        ref = {"det": out[0], "loc": out[1], "landmark": out[2]}
    else:
        ref = {"det": out[0], "loc": out[2], "landmark": out[3], "lmloc": out[4]}
    assert set(ref) == {h for h, _ in O.HEADS[variant]}
    fusion = inter["fusion"].double()
    for h, r in ref.items():
        w1, b1 = P["conv5_1_%s.weight" % h].double()[:, :, 0, 0], P["conv5_1_%s.bias" % h].double()
        w2, b2 = P["conv5_2_%s.weight" % h].double()[:, :, 0, 0], P["conv5_2_%s.bias" % h].double()
        got = F.conv2d(fusion, (w2 @ w1)[:, :, None, None], b2 + w2 @ b1).float()
        s = r.abs().max().item()
        assert s > 0 and got.shape == r.shape
        assert (got - r).abs().max().item() <= 2e-6 * s + 1e-7, h
