"""GPU parity of the fused loss kernel: (1) against the fixtures produced by the unmodified reference loop bodies
(tests/golden/loss_*.npz) — masks and quotas bit-exact, loss and gradients to fp32 round-off; (2) against the oracle
on random cases incl. boxes at the border (python-slice clipping), large boxes (big top-k), empty boxes and pos/neg
patches; (3) data-parallel quota (global positive count) equals the single-batch result."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a, shape):
    return np.unpackbits(a)[:int(np.prod(shape))].reshape(shape)


def run_cuda(variant, outs, bbox, rand, vertices=None, lm_rand=None, labels=None, **kw):
    from densebox_b200 import densebox_loss
    o = [t.detach().cuda().requires_grad_(True) for t in outs]
    if variant == "densebox":
        args = dict()
        score, loc = o
    elif variant == "lm":
        score, loc, lm, rf = o
        args = dict(lm=lm, rf=rf, vertices=vertices, lm_rand_neg_idx=lm_rand)
    else:
        score, rf, loc, lm, lmloc = o
        args = dict(lm=lm, rf=rf, lm_loc=lmloc, vertices=vertices, lm_rand_neg_idx=lm_rand)
    L, info = densebox_loss(score, loc, bbox, rand_neg_idx=rand, labels=labels, return_info=True, **args, **kw)
    L.backward()
    return L, info, [t.grad.cpu() for t in o]


@pytest.mark.parametrize("case", ["densebox", "lm", "lmloc", "lmloc_pn"])
def test_loss_kernel_vs_reference_fixture(case):
    d = np.load(os.path.join(G, "loss_%s.npz" % case))
    variant = case.split("_")[0]
    n_out = {"densebox": 2, "lm": 4, "lmloc": 5}[variant]
    outs = [torch.from_numpy(d["out%d" % i].astype(np.float32)) for i in range(n_out)]
    labels = d["labels"] if "labels" in d.files else None
    L, info, grads = run_cuda(variant, outs, d["bbox"], d["rand"], vertices=d["vertices"], lm_rand=d["lm_rand"],
                              labels=labels)
    B = d["bbox"].shape[0]
    assert info["half"] == int(d["half"]) and info["pos"] == int(d["pos"])
    assert np.array_equal(info["mask"].cpu().numpy().reshape(B, 1, 60, 60), bits(d["mask"], (B, 1, 60, 60)))
    if variant != "densebox":
        assert np.array_equal(info["lm_mask"].cpu().numpy().reshape(B, 4, 60, 60), bits(d["lm_mask"], (B, 4, 60, 60)))
    assert abs(L.item() - float(d["loss"])) <= 2e-6 * abs(float(d["loss"]))
    for i, gr in enumerate(grads):
        np.testing.assert_allclose(gr.numpy()[:, :, ::5, ::5], d["grad%d" % i], rtol=1e-5, atol=1e-5)


def _random_case(variant, B, seed, kind):
    rs = np.random.RandomState(seed)
    if kind == "border":      # boxes touching / leaving the map: python slice clipping and wrap-around
        x0 = rs.uniform(-3, 50, B); y0 = rs.uniform(-3, 50, B); w = rs.uniform(4, 30, B); h = rs.uniform(4, 30, B)
    elif kind == "large":     # big boxes -> hundreds of hard negatives per sample
        x0 = rs.uniform(0, 5, B); y0 = rs.uniform(0, 5, B); w = rs.uniform(45, 55, B); h = rs.uniform(45, 55, B)
    else:
        x0 = rs.uniform(5, 30, B); y0 = rs.uniform(5, 30, B); w = rs.uniform(2, 25, B); h = rs.uniform(2, 20, B)
    bbox = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float32)
    verts = np.clip(np.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1) + rs.uniform(-1, 1, (B, 8)), 2.6,
                    56.4).astype(np.float32)
    ch = {"densebox": [1, 4], "lm": [1, 4, 4, 1], "lmloc": [1, 1, 4, 4, 8]}[variant]
    g = torch.Generator().manual_seed(seed)
    outs = [torch.randn(B, c, 60, 60, generator=g) for c in ch]
    rand = np.stack([rs.choice(3600, 3600, replace=False) for _ in range(B)]).astype(np.int64)
    lm_rand = rs.randint(0, 3600, (B, 4)).astype(np.int64)
    return outs, bbox, verts, rand, lm_rand


@pytest.mark.parametrize("variant", ["densebox", "lm", "lmloc"])
@pytest.mark.parametrize("kind", ["interior", "border", "large"])
def test_loss_kernel_vs_oracle(variant, kind):
    B = 5
    outs, bbox, verts, rand, lm_rand = _random_case(variant, B, 100 + len(kind), kind)
    ro = [t.clone().requires_grad_(True) for t in outs]
    L_ref, info_ref = O.loss(tuple(ro), variant, bbox, rand, vertices=verts, lm_rand_idx=lm_rand)
    L_ref.backward()
    L, info, grads = run_cuda(variant, outs, bbox, rand, vertices=verts, lm_rand=lm_rand)
    assert info["half"] == info_ref["half"] and info["pos"] == info_ref["pos"]
    assert np.array_equal(info["mask"].cpu().numpy().reshape(B, 1, 60, 60), info_ref["mask"].astype(np.uint8))
    if variant != "densebox":
        assert np.array_equal(info["lm_mask"].cpu().numpy().reshape(B, 4, 60, 60), info_ref["lm_mask"].astype(np.uint8))
    assert abs(L.item() - L_ref.item()) <= 2e-6 * abs(L_ref.item())
    for gr, r in zip(grads, ro):
        np.testing.assert_allclose(gr.numpy(), r.grad.numpy(), rtol=1e-5, atol=1e-4)


def test_data_parallel_quota_matches_single_batch():
    """Sharding the batch over 2 'ranks' with the batch-global positive count reproduces the single-batch masks and
    loss (DenseBox.py:2864-2868 uses the whole batch)."""
    B = 6
    outs, bbox, verts, rand, lm_rand = _random_case("densebox", B, 7, "interior")
    L, info, _ = run_cuda("densebox", outs, bbox, rand)
    tot, masks = 0.0, []
    for lo, hi in ((0, 3), (3, 6)):
        Ls, infos, _ = run_cuda("densebox", [t[lo:hi] for t in outs], bbox[lo:hi], rand[lo:hi],
                                global_pos_count=info["pos"], global_batch=B)
        assert infos["half"] == info["half"]
        tot += Ls.item()
        masks.append(infos["mask"].cpu())
    assert torch.equal(torch.cat(masks), info["mask"].cpu())
    assert abs(tot - L.item()) <= 1e-5 * abs(L.item())
