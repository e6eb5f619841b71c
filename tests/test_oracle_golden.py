"""CPU: the oracle restatement (oracle/densebox_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container).  Integer work (boxes, masks, quotas, NMS) is bit-exact;
floating point: loss 1e-6 relative, forward samples 1e-4 absolute (same torch, same ops, different call path)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a, shape):
    return np.unpackbits(a)[:int(np.prod(shape))].reshape(shape)


def test_geometry_bit_exact():
    d = np.load(os.path.join(G, "geometry.npz"))
    bbox, verts = d["bbox"], d["vertices"]
    B = bbox.shape[0]
    gts = O.gt_maps(bbox, verts)
    assert np.array_equal(gts["score"].astype(np.uint8), bits(d["score"], (B, 1, 60, 60)))
    assert np.array_equal(gts["lm"].astype(np.uint8), bits(d["lmheat"], (B, 4, 60, 60)))
    gray = np.ones((B, 1, 60, 60), np.float32)
    for b in range(B):
        (gx, Gx, gy, Gy), (ix, Ix, iy, Iy) = O.gray_box(bbox[b])
        ya, yb = O._pyslice(gy, Gy, 60); xa, xb = O._pyslice(gx, Gx, 60)
        gray[b, 0, ya:yb, xa:xb] = 0
        ya, yb = O._pyslice(iy, Iy, 60); xa, xb = O._pyslice(ix, Ix, 60)
        gray[b, 0, ya:yb, xa:xb] = 1
    assert np.array_equal(gray.astype(np.uint8), bits(d["gray"], (B, 1, 60, 60)))
    lmg = np.ones((B, 4, 60, 60), np.float32)
    for b in range(B):
        for k, (x, y) in enumerate(O.lm_points(verts[b])):
            ya, yb = O._pyslice(y - 2, y + 3, 60); xa, xb = O._pyslice(x - 2, x + 3, 60)
            lmg[b, k, ya:yb, xa:xb] = 0
            lmg[b, k, y, x] = 1
    assert np.array_equal(lmg.astype(np.uint8), bits(d["lmgray"], (B, 4, 60, 60)))
    g4 = O.gt_maps(bbox[:4], verts[:4])
    assert np.array_equal(g4["loc"], d["loc4"]) and np.array_equal(g4["lmloc"], d["lmloc4"])


@pytest.mark.parametrize("case", ["densebox", "lm", "lmloc", "lmloc_pn"])
def test_loss_against_reference_loop_body(case):
    d = np.load(os.path.join(G, "loss_%s.npz" % case))
    variant = case.split("_")[0]
    n_out = {"densebox": 2, "lm": 4, "lmloc": 5}[variant]
    outs = [torch.from_numpy(d["out%d" % i].astype(np.float32)).requires_grad_(True) for i in range(n_out)]
    labels = d["labels"] if "labels" in d.files else None
    L, info = O.loss(tuple(outs), variant, d["bbox"], d["rand"], vertices=d["vertices"] if variant != "densebox" else None,
                     lm_rand_idx=d["lm_rand"], labels=labels)
    B = d["bbox"].shape[0]
    assert info["half"] == int(d["half"]) and info["pos"] == int(d["pos"])
    assert np.array_equal(info["mask"].astype(np.uint8), bits(d["mask"], (B, 1, 60, 60)))
    if variant != "densebox":
        assert np.array_equal(info["lm_mask"].astype(np.uint8), bits(d["lm_mask"], (B, 4, 60, 60)))
    assert abs(L.item() - float(d["loss"])) <= 1e-6 * abs(float(d["loss"]))
    L.backward()
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.grad.numpy()[:, :, ::5, ::5], d["grad%d" % i], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("variant", ["densebox", "lm", "lmloc"])
def test_forward_against_reference_modules(variant):
    d = np.load(os.path.join(G, "forward_%s.npz" % variant))
    vgg = O.seeded_vgg19(0)
    P = O.params_from_vgg(vgg, variant, seed_heads=1)
    x = torch.randn(1, 3, 240, 240, generator=torch.Generator().manual_seed(2))
    if variant == "densebox":
        for v in P.values():
            v.requires_grad_(True)
    outs = O.forward(P, x, variant)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().numpy()[:, :, ::6, ::6], d["sample%d" % i], rtol=1e-4, atol=1e-4)
        assert abs(o.double().sum().item() - float(d["sum%d" % i])) <= 1e-3 * max(1.0, abs(float(d["sum%d" % i])))
    if variant == "densebox":  # the BASELINE.md plumbing known-answer test
        bbox = np.array([[20, 25, 40, 32]], np.float32)
        rand = np.random.RandomState(3).choice(3600, 64, replace=False)[None]
        L, info = O.loss(outs, variant, bbox, rand)
        L.backward()
        assert info["half"] == int(d["kat_half"]) == 11
        assert int(info["mask"].sum()) == int(d["kat_mask_nnz"]) == 43
        assert abs(L.item() - float(d["kat_loss"])) <= 1e-5 * float(d["kat_loss"])
        assert abs(L.item() - 14466.071289) <= 0.05  # BASELINE.md value
        assert abs(P["conv1_1.weight"].grad.norm().item() - float(d["kat_gnorm_conv1_1"])) <= 1e-3 * float(d["kat_gnorm_conv1_1"])
        assert abs(P["conv5_2_det.weight"].grad.norm().item() - float(d["kat_gnorm_conv5_2_det"])) <= 1e-3 * float(d["kat_gnorm_conv5_2_det"])
        assert P["conv3_3.weight"].grad is None and bool(d["kat_conv3_3_grad_none"])


def test_decode_and_nms_bit_exact():
    d = np.load(os.path.join(G, "decode_nms.npz"))
    t = lambda k: torch.from_numpy(d[k])
    assert np.array_equal(O.decode(t("score"), t("loc"), None, K=10), d["dets_mn"])
    assert np.array_equal(O.decode(t("score"), t("loc"), t("lml"), K=10), d["dets_lmloc"])
    assert np.array_equal(O.decode(t("score"), t("loc"), None, K=10, lmheat_map=t("lmh")), d["dets_lm"])
    for key, th in (("keep02", 0.2), ("keep04", 0.4), ("keep06", 0.6)):
        assert O.nms(d["boxes"], th) == d[key].tolist()
    kat = np.array([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [50, 50, 60, 60, .7]])
    assert O.nms(kat, 0.4) == d["kat_keep"].tolist() == [0, 2]


def test_state_dict_keys_match_reference():
    """Drop-in boundary: the module classes expose exactly the reference's state_dict keys (64 / 78 / 86)."""
    import densebox_b200
    vgg = O.seeded_vgg19(0)
    for variant, cls, n in (("densebox", "DenseBox", 64), ("lm", "DenseBoxLM", 78), ("lmloc", "DenseBoxLMLOC", 86)):
        d = np.load(os.path.join(G, "forward_%s.npz" % variant))
        net = getattr(densebox_b200, cls)(vgg)
        keys = sorted(net.state_dict().keys())
        assert len(keys) == n == int(d["n_keys"])
        assert keys == [str(k) for k in d["keys"]]


def test_state_dict_shapes_and_checkpoint_round_trip():
    """Checkpoint compatibility (DenseBox.py:1938-1945, :1989-1994): same keys IN THE SAME ORDER with the same shapes
    as the reference's state_dict(); make_golden.py loaded a reference checkpoint into the drop-in modules and a
    drop-in checkpoint into the reference modules (strict, values compared) and stored the outcome.  When the
    reference is present (the build container) the round trip is repeated live."""
    import densebox_b200
    d = np.load(os.path.join(G, "state_dict.npz"))
    vgg = O.seeded_vgg19(0)
    for variant, cls in (("densebox", "DenseBox"), ("lm", "DenseBoxLM"), ("lmloc", "DenseBoxLMLOC")):
        assert d[variant + "_roundtrip_ok"].all()
        sd = getattr(densebox_b200, cls)(vgg).state_dict()
        assert list(sd.keys()) == [str(k) for k in d[variant + "_keys"]]
        assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(x) for x in d[variant + "_shapes"]]
    if not os.path.isdir("/root/reference"):
        return
    sys.path.insert(0, G)
    import make_golden
    REF = make_golden.load_reference()
    torch.manual_seed(5)
    ref = REF.DenseBoxLM(vgg)
    torch.manual_seed(6)
    ours = densebox_b200.DenseBoxLM(vgg)
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(1, 3, 240, 240, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = ref.eval()(x)
        got = O.forward(O.params_from_state_dict(ours.state_dict(), "lm"), x, "lm")
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-4)


def test_label_parser_and_ingest_table_match_reference_datasets():
    """f-3, callers' side: densebox_b200.data against the reference's datasets (fixture written by make_golden.py from
    DenseBoxDataset / LPPatchLM_Online / LPPatch_Online and DenseBoxDataset's default transform): labels bit-exact,
    the ToTensor + Normalize table bit-exact on every byte value that occurs."""
    from densebox_b200 import data
    d = np.load(os.path.join(G, "ingest.npz"))
    lab, bbox, verts = data.parse_label_names([str(n) for n in d["db_names"]], kind="densebox")
    assert np.array_equal(lab.numpy(), d["db_labels"]) and (lab == 0).sum() == 1
    assert np.array_equal(bbox.numpy(), d["db_bbox"]) and np.array_equal(verts.numpy(), d["db_vertices"])
    _, bbox, verts = data.parse_label_names([str(n) for n in d["lm_names"]], kind="lm")
    assert np.array_equal(bbox.numpy(), d["lm_bbox"]) and np.array_equal(verts.numpy(), d["lm_vertices"])
    _, bbox, _ = data.parse_label_names([str(n) for n in d["b_names"]], kind="bbox")
    assert np.array_equal(bbox.numpy(), d["b_bbox"])
    with pytest.raises(ValueError):
        data.parse_label_name("no_label_here.jpg")
    u8 = torch.from_numpy(d["img_u8"])
    tab = data.ingest_table()
    via_table = torch.stack([tab[c][u8[..., c].long()] for c in range(3)])
    assert torch.equal(via_table, torch.from_numpy(d["img_norm"]))
    assert torch.equal(data.normalize_u8(u8[None])[0], torch.from_numpy(d["img_norm"]))


def test_perspective_matrix_matches_cv2():
    """f-4: the homography of perspective_transform (DenseBox.py:3455-3476) against cv2.getPerspectiveTransform
    (fixture): 1e-12 relative (two LU solves of the same 8 x 8 system)."""
    from densebox_b200.postproc import perspective_matrix
    d = np.load(os.path.join(G, "perspective.npz"))
    for pts, want in zip(d["pts"], d["mats"]):
        M, Minv = perspective_matrix(pts)
        assert np.abs(M - want).max() <= 1e-12 * np.abs(want).max()
        assert np.abs(M @ Minv - np.eye(3)).max() <= 1e-9
