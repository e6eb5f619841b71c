"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (/root/reference/DenseBox.py).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the oracle is pinned against the reference itself:
  * geometry.npz   — init_score_map / mask_gray_zone_cls / init_lm_heatmap / mask_gray_zone_lm on 400 random labels
  * loss_*.npz     — the loop bodies train_online / train_LM_online / train_LMLOC_online / train_densebox_online driven
                     with the reference's own helper functions on random head maps (np.random.choice draws injected)
  * forward_*.npz  — DenseBox / DenseBoxLM / DenseBoxLMLOC modules of the reference on the seeded KAT input
  * decode_nms.npz — parse_out_MN / parse_DetLMLOC / parse_DetLM / NMS
  * ingest.npz     — file-name label parsers of DenseBoxDataset / LPPatchLM_Online / LPPatch_Online and the default
                     Resize + CenterCrop + ToTensor + Normalize transform on a random image
  * perspective.npz — perspective_transform (cv2 homography + warpPerspective) on a random image
  * state_dict.npz — key -> shape of the three modules' state_dict() and the outcome of a strict load_state_dict round
                     trip between the reference modules and the drop-in modules (both directions)
Loading shims: matplotlib stub, CUDA hidden during import, float64 labels for the loc-map
generators under NumPy >= 2.  Nothing of the reference is copied; only its outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("DBX_REFERENCE", "/root/reference")


def load_reference():
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF_DIR)
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    import DenseBox as REF
    torch.cuda.is_available = avail
    if cvd is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    return REF


def rand_labels(rs, B, mode):
    """mode 'quarter': 240-space integer labels / 4 (what the datasets produce); 'float': arbitrary floats."""
    if mode == "quarter":
        x0 = rs.randint(8, 150, B); y0 = rs.randint(8, 170, B)
        w = rs.randint(12, 90, B); h = rs.randint(8, 60, B)
        bbox = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float32) / np.float32(4.0)
        jit = rs.randint(-3, 4, (B, 8))
        verts = (np.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1) + jit).astype(np.float32) / np.float32(4)
    else:
        x0 = rs.uniform(3, 35, B); y0 = rs.uniform(3, 40, B); w = rs.uniform(3, 22, B); h = rs.uniform(2, 15, B)
        bbox = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float32)
        verts = (np.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1) + rs.uniform(-1, 1, (B, 8))).astype(
            np.float32)
    verts = np.clip(verts, 2.6, 56.4).astype(np.float32)  # landmarks >= 2 px inside (reference raises otherwise)
    return bbox, verts


def gen_geometry(REF):
    rs = np.random.RandomState(7)
    bb, vv, score, gray, lmheat, lmgray = [], [], [], [], [], []
    for mode, B in (("quarter", 200), ("float", 200)):
        bbox, verts = rand_labels(rs, B, mode)
        tb, tv = torch.from_numpy(bbox), torch.from_numpy(verts)
        s = REF.init_score_map(bbox=tb, batch_size=B, ratio=0.3)
        m = torch.ones(B, 1, 60, 60)
        REF.mask_gray_zone_cls(loss_mask=m, bboxes=tb, ratio=0.3, gray_border=2.0)
        hm = REF.init_lm_heatmap(vertices=tv, batch_size=B)
        lg = torch.ones(B, 4, 60, 60)
        for k in range(4):
            REF.mask_gray_zone_lm(loss_mask=lg[:, k].unsqueeze(1), pos_indices=torch.nonzero(hm[:, k].unsqueeze(1)),
                                  lm_id=k, gray_border=2.0)
        bb.append(bbox); vv.append(verts); score.append(s.numpy()); gray.append(m.numpy())
        lmheat.append(hm.numpy()); lmgray.append(lg.numpy())
    pk = lambda a: np.packbits(np.concatenate(a).astype(np.uint8).reshape(-1))
    loc = REF.init_loc_map(bboxes=torch.from_numpy(bb[0][:4]).double(), batch_size=4).numpy()
    lmloc = REF.init_lm_locmap(vertices=torch.from_numpy(vv[0][:4]).double(), batch_size=4).numpy()
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), bbox=np.concatenate(bb), vertices=np.concatenate(vv),
                        score=pk(score), gray=pk(gray), lmheat=pk(lmheat), lmgray=pk(lmgray), loc4=loc, lmloc4=lmloc)


def ref_loss(REF, variant, outs, bbox, verts, labels, rand, lm_rand, lam=(3.0, 1.0, 0.5)):
    """The reference loop bodies (DenseBox.py:2843-2918, :2575-2723, :2300-2456, :2023-2180) with the reference's own
    helpers; only np.random.choice is replaced by the injected draws."""
    lambda_loc, lambda_det, lambda_lm = lam
    B = bbox.shape[0]
    loss_func = torch.nn.MSELoss(reduction="none")
    tb, tv = torch.from_numpy(bbox), (torch.from_numpy(verts) if verts is not None else None)
    pn = labels is not None
    tl = torch.from_numpy(labels).view(B, 1) if pn else None
    if pn:
        cls_gt = REF.init_score(bboxes=tb, labels=tl, ratio=0.3)
        loc_gt = REF.init_loc(bboxes=tb.double(), labels=tl)
    else:
        cls_gt = REF.init_score_map(bbox=tb, batch_size=B, ratio=0.3)
        loc_gt = REF.init_loc_map(bboxes=tb.double(), batch_size=B)
    mask_cls = cls_gt.clone()
    if variant == "densebox":
        score, loc = outs
    elif variant == "lm":
        score, loc, lm, rf = outs
    else:
        score, rf, loc, lm, lmloc = outs
    cls_loss = loss_func(score, cls_gt)
    loc_loss = loss_func(loc, loc_gt)
    pos_indices = torch.nonzero(cls_gt)
    neg_num = int(float(pos_indices.size(0)) / float(B) + 0.5)
    half = int(neg_num * 0.5 + 0.5)
    neg_cls = (cls_loss * (torch.ones(B, 1, 60, 60) - cls_gt)).view(B, -1)
    _, hard = torch.topk(input=neg_cls, k=half, dim=1)
    neg_indices = torch.cat((hard, torch.from_numpy(rand[:, :half]).long()), dim=1)
    REF.mask_by_sel(loss_mask=mask_cls, pos_indices=pos_indices, neg_indices=neg_indices)
    if pn:
        REF.mask_gray_zone_cls_pn(loss_mask=mask_cls, bboxes=tb, labels=tl, ratio=0.3, gray_border=2.0)
    else:
        REF.mask_gray_zone_cls(loss_mask=mask_cls, bboxes=tb, ratio=0.3, gray_border=2.0)
    out = {"half": half, "pos": pos_indices.size(0), "mask": mask_cls.numpy().copy()}
    if variant == "densebox":
        loss = torch.sum(mask_cls * cls_loss) + torch.sum(lambda_loc * (mask_cls * cls_gt * loc_loss))
        out["loss"] = loss
        return out
    lm_gt = REF.init_lm_heatmap_pn(vertices=tv, labels=tl) if pn else REF.init_lm_heatmap(vertices=tv, batch_size=B)
    mask_lm = lm_gt.clone()
    lm_loss = loss_func(lm, lm_gt)
    rf_loss = loss_func(rf, cls_gt)
    for k in range(4):
        gt_k = lm_gt[:, k].unsqueeze(1)
        mask_k = mask_lm[:, k].unsqueeze(1)
        neg_k = REF.gen_neg_loss(loss_orig=lm_loss[:, k].unsqueeze(1), map_gt=gt_k).view(B, -1)
        pos_k = torch.nonzero(gt_k)
        _, hard_k = torch.topk(input=neg_k, k=1, dim=1)
        neg_k_idx = torch.cat((hard_k, torch.from_numpy(lm_rand[:, k:k + 1]).long()), dim=1)
        REF.mask_by_sel(loss_mask=mask_k, pos_indices=pos_k, neg_indices=neg_k_idx)
        REF.mask_gray_zone_lm(loss_mask=mask_k, pos_indices=pos_k, lm_id=k, gray_border=2.0)
    det = lambda_det * (torch.sum(mask_cls * cls_loss) + lambda_loc * torch.sum(mask_cls * cls_gt * loc_loss))
    lml = lambda_lm * torch.sum(mask_lm * lm_loss)
    if variant == "lmloc":
        lmloc_gt = (REF.init_lm_locmap_pn(vertices=tv.double(), labels=tl) if pn
                    else REF.init_lm_locmap(vertices=tv.double(), batch_size=B))
        lml = lml + torch.sum(mask_cls * cls_gt * loss_func(lmloc, lmloc_gt))
    out["loss"] = det + lml + torch.sum(mask_cls * rf_loss)
    out["lm_mask"] = mask_lm.numpy().copy()
    return out


def gen_loss(REF):
    chans = {"densebox": [1, 4], "lm": [1, 4, 4, 1], "lmloc": [1, 1, 4, 4, 8]}
    for case, variant, pn in (("densebox", "densebox", False), ("lm", "lm", False), ("lmloc", "lmloc", False),
                              ("lmloc_pn", "lmloc", True)):
        rs = np.random.RandomState(11)
        B = 4
        bbox, verts = rand_labels(rs, B, "quarter")
        labels = np.array([1, 0, 1, 1], np.float32) if pn else None
        g = torch.Generator().manual_seed(21)
        outs = [(torch.randn(B, c, 60, 60, generator=g) * 0.7).half().float().requires_grad_(True) for c in chans[variant]]
        rand = np.stack([rs.choice(3600, 256, replace=False) for _ in range(B)]).astype(np.int64)
        lm_rand = rs.randint(0, 3600, (B, 4)).astype(np.int64)
        r = ref_loss(REF, variant, outs, bbox, verts if variant != "densebox" else None, labels, rand, lm_rand)
        r["loss"].backward()
        d = {"bbox": bbox, "vertices": verts, "rand": rand, "lm_rand": lm_rand, "loss": np.float32(r["loss"].item()),
             "half": r["half"], "pos": r["pos"], "mask": np.packbits(r["mask"].astype(np.uint8).reshape(-1))}
        if labels is not None:
            d["labels"] = labels
        if "lm_mask" in r:
            d["lm_mask"] = np.packbits(r["lm_mask"].astype(np.uint8).reshape(-1))
        for i, o in enumerate(outs):
            d["out%d" % i] = o.detach().numpy().astype(np.float16)  # inputs stored as fp16 (exactly representable)
            d["grad%d" % i] = o.grad.numpy()[:, :, ::5, ::5].copy()
        np.savez_compressed(os.path.join(HERE, "loss_%s.npz" % case), **d)


def gen_forward(REF):
    import torchvision
    for variant, cls in (("densebox", "DenseBox"), ("lm", "DenseBoxLM"), ("lmloc", "DenseBoxLMLOC")):
        torch.manual_seed(0)
        vgg = torchvision.models.vgg19(weights=None)
        torch.manual_seed(1)
        net = getattr(REF, cls)(vgg).eval()
        x = torch.randn(1, 3, 240, 240, generator=torch.Generator().manual_seed(2))
        outs = net(x)
        d = {"n_keys": len(net.state_dict()), "keys": np.array(sorted(net.state_dict().keys()))}
        for i, o in enumerate(outs):
            d["sum%d" % i] = np.float64(o.double().sum().item())
            d["sample%d" % i] = o.detach().numpy()[:, :, ::6, ::6].copy()
        if variant == "densebox":  # the BASELINE.md plumbing KAT
            bbox = np.array([[20, 25, 40, 32]], np.float32)
            rand = np.random.RandomState(3).choice(3600, 64, replace=False)[None].astype(np.int64)
            r = ref_loss(REF, variant, outs, bbox, None, None, rand, None)
            r["loss"].backward()
            d.update(kat_loss=np.float32(r["loss"].item()), kat_half=r["half"], kat_mask_nnz=int(r["mask"].sum()),
                     kat_gnorm_conv1_1=np.float32(net.conv1_1_1.weight.grad.norm().item()),
                     kat_gnorm_conv5_2_det=np.float32(net.conv5_2_det.weight.grad.norm().item()),
                     kat_conv3_3_grad_none=net.conv3_3_1.weight.grad is None)
        np.savez_compressed(os.path.join(HERE, "forward_%s.npz" % variant), **d)


def gen_decode(REF):
    g = torch.Generator().manual_seed(31)
    M, N = 64, 96
    score = torch.randn(1, 1, M // 4, N // 4, generator=g)
    loc = torch.randn(1, 4, M // 4, N // 4, generator=g) * 5
    lmh = torch.randn(1, 4, M // 4, N // 4, generator=g)
    lml = torch.randn(1, 8, M // 4, N // 4, generator=g) * 5
    d1 = REF.parse_out_MN(score, loc, M, N, K=10)
    d2 = REF.parse_DetLMLOC(score, loc, lmh, lml, M, N, K=10)
    d3 = REF.parse_DetLM(score, loc, lmh, M, N, K=10)
    rs = np.random.RandomState(5)
    boxes = []
    for _ in range(40):
        x, y = rs.uniform(0, 200, 2); w, h = rs.uniform(10, 80, 2)
        boxes.append([x, y, x + w, y + h, rs.uniform(0, 1)])
    boxes = np.array(boxes)
    keep = [REF.NMS(boxes, t) for t in (0.2, 0.4, 0.6)]
    kat = REF.NMS(np.array([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [50, 50, 60, 60, .7]]), 0.4)
    np.savez_compressed(os.path.join(HERE, "decode_nms.npz"), score=score.numpy(), loc=loc.numpy(), lmh=lmh.numpy(),
                        lml=lml.numpy(), dets_mn=np.asarray(d1), dets_lmloc=np.asarray(d2), dets_lm=np.asarray(d3), boxes=boxes,
                        keep02=np.array(keep[0]), keep04=np.array(keep[1]), keep06=np.array(keep[2]),
                        kat_keep=np.array(kat))


def gen_state_dict(REF):
    """Checkpoint compatibility (DenseBox.py:1938-1945, :1989-1994): the drop-in modules must load a reference
    checkpoint and save one the reference loads — strict, both ways — and carry the same values afterwards."""
    import torchvision
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import densebox_b200
    d = {}
    for variant, cls in (("densebox", "DenseBox"), ("lm", "DenseBoxLM"), ("lmloc", "DenseBoxLMLOC")):
        torch.manual_seed(0)
        vgg = torchvision.models.vgg19(weights=None)
        torch.manual_seed(1)
        ref = getattr(REF, cls)(vgg)
        torch.manual_seed(2)
        ours = getattr(densebox_b200, cls)(vgg)
        sd_ref = ref.state_dict()
        r1 = ours.load_state_dict(sd_ref, strict=True)          # reference checkpoint -> drop-in module
        same1 = all(torch.equal(v, sd_ref[k]) for k, v in ours.state_dict().items())
        torch.manual_seed(3)
        ours2 = getattr(densebox_b200, cls)(vgg)
        r2 = ref.load_state_dict(ours2.state_dict(), strict=True)  # drop-in checkpoint -> reference module
        same2 = all(torch.equal(v, ours2.state_dict()[k]) for k, v in ref.state_dict().items())
        d[variant + "_keys"] = np.array(list(sd_ref.keys()))
        d[variant + "_shapes"] = np.array([",".join(map(str, v.shape)) for v in sd_ref.values()])
        d[variant + "_roundtrip_ok"] = np.array([not r1.missing_keys and not r1.unexpected_keys and same1,
                                                 not r2.missing_keys and not r2.unexpected_keys and same2])
    np.savez_compressed(os.path.join(HERE, "state_dict.npz"), **d)


def gen_ingest(REF):
    """Callers' side (f-3): the file-name label parsers of the three online datasets and the default transform
    (Resize + CenterCrop to the stored size, ToTensor, Normalize) of DenseBoxDataset on a random RGB image."""
    import tempfile
    from PIL import Image
    rs = np.random.RandomState(13)
    names12 = []
    for i in range(24):
        v = rs.randint(1, 240, 12)
        names12.append("img%03d_label_%s.jpg" % (i, "_".join(str(int(x)) for x in v)))
    names12 += ["neg_7_label_0_0_0_0_0_0_0_0_0_0_0_0.jpg", "a_b_label_1_2_3_4_5_6_7_8_9_10_11_12_x.png"]
    names4 = ["p%02d_label_%d_%d_%d_%d.jpg" % (i, *rs.randint(1, 240, 4)) for i in range(12)]
    d = {}
    with tempfile.TemporaryDirectory() as root:
        r12, r4 = os.path.join(root, "a"), os.path.join(root, "b")
        os.makedirs(r12); os.makedirs(r4)
        for n in names12:
            open(os.path.join(r12, n), "wb").close()
        for n in names4:
            open(os.path.join(r4, n), "wb").close()
        ds = REF.DenseBoxDataset(r12, transform=None, size=(48, 64))
        d["db_names"] = np.array([os.path.split(p)[1] for p in ds.imgs_path])
        d["db_bbox"] = torch.stack(ds.bboxes).numpy(); d["db_vertices"] = torch.stack(ds.vertices).numpy()
        d["db_labels"] = torch.stack(ds.labels).numpy().reshape(-1)
        lm_names = [n for n in names12 if not n.startswith("neg")]
        r12b = os.path.join(root, "c"); os.makedirs(r12b)
        for n in lm_names:
            open(os.path.join(r12b, n), "wb").close()
        dl = REF.LPPatchLM_Online(r12b, transform=None, size=(48, 64))
        d["lm_names"] = np.array([os.path.split(p)[1] for p in dl.imgs_path])
        d["lm_bbox"] = torch.stack(dl.bboxes).numpy(); d["lm_vertices"] = torch.stack(dl.vertices).numpy()
        d4 = REF.LPPatch_Online(r4, transform=None, size=(48, 64))
        d["b_names"] = np.array([os.path.split(p)[1] for p in d4.imgs_path])
        d["b_bbox"] = torch.stack(d4.labels).numpy()
        img = rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)
        img[0, 0] = (0, 0, 0); img[0, 1] = (255, 255, 255)
        d["img_u8"] = img
        d["img_norm"] = ds.transform(Image.fromarray(img)).numpy()   # the reference's own Compose
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **d)


def gen_perspective(REF):
    """perspective_transform (DenseBox.py:3446-3481) of the reference (cv2.getPerspectiveTransform + warpPerspective)
    on a random RGB image for three landmark quadrilaterals, plus the homographies cv2 computes."""
    import cv2
    rs = np.random.RandomState(21)
    img = rs.randint(0, 256, (96, 128, 3)).astype(np.uint8)
    cases = [[[30, 20], [100, 26], [98, 60], [28, 52]], [[10.5, 8.25], [90, 5], [95.5, 70], [12, 66]],
             [[40, 40], [80, 38], [84, 62], [37, 66]]]
    outs = [REF.perspective_transform(img, c) for c in cases]
    mats = []
    for c in cases:
        s = np.float32(c)
        lu, ru, rd, ld = s
        mnx, mxx, mny, mxy = min(lu[0], ld[0]), max(ru[0], rd[0]), min(lu[1], ru[1]), max(ld[1], rd[1])
        mats.append(cv2.getPerspectiveTransform(s, np.float32([[mnx, mny], [mxx, mny], [mxx, mxy], [mnx, mxy]])))
    np.savez_compressed(os.path.join(HERE, "perspective.npz"), img=img, pts=np.array(cases, np.float64), out0=outs[0],
                        out1=outs[1], out2=outs[2], mats=np.array(mats))


if __name__ == "__main__":
    REF = load_reference()
    gen_perspective(REF)
    gen_ingest(REF)
    gen_state_dict(REF)
    gen_geometry(REF)
    gen_loss(REF)
    gen_decode(REF)
    gen_forward(REF)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
