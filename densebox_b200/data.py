"""Callers' side of the hot path: where a batch comes from.

  * `parse_label_name` / `parse_label_names` — the file-name label convention of the reference's datasets
    (`DenseBoxDataset` DenseBox.py:787-850, `LPPatchLM_Online` :928-972, `LPPatch_Online` :1038-1049):
    `<anything>_label_x0_y0_x1_y1[_v0x_v0y_..._v3x_v3y]<anything>` in 240-space integers, divided by 4.0 into the 60x60
    output space; twelve zeros = a pure-negative patch (label 0, :809-816).
  * `load_patch_u8` — `Image.open` + gray -> RGB (:862-868) as a uint8 HWC array: what leaves the host.
  * `ingest_table` — `ToTensor()` + `Normalize(mean, std)` (:766-772, :3609-3621) as a 3 x 256 fp32 table, computed with
    the same float32 operations torchvision performs ((u / 255 - mean) / std), so the fused on-GPU ingest
    (`dbx_net_forward_u8`: table lookup inside the im2col kernel of conv1_1) is bit-identical to the reference pipeline
    followed by the fp32 forward — and the host->device copy of a batch is 4x smaller (uint8 HWC vs fp32 CHW).
    `Resize(size)` / `CenterCrop(size)` are identities in every use the reference makes of them (patches are stored at
    240 x 240, `test_*` pass the image's own size), so they are asserted, not implemented.
"""
import os
import re

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_P12 = re.compile(".*_label_" + "_".join(["([0-9]+)"] * 12) + ".*")
_P4 = re.compile(".*_label_([0-9]+)_([0-9]+)_([0-9]+)_([0-9]+)")


def parse_label_name(name, kind="densebox"):
    """-> (label, bbox[4], vertices[8]) as python floats in 60-space.

    kind "densebox": `DenseBoxDataset` (12 numbers; all zero -> label 0 with zero boxes), "lm": `LPPatchLM_Online`
    (12 numbers, always positive), "bbox": `LPPatch_Online` (4 numbers, no vertices -> zeros).  A name that does not
    match raises ValueError (the reference prints the name and then fails on the unbound match, :793-808)."""
    name = os.path.split(name)[1]
    if kind == "bbox":
        m = _P4.match(name)
        if m is None:
            raise ValueError("no _label_x0_y0_x1_y1 in %r" % name)
        v = [float(g) for g in m.groups()]
        return 1.0, [x / 4.0 for x in v], [0.0] * 8
    m = _P12.match(name)
    if m is None:
        raise ValueError("no 12-number _label_ field in %r" % name)
    v = [float(g) for g in m.groups()]
    if kind == "densebox" and all(x == 0.0 for x in v):
        return 0.0, [0.0] * 4, [0.0] * 8
    return 1.0, [x / 4.0 for x in v[:4]], [x / 4.0 for x in v[4:]]


def parse_label_names(names, kind="densebox", pin=False):
    """A batch of file names -> (labels [B], bbox [B,4], vertices [B,8]) fp32 tensors (pinned on request): the label
    arguments of `DenseBoxTrainer.step` / `densebox_loss`."""
    rows = [parse_label_name(n, kind) for n in names]
    labels = torch.tensor([r[0] for r in rows], dtype=torch.float32)
    bbox = torch.tensor(np.asarray([r[1] for r in rows], dtype=np.float64).astype(np.float32))
    verts = torch.tensor(np.asarray([r[2] for r in rows], dtype=np.float64).astype(np.float32))
    if pin:
        labels, bbox, verts = labels.pin_memory(), bbox.pin_memory(), verts.pin_memory()
    return labels, bbox, verts


def load_patch_u8(path, size=None):
    """`Image.open(path)`, gray -> RGB (:862-868) -> uint8 array [H,W,3].  size=(H,W): asserted (Resize + CenterCrop to
    the stored size are identities in the reference's use)."""
    from PIL import Image
    img = Image.open(path)
    if img.mode != "RGB":
        img = img.convert("RGB")
    a = np.asarray(img, dtype=np.uint8)
    if size is not None and tuple(a.shape[:2]) != tuple(size):
        raise ValueError("patch %s is %s, expected %s (the reference stores patches at their training size)"
                         % (path, a.shape[:2], tuple(size)))
    return a


def ingest_table(mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """fp32 [3,256]: table[c][u] = (u / 255 - mean[c]) / std[c] in the float32 arithmetic of ToTensor + Normalize."""
    u = torch.arange(256, dtype=torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1)
    return ((u.view(1, 256) - m) / s).contiguous()


def normalize_u8(x_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """uint8 [N,H,W,3] -> fp32 [N,3,H,W] exactly as ToTensor + Normalize do (a torch op on the tensor's device; the
    reference-format input of `net.forward`).  The fused path never materialises this tensor."""
    t = x_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32, device=x_u8.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=x_u8.device).view(1, 3, 1, 1)
    return (t - m) / s
