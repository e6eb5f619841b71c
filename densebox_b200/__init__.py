"""densebox_b200 — B200-native (sm_100a) DenseBox training / inference hot path.

Drop-in for the hot path of CaptainEven/DenseBox (DenseBox.py): the `DenseBox`, `DenseBoxLM`, `DenseBoxLMLOC`
modules, the multi-task loss of its training loops (`densebox_loss`), a fused native training step
(`DenseBoxTrainer`) and the detection post-processing (`decode_nms`).  All math runs in hand-written CUDA kernels
behind the C ABI of include/densebox_b200.h; there is no CPU fallback.
"""
from ._lib import DbxError, LIB_PATH  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA (build scripts, CPU-only tests)
    if name in ("DenseBox", "DenseBoxLM", "DenseBoxLMLOC"):
        from . import modules
        return getattr(modules, name)
    if name == "densebox_loss":
        from .loss import densebox_loss
        return densebox_loss
    if name == "DenseBoxTrainer":
        from .trainer import DenseBoxTrainer
        return DenseBoxTrainer
    if name == "NetEngine":
        from .engine import NetEngine
        return NetEngine
    if name in ("parse_label_name", "parse_label_names", "load_patch_u8", "ingest_table", "normalize_u8"):
        from . import data
        return getattr(data, name)
    if name in ("decode_nms", "perspective_transform", "perspective_matrix"):
        from . import postproc
        return getattr(postproc, name)
    raise AttributeError(name)
