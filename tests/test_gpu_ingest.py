"""GPU, callers' side of the path: the batch as decoded image bytes.  ToTensor + Normalize fused into
the first kernel (dbx_net_forward_u8) must be BIT-identical to normalising with torch and calling the fp32 forward;
the trainer fed uint8 batches (a quarter of the host->device bytes) must produce the same losses as the trainer fed
the normalised fp32 batches."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def _net(variant="densebox"):
    import densebox_b200
    vgg = O.seeded_vgg19(0)
    torch.manual_seed(1)
    return getattr(densebox_b200, {"densebox": "DenseBox", "lm": "DenseBoxLM"}[variant])(vgg).cuda()


@pytest.mark.parametrize("H,W", [(240, 240), (64, 88)])
def test_fused_u8_ingest_is_bit_identical(H, W):
    from densebox_b200 import NetEngine, normalize_u8
    net = _net()
    g = torch.Generator().manual_seed(4)
    u8 = torch.randint(0, 256, (2, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    outs = []
    for x in (u8, normalize_u8(u8).contiguous()):
        eng = NetEngine("densebox", 2, H, W, train=False)
        eng.set_params(net)
        eng.forward(x)
        torch.cuda.synchronize()
        outs.append(eng.head_out().clone())
    assert torch.equal(outs[0], outs[1])
    assert outs[0].abs().max().item() > 0


def test_trainer_u8_batches_match_fp32_batches():
    from densebox_b200 import DenseBoxTrainer, normalize_u8, parse_label_names
    B = 2
    names = ["p0_label_40_60_120_92_38_58_122_60_121_93_41_90.jpg", "p1_label_60_80_140_108_59_79_141_80_142_109_61_107.jpg"]
    _, bbox, _ = parse_label_names(names, kind="lm", pin=True)
    rs = np.random.RandomState(3)
    rand = torch.tensor(np.stack([rs.choice(3600, 256, replace=False) for _ in range(B)])).pin_memory()
    g = torch.Generator().manual_seed(5)
    batches = [torch.randint(0, 256, (B, 240, 240, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(3)]
    losses = []
    for u8_mode in (True, False):
        tr = DenseBoxTrainer(_net(), B, lr=1e-7, dropout=False, input_u8=u8_mode)
        cur = []
        for i, x in enumerate(batches):
            xin = x if u8_mode else normalize_u8(x).contiguous().pin_memory()
            if u8_mode and i + 1 < len(batches):
                pass
            cur.append(tr.step(xin, bbox, rand_neg_idx=rand).item())
            if u8_mode and i + 1 < len(batches):
                tr.prefetch(batches[i + 1], bbox, rand_neg_idx=rand)
        losses.append(cur)
    assert losses[0] == losses[1], losses
