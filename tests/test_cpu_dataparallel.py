"""CPU, world_size 2 over gloo: the data-parallel contract of SURVEY.md §8(e) at the oracle level — shards use the
batch-GLOBAL positive count (1-int all-reduce, DenseBox.py:2864-2868), the loss is the SUM over ranks (:2917) and the
gradient all-reduce is a SUM; the result equals the single-process batch.  (The CUDA trainer applies the same
protocol with NCCL: densebox_b200/trainer.py.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402


def _case(B):
    g = torch.Generator().manual_seed(0)
    outs = [torch.randn(B, 1, 60, 60, generator=g), torch.randn(B, 4, 60, 60, generator=g)]
    lab = O.synth_batch(B, seed=4)
    rs = np.random.RandomState(1)
    rand = np.stack([rs.choice(3600, 128, replace=False) for _ in range(B)])
    w = torch.randn(5, generator=g)  # a shared "parameter": per-channel scale, so that gradients need an all-reduce
    return outs, lab["bbox"], rand, w


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 8
    outs, bbox, rand, w = _case(B)
    lo, hi = rank * B // world, (rank + 1) * B // world
    w = w.clone().requires_grad_(True)
    gt = O.gt_maps(bbox[lo:hi])["score"]
    pos = torch.tensor([int(np.count_nonzero(gt))])
    dist.all_reduce(pos)  # the one exchange the loss needs
    score = outs[0][lo:hi] * w[0]
    loc = outs[1][lo:hi] * w[1:].view(1, 4, 1, 1)
    L, info = O.loss((score, loc), "densebox", bbox[lo:hi], rand[lo:hi], global_pos=int(pos), global_batch=B)
    L.backward()
    g = w.grad.clone()
    dist.all_reduce(g)        # gradient SUM
    Ls = L.detach().clone()
    dist.all_reduce(Ls)
    q.put((rank, float(Ls), g.numpy(), info["half"], info["mask"]))
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_batch():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = 8
    outs, bbox, rand, w = _case(B)
    w = w.clone().requires_grad_(True)
    L, info = O.loss((outs[0] * w[0], outs[1] * w[1:].view(1, 4, 1, 1)), "densebox", bbox, rand)
    L.backward()
    for rank, Ls, g, half, mask in res:
        assert half == info["half"]
        assert abs(Ls - L.item()) <= 1e-5 * abs(L.item())
        np.testing.assert_allclose(g, w.grad.numpy(), rtol=1e-5)
    assert np.array_equal(np.concatenate([r[4] for r in res]), info["mask"])
