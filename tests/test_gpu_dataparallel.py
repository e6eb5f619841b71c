"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): DenseBoxTrainer data parallel == one process on the
concatenated batch — same global negative quota, loss = sum over ranks, identical parameter update up to the
summation order of the gradient all-reduce (DenseBox.py:2864-2868, :2917)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    import densebox_b200
    from oracle import densebox_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B = 2
    vgg = O.seeded_vgg19(0)
    torch.manual_seed(1)
    net = densebox_b200.DenseBox(vgg).cuda()
    tr = densebox_b200.DenseBoxTrainer(net, B, lr=1e-7, dropout=False, use_cuda_graph=True,
                                       process_group=dist.group.WORLD)
    losses = []
    for step in range(3):
        g = torch.Generator().manual_seed(100 + step)
        x = torch.randn(world * B, 3, 240, 240, generator=g)
        lab = O.synth_batch(world * B, seed=step)
        rs = np.random.RandomState(step)
        rand = np.stack([rs.choice(3600, 64, replace=False) for _ in range(world * B)]).astype(np.int64)
        sl = slice(rank * B, (rank + 1) * B)
        L = tr.step(x[sl], lab["bbox"][sl], rand_neg_idx=rand[sl])
        Ls = L.detach().clone().reshape(1)
        dist.all_reduce(Ls)
        losses.append(float(Ls))
    tr.store_to_module()
    q.put((rank, losses, net.conv5_2_loc.weight.detach().cpu(), net.conv4_4_1.bias.detach().cpu()))
    dist.destroy_process_group()


def test_two_gpu_trainer_equals_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import numpy as np
    import torch.multiprocessing as mp
    import densebox_b200
    from oracle import densebox_oracle as O
    world, port = 2, 29741
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=600) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process on the concatenated batch
    B = 4
    vgg = O.seeded_vgg19(0)
    torch.manual_seed(1)
    net = densebox_b200.DenseBox(vgg).cuda()
    w0 = net.conv5_2_loc.weight.detach().cpu().clone()
    tr = densebox_b200.DenseBoxTrainer(net, B, lr=1e-7, dropout=False, use_cuda_graph=True)
    ref = []
    for step in range(3):
        g = torch.Generator().manual_seed(100 + step)
        x = torch.randn(B, 3, 240, 240, generator=g)
        lab = O.synth_batch(B, seed=step)
        rs = np.random.RandomState(step)
        rand = np.stack([rs.choice(3600, 64, replace=False) for _ in range(B)]).astype(np.int64)
        ref.append(float(tr.step(x, lab["bbox"], rand_neg_idx=rand)))
    tr.store_to_module()
    for rank, losses, w, b in res:
        for a, r in zip(losses, ref):
            assert abs(a - r) <= 1e-4 * abs(r), (losses, ref)
        d_ref = net.conv5_2_loc.weight.detach().cpu() - w0
        assert float((w - w0 - d_ref).norm() / d_ref.norm()) <= 2e-2
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])  # replicas stay in sync
