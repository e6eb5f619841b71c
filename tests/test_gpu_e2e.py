"""GPU parity of the whole path (forward -> loss -> backward -> SGD) against the CPU oracle, through the public API.

Tolerances (bf16 storage of activations/weights, fp32 accumulation; the oracle is fed the SAME bf16-rounded weights
and inputs so only activation rounding and summation order differ):
  * head maps:            max|err| <= 3e-2 * max|ref|
  * loss:                 |L - L_ref| / |L_ref| <= 1e-3          (BASELINE.json: "loss match <= 1e-3")
  * masks / quotas:       bit-exact (integer work)
  * parameter gradients vs the fp32 oracle: ||g - g_ref|| / ||g_ref|| <= 0.35 per tensor.  This is NOT an
    implementation slack: storing activations in bf16 moves ~0.4 % of the pre-activations across the ReLU threshold
    (and a similar share of 2x2 pool arg-maxes) per layer, each flip re-routes a full-size gradient element, so the
    error grows like sqrt(layers * flipped fraction): ~3 % at conv4_4, ~20 % at conv1_1 (measured).
  * parameter gradients vs the oracle run with bf16-STORAGE emulation (same rounding points as the engine):
    <= 2e-2 for the head/refine tensors, <= 0.15 for the backbone.  bf16 storage makes the backward chaotic: the
    emulated oracle perturbed by 2e-7 (relative, on the input) differs from ITSELF by 0.6 % (conv4_4) .. 9.6 %
    (conv1_1) — measured on CPU, see DESIGN.md "Gradient parity" — so tighter bounds on early layers test noise.
    The kernels of the backward are checked tightly one by one in test_gpu_gemm.py / test_gpu_elementwise.py.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densebox_oracle as O  # noqa: E402  (tests may import the oracle — as the checker only)

pytestmark = pytest.mark.gpu

CLS = {"densebox": "DenseBox", "lm": "DenseBoxLM", "lmloc": "DenseBoxLMLOC"}


def build(variant, seed_heads=1):
    import densebox_b200
    vgg = O.seeded_vgg19(0)
    torch.manual_seed(seed_heads)
    net = getattr(densebox_b200, CLS[variant])(vgg)
    return vgg, net


def oracle_params(net, variant):
    P = O.params_from_state_dict(net.state_dict(), variant)
    for k in P:  # same rounding as the engine: bf16 weights, fp32 biases
        if k.endswith(".weight"):
            P[k] = P[k].bfloat16().float()
    return P


def make_inputs(B, variant, seed=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 240, 240, generator=g).bfloat16().float()
    lab = O.synth_batch(B, seed=0, with_vertices=variant != "densebox")
    rs = np.random.RandomState(3)
    rand = np.stack([rs.choice(3600, 64, replace=False) for _ in range(B)]).astype(np.int64)
    lm_rand = rs.randint(0, 3600, (B, 4)).astype(np.int64)
    return x, lab, rand, lm_rand


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("variant", ["densebox", "lm", "lmloc"])
def test_forward_loss_backward_eval(variant):
    from densebox_b200 import densebox_loss
    B = 2
    vgg, net = build(variant)
    net = net.cuda().eval()
    x, lab, rand, lm_rand = make_inputs(B, variant)
    # ---- oracle
    P = oracle_params(net, variant)
    for v in P.values():
        v.requires_grad_(True)
    outs_ref = O.forward(P, x, variant)
    verts = lab.get("vertices")
    L_ref, info_ref = O.loss(outs_ref, variant, lab["bbox"], rand, vertices=verts, lm_rand_idx=lm_rand)
    L_ref.backward()
    g_fp32 = {k: v.grad.clone() for k, v in P.items() if v.grad is not None}
    for v in P.values():
        v.grad = None
    L_em, _ = O.loss(O.forward(P, x, variant, emulate_bf16_storage=True), variant, lab["bbox"], rand, vertices=verts,
                     lm_rand_idx=lm_rand)
    L_em.backward()
    g_em = {k: v.grad.clone() for k, v in P.items() if v.grad is not None}
    # ---- CUDA path through the drop-in API
    outs = net(x.cuda())
    for o, r in zip(outs, outs_ref):
        assert o.shape == r.shape
        err = (o.detach().cpu() - r.detach()).abs().max().item()
        assert err <= 3e-2 * r.detach().abs().max().item() + 1e-6, (variant, err, r.abs().max().item())
    kw = {}
    if variant == "lm":
        score, loc, lm, rf = outs
        kw = dict(lm=lm, rf=rf, vertices=verts, lm_rand_neg_idx=lm_rand)
    elif variant == "lmloc":
        score, rf, loc, lm, lmloc = outs
        kw = dict(lm=lm, rf=rf, lm_loc=lmloc, vertices=verts, lm_rand_neg_idx=lm_rand)
    else:
        score, loc = outs
    L, info = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand, return_info=True, **kw)
    assert info["half"] == info_ref["half"] and info["pos"] == info_ref["pos"]
    mask = info["mask"].cpu().numpy().reshape(B, 1, 60, 60)
    # hard negatives are picked from the network's own scores: identical unless two candidates are within rounding
    mism = int((mask != info_ref["mask"].astype(np.uint8)).sum())
    assert mism <= 2 * B, ("mask mismatch", mism)
    lrel = abs(L.item() - L_ref.item()) / abs(L_ref.item())
    assert lrel <= 1e-3, (variant, L.item(), L_ref.item(), lrel)
    L.backward()
    e32, eem = {}, {}
    for name in [n[:-7] for n in P if n.endswith(".weight") and n != "conv3_3.weight"]:
        w, b = net._wb(name)
        for kind, t in ((".weight", w), (".bias", b)):
            e32[name + kind] = rel(t.grad.cpu(), g_fp32[name + kind])
            eem[name + kind] = rel(t.grad.cpu(), g_em[name + kind])
    print(variant, "grad rel err vs fp32 oracle: max %.3f; vs bf16-storage emulation: max %.4f (%s)" % (
        max(e32.values()), max(eem.values()), max(eem, key=eem.get)), flush=True)
    bad = {k: round(v, 4) for k, v in e32.items() if not v <= 0.35}
    assert not bad, (variant, "vs fp32 oracle", bad)
    bad = {k: round(v, 4) for k, v in eem.items() if not v <= (2e-2 if k.startswith(("conv5", "conv6")) else 0.15)}
    assert not bad, (variant, "vs bf16-storage emulation", bad)
    assert net.conv3_3_1.weight.grad is None  # conv3_3 is never run (DenseBox.py:193-195)


def test_train_mode_dropout_injected():
    """train(): inject the oracle's {0,2} dropout masks (DenseBox.py:160,176) and compare loss + head gradients."""
    from densebox_b200 import densebox_loss
    variant, B = "densebox", 2
    vgg, net = build(variant)
    net = net.cuda().train()
    x, lab, rand, _ = make_inputs(B, variant)
    g = torch.Generator().manual_seed(5)
    drop = {h: (torch.rand(B, 512, 60, 60, generator=g) < 0.5).float() * 2 for h in ("det", "loc")}
    net.dropout_mask = drop
    P = oracle_params(net, variant)
    for v in P.values():
        v.requires_grad_(True)
    outs_ref = O.forward(P, x, variant, dropout=drop)
    L_ref, _ = O.loss(outs_ref, variant, lab["bbox"], rand)
    L_ref.backward()
    score, loc = net(x.cuda())
    L = densebox_loss(score, loc, lab["bbox"], rand_neg_idx=rand)
    assert abs(L.item() - L_ref.item()) / abs(L_ref.item()) <= 1e-3
    L.backward()
    assert rel(net.conv5_1_loc.weight.grad.cpu(), P["conv5_1_loc.weight"].grad) <= 3e-2
    assert rel(net.conv4_4_1.weight.grad.cpu(), P["conv4_4.weight"].grad) <= 6e-2


def test_trainer_matches_torch_sgd():
    """Three native training steps == three steps of oracle forward/loss + torch.optim.SGD (DenseBox.py:2821-2926)."""
    from densebox_b200 import DenseBoxTrainer
    variant, B = "densebox", 2
    vgg, net = build(variant)
    net = net.cuda()
    lr = 1e-7
    tr = DenseBoxTrainer(net, B, lr=lr, dropout=False, use_cuda_graph=True)
    P = O.params_from_state_dict(net.state_dict(), variant)
    w0 = {k: v.clone() for k, v in P.items()}
    for v in P.values():
        v.requires_grad_(True)
    opt = torch.optim.SGD(list(P.values()), lr=lr, momentum=0.9, weight_decay=5e-8)
    losses, losses_ref = [], []
    for step in range(3):
        x, lab, rand, _ = make_inputs(B, variant, seed=10 + step)
        losses.append(tr.step(x, lab["bbox"], rand_neg_idx=rand).item())
        prev = {k: v.detach().clone() for k, v in P.items()}
        Pr = {k: (v.bfloat16().float() if k.endswith(".weight") else v) for k, v in P.items()}
        Pr = {k: v.detach() + (P[k] - P[k].detach()) for k, v in Pr.items()}  # straight-through: grads reach the fp32 masters
        opt.zero_grad()
        L_ref, _ = O.loss(O.forward(Pr, x, variant), variant, lab["bbox"], rand)
        L_ref.backward()
        opt.step()
        losses_ref.append(L_ref.item())
        tr.store_to_module()
        bad = {}
        for name in [n[:-7] for n in P if n.endswith(".weight") and n != "conv3_3.weight"]:
            w, b = net._wb(name)
            for kind, t in ((".weight", w), (".bias", b)):
                e = rel(t.detach().cpu() - prev[name + kind], P[name + kind].detach() - prev[name + kind])
                if not e <= (5e-2 if name.startswith("conv5") else 0.6):  # backbone: bf16 ReLU-flip noise, see header
                    bad[name + kind] = round(e, 4)
        assert not bad, ("step", step, bad)
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) / abs(b) <= 1e-3, (losses, losses_ref)
    tr.store_to_module()
    for name, mod in (("conv5_2_loc", net.conv5_2_loc), ("conv5_1_det", net.conv5_1_det), ("conv4_4", net.conv4_4_1),
                      ("conv1_1", net.conv1_1_1)):
        for kind, t in ((".weight", mod.weight), (".bias", mod.bias)):
            upd = t.detach().cpu() - w0[name + kind]
            ref = P[name + kind].detach() - w0[name + kind]
            assert rel(upd, ref) <= (5e-2 if name.startswith("conv5") else 0.6), (name + kind, rel(upd, ref))


if __name__ == "__main__":
    for v in ["densebox", "lm", "lmloc"]:
        try:
            test_forward_loss_backward_eval(v)
            print("e2e", v, "OK", flush=True)
        except Exception as e:
            import traceback; traceback.print_exc()
            print("e2e", v, "FAIL", repr(e)[:600], flush=True)
    for fn in (test_train_mode_dropout_injected, test_trainer_matches_torch_sgd):
        try:
            fn()
            print(fn.__name__, "OK", flush=True)
        except Exception as e:
            import traceback; traceback.print_exc()
            print(fn.__name__, "FAIL", repr(e)[:600], flush=True)


def test_inplace_philox_dropout_equals_explicit_mask():
    """dropout_mode 1 (mask drawn inside the conv epilogues) == dropout_mode 2 fed the mask that dbx_dropout_mask
    materialises from the same (seed, offset): identical head outputs and identical gradients, bit for bit."""
    import ctypes
    from densebox_b200 import NetEngine
    from densebox_b200._lib import check, lib, ptr, stream_ptr
    vgg, net = build("densebox")
    net = net.cuda()
    B = 2
    x, lab, rand, _ = make_inputs(B, "densebox")
    res = []
    for mode in (1, 2):
        eng = NetEngine("densebox", B, 240, 240, train=True)
        for name in ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_4", "conv4_1", "conv4_2",
                     "conv4_3", "conv4_4", "conv5_1_det", "conv5_1_loc", "conv5_2_det", "conv5_2_loc"]:
            w, b = net._wb(name)
            eng.set_param(name, w, b)
        eng.refresh_dgrad()
        if mode == 2:
            drop = eng.buffer("drop", torch.bfloat16)
            check(lib().dbx_dropout_mask(ptr(drop), ctypes.c_ulonglong(drop.numel()), ctypes.c_ulonglong(77),
                                         ctypes.c_ulonglong(5), stream_ptr()), "dropout_mask")
            frac = drop.float().mean().item() / 2
            assert 0.49 < frac < 0.51
        eng.forward(x.cuda(), dropout_mode=mode, seed=77, offset=5)
        eng.loss(torch.tensor(lab["bbox"]).cuda(), rand_idx=torch.tensor(rand).cuda())
        eng.zero_grad()
        eng.backward()
        torch.cuda.synchronize()
        res.append((eng.head_out().clone(), eng.flat_grads().clone(), float(eng.loss_value())))
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][2] == res[1][2]
    # wgrad accumulates with fp32 atomics (order varies run to run): compare to accumulation noise
    assert (res[0][1] - res[1][1]).abs().max().item() <= 1e-4 * res[1][1].abs().max().item()


def test_trainer_prefetch_matches_direct_step():
    """trainer.prefetch(batch) + step(batch) (host->device copy on a side stream into staging buffers) must give the
    same losses as step(batch) alone, over several steps with changing batches (staging-buffer reuse, events)."""
    from densebox_b200 import DenseBoxTrainer
    variant, B = "densebox", 2
    batches = []
    for s in range(4):
        x, lab, rand, _ = make_inputs(B, variant, seed=20 + s)
        batches.append((x.pin_memory(), torch.tensor(lab["bbox"]).pin_memory(), torch.tensor(rand).pin_memory()))
    out = []
    for use_prefetch in (False, True):
        _, net = build(variant)
        tr = DenseBoxTrainer(net.cuda(), B, lr=1e-7, dropout=False, use_cuda_graph=True)
        losses = []
        if use_prefetch:
            tr.prefetch(batches[0][0], batches[0][1], rand_neg_idx=batches[0][2])
        for i, (x, bbox, rand) in enumerate(batches):
            loss = tr.step(x, bbox, rand_neg_idx=rand)
            if use_prefetch and i + 1 < len(batches):
                nx = batches[i + 1]
                tr.prefetch(nx[0], nx[1], rand_neg_idx=nx[2])
            losses.append(loss.item())
        out.append(losses)
    assert out[0] == out[1], out


@pytest.mark.parametrize("variant,H,W", [("densebox", 264, 328), ("lmloc", 136, 200)])
def test_inference_forward_other_sizes(variant, H, W):
    """Inference forward (eval, no grad) on non-square inputs whose maps are NOT multiples of the 8 x 16 column-box
    tiles (33 x 41 at conv4 for 264 x 328): ragged tiles, the pairs layout of conv1_1, CTA pairs with an odd tile
    count.  Head maps within 3e-2 of the oracle's largest entry (bf16 storage, same bound as the 240 x 240 test)."""
    _, net = build(variant)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, H, W, generator=g).bfloat16().float()
    with torch.no_grad():
        outs = net(x.cuda())
    P = oracle_params(net, variant)
    with torch.no_grad():
        ref = O.forward(P, x, variant)
    assert len(outs) == len(ref)
    for o, r in zip(outs, ref):
        assert tuple(o.shape) == tuple(r.shape)
        err = (o.float().cpu() - r).abs().max().item()
        assert err <= 3e-2 * r.abs().max().item(), (variant, tuple(o.shape), err, r.abs().max().item())
