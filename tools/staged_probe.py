"""One GPU: the single-graph step against the data-parallel step STRUCTURE (four graphs + SGD graph, no communication),
with and without the SM reservation of the overlapped stages — isolates what the data-parallel mode costs by itself."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

c = bench.Ctx()
c.dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
c.rank, c.world, c.dist, c.pk, c.sampler = 0, 1, None, bench.peaks(), None
c.barrier = torch.cuda.synchronize
import densebox_b200
variant = sys.argv[1] if len(sys.argv) > 1 else "densebox"
for name, kw in (("single graph", {}), ("staged, no reserve", dict(staged=True, sm_reserve=0)),
                 ("staged, reserve 4 in stages 1,2", dict(staged=True, sm_reserve=4)),
                 ("staged, reserve 8 in stages 1,2", dict(staged=True, sm_reserve=8)),
                 ("staged, reserve 8 in stage 1", dict(staged=True, sm_reserve=8, reserve_stages=(1,))),
                 ("single graph", {})):
    net = bench.make_net(variant, c.dev)
    tr = densebox_b200.DenseBoxTrainer(net, 32, lr=1e-9, device=c.dev, **kw)
    bs = [{k: v.to(c.dev) for k, v in b.items()} for b in bench.synth(variant, 32, 0, 4)]
    for i in range(6):
        b = bs[i % 4]
        tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
    res = [bench.timed_steps(c, tr, bs, 20, host=False)[0] for _ in range(2)]
    print("%-36s %.3f %.3f ms" % (name, res[0], res[1]), flush=True)
    del tr, net, bs
    torch.cuda.empty_cache()
