"""Data-parallel step time with / without the gradient all-reduces, for both variants and for the two overlap modes
(dedicated low-CTA NCCL communicator + SM reservation, or plain overlap on the default group).  Run under torchrun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    c = bench.Ctx()
    c.dev = torch.device("cuda", lr)
    c.rank, c.world, c.dist, c.pk, c.sampler = rank, world, dist, bench.peaks(), None
    dist.init_process_group("nccl", device_id=c.dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    c.barrier = barrier
    import densebox_b200
    order = sys.argv[1:] or ["lm", "densebox", "lm"]
    for variant in order:
        for ctas, stages in ((8, (1, 2)),):
            net = bench.make_net(variant, c.dev)
            tr = densebox_b200.DenseBoxTrainer(net, 32, lr=1e-9, process_group=dist.group.WORLD, device=c.dev,
                                               nccl_max_ctas=ctas, reserve_stages=stages)
            bs = [{k: v.to(c.dev) for k, v in b.items()} for b in bench.synth(variant, 32, rank, 4)]
            for i in range(6):
                b = bs[i % 4]
                tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
            res = []
            for mask in (0, 15, 0):
                tr._skip_mask = mask
                for i in range(2):
                    b = bs[i % 4]
                    tr.step(b["x"], b["bbox"], vertices=b.get("vertices"), rand_neg_idx=b["rand"], lm_rand_neg_idx=b.get("lm_rand"))
                ms, _, _ = bench.timed_steps(c, tr, bs, 20, host=False)
                res.append("skip=%2d: %.3f" % (mask, bench.max_over_ranks(c, [ms])[0]))
            if rank == 0:
                print("%-9s world %d nccl_max_ctas %2d (reserve %d SMs in stages %s, peer count exchange %s): %s" % (variant, world, ctas, tr.sm_reserve, stages, tr._slots is not None, " | ".join(res)),
                      flush=True)
            del tr, net, bs
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
