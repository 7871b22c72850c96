#!/usr/bin/env python
"""Times the collectives of the data-parallel step on their own (CUDA events, max over ranks):
reduce-scatter / all-gather of one discriminator half (110 MB at cfg4), the item-factor all-reduce (27 MB)
and the 7-scalar all-reduce.  torchrun --nproc-per-node N tools/nccl_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    from ganmf_b200.parallel import init_nccl
    init_nccl(lr)
    n_half = 27000 * 1024 + 1024
    n_half -= n_half % (4 * world)
    half = torch.zeros(n_half, device="cuda")
    dv = torch.zeros(27000 * 256, device="cuda")
    sc = torch.zeros(7, device="cuda", dtype=torch.float64)
    c = n_half // world
    ops = {
        "reduce_scatter %d MB" % (n_half * 4 >> 20): lambda: dist.reduce_scatter_tensor(half[rank * c:(rank + 1) * c], half),
        "all_gather     %d MB" % (n_half * 4 >> 20): lambda: dist.all_gather_into_tensor(half, half[rank * c:(rank + 1) * c]),
        "all_reduce     %d MB" % (n_half * 4 >> 20): lambda: dist.all_reduce(half),
        "all_reduce     %d MB" % (dv.numel() * 4 >> 20): lambda: dist.all_reduce(dv),
        "all_reduce 56 B": lambda: dist.all_reduce(sc),
    }
    for name, fn in ops.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("N=%d ctas=%s  %-24s %.3f ms" % (world, os.environ.get("GANMF_NCCL_MAX_CTAS", "default"), name, t.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
