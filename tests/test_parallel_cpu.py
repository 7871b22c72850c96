"""World-size-2 gloo test of the data-parallel host logic (no GPU): the collectives sit at the right
points of the step, sums are global, per-rank shards are disjoint, sharded evaluation reduces only sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ganmf_b200.parallel import DataParallelTrainer, shard_rows, sharded_eval_sums


class FakeEngine(object):
    """Stand-in for the device engine: each phase writes rank-dependent values into the shared buffers
    and records the call order."""

    def __init__(self, rank, bufs):
        self.rank, self.b, self.log = rank, bufs, []

    class _Cfg(object):
        kind = 1
    cfg = _Cfg()

    def d_forward(self, off, B):
        self.log.append("d_forward")
        self.b["step_scalars"][:] = torch.tensor([1.0 + self.rank, 10.0 * (1 + self.rank), 0, 0, 0, 0, 0],
                                                 dtype=torch.float64)

    def d_backward(self, B, n_global, m):
        self.log.append(("d_backward", n_global, self.b["step_scalars"][:2].tolist()))
        self.b["d_grads"][:] = float(self.rank + 1)

    def d_apply(self, lr, reg, slot):
        self.log.append(("d_apply", self.b["d_grads"].tolist()))

    def g_forward_backward(self, off, B, n_global, a):
        self.log.append(("g_fb", n_global))
        self.b["g_shared_grad"][:] = float(10 * (self.rank + 1))
        self.b["step_scalars"][:] = float(self.rank + 1)

    def g_apply(self, B, n_global, lr, reg, a, slot):
        self.log.append(("g_apply", self.b["g_shared_grad"].tolist(), self.b["step_scalars"][0].item()))

    def finalize_loss(self, reg, slot):
        self.log.append("finalize")

    def evaluate(self, users, cutoffs, remove_seen=True):
        return np.full((len(cutoffs), 3), float(len(users))), np.full((len(cutoffs), 5), self.rank + 1, dtype=np.int64)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bufs = {"d_grads": torch.zeros(4), "g_shared_grad": torch.zeros(3), "step_scalars": torch.zeros(7, dtype=torch.float64)}
    eng = FakeEngine(rank, bufs)
    tr = DataParallelTrainer(eng, world, buffers=bufs)
    tr.d_step(0, 8, 1e-3, 0.0, 1.0, 0)
    tr.g_step(0, 8, 1e-3, 0.0, 0.1, 1)
    sums, counts, n = sharded_eval_sums(eng, np.arange(3 + rank), [5, 10], True, dist=dist)
    q.put((rank, eng.log, sums.tolist(), counts.tolist(), n))
    dist.destroy_process_group()


def test_dp_step_collectives_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, log, sums, counts, n in out:
        assert log[0] == "d_forward"
        assert log[1] == ("d_backward", 16, [3.0, 30.0])          # scalars summed BEFORE the backward
        assert log[2] == ("d_apply", [3.0] * 4)                    # gradients summed before Adam
        assert log[3] == ("g_fb", 16)
        assert log[4] == ("g_apply", [30.0] * 3, 3.0) and log[5] == "finalize"
        assert n == 7 and sums == [[7.0] * 3] * 2 and counts == [[3] * 5] * 2


def test_shard_rows_partition():
    for n, w in [(10, 3), (138000, 8), (7, 8), (2000000, 4)]:
        spans = [shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
