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

    n_items = 5

    def evaluate_values(self, users, cutoffs, remove_seen=True):
        self._ev = (len(users), len(cutoffs))

    def evaluate_sums(self, carry_in=None):
        from ganmf_b200._lib import MC_NCOL
        n, nc = self._ev
        base = np.zeros((nc, MC_NCOL)) if carry_in is None else np.array(carry_in)
        # NOT associative on purpose: the chain must be continued in rank order from the carried value
        return base * 2.0 + float(n), np.full((nc, 5), self.rank + 1, dtype=np.int64)


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
        # rank 0: 0*2 + 3 = 3; rank 1 continues from it: 3*2 + 4 = 10; every rank ends up with the last rank's sums
        assert n == 7 and sums == [[10.0] * 13] * 2 and counts == [[3] * 5] * 2


class FakeGanmfEngine(object):
    """Stand-in for a GANMF device engine with the discriminator slab on the host: [enc | dec] parameters and
    gradients.  Each backward phase writes rank-dependent gradients, d_apply_ranges does theta -= lr * g on the
    ranges it is given; every call is logged."""

    class _Cfg(object):
        kind = 0
    cfg = _Cfg()

    def __init__(self, rank, ne, nd):
        self.rank, self.ne, self.nd, self.log = rank, ne, nd, []
        self.params = torch.arange(ne + nd, dtype=torch.float32)
        self.grads = torch.zeros(ne + nd)
        self.b = {"d_grads": self.grads, "g_shared_grad": torch.zeros(3),
                  "step_scalars": torch.zeros(7, dtype=torch.float64),
                  "d_grads_enc": self.grads[:ne], "d_grads_dec": self.grads[ne:],
                  "d_params_enc": self.params[:ne], "d_params_dec": self.params[ne:]}
        self.step = 0

    def _zero_scalars_and_fill_losses(self):
        self.b["step_scalars"][:] = 0.0                 # (the device forward zeroes the step scalars ...)
        self.b["step_scalars"][:2] = float(self.rank + 1)   # (... and accumulates the reconstruction sums)

    def d_forward(self, off, B):
        self.log.append("fwd")
        self._zero_scalars_and_fill_losses()

    def d_forward_phase(self, off, B, phase):
        self.log.append("fwd%d" % phase)
        if phase == 2:
            self.seen_params = self.params.clone()      # what the discriminator forward reads
            self._zero_scalars_and_fill_losses()

    def d_backward_phase(self, B, n_global, m, phase):
        self.log.append("bwd%d" % phase)
        s = self.step + 1
        if phase == 1:
            self.grads[self.ne:] = s * (self.rank + 1) * torch.arange(1, self.nd + 1, dtype=torch.float32)
        if phase == 4:
            self.grads[:self.ne] = s * 10.0 * (self.rank + 1) * torch.arange(1, self.ne + 1, dtype=torch.float32)
            self.step += 1

    def d_apply_ranges(self, lr, reg, offsets, counts, new_step=True):
        self.log.append(("apply", list(offsets), list(counts), new_step))
        for o, c in zip(offsets, counts):
            self.params[o:o + c] -= lr * self.grads[o:o + c]
            self.b["step_scalars"][6] += float(c)       # stands for this rank's share of sum(theta^2)

    def finalize_loss(self, reg, slot):
        self.log.append(("finalize", slot, self.b["step_scalars"][6].item()))

    def g_forward_backward_part(self, off, B, n_global, a, part):
        self.log.append("g_part%d" % part)
        if part == 1:
            self.b["g_shared_grad"][:] = float(self.rank + 1)
            self.b["step_scalars"][:] = float(self.rank + 1)

    def g_apply(self, B, n_global, lr, reg, a, slot):
        self.log.append(("g_apply", self.b["g_shared_grad"].tolist(), self.b["step_scalars"][0].item()))

    def set_gemm_sms(self, n):
        self.log.append(("sms", n))


def _sharded_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GANMF_DP_RESERVE_SMS"] = "32"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ne, nd = 16, 24
    eng = FakeGanmfEngine(rank, ne, nd)
    tr = DataParallelTrainer(eng, world, buffers=eng.b)
    assert tr.overlap and tr.sharded_adam
    lr = 0.5
    tr.d_step(0, 8, lr, 0.0, 1.0, 0)
    tr.d_step(8, 8, lr, 0.0, 1.0, 1)           # starts while the first step's all-gathers are pending
    seen_by_second_forward = eng.seen_params.clone()
    tr.g_step(0, 8, lr, 0.0, 0.1, 2)
    q.put((rank, eng.log, eng.params.tolist(), seen_by_second_forward.tolist()))
    dist.destroy_process_group()


def test_sharded_overlapped_d_step_equals_allreduce_update_gloo():
    """reduce-scatter -> optimiser on this rank's chunk of each half -> all-gather must leave every rank with
    theta - lr * sum_over_ranks(grad), the weights must have landed before the next discriminator forward reads
    them, the loss of step i is settled (sum(theta^2) shares summed) before step i+1 reuses the scalars, and the
    GEMM SM cap brackets exactly the phases that overlap a collective."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ne, nd, lr = 16, 24, 0.5
    theta0 = np.arange(ne + nd, dtype=np.float64)
    g_enc = lambda step: step * 10.0 * 3 * np.arange(1, ne + 1)       # summed over ranks 1 + 2
    g_dec = lambda step: step * 3 * np.arange(1, nd + 1)
    after1 = theta0 - lr * np.concatenate([g_enc(1), g_dec(1)])
    after2 = after1 - lr * np.concatenate([g_enc(2), g_dec(2)])
    for rank, log, params, seen in out:
        np.testing.assert_allclose(params, after2)
        np.testing.assert_allclose(seen, after1)                       # step 2's forward saw step 1's update
        names = [e if isinstance(e, str) else e[0] for e in log]
        assert names[:3] == ["fwd", "bwd1", "sms"]                     # first step: whole forward, then decoder half
        first_apply = log[names.index("apply")]
        assert first_apply[1] == [ne + rank * nd // 2] and first_apply[2] == [nd // 2] and first_apply[3] is True
        # second step: generator part, then the pending work is awaited, the first loss finalised, cap off
        i = names.index("fwd1")
        assert names[i:i + 4] == ["fwd1", "finalize", "sms", "fwd2"]
        assert log[i + 1] == ("finalize", 0, float(ne + nd))           # both halves of both ranks counted once
        assert log[i + 2] == ("sms", 0)
        fin = [e for e in log if not isinstance(e, str) and e[0] == "finalize"]
        assert [f[1] for f in fin] == [0, 1, 2]
        assert ("g_apply", [3.0] * 3, 3.0) in log
        assert names.index("g_part1") < names.index("g_part2") < names.index("g_apply")


def test_shard_rows_partition():
    for n, w in [(10, 3), (138000, 8), (7, 8), (2000000, 4)]:
        spans = [shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
