"""Data parallelism over users (one process per GPU, torch.distributed / NCCL for the plumbing).

Partitioning: rank g owns a contiguous shard of the training rows -- its CSR
rows, its rows of the user-factor matrix P and their Adam moments.  The item factors V and the
discriminator are replicated.  Per step every rank takes its own minibatch from its own shard; the
exchanged quantities are sums:
  D step: [sum (Dr-R)^2, sum (Df-F)^2] (hinge gate needs the GLOBAL batch means), then all
          discriminator gradients (one contiguous buffer);
  G step: the item-factor gradient and the loss scalars.  dP rows never leave their owner.
Gradients are already normalised by the global element count, so the collective is a plain SUM."""
import os
import time

import numpy as np


# SMs set aside for NCCL while a collective overlaps the GEMMs, and the matching cap on NCCL's CTAs.  Measured
# at N=2 on cfg4 (profiles/r01_dp_overlap_ab_n2.txt): no reservation 2.73 ms/step, 16: 2.99, 24: 2.79,
# 32: 2.59, 40: 2.57 -- below 32 CTAs NCCL's own bandwidth drops (reduce-scatter of 105 MB: 0.155 -> 0.206 ms
# at 16), without a reservation its kernels wait for the persistent GEMM to drain.  At N=8 the collectives run
# through the switch (NVLS) and do not compete for SMs: 2.722 ms with the reservation, 2.701 without, so it is
# applied to two-rank groups only (GANMF_DP_RESERVE_SMS overrides).
NCCL_CTAS = 32


def init_nccl(local_rank):
    """One process per GPU over NCCL/NVLink.  GANMF_NCCL_MAX_CTAS caps the SMs NCCL may occupy: its kernels run
    concurrently with the tcgen05 GEMMs (whose CTAs need a whole SM each), and over NVSwitch a collective does
    not need many CTAs to saturate the links."""
    import torch
    import torch.distributed as dist
    kw = {}
    cap = int(os.environ.get("GANMF_NCCL_MAX_CTAS", str(NCCL_CTAS)))
    if cap > 0:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = cap
        opts.config.min_ctas = min(cap, max(1, int(os.environ.get("GANMF_NCCL_MIN_CTAS", "1"))))
        kw["pg_options"] = opts
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), **kw)


class DataParallelTrainer(object):
    def __init__(self, engine, world_size, group=None, buffers=None):
        """buffers: optional {"d_grads", "g_shared_grad", "step_scalars"} tensors (the CPU/gloo tests pass
        host tensors of a stand-in engine); by default the engine's device buffers are wrapped zero-copy."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.eng, self.world, self.group = engine, world_size, group
        self.overlap = False
        self.sharded_adam = False
        self._pending = []
        self._pending_loss = None
        if buffers is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            names = ["d_grads", "g_shared_grad", "step_scalars"]
            if engine.cfg.kind == 0:                 # GANMF: decoder / encoder halves for the overlapped sum
                names += ["d_grads_dec", "d_grads_enc", "d_params_dec", "d_params_enc"]
                self.overlap = True
            buffers = {n: torch.as_tensor(engine.device_buffer(n), device=dev) for n in names}
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        self.d_grads, self.g_shared, self.scalars = (buffers["d_grads"], buffers["g_shared_grad"],
                                                     buffers["step_scalars"])
        self.d_dec, self.d_enc = buffers.get("d_grads_dec"), buffers.get("d_grads_enc")
        self.p_dec, self.p_enc = buffers.get("d_params_dec"), buffers.get("d_params_enc")
        if engine.cfg.kind == 0 and None not in (self.d_dec, self.d_enc, self.p_dec, self.p_enc):
            self.overlap = True                      # (also for a stand-in engine that hands in all four halves)
        if self.overlap and self.p_dec is not None:
            # reduce-scatter -> Adam on this rank's 1/N of every half -> all-gather of the parameters:
            # same bytes on the wire as an all-reduce, but the (HBM-bound) optimiser work is divided by N
            ne, nd = self.d_enc.numel(), self.d_dec.numel()
            self.sharded_adam = world_size > 1 and ne % (4 * world_size) == 0 and nd % (4 * world_size) == 0
            if self.sharded_adam:
                r = dist.get_rank(group)
                ce, cd = ne // world_size, nd // world_size
                self._enc_chunk, self._dec_chunk = slice(r * ce, (r + 1) * ce), slice(r * cd, (r + 1) * cd)
                # slab layout: [We | be | Wd | bd]  (enc half first)
                self._ranges = ([r * ce, ne + r * cd], [ce, cd])
                self._rank = r
                # SMs the GEMMs leave to NCCL while a collective is in flight (0 = no reservation)
                self._reserve = int(os.environ.get("GANMF_DP_RESERVE_SMS", str(NCCL_CTAS if world_size == 2 else 0)))

    def _cap(self, on):
        if getattr(self, "_reserve", 0) > 0:
            self.eng.set_gemm_sms(148 - self._reserve if on else 0)

    def _sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def _wait_pending(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self._pending_loss is not None:
            self.eng.finalize_loss(*self._pending_loss)
            self._pending_loss = None

    def d_step(self, ids_offset, B, lr, reg, m_hinge, loss_slot):
        n_global = B * self.world
        if self._pending:
            self.eng.d_forward_phase(ids_offset, B, 1)    # profiles + generator overlap the weight all-gather
            self._wait_pending()
            self._cap(False)
            self.eng.d_forward_phase(ids_offset, B, 2)
        else:
            self.eng.d_forward(ids_offset, B)
        self._sum(self.scalars)
        if self.sharded_adam:
            # Schedule (NCCL ops run on their own stream; -> = stream order, || = concurrent):
            #   dWd,dbd -> RS(dec) || dH                      (RS of the decoder half hides behind dH)
            #   Adam(dec chunk) -> AG(dec) || dWe             (decoder weights travel while dWe is computed)
            #   RS(enc) -> Adam(enc chunk) -> AG(enc) || next step's profiles + generator
            d = self.dist
            self.eng.d_backward_phase(B, n_global, m_hinge, 1)
            w = d.reduce_scatter_tensor(self.d_dec[self._dec_chunk], self.d_dec, op=d.ReduceOp.SUM, group=self.group,
                                        async_op=True)
            self._cap(True)                                               # until the last all-gather has landed
            self.eng.d_backward_phase(B, n_global, m_hinge, 3)            # dH: last reader of the old Wd
            w.wait()
            self.eng.d_apply_ranges(lr, reg, [self._ranges[0][1]], [self._ranges[1][1]], new_step=True)
            ag_dec = d.all_gather_into_tensor(self.p_dec, self.p_dec[self._dec_chunk], group=self.group,
                                              async_op=True)
            self.eng.d_backward_phase(B, n_global, m_hinge, 4)            # dWe
            d.reduce_scatter_tensor(self.d_enc[self._enc_chunk], self.d_enc, op=d.ReduceOp.SUM, group=self.group)
            self.eng.d_apply_ranges(lr, reg, [self._ranges[0][0]], [self._ranges[1][0]], new_step=False)
            ag_enc = d.all_gather_into_tensor(self.p_enc, self.p_enc[self._enc_chunk], group=self.group,
                                              async_op=True)
            # sum(theta^2) lives in rank shards; its sum and the loss bookkeeping trail the all-gather (NCCL ops
            # are serial) and are settled when the weights are awaited, before the scalars are reused
            w_l2 = d.all_reduce(self.scalars[6:7], op=d.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending = [ag_dec, ag_enc, w_l2]                       # awaited before the weights are read
            self._pending_loss = (reg, loss_slot)
            return
        if self.overlap:
            # decoder gradients are summed on NCCL's stream while the encoder half is still computed
            self.eng.d_backward_phase(B, n_global, m_hinge, 1)
            w = self.dist.all_reduce(self.d_dec, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.eng.d_backward_phase(B, n_global, m_hinge, 2)
            self._sum(self.d_enc)
            w.wait()
        else:
            self.eng.d_backward(B, n_global, m_hinge)
            self._sum(self.d_grads)
            if self.eng.cfg.kind != 0:
                self._sum(self.scalars)       # DisGANMF forms its loss sums in the backward (no gradient needs them)
        self.eng.d_apply(lr, reg, loss_slot)

    def g_step(self, ids_offset, B, lr, reg, recon_coefficient, loss_slot):
        n_global = B * self.world
        self._wait_pending()
        self._cap(False)
        if self.overlap:
            # dV is complete after part 1: its sum over ranks (and the loss scalars) travel while dPb is computed
            d = self.dist
            self.eng.g_forward_backward_part(ids_offset, B, n_global, recon_coefficient, 1)
            w1 = d.all_reduce(self.g_shared, op=d.ReduceOp.SUM, group=self.group, async_op=True)
            w2 = d.all_reduce(self.scalars, op=d.ReduceOp.SUM, group=self.group, async_op=True)
            self._cap(True)
            self.eng.g_forward_backward_part(ids_offset, B, n_global, recon_coefficient, 2)
            self._cap(False)
            w1.wait()
            w2.wait()
        else:
            self.eng.g_forward_backward(ids_offset, B, n_global, recon_coefficient)
            self._sum(self.g_shared)
            self._sum(self.scalars)
        self.eng.g_apply(B, n_global, lr, reg, recon_coefficient, loss_slot)
        if reg != 0.0:
            self._sum(self.scalars[6:7])          # ||P||^2 lives in row shards
        self.eng.finalize_loss(reg, loss_slot)

    def train_epoch(self, perm_local, batch_size, d_steps, g_steps, hp):
        """Reference schedule (GANMF.py:172-203) on this rank's shard; every rank must pass the same
        number of ids.  Returns (d_losses, g_losses): the GLOBAL per-step losses."""
        n = len(perm_local)
        nb = (n + batch_size - 1) // batch_size
        self.eng.upload_ids(perm_local)
        slot = 0
        for _ in range(d_steps):
            for b in range(nb):
                off = b * batch_size
                self.d_step(off, min(batch_size, n - off), hp["d_lr"], hp["d_reg"], hp["m"], slot)
                slot += 1
        nd = slot
        for _ in range(g_steps):
            for b in range(nb):
                off = b * batch_size
                self.g_step(off, min(batch_size, n - off), hp["g_lr"], hp["g_reg"], hp["alpha"], slot)
                slot += 1
        self._wait_pending()
        self._cap(False)
        losses = self.eng.read_losses(slot)
        return losses[:nd], losses[nd:]

    def e2e_epoch(self, rs, n_rows, K, B, hp):
        """bench.py: host ids in, losses out, wall clock around the whole call (ms)."""
        perm = rs.permutation(n_rows)[:K * B].astype(np.int32)
        self.dist.barrier()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.train_epoch(perm, B, 1, 1, hp)
        self.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        return {"ms": ms, "unit": "rows/s", "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 8,
                "api": "DataParallelTrainer.train_epoch: host row ids in, per-step losses out"}


class ItemShardedTrainer(object):
    """GANMF training partitioned over ITEMS: rank r holds columns [lo_r, hi_r) of the
    training matrix and the matching slices of We (rows), Wd / bd (columns) and V (rows); the user factors and the
    encoder bias are replicated.  Every rank runs the WHOLE minibatch on its slice, so

      * everything indexed by items -- residuals, dWe, dWd, dbd, dV and their Adam updates (fused into the
        weight-gradient GEMM epilogues, as on one GPU) -- is local: no weight or weight gradient crosses NVLink;
      * the contractions over items are partial sums, all-reduced (SUM) per step:
          D: codes [2B, E] -> 2 energy sums (the hinge gate is global) -> code gradients [2B, E] + dbe [E]
          G: codes [2B, E] -> fake code gradients [B, E] -> user-factor gradients [B, k]

    At cfg5 (I = 200 000, E = 1024, B = 8 x 1024) that is 0.23 GB of activations per D+G step pair instead of the
    1.64 GB reduce-scatter + 1.64 GB all-gather of the data-parallel weight exchange (and the optimiser traffic of
    the discriminator is divided by N).  Same arithmetic as one GPU stepping on the whole minibatch, up to the
    summation order over items.  The loss log holds per-rank partial losses; train_epoch sums them once.

    `engines`: the engine of this process, or (tests, one process) a list of the engines of ALL ranks living on
    one device -- the sums are then formed in place with torch instead of NCCL.  `buffers`: per-engine dicts of
    host tensors for a stand-in engine (CPU/gloo tests)."""

    NAMES = ("tp_h2", "tp_dh2", "tp_dpb", "tp_m1", "step_scalars")

    def __init__(self, engines, group=None, buffers=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.engines = list(engines) if isinstance(engines, (list, tuple)) else [engines]
        self.local = len(self.engines) > 1
        if buffers is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            buffers = []
            for e in self.engines:
                e.set_stream(torch.cuda.current_stream().cuda_stream)
                buffers.append({n: torch.as_tensor(e.device_buffer(n), device=dev) for n in self.NAMES})
        elif isinstance(buffers, dict):
            buffers = [buffers]
        self.buf = buffers
        e0 = self.engines[0]
        self.ld_h, self.ld_p = e0.device_buffer_ld("tp_h2"), e0.device_buffer_ld("tp_dpb")
        self.world = len(self.engines) if self.local else dist.get_world_size(group)
        # low-rank generator route (ganmf_step_routes): the fake rows' codes are Pb . (V^T . We), so after phase 1 only
        # the real rows' partial codes and the [k, E] partial of V^T . We are summed (B + k rows instead of 2B)
        self.lowrank = bool(e0.step_routes()["lowrank_fake"])
        self.k = e0.cfg.num_factors

    def _sum(self, name, lo, hi, async_op=False):
        if self.local:
            views = [b[name][lo:hi] for b in self.buf]
            tot = views[0].clone()
            for v in views[1:]:
                tot += v
            for v in views:
                v.copy_(tot)
            return None
        return self.dist.all_reduce(self.buf[0][name][lo:hi], op=self.dist.ReduceOp.SUM, group=self.group,
                                    async_op=async_op)

    def _forward_codes(self, fn, a, B):
        """Phase 1 and the sums behind it.  Low-rank route: the real rows' codes travel while the generator GEMM and
        V^T.We are computed (phases 6 / 7), then the [k, E] partial of V^T.We (1 MB) is summed."""
        if not self.lowrank:
            self._phase(fn, 1, *a)
            self._sum("tp_h2", 0, 2 * B * self.ld_h)
            return
        self._phase(fn, 6, *a)
        w = self._sum("tp_h2", 0, B * self.ld_h, async_op=True)
        self._phase(fn, 7, *a)
        self._sum("tp_m1", 0, self.k * self.ld_h)
        if w is not None:
            w.wait()

    def _phase(self, fn, *a):
        for e in self.engines:
            getattr(e, fn)(*a)

    def d_step(self, ids_offset, B, lr, reg, m_hinge, loss_slot):
        a = (ids_offset, B, lr, reg, m_hinge, loss_slot)
        self._forward_codes("tp_d_phase", a, B)
        self._phase("tp_d_phase", 2, *a)
        self._sum("step_scalars", 0, 2)
        self._phase("tp_d_phase", 3, *a)
        w = self._sum("tp_dh2", 0, (2 * B + 1) * self.ld_h, async_op=True)    # travels while dWd (+ Adam) is computed
        self._phase("tp_d_phase", 4, *a)
        if w is not None:
            w.wait()
        self._phase("tp_d_phase", 5, *a)

    def g_step(self, ids_offset, B, lr, reg, recon_coefficient, loss_slot):
        a = (ids_offset, B, lr, reg, recon_coefficient, loss_slot)
        self._forward_codes("tp_g_phase", a, B)
        self._phase("tp_g_phase", 2, *a)
        if self.lowrank:
            # the fake code gradients travel while the halves of dV / dPb that do not need them are computed
            w = self._sum("tp_dh2", B * self.ld_h, 2 * B * self.ld_h, async_op=True)
            self._phase("tp_g_phase", 8, *a)
            if w is not None:
                w.wait()
            self._phase("tp_g_phase", 9, *a)
            w = self._sum("tp_dpb", 0, B * self.ld_p, async_op=True)  # travels while dV is completed
            self._phase("tp_g_phase", 10, *a)
        else:
            self._sum("tp_dh2", B * self.ld_h, 2 * B * self.ld_h)
            self._phase("tp_g_phase", 3, *a)
            w = self._sum("tp_dpb", 0, B * self.ld_p, async_op=True)  # travels while the item-factor gradient is computed
            self._phase("tp_g_phase", 4, *a)
        if w is not None:
            w.wait()
        self._phase("tp_g_phase", 5, *a)

    def read_losses(self, n):
        """Per-step losses = sum over ranks of the per-rank partial losses."""
        if self.local:
            return np.sum([e.read_losses(n) for e in self.engines], axis=0, dtype=np.float32)
        part = self.engines[0].read_losses(n)
        if self.world > 1:
            dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
            t = self.torch.from_numpy(part).to(dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            part = t.cpu().numpy()
        return part

    def train_epoch(self, perm, batch_size, d_steps, g_steps, hp):
        """Reference schedule (GANMF.py:172-203) on the GLOBAL id stream `perm` (the same on every rank;
        batch_size = rows of the whole minibatch).  Returns (d_losses, g_losses)."""
        n = len(perm)
        nb = (n + batch_size - 1) // batch_size
        for e in self.engines:
            e.upload_ids(perm)
        slot = 0
        for _ in range(d_steps):
            for b in range(nb):
                off = b * batch_size
                self.d_step(off, min(batch_size, n - off), hp["d_lr"], hp["d_reg"], hp["m"], slot)
                slot += 1
        nd = slot
        for _ in range(g_steps):
            for b in range(nb):
                off = b * batch_size
                self.g_step(off, min(batch_size, n - off), hp["g_lr"], hp["g_reg"], hp["alpha"], slot)
                slot += 1
        losses = self.read_losses(slot)
        return losses[:nd], losses[nd:]

    def encode(self, rows):
        """autoencoder_codes (GANMF.py:304-307) of an item-sharded model: R[rows] . We + be, the partial products over
        the item slices summed like the codes of a training step (phase 1 of the D step)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        e0 = self.engines[0]
        B, E = e0.cfg.max_batch, e0.cfg.emb_dim
        out = np.empty((rows.size, E), dtype=np.float32)
        for s in range(0, rows.size, B):
            ids = rows[s:s + B]
            for e in self.engines:
                e.upload_ids(ids)
            self._phase("tp_d_phase", 6 if self.lowrank else 1, 0, ids.size, 0.0, 0.0, 1.0, 0)
            self._sum("tp_h2", 0, ids.size * self.ld_h)
            h = self.buf[0]["tp_h2"][:ids.size * self.ld_h].reshape(ids.size, self.ld_h)[:, :E]
            out[s:s + ids.size] = h.cpu().numpy()
        return out


def item_slices(n_items, world_size):
    """[lo, hi) column ranges of the item-sharded layout (contiguous, sizes differ by at most one)."""
    return [shard_rows(n_items, world_size, r) for r in range(world_size)]


def shard_rows(n_rows, world_size, rank):
    """Contiguous row range [lo, hi) owned by `rank`."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_eval_sums(engine, users_local, cutoffs, remove_seen, dist=None, group=None):
    """Evaluation sharded by user rows: only the metric sums (and the per-item histograms) cross GPUs.

    `users_local` must be this rank's CONTIGUOUS range of the ascending global user list (rank order = user order).
    Every rank computes the per-user metric values of its users in parallel; the running sums are then continued
    rank after rank (rank r starts from the sums of ranks < r), so the result equals the single-GPU running sums over
    all users bit for bit (the reference keeps one running float per metric, Evaluator.py:305-335).  The serial part
    is the ordered accumulation alone (~12 ns per user)."""
    import torch
    engine.evaluate_values(users_local, cutoffs, remove_seen=remove_seen)
    n = np.array([len(users_local)], dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        sums, counts = engine.evaluate_sums(None)
        return sums, counts, int(n[0])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    ranks = [dist.get_global_rank(group, r) for r in range(world)] if group is not None else list(range(world))
    carry = None
    from . import _lib as L
    shape = (len(cutoffs), L.MC_NCOL)
    if rank > 0:
        t = torch.zeros(shape, dtype=torch.float64, device=dev)
        dist.recv(t, src=ranks[rank - 1], group=group)
        carry = t.cpu().numpy()
    sums, counts = engine.evaluate_sums(carry)
    t = torch.from_numpy(sums).to(dev)
    if rank < world - 1:
        dist.send(t, dst=ranks[rank + 1], group=group)
    dist.broadcast(t, src=ranks[world - 1], group=group)          # the last rank holds the sums over all users
    sums = t.cpu().numpy()
    for a in (counts, n):
        t = torch.from_numpy(a).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        a[...] = t.cpu().numpy()
    return sums, counts, int(n[0])
