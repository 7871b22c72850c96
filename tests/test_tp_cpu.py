"""World-size-2 gloo test of ItemShardedTrainer (no GPU): a NumPy stand-in engine performs the per-rank arithmetic
of every phase of ganmf_tp_d_phase / ganmf_tp_g_phase on its item slice (float64), the trainer places the
all-reduces; the gathered result must equal the single-stream oracle stepping on the whole minibatch."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ganmf_b200.parallel import ItemShardedTrainer, item_slices
from oracle import train_oracle as to

WE, BE, WD, BD = to.GANMF_D
P_, V_ = to.GANMF_G


class NumpyTPEngine(object):
    """The arithmetic of one rank of an item-sharded GANMF context (csrc/capi.cu, ganmf_tp_*_phase), float64."""

    def __init__(self, rank, world, urm_slice, width_global, p0_slice, hp, B, lowrank=False):
        self.rank, self.world, self.urm, self.Wg, self.hp = rank, world, urm_slice, width_global, hp
        self.lowrank = lowrank              # ganmf_ctx::lowrank: products over the fake profiles through V^T . We
        self.p = {k: np.array(v, dtype=np.float64) for k, v in p0_slice.items()}
        E, k = self.p[BE].shape[0], self.p[P_].shape[1]
        self.ld_h, self.ld_p = E, k
        self.b = {"tp_h2": torch.zeros(2 * B * E, dtype=torch.float64),
                  "tp_dh2": torch.zeros((2 * B + 1) * E, dtype=torch.float64),
                  "tp_dpb": torch.zeros(B * k, dtype=torch.float64),
                  "tp_m1": torch.zeros(k * E, dtype=torch.float64),
                  "step_scalars": torch.zeros(7, dtype=torch.float64)}
        self.opt_d = to.TFAdam(self.p, to.GANMF_D, hp["d_lr"], np.float64)
        self.opt_g = to.TFAdam(self.p, to.GANMF_G, hp["g_lr"], np.float64)
        self.losses = {}
        self.ids = None

        class _Cfg(object):
            num_factors = k
        self.cfg = _Cfg()

    def step_routes(self):
        return {"sparse_real": False, "bias_grad_from_gemm": False, "lowrank_fake": self.lowrank}

    def device_buffer_ld(self, name):
        return self.ld_p if name == "tp_dpb" else self.ld_h

    def upload_ids(self, ids):
        self.ids = np.asarray(ids)

    def _view(self, name, rows, ld):
        return self.b[name].numpy()[:rows * ld].reshape(rows, ld)

    def _forward_codes(self, off, B, part=0):
        bias = self.p[BE] if self.rank == 0 else 0.0                             # bias joins the sum once
        if part in (0, 6):
            ids = self.ids[off:off + B]
            self.R = np.asarray(self.urm[ids].toarray(), dtype=np.float64)
            self.Pb = self.p[P_][ids]
            self.b["step_scalars"].zero_()
        if part == 6:                      # low-rank route: the real rows' partial codes ...
            self._view("tp_h2", 2 * B, self.ld_h)[:B] = self.R @ self.p[WE] + bias
            return
        self.F = self.Pb @ self.p[V_].T
        self.X2 = np.concatenate([self.R, self.F], 0)
        if part == 7:                      # ... and the [k, E] partial of V^T . We; the fake codes follow in phase 2
            assert self.lowrank
            self._view("tp_m1", self.ld_p, self.ld_h)[:] = self.p[V_].T @ self.p[WE]
            return
        assert not self.lowrank            # (the trainer drives the low-rank route through phases 6 / 7)
        self._view("tp_h2", 2 * B, self.ld_h)[:] = self.X2 @ self.p[WE] + bias

    def _finish_codes(self, B):
        if self.lowrank:                   # Hf = Pb . M1 + be from the summed M1 (the same on every rank)
            M1 = self._view("tp_m1", self.ld_p, self.ld_h)
            self._view("tp_h2", 2 * B, self.ld_h)[B:] = self.Pb @ M1 + self.p[BE]

    def tp_d_phase(self, phase, off, B, lr, reg, m, slot):
        E = self.ld_h
        H2 = self._view("tp_h2", 2 * B, E)
        if phase in (1, 6, 7):
            self._forward_codes(off, B, 0 if phase == 1 else phase)
        elif phase == 2:
            self._finish_codes(B)
            self.Res = H2 @ self.p[WD] + self.p[BD] - self.X2
            sc = self.b["step_scalars"].numpy()
            sc[0], sc[1] = (self.Res[:B] ** 2).sum(), (self.Res[B:] ** 2).sum()
        elif phase == 3:
            sc = self.b["step_scalars"].numpy()
            N = float(B) * self.Wg
            Lr, Lf = sc[0] / N, sc[1] / N
            gate = 1.0 if m * Lr - Lf > 0 else 0.0
            self.loss_main = Lr + max(0.0, m * Lr - Lf)
            self.rs = np.concatenate([np.full(B, (1 + gate * m) * 2 / N), np.full(B, -gate * 2 / N)])[:, None]
            G = self.rs * self.Res
            self.dbd = G.sum(0)
            dh = self._view("tp_dh2", 2 * B + 1, E)
            dh[:2 * B] = G @ self.p[WD].T                       # partial over items
            dh[2 * B] = self.p[WD] @ self.dbd                   # partial dbe
        elif phase == 4:                                        # (on the device: dWd + Adam(Wd), no summed input)
            self.dWd = (self.rs * H2).T @ self.Res
        elif phase == 5:
            dh = self._view("tp_dh2", 2 * B + 1, E)
            l2 = (self.p[WE] ** 2).sum() + (self.p[WD] ** 2).sum() + (self.p[BD] ** 2).sum()
            if self.rank == 0:
                l2 += (self.p[BE] ** 2).sum()
            if self.lowrank:               # dWe = R^T.dH_r + V.(Pb^T.dH_f)  (ganmf_ctx::lowrank_dwe)
                dwe = self.R.T @ dh[:B] + self.p[V_] @ (self.Pb.T @ dh[B:2 * B])
            else:
                dwe = self.X2.T @ dh[:2 * B]
            grads = {WD: self.dWd + reg * self.p[WD], WE: dwe + reg * self.p[WE],
                     BE: dh[2 * B] + reg * self.p[BE], BD: self.dbd + reg * self.p[BD]}
            self.opt_d.apply(self.p, grads)
            self.losses[slot] = (self.loss_main if self.rank == 0 else 0.0) + reg * 0.5 * l2

    def tp_g_phase(self, phase, off, B, lr, reg, a, slot):
        E, k = self.ld_h, self.ld_p
        H2 = self._view("tp_h2", 2 * B, E)
        N, M = float(B) * self.Wg, float(B) * E
        c1, c2 = (1 - a) * 2 / N, a * 2 / M
        if phase in (1, 6, 7):
            self._forward_codes(off, B, 0 if phase == 1 else phase)
        elif phase == 2:
            self._finish_codes(B)
            self.Resf = H2[B:] @ self.p[WD] + self.p[BD] - self.F
            self.sumsq = (self.Resf ** 2).sum()
            self.fm = ((H2[:B] - H2[B:]) ** 2).sum()
            dhf = c1 * (self.Resf @ self.p[WD].T)
            if self.rank == 0:
                dhf = dhf + c2 * (H2[B:] - H2[:B])
            self._view("tp_dh2", 2 * B + 1, E)[B:2 * B] = dhf
        elif phase == 3:
            dhf = self._view("tp_dh2", 2 * B + 1, E)[B:2 * B]
            self.dF = dhf @ self.p[WE].T - c1 * self.Resf
            self._view("tp_dpb", B, k)[:] = self.dF @ self.p[V_]
        elif phase == 4:
            self.dV = self.dF.T @ self.Pb
        # low-rank route (dF is never formed): dPb = [dHf . M1^T on one rank] - c1 * Res_f . V,
        # dV = We . (dHf^T . Pb) - c1 * Res_f^T . Pb; phase 8 runs BEFORE the code gradients are summed
        elif phase == 8:
            self.dV = -c1 * (self.Resf.T @ self.Pb)
            self._view("tp_dpb", B, k)[:] = -c1 * (self.Resf @ self.p[V_])
        elif phase == 9:
            if self.rank == 0:
                dhf = self._view("tp_dh2", 2 * B + 1, E)[B:2 * B]
                self._view("tp_dpb", B, k)[:] += dhf @ self._view("tp_m1", self.ld_p, self.ld_h).T
        elif phase == 10:
            dhf = self._view("tp_dh2", 2 * B + 1, E)[B:2 * B]
            self.dV = self.dV + self.p[WE] @ (dhf.T @ self.Pb)
        elif phase == 5:
            ids = self.ids[off:off + B]
            dP = reg * self.p[P_]
            np.add.at(dP, ids, self._view("tp_dpb", B, k))
            l2v, l2p = (self.p[V_] ** 2).sum(), (self.p[P_] ** 2).sum()
            self.opt_g.apply(self.p, {P_: dP, V_: self.dV + reg * self.p[V_]})
            main = (1 - a) * self.sumsq / N + (a * self.fm / M if self.rank == 0 else 0.0)
            self.losses[slot] = main + reg * 0.5 * (l2v + (l2p if self.rank == 0 else 0.0))

    def read_losses(self, n):
        return np.array([self.losses[i] for i in range(n)], dtype=np.float32)


def _problem():
    rs = np.random.RandomState(0)
    n_rows, width, k, E, B = 96, 131, 6, 10, 32
    urm = sps.random(n_rows, width, 0.1, format="csr", dtype=np.float32, random_state=rs)
    urm.data[:] = 1.0
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=3, dtype=np.float64)
    p0[BE] = rs.standard_normal(E) * 0.01
    p0[BD] = rs.standard_normal(width) * 0.01
    hp = dict(d_lr=1e-3, g_lr=2e-3, d_reg=1e-3, g_reg=1e-3, m=0.3, alpha=0.2)
    return urm, p0, hp, B


def _worker(rank, world, port, q, lowrank=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    urm, p0, hp, B = _problem()
    lo, hi = item_slices(urm.shape[1], world)[rank]
    sl = {WE: p0[WE][lo:hi], BE: p0[BE], WD: p0[WD][:, lo:hi], BD: p0[BD][lo:hi], P_: p0[P_], V_: p0[V_][lo:hi]}
    eng = NumpyTPEngine(rank, world, urm[:, lo:hi].tocsr(), urm.shape[1], sl, hp, B, lowrank)
    tr = ItemShardedTrainer(eng, buffers=eng.b)
    dl, gl = [], []
    for _, batches in to.epoch_index_stream(urm.shape[0], B, 3, seed=5):
        a, b = tr.train_epoch(np.concatenate(batches), B, 1, 1, hp)
        dl += list(a)
        gl += list(b)
    q.put((rank, dl, gl, {n: v.tolist() for n, v in eng.p.items()}))
    dist.destroy_process_group()


@pytest.mark.parametrize("lowrank", [False, True])
def test_item_sharded_trainer_gloo_equals_oracle(lowrank):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, lowrank)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    urm, p0, hp, B = _problem()
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float64)
    odl, ogl = [], []
    for _, batches in to.epoch_index_stream(urm.shape[0], B, 3, seed=5):
        for b in batches:
            odl.append(orc.d_step(b, to.csr_rows_to_dense(urm, b, np.float64), d_reg=hp["d_reg"], m=hp["m"]))
        for b in batches:
            ogl.append(orc.g_step(b, to.csr_rows_to_dense(urm, b, np.float64), g_reg=hp["g_reg"],
                                  recon_coefficient=hp["alpha"]))
    (_, dl0, gl0, p_r0), (_, dl1, gl1, p_r1) = out
    assert dl0 == dl1 and gl0 == gl1                               # every rank reports the summed losses
    np.testing.assert_allclose(dl0, odl, rtol=1e-6)
    np.testing.assert_allclose(gl0, ogl, rtol=1e-6)
    got = {WE: np.concatenate([p_r0[WE], p_r1[WE]], 0), BE: np.array(p_r0[BE]),
           WD: np.concatenate([p_r0[WD], p_r1[WD]], 1), BD: np.concatenate([p_r0[BD], p_r1[BD]]),
           P_: np.array(p_r0[P_]), V_: np.concatenate([p_r0[V_], p_r1[V_]], 0)}
    assert np.array_equal(np.array(p_r0[P_]), np.array(p_r1[P_]))   # replicated tensors stay identical
    assert np.array_equal(np.array(p_r0[BE]), np.array(p_r1[BE]))
    for n in orc.p:
        np.testing.assert_allclose(got[n], orc.p[n], rtol=1e-9, atol=1e-12, err_msg=n)


def test_item_slices_cover_the_columns():
    for n, w in ((200000, 8), (27000, 8), (517, 3), (5, 8)):
        sl = item_slices(n, w)
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= 1
