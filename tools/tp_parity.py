#!/usr/bin/env python
"""Item-sharded parity over NCCL (run under torchrun, one rank per GPU): every rank holds an item slice and
runs ItemShardedTrainer on the GLOBAL minibatch stream; the single-stream oracle replays the same minibatches
and must agree on losses and on the gathered weights (rel <= 1e-3).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/tp_parity.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


def run_case(rank, world, g_reg, m):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    from ganmf_b200.parallel import ItemShardedTrainer, item_slices
    from oracle import train_oracle as to
    WE, BE, WD, BD = to.GANMF_D
    P_, V_ = to.GANMF_G
    n_rows, width, k, E, B = 512, 1301, 24, 48, 64            # B = rows of the whole minibatch
    rs = np.random.RandomState(0)
    urm = sps.random(n_rows, width, 0.05, format="csr", dtype=np.float32, random_state=rs)
    urm.data[:] = 1.0
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=3)
    p0[BE] = (rs.standard_normal(E) * 0.01).astype(np.float32)
    p0[BD] = (rs.standard_normal(width) * 0.01).astype(np.float32)
    hp = dict(d_lr=1e-4, g_lr=2e-4, d_reg=1e-4, g_reg=g_reg, m=m, alpha=0.1)
    lo, hi = item_slices(width, world)[rank]
    eng = Engine(L.KIND_GANMF, n_rows, hi - lo, k, emb_dim=E, max_batch=B, device=torch.cuda.current_device(),
                 global_width=width, item_offset=lo, tp_rank=rank, tp_world=world)
    eng.set_csr(L.CSR_TRAIN, urm[:, lo:hi].tocsr())
    eng.set_params({WE: p0[WE][lo:hi], BE: p0[BE], WD: p0[WD][:, lo:hi], BD: p0[BD][lo:hi], P_: p0[P_], V_: p0[V_][lo:hi]})
    eng.reset_optimizers()
    tr = ItemShardedTrainer(eng)
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    dl, gl, odl, ogl = [], [], [], []
    for _, batches in to.epoch_index_stream(n_rows, B, 4, seed=1337):
        a, b = tr.train_epoch(np.concatenate(batches).astype(np.int32), B, 1, 1, hp)
        dl += list(a)
        gl += list(b)
        if rank == 0:
            for bt in batches:
                odl.append(orc.d_step(bt, to.csr_rows_to_dense(urm, bt), d_reg=hp["d_reg"], m=hp["m"]))
            for bt in batches:
                ogl.append(orc.g_step(bt, to.csr_rows_to_dense(urm, bt), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"]))
    parts = [None] * world
    dist.all_gather_object(parts, eng.get_params())
    ok = True
    if rank == 0:
        got = {WE: np.concatenate([q[WE] for q in parts], 0), BE: parts[0][BE],
               WD: np.concatenate([q[WD] for q in parts], 1), BD: np.concatenate([q[BD] for q in parts]),
               P_: parts[0][P_], V_: np.concatenate([q[V_] for q in parts], 0)}
        e_d = float(np.max(np.abs(np.array(dl) - np.array(odl)) / np.abs(odl)))
        e_g = float(np.max(np.abs(np.array(gl) - np.array(ogl)) / np.abs(ogl)))
        errs = {n: rel_err(got[n], orc.p[n]) for n in orc.p}
        same = all(np.array_equal(q[P_], parts[0][P_]) and np.array_equal(q[BE], parts[0][BE]) for q in parts[1:])
        ok = e_d <= 1e-3 and e_g <= 1e-3 and max(errs.values()) <= 1e-3 and same
        print("N=%d g_reg=%g m=%g: %d D + %d G steps, max rel loss err D %.2e G %.2e, worst tensor err %.2e, "
              "replicated tensors bit-identical: %s -> %s" % (world, g_reg, m, len(dl), len(gl), e_d, e_g,
                                                               max(errs.values()), same, "ok" if ok else "FAIL"))
    eng.close()
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    from ganmf_b200.parallel import init_nccl
    init_nccl(int(os.environ["LOCAL_RANK"]))
    ok = True
    for g_reg, m in ((0.0, 10.0), (1e-3, 0.05)):
        ok = run_case(rank, world, g_reg, m) and ok
    if rank == 0:
        print("TP PARITY", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
