#!/usr/bin/env python
"""Data-parallel parity (run under torchrun, one rank per GPU):
every rank trains its own user shard with DataParallelTrainer; the single-stream oracle replays the
GLOBAL minibatches (concatenation of the ranks' local batches) and must agree on losses and weights.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    from ganmf_b200.parallel import init_nccl
    init_nccl(int(os.environ["LOCAL_RANK"]))
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    from ganmf_b200.parallel import DataParallelTrainer, shard_rows
    from oracle import train_oracle as to

    ok = True
    for g_reg in (1e-3, 0.0):        # 0.0: the lazy user-factor optimiser (kernels.cuh K6b) under row sharding
        ok = run_case(rank, world, L, Engine, DataParallelTrainer, shard_rows, to, g_reg) and ok
    ok = run_case(rank, world, L, Engine, DataParallelTrainer, shard_rows, to, 1e-4, dis=True) and ok
    if rank == 0:
        print("DP PARITY", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def run_case(rank, world, L, Engine, DataParallelTrainer, shard_rows, to, g_reg, dis=False):
    """dis: DisGANMF (2 tanh layers of 48 units; the GLOBAL row id is a discriminator input feature, so every
    rank's engine carries row_id_offset = first row of its shard)."""
    n_rows, width, k, E, B = 512, 700, 24, 48, 32            # B per rank
    rs = np.random.RandomState(0)
    urm = sps.random(n_rows, width, 0.05, format="csr", dtype=np.float32, random_state=rs)
    urm.data[:] = 1.0
    p0 = to.init_disganmf_params(n_rows, width, k, 2, E, seed=3) if dis else to.init_ganmf_params(n_rows, width, k, E, seed=3)
    hp = dict(d_lr=1e-4, g_lr=2e-4, d_reg=1e-4, g_reg=g_reg, m=10.0, alpha=0.1)
    lo, hi = shard_rows(n_rows, world, rank)
    if dis:
        eng = Engine(L.KIND_DISGANMF, hi - lo, width, k, d_layers=2, d_nodes=E, d_act="tanh", max_batch=B,
                     row_id_offset=lo, device=torch.cuda.current_device())
    else:
        eng = Engine(L.KIND_GANMF, hi - lo, width, k, emb_dim=E, max_batch=B, device=torch.cuda.current_device())
    eng.set_csr(L.CSR_TRAIN, urm[lo:hi])
    local = dict(p0)
    local["generator/user_embeddings"] = p0["generator/user_embeddings"][lo:hi]
    eng.set_params(local)
    eng.reset_optimizers()
    tr = DataParallelTrainer(eng, world)
    epochs = 3
    prs = np.random.RandomState(100 + rank)
    n_local = ((hi - lo) // B) * B                            # equal step counts on every rank
    perms = [prs.permutation(hi - lo)[:n_local].astype(np.int32) for _ in range(epochs)]
    dl, gl = [], []
    for p in perms:
        a, b = tr.train_epoch(p, B, 1, 1, hp)
        dl += list(a)
        gl += list(b)
    # gather every rank's local permutations (as global ids) and user-factor shards on rank 0
    all_perms = [None] * world
    dist.all_gather_object(all_perms, [(p + lo).tolist() for p in perms])
    shards = [None] * world
    dist.all_gather_object(shards, eng.get_param("generator/user_embeddings"))
    got = eng.get_params()
    ok = True
    if rank == 0:
        orc = (to.DisGanmfOracle(p0, 2, "tanh", hp["d_lr"], hp["g_lr"], dtype=np.float32) if dis else
               to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32))
        odl, ogl = [], []
        nb = n_local // B
        for e in range(epochs):
            batches = [np.concatenate([np.array(all_perms[r][e][b * B:(b + 1) * B]) for r in range(world)])
                       for b in range(nb)]
            for ids in batches:
                odl.append(orc.d_step(ids, to.csr_rows_to_dense(urm, ids), d_reg=hp["d_reg"]) if dis else
                           orc.d_step(ids, to.csr_rows_to_dense(urm, ids), d_reg=hp["d_reg"], m=hp["m"]))
            for ids in batches:
                ogl.append(orc.g_step(ids, to.csr_rows_to_dense(urm, ids), g_reg=hp["g_reg"],
                                      recon_coefficient=hp["alpha"]))
        got["generator/user_embeddings"] = np.concatenate(shards, axis=0)
        dmax = float(np.max(np.abs(np.array(dl) / np.array(odl) - 1)))
        gmax = float(np.max(np.abs(np.array(gl) / np.array(ogl) - 1)))
        print("world=%d %s g_reg=%g steps=%d  max rel loss diff: D %.2e  G %.2e" %
              (world, "DisGANMF" if dis else "GANMF", g_reg, len(dl), dmax, gmax))
        ok = dmax < 1e-3 and gmax < 1e-3
        for n in orc.p:
            err = np.linalg.norm(got[n] - orc.p[n]) / np.linalg.norm(orc.p[n])
            print("  %-34s rel err %.2e" % (n, err))
            ok = ok and err < 1e-3
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(okt, 0)
    eng.close()
    return bool(okt.item())


if __name__ == "__main__":
    sys.exit(main())
