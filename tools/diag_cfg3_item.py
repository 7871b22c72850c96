#!/usr/bin/env python
"""Per-tensor diagnosis of one D update of DisGANMF at the cfg3-item shape (teacher-forced, device vs fp32/fp64 oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import train_oracle as to  # noqa: E402
from tests.helpers import load_quality_targets, load_split  # noqa: E402
from tests.test_gpu_baseline_shapes import pick_batches  # noqa: E402


def main():
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    bp = load_quality_targets()["DisGANMF_item_hetrec2011"]["best_params"]
    urm = load_split("Movielenshetrec2011")["train"].T.tocsr()
    n_rows, width = urm.shape
    k, B, layers, nodes = int(bp["num_factors"]), int(bp["batch_size"]), int(bp["d_layers"]), int(bp["d_nodes"])
    act = bp["d_hidden_act"]
    path = getattr(L, sys.argv[1] if len(sys.argv) > 1 else "GEMM_SIMT")
    p = to.init_disganmf_params(n_rows, width, k, layers, nodes, seed=3)
    eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=layers, d_nodes=nodes, d_act=act, max_batch=B, item_mode=True,
                 gemm_path=path)
    eng.set_csr(L.CSR_TRAIN, urm)
    b = pick_batches(n_rows, B, 1)[0]
    R = to.csr_rows_to_dense(urm, b)
    eng.set_params(p)
    eng.reset_optimizers()
    o32 = to.DisGanmfOracle(p, layers, act, bp["d_lr"], bp["g_lr"], dtype=np.float32)
    o64 = to.DisGanmfOracle(p, layers, act, bp["d_lr"], bp["g_lr"], dtype=np.float64)
    eng.upload_ids(b.astype(np.int32))
    eng.d_step(0, len(b), bp["d_lr"], bp["d_reg"], 1.0, loss_slot=0)
    l32, l64 = o32.d_step(b, R, d_reg=bp["d_reg"]), o64.d_step(b, R, d_reg=bp["d_reg"])
    print("loss dev %.6f o32 %.6f o64 %.6f" % (eng.read_losses(1)[0], l32, l64))
    got = eng.get_params()
    for n in to.disganmf_d_names(layers):
        du = got[n].astype(np.float64) - p[n]
        d32 = o32.p[n].astype(np.float64) - p[n]
        d64 = o64.p[n] - p[n]
        nz = np.abs(d64) > 0
        flips_dev = float(np.mean(np.sign(du[nz]) != np.sign(d64[nz]))) if nz.any() else 0.0
        flips_o32 = float(np.mean(np.sign(d32[nz]) != np.sign(d64[nz]))) if nz.any() else 0.0
        print("%-32s shape %-14s |d64| mean %.3e  err dev %.3e  err o32 %.3e  sign flips dev %.4f o32 %.6f  zero-upd frac dev %.4f o64 %.4f"
              % (n, got[n].shape, np.mean(np.abs(d64)), np.linalg.norm(du - d64) / max(np.linalg.norm(d64), 1e-300),
                 np.linalg.norm(d32 - d64) / max(np.linalg.norm(d64), 1e-300), flips_dev, flips_o32,
                 float(np.mean(du == 0)), float(np.mean(d64 == 0))))
        if n.endswith("layer_0/kernel"):
            for name, sl in (("id row", slice(0, 1)), ("profile rows", slice(1, None))):
                a, c = du[sl], d64[sl]
                print("      %-14s err dev %.3e" % (name, np.linalg.norm(a - c) / max(np.linalg.norm(c), 1e-300)))


if __name__ == "__main__":
    main()
