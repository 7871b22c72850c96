#!/usr/bin/env python
"""cfg5-scale run (BASELINE.json configs[4]: 2M users x 200k items, 0.1 % density, k=250, E=1024, B=1024):
one GPU's share of the 8-way user partition (250 000 users, all 200 000 items, replicated D and V).
Reports train rows/s, the tcgen05 GEMM rate, the top-10 evaluator and the top-k kernel at I = 200 000.

    python tools/run_cfg5.py [--steps 6] [--users 250000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--users", type=int, default=250000)
    ap.add_argument("--eval-users", type=int, default=4000)
    a = ap.parse_args()
    import torch
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    c = dict(users=a.users, items=200000, density=0.001, k=250, E=1024, B=1024)
    t0 = time.time()
    urm = bench.synthetic_urm(c["users"], c["items"], c["density"], 1337)
    t_gen = time.time() - t0
    eng = Engine(L.KIND_GANMF, c["users"], c["items"], c["k"], emb_dim=c["E"], max_batch=c["B"])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_csr(L.CSR_SEEN, urm, with_data=False)
    eng.init_params(1234)
    K, W, B = a.steps, 3, c["B"]
    perm = np.random.RandomState(1).permutation(c["users"])[:(K + W) * B].astype(np.int32)
    eng.upload_ids(perm)
    HP = bench.HP

    def run(first, count):
        for i in range(count):
            eng.d_step((first + i) * B, B, HP["d_lr"], HP["d_reg"], HP["m"], loss_slot=i)
        for i in range(count):
            eng.g_step((first + i) * B, B, HP["g_lr"], HP["g_reg"], HP["alpha"], loss_slot=count + i)
    run(0, W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(W, K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    losses = eng.read_losses(2 * K)
    eng.profile(True)
    run(W, K)
    rec_ms, rec_shape = eng.profile_records()
    gemm_ms, gemm_flops, n_gemm = eng.profile_read()
    eng.profile(False)
    pk = bench.peaks()
    out = {"workload": "cfg5 shard: %d users x 200000 items, 0.1%% density, k=250, E=1024, B=1024" % c["users"],
           "csr_nnz": int(urm.nnz), "host_gen_s": t_gen, "rows_per_s": B * K / (ms * 1e-3), "ms_per_step": ms / K,
           "gemm_tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12, "gemm_share": gemm_ms / ms,
           "step_algorithmic_tflops": c["items"] * (8 * c["k"] + 30 * c["E"]) * B * K / (ms * 1e-3) / 1e12,
           "losses_finite": bool(np.all(np.isfinite(losses))), "loss_last": [float(losses[K - 1]), float(losses[-1])],
           "mem_allocated_gb": (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2 ** 30}
    agg = {}
    for t, s in zip(rec_ms, rec_shape):
        k_ = tuple(int(x) for x in s)
        v = agg.setdefault(k_, [0, 0.0])
        v[0] += 1
        v[1] += t
    out["gemm_table"] = [{"M": k_[0], "N": k_[1], "K": k_[2], "splits": k_[3], "ms": v[1] / v[0],
                          "tflops": 2.0 * k_[0] * k_[1] * k_[2] / (v[1] / v[0] * 1e-3) / 1e12}
                         for k_, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    # evaluator on this shard
    test = bench.synthetic_urm(c["users"], c["items"], c["density"] / 4, 4242)
    test = sps.csr_matrix(test - test.multiply(urm))
    test.eliminate_zeros()
    test.sort_indices()
    eng.set_test(test, urm)
    users = np.flatnonzero(np.diff(test.indptr) > 0)[:a.eval_users].astype(np.int32)
    eng.evaluate(users, [10], remove_seen=True, want_counts=False)      # warm-up sizes the device buffers
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sums, _ = eng.evaluate(users, [10], remove_seen=True, want_counts=False)
    dt = time.perf_counter() - t0
    out["eval_users_per_s"] = len(users) / dt
    out["eval_hbm_frac_4I"] = len(users) / dt * 4 * c["items"] / 1e9 / pk["hbm"]
    # top-k kernel alone at I = 200000
    n, I = 1024, c["items"]
    sc = torch.randn((n, I), device="cuda", dtype=torch.float32)
    idx = torch.empty((n, 10), device="cuda", dtype=torch.int32)
    val = torch.empty((n, 10), device="cuda", dtype=torch.float32)
    f = lambda: L.check(eng.lib.ganmf_k_topk(eng.ctx, sc.data_ptr(), I, n, I, 10, idx.data_ptr(), val.data_ptr()))
    f()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 * 1e-3
    out["topk_kernel"] = {"GBps": n * I * 4 / t / 1e9, "frac_of_hbm": n * I * 4 / t / 1e9 / pk["hbm"],
                          "users_per_s": n / t}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
