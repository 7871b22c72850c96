#!/usr/bin/env python
"""Evaluator timeline: ganmf_evaluate on random factors at a given shape (run under ncu for the launch list).

    python tools/eval_profile.py --items 200000 --users 32768 [--k 250] [--reps 3]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=200000)
    ap.add_argument("--users", type=int, default=32768)
    ap.add_argument("--k", type=int, default=250)
    ap.add_argument("--density", type=float, default=0.001)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cutoff", type=int, default=10)
    ap.add_argument("--block", type=int, default=0)
    a = ap.parse_args()
    import torch
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    U, I = a.users, a.items
    per = max(1, int(round(a.density * I)))
    e = Engine(L.KIND_MF, U, I, a.k, max_batch=1)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.init_params(7)
    ip, ix = bench.synth_csr_device(torch, 0, U, I, per, bench.SALT_TRAIN)
    e.set_csr_device(L.CSR_SEEN, U, I, ip, ix)
    rows = torch.repeat_interleave(torch.arange(U, device=ix.device, dtype=torch.int64), (ip[1:] - ip[:-1]).long())
    tip, tix = bench.synth_csr_device(torch, 0, U, I, max(1, per // 4), bench.SALT_TEST, drop_keys=rows * I + ix.long())
    pop = torch.bincount(ix.long(), minlength=I).cpu().numpy()
    test = bench.csr_to_host(tip, tix, (U, I))
    e.set_test(test, item_popularity=pop)
    users = np.flatnonzero(np.diff(test.indptr) > 0).astype(np.int32)
    e.evaluate(users, [a.cutoff], remove_seen=True, want_counts=False, block_size=a.block)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter()
        if os.environ.get("EVAL_PROFILE_RANGE") == "1":
            torch.cuda.profiler.start()
        e.evaluate(users, [a.cutoff], remove_seen=True, want_counts=False, block_size=a.block)
        torch.cuda.synchronize()
        if os.environ.get("EVAL_PROFILE_RANGE") == "1":
            torch.cuda.profiler.stop()
        best = min(best, time.perf_counter() - t0)
    fused, fb = e.eval_stats()
    pk = bench.peaks()
    print("block=%d segs=%s " % (a.block, os.environ.get("GANMF_EVAL_SEGS", "auto")), end="")
    print("items=%d users=%d k=%d: %.3f ms, %.3g users/s, frac of 4*I HBM roofline %.3f, fused rows %d, fallback %d" %
          (I, len(users), a.k, best * 1e3, len(users) / best, len(users) / best * 4 * I / 1e9 / pk["hbm"], fused, fb))


if __name__ == "__main__":
    main()
