#!/usr/bin/env python
"""Times ganmf_evaluate / ganmf_recommend on the cfg4 shape (run plain, or under ncu for a launch list)."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_users, n_items = 20000, 27000
    urm = bench.synthetic_urm(n_users, n_items, 0.005, 1)
    test = bench.synthetic_urm(n_users, n_items, 0.00125, 2)
    test = sps.csr_matrix(test - test.multiply(urm))
    test.eliminate_zeros()
    test.sort_indices()
    eng = Engine(L.KIND_GANMF, n_users, n_items, 250, emb_dim=64, max_batch=64)
    eng.set_csr(L.CSR_SEEN, urm, with_data=False)
    eng.init_params(1)
    eng.set_test(test, urm)
    users = np.flatnonzero(np.diff(test.indptr) > 0)[:8192].astype(np.int32)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sums, _ = eng.evaluate(users, [10], remove_seen=True, want_counts=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("evaluate %d users: %.2f ms -> %.0f users/s" % (len(users), dt * 1e3, len(users) / dt))
    for rep in range(2):
        t0 = time.perf_counter()
        idx, val, _ = eng.recommend(users[:1000], 10, remove_seen=True)
        dt = time.perf_counter() - t0
        print("recommend 1000 users: %.2f ms" % (dt * 1e3))


if __name__ == "__main__":
    main()
