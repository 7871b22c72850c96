"""Fused scorer of the evaluator / recommend() (csrc/score_select.cuh): one TF32 pass with top-K candidate lists
in the GEMM epilogue, exact re-scoring, per-row certificate, exact fallback.

Contract: the returned lists ARE the top K (ties -> lowest item index) of the correctly rounded fp32 score row
fl32(sum_k fp64(p_k * v_k)) with the seen items removed -- bit-exact, whichever of the two device routes (certified
candidates / exact fallback) produced them.  The host side of these tests computes that score row in float64."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import eval_oracle as eo

pytestmark = pytest.mark.gpu


def exact_scores(P, V):
    return (P.astype(np.float64) @ V.astype(np.float64).T).astype(np.float32)


def host_topk(scores, train, users, K, remove_seen=True):
    out = np.full((len(users), K), -1, dtype=np.int32)
    for r, u in enumerate(users):
        s = scores[r].copy()
        if remove_seen:
            s[train.indices[train.indptr[u]:train.indptr[u + 1]]] = -np.inf
        order = np.argsort(-s, kind="stable")[:K]           # ties -> lowest index
        order = order[np.isfinite(s[order])]
        out[r, :len(order)] = order
    return out


def make_engine(n_users, n_items, k, P, V, train, item_mode=False):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rows, width = (n_items, n_users) if item_mode else (n_users, n_items)
    e = Engine(L.KIND_MF, rows, width, k, max_batch=1, item_mode=item_mode)
    e.set_csr(L.CSR_SEEN, train, with_data=False)
    e.set_param("generator/user_embeddings", V if item_mode else P)      # row factors of the TRAINING orientation
    e.set_param("generator/item_embeddings", P if item_mode else V)
    return e


@pytest.mark.parametrize("n_users,n_items,k,K,item_mode", [
    (1000, 27000, 250, 10, False),        # cfg4 width, full k, KP = 16
    (777, 5003, 33, 10, False),           # ragged everything: k tail (2 k-blocks), last item tile, last row block
    (600, 4100, 96, 20, False),           # KP = 32
    (500, 3000, 64, 5, True),             # item mode: the ranked side is the user-factor matrix of the training view
    (300, 200, 16, 10, False),            # fewer items than one tile
])
def test_fused_lists_equal_exact_topk(n_users, n_items, k, K, item_mode):
    rs = np.random.RandomState(n_items)
    P = (rs.standard_normal((n_users, k)) * 0.3).astype(np.float32)
    V = (rs.standard_normal((n_items, k)) * 0.3).astype(np.float32)
    train = sps.random(n_users, n_items, min(0.01, 40.0 / n_items), format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    train.sort_indices()
    e = make_engine(n_users, n_items, k, P, V, train, item_mode)
    users = np.sort(rs.permutation(n_users)[:min(n_users, 700)]).astype(np.int32)
    sc = exact_scores(P[users], V)
    for remove_seen in (True, False):
        idx, val, _ = e.recommend(users, K, remove_seen=remove_seen)
        want = host_topk(sc, train, users, K, remove_seen)
        assert np.array_equal(idx, want)
        ok = want >= 0
        assert np.array_equal(val[ok], np.take_along_axis(sc, np.maximum(want, 0), 1)[ok])   # values = exact scores
    fused, fb = e.eval_stats()
    assert fused == 2 * len(users)
    assert fb <= 0.02 * fused, (fused, fb)         # the certificate holds for (nearly) every row of a generic model
    e.close()


def test_fused_ties_and_degenerate_rows_fall_back_to_the_exact_path():
    """Exact ties across the K-th place (duplicated item rows, all-zero users) cannot be certified from candidate
    lists; those rows must take the exact route and still follow the lowest-index rule."""
    rs = np.random.RandomState(5)
    n_users, n_items, k, K = 400, 3000, 48, 10
    P = (rs.standard_normal((n_users, k)) * 0.3).astype(np.float32)
    V = (rs.standard_normal((n_items, k)) * 0.3).astype(np.float32)
    V[1500:] = V[:1500]                               # every item has an exact twin 1500 columns later
    P[:50] = 0                                        # 50 users with all scores equal
    train = sps.random(n_users, n_items, 0.01, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    train.sort_indices()
    e = make_engine(n_users, n_items, k, P, V, train)
    users = np.arange(n_users, dtype=np.int32)
    idx, _, _ = e.recommend(users, K, remove_seen=True)
    assert np.array_equal(idx, host_topk(exact_scores(P, V), train, users, K))
    fused, fb = e.eval_stats()
    assert fb >= 50                                   # at least the all-equal rows went the exact way
    e.close()


def test_fused_rows_with_fewer_than_k_rankable_items():
    rs = np.random.RandomState(9)
    n_users, n_items, k, K = 64, 300, 16, 10
    P = rs.standard_normal((n_users, k)).astype(np.float32)
    V = rs.standard_normal((n_items, k)).astype(np.float32)
    dense = np.zeros((n_users, n_items), dtype=np.float32)
    dense[:, :] = 1
    for u in range(n_users):                          # user u has only u % 13 unseen items
        dense[u, rs.permutation(n_items)[:u % 13]] = 0
    train = sps.csr_matrix(dense)
    e = make_engine(n_users, n_items, k, P, V, train)
    users = np.arange(n_users, dtype=np.int32)
    idx, val, _ = e.recommend(users, K, remove_seen=True)
    want = host_topk(exact_scores(P, V), train, users, K)
    assert np.array_equal(idx, want)
    assert np.all(np.isneginf(val[want < 0]))
    e.close()


def test_fused_evaluate_sums_bit_exact_vs_oracle(monkeypatch):
    """ganmf_evaluate on the fused route == the oracle (pinned to the reference evaluator) fed with the exact
    score rows; and == the materialised route (GANMF_EVAL_FUSED=0) wherever the two score definitions rank alike."""
    from ganmf_b200 import _lib as L
    rs = np.random.RandomState(11)
    n_users, n_items, k = 900, 2500, 40
    P = (rs.standard_normal((n_users, k)) * 0.4).astype(np.float32)
    V = (rs.standard_normal((n_items, k)) * 0.4).astype(np.float32)
    train = sps.random(n_users, n_items, 0.02, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    test = sps.random(n_users, n_items, 0.01, format="csr", dtype=np.float32, random_state=rs)
    test = sps.csr_matrix(test - test.multiply(train))
    test.eliminate_zeros()
    test.data[:] = rs.randint(1, 6, size=test.nnz).astype(np.float32)       # explicit ratings: gains and RMSE matter
    test.sort_indices()
    train.sort_indices()
    e = make_engine(n_users, n_items, k, P, V, train)
    e.set_test(test, train)
    users = eo.users_to_evaluate(test).astype(np.int32)
    cut = [5, 10, 20]
    sums, counts = e.evaluate(users, cut, remove_seen=True, block_size=256)   # several blocks, ragged last one
    assert e.eval_stats()[0] == len(users)
    sc = exact_scores(P, V)
    ores, n_eval = eo.evaluate(lambda u: sc[u], train, test, cut, promotion="legacy")
    assert n_eval == len(users)
    for ci, c in enumerate(cut):
        for m in ("PRECISION", "RECALL", "PRECISION_RECALL_MIN_DEN", "MAP", "NDCG", "MRR", "ROC_AUC", "HIT_RATE",
                  "NOVELTY", "AVERAGE_POPULARITY"):
            assert sums[ci, L.MC_NAMES.index(m)] == float(ores[c]["_sums"][m]), (c, m)
        assert sums[ci, L.MC_NAMES.index("RMSE")] == pytest.approx(float(ores[c]["_sums"]["RMSE"]), rel=1e-12)
        assert np.array_equal(counts[ci], ores[c]["_counts"])
    e.close()


def test_tf32_error_model_of_the_certificate():
    """The certificate needs |tf32 score - exact| <= gamma * ||p|| * ||v||.  Measure the left side with the
    tensor-core GEMM primitive (same TMA rounding, same MMA) on data with a wide dynamic range."""
    import ctypes as C
    import torch
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(2)
    n, m, k = 512, 1024, 250
    P = (rs.standard_normal((n, k)) * np.exp(rs.standard_normal((n, 1)))).astype(np.float32)
    V = (rs.standard_normal((m, k)) * np.exp(rs.standard_normal((1, k)))).astype(np.float32)
    e = Engine(L.KIND_MF, 8, 8, 8, max_batch=1)
    ld = 256
    a = torch.zeros((n, ld), device="cuda"); a[:, :k] = torch.from_numpy(P)
    b = torch.zeros((m, ld), device="cuda"); b[:, :k] = torch.from_numpy(V)
    out = torch.empty((n, m), device="cuda")
    L.check(e.lib.ganmf_k_gemm(e.ctx, C.c_void_p(a.data_ptr()), ld, 0, C.c_void_p(b.data_ptr()), ld, 0, n, m, k,
                               C.c_void_p(out.data_ptr()), m, L.GEMM_TC))
    torch.cuda.synchronize()
    exact = P.astype(np.float64) @ V.astype(np.float64).T
    bound = np.linalg.norm(P.astype(np.float64), axis=1)[:, None] * np.linalg.norm(V.astype(np.float64), axis=1)[None, :]
    ratio = np.abs(out.cpu().numpy().astype(np.float64) - exact) / bound
    gamma = 1.05 * (2.0 / 1024.0 + (k + 8) * 1.2e-7)          # capi.cu: fused_gamma
    assert ratio.max() < 0.5 * gamma, ratio.max()             # round-to-nearest operands: <= 2^-10 even in the worst case
    e.close()


def test_mf_context_scores_like_a_ganmf_context():
    """GANMF_KIND_MF (factor matrices only; the device scorer behind BaseMatrixFactorizationRecommender and the
    evaluation context of item-sharded runs) == a GANMF context holding the same factors."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(4)
    n_users, n_items, k = 300, 700, 24
    P = rs.standard_normal((n_users, k)).astype(np.float32)
    V = rs.standard_normal((n_items, k)).astype(np.float32)
    train = sps.random(n_users, n_items, 0.03, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    a = make_engine(n_users, n_items, k, P, V, train)
    b = Engine(L.KIND_GANMF, n_users, n_items, k, emb_dim=8, max_batch=4)
    b.set_csr(L.CSR_SEEN, train, with_data=False)
    b.set_param("generator/user_embeddings", P)
    b.set_param("generator/item_embeddings", V)
    users = np.arange(0, n_users, 3, dtype=np.int32)
    assert np.array_equal(a.score(users), b.score(users))
    ia, va, sa = a.recommend(users, 30, remove_seen=True, return_scores=True)      # K > 24: materialised route
    ib, vb, sb = b.recommend(users, 30, remove_seen=True, return_scores=True)
    assert np.array_equal(ia, ib) and np.array_equal(sa, sb)
    assert [n for n, _, _, _ in a.param_infos()] == ["generator/user_embeddings", "generator/item_embeddings"]
    a.upload_ids(np.arange(4, dtype=np.int32))
    with pytest.raises(L.GanmfError, match="does not train"):
        a.d_step(0, 4, 1e-4, 0.0, 1.0)
    a.close()
    b.close()


def test_matrix_factorization_base_recommender_on_device():
    """Base/BaseMatrixFactorizationRecommender.py:94-143 mirror: factors set by a subclass' fit(), ranking and
    evaluation on the device; biases and cold users take the host-edited score route of the reference."""
    from ganmf_b200.Base.BaseMatrixFactorizationRecommender import BaseMatrixFactorizationRecommender
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    rs = np.random.RandomState(21)
    n_users, n_items, k = 240, 900, 20
    train = sps.random(n_users, n_items, 0.03, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    train = train.tolil()
    train[7, :] = 0                                     # a cold user
    train = sps.csr_matrix(train)
    train.eliminate_zeros()
    test = sps.random(n_users, n_items, 0.02, format="csr", dtype=np.float32, random_state=rs)
    test = sps.csr_matrix(test - test.multiply(train))
    test.eliminate_zeros()
    test.data[:] = 1

    class PureSVDLike(BaseMatrixFactorizationRecommender):
        RECOMMENDER_NAME = "PureSVDLike"

        def fit(self):
            self.USER_factors = rs.standard_normal((n_users, k)).astype(np.float32)
            self.ITEM_factors = rs.standard_normal((n_items, k)).astype(np.float32)

    rec = PureSVDLike(train)
    rec.fit()
    sc = exact_scores(rec.USER_factors, rec.ITEM_factors)
    warm = np.array([u for u in range(40) if u != 7], dtype=np.int32)
    lists = rec.recommend(warm, cutoff=10)
    want = host_topk(sc[warm], train, warm, 10)
    assert lists == [[int(i) for i in row if i >= 0] for row in want]
    assert rec.recommend(np.array([7]), cutoff=10) == [[]]                    # cold user: -inf everywhere (:124-139)
    raw = rec._compute_item_score(np.array([3, 7]))
    assert np.all(np.isneginf(raw[1])) and np.allclose(raw[0], sc[3], rtol=1e-5, atol=1e-5)
    only = np.array([5, 17, 300])
    lim = rec._compute_item_score(np.array([3]), items_to_compute=only)      # :112-114
    assert np.all(np.isneginf(np.delete(lim[0], only))) and np.all(np.isfinite(lim[0, only]))
    # evaluator: equals the oracle on the recommender's own score rows (cold user included -> host-edited route)
    res, _ = EvaluatorHoldout(test, cutoff_list=[5, 10], exclude_seen=True).evaluateRecommender(rec)
    ores, _ = eo.evaluate(lambda u: rec._compute_item_score(u), train, test, [5, 10], promotion="legacy")
    for c in (5, 10):
        for m in ("PRECISION", "RECALL", "MAP", "NDCG", "MRR", "HIT_RATE", "COVERAGE_ITEM", "NOVELTY"):
            assert float(res[c][m]) == pytest.approx(float(ores[c][m]), rel=1e-12), (c, m)
    # biases (:119-122)
    rec.use_bias = True
    rec.ITEM_bias = rs.standard_normal(n_items).astype(np.float32)
    rec.USER_bias = rs.standard_normal(n_users).astype(np.float32)
    rec.GLOBAL_bias = np.float32(0.3)
    b = rec._compute_item_score(np.array([3]))
    assert np.allclose(b[0], sc[3] + rec.ITEM_bias + rec.GLOBAL_bias + rec.USER_bias[3], rtol=1e-5, atol=1e-5)
    got = rec.recommend(np.array([3]), cutoff=5)[0]
    s3 = b[0].copy()
    s3[train.indices[train.indptr[3]:train.indptr[4]]] = -np.inf
    assert got == [int(i) for i in np.argsort(-s3, kind="stable")[:5]]
