"""Kernel-level parity on the B200 (through the C ABI): GEMM (tcgen05 and SIMT) vs fp64, CSR gather
vs scipy, fused Adam vs the oracle's TF-Adam, top-k vs the oracle (bit-exact, ties -> lowest index)."""
import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    e = Engine(L.KIND_GANMF, 64, 96, 8, emb_dim=8, max_batch=16)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_pair():
    """GANMF_PAIR=2: every tcgen05 GEMM that legally can runs on CTA pairs (cta_group::2, 256 x 256 tiles)."""
    import os
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    old = os.environ.get("GANMF_PAIR")
    os.environ["GANMF_PAIR"] = "2"
    try:
        e = Engine(L.KIND_GANMF, 64, 96, 8, emb_dim=8, max_batch=16)
    finally:
        if old is None:
            del os.environ["GANMF_PAIR"]
        else:
            os.environ["GANMF_PAIR"] = old
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rup(x, a=32):
    return (x + a - 1) // a * a


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(200, 300, 100), (64, 3706, 250), (257, 130, 1100), (128, 4, 3000)])
def test_gemm_paths_match_fp64(eng, a_mn, b_mn, shape):
    from ganmf_b200 import _lib as L
    M, N, K = shape
    rs = np.random.RandomState(M + N + K + a_mn * 2 + b_mn)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = rs.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T

    def padded(x):          # row-major with ld = roundup(cols, 32), padding poisoned
        out = np.full((x.shape[0], rup(x.shape[1])), 1e30, dtype=np.float32)
        out[:, :x.shape[1]] = x
        return out
    As = padded(A.T if a_mn else A)
    Bs = padded(B.T if b_mn else B)
    dA, dB = dev(As), dev(Bs)
    ldo = rup(N)
    for path, tol in ((L.GEMM_SIMT, 2e-5), (L.GEMM_TC, 2e-3)):
        out = torch.zeros((M, ldo), dtype=torch.float32, device="cuda")
        L.check(eng.lib.ganmf_k_gemm(eng.ctx, dA.data_ptr(), As.shape[1], a_mn, dB.data_ptr(), Bs.shape[1], b_mn,
                                     M, N, K, out.data_ptr(), ldo, path))
        got = out.cpu().numpy()[:, :N].astype(np.float64)
        scale = np.sqrt(K)          # |A||B| row norms ~ sqrt(K)
        assert np.max(np.abs(got - want)) / scale < tol, (path, np.max(np.abs(got - want)) / scale)
        assert np.all(out.cpu().numpy()[:, N:] == 0)      # padding columns untouched


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 700, 260), (1000, 520, 100), (129, 300, 40), (2048, 3000, 96)])
def test_gemm_cta_pairs_match_fp64(eng_pair, a_mn, b_mn, shape):
    """The cta_group::2 tiles (two SMs per 256 x 256 tile, B split between them): ragged M / N / K tails, tiles
    whose second CTA is entirely out of range, several tiles per pair, all operand major-ness combinations."""
    from ganmf_b200 import _lib as L
    M, N, K = shape
    rs = np.random.RandomState(M + N + K + a_mn * 2 + b_mn)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = rs.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T

    def padded(x):
        out = np.full((x.shape[0], rup(x.shape[1])), 1e30, dtype=np.float32)
        out[:, :x.shape[1]] = x
        return out
    As = padded(A.T if a_mn else A)
    Bs = padded(B.T if b_mn else B)
    dA, dB = dev(As), dev(Bs)
    ldo = rup(N)
    out = torch.zeros((M, ldo), dtype=torch.float32, device="cuda")
    for _ in range(2):                                     # twice: barrier phases / TMEM hand-back between launches
        L.check(eng_pair.lib.ganmf_k_gemm(eng_pair.ctx, dA.data_ptr(), As.shape[1], a_mn, dB.data_ptr(), Bs.shape[1],
                                          b_mn, M, N, K, out.data_ptr(), ldo, L.GEMM_TC))
    got = out.cpu().numpy()[:, :N].astype(np.float64)
    assert np.max(np.abs(got - want)) / np.sqrt(K) < 2e-3
    assert np.all(out.cpu().numpy()[:, N:] == 0)


def test_csr_gather_dense_matches_scipy():
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(0)
    for n_rows, width, B in [(300, 3706, 64), (50, 70, 24), (40, 5000, 7)]:
        urm = sps.random(n_rows, width, 0.05, format="csr", dtype=np.float32, random_state=rs)
        urm.data[:] = rs.randint(1, 6, size=urm.nnz)
        e = Engine(L.KIND_GANMF, n_rows, width, 4, emb_dim=4, max_batch=B)
        e.set_csr(L.CSR_TRAIN, urm)
        ids = rs.permutation(n_rows)[:B].astype(np.int32)
        e.upload_ids(ids)
        ld = rup(width)
        out = torch.full((B, ld), 7.0, dtype=torch.float32, device="cuda")
        L.check(e.lib.ganmf_k_csr_gather_dense(e.ctx, 0, B, out.data_ptr(), ld))
        got = out.cpu().numpy()
        assert np.array_equal(got[:, :width], urm[ids].toarray())
        assert np.all(got[:, width:] == 0)
        e.close()


def test_fused_adam_matches_tf_adam_oracle(eng):
    from ganmf_b200 import _lib as L
    from oracle.train_oracle import TFAdam
    rs = np.random.RandomState(1)
    n = 4096 + 64
    theta = rs.standard_normal(n).astype(np.float32) * 0.05
    params = {"w": theta.copy()}
    opt = TFAdam(params, ["w"], 1e-3, np.float32)
    t, m, v = dev(theta), dev(np.zeros(n, np.float32)), dev(np.zeros(n, np.float32))
    b1p, b2p = np.float32(0.9), np.float32(0.999)
    reg = 1e-2
    for step in range(20):
        g = rs.standard_normal(n).astype(np.float32) * (0.1 if step % 3 else 1e-4)
        alpha = np.float32(np.float32(1e-3) * np.sqrt(np.float32(1) - b2p) / (np.float32(1) - b1p))
        L.check(eng.lib.ganmf_k_adam(eng.ctx, t.data_ptr(), m.data_ptr(), v.data_ptr(), dev(g).data_ptr(), n,
                                     float(alpha), reg))
        opt.apply(params, {"w": g + np.float32(reg) * params["w"]})
        b1p, b2p = np.float32(b1p * np.float32(0.9)), np.float32(b2p * np.float32(0.999))
    got = t.cpu().numpy()
    # same formula, fp32; FMA contraction on the device moves last bits only
    np.testing.assert_allclose(got, params["w"], rtol=2e-5, atol=2e-7)


@pytest.mark.parametrize("n_items,K", [(3706, 50), (500, 5), (40, 50), (17632, 20), (200000, 10), (1000, 128)])
def test_topk_bit_exact_vs_oracle(eng, n_items, K):
    from ganmf_b200 import _lib as L
    from oracle.eval_oracle import topk_lowest_index
    rs = np.random.RandomState(n_items + K)
    n = 37
    s = rs.standard_normal((n, n_items)).astype(np.float32)
    s[3] = np.round(s[3] * 2) / 2                      # heavy ties
    s[4] = 0.0                                         # all equal -> indices 0..K-1
    s[5, rs.permutation(n_items)[: n_items - 3]] = -np.inf    # only 3 finite entries
    s[6] = -np.inf
    s[7] = np.sort(s[7])                               # ascending: worst case for threshold filters
    s[8] = np.sort(s[8])[::-1]
    ld = rup(n_items)
    sp = np.zeros((n, ld), dtype=np.float32)
    sp[:, :n_items] = s
    d = dev(sp)
    idx = torch.zeros((n, K), dtype=torch.int32, device="cuda")
    val = torch.zeros((n, K), dtype=torch.float32, device="cuda")
    L.check(eng.lib.ganmf_k_topk(eng.ctx, d.data_ptr(), ld, n, n_items, K, idx.data_ptr(), val.data_ptr()))
    gi, gv = idx.cpu().numpy(), val.cpu().numpy()
    kk = min(K, n_items)
    wi, wv = topk_lowest_index(s, kk)
    wi = np.where(np.isinf(wv) & (wv < 0), -1, wi)
    assert np.array_equal(gi[:, :kk], wi)
    assert np.array_equal(gv[:, :kk], wv)
    assert np.all(gi[:, kk:] == -1)


def test_csr_transposed_on_device_is_canonical_and_equals_scipy():
    """ganmf_set_csr_transposed (GANMF.py:32-33 `URM_train.T.tocsr()` on the GPU): indptr, indices (ascending inside every
    row of the transpose) and values bit-identical to scipy's canonical transpose -- empty columns, explicit ratings,
    implicit ones, a column longer than the shared-memory sort (6000 > 4096 entries) and one past the next power of two."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(3)
    for n_rows, n_cols, dens, with_data in [(300, 517, 0.05, True), (9000, 40, 0.02, False), (64, 1, 0.5, True)]:
        m = sps.random(n_rows, n_cols, dens, format="csr", dtype=np.float32, random_state=rs)
        m.data[:] = rs.randint(1, 6, size=m.nnz)
        m = m.tolil()
        if n_rows == 9000:
            m[rs.permutation(n_rows)[:6000], 7] = 3.0           # popular item: global-memory sort path
            m[rs.permutation(n_rows)[:4097], 9] = 2.0
            m[:, 11] = 0.0                                      # empty column
        m = sps.csr_matrix(m)
        m.eliminate_zeros()
        m.sort_indices()
        e = Engine(L.KIND_GANMF, n_cols, n_rows, 4, emb_dim=4, max_batch=8)
        e.set_csr_transposed(L.CSR_TRAIN, m, with_data=with_data)
        got = e.get_csr(L.CSR_TRAIN, with_data=with_data)
        want = m.T.tocsr()
        want.sort_indices()
        assert got.shape == want.shape
        assert np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices)
        if with_data:
            assert np.array_equal(got.data, want.data)
        e.close()


def test_csr_encode_rows_matches_float64():
    """csr_encode_rows_kernel alone: be + sum of the weight rows of the interactions, against float64."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(5)
    for n_rows, width, E, B in [(200, 3000, 100, 64), (50, 700, 1024, 17), (30, 90, 8, 30)]:
        urm = sps.random(n_rows, width, 0.03, format="csr", dtype=np.float32, random_state=rs)
        urm.data[:] = rs.randint(1, 6, size=urm.nnz)
        e = Engine(L.KIND_GANMF, n_rows, width, 4, emb_dim=E, max_batch=B)
        e.set_csr(L.CSR_TRAIN, urm)
        e.init_params(9)
        p = e.get_params()
        be = (rs.standard_normal(E) * 0.1).astype(np.float32)
        e.set_params({"autoencoder/encoding/bias": be})
        ids = rs.permutation(n_rows)[:B].astype(np.int32)
        e.upload_ids(ids)
        ldo = rup(E)
        out = torch.full((B, ldo), 7.0, dtype=torch.float32, device="cuda")
        L.check(e.lib.ganmf_k_csr_encode_rows(e.ctx, 0, B, out.data_ptr(), ldo))
        got = out.cpu().numpy()
        want = urm[ids].astype(np.float64) @ p["autoencoder/encoding/kernel"].astype(np.float64) + be
        assert np.max(np.abs(got[:, :E] - want)) <= 1e-5 * max(1.0, np.abs(want).max())
        assert np.all(got[:, E:] == 0)
        e.close()


@pytest.mark.parametrize("shape", [(256, 256, 64), (1024, 5000, 250), (300, 700, 250), (129, 40000, 96), (2048, 3000, 256),
                                   (64, 517, 24)])
def test_resident_a_generator_gemm_matches_fp64(eng_pair, shape):
    """gen_gemm.cuh: F = Pb . V^T with the A tile resident in shared memory (K <= 256, both operands K-major): ragged
    M / N / K tails, fewer rows than a pair tile, several row blocks and column segments per pair, launched twice
    (barrier phases / TMEM hand-back between launches); padding columns stay zero."""
    from ganmf_b200 import _lib as L
    M, N, K = shape
    rs = np.random.RandomState(M + N + K)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = rs.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T

    def padded(x):
        out = np.full((x.shape[0], rup(x.shape[1])), 1e30, dtype=np.float32)      # TMA must never read the padding
        out[:, :x.shape[1]] = x
        return out
    As, Bs = padded(A), padded(B)
    dA, dB = torch.from_numpy(As).cuda(), torch.from_numpy(Bs).cuda()
    ldo = rup(N)
    out = torch.zeros((M, ldo), dtype=torch.float32, device="cuda")
    for _ in range(2):
        L.check(eng_pair.lib.ganmf_k_gemm(eng_pair.ctx, dA.data_ptr(), As.shape[1], 0, dB.data_ptr(), Bs.shape[1], 0,
                                          M, N, K, out.data_ptr(), ldo, L.GEMM_RESIDENT_A))
    got = out.cpu().numpy()[:, :N].astype(np.float64)
    assert np.max(np.abs(got - want)) / np.sqrt(K) < 2e-3
    assert np.all(out.cpu().numpy()[:, N:] == 0)
