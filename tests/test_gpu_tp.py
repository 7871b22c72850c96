"""Item-sharded (tensor-parallel over items) GANMF training, SURVEY.md section 8f-3, on ONE GPU: the contexts of
all ranks live on the same device and the partial sums that NCCL would all-reduce are formed in place by
ItemShardedTrainer's in-process mode.  Same contract as the single-GPU parity tests: per-step losses and
parameters against the fp32 oracle stepping on the whole minibatch, rel <= 1e-3."""
import numpy as np
import pytest

from oracle import train_oracle as to
from tests.test_gpu_train_parity import HP, REL, make_urm, rel_err

pytestmark = pytest.mark.gpu

WE, BE, WD, BD = to.GANMF_D
P_, V_ = to.GANMF_G


def slice_params(p, lo, hi):
    return {WE: p[WE][lo:hi], BE: p[BE], WD: p[WD][:, lo:hi], BD: p[BD][lo:hi], P_: p[P_], V_: p[V_][lo:hi]}


def gather_params(parts):
    return {WE: np.concatenate([q[WE] for q in parts], 0), BE: parts[0][BE],
            WD: np.concatenate([q[WD] for q in parts], 1), BD: np.concatenate([q[BD] for q in parts]),
            P_: parts[0][P_], V_: np.concatenate([q[V_] for q in parts], 0)}


def make_engines(urm, k, E, B, world, gemm_path=None):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    from ganmf_b200.parallel import item_slices
    n_rows, width = urm.shape
    engines = []
    for r, (lo, hi) in enumerate(item_slices(width, world)):
        e = Engine(L.KIND_GANMF, n_rows, hi - lo, k, emb_dim=E, max_batch=B, global_width=width, item_offset=lo,
                   tp_rank=r, tp_world=world, gemm_path=L.GEMM_AUTO if gemm_path is None else gemm_path)
        e.set_csr(L.CSR_TRAIN, urm[:, lo:hi].tocsr())
        engines.append(e)
    return engines, item_slices(width, world)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("hp", [HP, dict(HP, m=0.05, g_reg=1e-3, alpha=0.3)])     # gate open / closed, dense P update
# sparse real codes x low-rank generator route x low-rank encoder weight gradient (ranks of >= 4-GPU groups)
@pytest.mark.parametrize("routes", ["000", "100", "010", "110", "011", "111"])
def test_item_sharded_steps_parity(monkeypatch, world, hp, routes):
    from ganmf_b200 import _lib as L
    monkeypatch.setenv("GANMF_SPARSE_REAL", routes[0])
    monkeypatch.setenv("GANMF_LOWRANK", routes[1])
    monkeypatch.setenv("GANMF_LOWRANK_DWE", routes[2])
    from ganmf_b200.parallel import ItemShardedTrainer
    n_rows, width, k, E, B, epochs = 300, 517, 24, 40, 64, 10
    urm = make_urm(n_rows, width, 0.05, 0)
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=1)
    rs = np.random.RandomState(2)
    p0[BE] = (rs.standard_normal(E) * 0.01).astype(np.float32)
    p0[BD] = (rs.standard_normal(width) * 0.01).astype(np.float32)
    engines, slices = make_engines(urm, k, E, B, world, L.GEMM_TC)
    for e, (lo, hi) in zip(engines, slices):
        e.set_params(slice_params(p0, lo, hi))
        e.reset_optimizers()
    tr = ItemShardedTrainer(engines)
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    dl_all, gl_all, odl, ogl = [], [], [], []
    for _, batches in to.epoch_index_stream(n_rows, B, epochs, seed=1337):
        dl, gl = tr.train_epoch(np.concatenate(batches).astype(np.int32), B, 1, 1, hp)
        dl_all += list(dl)
        gl_all += list(gl)
        for b in batches:
            odl.append(orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"], m=hp["m"]))
        for b in batches:
            ogl.append(orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"]))
    assert len(dl_all) == 50
    np.testing.assert_allclose(dl_all, odl, rtol=REL)
    np.testing.assert_allclose(gl_all, ogl, rtol=REL)
    parts = [e.get_params() for e in engines]
    for q in parts[1:]:                              # replicated tensors stay bit-identical on every rank
        assert np.array_equal(q[P_], parts[0][P_]) and np.array_equal(q[BE], parts[0][BE])
    got = gather_params(parts)
    for n in orc.p:
        assert rel_err(got[n], orc.p[n]) <= REL, (n, rel_err(got[n], orc.p[n]))
    for e in engines:
        e.close()


def test_item_sharded_init_equals_unsharded_init():
    """ganmf_init_params draws element (r, c) of a slice from the position it has in the whole tensor: a sharded
    run and a single-GPU run with the same seed start from the same model."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_rows, width, k, E, B = 200, 333, 16, 48, 32
    urm = make_urm(n_rows, width, 0.05, 3)
    one = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B)
    one.init_params(77)
    want = one.get_params()
    one.close()
    engines, _ = make_engines(urm, k, E, B, 3)
    for e in engines:
        e.init_params(77)
    got = gather_params([e.get_params() for e in engines])
    for n in want:
        assert np.array_equal(got[n], want[n]), n
    for e in engines:
        e.close()


def test_item_sharded_matches_single_gpu_on_larger_shapes():
    """tcgen05 path with CTA pairs, split-K and ragged slices (I = 4100 over 3 ranks, E = 256, B = 512): the
    sharded run against ONE context stepping on the same minibatches (both TF32; they differ by summation order)."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    from ganmf_b200.parallel import ItemShardedTrainer
    n_rows, width, k, E, B = 2048, 4100, 64, 256, 512
    urm = make_urm(n_rows, width, 0.02, 5)
    hp = HP
    one = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B)
    one.set_csr(L.CSR_TRAIN, urm)
    one.init_params(11)
    engines, _ = make_engines(urm, k, E, B, 3)
    for e in engines:
        e.init_params(11)
    tr = ItemShardedTrainer(engines)
    rs = np.random.RandomState(0)
    for _ in range(3):
        perm = rs.permutation(n_rows).astype(np.int32)
        dl1, gl1 = one.train_epoch(perm, B, 1, 1, hp["d_lr"], hp["g_lr"], hp["d_reg"], hp["g_reg"], hp["m"], hp["alpha"])
        dl, gl = tr.train_epoch(perm, B, 1, 1, hp)
        np.testing.assert_allclose(dl, dl1, rtol=REL)
        np.testing.assert_allclose(gl, gl1, rtol=REL)
    want = one.get_params()
    got = gather_params([e.get_params() for e in engines])
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))
    one.close()
    for e in engines:
        e.close()


def test_unsharded_entry_points_refuse_a_sharded_context():
    from ganmf_b200 import _lib as L
    urm = make_urm(64, 96, 0.1, 1)
    engines, _ = make_engines(urm, 8, 16, 16, 2)
    engines[0].upload_ids(np.arange(16, dtype=np.int32))
    with pytest.raises(L.GanmfError, match="item-sharded"):
        engines[0].d_step(0, 16, 1e-4, 0.0, 1.0)
    for e in engines:
        e.close()
