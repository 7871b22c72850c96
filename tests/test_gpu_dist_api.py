"""Multi-GPU behind the drop-in API: under a torch.distributed process group,
GANMF.fit() trains item-sharded and EvaluatorHoldout.evaluateRecommender() shards the users, and the caller's
code is the single-GPU code.  Two ranks are spawned on ONE device over gloo (NCCL refuses two ranks per GPU), so
the test runs on a single-GPU box; the NCCL path of the same trainer is exercised by tools/tp_parity.py and bench.py.

Checks: (1) the 2-rank fit follows the 1-rank fit (losses and weights within the TF32 parity tolerance: the two
differ only by the summation order over items); (2) with identical weights, the sharded evaluation returns the
single-GPU results bit for bit (the running sums are continued rank after rank)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu

FIT = dict(num_factors=16, emb_dim=32, epochs=6, batch_size=64, d_lr=1e-3, g_lr=5e-3, d_reg=1e-4, m=10,
           recon_coefficient=0.05, validation_set=None, sample_every=None, validation_evaluator=None)


def small_data(seed=0, n_users=260, n_items=333):
    rs = np.random.RandomState(seed)
    a, b = rs.standard_normal((n_users, 6)), rs.standard_normal((n_items, 6))
    full = (a @ b.T + 0.5 * rs.standard_normal((n_users, n_items))) > 1.8
    mask = rs.rand(n_users, n_items) < 0.75
    return sps.csr_matrix((full & mask).astype(np.float32)), sps.csr_matrix((full & ~mask).astype(np.float32))


def run(mode):
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    from ganmf_b200.GANRec.GANMF import GANMF
    train, test = small_data()
    np.random.seed(1337)
    rec = GANMF(train, mode=mode, seed=1337, is_experiment=True)
    rec.fit(**FIT)
    ev = EvaluatorHoldout(test, cutoff_list=[5, 10], exclude_seen=True)
    res, txt = ev.evaluateRecommender(rec)
    users = np.arange(0, 50)
    lists = rec.recommend(users, cutoff=7)
    return dict(d=rec.train_d_loss, g=rec.train_g_loss, res=res, txt=txt, params=rec.get_weights(), lists=lists,
                codes=rec.autoencoder_codes()[:20], sharded=rec._trainer is not None)


def _worker(rank, world, port, q, mode):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = run(mode)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["user", "item"])
def test_fit_and_evaluate_under_a_process_group(mode):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one = run(mode)                                    # the same code in this (single) process
    assert not one["sharded"] and outs[0]["sharded"] and outs[1]["sharded"]
    for r in (0, 1):
        np.testing.assert_allclose(outs[r]["d"], one["d"], rtol=1e-3)
        np.testing.assert_allclose(outs[r]["g"], one["g"], rtol=1e-3)
        for n, w in one["params"].items():
            g = outs[r]["params"][n]
            assert g.shape == w.shape
            err = np.linalg.norm(g.astype(np.float64) - w) / max(np.linalg.norm(w), 1e-30)
            assert err < 2e-3, (n, err)
        np.testing.assert_allclose(outs[r]["codes"], one["codes"], rtol=2e-2, atol=2e-3)
    assert outs[0]["res"] == outs[1]["res"] and outs[0]["txt"] == outs[1]["txt"]      # every rank gets the full result
    assert outs[0]["lists"] == outs[1]["lists"]
    # identical weights -> the sharded evaluation equals the single-GPU evaluation bit for bit
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    from ganmf_b200.GANRec.GANMF import GANMF
    train, test = small_data()
    rec = GANMF(train, mode=mode, seed=1337, is_experiment=True)
    rec.build(FIT["num_factors"], FIT["emb_dim"])
    rec._build_engine(FIT["batch_size"])
    rec.set_weights(outs[0]["params"])
    rec._finish_fit()
    res1, txt1 = EvaluatorHoldout(test, cutoff_list=[5, 10], exclude_seen=True).evaluateRecommender(rec)
    assert txt1 == outs[0]["txt"]
    for c in (5, 10):
        for m, v in res1[c].items():
            assert outs[0]["res"][c][m] == v or (np.isnan(v) and np.isnan(outs[0]["res"][c][m])), (c, m)
    assert rec.recommend(np.arange(0, 50), cutoff=7) == outs[0]["lists"]
