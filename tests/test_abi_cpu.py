"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/ganmf_b200.h declares, the ctypes binding covers exactly that set, and compute entry points
fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from ganmf_b200 import build
    return build.build()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ganmf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ganmf_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_build())
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_binding_matches_header():
    from ganmf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _build()
    _lib.load()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    from ganmf_b200 import _lib
    from ganmf_b200.GANRec.GANMF import GANMF
    m = sps.random(30, 20, 0.2, format="csr", dtype=np.float32)
    rec = GANMF(m, is_experiment=True)
    with pytest.raises(_lib.GanmfError, match="no CUDA device"):
        rec.fit(epochs=1, batch_size=8)
    with pytest.raises(RuntimeError):
        rec.recommend(np.arange(3), cutoff=5)


def test_product_never_imports_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "ganmf_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_constructor_semantics():
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.GANRec.GANMF import GANMF
    m = sps.random(30, 20, 0.2, format="csr", dtype=np.float32)
    with pytest.raises(ValueError):
        GANMF(m, mode="both", is_experiment=True)
    g = GANMF(m, mode="item", is_experiment=True)
    assert (g.num_users, g.num_items) == (20, 30) and g.URM_train.shape == (20, 30)   # GANMF.py:32-36
    d = DisGANMF(m, "user", 7, False, True)           # positional order of DisGANMF.py:24
    assert d.seed == 7 and d.RECOMMENDER_NAME == "DisGANMF" and g.RECOMMENDER_NAME == "GANMF"
