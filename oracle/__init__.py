"""CPU oracle for the GANMF / DisGANMF hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and there only as the checker or as the timed CPU
baseline.  The product (``ganmf_b200``) never imports this package and has no CPU
fallback.

Parity status:
  * evaluation path (seen mask -> top-k -> metrics): PINNED against the reference's
    own ``Base/Evaluation`` code run in the build container (tests/golden/*, made by
    tests/golden/make_golden.py), the reference's ``metrics_Test.py`` vectors and the
    surviving LastFM checkpoint + its stored ``test_results.pkl``.
  * training arithmetic (losses, gradients, TF-Adam): PARITY UNPINNED at step level --
    TensorFlow 1.12 cannot be installed here and the reference has no test of it.  The
    restatement follows the TF graph in GANRec/GANMF.py / GANRec/DisGANMF.py, is
    cross-checked against torch.autograd in fp64, and is pinned end to end only through
    the committed quality numbers (test_results/*).
"""
