"""ganmf_b200 -- B200-native (sm_100a) GANMF / DisGANMF training, scoring and evaluation behind the
reference's recommender API (edervishaj/GANMF).  Import paths mirror the reference:
    from ganmf_b200.GANRec.GANMF import GANMF
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
The CUDA library (libganmf_b200.so) is loaded on first use; there is no CPU fallback."""
__all__ = ["GANMF", "DisGANMF", "EvaluatorHoldout", "EarlyStoppingScheduler"]


def __getattr__(name):
    if name == "GANMF":
        from .GANRec.GANMF import GANMF
        return GANMF
    if name == "DisGANMF":
        from .GANRec.DisGANMF import DisGANMF
        return DisGANMF
    if name == "EvaluatorHoldout":
        from .Base.Evaluation.Evaluator import EvaluatorHoldout
        return EvaluatorHoldout
    if name == "EarlyStoppingScheduler":
        from .Utils_ import EarlyStoppingScheduler
        return EarlyStoppingScheduler
    raise AttributeError(name)
