"""Host-side state machines against the oracle's restatement (no GPU needed)."""
import numpy as np

from ganmf_b200.Base.Incremental_Training_Early_Stopping import Incremental_Training_Early_Stopping
from ganmf_b200.Utils_ import EarlyStoppingScheduler
from oracle.eval_oracle import EarlyStoppingOracle


class FakeModel(object):
    def __init__(self):
        self.log = []
        self.stopped = False

    def stop_fit(self):
        self.stopped = True
        self.log.append("stop")

    def load_model(self):
        self.log.append("load")

    def save_current_model(self):
        self.log.append("save")


class FakeEvaluator(object):
    def __init__(self, seq):
        self.seq, self.i = list(seq), 0

    def evaluateRecommender(self, model):
        v = self.seq[self.i]
        self.i += 1
        return {5: {"MAP": v, "NDCG": v / 2}}, ""


def run(cls_factory, seq, **kw):
    model = FakeModel()
    sched = cls_factory(model, seq, **kw)
    epoch = 1
    while not model.stopped and epoch <= 40:
        sched(epoch)
        epoch += 1
    return model.log, epoch - 1


def test_scheduler_matches_oracle_state_machine():
    seqs = [[0.1, 0.2, 0.2, 0.15, 0.1, 0.05, 0.3, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1] * 3,
            [0.0] * 40, [0.5, 0.4, 0.3, 0.6] + [0.1] * 40]
    for seq in seqs:
        for allow_worse, freq, after in [(0, 1, 0), (2, 1, 0), (2, 3, 4), (5, 5, 0)]:
            a = run(lambda m, s, **k: EarlyStoppingScheduler(m, FakeEvaluator(s), **k), seq,
                    metrics=["MAP", "NDCG"], freq=freq, allow_worse=allow_worse, after=after)
            ev = FakeEvaluator(seq)
            b = run(lambda m, s, **k: EarlyStoppingOracle(m, lambda mm: ev.evaluateRecommender(mm)[0], **k), seq,
                    metrics=["MAP", "NDCG"], freq=freq, allow_worse=allow_worse, after=after)
            assert a == b, (seq[:6], allow_worse, freq, after)


def test_scheduler_stops_on_allow_worse_plus_one():
    log, last = run(lambda m, s, **k: EarlyStoppingScheduler(m, FakeEvaluator(s), **k),
                    [0.3, 0.2, 0.2, 0.2, 0.2], metrics=["MAP"], freq=1, allow_worse=2, after=0)
    assert log == ["save", "stop", "load"] and last == 4       # bad evals at 2,3,4 -> third one stops


class Incr(Incremental_Training_Early_Stopping):
    def __init__(self):
        self.ran, self.best_at = [], []

    def _run_epoch(self, n):
        self.ran.append(n)

    def _prepare_model_for_validation(self):
        pass

    def _update_best_model(self):
        self.best_at.append(len(self.ran))


class SeqEval(object):
    def __init__(self, seq):
        self.seq, self.i = seq, 0

    def evaluateRecommender(self, m):
        v = self.seq[self.i]
        self.i += 1
        return {10: {"MAP": v}, 20: {"MAP": -1}}, ""


def test_incremental_mixin_semantics():
    m = Incr()
    m._train_with_early_stopping(50, epochs_min=0, validation_every_n=2, stop_on_validation=True,
                                 validation_metric="MAP", lower_validations_allowed=2,
                                 evaluator_object=SeqEval([0.1, 0.3, 0.2, 0.3, 0.1]))
    assert m.ran == list(range(8))          # validations at epochs 2,4,6,8 -> second non-improvement stops
    assert m.epochs_best == 4 and m.best_at == [2, 4]
    assert m.get_early_stopping_final_epochs_dict() == {"epochs": 4}
    m2 = Incr()
    m2._train_with_early_stopping(3)
    assert m2.ran == [0, 1, 2] and m2.epochs_best == 2 and m2.best_at == [3]
