"""Validation-driven early stopping used by GANMF.fit / DisGANMF.fit -- the host-side state machine of the
reference's `Utils_.EarlyStoppingScheduler` (Utils_.py:25-88), written for this package (the reference module
itself cannot be imported: it pulls seaborn / matplotlib at import time).

Contract kept: called once per epoch with the 1-based epoch number; a validation runs
when `epoch > after` and `epoch % freq == 0` and reads the watched metrics at cutoff 5 (hard-coded in the
reference, Utils_.py:64); a validation is a STALL when no watched metric exceeds the best vector seen so far
(i.e. all are <=); `allow_worse` stalls are tolerated, the next one calls `model.stop_fit()` and
`model.load_model()`; any other validation becomes the new best vector, refills the tolerance and calls
`model.save_current_model()`."""
import numpy as np


class EarlyStoppingScheduler(object):
    _CUTOFF = 5

    def __init__(self, model, evaluator, metrics=['PRECISION', 'RECALL', 'MAP', 'NDCG'], freq=1, allow_worse=5,
                 after=0):
        self.model, self.evaluator = model, evaluator
        self.metrics, self.freq, self.after, self.allow_worse = metrics, freq, after, allow_worse
        self.scores = []                                   # every validation vector, oldest first
        self.best_scores = np.zeros(len(metrics))          # starts at 0: an all-zero validation is a stall
        self.reset()

    # -- the public surface the reference's callers use ---------------------------------------------
    def __call__(self, epoch):
        if epoch > self.after:
            self.score(epoch)

    def reset(self):
        self.worse_left = self.allow_worse

    def load_best(self):
        self.model.load_model()

    def get_scores(self):
        return self.scores

    def score(self, epoch):
        if epoch % self.freq:
            return
        results, _ = self.evaluator.evaluateRecommender(self.model)
        now = np.array([results[self._CUTOFF][name] for name in self.metrics])
        self.scores.append(now)
        stalled = bool(np.all(now <= self.best_scores))
        if not stalled:
            self.best_scores = now
            self.reset()
            self.model.save_current_model()
        elif self.worse_left > 0:
            self.worse_left -= 1
        else:
            self.model.stop_fit()
            self.load_best()
