"""Epoch loop with periodic validation, as the PoliMi base classes the reference builds on expect it
(reference: Base/Incremental_Training_Early_Stopping.py:93-259; `north_star` names it as part of the drop-in
surface).  Written for this package; the contract that is kept:

* the recommender supplies `_run_epoch(i)` (0-based), `_prepare_model_for_validation()` and
  `_update_best_model()` and calls `_train_with_early_stopping(...)`;
* three legal argument sets: no evaluator (train `epochs_max` epochs, the last model is the best one);
  evaluator + `validation_every_n` + `validation_metric` (validate, never stop early); the same plus
  `stop_on_validation=True` and `lower_validations_allowed` (stop early);
* a validation happens after every `validation_every_n`-th epoch, on the FIRST cutoff the evaluator returns;
  only a strictly larger metric value makes a new best model; training stops once
  `lower_validations_allowed` validations in a row failed to improve and at least `epochs_min` epochs ran;
* `self.epochs_best` (number of epochs of the best model; `epochs_max - 1` without an evaluator, as in the
  reference) feeds `get_early_stopping_final_epochs_dict()`; the return value is the number of epochs run."""


class Incremental_Training_Early_Stopping(object):
    def __init__(self):
        super(Incremental_Training_Early_Stopping, self).__init__()

    # -- hooks of the concrete recommender ------------------------------------------------------------
    def _run_epoch(self, num_epoch):
        raise NotImplementedError()

    def _prepare_model_for_validation(self):
        raise NotImplementedError()

    def _update_best_model(self):
        raise NotImplementedError()

    def get_early_stopping_final_epochs_dict(self):
        return {"epochs": self.epochs_best}

    # -- the loop ---------------------------------------------------------------------------------------
    @staticmethod
    def _check_arguments(name, epochs_max, epochs_min, validation_every_n, stop_on_validation, validation_metric,
                         lower_validations_allowed, evaluator_object):
        assert epochs_max > 0, "{}: Number of epochs_max must be > 0, passed was {}".format(name, epochs_max)
        assert epochs_min >= 0, "{}: Number of epochs_min must be >= 0, passed was {}".format(name, epochs_min)
        assert epochs_min <= epochs_max, \
            "{}: epochs_min must be <= epochs_max, passed are epochs_min {}, epochs_max {}".format(name, epochs_min,
                                                                                                 epochs_max)
        if evaluator_object is None:
            return
        validating = validation_every_n is not None and validation_metric is not None
        consistent = validating and (not stop_on_validation or lower_validations_allowed is not None)
        assert consistent, "{}: Inconsistent parameters passed, please check the supported uses".format(name)

    def _validate(self, evaluator_object, validation_metric):
        """One validation: metric value at the evaluator's first cutoff."""
        self._prepare_model_for_validation()
        per_cutoff, _ = evaluator_object.evaluateRecommender(self)
        first_cutoff = next(iter(per_cutoff))
        return per_cutoff[first_cutoff][validation_metric]

    def _train_with_early_stopping(self, epochs_max, epochs_min=0, validation_every_n=None, stop_on_validation=False,
                                   validation_metric=None, lower_validations_allowed=None, evaluator_object=None,
                                   algorithm_name="Incremental_Training_Early_Stopping"):
        self._check_arguments(algorithm_name, epochs_max, epochs_min, validation_every_n, stop_on_validation,
                              validation_metric, lower_validations_allowed, evaluator_object)
        self.best_validation_metric = None
        self.epochs_best = 0
        misses = 0                                   # validations in a row without a new best
        done = 0                                     # epochs run so far
        while done < epochs_max:
            self._run_epoch(done)
            done += 1
            if evaluator_object is None:
                self.epochs_best = done - 1          # (the reference records the 0-based index here)
                continue
            if done % validation_every_n:
                continue
            value = self._validate(evaluator_object, validation_metric)
            if self.best_validation_metric is None or self.best_validation_metric < value:
                self.best_validation_metric, self.epochs_best, misses = value, done, 0
                self._update_best_model()
            else:
                misses += 1
            if stop_on_validation and misses >= lower_validations_allowed and done - 1 >= epochs_min:
                break
        if evaluator_object is None:                 # no validation: the final model is the one to keep
            self._prepare_model_for_validation()
            self._update_best_model()
        return done
