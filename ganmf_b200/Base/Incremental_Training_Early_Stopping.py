"""PoliMi early-stopping mixin (reference: Base/Incremental_Training_Early_Stopping.py:93-259).

Same contract: the recommender implements _run_epoch / _prepare_model_for_validation /
_update_best_model and calls _train_with_early_stopping(...).  0-based epochs, validation when
(epoch+1) % validation_every_n == 0 on the FIRST cutoff of the evaluator, strict improvement
`best < current`, stop when lower_validations_count >= lower_validations_allowed and
epoch >= epochs_min; self.epochs_best is what get_early_stopping_final_epochs_dict() returns."""


class Incremental_Training_Early_Stopping(object):
    def __init__(self):
        super(Incremental_Training_Early_Stopping, self).__init__()

    def get_early_stopping_final_epochs_dict(self):
        return {"epochs": self.epochs_best}

    def _run_epoch(self, num_epoch):
        raise NotImplementedError()

    def _prepare_model_for_validation(self):
        raise NotImplementedError()

    def _update_best_model(self):
        raise NotImplementedError()

    def _train_with_early_stopping(self, epochs_max, epochs_min=0, validation_every_n=None, stop_on_validation=False,
                                   validation_metric=None, lower_validations_allowed=None, evaluator_object=None,
                                   algorithm_name="Incremental_Training_Early_Stopping"):
        assert epochs_max > 0, "{}: Number of epochs_max must be > 0, passed was {}".format(algorithm_name, epochs_max)
        assert epochs_min >= 0, "{}: Number of epochs_min must be >= 0, passed was {}".format(algorithm_name,
                                                                                              epochs_min)
        assert epochs_min <= epochs_max, "{}: epochs_min must be <= epochs_max, passed are epochs_min {}, " \
                                         "epochs_max {}".format(algorithm_name, epochs_min, epochs_max)
        assert evaluator_object is None or \
            (not stop_on_validation and validation_every_n is not None and validation_metric is not None) or \
            (stop_on_validation and validation_every_n is not None and validation_metric is not None and
             lower_validations_allowed is not None), \
            "{}: Inconsistent parameters passed, please check the supported uses".format(algorithm_name)

        self.best_validation_metric = None
        lower_validations_count = 0
        convergence = False
        self.epochs_best = 0
        epochs_current = 0
        while epochs_current < epochs_max and not convergence:
            self._run_epoch(epochs_current)
            if evaluator_object is None:
                self.epochs_best = epochs_current
            elif (epochs_current + 1) % validation_every_n == 0:
                self._prepare_model_for_validation()
                results_run, _ = evaluator_object.evaluateRecommender(self)
                results_run = results_run[list(results_run.keys())[0]]
                current_metric_value = results_run[validation_metric]
                if self.best_validation_metric is None or self.best_validation_metric < current_metric_value:
                    self.best_validation_metric = current_metric_value
                    self._update_best_model()
                    self.epochs_best = epochs_current + 1
                    lower_validations_count = 0
                else:
                    lower_validations_count += 1
                if stop_on_validation and lower_validations_count >= lower_validations_allowed and \
                        epochs_current >= epochs_min:
                    convergence = True
            epochs_current += 1
        if evaluator_object is None:
            self._prepare_model_for_validation()
            self._update_best_model()
        return epochs_current
