"""EvaluatorHoldout of the reference (Base/Evaluation/Evaluator.py:219-414) with the hot loop on the GPU.

Same constructor, same `evaluateRecommender(recommender) -> (results_dict, results_string)`, same
metric keys, averaging and F1 rule.  What changed underneath: instead of the per-user Python loop
over 20 metric updates (Evaluator.py:291-335, ~85 % of the reference's evaluation time) each user
block is scored, seen-masked, top-k'ed and reduced to metric sums on the device; only the sums and the
per-item recommendation histograms come back."""
import numpy as np
import scipy.sparse as sps

from ... import _lib as L
from .metrics import finalize_count_metrics

# key order of the reference's EvaluatorMetrics enum (Evaluator.py:21-43) minus DIVERSITY_SIMILARITY,
# which only exists when a diversity_object is passed
RESULT_KEYS = ["ROC_AUC", "PRECISION", "PRECISION_RECALL_MIN_DEN", "RECALL", "MAP", "MRR", "NDCG", "F1",
               "HIT_RATE", "ARHR", "RMSE", "NOVELTY", "AVERAGE_POPULARITY", "DIVERSITY_MEAN_INTER_LIST",
               "DIVERSITY_HERFINDAHL", "COVERAGE_ITEM", "COVERAGE_USER", "DIVERSITY_GINI", "SHANNON_ENTROPY"]


def get_result_string(results_run, n_decimals=7):                      # Evaluator.py:95-110
    output_str = ""
    for cutoff in results_run.keys():
        output_str += "CUTOFF: {} - ".format(cutoff)
        for metric in results_run[cutoff].keys():
            output_str += "{}: {:.{n_decimals}f}, ".format(metric, results_run[cutoff][metric], n_decimals=n_decimals)
        output_str += "\n"
    return output_str


class Evaluator(object):
    EVALUATOR_NAME = "Evaluator_Base_Class"

    def __init__(self, URM_test_list, cutoff_list, minRatingsPerUser=1, exclude_seen=True, diversity_object=None,
                 ignore_items=None, ignore_users=None):
        super(Evaluator, self).__init__()
        if ignore_items is None:                                       # Evaluator.py:128-134
            self.ignore_items_flag = False
            self.ignore_items_ID = np.array([])
        else:
            print("Ignoring {} Items".format(len(ignore_items)))
            self.ignore_items_flag = True
            self.ignore_items_ID = np.array(ignore_items)
        self.cutoff_list = cutoff_list.copy()
        self.max_cutoff = max(self.cutoff_list)
        self.minRatingsPerUser = minRatingsPerUser
        self.exclude_seen = exclude_seen
        if isinstance(URM_test_list, list):
            raise ValueError("List of URM_test not supported")
        if diversity_object is not None:
            raise NotImplementedError("DIVERSITY_SIMILARITY (diversity_object) is outside the GANMF hot path")
        self.diversity_object = None
        self.URM_test = sps.csr_matrix(URM_test_list.copy(), dtype=np.float32)
        self.URM_test.sort_indices()
        self.n_users, self.n_items = self.URM_test.shape
        numRatings = np.ediff1d(self.URM_test.indptr)
        self.usersToEvaluate = np.arange(self.n_users)[numRatings >= minRatingsPerUser]
        if ignore_users is not None:
            print("Ignoring {} Users".format(len(ignore_users)))
            self.ignore_users_ID = np.array(ignore_users)
            self.usersToEvaluate = sorted(set(self.usersToEvaluate) - set(ignore_users))
        else:
            self.ignore_users_ID = np.array([])
        self.usersToEvaluate = list(self.usersToEvaluate)

    def get_user_relevant_items(self, user_id):
        return self.URM_test.indices[self.URM_test.indptr[user_id]:self.URM_test.indptr[user_id + 1]]

    def get_user_test_ratings(self, user_id):
        return self.URM_test.data[self.URM_test.indptr[user_id]:self.URM_test.indptr[user_id + 1]]


class EvaluatorHoldout(Evaluator):
    EVALUATOR_NAME = "EvaluatorHoldout"

    def __init__(self, URM_test_list, cutoff_list, minRatingsPerUser=1, exclude_seen=True, diversity_object=None,
                 ignore_items=None, ignore_users=None):
        super(EvaluatorHoldout, self).__init__(URM_test_list, cutoff_list, diversity_object=diversity_object,
                                               minRatingsPerUser=minRatingsPerUser, exclude_seen=exclude_seen,
                                               ignore_items=ignore_items, ignore_users=ignore_users)
        self._scores_engine = None      # lazily created device context for recommenders without their own

    # ------------------------------------------------------------------------------------------
    def _device_sums(self, recommender_object, users):
        """(sums[n_cut, MC_NCOL], counts[n_cut, n_items]) over `users` in ascending order."""
        eng = (recommender_object._device_engine() if hasattr(recommender_object, "_device_engine")
               else getattr(recommender_object, "_engine", None))
        # popularity tables need the users x items matrix: an engine-backed item-mode model may hold URM_train
        # transposed (after loadModel the reference leaves it so, GANMF.py:32-33,337-342)
        URM_train = getattr(recommender_object, "_URM_users_items", None) if eng is not None else None
        if URM_train is None:
            URM_train = recommender_object.get_URM_train()
        block = min(1000, max(1, int(1e8 / self.n_items)))             # Evaluator.py:238
        score_fn = lambda u: recommender_object._compute_item_score(u)
        if self.ignore_items_flag:
            # Evaluator.py:369-370 / BaseRecommender.py:103-106,210-211: the custom items get -inf before the
            # ranking.  Rare API-edge path: the score rows make a round trip through the host, where the columns
            # are masked, and re-enter the device mask -> top-k -> metric stage block by block.
            ignore = np.asarray(self.ignore_items_ID, dtype=np.int64)
            base_fn = score_fn

            def score_fn(u):
                sc = np.array(base_fn(u), dtype=np.float32, copy=True)
                sc[:, ignore] = -np.inf
                return sc
        host_edit = getattr(recommender_object, "_scores_need_host_edit", None)
        if eng is not None and host_edit is not None and host_edit(users):
            # a factor model whose score rows are edited on the host (biases, cold users): stream them through the
            # device mask -> top-k -> metric stage like any foreign recommender's
            eng.set_test(self.URM_test, URM_train)
            return eng.evaluate_scores(score_fn, users, self.cutoff_list, remove_seen=self.exclude_seen,
                                       block_size=block)
        if eng is None:
            # any recommender exposing _compute_item_score (e.g. the reference's own baselines): its host score
            # rows are pushed, block by block as Evaluator.py:238 sizes them, through the device
            # mask -> top-k -> metric stage.  There is no CPU evaluation path.
            if self._scores_engine is None:
                from ...engine import Engine
                self._scores_engine = Engine(L.KIND_GANMF, self.n_users, self.n_items, 1, emb_dim=1, max_batch=1)
            eng = self._scores_engine
            eng.set_csr(L.CSR_SEEN, URM_train, with_data=False)
            eng.set_test(self.URM_test, URM_train)
            return eng.evaluate_scores(score_fn, users, self.cutoff_list, remove_seen=self.exclude_seen,
                                       block_size=block)
        eng.set_test(self.URM_test, URM_train)
        if self.ignore_items_flag:
            return eng.evaluate_scores(score_fn, users, self.cutoff_list, remove_seen=self.exclude_seen,
                                       block_size=block)
        try:
            import torch.distributed as dist
            sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except ImportError:
            sharded = False
        if sharded:
            # under torchrun every rank holds the model (replicated factors): each evaluates a contiguous range of
            # the ascending user list, only metric sums and item histograms cross GPUs, and
            # the running sums are continued rank after rank -- same bits as the single-GPU evaluation
            from ...parallel import shard_rows, sharded_eval_sums
            lo, hi = shard_rows(len(users), dist.get_world_size(), dist.get_rank())
            sums, counts, _ = sharded_eval_sums(eng, users[lo:hi], self.cutoff_list, self.exclude_seen, dist=dist)
            return sums, counts
        return eng.evaluate(users, self.cutoff_list, remove_seen=self.exclude_seen)

    def evaluateRecommender(self, recommender_object):
        users = np.asarray(self.usersToEvaluate, dtype=np.int32)
        n_eval = len(users)
        results_dict = {}
        if n_eval > 0:
            if self.ignore_items_flag and hasattr(recommender_object, "set_items_to_ignore"):
                recommender_object.set_items_to_ignore(self.ignore_items_ID)          # Evaluator.py:369-370
            sums, counts = self._device_sums(recommender_object, users)
            if self.ignore_items_flag and hasattr(recommender_object, "reset_items_to_ignore"):
                recommender_object.reset_items_to_ignore()                            # Evaluator.py:410-411
            for ci, cutoff in enumerate(self.cutoff_list):
                s = dict(zip(L.MC_NAMES, sums[ci]))
                res = {k: s[k] / n_eval for k in ("ROC_AUC", "PRECISION", "PRECISION_RECALL_MIN_DEN", "RECALL", "MAP",
                                                  "MRR", "NDCG", "HIT_RATE", "ARHR", "RMSE", "NOVELTY",
                                                  "AVERAGE_POPULARITY")}
                res.update(finalize_count_metrics(counts[ci], n_eval, cutoff, self.n_items,
                                                  self.ignore_items_ID if self.ignore_items_flag else None))
                res["COVERAGE_USER"] = s["COVERED"] / (self.n_users - len(self.ignore_users_ID))   # metrics.py:57-80
                p_, r_ = res["PRECISION"], res["RECALL"]
                res["F1"] = 2 * (p_ * r_) / (p_ + r_) if p_ + r_ != 0 else 0.0   # Evaluator.py:392-397
                results_dict[cutoff] = {k: res[k] for k in RESULT_KEYS}
        else:
            print("WARNING: No users had a sufficient number of relevant items")
            for cutoff in self.cutoff_list:
                results_dict[cutoff] = {k: 0.0 for k in RESULT_KEYS}
        return results_dict, get_result_string(results_dict)
