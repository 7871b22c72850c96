"""Recommender base API of the reference (Base/BaseRecommender.py) on top of the device engine.

Same public surface: recommend(user_id_array, cutoff, remove_seen_flag, items_to_compute,
remove_top_pop_flag, remove_CustomItems_flag, return_scores), get_URM_train(),
set_items_to_ignore()/reset_items_to_ignore(), saveModel()/loadModel().  The body of recommend()
(BaseRecommender.py:155-247: score -> -inf on seen -> argpartition -> argsort -> drop -inf) runs as
score GEMM -> seen mask -> per-row top-k on the GPU; ties go to the lowest item index."""
import numpy as np

from .. import _lib as L


class BaseRecommender(object):
    RECOMMENDER_NAME = "Recommender_Base_Class"

    def __init__(self, URM_train=None):
        super(BaseRecommender, self).__init__()
        self.URM_train = URM_train
        self.items_to_ignore_flag = False
        self.items_to_ignore_ID = np.array([], dtype=int)
        self.filterTopPop = False
        self.filterTopPop_ItemsID = np.array([], dtype=int)

    # -- reference API -----------------------------------------------------------------------
    def fit(self):
        pass

    def get_URM_train(self):                                   # BaseRecommender.py:51-52
        return self.URM_train.copy()

    def set_items_to_ignore(self, items_to_ignore):            # :72-75
        self.items_to_ignore_flag = True
        self.items_to_ignore_ID = np.array(items_to_ignore, dtype=int)

    def reset_items_to_ignore(self):                           # :77-80
        self.items_to_ignore_flag = False
        self.items_to_ignore_ID = np.array([], dtype=int)

    def _compute_item_score(self, user_id_array, items_to_compute=None):
        raise NotImplementedError("BaseRecommender: compute_item_score not assigned for current recommender, "
                                  "unable to compute prediction scores")

    def _device_engine(self):
        """Engine that owns the model on the GPU (None for recommenders without one)."""
        return getattr(self, "_engine", None)

    def recommend(self, user_id_array, cutoff=None, remove_seen_flag=True, items_to_compute=None,
                  remove_top_pop_flag=False, remove_CustomItems_flag=False, return_scores=False):
        if np.isscalar(user_id_array):                         # :159-163
            user_id_array = np.atleast_1d(user_id_array)
            single_user = True
        else:
            single_user = False
        user_id_array = np.asarray(user_id_array)
        if cutoff is None:                                     # :166-167
            cutoff = self.URM_train.shape[1] - 1
        eng = self._device_engine()
        if eng is None:
            raise RuntimeError("%s has no device engine (call fit() or loadModel() first); ganmf_b200 has no "
                               "CPU recommend path" % self.RECOMMENDER_NAME)
        extra_mask = []
        if remove_top_pop_flag:                                # :207-208
            extra_mask.append(np.asarray(self.filterTopPop_ItemsID, dtype=np.int64))
        if remove_CustomItems_flag:                            # :210-211
            extra_mask.append(np.asarray(self.items_to_ignore_ID, dtype=np.int64))
        host_edit = getattr(self, "_scores_need_host_edit", None)
        own_scores = items_to_compute is not None or (host_edit is not None and host_edit(user_id_array))
        if cutoff > L.TOPK_MAX or extra_mask or own_scores:
            # rare API-edge cases (full rankings, custom item filters, restricted item sets, biases / cold users of a
            # matrix-factorisation baseline): the score rows are edited on the host as the reference does and the
            # device mask -> top-k kernel runs on the edited matrix in chunks of 128 ranks
            scores = (np.ascontiguousarray(self._compute_item_score(user_id_array, items_to_compute=items_to_compute),
                                           dtype=np.float32) if own_scores else eng.score(user_id_array))
            if extra_mask:
                scores[:, np.concatenate(extra_mask)] = -np.inf
            idx = self._topk_large(eng, scores, user_id_array, int(cutoff), remove_seen_flag)
            scores_batch = scores
        else:
            idx, _, scores_batch = eng.recommend(user_id_array, int(cutoff), remove_seen=remove_seen_flag,
                                                 return_scores=return_scores)
        ranking_list = [row[row >= 0].tolist() for row in idx]  # -inf entries dropped (:227-234)
        if single_user:
            ranking_list = ranking_list[0]
        if return_scores:
            return ranking_list, scores_batch
        return ranking_list

    @staticmethod
    def _topk_large(eng, scores, users, cutoff, remove_seen):
        """cutoff > 128: peel the ranking off 128 ranks at a time with the device mask+top-k kernel
        (each pass masks what the previous ones returned).  `scores` ends up seen-masked only."""
        n, n_items = scores.shape
        cutoff = min(cutoff, n_items)
        work = np.ascontiguousarray(scores, dtype=np.float32).copy()
        out = np.full((n, cutoff), -1, dtype=np.int32)
        done = 0
        first = True
        while done < cutoff:
            k = min(L.TOPK_MAX, cutoff - done)
            idx, _ = eng.mask_topk(work, k, users=users, remove_seen=remove_seen and first, write_back=True)
            if first and remove_seen:
                scores[...] = work
            first = False
            out[:, done:done + k] = idx
            rows = np.repeat(np.arange(n), k)
            cols = idx.reshape(-1)
            ok = cols >= 0
            work[rows[ok], cols[ok]] = -np.inf
            done += k
            if not ok.any():
                break
        return out

    def saveModel(self, folder_path, file_name=None):
        raise NotImplementedError("BaseRecommender: saveModel not implemented")

    def loadModel(self, folder_path, file_name=None):
        raise NotImplementedError("BaseRecommender: loadModel not implemented")
