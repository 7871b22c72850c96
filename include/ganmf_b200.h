/* ganmf_b200 -- C ABI of the B200-native GANMF / DisGANMF hot path.
 *
 * Drop-in boundary.  The reference (edervishaj/GANMF) is Python over TensorFlow 1.12; the
 * "FFI" it crosses for this path is tf.Session.run() on the graph built in
 *   GANRec/GANMF.py:53-139 / GANRec/DisGANMF.py:51-140           (build + losses + minimize)
 * plus numpy inside Base/BaseRecommender.py:155-247 (recommend) and the Python loop of
 *   Base/Evaluation/Evaluator.py:234-414 (EvaluatorHoldout).
 * Each entry point below names the reference call it replaces.  Plain pointers and sizes only;
 * all *_host pointers are host memory, everything else lives in the context on the GPU.
 * Every function returns 0 on success, non-zero on error (ganmf_last_error() has the text).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Storage convention: matrices are row-major fp32 [rows][cols]; inside the context they are
 * padded to a leading dimension of roundup(cols, 32) floats, at the ABI they are dense.
 */
#ifndef GANMF_B200_H
#define GANMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ganmf_ctx ganmf_ctx;

enum { GANMF_KIND_GANMF = 0, GANMF_KIND_DISGANMF = 1,
       /* factor matrices only (no discriminator, no training): the device scorer / evaluator behind
        * Base/BaseMatrixFactorizationRecommender.py:94-143 (USER_factors . ITEM_factors^T), and the evaluation
        * context of an item-sharded training run */
       GANMF_KIND_MF = 2 };
enum { GANMF_ACT_LINEAR = 0, GANMF_ACT_TANH = 1, GANMF_ACT_RELU = 2, GANMF_ACT_SIGMOID = 3 };
enum { GANMF_CSR_TRAIN = 0,   /* rows of the TRAINING orientation (items x users in --item mode) */
       GANMF_CSR_SEEN = 1,    /* users x items, masks seen items in recommend()                   */
       GANMF_CSR_TEST = 2 };  /* users x items held-out interactions (evaluator)                  */
enum { GANMF_GEMM_AUTO = 0, GANMF_GEMM_SIMT = 1, GANMF_GEMM_TC = 2,
       /* tcgen05 on split-TF32 operands (hi.hi + hi.lo + lo.hi): fp32-accurate products at 3x the MMA work */
       GANMF_GEMM_TC3 = 3,
       /* ganmf_k_gemm only: the generator kernel with a resident A tile (both operands K-major, K <= 256) */
       GANMF_GEMM_RESIDENT_A = 4 };

/* Per-(user, cutoff) metric columns produced by the device evaluator. */
enum { GANMF_MC_PRECISION = 0, GANMF_MC_RECALL, GANMF_MC_PRMD, GANMF_MC_MAP, GANMF_MC_NDCG,
       GANMF_MC_MRR, GANMF_MC_ARHR, GANMF_MC_ROC_AUC, GANMF_MC_HIT, GANMF_MC_NOVELTY,
       GANMF_MC_AVGPOP, GANMF_MC_COVERED, GANMF_MC_RMSE, GANMF_MC_NCOL };

typedef struct ganmf_config {
  int kind;            /* GANMF_KIND_*                                                         */
  int n_rows;          /* rows of the training matrix held by THIS context (user shard)         */
  int width;           /* profile width (columns of the training matrix)                        */
  int num_factors;     /* k            (GANMF.py:53 build(num_factors, emb_dim))                */
  int emb_dim;         /* GANMF: autoencoder code size E                                        */
  int d_layers;        /* DisGANMF: hidden layers          (DisGANMF.py:51)                     */
  int d_nodes;         /* DisGANMF: units per hidden layer                                      */
  int d_act;           /* DisGANMF: GANMF_ACT_*                                                 */
  int max_batch;       /* largest minibatch (rows per step on this GPU)                         */
  int item_mode;       /* 1: rows are items; scoring uses V[u] . P^T (GANMF.py:288-290)         */
  int row_id_offset;   /* global id of local row 0 (data-parallel shards; DisGANMF id feature)  */
  int device;          /* CUDA device ordinal                                                   */
  int gemm_path;       /* GANMF_GEMM_*; AUTO picks tcgen05 unless the shape is tiny             */
  /* Item-sharded (tensor-parallel) training, GANMF only (no reference counterpart,
   * the reference trains on one device): this context holds columns [item_offset, item_offset + width) of the
   * training matrix and the matching slices of We (rows), Wd / bd (columns) and V (rows); the user factors and
   * the encoder bias are replicated.  Mean losses and glorot limits use global_width.  All zero = unsharded. */
  int global_width;    /* columns of the WHOLE training matrix (0: = width)                     */
  int item_offset;     /* first global column held by this context                              */
  int tp_rank;         /* rank inside the item-sharded group                                    */
  int tp_world;        /* size of the group (0 or 1: unsharded)                                 */
} ganmf_config;

/* ---- lifecycle -------------------------------------------------------------------------- */
const char* ganmf_last_error(void);
int ganmf_create(const ganmf_config* cfg, ganmf_ctx** out);       /* ~ GANMF.build + Session   */
void ganmf_destroy(ganmf_ctx* ctx);
int ganmf_set_stream(ganmf_ctx* ctx, void* cuda_stream);          /* cudaStream_t, NULL = default */
/* Cap on the number of SMs the persistent tensor-core GEMMs occupy (0 = all).  A data-parallel caller lowers
 * it while a collective is in flight: a GEMM CTA needs a whole SM, so NCCL's CTAs can only run next to a GEMM
 * on SMs the GEMM grid leaves free.  (No reference counterpart: the reference trains on one device,
 * GANMF.py:142-150; this belongs to the data-parallel extension) */
int ganmf_set_gemm_sms(ganmf_ctx* ctx, int n_sms);
int ganmf_synchronize(ganmf_ctx* ctx);

/* ---- data ------------------------------------------------------------------------------- */
/* ~ URM_train[uids].toarray() source (GANMF.py:184) / seen filter (BaseRecommender.py:93-100) /
 *   URM_test (Evaluator.py:196-207).  data_host may be NULL (all ones).  Indices of a row must be
 *   sorted for GANMF_CSR_TEST. */
int ganmf_set_csr(ganmf_ctx* ctx, int which, int n_rows, int n_cols, const int32_t* indptr_host,
                  const int32_t* indices_host, const float* data_host);
/* The same from DEVICE arrays (copied device-to-device; nnz = indptr[n_rows]).  ~ sps.load_npz + upload of
 * experiments/datasets/*.npz without a host detour when the matrix is built or transposed on the GPU
 * (GANMF.py:32-33 transposes on the host). */
/* CSR `which` := the TRANSPOSE of the host CSR [n_rows x n_cols] given here, transposed on the device (count, scan,
 * scatter, per-column sort: the result is canonical, bit-identical to scipy's).  ~ `URM_train.T.tocsr()` of item mode
 * (GANRec/GANMF.py:32-33): the users x items matrix is uploaded as it is and the items x users training matrix is built on
 * the GPU.  For GANMF_CSR_TRAIN the context expects n_cols x n_rows = config.n_rows x config.width. */
int ganmf_set_csr_transposed(ganmf_ctx* ctx, int which, int n_rows, int n_cols, const int32_t* indptr_host,
                             const int32_t* indices_host, const float* data_host);
/* Read a resident CSR back (sizes first with NULL arrays, then the arrays): the matrix `get_URM_train()` returns
 * (Base/BaseRecommender.py:51-52) as the device holds it, e.g. after ganmf_set_csr_transposed.  data_host must be NULL
 * for a CSR that was set without values. */
int ganmf_get_csr(ganmf_ctx* ctx, int which, int32_t* n_rows, int32_t* n_cols, int64_t* nnz, int32_t* indptr_host,
                  int32_t* indices_host, float* data_host);
int ganmf_set_csr_device(ganmf_ctx* ctx, int which, int n_rows, int n_cols, const int32_t* indptr_dev,
                         const int32_t* indices_dev, const float* data_dev, int64_t nnz);

/* ---- parameters (TF variable names, GANMF.py:119-121 / DisGANMF.py:58-64) ---------------- */
int ganmf_param_count(ganmf_ctx* ctx);
int ganmf_param_info(ganmf_ctx* ctx, int i, char* name_out, int name_cap, int* rows, int* cols,
                     int* is_generator);
int ganmf_set_param(ganmf_ctx* ctx, const char* name, const float* host, int64_t count);
int ganmf_get_param(ganmf_ctx* ctx, const char* name, float* host, int64_t count);
int ganmf_init_params(ganmf_ctx* ctx, uint64_t seed);   /* glorot_uniform + zero biases (GANMF.py:57) */
int ganmf_reset_optimizers(ganmf_ctx* ctx);             /* zero Adam moments, t = 0 (new fit())      */
int ganmf_snapshot(ganmf_ctx* ctx);                     /* save_current_model (GANMF.py:249-251)     */
int ganmf_restore(ganmf_ctx* ctx);                      /* load_model         (GANMF.py:253-255)     */

/* ---- training --------------------------------------------------------------------------- */
/* One sess.run([dtrain, dloss]) / sess.run([gtrain, gloss]) (GANMF.py:186-187,200-201;
 * DisGANMF.py:187-188,201).  row_ids are LOCAL row indices already resident on the device
 * (ganmf_upload_ids) at offset ids_offset.  The loss goes to the device-side loss log at
 * loss_slot; nothing is copied back, nothing synchronises.  n_rows_global is the row count of the
 * whole (all-GPU) minibatch: it normalises the mean losses under data parallelism.
 * m_hinge is ignored by DisGANMF. */
int ganmf_upload_ids(ganmf_ctx* ctx, const int32_t* ids_host, int n);
int ganmf_d_step(ganmf_ctx* ctx, int ids_offset, int B, int n_rows_global, float lr, float reg,
                 float m_hinge, int loss_slot);
int ganmf_g_step(ganmf_ctx* ctx, int ids_offset, int B, int n_rows_global, float lr, float reg,
                 float recon_coefficient, int loss_slot);
/* The same steps cut at the points where data-parallel ranks must exchange sums:
 *   d_forward -> allreduce(step scalars) -> d_backward -> allreduce(d grads) -> d_apply
 *   g_forward_backward -> allreduce(g shared grad + step scalars) -> g_apply */
int ganmf_d_forward(ganmf_ctx* ctx, int ids_offset, int B);
int ganmf_d_backward(ganmf_ctx* ctx, int B, int n_rows_global, float m_hinge);
/* GANMF: the D backward in two halves so the sum of the decoder gradients ("d_grads_dec") can overlap the
 * computation of the encoder half ("d_grads_enc"): phase 1 = gate + dWd,dbd; phase 2 = dH + dWe,dbe; 0 = both */
int ganmf_d_backward_phase(ganmf_ctx* ctx, int B, int n_rows_global, float m_hinge, int phase);
/* ... phases 3 / 4 split phase 2 into dH (still reads the decoder weights) and dWe, so the decoder can be
 * updated and all-gathered while dWe is computed.  Forward: phase 1 = profiles + generator (independent of
 * the discriminator weights), phase 2 = discriminator forward, 0 = both. */
int ganmf_d_forward_phase(ganmf_ctx* ctx, int ids_offset, int B, int phase);
int ganmf_d_apply(ganmf_ctx* ctx, float lr, float reg, int loss_slot);
int ganmf_g_forward_backward(ganmf_ctx* ctx, int ids_offset, int B, int n_rows_global,
                             float recon_coefficient);
/* GANMF: the same in two parts so the all-reduce of the item-factor gradient ("g_shared_grad", complete after
 * part 1) overlaps the user-factor gradient GEMM (part 2).  Both parts together are one
 * sess.run([gtrain, gloss]) up to the optimiser (GANMF.py:200-201); data-parallel extension. */
int ganmf_g_forward_backward_part(ganmf_ctx* ctx, int ids_offset, int B, int n_rows_global,
                                  float recon_coefficient, int part);
int ganmf_g_apply(ganmf_ctx* ctx, int B, int n_rows_global, float lr, float reg,
                  float recon_coefficient, int loss_slot);
/* Data-parallel D step with sharded optimiser work: after a reduce-scatter of the gradient slab, run
 * Adam on [offsets[i], offsets[i]+counts[i]) (elements of the discriminator slab, multiples of 4) only;
 * the caller all-gathers "d_params" afterwards, sums step_scalars[6] and calls ganmf_finalize_loss. */
int ganmf_d_apply_ranges(ganmf_ctx* ctx, float lr, float reg, const int64_t* offsets, const int64_t* counts,
                         int n_ranges, int new_step /* 1: first call of this optimiser step */);
/* Data-parallel G step only (n_rows_global != B): after ganmf_g_apply, sum step_scalars[6] (the l2 of
 * the row-sharded user factors) over ranks, then write loss_slot. */
int ganmf_finalize_loss(ganmf_ctx* ctx, float reg, int loss_slot);
/* Item-sharded training (config.tp_world > 1): the same sess.run([dtrain, dloss]) / sess.run([gtrain, gloss])
 * (GANMF.py:186-187,200-201) on the WHOLE minibatch of B rows restricted to this context's item slice, cut
 * where partial sums over the item slices must be added up by the caller (NCCL all-reduce, SUM):
 *   D: phase 1 (profiles, F = Pb.V^T, partial codes)      -> sum "tp_h2"[0 : 2B*ld]
 *      phase 2 (residuals, energy sums)                    -> sum "step_scalars"[0:2]   (the hinge gate is global)
 *      phase 3 (hinge gate, dbd, partial dH and dbe)       -> sum "tp_dh2"[0 : (2B+1)*ld]  (row 2B = dbe; may overlap phase 4)
 *      phase 4 (dWd with Adam in the epilogue: needs no summed quantity)
 *      phase 5 (dWe with Adam in the epilogue, biases, loss log)
 *   G: phase 1 (as D)                                      -> sum "tp_h2"[0 : 2B*ld]
 *      phase 2 (fake residual, feature matching, partial dHf) -> sum "tp_dh2"[B*ld : 2B*ld]
 *      phase 3 (dF, partial dPb)                           -> sum "tp_dpb"[0 : B*ld]  (may overlap phase 4)
 *      phase 4 (dV)
 *      phase 5 (Adam on the batch rows of P and on the V slice, loss log)
 * Low-rank generator route (ganmf_step_routes reports lowrank_fake = 1): after phase 1 only the real rows of the codes
 * and the [k, E] partial of V^T.We are summed -- "tp_h2"[0 : B*ld] and "tp_m1"[0 : k*ld] -- and phase 2 forms the fake
 * rows' codes from the summed M1.  Finer cuts of the same step, for overlapping the sums with work that does not need
 * them (both steps: 6 then 7 instead of 1; G step: 8, 9, 10 instead of 3, 4):
 *      phase 6  (profiles, real rows' partial codes)          -> sum "tp_h2"[0 : B*ld]   (may overlap phase 7)
 *      phase 7  (F = Pb.V^T, partial V^T.We)                  -> sum "tp_m1"[0 : k*ld]
 *      phase 8  (after G phase 2; -c1 Res_f^T.Pb -> dV, -c1 Res_f.V -> partial dPb: overlaps the sum of "tp_dh2")
 *      phase 9  (rank 0 adds dHf.M1^T to its partial dPb)      -> sum "tp_dpb"[0 : B*ld]  (may overlap phase 10)
 *      phase 10 (dV += We.(dHf^T.Pb))
 * ld = leading dimension reported by ganmf_device_buffer_ld.  Every rank passes the same ids.  The loss log holds
 * per-rank PARTIAL losses (rank 0: the data term + its l2 share; others: their l2 / reconstruction share): the
 * step's loss is their SUM over ranks, formed by the caller once per epoch. */
int ganmf_tp_d_phase(ganmf_ctx* ctx, int phase, int ids_offset, int B, float lr, float reg, float m_hinge,
                     int loss_slot);
int ganmf_tp_g_phase(ganmf_ctx* ctx, int phase, int ids_offset, int B, float lr, float reg,
                     float recon_coefficient, int loss_slot);
/* One epoch of the reference schedule (GANMF.py:172-203): shuffled row ids in, d_steps full D
 * passes then g_steps full G passes over the same batches, per-batch losses out (host).
 * H2D: n_rows ids; D2H: the losses.  Synchronises once at the end. */
int ganmf_train_epoch(ganmf_ctx* ctx, const int32_t* perm_host, int n_ids, int batch_size,
                      int d_steps, int g_steps, float d_lr, float g_lr, float d_reg, float g_reg,
                      float m_hinge, float recon_coefficient, float* d_losses_host,
                      float* g_losses_host);
int ganmf_read_losses(ganmf_ctx* ctx, float* host, int n);        /* loss log [0, n) -> host    */
/* Raw device buffers for the collectives (wrap with __cuda_array_interface__; fp32 unless noted):
 * "d_grads" (all discriminator gradients, contiguous; GANMF halves: "d_grads_enc" = dWe|dbe,
 * "d_grads_dec" = dWd|dbd), "d_params" (+ "_enc"/"_dec": the matching parameter ranges),
 * "g_shared_grad" (item-factor gradient),
 * "step_scalars" (7 float64: sumsq_real, sumsq_fake, feature-matching, l2 of replicated tensors,
 * bce_real, bce_fake, l2 of the row-sharded user factors).
 * Item-sharded training: "tp_h2" ([2*max_batch][ld] codes), "tp_dh2" ([2*max_batch+1][ld] code gradients, the
 * last used row carries the partial encoder-bias gradient), "tp_dpb" ([max_batch][ld] user-factor gradients).
 * "user_factors" / "item_factors": the factor matrices [rows][ld] (deferred optimiser steps are applied first). */
int ganmf_device_buffer(ganmf_ctx* ctx, const char* name, void** dev_ptr, int64_t* n_elems);
int ganmf_device_buffer_ld(ganmf_ctx* ctx, const char* name, int* ld);

/* ---- scoring / recommendation / evaluation ---------------------------------------------- */
/* ~ _compute_item_score (GANMF.py:285-292): scores_host[n][n_items], user ids in scoring
 * orientation.  H2D: ids; D2H: n * n_items floats. */
int ganmf_score(ganmf_ctx* ctx, const int32_t* user_ids_host, int n, float* scores_host);
/* ~ autoencoder_codes (GANMF.py:304-307): codes_host[n][emb_dim] = R[row_ids] . We + be for training rows
 * (GANMF only). */
int ganmf_encode(ganmf_ctx* ctx, const int32_t* row_ids_host, int n, float* codes_host);
/* ~ BaseRecommender.recommend (:155-247) on a given fp32 score matrix: optional seen mask from
 * GANMF_CSR_SEEN, then top-K (descending score, ties -> lowest index).  idx = -1 where the score is
 * -inf.  scores_host is updated in place with the mask when write_back != 0. */
int ganmf_mask_topk(ganmf_ctx* ctx, float* scores_host, int n, int n_items,
                    const int32_t* user_ids_host, int remove_seen, int K, int32_t* idx_host,
                    float* val_host, int write_back);
/* score -> mask -> top-K without leaving the device; only ids/values come back. */
int ganmf_recommend(ganmf_ctx* ctx, const int32_t* user_ids_host, int n, int remove_seen, int K,
                    int32_t* idx_host, float* val_host, float* masked_scores_host /* nullable */);
/* ~ EvaluatorHoldout._run_evaluation_on_selected_users (Evaluator.py:234-357): for the given users
 * (ascending, each with >= 1 test item) computes, per cutoff, float64 sums over users IN ORDER of
 * the GANMF_MC_* columns, and per-item recommendation counts.  Tables: see ganmf_set_eval_tables.
 * sums_host[n_cutoffs][GANMF_MC_NCOL]; item_counts_host[n_cutoffs][n_items] (nullable). */
int ganmf_set_eval_tables(ganmf_ctx* ctx, const float* test_gain_host, const float* test_gain_desc_host,
                          const float* logtab_host, int logtab_n, const double* item_novelty_host,
                          const uint8_t* item_has_pop_host, const double* item_popnorm_host,
                          int n_items_tables /* length of the three per-item tables; must equal the test matrix's columns */);
int ganmf_evaluate(ganmf_ctx* ctx, const int32_t* user_ids_host, int n_users, const int32_t* cutoffs_host,
                   int n_cutoffs, int remove_seen, int block_size, double* sums_host,
                   int64_t* item_counts_host);
/* The same in two calls, for evaluation sharded by user rows over several GPUs: every rank
 * computes the per-user metric values of ITS contiguous range of the ascending user list (parallel), then the
 * running sums are formed rank after rank -- rank r starts from the sums of ranks < r (carry_in_host,
 * [n_cutoffs][GANMF_MC_NCOL]; NULL = zeros) -- so the final sums equal the single-GPU running sums over all users
 * bit for bit (Evaluator.py:305-335 keeps ONE running sum per metric).  Item counts are plain integer sums. */
int ganmf_evaluate_values(ganmf_ctx* ctx, const int32_t* user_ids_host, int n_users, const int32_t* cutoffs_host,
                          int n_cutoffs, int remove_seen, int block_size);
int ganmf_evaluate_sums(ganmf_ctx* ctx, const double* carry_in_host, double* sums_host, int64_t* item_counts_host);
/* ganmf_recommend (without score rows) and ganmf_evaluate rank with the fused scorer when the largest cutoff is
 * <= 24 and num_factors <= 256: one TF32 tensor-core pass keeps per-row candidate lists in its epilogue (scores
 * never reach HBM), the candidates are re-scored exactly -- fl32(sum_k fp64(p_k v_k)) -- and each list carries a
 * certificate that it equals the top K of the exact score row; rows without one fall back to exact score rows +
 * the materialised top-k.  Counters since ganmf_create: rows ranked by the fused scorer / rows that fell back. */
int ganmf_eval_stats(ganmf_ctx* ctx, int64_t* fused_rows, int64_t* fallback_rows);
/* The same evaluation for ANY recommender that can produce host score rows (the reference's
 * recommender.recommend(...) call inside Evaluator.py:271-277): begin, then feed blocks of users IN ORDER
 * with their fp32 score rows [n][n_items] (masked in place with -inf on seen items when remove_seen),
 * then end.  Sums / counts as in ganmf_evaluate. */
int ganmf_eval_begin(ganmf_ctx* ctx, int n_users_total, const int32_t* cutoffs_host, int n_cutoffs);
int ganmf_eval_scores_block(ganmf_ctx* ctx, float* scores_host, const int32_t* user_ids_host, int n,
                            int remove_seen, int write_back);
int ganmf_eval_end(ganmf_ctx* ctx, double* sums_host, int64_t* item_counts_host);
/* Same metric stage on caller-supplied top-K lists (bit-exact contract tests). */
int ganmf_metrics_from_topk(ganmf_ctx* ctx, const int32_t* topk_idx_host, int K,
                            const int32_t* user_ids_host, int n_users, const int32_t* cutoffs_host,
                            int n_cutoffs, double* per_user_host /* [n][n_cut][NCOL] nullable */,
                            double* sums_host, int64_t* item_counts_host);

/* Which routes the training step of this context takes (decided when the train CSR is set; GANRec/GANMF.py:62-70,
 * 184-187 is one dense graph): *sparse_real = 1 when the codes of the real rows are the CSR gather-sum instead of
 * the real half of the dense encode GEMM (density <= 0.55 %, or GANMF_SPARSE_REAL=1); *bias_grad_from_gemm = 1 when
 * the decoder-bias gradient is formed from the residual GEMM's per-32-row column sums (GANMF_COLPART != 0);
 * *lowrank_fake = 1 when the products that contract the generated profiles `fake_profile = P[u] . V^T`
 * (GANRec/GANMF.py:82-84) over the items go through the [k, emb_dim] matrix V^T . W_enc (2 * num_factors <=
 * max_batch, or GANMF_LOWRANK=1): an item-sharded caller then all-reduces the buffer "tp_m1" and only the real rows
 * of "tp_h2" after phase 1 of ganmf_tp_d_phase / ganmf_tp_g_phase. */
int ganmf_step_routes(ganmf_ctx* ctx, int32_t* sparse_real, int32_t* bias_grad_from_gemm, int32_t* lowrank_fake);

/* ---- primitive kernels (unit tests, ncu) ------------------------------------------------- */
/* All pointers are DEVICE pointers here. */
int ganmf_k_gemm(ganmf_ctx* ctx, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn,
                 int M, int N, int K, float* out, int ldo, int path);
int ganmf_k_csr_gather_dense(ganmf_ctx* ctx, int ids_offset, int B, float* out, int ld);
/* Sparse real-profile encode on its own: out[b, :] = be + sum over the interactions j of training row
 * ids[ids_offset + b] of data[j] * We[indices[j], :] -- `real_histories . W_enc + b_enc` of GANRec/GANMF.py:64 without
 * the dense tile of GANMF.py:184.  out: [B][ldo] device floats, ldo >= roundup(emb_dim, 32). */
int ganmf_k_csr_encode_rows(ganmf_ctx* ctx, int ids_offset, int B, float* out, int ldo);
int ganmf_k_adam(ganmf_ctx* ctx, float* theta, float* m, float* v, const float* g, int64_t n,
                 float alpha, float reg);
int ganmf_k_topk(ganmf_ctx* ctx, const float* scores, int ld, int n, int n_items, int K,
                 int32_t* idx, float* val);
/* Live CUDA-event timing of the tensor-core GEMM launches on the context's stream (bench.py's
 * roofline): enable, run steps, then read the summed device time, algorithmic FLOPs (2*M*N*K) and
 * launch count of those GEMMs since the last read. */
int ganmf_profile(ganmf_ctx* ctx, int enable);
int ganmf_profile_read(ganmf_ctx* ctx, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
/* Per-launch records of the same window (call BEFORE ganmf_profile_read): ms[i] and the GEMM shape
 * shape[4*i..] = {M, N, K, splits}; returns the number of records written (<= cap) in *n. */
int ganmf_profile_records(ganmf_ctx* ctx, double* ms, int32_t* shape, int cap, int* n);
/* number of kernels this library has launched since ganmf_create (bench.py's gpu_launches) */
int64_t ganmf_launch_count(ganmf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GANMF_B200_H */
