// C ABI of ganmf_b200 (include/ganmf_b200.h): device-resident model state + the step, scoring
// and evaluation drivers that sequence the sm_100a kernels.  No CPU compute path exists here.
#include "../../include/ganmf_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "csr_kernels.cuh"
#include "eval_kernels.cuh"
#include "gen_gemm.cuh"
#include "kernels.cuh"
#include "score_select.cuh"
#include "tc_gemm.cuh"

using namespace ganmf;

static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CU(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess)                                                                 \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));       \
  } while (0)
#define RC(x)                \
  do {                       \
    int r_ = (x);            \
    if (r_) return r_;       \
  } while (0)

static inline int rup(int x, int a) { return (x + a - 1) / a * a; }

struct Mat {
  float* p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  size_t elems() const { return (size_t)rows * ld; }
  float* row(int r) const { return p + (size_t)r * ld; }
};
struct Param {
  std::string name;
  Mat w;                       // view into the optimiser group's slab
  float *m = nullptr, *v = nullptr, *g = nullptr, *best = nullptr;
  int is_gen = 0;
  // item-sharded contexts: position of this slice inside the whole tensor (glorot init draws the SAME values a
  // single-GPU context would, so sharded and unsharded runs start from one model)
  int row_off = 0, col_off = 0, rows_g = 0, cols_g = 0;
};
struct Csr {
  int* indptr = nullptr;
  int* indices = nullptr;
  float* data = nullptr;
  int n_rows = 0, n_cols = 0;
  long long nnz = 0;
};

struct ganmf_ctx {
  ganmf_config cfg;
  cudaStream_t st = 0;
  std::vector<Param> params;           // discriminator tensors first, then P, V
  int n_d = 0;                         // number of discriminator tensors
  // slabs: [theta | m | v | grad | best] for D; P and V separate (P can be very large)
  float* d_slab = nullptr; size_t d_elems = 0;
  float* p_slab = nullptr; size_t p_elems = 0;     // theta, m, v, best (no dense grad)
  float* v_slab = nullptr; size_t v_elems = 0;     // theta, m, v, grad, best
  bool have_best = false;
  // activations / workspaces
  int B = 0, W = 0, Wp = 0, k = 0, kp = 0, E = 0, Ep = 0;
  int Wg = 0;                          // columns of the WHOLE training matrix (= W unless item-sharded)
  int tp_rank = 0, tp_world = 1;       // item-sharded group
  float tp_g_alpha = 0.f, tp_d_alpha = 0.f;
  Mat X2, H2, H2s, Res2, dH2, dF, Pb, dPb;
  // DisGANMF: per layer activations h[l] [2B, Hp], dz [2B, Hp], out2/dout2 [2B]
  std::vector<Mat> hs, dzs;
  Mat dhtmp;
  float *out2 = nullptr, *dout2 = nullptr, *idf = nullptr;
  float* ws = nullptr; size_t ws_floats = 0;
  bool presplit = false;               // the GEMM being issued already has split-TF32 operands (scoring)
  float* s3[2] = {nullptr, nullptr}; size_t s3_cap[2] = {0, 0};     // split-TF32 operand copies (GANMF_GEMM_TC3)
  StepScalars* sc = nullptr;
  float* losses = nullptr; int losses_cap = 0;
  int* ids = nullptr; int ids_cap = 0;
  int* slot = nullptr;
  Csr csr[3];
  float b1p[2] = {ADAM_B1, ADAM_B1}, b2p[2] = {ADAM_B2, ADAM_B2};   // [0]=D optimiser, [1]=G
  // evaluator
  EvalTables tb{};
  float *tb_gain = nullptr, *tb_gain_desc = nullptr, *tb_logtab = nullptr;
  double *tb_nov = nullptr, *tb_popn = nullptr;
  unsigned char* tb_haspop = nullptr;
  float* rmse_scratch = nullptr;
  bool have_tables = false;
  float* scores = nullptr; size_t scores_elems = 0;   // [block][items_ld]
  Mat Fb;                                              // gathered query rows, split-TF32 [hi|hi|lo]
  Mat Ob;                                              // ranked-item factors, split-TF32 [hi|lo|hi]
  int* topk_idx = nullptr; float* topk_val = nullptr; size_t topk_cap = 0;
  double* uvals = nullptr; size_t uvals_cap = 0;
  double* usums = nullptr; int* icounts = nullptr; size_t icounts_cap = 0;
  int* cut_dev = nullptr;
  int* eval_users = nullptr; int eval_users_cap = 0;
  // fused scorer (score_select.cuh): gathered query factors, candidate lists, fallback bookkeeping
  bool eval_fused = true;              // GANMF_EVAL_FUSED=0: always the materialised split-TF32 scorer (A/B, tests)
  // two buffer sets: while the tensor cores score block i+1, the re-scoring / metric kernels of block i run on
  // side streams (they fit next to the one persistent GEMM CTA per SM)
  struct FusedBuf {
    Mat Qg;                              // gathered query factors of the block
    float* cand_val = nullptr; int* cand_idx = nullptr; size_t cand_cap = 0;
    unsigned int* row_thr = nullptr;     // shared per-row rejection thresholds
    int* fb_rows = nullptr; size_t rows_cap = 0;
    int* fb_count = nullptr; int* h_count = nullptr;     // device counter + pinned host copy
    int* topk_idx = nullptr; float* topk_val = nullptr; size_t topk_cap = 0;
    cudaEvent_t ev_sel = nullptr, ev_res = nullptr, ev_fin = nullptr;
  } fbuf[2];
  cudaStream_t st_res = nullptr, st_fin = nullptr;
  cudaEvent_t ev_setup = nullptr, ev_pipe_done = nullptr;
  unsigned int* vmax_bits = nullptr;   // max_j ||ranked factor row j|| (float bits)
  int* fb_users = nullptr;
  int* fb_idx = nullptr; float* fb_val = nullptr; size_t fb_topk_cap = 0;
  long long fused_rows = 0, fallback_rows = 0;      // statistics since ganmf_create
  long long launches = 0;
  int ev_total = 0, ev_done = 0, ev_ncut = 0, ev_K = 0;     // streaming evaluation (ganmf_eval_begin..end)
  int ev_pending_users = -1, ev_pending_ncut = 0;           // ganmf_evaluate_values done, sums not yet formed
  bool ev_sums_done = false;                                // ... or formed block by block already (ganmf_evaluate)
  int last_ids_offset = 0;
  int last_n_global = 0;      // DisGANMF: rows of the global minibatch of the pending D update
  float last_alpha_d = 0.f;
  int gemm_sm_cap = 0;        // > 0: persistent GEMM grids leave SMs free (ganmf_set_gemm_sms)
  int pair_mode = 1;          // CTA-pair (cta_group::2) GEMM tiles: 0 = never, 1 = GEMMs with >= 148 tiles, 2 = whenever legal
  bool fuse_adam = true;      // ganmf_d_step / ganmf_g_step: optimiser inside the weight-gradient GEMMs
  // sparse real-profile encode: codes of the real rows as a gather-sum over their CSR entries instead
  // of the real half of the dense G2 product.  sparse_mode: -1 = by density (GANMF_SPARSE_REAL unset), 0 / 1 = forced
  int sparse_mode = -1;
  bool sparse_real = false;
  // Side stream: small-footprint kernels that are independent of the main chain run beside it -- the deferred
  // user-factor steps (compute-bound, 64-thread CTAs) next to the profile gather and the real rows' encode
  // (HBM-bound).
  // Fork = event on st, join = event on st_aux; every entry point leaves with the side stream joined or joins it at
  // the first consumer (aux_pending).  GANMF_AUX_STREAM=0: everything on st (A/B switch).
  cudaStream_t st_aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool aux_on = true, aux_pending = false;
  // generator GEMM F = Pb . V^T on the resident-A kernel (gen_gemm.cuh): -1 = when the minibatch fills a pair tile
  // (B >= 256) and k <= 256, 0 / 1 forced (GANMF_GEN_RESIDENT)
  int gen_resident_mode = -1;
  // low-rank generator route: the fake profiles F = Pb . V^T have rank k, so every product that contracts F (or a
  // gradient flowing back into it) over the items can go through a [k, E] matrix instead of a [B, I] one:
  //   codes      Hf  = F . We            = Pb . (V^T . We)                        = Pb . M1
  //   dPb        = dF . V                = dHf . M1^T - c1 * (Res_f . V)
  //   dV         = dF^T . Pb             = We . (dHf^T . Pb) - c1 * (Res_f^T . Pb)
  // (dF = dHf . We^T - c1 * Res_f is never formed).  2*k*I*E flops for M1 replace 2*B*I*E for the fake codes, and the
  // [B, I] x [I, E] product behind dF disappears: worth it when k is well below the minibatch rows (always for the
  // item-sharded step, whose minibatch is N * 1024 rows).  lowrank_mode: -1 = by shape (2k <= max_batch), 0 / 1 forced
  // (GANMF_LOWRANK).
  int lowrank_mode = -1;
  bool lowrank = false;
  // ... and the encoder weight gradient of an item-sharded rank: dWe = X2^T.dH2 = R^T.dH_r + V.(Pb^T.dH_f) halves the
  // K of that GEMM (K = 2B -> B + k).  Worth it where the GEMM is tensor-bound, i.e. where the minibatch is long and
  // the optimiser traffic (then a separate pass over this rank's slice) is small: tp_world >= 4.  GANMF_LOWRANK_DWE=0/1.
  int lowrank_dwe_mode = -1;
  bool lowrank_dwe = false;
  Mat M1, T1t;                // V^T . We  [k, E];  Pb^T . dHf  [k, E]
  // decoder-bias gradient from the residual GEMM's per-32-row column sums (Epilogue::colpart)
  float* colpart = nullptr; int colpart_rows = 0;
  bool colpart_on = true;     // GANMF_COLPART=0: always the colsum pass over the residual (A/B switch)
  bool colpart_ok = false;    // the last GEMM that was asked for partial column sums produced them
  // lazy user-factor optimiser (kernels.cuh K6b): p_last[row] = G step the row is current at, alpha_log[t -
  // log_base] = step size of G step t+1, g_T = G steps taken, p_stale = some row may lag behind g_T
  bool lazy_p = true;
  bool p_stale = false;
  int* p_last = nullptr;
  float* alpha_log = nullptr;
  int log_cap = 1 << 16, log_base = 0, g_T = 0;
  // live GEMM timing (bench roofline)
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  double prof_flops = 0;
  long long prof_launches = 0;
  std::vector<int> prof_shapes;
  TmapCache tmaps;
};

const char* ganmf_last_error(void) { return g_err; }
static int aux_join(ganmf_ctx* c);

template <typename T>
static int dalloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
  e = cudaMemset(*p, 0, n * sizeof(T));
  if (e != cudaSuccess) return fail("cudaMemset -> %s", cudaGetErrorString(e));
  return 0;
}
static int mat_alloc(Mat* m, int rows, int cols) {
  m->rows = rows; m->cols = cols; m->ld = rup(cols, 32);
  return dalloc(&m->p, m->elems());
}

// ------------------------------------------------------------------------------ GEMM dispatch
// out[M,N] = epilogue(A . B^T), A logical [M,K], B logical [N,K]; *_mn = stored transposed.
static int gemm(ganmf_ctx* c, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn,
                int M, int N, int K, const Epilogue& ep, int force_path = GANMF_GEMM_AUTO) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int path = force_path != GANMF_GEMM_AUTO ? force_path : c->cfg.gemm_path;
  const bool aligned = !(lda & 3) && !(ldb & 3) && !((uintptr_t)A & 15) && !((uintptr_t)B & 15);
  if (path == GANMF_GEMM_AUTO) {
    // DisGANMF trains on split-TF32 operands by default: the BCE gradients of the real and the fake half cancel in
    // the weight-gradient sums, so plain TF32 rounding shows at ~1e-2 on wide nets (tests/test_gpu_baseline_shapes.py)
    const int tc = c->cfg.kind == GANMF_KIND_DISGANMF ? GANMF_GEMM_TC3 : GANMF_GEMM_TC;
    path = (aligned && (double)M * N * K >= (double)(1 << 18)) ? tc : GANMF_GEMM_SIMT;
  }
  if (path == GANMF_GEMM_TC3 && c->presplit) path = GANMF_GEMM_TC;
  if (path == GANMF_GEMM_TC3 && (double)M * N * K < (double)(1 << 18)) path = GANMF_GEMM_SIMT;   // tiny: exact fp32 FMA
  if (path == GANMF_GEMM_TC && !aligned) return fail("tcgen05 GEMM needs 16-byte aligned operands");
  if (path == GANMF_GEMM_TC3) {
    // fp32-accurate tensor-core GEMM: x = hi + lo (tf32 each); A -> [hi|hi|lo], B -> [hi|lo|hi], one pass over 3K
    // computes hi.hi + hi.lo + lo.hi (the dropped lo.lo term is ~2^-22 relative).  3x the MMA work plus one
    // read/write of each operand; used where TF32 rounding is visible through cancellation (DisGANMF).
    const int ld3 = rup(3 * K, 32);
    const float* src[2] = {A, B};
    const int lds[2] = {lda, ldb}, mns[2] = {a_mn, b_mn}, ext[2] = {M, N};
    for (int o = 0; o < 2; ++o) {
      const size_t need = (size_t)ext[o] * ld3;
      if (need > c->s3_cap[o]) {
        CU(cudaStreamSynchronize(c->st));
        cudaFree(c->s3[o]);
        RC(dalloc(&c->s3[o], need));
        c->s3_cap[o] = need;
        c->tmaps = TmapCache();               // (descriptors of the freed buffer must not be reused)
      }
      split3_any_kernel<<<dim3((ext[o] + 31) / 32, (K + 31) / 32), 256, 0, c->st>>>(src[o], lds[o], mns[o], ext[o], K,
                                                                                 c->s3[o], ld3, o);
      CU(cudaGetLastError());
    }
    c->launches += 2;
    return gemm(c, c->s3[0], ld3, 0, c->s3[1], ld3, 0, M, N, 3 * K, ep, GANMF_GEMM_TC);
  }
  if (path == GANMF_GEMM_SIMT) {
    c->launches += 1;
    if (ep.colpart) c->colpart_ok = false;              // (tensor-core epilogue only: the caller falls back to colsum)
    CU(simt_gemm(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, c->st));
    return 0;
  }
  TcGemmCall g;
  g.A = A; g.lda = lda; g.a_mn = a_mn;
  g.B = B; g.ldb = ldb; g.b_mn = b_mn;
  g.M = M; g.N = N; g.K = K;
  g.ep = ep;
  g.bn = N > 128 ? 256 : 128;
  int tiles = ((M + TC_BM - 1) / TC_BM) * ((N + g.bn - 1) / g.bn);
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  // Small outputs with a long K (the split-K GEMMs) run 256-row CTA tiles: two accumulators share every
  // B stage (1.5x flops per byte from L2; measured +9..11 % at 2048x1024x27000); no accumulator
  // double-buffering is needed there because the epilogue is a small part of a long K loop.
  // (The same 256 x 256 tiles on CTA pairs -- six 32 KB stages instead of three of 64 KB -- were measured at cfg5:
  //  the split-K code-gradient GEMMs lose 3-13 %, so split-K stays on single-CTA 256-row tiles.)
  if (tiles < 148 && g.bn == 256 && M >= 256 && N >= 512 && total_kb >= 256) {
    g.mt = 2;
    tiles = ((M + 2 * TC_BM - 1) / (2 * TC_BM)) * ((N + g.bn - 1) / g.bn);
  }
  // Split K when the output has too few tiles to fill the 148 SMs.  Cost model per candidate split
  // count s: MMA rounds ceil(tiles*s/148) * (k-blocks per split) in units of one k-block of one tile,
  // plus the partial-sum traffic (s copies of the M x N output written and read again) converted to the
  // same unit (a 128 x BN x 32 tf32 k-block at ~600 TFLOP/s vs ~5 TB/s of workspace traffic).
  int splits = 1;
  if (tiles < 148 && total_kb >= 16) {
    const size_t per = (size_t)M * rup(N, 32);
    const int smax = std::min(std::min(total_kb / 8, 64), (int)std::min<size_t>(c->ws_floats / per, 64));
    const double kb_us = 2.0 * g.mt * TC_BM * g.bn * TC_BK / (600e6 / 148.0);   // us per k-block per tile on one SM
    double best = 1e30;
    for (int sp = 1; sp <= smax; ++sp) {
      const int kbps = (total_kb + sp - 1) / sp;
      if ((total_kb + kbps - 1) / kbps != sp) continue;
      const int rounds = (tiles * sp + 147) / 148;
      double cost = rounds * kbps * kb_us;
      if (sp > 1) cost += 2.0 * sp * per * 4 / 5e6 + 3.0;                     // workspace traffic + reduce launch
      if (cost < best) { best = cost; splits = sp; }
    }
  }
  // Many-tile GEMMs are bound by the L2 -> SM fill rate: CTA pairs (256 x 256 tiles over two SMs, each CTA
  // loading half of B) need 1.5x fewer bytes per flop and keep the accumulator double-buffered.
  if (c->pair_mode && g.bn == 256 && g.mt == 1 && splits == 1 && M > TC_BM && (c->pair_mode == 2 || tiles >= 148))
    g.cg = 2;
  if (ep.colpart) {                                     // partial column sums: unsplit launches only
    c->colpart_ok = splits == 1 && tc_lean_epilogue(ep);
    if (!c->colpart_ok) g.ep.colpart = nullptr;
  }
  g.splits = splits;
  g.ws = c->ws;
  g.cache = &c->tmaps;
  g.max_ctas = c->gemm_sm_cap;
  c->launches += splits > 1 ? 2 : 1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (c->profile) {
    while (c->ev_pool.size() < c->ev_used + 2) {
      cudaEvent_t ev;
      CU(cudaEventCreate(&ev));
      c->ev_pool.push_back(ev);
    }
    ev0 = c->ev_pool[c->ev_used++];
    ev1 = c->ev_pool[c->ev_used++];
    CU(cudaEventRecord(ev0, c->st));
  }
  cudaError_t e = tc_gemm(g, c->st);
  if (e != cudaSuccess) return fail("tc_gemm(M=%d N=%d K=%d) -> %s", M, N, K, cudaGetErrorString(e));
  if (c->profile) {
    CU(cudaEventRecord(ev1, c->st));
    c->prof_flops += 2.0 * M * N * K;
    c->prof_launches += 1;
    c->prof_shapes.insert(c->prof_shapes.end(), {M, N, K, splits});
  }
  return 0;
}

// ------------------------------------------------------------------------------ create/destroy
// sliced: 0 = whole tensor, 1 = rows are items (slice of the rows), 2 = columns are items
static void add_param(ganmf_ctx* c, const char* name, int rows, int cols, int is_gen, int sliced = 0) {
  Param p;
  p.name = name;
  p.w.rows = rows; p.w.cols = cols; p.w.ld = rup(cols, 32);
  p.is_gen = is_gen;
  p.rows_g = sliced == 1 ? c->Wg : rows;
  p.cols_g = sliced == 2 ? c->Wg : cols;
  p.row_off = sliced == 1 ? c->cfg.item_offset : 0;
  p.col_off = sliced == 2 ? c->cfg.item_offset : 0;
  c->params.push_back(p);
}

static int create_buffers(ganmf_ctx* c);

int ganmf_create(const ganmf_config* cfg, ganmf_ctx** out) {
  if (!cfg || !out) return fail("null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("no CUDA device: ganmf_b200 has no CPU path (%s)", cudaGetErrorString(e));
  CU(cudaSetDevice(cfg->device));
  if (cfg->n_rows <= 0 || cfg->width <= 0 || cfg->num_factors <= 0 || cfg->max_batch <= 0)
    return fail("bad config");
  ganmf_ctx* c = new ganmf_ctx();
  c->cfg = *cfg;
  if (const char* nf = getenv("GANMF_NO_FUSED_ADAM")) c->fuse_adam = !(nf[0] == '1');   // A/B switch
  if (const char* pm = getenv("GANMF_PAIR")) c->pair_mode = atoi(pm);                  // A/B switch / tests
  if (const char* nl = getenv("GANMF_NO_LAZY_ADAM")) c->lazy_p = !(nl[0] == '1');       // A/B switch
  if (const char* lc = getenv("GANMF_LAZY_LOG_CAP")) c->log_cap = std::max(1, atoi(lc)); // tests: force log wrap
  if (const char* ef = getenv("GANMF_EVAL_FUSED")) c->eval_fused = !(ef[0] == '0');     // A/B switch / tests
  if (const char* sr = getenv("GANMF_SPARSE_REAL")) c->sparse_mode = atoi(sr);          // A/B switch / tests
  if (const char* cp = getenv("GANMF_COLPART")) c->colpart_on = !(cp[0] == '0');        // A/B switch / tests
  if (const char* lr = getenv("GANMF_LOWRANK")) c->lowrank_mode = atoi(lr);             // A/B switch / tests
  if (const char* ld = getenv("GANMF_LOWRANK_DWE")) c->lowrank_dwe_mode = atoi(ld);     // A/B switch / tests
  if (const char* gr = getenv("GANMF_GEN_RESIDENT")) c->gen_resident_mode = atoi(gr);   // A/B switch / tests
  if (const char* ax = getenv("GANMF_AUX_STREAM")) c->aux_on = !(ax[0] == '0');         // A/B switch / tests
  c->B = cfg->max_batch; c->W = cfg->width; c->Wp = rup(cfg->width, 32);
  c->k = cfg->num_factors; c->kp = rup(c->k, 32);
  c->Wg = cfg->global_width > 0 ? cfg->global_width : cfg->width;
  c->tp_world = cfg->tp_world > 1 ? cfg->tp_world : 1;
  c->tp_rank = c->tp_world > 1 ? cfg->tp_rank : 0;
  if (c->tp_world > 1 && (cfg->kind != GANMF_KIND_GANMF || cfg->tp_rank < 0 ||
                          cfg->tp_rank >= c->tp_world || cfg->item_offset < 0 ||
                          cfg->item_offset + cfg->width > c->Wg)) {
    delete c;
    return fail("item-sharded contexts: GANMF only, 0 <= tp_rank < tp_world, slice inside global_width");
  }
  if (c->tp_world == 1 && (c->Wg != c->W || cfg->item_offset != 0)) { delete c; return fail("global_width / item_offset need tp_world > 1"); }
  const int sl = c->tp_world > 1;
  if (cfg->kind == GANMF_KIND_GANMF) {
    if (cfg->emb_dim <= 0) { delete c; return fail("emb_dim must be > 0"); }
    c->E = cfg->emb_dim; c->Ep = rup(c->E, 32);
    add_param(c, "autoencoder/encoding/kernel", c->W, c->E, 0, sl ? 1 : 0);
    add_param(c, "autoencoder/encoding/bias", 1, c->E, 0);
    add_param(c, "autoencoder/decoding/kernel", c->E, c->W, 0, sl ? 2 : 0);
    add_param(c, "autoencoder/decoding/bias", 1, c->W, 0, sl ? 2 : 0);
  } else if (cfg->kind == GANMF_KIND_MF) {
    c->E = 1; c->Ep = 32;                 // no discriminator: factor matrices only
  } else if (cfg->kind == GANMF_KIND_DISGANMF) {
    if (cfg->d_layers < 1 || cfg->d_nodes < 1) { delete c; return fail("bad discriminator shape"); }
    c->E = cfg->d_nodes; c->Ep = rup(c->E, 32);
    int fan_in = c->W + 1;
    char nm[96];
    for (int l = 0; l < cfg->d_layers; ++l) {
      snprintf(nm, sizeof nm, "discriminator/layer_%d/kernel", l);
      add_param(c, nm, fan_in, cfg->d_nodes, 0);
      snprintf(nm, sizeof nm, "discriminator/layer_%d/bias", l);
      add_param(c, nm, 1, cfg->d_nodes, 0);
      fan_in = cfg->d_nodes;
    }
    add_param(c, "discriminator/D_output/kernel", fan_in, 1, 0);
    add_param(c, "discriminator/D_output/bias", 1, 1, 0);
  } else {
    delete c;
    return fail("unknown kind %d", cfg->kind);
  }
  c->n_d = (int)c->params.size();
  add_param(c, "generator/user_embeddings", cfg->n_rows, c->k, 1);
  add_param(c, "generator/item_embeddings", c->W, c->k, 1, sl ? 1 : 0);

  const int rc_alloc = create_buffers(c);
  if (rc_alloc) { ganmf_destroy(c); return rc_alloc; }     // nothing allocated so far is leaked
  *out = c;
  return 0;
}

static int create_buffers(ganmf_ctx* c) {
  const ganmf_config* cfg = &c->cfg;
  // discriminator slab: theta | m | v | grad | best, each d_elems floats, tensors back to back
  for (int i = 0; i < c->n_d; ++i) c->d_elems += c->params[i].w.elems();
  RC(dalloc(&c->d_slab, 5 * c->d_elems));
  size_t off = 0;
  for (int i = 0; i < c->n_d; ++i) {
    Param& p = c->params[i];
    p.w.p = c->d_slab + off;
    p.m = c->d_slab + c->d_elems + off;
    p.v = c->d_slab + 2 * c->d_elems + off;
    p.g = c->d_slab + 3 * c->d_elems + off;
    p.best = c->d_slab + 4 * c->d_elems + off;
    off += p.w.elems();
  }
  Param& P = c->params[c->n_d];
  Param& V = c->params[c->n_d + 1];
  c->p_elems = P.w.elems();
  RC(dalloc(&c->p_slab, 4 * c->p_elems));
  P.w.p = c->p_slab; P.m = c->p_slab + c->p_elems; P.v = c->p_slab + 2 * c->p_elems;
  P.best = c->p_slab + 3 * c->p_elems;
  c->v_elems = V.w.elems();
  RC(dalloc(&c->v_slab, 5 * c->v_elems));
  V.w.p = c->v_slab; V.m = c->v_slab + c->v_elems; V.v = c->v_slab + 2 * c->v_elems;
  V.g = c->v_slab + 3 * c->v_elems; V.best = c->v_slab + 4 * c->v_elems;

  const int B = c->B;
  RC(mat_alloc(&c->X2, 2 * B, c->W));
  RC(mat_alloc(&c->Pb, B, c->k));
  RC(mat_alloc(&c->dPb, B, c->k));
  RC(mat_alloc(&c->dF, B, c->W));
  if (cfg->kind == GANMF_KIND_GANMF) {
    RC(mat_alloc(&c->H2, 2 * B, c->E));
    RC(mat_alloc(&c->H2s, 2 * B, c->E));
    RC(mat_alloc(&c->dH2, 2 * B + 1, c->E));      // + one row: partial encoder-bias gradient of an item-sharded step
    RC(mat_alloc(&c->Res2, 2 * B, c->W));
    c->colpart_rows = (2 * B + 31) / 32;
    RC(dalloc(&c->colpart, (size_t)c->colpart_rows * c->Wp));
    c->lowrank = c->lowrank_mode == 1 || (c->lowrank_mode < 0 && 2 * c->k <= B);
    c->lowrank_dwe = c->tp_world > 1 && (c->lowrank_dwe_mode == 1 || (c->lowrank_dwe_mode < 0 && c->lowrank && c->tp_world >= 4));
    RC(mat_alloc(&c->M1, c->k, c->E));
    RC(mat_alloc(&c->T1t, c->k, c->E));
  } else if (cfg->kind == GANMF_KIND_DISGANMF) {
    c->hs.resize(cfg->d_layers);
    c->dzs.resize(cfg->d_layers);
    for (int l = 0; l < cfg->d_layers; ++l) {
      RC(mat_alloc(&c->hs[l], 2 * B, cfg->d_nodes));
      RC(mat_alloc(&c->dzs[l], 2 * B, cfg->d_nodes));
    }
    RC(mat_alloc(&c->dhtmp, 2 * B, cfg->d_nodes));
    RC(dalloc(&c->out2, (size_t)rup(2 * B, 32)));
    RC(dalloc(&c->dout2, (size_t)rup(2 * B, 32)));
    RC(dalloc(&c->idf, (size_t)rup(2 * B, 32)));
  }
  if (c->aux_on) {
    CU(cudaStreamCreateWithFlags(&c->st_aux, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  c->ws_floats = (size_t)48 << 20;     // 192 MB of split-K partials
  RC(dalloc(&c->ws, c->ws_floats));
  RC(dalloc(&c->sc, 1));
  c->losses_cap = 1 << 16;
  RC(dalloc(&c->losses, (size_t)c->losses_cap));
  c->ids_cap = std::max(cfg->n_rows, 2 * B);
  RC(dalloc(&c->ids, (size_t)c->ids_cap));
  RC(dalloc(&c->slot, (size_t)cfg->n_rows));
  RC(dalloc(&c->p_last, (size_t)cfg->n_rows));
  RC(dalloc(&c->alpha_log, (size_t)c->log_cap));
  fill_int_kernel<<<(cfg->n_rows + 255) / 256, 256>>>(c->slot, -1, (size_t)cfg->n_rows);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return 0;
}

static void csr_free(Csr& m) {
  cudaFree(m.indptr); cudaFree(m.indices); cudaFree(m.data);
  m = Csr();
}

void ganmf_destroy(ganmf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  cudaFree(c->d_slab); cudaFree(c->p_slab); cudaFree(c->v_slab);
  cudaFree(c->p_last); cudaFree(c->alpha_log); cudaFree(c->colpart);
  for (Mat* m : {&c->X2, &c->H2, &c->H2s, &c->Res2, &c->dH2, &c->dF, &c->Pb, &c->dPb, &c->dhtmp, &c->Fb, &c->Ob, &c->M1,
                 &c->T1t})
    cudaFree(m->p);
  for (auto& m : c->hs) cudaFree(m.p);
  for (auto& m : c->dzs) cudaFree(m.p);
  cudaFree(c->out2); cudaFree(c->dout2); cudaFree(c->idf);
  cudaFree(c->s3[0]); cudaFree(c->s3[1]);
  cudaFree(c->ws); cudaFree(c->sc); cudaFree(c->losses); cudaFree(c->ids); cudaFree(c->slot);
  for (int i = 0; i < 3; ++i) csr_free(c->csr[i]);
  cudaFree(c->tb_gain); cudaFree(c->tb_gain_desc); cudaFree(c->tb_logtab); cudaFree(c->tb_nov);
  cudaFree(c->tb_popn); cudaFree(c->tb_haspop); cudaFree(c->rmse_scratch);
  cudaFree(c->scores); cudaFree(c->topk_idx); cudaFree(c->topk_val); cudaFree(c->uvals);
  cudaFree(c->usums); cudaFree(c->icounts); cudaFree(c->cut_dev); cudaFree(c->eval_users);
  for (auto& b : c->fbuf) {
    cudaFree(b.Qg.p); cudaFree(b.cand_val); cudaFree(b.cand_idx); cudaFree(b.row_thr); cudaFree(b.fb_rows);
    cudaFree(b.fb_count); cudaFree(b.topk_idx); cudaFree(b.topk_val);
    if (b.h_count) cudaFreeHost(b.h_count);
    for (cudaEvent_t ev : {b.ev_sel, b.ev_res, b.ev_fin}) if (ev) cudaEventDestroy(ev);
  }
  if (c->st_aux) cudaStreamDestroy(c->st_aux);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->st_res) cudaStreamDestroy(c->st_res);
  if (c->st_fin) cudaStreamDestroy(c->st_fin);
  if (c->ev_setup) cudaEventDestroy(c->ev_setup);
  if (c->ev_pipe_done) cudaEventDestroy(c->ev_pipe_done);
  cudaFree(c->vmax_bits); cudaFree(c->fb_users); cudaFree(c->fb_idx); cudaFree(c->fb_val);
  for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
  delete c;
}

int ganmf_set_stream(ganmf_ctx* c, void* s) {
  if (!c) return fail("null ctx");
  c->st = (cudaStream_t)s;
  return 0;
}
int ganmf_set_gemm_sms(ganmf_ctx* c, int n_sms) {
  if (!c || n_sms < 0) return fail("bad argument");
  c->gemm_sm_cap = n_sms;
  return 0;
}
int ganmf_synchronize(ganmf_ctx* c) {
  RC(aux_join(c));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}
int64_t ganmf_launch_count(ganmf_ctx* c) { return c ? c->launches : 0; }
int ganmf_profile(ganmf_ctx* c, int enable) {
  if (!c) return fail("null ctx");
  CU(cudaStreamSynchronize(c->st));
  c->profile = enable != 0;
  c->ev_used = 0; c->prof_flops = 0; c->prof_launches = 0; c->prof_shapes.clear();
  return 0;
}
int ganmf_profile_records(ganmf_ctx* c, double* ms, int32_t* shape, int cap, int* n) {
  if (!c || !n) return fail("null argument");
  CU(cudaStreamSynchronize(c->st));
  int k = 0;
  for (size_t i = 0; i + 1 < c->ev_used && k < cap; i += 2, ++k) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, c->ev_pool[i], c->ev_pool[i + 1]));
    if (ms) ms[k] = t;
    if (shape) memcpy(shape + 4 * k, c->prof_shapes.data() + 4 * k, 16);
  }
  *n = k;
  return 0;
}
int ganmf_profile_read(ganmf_ctx* c, double* ms, double* flops, int64_t* launches) {
  if (!c) return fail("null ctx");
  CU(cudaStreamSynchronize(c->st));
  double total = 0;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, c->ev_pool[i], c->ev_pool[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (flops) *flops = c->prof_flops;
  if (launches) *launches = c->prof_launches;
  c->ev_used = 0; c->prof_flops = 0; c->prof_launches = 0; c->prof_shapes.clear();
  return 0;
}

// ------------------------------------------------------------------------------ data
// Route of the real rows' codes.  The gather-sum reads 4*E bytes of We per interaction, the dense
// product spends 2*E flops per matrix cell: on a B200 (~5.6 TB/s of gathered weight rows against ~600 TFLOP/s of tf32
// MMAs inside the step) the two meet near 0.5 % density.  Measured on both benchmark shapes: cfg5 (0.1 %) 12.14 ->
// 11.36 ms per step pair, cfg4 (0.5 %, We = 110 MB: partly L2-resident) 1.750 -> 1.644 ms; the sparse route is taken
// up to 0.55 % (the committed MovieLens / LastFM splits, 0.3-4.5 % dense with tiny We, stay dense or do not care).
static void note_train_csr(ganmf_ctx* c) {
  const Csr& m = c->csr[GANMF_CSR_TRAIN];
  const double cells = (double)m.n_rows * (double)m.n_cols;
  const double density = cells > 0 ? (double)m.nnz / cells : 1.0;
  c->sparse_real = c->cfg.kind == GANMF_KIND_GANMF &&
                   (c->sparse_mode == 1 || (c->sparse_mode < 0 && density <= 0.0055));
}

int ganmf_set_csr(ganmf_ctx* c, int which, int n_rows, int n_cols, const int32_t* indptr,
                  const int32_t* indices, const float* data) {
  if (!c || which < 0 || which > 2 || !indptr) return fail("bad argument");
  if (which == GANMF_CSR_TRAIN && (n_rows != c->cfg.n_rows || n_cols != c->W))
    return fail("train CSR is %dx%d, context expects %dx%d", n_rows, n_cols, c->cfg.n_rows, c->W);
  Csr& m = c->csr[which];
  csr_free(m);
  m.n_rows = n_rows; m.n_cols = n_cols; m.nnz = indptr[n_rows];
  RC(dalloc(&m.indptr, (size_t)n_rows + 1));
  RC(dalloc(&m.indices, (size_t)m.nnz));
  CU(cudaMemcpy(m.indptr, indptr, ((size_t)n_rows + 1) * 4, cudaMemcpyHostToDevice));
  if (m.nnz) CU(cudaMemcpy(m.indices, indices, (size_t)m.nnz * 4, cudaMemcpyHostToDevice));
  if (data) {
    RC(dalloc(&m.data, (size_t)m.nnz));
    if (m.nnz) CU(cudaMemcpy(m.data, data, (size_t)m.nnz * 4, cudaMemcpyHostToDevice));
  }
  if (which == GANMF_CSR_TEST) c->have_tables = false;
  if (which == GANMF_CSR_TRAIN) note_train_csr(c);
  return 0;
}

int ganmf_set_csr_device(ganmf_ctx* c, int which, int n_rows, int n_cols, const int32_t* indptr_dev,
                         const int32_t* indices_dev, const float* data_dev, int64_t nnz) {
  if (!c || which < 0 || which > 2 || !indptr_dev || nnz < 0 || (nnz > 0 && !indices_dev)) return fail("bad argument");
  if (which == GANMF_CSR_TRAIN && (n_rows != c->cfg.n_rows || n_cols != c->W))
    return fail("train CSR is %dx%d, context expects %dx%d", n_rows, n_cols, c->cfg.n_rows, c->W);
  int last = 0;
  CU(cudaMemcpy(&last, indptr_dev + n_rows, 4, cudaMemcpyDeviceToHost));
  if ((int64_t)last != nnz) return fail("indptr[n_rows] = %d but nnz = %lld", last, (long long)nnz);
  Csr& m = c->csr[which];
  csr_free(m);
  m.n_rows = n_rows; m.n_cols = n_cols; m.nnz = nnz;
  RC(dalloc(&m.indptr, (size_t)n_rows + 1));
  RC(dalloc(&m.indices, (size_t)nnz));
  CU(cudaMemcpy(m.indptr, indptr_dev, ((size_t)n_rows + 1) * 4, cudaMemcpyDeviceToDevice));
  if (nnz) CU(cudaMemcpy(m.indices, indices_dev, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
  if (data_dev) {
    RC(dalloc(&m.data, (size_t)nnz));
    if (nnz) CU(cudaMemcpy(m.data, data_dev, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
  }
  if (which == GANMF_CSR_TEST) c->have_tables = false;
  if (which == GANMF_CSR_TRAIN) note_train_csr(c);
  return 0;
}

// ------------------------------------------------------------------------------ side stream
// work enqueued on the returned stream starts after everything enqueued on c->st so far
static int aux_fork(ganmf_ctx* c, cudaStream_t* s) {
  *s = c->st;
  if (!c->aux_on) return 0;
  CU(cudaEventRecord(c->ev_fork, c->st));
  CU(cudaStreamWaitEvent(c->st_aux, c->ev_fork, 0));
  c->aux_pending = true;
  *s = c->st_aux;
  return 0;
}
// everything enqueued on c->st from here on waits for the side stream
static int aux_join(ganmf_ctx* c) {
  if (!c->aux_pending) return 0;
  CU(cudaEventRecord(c->ev_join, c->st_aux));
  CU(cudaStreamWaitEvent(c->st, c->ev_join, 0));
  c->aux_pending = false;
  return 0;
}

int ganmf_get_csr(ganmf_ctx* c, int which, int32_t* n_rows, int32_t* n_cols, int64_t* nnz, int32_t* indptr,
                  int32_t* indices, float* data) {
  if (!c || which < 0 || which > 2) return fail("bad argument");
  const Csr& m = c->csr[which];
  if (!m.indptr) return fail("CSR %d not set", which);
  if (n_rows) *n_rows = m.n_rows;
  if (n_cols) *n_cols = m.n_cols;
  if (nnz) *nnz = m.nnz;
  CU(cudaStreamSynchronize(c->st));
  if (indptr) CU(cudaMemcpy(indptr, m.indptr, ((size_t)m.n_rows + 1) * 4, cudaMemcpyDeviceToHost));
  if (indices && m.nnz) CU(cudaMemcpy(indices, m.indices, (size_t)m.nnz * 4, cudaMemcpyDeviceToHost));
  if (data && m.nnz) {
    if (!m.data) return fail("CSR %d holds no values (implicit ones)", which);
    CU(cudaMemcpy(data, m.data, (size_t)m.nnz * 4, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// CSR `which` := transpose of the given host CSR [n_rows x n_cols], built on the device (csr_kernels.cuh).
int ganmf_set_csr_transposed(ganmf_ctx* c, int which, int n_rows, int n_cols, const int32_t* indptr,
                             const int32_t* indices, const float* data) {
  if (!c || which < 0 || which > 2 || !indptr || n_rows < 0 || n_cols < 0) return fail("bad argument");
  if (which == GANMF_CSR_TRAIN && (n_cols != c->cfg.n_rows || n_rows != c->W))
    return fail("transposed train CSR is %dx%d, context expects %dx%d", n_cols, n_rows, c->cfg.n_rows, c->W);
  const long long nnz = indptr[n_rows];
  if (nnz < 0 || (nnz > 0 && !indices)) return fail("bad argument");
  for (long long i = 0; i < nnz; ++i)
    if (indices[i] < 0 || indices[i] >= n_cols) return fail("column index %d at position %lld outside [0, %d)", indices[i], i, n_cols);
  int *s_ip = nullptr, *s_ix = nullptr, *cursor = nullptr;
  float* s_dat = nullptr;
  Csr t;
  t.n_rows = n_cols; t.n_cols = n_rows; t.nnz = nnz;
  auto cleanup = [&]() { cudaFree(s_ip); cudaFree(s_ix); cudaFree(s_dat); cudaFree(cursor); };
  auto bail = [&](int rc) { cleanup(); csr_free(t); return rc; };
#define TRY(x) do { int r__ = (x); if (r__) return bail(r__); } while (0)
#define TRYCU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return bail(fail("%s -> %s", #x, cudaGetErrorString(e__))); } while (0)
  TRY(dalloc(&s_ip, (size_t)n_rows + 1));
  TRY(dalloc(&s_ix, (size_t)nnz));
  TRY(dalloc(&cursor, (size_t)n_cols + 1));
  TRY(dalloc(&t.indptr, (size_t)n_cols + 1));
  TRY(dalloc(&t.indices, (size_t)nnz));
  TRYCU(cudaMemcpyAsync(s_ip, indptr, ((size_t)n_rows + 1) * 4, cudaMemcpyHostToDevice, c->st));
  if (nnz) TRYCU(cudaMemcpyAsync(s_ix, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->st));
  if (data) {
    TRY(dalloc(&s_dat, (size_t)nnz));
    TRY(dalloc(&t.data, (size_t)nnz));
    if (nnz) TRYCU(cudaMemcpyAsync(s_dat, data, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->st));
  }
  if (nnz) {
    csr_col_count_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c->st>>>(s_ix, nnz, cursor);
    TRYCU(cudaGetLastError());
  }
  exclusive_scan_kernel<<<1, 1024, 0, c->st>>>(cursor, n_cols, t.indptr);
  TRYCU(cudaGetLastError());
  TRYCU(cudaMemsetAsync(cursor, 0, ((size_t)n_cols + 1) * 4, c->st));
  if (nnz && n_rows) {
    csr_scatter_kernel<<<(unsigned)(((long long)n_rows * 32 + 255) / 256), 256, 0, c->st>>>(s_ip, s_ix, s_dat, n_rows, t.indptr,
                                                                                          cursor, t.indices, t.data);
    TRYCU(cudaGetLastError());
    csr_sort_segments_kernel<<<n_cols, TR_THREADS, 0, c->st>>>(t.indptr, t.indices, t.data);
    TRYCU(cudaGetLastError());
  }
  c->launches += 4;
  // columns with more than TR_CAP entries (popular items): found on the host from the new indptr, sorted in a
  // padded global scratch, one CTA each
  std::vector<int> tip((size_t)n_cols + 1);
  TRYCU(cudaMemcpyAsync(tip.data(), t.indptr, ((size_t)n_cols + 1) * 4, cudaMemcpyDeviceToHost, c->st));
  TRYCU(cudaStreamSynchronize(c->st));
  std::vector<int> longs;
  int max_len = 0;
  for (int j = 0; j < n_cols; ++j) {
    const int len = tip[j + 1] - tip[j];
    if (len > TR_CAP) { longs.push_back(j); max_len = std::max(max_len, len); }
  }
  if (!longs.empty()) {
    int cap2 = 2;
    while (cap2 < max_len) cap2 <<= 1;
    int *seg = nullptr, *sk = nullptr;
    float* sv = nullptr;
    auto bail2 = [&](int rc) { cudaFree(seg); cudaFree(sk); cudaFree(sv); return bail(rc); };
    // (scratch for a bounded number of long columns at a time: a CTA each)
    const size_t per_launch = std::max<size_t>(1, ((size_t)256 << 20) / ((size_t)cap2 * 8));
    int r1 = dalloc(&seg, longs.size());
    if (!r1) r1 = dalloc(&sk, std::min(per_launch, longs.size()) * cap2);
    if (!r1 && t.data) r1 = dalloc(&sv, std::min(per_launch, longs.size()) * cap2);
    if (r1) return bail2(r1);
    cudaError_t e = cudaMemcpyAsync(seg, longs.data(), longs.size() * 4, cudaMemcpyHostToDevice, c->st);
    for (size_t o = 0; e == cudaSuccess && o < longs.size(); o += per_launch) {
      const unsigned nb = (unsigned)std::min(per_launch, longs.size() - o);
      csr_sort_long_segments_kernel<<<nb, TR_THREADS, 0, c->st>>>(seg + o, t.indptr, t.indices, t.data, sk, sv, cap2);
      e = cudaGetLastError();
      c->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    cudaFree(seg); cudaFree(sk); cudaFree(sv);
    if (e != cudaSuccess) return bail(fail("long-segment sort -> %s", cudaGetErrorString(e)));
  }
#undef TRY
#undef TRYCU
  cleanup();
  Csr& m = c->csr[which];
  csr_free(m);
  m = t;
  if (which == GANMF_CSR_TEST) c->have_tables = false;
  if (which == GANMF_CSR_TRAIN) note_train_csr(c);
  return 0;
}

// ------------------------------------------------------------------------------ lazy user factors
// Replay the deferred zero-gradient Adam steps: for the rows in ids (device), or for every row.
static int p_catchup(ganmf_ctx* c, const int* ids_dev, int n, cudaStream_t st = nullptr) {
  if (!c->p_stale || n <= 0) return 0;
  if (!st) st = c->st;
  const Param& P = c->params[c->n_d];
  const int grid = ids_dev ? n : std::min(n, 148 * 16);
  p_catchup_kernel<<<grid, 64, 0, st>>>(P.w.p, P.m, P.v, P.w.ld, ids_dev, n, c->p_last, c->alpha_log, c->g_T,
                                          c->log_base);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}
// Every reader of the whole matrix (export, snapshot, scoring, a dense optimiser step) calls this first.
static int p_flush(ganmf_ctx* c) {
  RC(aux_join(c));
  if (!c->p_stale) return 0;
  RC(p_catchup(c, nullptr, c->cfg.n_rows));
  c->p_stale = false;
  c->log_base = c->g_T;                  // nothing older than g_T is ever replayed again
  return 0;
}

// ------------------------------------------------------------------------------ parameters
static Param* find_param(ganmf_ctx* c, const char* name) {
  for (auto& p : c->params)
    if (p.name == name) return &p;
  return nullptr;
}
int ganmf_param_count(ganmf_ctx* c) { return c ? (int)c->params.size() : 0; }
int ganmf_param_info(ganmf_ctx* c, int i, char* name, int cap, int* rows, int* cols, int* is_gen) {
  if (!c || i < 0 || i >= (int)c->params.size()) return fail("bad index");
  const Param& p = c->params[i];
  if (name && cap > 0) { strncpy(name, p.name.c_str(), cap - 1); name[cap - 1] = 0; }
  if (rows) *rows = p.w.rows;
  if (cols) *cols = p.w.cols;
  if (is_gen) *is_gen = p.is_gen;
  return 0;
}
int ganmf_set_param(ganmf_ctx* c, const char* name, const float* host, int64_t count) {
  Param* p = c ? find_param(c, name) : nullptr;
  if (!p) return fail("unknown parameter %s", name ? name : "(null)");
  if (count != (int64_t)p->w.rows * p->w.cols) return fail("%s: expected %d x %d", name, p->w.rows, p->w.cols);
  RC(p_flush(c));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaMemcpy2D(p->w.p, (size_t)p->w.ld * 4, host, (size_t)p->w.cols * 4, (size_t)p->w.cols * 4,
                  p->w.rows, cudaMemcpyHostToDevice));
  return 0;
}
int ganmf_get_param(ganmf_ctx* c, const char* name, float* host, int64_t count) {
  Param* p = c ? find_param(c, name) : nullptr;
  if (!p) return fail("unknown parameter %s", name ? name : "(null)");
  if (count != (int64_t)p->w.rows * p->w.cols) return fail("%s: expected %d x %d", name, p->w.rows, p->w.cols);
  RC(p_flush(c));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaMemcpy2D(host, (size_t)p->w.cols * 4, p->w.p, (size_t)p->w.ld * 4, (size_t)p->w.cols * 4,
                  p->w.rows, cudaMemcpyDeviceToHost));
  return 0;
}

// counter-based uniform generator (splitmix64 finaliser): U(-lim, lim) on the real columns only
// (element (r, cc) of a slice draws the value of element (r + row_off, cc + col_off) of the whole tensor)
__global__ void glorot_kernel(float* w, int rows, int cols, int ld, float lim, unsigned long long seed,
                              int row_off, int col_off, int cols_g) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), cc = (int)(i % cols);
  const unsigned long long ig = (unsigned long long)(r + row_off) * (unsigned long long)cols_g + (cc + col_off);
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ig + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);     // [0, 1)
  w[(size_t)r * ld + cc] = (2.f * u - 1.f) * lim;
}
int ganmf_init_params(ganmf_ctx* c, uint64_t seed) {
  if (!c) return fail("null ctx");
  c->p_stale = false;                     // every tensor and moment is replaced: nothing deferred survives
  int i = 0;
  for (auto& p : c->params) {
    CU(cudaMemsetAsync(p.w.p, 0, p.w.elems() * 4, c->st));
    if (p.w.rows > 1 || p.name.find("kernel") != std::string::npos) {     // matrices; biases stay zero
      const float lim = sqrtf(6.0f / (float)(p.rows_g + p.cols_g));
      const size_t n = (size_t)p.w.rows * p.w.cols;
      glorot_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(p.w.p, p.w.rows, p.w.cols, p.w.ld, lim,
                                                                   seed * 1000003ull + 7919ull * (++i), p.row_off,
                                                                   p.col_off, p.cols_g);
      CU(cudaGetLastError());
      c->launches++;
    }
  }
  return ganmf_reset_optimizers(c);
}
int ganmf_reset_optimizers(ganmf_ctx* c) {
  if (!c) return fail("null ctx");
  RC(p_flush(c));                         // deferred steps still use the OLD moments: apply them before zeroing
  c->g_T = 0; c->log_base = 0;
  CU(cudaMemsetAsync(c->p_last, 0, (size_t)c->cfg.n_rows * 4, c->st));
  CU(cudaMemsetAsync(c->d_slab + c->d_elems, 0, 3 * c->d_elems * 4, c->st));   // m, v, grad
  CU(cudaMemsetAsync(c->p_slab + c->p_elems, 0, 2 * c->p_elems * 4, c->st));
  CU(cudaMemsetAsync(c->v_slab + c->v_elems, 0, 3 * c->v_elems * 4, c->st));
  c->b1p[0] = c->b1p[1] = ADAM_B1;
  c->b2p[0] = c->b2p[1] = ADAM_B2;
  return 0;
}
int ganmf_snapshot(ganmf_ctx* c) {
  RC(p_flush(c));
  CU(cudaMemcpyAsync(c->d_slab + 4 * c->d_elems, c->d_slab, c->d_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->p_slab + 3 * c->p_elems, c->p_slab, c->p_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->v_slab + 4 * c->v_elems, c->v_slab, c->v_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  c->have_best = true;
  return 0;
}
int ganmf_restore(ganmf_ctx* c) {
  // the reference's shadow variables exist from graph construction with their own random init
  // (GANMF.py:123-128); restoring before any snapshot is therefore refused instead of guessed
  if (!c->have_best) return fail("load_model() before any save_current_model()");
  RC(p_flush(c));                          // theta is replaced, the moments stay: bring them to the current step first
  CU(cudaMemcpyAsync(c->d_slab, c->d_slab + 4 * c->d_elems, c->d_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->p_slab, c->p_slab + 3 * c->p_elems, c->p_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->v_slab, c->v_slab + 4 * c->v_elems, c->v_elems * 4, cudaMemcpyDeviceToDevice, c->st));
  return 0;
}

// ------------------------------------------------------------------------------ training
int ganmf_upload_ids(ganmf_ctx* c, const int32_t* ids, int n) {
  if (!c || n < 0 || n > c->ids_cap) return fail("upload_ids: n=%d exceeds capacity %d", n, c ? c->ids_cap : 0);
  for (int i = 0; i < n; ++i)              // the gather / optimiser kernels index with these unchecked
    if (ids[i] < 0 || ids[i] >= c->cfg.n_rows) return fail("row id %d at position %d outside [0, %d)", ids[i], i, c->cfg.n_rows);
  CU(cudaMemcpyAsync(c->ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  return 0;
}

static float adam_alpha(ganmf_ctx* c, int which, float lr) {
  // TF: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), beta powers kept as fp32 variables
  const float a = lr * sqrtf(1.f - c->b2p[which]) / (1.f - c->b1p[which]);
  c->b1p[which] *= ADAM_B1;
  c->b2p[which] *= ADAM_B2;
  return a;
}

static int check_batch(ganmf_ctx* c, int ids_offset, int B, bool tp_call = false) {
  if (!c) return fail("null ctx");
  if (c->cfg.kind == GANMF_KIND_MF) return fail("a factor-only (GANMF_KIND_MF) context does not train");
  if (c->tp_world > 1 && !tp_call)
    return fail("item-sharded context: use ganmf_tp_d_phase / ganmf_tp_g_phase (partial sums must be all-reduced)");
  if (B <= 0 || B > c->B) return fail("batch %d outside (0, %d]", B, c->B);
  if (ids_offset < 0 || ids_offset + B > c->ids_cap) return fail("ids range out of bounds");
  if (!c->csr[GANMF_CSR_TRAIN].indptr) return fail("train CSR not set");
  return 0;
}

// real profiles -> X2[0:B] (when some GEMM of the step reads the dense tile), P[ids] -> Pb.  The replay of the deferred
// user-factor steps and the row gather go to the side stream: the first consumer of Pb joins it (forward_fake).
static int forward_profiles(ganmf_ctx* c, int ids_offset, int B, bool dense_real = true) {
  const Csr& tr = c->csr[GANMF_CSR_TRAIN];
  const int* ids = c->ids + ids_offset;
  cudaStream_t s2;
  RC(aux_fork(c, &s2));
  const Param& P = c->params[c->n_d];
  RC(p_catchup(c, ids, B, s2));            // deferred optimiser steps of exactly these rows
  gather_rows_kernel<<<B, 64, 0, s2>>>(P.w.p, ids, c->Pb.p, c->Pb.ld);
  CU(cudaGetLastError());
  c->launches += 1;
  if (dense_real) {
    CU(csr_gather_dense(tr.indptr, tr.indices, tr.data, ids, B, c->X2.p, c->X2.ld, 0, c->st));
    c->launches += 1;
  }
  return 0;
}
// out[M, N] = A . B^T on the resident-A generator kernel (gen_gemm.cuh); accounted like every tensor-core GEMM
static int gen_gemm(ganmf_ctx* c, const GenGemmCall& g) {
  c->launches += 1;
  cudaEvent_t ev1 = nullptr;
  if (c->profile) {
    while (c->ev_pool.size() < c->ev_used + 2) {
      cudaEvent_t ev;
      CU(cudaEventCreate(&ev));
      c->ev_pool.push_back(ev);
    }
    cudaEvent_t ev0 = c->ev_pool[c->ev_used++];
    ev1 = c->ev_pool[c->ev_used++];
    CU(cudaEventRecord(ev0, c->st));
  }
  cudaError_t e = resident_a_gemm(g, c->st);
  if (e != cudaSuccess) return fail("resident_a_gemm(M=%d N=%d K=%d) -> %s", g.M, g.N, g.K, cudaGetErrorString(e));
  if (c->profile) {
    CU(cudaEventRecord(ev1, c->st));
    c->prof_flops += 2.0 * g.M * g.N * g.K;
    c->prof_launches += 1;
    c->prof_shapes.insert(c->prof_shapes.end(), {g.M, g.N, g.K, 1});
  }
  return 0;
}
// fake profiles F = Pb . V^T -> X2[B:2B]
static int forward_fake(ganmf_ctx* c, int B) {
  RC(aux_join(c));
  const Param& V = c->params[c->n_d + 1];
  if (c->gen_resident_mode != 0 && c->cfg.gemm_path != GANMF_GEMM_SIMT && (c->gen_resident_mode == 1 || B >= 256)) {
    GenGemmCall g;
    g.A = c->Pb.p; g.lda = c->Pb.ld; g.B = V.w.p; g.ldb = V.w.ld;
    g.M = B; g.N = c->W; g.K = c->k;
    g.out = c->X2.row(B); g.ldo = c->X2.ld;
    g.cache = &c->tmaps; g.max_ctas = c->gemm_sm_cap;
    if (resident_a_gemm_ok(g)) return gen_gemm(c, g);                                // G1
  }
  Epilogue ep;
  ep.out = c->X2.row(B); ep.ldo = c->X2.ld;
  return gemm(c, c->Pb.p, c->Pb.ld, 0, V.w.p, V.w.ld, 0, B, c->W, c->k, ep);      // G1
}
static int forward_generator(ganmf_ctx* c, int ids_offset, int B) {
  RC(forward_profiles(c, ids_offset, B));
  return forward_fake(c, B);
}

static int adam_group(ganmf_ctx* c, int first, int count, float alpha, float reg, int slot_param) {
  // one launch per ADAM_MAX_SEG tensors (a 4-layer DisGANMF discriminator has 10)
  for (int f0 = 0; f0 < count; f0 += ADAM_MAX_SEG) {
    const int n = std::min(ADAM_MAX_SEG, count - f0);
    AdamArgs a;
    memset(&a, 0, sizeof a);
    a.nseg = n;
    for (int i = 0; i < n; ++i) {
      Param& p = c->params[first + f0 + i];
      AdamSeg& s = a.seg[i];
      s.theta = p.w.p; s.m = p.m; s.v = p.v; s.g = p.g; s.slot = nullptr; s.ld = p.w.ld;
      s.n4 = p.w.elems() / 4;
      if (first + f0 + i == slot_param) { s.g = c->dPb.p; s.slot = c->slot; }
    }
    a.alpha = alpha; a.reg = reg;
    a.l2_out = &c->sc->l2;
    a.l2_shard_out = &c->sc->l2_shard;
    c->launches++;
    CU(fused_adam(a, c->st));
  }
  return 0;
}

// Codes of the stacked [real ; fake] rows, H2 = X2 . We (+ bias).  On the sparse route (note_train_csr) the real
// rows are a gather-sum over their CSR entries -- exact fp32, 4*E bytes per interaction -- and only the fake rows go
// through the tensor cores: the real half of G2 (2*B*I*E flops, twice per step pair) is gone.
// On the low-rank route (ganmf_ctx::lowrank) the fake rows' codes are Pb . M1 with M1 = V^T . We [k, E]:
// forward_codes_partial leaves M1 (a partial sum over the items of an item-sharded context, like the real codes),
// forward_codes_finish forms Hf = Pb . M1 + be from the complete M1.
static int forward_m1(ganmf_ctx* c) {
  Param* We = &c->params[0];
  const Param& V = c->params[c->n_d + 1];
  Epilogue em;                                                                     // M1 = V^T . We
  em.out = c->M1.p; em.ldo = c->M1.ld;
  return gemm(c, V.w.p, V.w.ld, 1, We->w.p, We->w.ld, 1, c->k, c->E, c->W, em);
}
// the real rows' codes alone (low-rank route: the only rows that are encoded at all)
static int forward_codes_real(ganmf_ctx* c, int ids_offset, int B, const float* bias) {
  Param* We = &c->params[0];
  if (c->sparse_real) {
    const Csr& tr = c->csr[GANMF_CSR_TRAIN];
    CU(csr_encode_rows(tr.indptr, tr.indices, tr.data, c->ids + ids_offset, B, We->w.p, We->w.ld, c->H2.ld, bias,
                       c->H2.p, c->H2.ld, c->st));
    c->launches++;
    return 0;
  }
  Epilogue e2;
  e2.out = c->H2.p; e2.ldo = c->H2.ld; e2.bias = bias;
  return gemm(c, c->X2.p, c->X2.ld, 0, We->w.p, We->w.ld, 1, B, c->E, c->W, e2);
}
static int forward_codes_partial(ganmf_ctx* c, int ids_offset, int B, const float* bias) {
  Param* We = &c->params[0];
  const int fake_rows = c->lowrank ? 0 : B;            // rows of X2 that still go through the dense encode GEMM
  Epilogue e2;                                                                     // G2
  e2.bias = bias;
  if (c->sparse_real) {
    const Csr& tr = c->csr[GANMF_CSR_TRAIN];
    CU(csr_encode_rows(tr.indptr, tr.indices, tr.data, c->ids + ids_offset, B, We->w.p, We->w.ld, c->H2.ld, bias,
                       c->H2.p, c->H2.ld, c->st));
    c->launches++;
    if (fake_rows) {
      e2.out = c->H2.row(B); e2.ldo = c->H2.ld;
      RC(gemm(c, c->X2.row(B), c->X2.ld, 0, We->w.p, We->w.ld, 1, fake_rows, c->E, c->W, e2));
    }
  } else {
    e2.out = c->H2.p; e2.ldo = c->H2.ld;
    RC(gemm(c, c->X2.p, c->X2.ld, 0, We->w.p, We->w.ld, 1, B + fake_rows, c->E, c->W, e2));
  }
  return c->lowrank ? forward_m1(c) : 0;
}
static int forward_codes_finish(ganmf_ctx* c, int B, const float* bias) {
  if (!c->lowrank) return 0;
  RC(aux_join(c));
  Epilogue ef;                                                                     // Hf = Pb . M1 + be
  ef.out = c->H2.row(B); ef.ldo = c->H2.ld; ef.bias = bias;
  return gemm(c, c->Pb.p, c->Pb.ld, 0, c->M1.p, c->M1.ld, 1, B, c->E, c->k, ef);
}
static int forward_codes(ganmf_ctx* c, int ids_offset, int B, const float* bias) {
  RC(forward_codes_partial(c, ids_offset, B, bias));
  return forward_codes_finish(c, B, bias);
}

// Profiles, generator and codes of one step, ordered so that the side stream has company: on the low-rank route the
// real rows' codes (HBM-bound gather-sum or GEMM) are formed before the generator GEMM, which is the first kernel that
// needs the (caught-up) user-factor rows.  dense_real: some later GEMM of the step reads the dense real tile X2[0:B].
static int forward_all(ganmf_ctx* c, int ids_offset, int B, const float* bias, bool dense_real) {
  RC(forward_profiles(c, ids_offset, B, dense_real || !c->sparse_real));
  CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
  if (c->lowrank) {
    RC(forward_codes_real(c, ids_offset, B, bias));
    RC(forward_fake(c, B));
    RC(forward_m1(c));
    return forward_codes_finish(c, B, bias);
  }
  RC(forward_fake(c, B));
  return forward_codes_partial(c, ids_offset, B, bias);
}

// G3 asks for the per-32-row column sums of the residual when the real / fake boundary falls on a row group
static void want_colpart(ganmf_ctx* c, Epilogue& e3, int B) {
  c->colpart_ok = false;
  if (c->colpart_on && c->colpart && (B & 31) == 0 && (2 * B + 31) / 32 <= c->colpart_rows) {
    e3.colpart = c->colpart; e3.ldcp = c->Wp;
  }
}
// dbd = colsum(rs (.) Res2): from G3's partial sums when they exist, else one pass over the residual
static int decoder_bias_grad(ganmf_ctx* c, int B, const float* rs, float* out) {
  if (c->colpart_ok)
    colsum_parts_kernel<<<(c->W + 255) / 256, 256, 0, c->st>>>(c->colpart, 2 * B / 32, B / 32, c->W, c->Wp, rs, out);
  else
    colsum_kernel<<<(c->W + 31) / 32, dim3(32, 8), 0, c->st>>>(c->Res2.p, 2 * B, c->W, c->Res2.ld, rs, B, nullptr, out);
  CU(cudaGetLastError());
  return 0;
}

// ---- GANMF ---------------------------------------------------------------------------------
// phase 1: profiles + generator (independent of the discriminator weights); phase 2: discriminator
// forward on [R ; F]; phase 0: both.  A data-parallel caller runs phase 1 of the NEXT step while the
// all-gather of the freshly updated discriminator weights is still in flight.
static int ganmf_d_forward_impl(ganmf_ctx* c, int ids_offset, int B, int phase = 0) {
  Param *be = &c->params[1], *Wd = &c->params[2], *bd = &c->params[3];
  if (phase == 0) {
    RC(forward_all(c, ids_offset, B, be->w.p, true));                              // G1, G2
  } else {
    if (phase != 2) RC(forward_generator(c, ids_offset, B));
    if (phase == 1) return 0;
    CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
    RC(forward_codes(c, ids_offset, B, be->w.p));                                  // G2
  }
  Epilogue e3;                                                                     // G3
  e3.out = c->Res2.p; e3.ldo = c->Res2.ld; e3.bias = bd->w.p;
  e3.c1 = c->X2.p; e3.ldc1 = c->X2.ld; e3.beta1 = -1.f;
  e3.sumsq2 = c->sc->sumsq; e3.row_split = B;
  want_colpart(c, e3, B);
  return gemm(c, c->H2.p, c->H2.ld, 0, Wd->w.p, Wd->w.ld, 1, 2 * B, c->W, c->E, e3);
}

// phase 1: hinge gate, decoder gradients (dWd, dbd); phase 2: dH, encoder gradients (dWe, dbe);
// phase 0: both.  The cut lets a data-parallel caller start summing the decoder gradients over ranks
// while the encoder half is still being computed.
static int ganmf_d_backward_impl(ganmf_ctx* c, int B, int n_global, float m_hinge, int phase) {
  Param *We = &c->params[0], *be = &c->params[1], *Wd = &c->params[2], *bd = &c->params[3];
  const float* rs = c->sc->row_scale;
  if (phase >= 2) goto second_half;
  {
  const double n_elems = (double)n_global * c->Wg;
  hinge_gate_kernel<<<1, 1, 0, c->st>>>(c->sc, m_hinge, n_elems);
  CU(cudaGetLastError());
  scale_rows_kernel<<<dim3(std::max(1, c->H2.ld / 4 / 128), 2 * B), 128, 0, c->st>>>(
      c->H2.p, c->H2s.p, c->H2.ld / 4, rs, B);
  CU(cudaGetLastError());
  Epilogue e4;                                                                     // G4: dWd
  e4.out = Wd->g; e4.ldo = Wd->w.ld;
  RC(gemm(c, c->H2s.p, c->H2s.ld, 1, c->Res2.p, c->Res2.ld, 1, c->E, c->W, 2 * B, e4));
  RC(decoder_bias_grad(c, B, rs, bd->g));                                          // dbd
  // dbe = colsum(dH2) = dbd . Wd^T, evaluated in fp32 from the LOCAL fp32 column sums (no tensor-core
  // rounding); it belongs to phase 1 because a data-parallel caller sums dbd in place right after it
  rowdot_kernel<<<c->E, 256, 0, c->st>>>(Wd->w.p, c->E, c->W, Wd->w.ld, bd->g, be->g);
  CU(cudaGetLastError());
  c->launches += 4;
  }
  if (phase == 1) return 0;
second_half:
  if (phase != 4) {
    Epilogue e5;                                                                   // G5: dH2 (reads Wd)
    e5.out = c->dH2.p; e5.ldo = c->dH2.ld; e5.row_scale2 = rs; e5.row_split = B;
    RC(gemm(c, c->Res2.p, c->Res2.ld, 0, Wd->w.p, Wd->w.ld, 0, 2 * B, c->E, c->W, e5));
  }
  if (phase != 3) {
    Epilogue e6;                                                                   // G6: dWe
    e6.out = We->g; e6.ldo = We->w.ld;
    RC(gemm(c, c->X2.p, c->X2.ld, 1, c->dH2.p, c->dH2.ld, 1, c->W, c->E, 2 * B, e6));
  }
  return 0;
}

// Single-GPU D update with the optimiser fused into the weight-gradient GEMMs: dH is formed first (it
// needs the OLD decoder weights), then the dWd / dWe GEMM epilogues run ApplyAdam on Wd / We in place
// (no gradient round trip through HBM, the parameter traffic hides under the MMAs); the two biases
// take a tiny Adam launch with the same step size.
static int ganmf_d_backward_apply_fused(ganmf_ctx* c, int B, int n_global, float m_hinge, float lr, float reg,
                                        int loss_slot) {
  if (loss_slot < 0 || loss_slot >= c->losses_cap) return fail("loss slot out of range");
  Param *We = &c->params[0], *be = &c->params[1], *Wd = &c->params[2], *bd = &c->params[3];
  const float* rs = c->sc->row_scale;
  hinge_gate_kernel<<<1, 1, 0, c->st>>>(c->sc, m_hinge, (double)n_global * c->Wg);
  CU(cudaGetLastError());
  scale_rows_kernel<<<dim3(std::max(1, c->H2.ld / 4 / 128), 2 * B), 128, 0, c->st>>>(
      c->H2.p, c->H2s.p, c->H2.ld / 4, rs, B);
  CU(cudaGetLastError());
  RC(decoder_bias_grad(c, B, rs, bd->g));                                          // dbd
  // dbe = dbd . Wd^T.  (Running these row dots on the side stream under G5 -- both stream the old Wd -- was measured:
  // G5 slows down by 0.23 ms to hide a 0.18 ms kernel, profiles/r02b_ab_aux_stream.txt.)
  rowdot_kernel<<<c->E, 256, 0, c->st>>>(Wd->w.p, c->E, c->W, Wd->w.ld, bd->g, be->g);
  CU(cudaGetLastError());
  c->launches += 4;
  Epilogue e5;                                                                     // G5: dH2 (old Wd)
  e5.out = c->dH2.p; e5.ldo = c->dH2.ld; e5.row_scale2 = rs; e5.row_split = B;
  RC(gemm(c, c->Res2.p, c->Res2.ld, 0, Wd->w.p, Wd->w.ld, 0, 2 * B, c->E, c->W, e5));
  const float alpha = adam_alpha(c, 0, lr);
  Epilogue e4;                                                                     // G4: dWd -> Adam(Wd)
  e4.out = Wd->w.p; e4.ldo = Wd->w.ld;
  e4.adam_m = Wd->m; e4.adam_v = Wd->v; e4.adam_alpha = alpha; e4.adam_reg = reg; e4.adam_l2 = &c->sc->l2;
  RC(gemm(c, c->H2s.p, c->H2s.ld, 1, c->Res2.p, c->Res2.ld, 1, c->E, c->W, 2 * B, e4));
  Epilogue e6;                                                                     // G6: dWe -> Adam(We)
  e6.out = We->w.p; e6.ldo = We->w.ld;
  e6.adam_m = We->m; e6.adam_v = We->v; e6.adam_alpha = alpha; e6.adam_reg = reg; e6.adam_l2 = &c->sc->l2;
  RC(gemm(c, c->X2.p, c->X2.ld, 1, c->dH2.p, c->dH2.ld, 1, c->W, c->E, 2 * B, e6));
  AdamArgs a;                                                                      // biases
  memset(&a, 0, sizeof a);
  a.nseg = 2;
  Param* bs[2] = {be, bd};
  for (int i = 0; i < 2; ++i) {
    a.seg[i].theta = bs[i]->w.p; a.seg[i].m = bs[i]->m; a.seg[i].v = bs[i]->v; a.seg[i].g = bs[i]->g;
    a.seg[i].ld = bs[i]->w.ld; a.seg[i].n4 = bs[i]->w.elems() / 4;
  }
  a.alpha = alpha; a.reg = reg; a.l2_out = &c->sc->l2; a.l2_shard_out = &c->sc->l2_shard;
  CU(fused_adam(a, c->st));
  finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot);
  CU(cudaGetLastError());
  c->launches += 2;
  return 0;
}

static int d_apply_impl(ganmf_ctx* c, float lr, float reg, int loss_slot) {
  if (loss_slot < 0 || loss_slot >= c->losses_cap) return fail("loss slot out of range");
  if (c->cfg.kind == GANMF_KIND_DISGANMF) {
    dis_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, 0, 0.f, (double)c->last_n_global, 1.0);
    CU(cudaGetLastError());
    c->launches++;
  }
  RC(adam_group(c, 0, c->n_d, adam_alpha(c, 0, lr), reg, -1));
  finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}

// Low-rank generator backward (ganmf_ctx::lowrank), from Res_f = Res2[B:2B], dHf = dH2[B:2B], Pb, M1:
//   dPb = [dHf . M1^T] - c1 * (Res_f . V)      (the bracket is complete on every rank of an item-sharded group:
//                                               `with_codes_term` is true on one of them only)
//   dV  = We . (dHf^T . Pb) - c1 * (Res_f^T . Pb)   -> V.g
// (the *_a halves need no code gradient: an item-sharded caller runs them while dHf is being summed over the ranks)
static int lowrank_dpb_a(ganmf_ctx* c, int B, float c1) {
  const Param& V = c->params[c->n_d + 1];
  Epilogue ea;                                                                     // -c1 * Res_f . V
  ea.out = c->dPb.p; ea.ldo = c->dPb.ld; ea.alpha = -c1;
  return gemm(c, c->Res2.row(B), c->Res2.ld, 0, V.w.p, V.w.ld, 1, B, c->k, c->W, ea);
}
static int lowrank_dpb_b(ganmf_ctx* c, int B) {
  Epilogue eb;                                                                     // + dHf . M1^T
  eb.out = c->dPb.p; eb.ldo = c->dPb.ld;
  eb.c1 = c->dPb.p; eb.ldc1 = c->dPb.ld; eb.beta1 = 1.f;
  return gemm(c, c->dH2.row(B), c->dH2.ld, 0, c->M1.p, c->M1.ld, 0, B, c->k, c->E, eb);
}
static int lowrank_dpb(ganmf_ctx* c, int B, float c1, bool with_codes_term) {
  RC(lowrank_dpb_a(c, B, c1));
  return with_codes_term ? lowrank_dpb_b(c, B) : 0;
}
static int lowrank_dv_a(ganmf_ctx* c, int B, float c1) {
  const Param& V = c->params[c->n_d + 1];
  Epilogue ea;                                                                     // -c1 * Res_f^T . Pb
  ea.out = V.g; ea.ldo = V.w.ld; ea.alpha = -c1;
  return gemm(c, c->Res2.row(B), c->Res2.ld, 1, c->Pb.p, c->Pb.ld, 1, c->W, c->k, B, ea);
}
static int lowrank_dv_b(ganmf_ctx* c, int B) {
  Param* We = &c->params[0];
  const Param& V = c->params[c->n_d + 1];
  Epilogue et;                                                                     // T1t = Pb^T . dHf  [k, E]
  et.out = c->T1t.p; et.ldo = c->T1t.ld;
  RC(gemm(c, c->Pb.p, c->Pb.ld, 1, c->dH2.row(B), c->dH2.ld, 1, c->k, c->E, B, et));
  Epilogue eb;                                                                     // + We . T1t^T
  eb.out = V.g; eb.ldo = V.w.ld;
  eb.c1 = V.g; eb.ldc1 = V.w.ld; eb.beta1 = 1.f;
  return gemm(c, We->w.p, We->w.ld, 0, c->T1t.p, c->T1t.ld, 0, c->W, c->k, c->E, eb);
}
static int lowrank_dv(ganmf_ctx* c, int B, float c1) {
  RC(lowrank_dv_a(c, B, c1));
  return lowrank_dv_b(c, B);
}

static int ganmf_g_fb_impl(ganmf_ctx* c, int ids_offset, int B, int n_global, float alpha,
                           bool fuse_adam = false, float adam_step = 0.f, float adam_reg = 0.f, int part = 0) {
  if (part == 2 && c->lowrank) {
    const double Ng = (double)n_global * c->Wg;
    return lowrank_dpb(c, B, (float)((1.0 - alpha) * 2.0 / Ng), true);
  }
  if (part == 2) {                                                                 // dPb only (after part 1)
    Param& V2 = c->params[c->n_d + 1];
    Epilogue e9;
    e9.out = c->dPb.p; e9.ldo = c->dPb.ld;
    return gemm(c, c->dF.p, c->dF.ld, 0, V2.w.p, V2.w.ld, 1, B, c->k, c->W, e9);
  }
  Param *We = &c->params[0], *be = &c->params[1], *Wd = &c->params[2], *bd = &c->params[3];
  Param& V = c->params[c->n_d + 1];
  RC(forward_all(c, ids_offset, B, be->w.p, false));                               // G1, G2 (real + fake codes)
  Epilogue e3;                                                                     // G3': fake residual
  e3.out = c->Res2.row(B); e3.ldo = c->Res2.ld; e3.bias = bd->w.p;
  e3.c1 = c->X2.row(B); e3.ldc1 = c->X2.ld; e3.beta1 = -1.f;
  e3.sumsq2 = c->sc->sumsq;
  RC(gemm(c, c->H2.row(B), c->H2.ld, 0, Wd->w.p, Wd->w.ld, 1, B, c->W, c->E, e3));
  sqdiff_kernel<<<std::min(B, 296), 256, 0, c->st>>>(c->H2.p, c->H2.row(B), B, c->E, c->H2.ld, &c->sc->fm);
  CU(cudaGetLastError());
  const double N = (double)n_global * c->Wg, M = (double)n_global * c->E;
  const float c1 = (float)((1.0 - alpha) * 2.0 / N), c2 = (float)(alpha * 2.0 / M);
  Epilogue e5;                                                                     // G5': dHf
  e5.out = c->dH2.row(B); e5.ldo = c->dH2.ld; e5.alpha = c1;
  e5.c1 = c->H2.row(B); e5.ldc1 = c->H2.ld; e5.beta1 = c2;
  e5.c2 = c->H2.p; e5.ldc2 = c->H2.ld; e5.beta2 = -c2;
  RC(gemm(c, c->Res2.row(B), c->Res2.ld, 0, Wd->w.p, Wd->w.ld, 0, B, c->E, c->W, e5));
  if (c->lowrank) {
    if (part != 1) RC(lowrank_dpb(c, B, c1, true));
    if (part == 0 || part == 1) RC(lowrank_dv(c, B, c1));
    return 0;
  }
  Epilogue e7;                                                                     // G7: dF
  e7.out = c->dF.p; e7.ldo = c->dF.ld;
  e7.c1 = c->Res2.row(B); e7.ldc1 = c->Res2.ld; e7.beta1 = -c1;
  RC(gemm(c, c->dH2.row(B), c->dH2.ld, 0, We->w.p, We->w.ld, 0, B, c->W, c->E, e7));
  // G9 (dPb, reads the old V) and G8 (dV): independent unless the optimiser is fused into G8's epilogue, which
  // rewrites V.  part 1 ends with dV so a data-parallel caller can sum it over ranks while part 2 (dPb) runs.
  auto g9 = [&]() -> int {
    Epilogue e9;                                                                   // G9: dPb
    e9.out = c->dPb.p; e9.ldo = c->dPb.ld;
    return gemm(c, c->dF.p, c->dF.ld, 0, V.w.p, V.w.ld, 1, B, c->k, c->W, e9);
  };
  if (fuse_adam) RC(g9());
  Epilogue e8;                                                                     // G8: dV [-> Adam(V)]
  if (fuse_adam) {
    e8.out = V.w.p; e8.ldo = V.w.ld;
    e8.adam_m = V.m; e8.adam_v = V.v; e8.adam_alpha = adam_step; e8.adam_reg = adam_reg; e8.adam_l2 = &c->sc->l2;
  } else {
    e8.out = V.g; e8.ldo = V.w.ld;
  }
  RC(gemm(c, c->dF.p, c->dF.ld, 1, c->Pb.p, c->Pb.ld, 1, c->W, c->k, B, e8));
  c->launches += 1;
  if (!fuse_adam && part != 1) RC(g9());
  return 0;
}

static int g_apply_impl(ganmf_ctx* c, int ids_offset, int B, int n_global, float lr, float reg, float alpha,
                        int loss_slot, bool v_done = false, float adam_step = 0.f) {
  if (loss_slot < 0 || loss_slot >= c->losses_cap) return fail("loss slot out of range");
  if (c->cfg.kind == GANMF_KIND_GANMF)
    gloss_kernel<<<1, 1, 0, c->st>>>(c->sc, alpha, c->tp_rank == 0 ? alpha : 0.f, (double)n_global * c->Wg,
                                     (double)n_global * c->E);
  else
    dis_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, 1, alpha, (double)n_global, (double)n_global * c->E);
  CU(cudaGetLastError());
  const int* ids = c->ids + ids_offset;
  if (c->lazy_p && reg == 0.f) {
    // user factors: Adam on the minibatch rows only, every other row defers its zero-gradient step
    // (kernels.cuh K6b); item factors: the usual dense update
    const float a = v_done ? adam_step : adam_alpha(c, 1, lr);
    if (c->g_T - c->log_base >= c->log_cap) RC(p_flush(c));
    Param& P = c->params[c->n_d];
    p_batch_adam_kernel<<<B, 64, 0, c->st>>>(P.w.p, P.m, P.v, P.w.ld, ids, c->dPb.p, c->dPb.ld, a, c->p_last,
                                             c->alpha_log, c->g_T, c->log_base);
    CU(cudaGetLastError());
    c->launches++;
    c->g_T++;
    c->p_stale = true;
    if (!v_done) RC(adam_group(c, c->n_d + 1, 1, a, reg, -1));
  } else {
    RC(p_flush(c));
    set_slots_kernel<<<(B + 127) / 128, 128, 0, c->st>>>(c->slot, ids, B, 0);
    CU(cudaGetLastError());
    if (v_done) RC(adam_group(c, c->n_d, 1, adam_step, reg, c->n_d));          // user factors only
    else RC(adam_group(c, c->n_d, 2, adam_alpha(c, 1, lr), reg, c->n_d));
    set_slots_kernel<<<(B + 127) / 128, 128, 0, c->st>>>(c->slot, ids, B, 1);
    CU(cudaGetLastError());
    c->g_T++;
    c->log_base = c->g_T;
    fill_int_kernel<<<(c->cfg.n_rows + 255) / 256, 256, 0, c->st>>>(c->p_last, c->g_T, (size_t)c->cfg.n_rows);
    CU(cudaGetLastError());
    c->launches += 4;
  }
  if (n_global == B && c->tp_world == 1) {       // under data / item parallelism the caller sums the l2 terms over ranks first
    finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot);
    CU(cudaGetLastError());
    c->launches++;
  }
  return 0;
}

// ---- DisGANMF ------------------------------------------------------------------------------
// discriminator on [float(id) | profile] for the stacked [real ; fake] rows (DisGANMF.py:57-65)
static int dis_forward(ganmf_ctx* c, int ids_offset, int B) {
  const int L = c->cfg.d_layers, H = c->cfg.d_nodes, act = c->cfg.d_act;
  const int* ids = c->ids + ids_offset;
  ids_to_float_kernel<<<(B + 127) / 128, 128, 0, c->st>>>(ids, c->idf, B, c->cfg.row_id_offset);
  ids_to_float_kernel<<<(B + 127) / 128, 128, 0, c->st>>>(ids, c->idf + B, B, c->cfg.row_id_offset);
  CU(cudaGetLastError());
  c->launches += 2;
  for (int l = 0; l < L; ++l) {
    Param& Wl = c->params[2 * l];
    Param& bl = c->params[2 * l + 1];
    Epilogue e;
    e.out = c->hs[l].p; e.ldo = c->hs[l].ld; e.bias = bl.w.p; e.act = act;
    if (l == 0) {
      // x . W0 = profile . W0[1:, :] + id_f (x) W0[0, :]
      e.r1_row = c->idf; e.r1_col = Wl.w.p;
      RC(gemm(c, c->X2.p, c->X2.ld, 0, Wl.w.row(1), Wl.w.ld, 1, 2 * B, H, c->W, e));
    } else {
      RC(gemm(c, c->hs[l - 1].p, c->hs[l - 1].ld, 0, Wl.w.p, Wl.w.ld, 1, 2 * B, H, H, e));
    }
  }
  Param& Wo = c->params[2 * L];
  Param& bo = c->params[2 * L + 1];
  Epilogue eo;
  eo.out = c->out2; eo.ldo = 1; eo.bias = bo.w.p;
  return gemm(c, c->hs[L - 1].p, c->hs[L - 1].ld, 0, Wo.w.p, Wo.w.ld, 1, 2 * B, 1, H, eo, GANMF_GEMM_SIMT);
}

// back-propagate dout2 (+ optional extra gradient already in dhtmp) through the MLP.
// want_param_grads: accumulate into the D gradient buffers; want_dx: dF (fake half only).
static int dis_backward(ganmf_ctx* c, int B, bool want_param_grads, bool extra_in_dhtmp, bool want_dx) {
  const int L = c->cfg.d_layers, H = c->cfg.d_nodes, act = c->cfg.d_act;
  Param& Wo = c->params[2 * L];
  Param& bo = c->params[2 * L + 1];
  const int R0 = want_param_grads ? 0 : B;          // G step only needs the fake rows
  const int R = want_param_grads ? 2 * B : B;
  if (want_param_grads) {
    Epilogue e;                                       // dWo = feat^T . dout
    e.out = Wo.g; e.ldo = Wo.w.ld;
    RC(gemm(c, c->hs[L - 1].p, c->hs[L - 1].ld, 1, c->dout2, 1, 1, H, 1, 2 * B, e, GANMF_GEMM_SIMT));
    colsum_kernel<<<1, dim3(32, 8), 0, c->st>>>(c->dout2, 2 * B, 1, 1, nullptr, 0x7fffffff, nullptr, bo.g);
    CU(cudaGetLastError());
    c->launches++;
  }
  // dh = dout . wo^T (+ extra)
  Epilogue eh;
  eh.out = c->dhtmp.row(R0); eh.ldo = c->dhtmp.ld;
  if (extra_in_dhtmp) { eh.c1 = c->dhtmp.row(R0); eh.ldc1 = c->dhtmp.ld; eh.beta1 = 1.f; }
  RC(gemm(c, c->dout2 + R0, 1, 0, Wo.w.p, Wo.w.ld, 0, R, H, 1, eh, GANMF_GEMM_SIMT));
  for (int l = L - 1; l >= 0; --l) {
    Param& Wl = c->params[2 * l];
    Param& bl = c->params[2 * l + 1];
    act_bwd_kernel<<<dim3((H + 127) / 128, R), 128, 0, c->st>>>(c->dhtmp.row(R0), c->hs[l].row(R0),
                                                              c->dzs[l].row(R0), R, H, c->dhtmp.ld, act);
    CU(cudaGetLastError());
    c->launches++;
    if (want_param_grads) {
      if (l == 0) {
        // dW0[1:, :] = profile^T . dz ; dW0[0, :] = sum_m id_f[m] dz[m, :]
        Epilogue e;
        e.out = Wl.g + Wl.w.ld; e.ldo = Wl.w.ld;
        RC(gemm(c, c->X2.p, c->X2.ld, 1, c->dzs[0].p, c->dzs[0].ld, 1, c->W, H, 2 * B, e));
        colsum_kernel<<<(H + 31) / 32, dim3(32, 8), 0, c->st>>>(c->dzs[0].p, 2 * B, H, c->dzs[0].ld, nullptr,
                                                               0x7fffffff, c->idf, Wl.g);
      } else {
        Epilogue e;
        e.out = Wl.g; e.ldo = Wl.w.ld;
        RC(gemm(c, c->hs[l - 1].p, c->hs[l - 1].ld, 1, c->dzs[l].p, c->dzs[l].ld, 1, H, H, 2 * B, e));
      }
      colsum_kernel<<<(H + 31) / 32, dim3(32, 8), 0, c->st>>>(c->dzs[l].p, 2 * B, H, c->dzs[l].ld, nullptr,
                                                             0x7fffffff, nullptr, bl.g);
      CU(cudaGetLastError());
      c->launches += 2;
    }
    if (l > 0) {
      Epilogue e;                                     // dh_{l-1} = dz_l . W_l^T
      e.out = c->dhtmp.row(R0); e.ldo = c->dhtmp.ld;
      RC(gemm(c, c->dzs[l].row(R0), c->dzs[l].ld, 0, Wl.w.p, Wl.w.ld, 0, R, H, H, e));
    } else if (want_dx) {
      Epilogue e;                                     // dF = dz_0[fake] . W0[1:, :]^T
      e.out = c->dF.p; e.ldo = c->dF.ld;
      RC(gemm(c, c->dzs[0].row(B), c->dzs[0].ld, 0, Wl.w.row(1), Wl.w.ld, 0, B, c->W, H, e));
    }
  }
  return 0;
}

static int dis_d_forward_impl(ganmf_ctx* c, int ids_offset, int B) {
  RC(forward_generator(c, ids_offset, B));
  CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
  return dis_forward(c, ids_offset, B);
}
// (the BCE sums are formed here; the reported loss is assembled in d_apply_impl, after a data-parallel caller
//  has summed the step scalars over ranks -- no gradient depends on them)
static int dis_d_backward_impl(ganmf_ctx* c, int B, int n_global) {
  bce_kernel<<<std::max(1, std::min(64, (2 * B + 255) / 256)), 256, 0, c->st>>>(c->out2, c->dout2, B, 0, c->sc,
                                                                              (float)n_global);
  CU(cudaGetLastError());
  c->last_n_global = n_global;
  c->launches += 1;
  return dis_backward(c, B, true, false, false);
}
static int dis_g_fb_impl(ganmf_ctx* c, int ids_offset, int B, int n_global, float alpha) {
  const int L = c->cfg.d_layers, H = c->cfg.d_nodes;
  RC(forward_generator(c, ids_offset, B));
  CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
  RC(dis_forward(c, ids_offset, B));
  bce_kernel<<<std::max(1, std::min(64, (2 * B + 255) / 256)), 256, 0, c->st>>>(c->out2, c->dout2, B, 1, c->sc,
                                                                              (float)n_global);
  const Mat& ft = c->hs[L - 1];
  sqdiff_kernel<<<std::min(B, 296), 256, 0, c->st>>>(ft.p, ft.row(B), B, H, ft.ld, &c->sc->fm);
  CU(cudaGetLastError());
  // extra gradient on the fake features: alpha * 2/(B*H) * (feat_f - feat_r) -> dhtmp[fake rows]
  const float c2 = (float)(alpha * 2.0 / ((double)n_global * H));
  axpby_kernel<<<dim3((H + 127) / 128, B), 128, 0, c->st>>>(ft.row(B), ft.p, c->dhtmp.row(B), B, H, ft.ld, c2,
                                                           -c2);
  CU(cudaGetLastError());
  c->launches += 2;
  RC(dis_backward(c, B, false, true, true));
  Param& V = c->params[c->n_d + 1];
  Epilogue e8;
  e8.out = V.g; e8.ldo = V.w.ld;
  RC(gemm(c, c->dF.p, c->dF.ld, 1, c->Pb.p, c->Pb.ld, 1, c->W, c->k, B, e8));
  Epilogue e9;
  e9.out = c->dPb.p; e9.ldo = c->dPb.ld;
  return gemm(c, c->dF.p, c->dF.ld, 0, V.w.p, V.w.ld, 1, B, c->k, c->W, e9);
}

// ---- item-sharded GANMF ------------------------------------------------------------
// Every rank runs the WHOLE minibatch on its item slice: X2 = [R ; F][:, slice], We[slice, :], Wd[:, slice],
// V[slice, :].  Contractions over items (codes H = X.We, dH = Res.Wd^T, dPb = dF.V, dbe = Wd.dbd) are partial
// sums the caller all-reduces; everything indexed by items (residuals, dWe, dWd, dbd, dV and their Adam
// updates) is local, so no weight or weight-gradient ever crosses NVLink -- only [2B, E] / [B, k] activations.
static int tp_check(ganmf_ctx* c, int ids_offset, int B) {
  if (!c) return fail("null ctx");
  if (c->tp_world <= 1) return fail("not an item-sharded context (config.tp_world <= 1)");
  return check_batch(c, ids_offset, B, true);
}
// Low-rank route, phase 1 in two parts so the all-reduce of the real rows' codes travels while the generator GEMM and
// V^T . We are computed: part 6 = profiles + the real rows' partial codes, part 7 = F = Pb . V^T + the partial M1.
static int tp_forward_split(ganmf_ctx* c, int phase, int ids_offset, int B, bool dense_real) {
  if (!c->lowrank) return fail("phases 6 / 7 belong to the low-rank route (ganmf_step_routes)");
  if (phase == 6) {
    RC(forward_profiles(c, ids_offset, B, dense_real || !c->sparse_real));
    CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
    return forward_codes_real(c, ids_offset, B, c->tp_rank == 0 ? c->params[1].w.p : nullptr);
  }
  RC(forward_fake(c, B));
  return forward_m1(c);
}
// profiles + generator + partial codes (the bias joins the sum once: on rank 0)
static int tp_forward_codes(ganmf_ctx* c, int ids_offset, int B) {
  RC(forward_generator(c, ids_offset, B));
  Param* be = &c->params[1];
  CU(cudaMemsetAsync(c->sc, 0, sizeof(StepScalars), c->st));
  // G2 (partial over items; on the low-rank route the fake rows' codes are formed from the summed M1 in phase 2)
  return forward_codes_partial(c, ids_offset, B, c->tp_rank == 0 ? be->w.p : nullptr);
}

int ganmf_tp_d_phase(ganmf_ctx* c, int phase, int ids_offset, int B, float lr, float reg, float m_hinge,
                     int loss_slot) {
  RC(tp_check(c, ids_offset, B));
  if (loss_slot < 0 || loss_slot >= c->losses_cap) return fail("loss slot out of range");
  Param *We = &c->params[0], *be = &c->params[1], *Wd = &c->params[2], *bd = &c->params[3];
  const float* rs = c->sc->row_scale;
  switch (phase) {
    case 1:
      return tp_forward_codes(c, ids_offset, B);
    case 6:
    case 7:
      return tp_forward_split(c, phase, ids_offset, B, true);
    case 2: {                                                                      // G3 on the summed codes
      RC(forward_codes_finish(c, B, be->w.p));
      Epilogue e3;
      e3.out = c->Res2.p; e3.ldo = c->Res2.ld; e3.bias = bd->w.p;
      e3.c1 = c->X2.p; e3.ldc1 = c->X2.ld; e3.beta1 = -1.f;
      e3.sumsq2 = c->sc->sumsq; e3.row_split = B;
      want_colpart(c, e3, B);
      return gemm(c, c->H2.p, c->H2.ld, 0, Wd->w.p, Wd->w.ld, 1, 2 * B, c->W, c->E, e3);
    }
    case 3: {                                                                      // gate (global sums), dbd, partial dH | dbe
      hinge_gate_kernel<<<1, 1, 0, c->st>>>(c->sc, m_hinge, (double)B * c->Wg);
      CU(cudaGetLastError());
      scale_rows_kernel<<<dim3(std::max(1, c->H2.ld / 4 / 128), 2 * B), 128, 0, c->st>>>(
          c->H2.p, c->H2s.p, c->H2.ld / 4, rs, B);
      CU(cudaGetLastError());
      RC(decoder_bias_grad(c, B, rs, bd->g));
      rowdot_kernel<<<c->E, 256, 0, c->st>>>(Wd->w.p, c->E, c->W, Wd->w.ld, bd->g, c->dH2.row(2 * B));
      CU(cudaGetLastError());
      c->launches += 4;
      Epilogue e5;                                                                 // G5 (partial over items)
      e5.out = c->dH2.p; e5.ldo = c->dH2.ld; e5.row_scale2 = rs; e5.row_split = B;
      return gemm(c, c->Res2.p, c->Res2.ld, 0, Wd->w.p, Wd->w.ld, 0, 2 * B, c->E, c->W, e5);
    }
    case 4: {                          // dWd needs no summed quantity: it runs while the code gradients are all-reduced
      const float alpha = c->tp_d_alpha = adam_alpha(c, 0, lr);
      Epilogue e4;                                                                 // G4: dWd[:, slice] -> Adam
      e4.out = Wd->w.p; e4.ldo = Wd->w.ld;
      e4.adam_m = Wd->m; e4.adam_v = Wd->v; e4.adam_alpha = alpha; e4.adam_reg = reg; e4.adam_l2 = &c->sc->l2;
      return gemm(c, c->H2s.p, c->H2s.ld, 1, c->Res2.p, c->Res2.ld, 1, c->E, c->W, 2 * B, e4);
    }
    case 5: {                                                                      // summed dH2 | dbe -> dWe, biases
      const float alpha = c->tp_d_alpha;
      CU(cudaMemcpyAsync(be->g, c->dH2.row(2 * B), (size_t)be->w.ld * 4, cudaMemcpyDeviceToDevice, c->st));
      if (c->lowrank_dwe) {
        // dWe[slice] = R^T.dH_r + V[slice].(Pb^T.dH_f): K = B + k instead of 2B; plain gradient + one Adam pass
        const Param& V = c->params[c->n_d + 1];
        Epilogue et;                                                               // T = Pb^T . dH_f  [k, E]
        et.out = c->T1t.p; et.ldo = c->T1t.ld;
        RC(gemm(c, c->Pb.p, c->Pb.ld, 1, c->dH2.row(B), c->dH2.ld, 1, c->k, c->E, B, et));
        Epilogue ea;                                                               // R^T . dH_r
        ea.out = We->g; ea.ldo = We->w.ld;
        RC(gemm(c, c->X2.p, c->X2.ld, 1, c->dH2.p, c->dH2.ld, 1, c->W, c->E, B, ea));
        Epilogue eb;                                                               // + V . T
        eb.out = We->g; eb.ldo = We->w.ld;
        eb.c1 = We->g; eb.ldc1 = We->w.ld; eb.beta1 = 1.f;
        RC(gemm(c, V.w.p, V.w.ld, 0, c->T1t.p, c->T1t.ld, 1, c->W, c->E, c->k, eb));
        AdamArgs aw;
        memset(&aw, 0, sizeof aw);
        aw.nseg = 1;
        aw.seg[0].theta = We->w.p; aw.seg[0].m = We->m; aw.seg[0].v = We->v; aw.seg[0].g = We->g;
        aw.seg[0].ld = We->w.ld; aw.seg[0].n4 = We->w.elems() / 4;
        aw.alpha = alpha; aw.reg = reg; aw.l2_out = &c->sc->l2;
        CU(fused_adam(aw, c->st));
        c->launches++;
      } else {
      Epilogue e6;                                                                 // G6: dWe[slice, :] -> Adam
      e6.out = We->w.p; e6.ldo = We->w.ld;
      e6.adam_m = We->m; e6.adam_v = We->v; e6.adam_alpha = alpha; e6.adam_reg = reg; e6.adam_l2 = &c->sc->l2;
      RC(gemm(c, c->X2.p, c->X2.ld, 1, c->dH2.p, c->dH2.ld, 1, c->W, c->E, 2 * B, e6));
      }
      // biases: bd is a slice, be is replicated (every rank applies the same update; its l2 counts once)
      Param* bs[2] = {be, bd};
      for (int i = 0; i < 2; ++i) {
        AdamArgs a;
        memset(&a, 0, sizeof a);
        a.nseg = 1;
        a.seg[0].theta = bs[i]->w.p; a.seg[0].m = bs[i]->m; a.seg[0].v = bs[i]->v; a.seg[0].g = bs[i]->g;
        a.seg[0].ld = bs[i]->w.ld; a.seg[0].n4 = bs[i]->w.elems() / 4;
        a.alpha = alpha; a.reg = reg;
        a.l2_out = (i == 1 || c->tp_rank == 0) ? &c->sc->l2 : nullptr;
        CU(fused_adam(a, c->st));
      }
      // per-rank partial loss: the hinge / reconstruction part (identical on every rank) is logged by rank 0
      finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot, c->tp_rank == 0 ? 1.f : 0.f, 0.f);
      CU(cudaGetLastError());
      c->launches += 3;
      return 0;
    }
    default:
      return fail("ganmf_tp_d_phase: phase must be 1..7");
  }
}

int ganmf_tp_g_phase(ganmf_ctx* c, int phase, int ids_offset, int B, float lr, float reg, float alpha,
                     int loss_slot) {
  RC(tp_check(c, ids_offset, B));
  if (loss_slot < 0 || loss_slot >= c->losses_cap) return fail("loss slot out of range");
  Param *We = &c->params[0], *Wd = &c->params[2], *bd = &c->params[3];
  Param& V = c->params[c->n_d + 1];
  const double N = (double)B * c->Wg, M = (double)B * c->E;
  const float c1 = (float)((1.0 - alpha) * 2.0 / N), c2 = (float)(alpha * 2.0 / M);
  switch (phase) {
    case 1:
      return tp_forward_codes(c, ids_offset, B);
    case 6:
    case 7:
      return tp_forward_split(c, phase, ids_offset, B, false);   // (no GEMM of the G step reads the dense real tile)
    case 8:                            // low-rank route: the halves of dV and dPb that need no summed code gradient
      if (!c->lowrank) return fail("phases 8..10 belong to the low-rank route (ganmf_step_routes)");
      RC(lowrank_dv_a(c, B, c1));
      return lowrank_dpb_a(c, B, c1);
    case 9:                            // ... the code-gradient term of dPb joins the sum once
      if (!c->lowrank) return fail("phases 8..10 belong to the low-rank route (ganmf_step_routes)");
      return c->tp_rank == 0 ? lowrank_dpb_b(c, B) : 0;
    case 10:                           // ... and the code-gradient term of dV (runs while dPb is summed)
      if (!c->lowrank) return fail("phases 8..10 belong to the low-rank route (ganmf_step_routes)");
      return lowrank_dv_b(c, B);
    case 2: {
      RC(forward_codes_finish(c, B, c->params[1].w.p));
      Epilogue e3;                                                                 // G3': fake residual (slice)
      e3.out = c->Res2.row(B); e3.ldo = c->Res2.ld; e3.bias = bd->w.p;
      e3.c1 = c->X2.row(B); e3.ldc1 = c->X2.ld; e3.beta1 = -1.f;
      e3.sumsq2 = c->sc->sumsq;
      RC(gemm(c, c->H2.row(B), c->H2.ld, 0, Wd->w.p, Wd->w.ld, 1, B, c->W, c->E, e3));
      // feature matching on the summed codes: the same on every rank, NOT summed again
      sqdiff_kernel<<<std::min(B, 296), 256, 0, c->st>>>(c->H2.p, c->H2.row(B), B, c->E, c->H2.ld, &c->sc->fm);
      CU(cudaGetLastError());
      c->launches += 1;
      Epilogue e5;                                                                 // G5': dHf (partial; the
      e5.out = c->dH2.row(B); e5.ldo = c->dH2.ld; e5.alpha = c1;                  //  feature-matching term joins once)
      if (c->tp_rank == 0) {
        e5.c1 = c->H2.row(B); e5.ldc1 = c->H2.ld; e5.beta1 = c2;
        e5.c2 = c->H2.p; e5.ldc2 = c->H2.ld; e5.beta2 = -c2;
      }
      return gemm(c, c->Res2.row(B), c->Res2.ld, 0, Wd->w.p, Wd->w.ld, 0, B, c->E, c->W, e5);
    }
    case 3: {
      if (c->lowrank) return lowrank_dpb(c, B, c1, c->tp_rank == 0);             // dPb (partial over items)
      Epilogue e7;                                                                 // G7: dF (slice)
      e7.out = c->dF.p; e7.ldo = c->dF.ld;
      e7.c1 = c->Res2.row(B); e7.ldc1 = c->Res2.ld; e7.beta1 = -c1;
      RC(gemm(c, c->dH2.row(B), c->dH2.ld, 0, We->w.p, We->w.ld, 0, B, c->W, c->E, e7));
      Epilogue e9;                                                                 // G9: dPb (partial over items)
      e9.out = c->dPb.p; e9.ldo = c->dPb.ld;
      return gemm(c, c->dF.p, c->dF.ld, 0, V.w.p, V.w.ld, 1, B, c->k, c->W, e9);
    }
    case 4: {                                                                      // (runs while dPb is summed)
      if (c->lowrank) return lowrank_dv(c, B, c1);
      Epilogue e8;                                                                 // G8: dV (slice)
      e8.out = V.g; e8.ldo = V.w.ld;
      return gemm(c, c->dF.p, c->dF.ld, 1, c->Pb.p, c->Pb.ld, 1, c->W, c->k, B, e8);
    }
    case 5:
      // Adam on the batch rows of the (replicated) user factors and on the item-factor slice; the logged loss
      // is this rank's share: (1-a) * (slice of the reconstruction sum) / N [+ a * fm on rank 0] + reg/2 * l2
      c->last_ids_offset = ids_offset;
      RC(g_apply_impl(c, ids_offset, B, B, lr, reg, alpha, loss_slot));
      finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot, 1.f, c->tp_rank == 0 ? 1.f : 0.f);
      CU(cudaGetLastError());
      c->launches++;
      return 0;
    default:
      return fail("ganmf_tp_g_phase: phase must be 1..10");
  }
}

// ---- public step API -------------------------------------------------------------------------
int ganmf_d_forward(ganmf_ctx* c, int ids_offset, int B) {
  RC(check_batch(c, ids_offset, B));
  return c->cfg.kind == GANMF_KIND_GANMF ? ganmf_d_forward_impl(c, ids_offset, B)
                                         : dis_d_forward_impl(c, ids_offset, B);
}
int ganmf_d_forward_phase(ganmf_ctx* c, int ids_offset, int B, int phase) {
  RC(check_batch(c, ids_offset, B));
  if (c->cfg.kind != GANMF_KIND_GANMF) return fail("phased forward: GANMF only");
  if (phase < 0 || phase > 2) return fail("phase must be 0, 1 or 2");
  return ganmf_d_forward_impl(c, ids_offset, B, phase);
}
int ganmf_d_backward(ganmf_ctx* c, int B, int n_global, float m_hinge) {
  if (!c) return fail("null ctx");
  return c->cfg.kind == GANMF_KIND_GANMF ? ganmf_d_backward_impl(c, B, n_global, m_hinge, 0)
                                         : dis_d_backward_impl(c, B, n_global);
}
int ganmf_d_backward_phase(ganmf_ctx* c, int B, int n_global, float m_hinge, int phase) {
  if (!c) return fail("null ctx");
  if (c->cfg.kind != GANMF_KIND_GANMF) return fail("phased backward: GANMF only");
  if (phase < 0 || phase > 4) return fail("phase must be 0..4");
  return ganmf_d_backward_impl(c, B, n_global, m_hinge, phase);
}
int ganmf_d_apply(ganmf_ctx* c, float lr, float reg, int loss_slot) {
  if (!c) return fail("null ctx");
  return d_apply_impl(c, lr, reg, loss_slot);
}

int ganmf_g_forward_backward(ganmf_ctx* c, int ids_offset, int B, int n_global, float alpha) {
  RC(check_batch(c, ids_offset, B));
  c->last_ids_offset = ids_offset;
  return c->cfg.kind == GANMF_KIND_GANMF ? ganmf_g_fb_impl(c, ids_offset, B, n_global, alpha)
                                         : dis_g_fb_impl(c, ids_offset, B, n_global, alpha);
}
int ganmf_g_forward_backward_part(ganmf_ctx* c, int ids_offset, int B, int n_global, float alpha, int part) {
  RC(check_batch(c, ids_offset, B));
  if (c->cfg.kind != GANMF_KIND_GANMF) return fail("two-part G backward: GANMF only");
  if (part < 1 || part > 2) return fail("part must be 1 or 2");
  c->last_ids_offset = ids_offset;
  return ganmf_g_fb_impl(c, ids_offset, B, n_global, alpha, false, 0.f, 0.f, part);
}
int ganmf_g_apply(ganmf_ctx* c, int B, int n_global, float lr, float reg, float alpha, int loss_slot) {
  if (!c) return fail("null ctx");
  return g_apply_impl(c, c->last_ids_offset, B, n_global, lr, reg, alpha, loss_slot);
}
int ganmf_d_step(ganmf_ctx* c, int ids_offset, int B, int n_global, float lr, float reg, float m_hinge,
                 int loss_slot) {
  RC(ganmf_d_forward(c, ids_offset, B));
  if (c->cfg.kind == GANMF_KIND_GANMF && c->fuse_adam)
    return ganmf_d_backward_apply_fused(c, B, n_global, m_hinge, lr, reg, loss_slot);
  RC(ganmf_d_backward(c, B, n_global, m_hinge));
  return d_apply_impl(c, lr, reg, loss_slot);
}
int ganmf_g_step(ganmf_ctx* c, int ids_offset, int B, int n_global, float lr, float reg, float alpha,
                 int loss_slot) {
  // (The item-factor gradient GEMM is too short -- K = B -- to hide an optimiser epilogue: measured
  //  0.076 -> 0.166 ms at cfg4, so the G step keeps its one fused Adam launch over {P, V}.)
  RC(ganmf_g_forward_backward(c, ids_offset, B, n_global, alpha));
  return g_apply_impl(c, ids_offset, B, n_global, lr, reg, alpha, loss_slot);
}

// Adam on arbitrary element ranges of the discriminator slab (theta, m, v and grad share offsets): under
// data parallelism each rank updates only the chunk whose summed gradient it received from the
// reduce-scatter, then the updated parameters are all-gathered.  sum(theta^2) of the ranges goes to
// l2_shard (summed over ranks by the caller before ganmf_finalize_loss).
int ganmf_d_apply_ranges(ganmf_ctx* c, float lr, float reg, const int64_t* offsets, const int64_t* counts,
                         int n_ranges, int new_step) {
  if (!c || !offsets || !counts || n_ranges < 1 || n_ranges > ADAM_MAX_SEG) return fail("bad argument");
  AdamArgs a;
  memset(&a, 0, sizeof a);
  a.nseg = n_ranges;
  for (int i = 0; i < n_ranges; ++i) {
    if (offsets[i] < 0 || counts[i] < 0 || (offsets[i] & 3) || (counts[i] & 3) ||
        (size_t)(offsets[i] + counts[i]) > c->d_elems)
      return fail("range %d outside the discriminator slab or not a multiple of 4", i);
    AdamSeg& s = a.seg[i];
    s.theta = c->d_slab + offsets[i];
    s.m = c->d_slab + c->d_elems + offsets[i];
    s.v = c->d_slab + 2 * c->d_elems + offsets[i];
    s.g = c->d_slab + 3 * c->d_elems + offsets[i];
    s.slot = nullptr; s.ld = 4;
    s.n4 = (unsigned long long)counts[i] / 4;
  }
  if (new_step) c->last_alpha_d = adam_alpha(c, 0, lr);     // one optimiser step may span several calls
  a.alpha = c->last_alpha_d;
  a.reg = reg;
  a.l2_out = &c->sc->l2_shard;
  a.l2_shard_out = &c->sc->l2_shard;
  c->launches++;
  CU(fused_adam(a, c->st));
  return 0;
}

int ganmf_finalize_loss(ganmf_ctx* c, float reg, int loss_slot) {
  if (!c || loss_slot < 0 || loss_slot >= c->losses_cap) return fail("bad argument");
  finalize_loss_kernel<<<1, 1, 0, c->st>>>(c->sc, reg, c->losses, loss_slot);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}

// the device-side loss log grows on demand (the reference has no limit on batches per epoch)
static int ensure_losses(ganmf_ctx* c, long long n) {
  if (n <= c->losses_cap) return 0;
  if (n > (1LL << 30)) return fail("loss log of %lld entries", n);
  CU(cudaStreamSynchronize(c->st));
  float* bigger = nullptr;
  RC(dalloc(&bigger, (size_t)n));
  CU(cudaMemcpy(bigger, c->losses, (size_t)c->losses_cap * 4, cudaMemcpyDeviceToDevice));
  cudaFree(c->losses);
  c->losses = bigger;
  c->losses_cap = (int)n;
  return 0;
}

int ganmf_read_losses(ganmf_ctx* c, float* host, int n) {
  if (!c || n < 0 || n > c->losses_cap) return fail("bad loss count");
  CU(cudaMemcpyAsync(host, c->losses, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int ganmf_train_epoch(ganmf_ctx* c, const int32_t* perm, int n_ids, int batch, int d_steps, int g_steps,
                      float d_lr, float g_lr, float d_reg, float g_reg, float m_hinge, float alpha,
                      float* d_losses, float* g_losses) {
  if (!c || !perm || n_ids <= 0 || batch <= 0) return fail("bad argument");
  if (batch > c->B) return fail("batch_size %d exceeds max_batch %d", batch, c->B);
  const int nb = (n_ids + batch - 1) / batch;
  RC(ensure_losses(c, (long long)nb * (d_steps + g_steps)));
  RC(ganmf_upload_ids(c, perm, n_ids));
  int slot = 0;
  for (int s = 0; s < d_steps; ++s)
    for (int b = 0; b < nb; ++b) {
      const int off = b * batch, B = std::min(batch, n_ids - off);
      RC(ganmf_d_step(c, off, B, B, d_lr, d_reg, m_hinge, slot++));
    }
  const int n_d = slot;
  for (int s = 0; s < g_steps; ++s)
    for (int b = 0; b < nb; ++b) {
      const int off = b * batch, B = std::min(batch, n_ids - off);
      RC(ganmf_g_step(c, off, B, B, g_lr, g_reg, alpha, slot++));
    }
  std::vector<float> tmp((size_t)slot);
  RC(ganmf_read_losses(c, tmp.data(), slot));
  if (d_losses) memcpy(d_losses, tmp.data(), (size_t)n_d * 4);
  if (g_losses) memcpy(g_losses, tmp.data() + n_d, (size_t)(slot - n_d) * 4);
  return 0;
}

int ganmf_device_buffer(ganmf_ctx* c, const char* name, void** ptr, int64_t* n) {
  if (!c || !name || !ptr || !n) return fail("null argument");
  if (!strcmp(name, "d_grads")) { *ptr = c->d_slab + 3 * c->d_elems; *n = (int64_t)c->d_elems; return 0; }
  if (!strcmp(name, "d_grads_enc")) {        // dWe | dbe
    *ptr = c->params[0].g; *n = (int64_t)(c->params[0].w.elems() + c->params[1].w.elems()); return 0;
  }
  if (!strcmp(name, "d_grads_dec")) {        // dWd | dbd
    *ptr = c->params[2].g; *n = (int64_t)(c->params[2].w.elems() + c->params[3].w.elems()); return 0;
  }
  if (!strcmp(name, "d_params")) { *ptr = c->d_slab; *n = (int64_t)c->d_elems; return 0; }
  if (!strcmp(name, "d_params_enc")) {
    *ptr = c->params[0].w.p; *n = (int64_t)(c->params[0].w.elems() + c->params[1].w.elems()); return 0;
  }
  if (!strcmp(name, "d_params_dec")) {
    *ptr = c->params[2].w.p; *n = (int64_t)(c->params[2].w.elems() + c->params[3].w.elems()); return 0;
  }
  if (!strcmp(name, "g_shared_grad")) { *ptr = c->v_slab + 3 * c->v_elems; *n = (int64_t)c->v_elems; return 0; }
  if (!strcmp(name, "step_scalars")) { *ptr = c->sc; *n = 7; return 0; }
  if (!strcmp(name, "tp_h2")) { *ptr = c->H2.p; *n = (int64_t)c->H2.elems(); return 0; }
  if (!strcmp(name, "tp_dh2")) { *ptr = c->dH2.p; *n = (int64_t)c->dH2.elems(); return 0; }
  if (!strcmp(name, "tp_dpb")) { *ptr = c->dPb.p; *n = (int64_t)c->dPb.elems(); return 0; }
  if (!strcmp(name, "tp_m1")) { *ptr = c->M1.p; *n = (int64_t)c->M1.elems(); return 0; }
  if (!strcmp(name, "user_factors")) {
    RC(p_flush(c));                          // deferred optimiser steps first: the caller reads the matrix
    *ptr = c->params[c->n_d].w.p; *n = (int64_t)c->params[c->n_d].w.elems(); return 0;
  }
  if (!strcmp(name, "item_factors")) {
    *ptr = c->params[c->n_d + 1].w.p; *n = (int64_t)c->params[c->n_d + 1].w.elems(); return 0;
  }
  return fail("unknown buffer %s", name);
}
int ganmf_device_buffer_ld(ganmf_ctx* c, const char* name, int* ld) {
  if (!c || !name || !ld) return fail("null argument");
  if (!strcmp(name, "tp_h2")) { *ld = c->H2.ld; return 0; }
  if (!strcmp(name, "tp_dh2")) { *ld = c->dH2.ld; return 0; }
  if (!strcmp(name, "tp_dpb")) { *ld = c->dPb.ld; return 0; }
  if (!strcmp(name, "tp_m1")) { *ld = c->M1.ld; return 0; }
  if (!strcmp(name, "user_factors")) { *ld = c->params[c->n_d].w.ld; return 0; }
  if (!strcmp(name, "item_factors")) { *ld = c->params[c->n_d + 1].w.ld; return 0; }
  return fail("no leading dimension for buffer %s", name);
}

// ------------------------------------------------------------------------------ scoring / eval
static int ensure_eval_buffers(ganmf_ctx* c, int block, int K, int n_cut) {
  const int n_items = c->cfg.item_mode ? c->cfg.n_rows : c->W;
  const int ild = rup(n_items, 32);
  const size_t need = (size_t)block * ild;
  if (need > c->scores_elems) {
    cudaFree(c->scores);
    RC(dalloc(&c->scores, need));
    c->scores_elems = need;
  }
  if (c->Fb.rows < block) {
    cudaFree(c->Fb.p);
    RC(mat_alloc(&c->Fb, block, 3 * c->k));
  }
  if (!c->Ob.p) RC(mat_alloc(&c->Ob, n_items, 3 * c->k));
  if (c->eval_users_cap < block) {
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)block));
    c->eval_users_cap = block;
  }
  if ((size_t)block * K > c->topk_cap) {
    cudaFree(c->topk_idx); cudaFree(c->topk_val);
    RC(dalloc(&c->topk_idx, (size_t)block * K));
    RC(dalloc(&c->topk_val, (size_t)block * K));
    c->topk_cap = (size_t)block * K;
  }
  if (n_cut > 0) {
    const size_t uv = (size_t)block * n_cut * MC_NCOL;
    if (uv > c->uvals_cap) {
      cudaFree(c->uvals);
      RC(dalloc(&c->uvals, uv));
      c->uvals_cap = uv;
    }
    const size_t ic = (size_t)n_cut * n_items;
    if (ic > c->icounts_cap) {
      cudaFree(c->icounts); cudaFree(c->usums); cudaFree(c->cut_dev);
      RC(dalloc(&c->icounts, ic));
      RC(dalloc(&c->usums, (size_t)64 * MC_NCOL));
      RC(dalloc(&c->cut_dev, (size_t)64));
      c->icounts_cap = ic;
    }
  }
  return 0;
}

// Ranking needs fp32-accurate scores (near-ties decide the order), so scoring runs the tensor cores
// on split-TF32 operands: one GEMM over K' = 3k (see split3_rows_kernel).
static int prepare_item_factors(ganmf_ctx* c) {
  RC(p_flush(c));
  const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
  split3_rows_kernel<<<other.w.rows, 64, 0, c->st>>>(other.w.p, other.w.ld, nullptr, c->Ob.p, c->Ob.ld, c->k, 1);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}
// scores[n, items] for the users already in c->eval_users (device); prepare_item_factors() first
static int score_block(ganmf_ctx* c, int n, const int* users_dev = nullptr) {
  if (!users_dev) users_dev = c->eval_users;
  const Param& rows_of = c->cfg.item_mode ? c->params[c->n_d + 1] : c->params[c->n_d];
  const int n_items = c->Ob.rows;
  split3_rows_kernel<<<n, 64, 0, c->st>>>(rows_of.w.p, rows_of.w.ld, users_dev, c->Fb.p, c->Fb.ld, c->k, 0);
  CU(cudaGetLastError());
  c->launches++;
  Epilogue e;
  e.out = c->scores; e.ldo = rup(n_items, 32);
  c->presplit = true;                       // operands are already split: never split them again
  const int rc = gemm(c, c->Fb.p, c->Fb.ld, 0, c->Ob.p, c->Ob.ld, 0, n, n_items, 3 * c->k, e);
  c->presplit = false;
  return rc;
}

static int n_items_of(ganmf_ctx* c) { return c->cfg.item_mode ? c->cfg.n_rows : c->W; }
static int n_users_of(ganmf_ctx* c) { return c->cfg.item_mode ? c->W : c->cfg.n_rows; }

static int check_users(ganmf_ctx* c, const int32_t* u, int n) {
  const int nu = n_users_of(c);
  for (int i = 0; i < n; ++i)
    if (u[i] < 0 || u[i] >= nu) return fail("user id %d outside [0, %d)", u[i], nu);
  return 0;
}

int ganmf_score(ganmf_ctx* c, const int32_t* users, int n, float* scores_host) {
  if (!c || !users || !scores_host || n <= 0) return fail("bad argument");
  RC(check_users(c, users, n));
  RC(ensure_eval_buffers(c, n, 1, 0));
  const int n_items = n_items_of(c), ild = rup(n_items, 32);
  CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  RC(prepare_item_factors(c));
  RC(score_block(c, n));
  CU(cudaMemcpy2DAsync(scores_host, (size_t)n_items * 4, c->scores, (size_t)ild * 4, (size_t)n_items * 4, n,
                       cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int ganmf_encode(ganmf_ctx* c, const int32_t* rows, int n, float* codes_host) {
  if (!c || !rows || !codes_host || n <= 0) return fail("bad argument");
  if (c->cfg.kind != GANMF_KIND_GANMF) return fail("encode: GANMF only");
  const Csr& tr = c->csr[GANMF_CSR_TRAIN];
  if (!tr.indptr) return fail("train CSR not set");
  for (int i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= c->cfg.n_rows) return fail("row id %d out of range", rows[i]);
  Param *We = &c->params[0], *be = &c->params[1];
  for (int s = 0; s < n; s += c->B) {
    const int b = std::min(c->B, n - s);
    CU(cudaMemcpyAsync(c->ids, rows + s, (size_t)b * 4, cudaMemcpyHostToDevice, c->st));
    // API edge: exact fp32 gather-sum over the rows' interactions (no dense profile tile, no tensor-core rounding)
    CU(csr_encode_rows(tr.indptr, tr.indices, tr.data, c->ids, b, We->w.p, We->w.ld, c->H2.ld, be->w.p, c->H2.p,
                       c->H2.ld, c->st));
    c->launches++;
    CU(cudaMemcpy2DAsync(codes_host + (size_t)s * c->E, (size_t)c->E * 4, c->H2.p, (size_t)c->H2.ld * 4,
                         (size_t)c->E * 4, b, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
  }
  return 0;
}

static int mask_and_topk(ganmf_ctx* c, int n, int n_items, int remove_seen, int K, const int* users_dev = nullptr) {
  if (!users_dev) users_dev = c->eval_users;
  if (K < 1 || K > TK_MAXK) return fail("top-K supports 1 <= K <= %d (got %d)", TK_MAXK, K);
  const int ild = rup(n_items, 32);
  if (remove_seen) {
    const Csr& seen = c->csr[GANMF_CSR_SEEN];
    if (!seen.indptr) return fail("seen CSR not set");
    if (seen.n_cols != n_items) return fail("seen CSR has %d columns, scores have %d", seen.n_cols, n_items);
    mask_seen_kernel<<<n, 128, 0, c->st>>>(c->scores, ild, users_dev, seen.indptr, seen.indices);
    CU(cudaGetLastError());
    c->launches++;
  }
  CU(topk_rows(c->scores, ild, n, n_items, K, c->topk_idx, c->topk_val, c->st));
  c->launches++;
  return 0;
}

// ---- fused scoring -> mask -> top-K (score_select.cuh) ----------------------------------------------------
constexpr int FB_BLOCK = 64;           // fallback rows scored exactly per pass
// |tf32 score - exact| <= gamma * ||q|| * ||v||: both operands rounded to tf32 (2^-11 each if the TMA rounds to
// nearest, 2^-10 if it truncated -- the bound assumes the worse), products summed in fp32 by the tensor core
static float fused_gamma(int k) { return 1.05f * (2.0f / 1024.0f + (float)(k + 8) * 1.2e-7f); }
static bool fused_ok(ganmf_ctx* c, int K) { return c->eval_fused && K <= 24 && c->k <= SS_MAX_KB * TC_BK; }

static int ensure_fused_buffers(ganmf_ctx* c, int block, int K, int n_sets) {
  const int n_items = c->cfg.item_mode ? c->cfg.n_rows : c->W;
  if (!c->st_res) {
    CU(cudaStreamCreateWithFlags(&c->st_res, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->st_fin, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_setup, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_pipe_done, cudaEventDisableTiming));
  }
  for (int i = 0; i < n_sets; ++i) {
    ganmf_ctx::FusedBuf& b = c->fbuf[i];
    if (!b.ev_sel) {
      CU(cudaEventCreateWithFlags(&b.ev_sel, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.ev_res, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.ev_fin, cudaEventDisableTiming));
      RC(dalloc(&b.fb_count, 1));
      CU(cudaMallocHost((void**)&b.h_count, 4));
    }
    if (b.Qg.rows < block) {
      cudaFree(b.Qg.p);
      RC(mat_alloc(&b.Qg, block, c->k));
    }
    if ((size_t)block > b.rows_cap) {
      cudaFree(b.fb_rows); cudaFree(b.row_thr);
      RC(dalloc(&b.fb_rows, (size_t)block));
      RC(dalloc(&b.row_thr, (size_t)block));
      b.rows_cap = block;
    }
    if ((size_t)block * K > b.topk_cap) {
      cudaFree(b.topk_idx); cudaFree(b.topk_val);
      RC(dalloc(&b.topk_idx, (size_t)block * K));
      RC(dalloc(&b.topk_val, (size_t)block * K));
      b.topk_cap = (size_t)block * K;
    }
  }
  if (c->eval_users_cap < block) {
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)block));
    c->eval_users_cap = block;
  }
  if (!c->vmax_bits) RC(dalloc(&c->vmax_bits, 1));
  if (!c->fb_users) RC(dalloc(&c->fb_users, (size_t)FB_BLOCK));
  if ((size_t)FB_BLOCK * K > c->fb_topk_cap) {
    cudaFree(c->fb_idx); cudaFree(c->fb_val);
    RC(dalloc(&c->fb_idx, (size_t)FB_BLOCK * K));
    RC(dalloc(&c->fb_val, (size_t)FB_BLOCK * K));
    c->fb_topk_cap = (size_t)FB_BLOCK * K;
  }
  const size_t sc_need = (size_t)FB_BLOCK * rup(n_items, 32);    // exact score rows of the fallback
  if (sc_need > c->scores_elems) {
    cudaFree(c->scores);
    RC(dalloc(&c->scores, sc_need));
    c->scores_elems = sc_need;
  }
  return 0;
}

// once per call: bring the lazily updated factors up to date, norm bound of the ranked side
static int prepare_fused(ganmf_ctx* c) {
  RC(p_flush(c));
  const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
  CU(cudaMemsetAsync(c->vmax_bits, 0, 4, c->st));
  row_norm_max_kernel<<<((other.w.rows + 3) / 4 * 32 + 255) / 256, 256, 0, c->st>>>(other.w.p, other.w.rows, c->k,
                                                                                    other.w.ld, c->vmax_bits);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}

// stage 1 (tensor cores): query rows of the block -> candidate lists.  prepare_fused() first.
static int fused_select(ganmf_ctx* c, ganmf_ctx::FusedBuf& b, int n, const int* users_dev, int remove_seen, int K,
                        cudaStream_t st) {
  const Param& rows_of = c->cfg.item_mode ? c->params[c->n_d + 1] : c->params[c->n_d];
  const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
  const int n_items = other.w.rows;
  const Csr& seen = c->csr[GANMF_CSR_SEEN];
  if (remove_seen) {
    if (!seen.indptr) return fail("seen CSR not set");
    if (seen.n_cols != n_items) return fail("seen CSR has %d columns, scores have %d", seen.n_cols, n_items);
  }
  gather_rows_kernel<<<n, 64, 0, st>>>(rows_of.w.p, users_dev, b.Qg.p, b.Qg.ld);
  CU(cudaGetLastError());
  ScoreSelectCall sc;
  sc.Q = b.Qg.p; sc.ldq = b.Qg.ld; sc.V = other.w.p; sc.ldv = other.w.ld;
  sc.n_rows = n; sc.n_items = n_items; sc.k = c->k;
  sc.KP = K <= 10 ? 16 : 32;
  sc.segs = score_select_segments(n, n_items, c->gemm_sm_cap > 0 ? c->gemm_sm_cap : 148);
  sc.users = users_dev;
  sc.seen_indptr = remove_seen ? seen.indptr : nullptr;
  sc.seen_indices = remove_seen ? seen.indices : nullptr;
  const size_t need = (size_t)n * 2 * sc.segs * sc.KP;             // candidate lists of this block
  if (need > b.cand_cap) {
    CU(cudaDeviceSynchronize());
    cudaFree(b.cand_val); cudaFree(b.cand_idx);
    RC(dalloc(&b.cand_val, need));
    RC(dalloc(&b.cand_idx, need));
    b.cand_cap = need;
  }
  sc.cand_val = b.cand_val; sc.cand_idx = b.cand_idx;
  sc.row_thr = b.row_thr;
  CU(cudaMemsetAsync(b.row_thr, 0, (size_t)n * 4, st));
  sc.cache = &c->tmaps; sc.max_ctas = c->gemm_sm_cap;
  cudaError_t e = score_select(sc, st);
  if (e != cudaSuccess) return fail("score_select(n=%d items=%d k=%d) -> %s", n, n_items, c->k, cudaGetErrorString(e));
  c->launches += 2;
  return 0;
}

// stage 2: exact re-scoring + certificate -> top-K lists of the block, number of uncertified rows -> b.h_count
static int fused_rescore(ganmf_ctx* c, ganmf_ctx::FusedBuf& b, int n, int K, cudaStream_t st) {
  const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
  const int n_items = other.w.rows;
  const int NL = 2 * score_select_segments(n, n_items, c->gemm_sm_cap > 0 ? c->gemm_sm_cap : 148);
  const float gamma = fused_gamma(c->k);
  CU(cudaMemsetAsync(b.fb_count, 0, 4, st));
  if (K <= 10)
    rescore_kernel<16><<<(n + 3) / 4, 128, 0, st>>>(b.cand_val, b.cand_idx, NL, n, K, b.Qg.p, b.Qg.ld, other.w.p,
                                                    other.w.ld, c->k, c->vmax_bits, b.row_thr, gamma, b.topk_idx,
                                                    b.topk_val, b.fb_count, b.fb_rows);
  else
    rescore_kernel<32><<<(n + 3) / 4, 128, 0, st>>>(b.cand_val, b.cand_idx, NL, n, K, b.Qg.p, b.Qg.ld, other.w.p,
                                                    other.w.ld, c->k, c->vmax_bits, b.row_thr, gamma, b.topk_idx,
                                                    b.topk_val, b.fb_count, b.fb_rows);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(b.h_count, b.fb_count, 4, cudaMemcpyDeviceToHost, st));
  c->launches++;
  return 0;
}

// stage 3 (needs the host-side count): rows without a certificate get exact score rows -> the materialised
// mask / top-k kernels -> back to their slots
static int fused_fallback(ganmf_ctx* c, ganmf_ctx::FusedBuf& b, int n, const int* users_dev, int remove_seen, int K,
                          int n_fb, cudaStream_t st) {
  const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
  const int n_items = other.w.rows;
  const Csr& seen = c->csr[GANMF_CSR_SEEN];
  c->fused_rows += n;
  c->fallback_rows += n_fb;
  const int ild = rup(n_items, 32);
  for (int f0 = 0; f0 < n_fb; f0 += FB_BLOCK) {
    const int nf = std::min(FB_BLOCK, n_fb - f0);
    exact_score_rows_kernel<<<dim3(std::max(1, 148 * 4 / nf), nf), 256, 0, st>>>(
        b.fb_rows, f0, nf, b.Qg.p, b.Qg.ld, other.w.p, other.w.ld, c->k, n_items, c->scores, ild);
    gather_ids_kernel<<<1, FB_BLOCK, 0, st>>>(users_dev, b.fb_rows, f0, nf, c->fb_users);
    if (remove_seen) mask_seen_kernel<<<nf, 128, 0, st>>>(c->scores, ild, c->fb_users, seen.indptr, seen.indices);
    CU(cudaGetLastError());
    CU(topk_rows(c->scores, ild, nf, n_items, K, c->fb_idx, c->fb_val, st));
    scatter_topk_kernel<<<(nf * K + 127) / 128, 128, 0, st>>>(c->fb_idx, c->fb_val, b.fb_rows, f0, nf, K, b.topk_idx,
                                                             b.topk_val);
    CU(cudaGetLastError());
    c->launches += 5;
  }
  return 0;
}

// recommend(): one block, everything in stream order on the context's stream; lists end up in c->fbuf[0]
static int fused_topk_block(ganmf_ctx* c, int n, const int* users_dev, int remove_seen, int K) {
  ganmf_ctx::FusedBuf& b = c->fbuf[0];
  RC(fused_select(c, b, n, users_dev, remove_seen, K, c->st));
  RC(fused_rescore(c, b, n, K, c->st));
  CU(cudaStreamSynchronize(c->st));
  return fused_fallback(c, b, n, users_dev, remove_seen, K, *b.h_count, c->st);
}

int ganmf_eval_stats(ganmf_ctx* c, int64_t* fused_rows, int64_t* fallback_rows) {
  if (!c) return fail("null ctx");
  if (fused_rows) *fused_rows = c->fused_rows;
  if (fallback_rows) *fallback_rows = c->fallback_rows;
  return 0;
}

int ganmf_mask_topk(ganmf_ctx* c, float* scores_host, int n, int n_items, const int32_t* users, int remove_seen,
                    int K, int32_t* idx_host, float* val_host, int write_back) {
  if (!c || !scores_host || n <= 0 || n_items <= 0) return fail("bad argument");
  if (remove_seen && !users) return fail("user ids required for the seen mask");
  const int ild = rup(n_items, 32);
  const size_t need = (size_t)n * ild;
  if (need > c->scores_elems) {
    cudaFree(c->scores);
    RC(dalloc(&c->scores, need));
    c->scores_elems = need;
  }
  if (c->eval_users_cap < n) {
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)n));
    c->eval_users_cap = n;
  }
  if ((size_t)n * K > c->topk_cap) {
    cudaFree(c->topk_idx); cudaFree(c->topk_val);
    RC(dalloc(&c->topk_idx, (size_t)n * K));
    RC(dalloc(&c->topk_val, (size_t)n * K));
    c->topk_cap = (size_t)n * K;
  }
  CU(cudaMemcpy2DAsync(c->scores, (size_t)ild * 4, scores_host, (size_t)n_items * 4, (size_t)n_items * 4, n,
                       cudaMemcpyHostToDevice, c->st));
  if (users) CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  RC(mask_and_topk(c, n, n_items, remove_seen, K));
  CU(cudaMemcpyAsync(idx_host, c->topk_idx, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
  CU(cudaMemcpyAsync(val_host, c->topk_val, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
  if (write_back)
    CU(cudaMemcpy2DAsync(scores_host, (size_t)n_items * 4, c->scores, (size_t)ild * 4, (size_t)n_items * 4, n,
                         cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int ganmf_recommend(ganmf_ctx* c, const int32_t* users, int n, int remove_seen, int K, int32_t* idx_host,
                    float* val_host, float* masked_scores_host) {
  if (!c || !users || n <= 0) return fail("bad argument");
  RC(check_users(c, users, n));
  if (K < 1 || K > TK_MAXK) return fail("top-K supports 1 <= K <= %d (got %d)", TK_MAXK, K);
  if (!masked_scores_host && fused_ok(c, K)) {
    // nobody asked for the score rows: fused scorer, the n x n_items matrix never exists
    RC(ensure_fused_buffers(c, n, K, 1));
    CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    RC(prepare_fused(c));
    RC(fused_topk_block(c, n, c->eval_users, remove_seen, K));
    CU(cudaMemcpyAsync(idx_host, c->fbuf[0].topk_idx, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
    if (val_host) CU(cudaMemcpyAsync(val_host, c->fbuf[0].topk_val, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
  }
  RC(ensure_eval_buffers(c, n, K, 0));
  const int n_items = n_items_of(c), ild = rup(n_items, 32);
  CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  RC(prepare_item_factors(c));
  RC(score_block(c, n));
  RC(mask_and_topk(c, n, n_items, remove_seen, K));
  CU(cudaMemcpyAsync(idx_host, c->topk_idx, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
  if (val_host) CU(cudaMemcpyAsync(val_host, c->topk_val, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->st));
  if (masked_scores_host)
    CU(cudaMemcpy2DAsync(masked_scores_host, (size_t)n_items * 4, c->scores, (size_t)ild * 4,
                         (size_t)n_items * 4, n, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

template <typename T>
static int upload(T** dev, const T* host, size_t n) {
  cudaFree(*dev);
  RC(dalloc(dev, n));
  CU(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int ganmf_set_eval_tables(ganmf_ctx* c, const float* gain, const float* gain_desc, const float* logtab,
                          int logtab_n, const double* nov, const uint8_t* haspop, const double* popn, int n_items_tables) {
  if (!c) return fail("null ctx");
  const Csr& te = c->csr[GANMF_CSR_TEST];
  if (!te.indptr) return fail("test CSR not set");
  if (logtab_n < TK_MAXK) return fail("logtab needs >= %d entries", TK_MAXK);
  const int n_items = te.n_cols;
  if (n_items_tables != n_items)
    return fail("per-item tables have %d entries, the test matrix has %d columns", n_items_tables, n_items);
  RC(upload(&c->tb_gain, gain, (size_t)te.nnz));
  RC(upload(&c->tb_gain_desc, gain_desc, (size_t)te.nnz));
  RC(upload(&c->tb_logtab, logtab, (size_t)logtab_n));
  RC(upload(&c->tb_nov, nov, (size_t)n_items));
  RC(upload(&c->tb_popn, popn, (size_t)n_items));
  RC(upload((uint8_t**)&c->tb_haspop, haspop, (size_t)n_items));
  cudaFree(c->rmse_scratch);
  RC(dalloc(&c->rmse_scratch, (size_t)te.nnz));
  c->tb.test_indptr = te.indptr; c->tb.test_indices = te.indices;
  c->tb.test_gain = c->tb_gain; c->tb.test_gain_desc = c->tb_gain_desc; c->tb.logtab = c->tb_logtab;
  c->tb.item_novelty = c->tb_nov; c->tb.item_has_pop = c->tb_haspop; c->tb.item_popnorm = c->tb_popn;
  c->have_tables = true;
  return 0;
}

// metric stage for n rows whose lists are in c->topk_idx and users in c->eval_users
// per-user metric values for n rows whose lists are in c->topk_idx; `uvals` points at the first of those
// rows inside the per-user value table.  The running sums are formed afterwards by accumulate_users().
static int metrics_block(ganmf_ctx* c, int n, int K, int n_cut, int n_items, bool with_rmse,
                         const int* users_dev, double* uvals, ganmf_ctx::FusedBuf* fb = nullptr, int remove_seen = 0,
                         cudaStream_t st = nullptr) {
  const int total = n * n_cut;
  const bool fused = fb != nullptr;
  if (!fused) st = c->st;
  user_metrics_kernel<<<(total + 127) / 128, 128, 0, st>>>(fused ? fb->topk_idx : c->topk_idx, K, users_dev, n, c->cut_dev,
                                                         n_cut, c->tb, uvals, c->icounts, n_items);
  CU(cudaGetLastError());
  c->launches++;
  if (with_rmse && fused) {
    // no score matrix exists: exact scores of the test items only (query rows are still in c->Qg)
    const Param& other = c->cfg.item_mode ? c->params[c->n_d] : c->params[c->n_d + 1];
    const Csr& te = c->csr[GANMF_CSR_TEST];
    const Csr& seen = c->csr[GANMF_CSR_SEEN];
    user_rmse_exact_kernel<<<(n + 3) / 4, 128, 0, st>>>(fb->Qg.p, fb->Qg.ld, other.w.p, other.w.ld, c->k, users_dev, n,
                                                         n_cut, c->tb, te.data, remove_seen ? seen.indptr : nullptr,
                                                         remove_seen ? seen.indices : nullptr, c->rmse_scratch, uvals);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
  }
  if (with_rmse) {
    const Csr& te = c->csr[GANMF_CSR_TEST];
    user_rmse_kernel<<<(n + 63) / 64, 64, 0, c->st>>>(c->scores, rup(n_items, 32), users_dev, n, n_cut, c->tb,
                                                    te.data, c->rmse_scratch, uvals);
    CU(cudaGetLastError());
    c->launches++;
  }
  return 0;
}
// sums[cutoff][metric] = sum over users IN ORDER (Evaluator.py:305-335 keeps one running sum)
static int accumulate_users(ganmf_ctx* c, int n_users, int n_cut, size_t first_user = 0, cudaStream_t st = nullptr) {
  const int ncols = n_cut * MC_NCOL;                     // <= 32 * 13 = 416 < OA_THREADS? no: handled below
  if (ncols > OA_THREADS) return fail("too many (cutoff, metric) columns for the ordered accumulation");
  // (continues the running sums already in c->usums: blocks of users are fed in order)
  ordered_accumulate_kernel<<<1, OA_THREADS, 0, st ? st : c->st>>>(c->uvals + first_user * ncols, n_users, ncols, c->usums);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}

static int eval_prologue(ganmf_ctx* c, const int32_t* cutoffs, int n_cut, int* Kout) {
  if (!c->have_tables) return fail("call ganmf_set_eval_tables first");
  if (n_cut < 1 || n_cut * MC_NCOL > OA_THREADS) return fail("1..%d cutoffs supported", OA_THREADS / MC_NCOL);
  int K = 0;
  for (int i = 0; i < n_cut; ++i) K = std::max(K, cutoffs[i]);
  if (K > TK_MAXK) return fail("cutoff %d > %d", K, TK_MAXK);
  *Kout = K;
  return 0;
}

static int evaluate_values_impl(ganmf_ctx* c, const int32_t* users, int n_users, const int32_t* cutoffs, int n_cut,
                                int remove_seen, int block, bool inline_sums);

int ganmf_evaluate(ganmf_ctx* c, const int32_t* users, int n_users, const int32_t* cutoffs, int n_cut,
                   int remove_seen, int block, double* sums_host, int64_t* counts_host) {
  if (!sums_host) return fail("bad argument");
  // one GPU: the ordered running sums are formed block by block behind the scoring of the next block
  RC(evaluate_values_impl(c, users, n_users, cutoffs, n_cut, remove_seen, block, true));
  return ganmf_evaluate_sums(c, nullptr, sums_host, counts_host);
}

int ganmf_evaluate_values(ganmf_ctx* c, const int32_t* users, int n_users, const int32_t* cutoffs, int n_cut,
                          int remove_seen, int block) {
  return evaluate_values_impl(c, users, n_users, cutoffs, n_cut, remove_seen, block, false);
}

static int evaluate_values_impl(ganmf_ctx* c, const int32_t* users, int n_users, const int32_t* cutoffs, int n_cut,
                                int remove_seen, int block, bool inline_sums) {
  if (!c || !users || !cutoffs || n_users < 0) return fail("bad argument");
  c->ev_pending_users = -1;
  c->ev_sums_done = false;
  int K;
  RC(eval_prologue(c, cutoffs, n_cut, &K));
  RC(check_users(c, users, n_users));
  const int n_items = n_items_of(c);
  if (c->csr[GANMF_CSR_TEST].n_cols != n_items) return fail("test CSR column count mismatch");
  if (!c->csr[GANMF_CSR_TEST].data) return fail("test CSR needs ratings (data) for RMSE");
  // The reference scores min(1000, 1e8/n_items) users at a time to bound HOST memory (Evaluator.py:238);
  // the result does not depend on the block size, so on the device a block is as many users as a
  // 1 GiB score buffer holds (at most 8192): fuller kernels, fewer launches.
  const bool fused = fused_ok(c, K);
  if (fused) {
    // fused scorer: no score matrix.  Blocks of <= 64 K users, two buffer sets: the re-scoring, metric and running-sum
    // kernels of block i run on side streams while the tensor cores score block i+1.
    // (measured at I = 200 000: one 131 072-user block 21.5 ms; two overlapped 65 536-user blocks 21.8 ms -- the side
    //  kernels share the SMs with the persistent GEMM CTAs, so the overlap buys little; blocks stay as large as the
    //  candidate lists allow and only longer user lists are pipelined)
    if (block <= 0) {
      const int nblk = (n_users + (1 << 17) - 1) >> 17;
      block = nblk <= 1 ? n_users : ((n_users + nblk - 1) / nblk + SS_ROWS - 1) / SS_ROWS * SS_ROWS;
    }
    block = std::min(block, std::max(n_users, 1));
    RC(ensure_fused_buffers(c, block, K, 2));
    RC(ensure_eval_buffers(c, 1, K, n_cut));           // metric tables (usums, icounts, cut_dev)
  } else {
    if (block <= 0) block = (int)std::min<long long>(8192, std::max<long long>(1, (1LL << 28) / rup(n_items, 32)));
    // buffers are sized for the full block even when this call has fewer users, so later (larger) calls
    // never re-allocate: cudaMalloc of a GiB-class buffer costs far more than the evaluation itself
    RC(ensure_eval_buffers(c, block, K, n_cut));
    block = std::min(block, std::max(n_users, 1));
  }
  if (c->eval_users_cap < n_users) {                       // all user ids go up once
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)n_users));
    c->eval_users_cap = n_users;
  }
  const size_t uv = (size_t)n_users * n_cut * MC_NCOL;     // per-user value table for the whole call
  if (uv > c->uvals_cap) {
    cudaFree(c->uvals);
    RC(dalloc(&c->uvals, uv));
    c->uvals_cap = uv;
  }
  CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n_users * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemcpyAsync(c->cut_dev, cutoffs, (size_t)n_cut * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->icounts, 0, (size_t)n_cut * n_items * 4, c->st));
  if (inline_sums) CU(cudaMemsetAsync(c->usums, 0, (size_t)n_cut * MC_NCOL * 8, c->st));
  if (fused) {
    RC(prepare_fused(c));
    // setup (ids, tables, norms) is ordered before everything the side streams do
    CU(cudaEventRecord(c->ev_setup, c->st));
    CU(cudaStreamWaitEvent(c->st_res, c->ev_setup, 0));
    CU(cudaStreamWaitEvent(c->st_fin, c->ev_setup, 0));
    const int nb = (n_users + block - 1) / block;
    for (int i = 0; i <= nb; ++i) {
      if (i < nb) {                                    // tensor-core pass of block i, re-scoring right behind it
        ganmf_ctx::FusedBuf& b = c->fbuf[i & 1];
        const int s0 = i * block, n = std::min(block, n_users - s0);
        if (i >= 2) CU(cudaStreamWaitEvent(c->st, b.ev_fin, 0));          // the set is free again
        RC(fused_select(c, b, n, c->eval_users + s0, remove_seen, K, c->st));
        CU(cudaEventRecord(b.ev_sel, c->st));
        CU(cudaStreamWaitEvent(c->st_res, b.ev_sel, 0));
        RC(fused_rescore(c, b, n, K, c->st_res));
        CU(cudaEventRecord(b.ev_res, c->st_res));
      }
      if (i >= 1) {                                    // block i-1: fallback rows, metric values, running sums
        ganmf_ctx::FusedBuf& b = c->fbuf[(i - 1) & 1];
        const int s0 = (i - 1) * block, n = std::min(block, n_users - s0);
        const int* ud = c->eval_users + s0;
        CU(cudaEventSynchronize(b.ev_res));            // (the host needs the count of uncertified rows)
        RC(fused_fallback(c, b, n, ud, remove_seen, K, *b.h_count, c->st_fin));
        RC(metrics_block(c, n, K, n_cut, n_items, true, ud, c->uvals + (size_t)s0 * n_cut * MC_NCOL, &b, remove_seen,
                         c->st_fin));
        if (inline_sums) RC(accumulate_users(c, n, n_cut, (size_t)s0, c->st_fin));
        CU(cudaEventRecord(b.ev_fin, c->st_fin));
      }
    }
    CU(cudaEventRecord(c->ev_pipe_done, c->st_fin));
    CU(cudaStreamWaitEvent(c->st, c->ev_pipe_done, 0));
  } else {
    RC(prepare_item_factors(c));
    for (int s = 0; s < n_users; s += block) {
      const int n = std::min(block, n_users - s);
      const int* ud = c->eval_users + s;
      RC(score_block(c, n, ud));
      RC(mask_and_topk(c, n, n_items, remove_seen, K, ud));
      RC(metrics_block(c, n, K, n_cut, n_items, true, ud, c->uvals + (size_t)s * n_cut * MC_NCOL));
    }
    if (inline_sums && n_users > 0) RC(accumulate_users(c, n_users, n_cut));
  }
  c->ev_pending_users = n_users; c->ev_pending_ncut = n_cut;
  c->ev_sums_done = inline_sums;
  return 0;
}

int ganmf_evaluate_sums(ganmf_ctx* c, const double* carry_in, double* sums_host, int64_t* counts_host) {
  if (!c || !sums_host) return fail("bad argument");
  if (c->ev_pending_users < 0) return fail("ganmf_evaluate_values first");
  const int n_users = c->ev_pending_users, n_cut = c->ev_pending_ncut;
  const int n_items = n_items_of(c);
  if (c->ev_sums_done) {
    if (carry_in) return fail("the running sums of this evaluation were already formed from zero");
  } else {
    if (carry_in) CU(cudaMemcpyAsync(c->usums, carry_in, (size_t)n_cut * MC_NCOL * 8, cudaMemcpyHostToDevice, c->st));
    else CU(cudaMemsetAsync(c->usums, 0, (size_t)n_cut * MC_NCOL * 8, c->st));
    if (n_users > 0) RC(accumulate_users(c, n_users, n_cut));
  }
  CU(cudaMemcpyAsync(sums_host, c->usums, (size_t)n_cut * MC_NCOL * 8, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (counts_host) {
    std::vector<int> tmp((size_t)n_cut * n_items);
    CU(cudaMemcpy(tmp.data(), c->icounts, tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) counts_host[i] = tmp[i];
  }
  c->ev_pending_users = -1;
  return 0;
}

int ganmf_eval_begin(ganmf_ctx* c, int n_users_total, const int32_t* cutoffs, int n_cut) {
  if (!c || !cutoffs || n_users_total < 0) return fail("bad argument");
  int K;
  RC(eval_prologue(c, cutoffs, n_cut, &K));
  const int n_items = c->csr[GANMF_CSR_TEST].n_cols;
  if (!c->csr[GANMF_CSR_TEST].data) return fail("test CSR needs ratings (data) for RMSE");
  if (c->eval_users_cap < n_users_total) {
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)std::max(n_users_total, 1)));
    c->eval_users_cap = std::max(n_users_total, 1);
  }
  const size_t uv = (size_t)std::max(n_users_total, 1) * n_cut * MC_NCOL;
  if (uv > c->uvals_cap) {
    cudaFree(c->uvals);
    RC(dalloc(&c->uvals, uv));
    c->uvals_cap = uv;
  }
  const size_t ic = (size_t)n_cut * n_items;
  if (ic > c->icounts_cap) {
    cudaFree(c->icounts); cudaFree(c->usums); cudaFree(c->cut_dev);
    RC(dalloc(&c->icounts, ic));
    RC(dalloc(&c->usums, (size_t)64 * MC_NCOL));
    RC(dalloc(&c->cut_dev, (size_t)64));
    c->icounts_cap = ic;
  }
  CU(cudaMemcpyAsync(c->cut_dev, cutoffs, (size_t)n_cut * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->usums, 0, (size_t)n_cut * MC_NCOL * 8, c->st));
  CU(cudaMemsetAsync(c->icounts, 0, ic * 4, c->st));
  c->ev_total = n_users_total; c->ev_done = 0; c->ev_ncut = n_cut; c->ev_K = K;
  return 0;
}

int ganmf_eval_scores_block(ganmf_ctx* c, float* scores_host, const int32_t* users, int n, int remove_seen,
                            int write_back) {
  if (!c || !scores_host || !users || n <= 0) return fail("bad argument");
  if (c->ev_ncut == 0) return fail("ganmf_eval_begin first");
  if (c->ev_done + n > c->ev_total) return fail("more users than announced to ganmf_eval_begin");
  const int n_items = c->csr[GANMF_CSR_TEST].n_cols, ild = rup(n_items, 32);
  const int nu = c->csr[GANMF_CSR_TEST].n_rows;
  for (int i = 0; i < n; ++i)
    if (users[i] < 0 || users[i] >= nu) return fail("user id %d outside [0, %d)", users[i], nu);
  const size_t need = (size_t)n * ild;
  if (need > c->scores_elems) {
    cudaFree(c->scores);
    RC(dalloc(&c->scores, need));
    c->scores_elems = need;
  }
  if ((size_t)n * c->ev_K > c->topk_cap) {
    cudaFree(c->topk_idx); cudaFree(c->topk_val);
    RC(dalloc(&c->topk_idx, (size_t)n * c->ev_K));
    RC(dalloc(&c->topk_val, (size_t)n * c->ev_K));
    c->topk_cap = (size_t)n * c->ev_K;
  }
  int* ud = c->eval_users + c->ev_done;
  CU(cudaMemcpyAsync(ud, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemcpy2DAsync(c->scores, (size_t)ild * 4, scores_host, (size_t)n_items * 4, (size_t)n_items * 4, n,
                       cudaMemcpyHostToDevice, c->st));
  RC(mask_and_topk(c, n, n_items, remove_seen, c->ev_K, ud));
  RC(metrics_block(c, n, c->ev_K, c->ev_ncut, n_items, true, ud,
                   c->uvals + (size_t)c->ev_done * c->ev_ncut * MC_NCOL));
  if (write_back)
    CU(cudaMemcpy2DAsync(scores_host, (size_t)n_items * 4, c->scores, (size_t)ild * 4, (size_t)n_items * 4, n,
                         cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));          // the caller reuses / frees its score block next
  c->ev_done += n;
  return 0;
}

int ganmf_eval_end(ganmf_ctx* c, double* sums_host, int64_t* counts_host) {
  if (!c || !sums_host) return fail("bad argument");
  if (c->ev_ncut == 0) return fail("ganmf_eval_begin first");
  const int n_items = c->csr[GANMF_CSR_TEST].n_cols;
  if (c->ev_done > 0) RC(accumulate_users(c, c->ev_done, c->ev_ncut));
  CU(cudaMemcpyAsync(sums_host, c->usums, (size_t)c->ev_ncut * MC_NCOL * 8, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (counts_host) {
    std::vector<int> tmp((size_t)c->ev_ncut * n_items);
    CU(cudaMemcpy(tmp.data(), c->icounts, tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) counts_host[i] = tmp[i];
  }
  c->ev_ncut = 0;
  return 0;
}

int ganmf_metrics_from_topk(ganmf_ctx* c, const int32_t* topk, int K, const int32_t* users, int n,
                            const int32_t* cutoffs, int n_cut, double* per_user, double* sums_host,
                            int64_t* counts_host) {
  if (!c || !topk || !users || !cutoffs || !sums_host || n <= 0) return fail("bad argument");
  int Kmax;
  RC(eval_prologue(c, cutoffs, n_cut, &Kmax));
  if (K < Kmax) return fail("lists have K=%d < largest cutoff %d", K, Kmax);
  const int n_items = c->csr[GANMF_CSR_TEST].n_cols;
  // buffers sized without touching the model's score workspace
  if (c->eval_users_cap < n) {
    cudaFree(c->eval_users);
    RC(dalloc(&c->eval_users, (size_t)n));
    c->eval_users_cap = n;
  }
  if ((size_t)n * K > c->topk_cap) {
    cudaFree(c->topk_idx); cudaFree(c->topk_val);
    RC(dalloc(&c->topk_idx, (size_t)n * K));
    RC(dalloc(&c->topk_val, (size_t)n * K));
    c->topk_cap = (size_t)n * K;
  }
  const size_t uv = (size_t)n * n_cut * MC_NCOL;
  if (uv > c->uvals_cap) {
    cudaFree(c->uvals);
    RC(dalloc(&c->uvals, uv));
    c->uvals_cap = uv;
  }
  const size_t ic = (size_t)n_cut * n_items;
  if (ic > c->icounts_cap) {
    cudaFree(c->icounts); cudaFree(c->usums); cudaFree(c->cut_dev);
    RC(dalloc(&c->icounts, ic));
    RC(dalloc(&c->usums, (size_t)64 * MC_NCOL));
    RC(dalloc(&c->cut_dev, (size_t)64));
    c->icounts_cap = ic;
  }
  CU(cudaMemcpyAsync(c->cut_dev, cutoffs, (size_t)n_cut * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemcpyAsync(c->eval_users, users, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemcpyAsync(c->topk_idx, topk, (size_t)n * K * 4, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->usums, 0, (size_t)n_cut * MC_NCOL * 8, c->st));
  CU(cudaMemsetAsync(c->icounts, 0, ic * 4, c->st));
  CU(cudaMemsetAsync(c->uvals, 0, uv * 8, c->st));
  RC(metrics_block(c, n, K, n_cut, n_items, false, c->eval_users, c->uvals));
  RC(accumulate_users(c, n, n_cut));
  CU(cudaMemcpyAsync(sums_host, c->usums, (size_t)n_cut * MC_NCOL * 8, cudaMemcpyDeviceToHost, c->st));
  if (per_user) CU(cudaMemcpyAsync(per_user, c->uvals, uv * 8, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (counts_host) {
    std::vector<int> tmp(ic);
    CU(cudaMemcpy(tmp.data(), c->icounts, ic * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ic; ++i) counts_host[i] = tmp[i];
  }
  return 0;
}

// ------------------------------------------------------------------------------ primitives
int ganmf_k_gemm(ganmf_ctx* c, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N,
                 int K, float* out, int ldo, int path) {
  if (!c) return fail("null ctx");
  if (path == GANMF_GEMM_RESIDENT_A) {
    GenGemmCall g;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.out = out; g.ldo = ldo;
    g.cache = &c->tmaps; g.max_ctas = c->gemm_sm_cap;
    if (a_mn || b_mn || !resident_a_gemm_ok(g))
      return fail("resident-A GEMM: K-major operands, K <= 256, ldo >= roundup(N, 32), 16-byte aligned");
    return gen_gemm(c, g);
  }
  Epilogue e;
  e.out = out; e.ldo = ldo;
  return gemm(c, A, lda, a_mn, B, ldb, b_mn, M, N, K, e, path);
}
int ganmf_k_csr_gather_dense(ganmf_ctx* c, int ids_offset, int B, float* out, int ld) {
  if (!c) return fail("null ctx");
  const Csr& tr = c->csr[GANMF_CSR_TRAIN];
  if (!tr.indptr) return fail("train CSR not set");
  if (ld < c->W || (ld & 3)) return fail("ld must be >= width and a multiple of 4");
  c->launches++;
  CU(csr_gather_dense(tr.indptr, tr.indices, tr.data, c->ids + ids_offset, B, out, ld, 0, c->st));
  return 0;
}
int ganmf_k_csr_encode_rows(ganmf_ctx* c, int ids_offset, int B, float* out, int ldo) {
  if (!c) return fail("null ctx");
  if (c->cfg.kind != GANMF_KIND_GANMF) return fail("csr_encode_rows: GANMF only");
  const Csr& tr = c->csr[GANMF_CSR_TRAIN];
  if (!tr.indptr) return fail("train CSR not set");
  if (ldo < c->Ep || (ldo & 3)) return fail("ldo must be >= roundup(emb_dim, 32) and a multiple of 4");
  if (B <= 0 || ids_offset < 0 || ids_offset + B > c->ids_cap) return fail("ids range out of bounds");
  const Param *We = &c->params[0], *be = &c->params[1];
  c->launches++;
  CU(csr_encode_rows(tr.indptr, tr.indices, tr.data, c->ids + ids_offset, B, We->w.p, We->w.ld, c->Ep, be->w.p, out,
                     ldo, c->st));
  return 0;
}
int ganmf_step_routes(ganmf_ctx* c, int32_t* sparse_real, int32_t* bias_grad_from_gemm, int32_t* lowrank_fake) {
  if (!c) return fail("null ctx");
  if (sparse_real) *sparse_real = c->sparse_real ? 1 : 0;
  if (bias_grad_from_gemm) *bias_grad_from_gemm = (c->colpart_on && c->colpart) ? 1 : 0;
  if (lowrank_fake) *lowrank_fake = c->lowrank ? 1 : 0;
  return 0;
}
int ganmf_k_adam(ganmf_ctx* c, float* theta, float* m, float* v, const float* g, int64_t n, float alpha,
                 float reg) {
  if (!c || (n & 3)) return fail("n must be a multiple of 4");
  AdamArgs a;
  memset(&a, 0, sizeof a);
  a.nseg = 1;
  a.seg[0].theta = theta; a.seg[0].m = m; a.seg[0].v = v; a.seg[0].g = g; a.seg[0].n4 = (unsigned long long)n / 4;
  a.seg[0].ld = 4;
  a.alpha = alpha; a.reg = reg; a.l2_out = nullptr;
  c->launches++;
  CU(fused_adam(a, c->st));
  return 0;
}
int ganmf_k_topk(ganmf_ctx* c, const float* scores, int ld, int n, int n_items, int K, int32_t* idx, float* val) {
  if (!c || K < 1 || K > TK_MAXK) return fail("bad K");
  CU(topk_rows(scores, ld, n, n_items, K, idx, val, c->st));
  c->launches++;
  return 0;
}
