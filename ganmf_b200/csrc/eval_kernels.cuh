// On-device evaluator kernels (sm_100a), HBM-bound:
//   K8  mask_seen + topk_rows   scores[u, seen(u)] = -inf, then per-row top-K by descending score,
//                               ties -> lowest item index   (BaseRecommender.py:189-194,214-234)
//   K9  user_metrics            per (user, cutoff) metric values in the reference's own arithmetic
//                               (metrics.py:576-722) + per-item recommendation histograms
//   --  ordered_accumulate      running float64 sums in USER ORDER, as Evaluator.py:305-335 does
// Bit-exactness notes: per-list sums follow numpy's pairwise summation (n <= 128: eight running
// accumulators, then the tail sequentially), ln(j+2) comes from a host table made with numpy, and
// the per-user dtypes are those of the numpy the reference pins (see oracle/eval_oracle.py).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ganmf {

__global__ void mask_seen_kernel(float* __restrict__ scores, int ld, const int* __restrict__ users,
                                 const int* __restrict__ indptr, const int* __restrict__ indices) {
  const int r = blockIdx.x;
  const int u = users[r];
  const int s = indptr[u], e = indptr[u + 1];
  float* row = scores + (size_t)r * ld;
  for (int i = s + threadIdx.x; i < e; i += blockDim.x) row[indices[i]] = -INFINITY;
}

// ----------------------------------------------------------------------------- top-K
// Total order key: higher score first, then lower index.  NaN sorts last (numpy argsort).
__device__ __forceinline__ unsigned long long topk_key(float f, int idx) {
  unsigned u = f == 0.f ? 0u : __float_as_uint(f);          // -0.0 == +0.0, as in numpy's comparison
  u = (f != f) ? 0u : (u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ float topk_key_score(unsigned long long k) {
  unsigned u = (unsigned)(k >> 32);
  u = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(u);
}

constexpr int TK_WARPS = 8;
constexpr int TK_THREADS = TK_WARPS * 32;
constexpr int TK_MAXK = 128;

// Warp-level selection: the warp keeps its running top-K as ONE sorted list spread over the lanes'
// registers (lane l holds ranks l*J .. l*J+J-1, J = ceil(K/32)).  Streaming a row, an element is
// looked at in detail only if its score bits reach the current K-th best; after the first few tiles
// that is rare (about K*ln(n/K) insertions per warp), so the kernel runs at the speed of the loads.
template <int J>
__device__ __forceinline__ void tk_insert(unsigned long long (&lst)[J], unsigned long long x, int lane) {
  unsigned long long left_last = __shfl_up_sync(0xffffffffu, lst[J - 1], 1);
  if (lane == 0) left_last = ~0ull;
#pragma unroll
  for (int j = J - 1; j >= 0; --j) {
    const unsigned long long left = j == 0 ? left_last : lst[j - 1];
    lst[j] = lst[j] > x ? lst[j] : (left > x ? x : left);
  }
}
template <int J>
__device__ __forceinline__ unsigned long long tk_kth(const unsigned long long (&lst)[J], int K) {
  const int slot = (K - 1) % J, owner = (K - 1) / J;
  unsigned long long v = lst[0];
#pragma unroll
  for (int j = 1; j < J; ++j) v = slot == j ? lst[j] : v;
  return __shfl_sync(0xffffffffu, v, owner);
}

// One CTA per score row; its 1..8 warps stream interleaved 512-byte tiles with 16-byte loads and the
// per-warp winners are merged by rank.  Fewer warps per row mean fewer (latency-bound) insertions per
// row, so the launcher uses as few as still give every SM ~32 warps.
// Algorithmic bytes: 4*n_items read + 8*K written per row.
template <int J>
__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float* __restrict__ scores, int ld, int n_items, int K, int* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  __shared__ unsigned long long sbuf[TK_WARPS][32 * J];
  __shared__ int scnt[TK_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;                       // warps working on this row
  const float* row = scores + (size_t)blockIdx.x * ld;
  unsigned long long lst[J];
#pragma unroll
  for (int j = 0; j < J; ++j) lst[j] = 0ull;
  unsigned long long thr = 0ull;
  float thr_f = -INFINITY;                              // score of the current K-th best
  const int n_tiles = (n_items + 127) / 128;
  const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  auto load_tile = [&](int t) -> float4 {
    const int c = t * 128 + lane * 4;
    if (t >= n_tiles) return ninf4;
    if (c + 3 < n_items) return *reinterpret_cast<const float4*>(row + c);
    float4 v = ninf4;
    if (c < n_items) v.x = row[c];
    if (c + 1 < n_items) v.y = row[c + 1];
    if (c + 2 < n_items) v.z = row[c + 2];
    return v;
  };
  // The warp walks its tiles (warp, warp+8, ...) in batches of TK_DEPTH with the next batch already in
  // flight: 2*TK_DEPTH 16-byte loads per lane outstanding keep the row streaming at HBM speed.
  constexpr int TK_DEPTH = 4;
  float4 nxt[TK_DEPTH];
#pragma unroll
  for (int d = 0; d < TK_DEPTH; ++d) nxt[d] = load_tile(warp + d * nw);
  for (int t0 = warp; t0 < n_tiles; t0 += nw * TK_DEPTH) {
    float4 cur[TK_DEPTH];
#pragma unroll
    for (int d = 0; d < TK_DEPTH; ++d) cur[d] = nxt[d];
#pragma unroll
    for (int d = 0; d < TK_DEPTH; ++d) nxt[d] = load_tile(t0 + (TK_DEPTH + d) * nw);
#pragma unroll
    for (int d = 0; d < TK_DEPTH; ++d) {
      const int t = t0 + d * nw;
      if (t >= n_tiles) break;                          // warp-uniform
      const float4 v = cur[d];
      // fast path (8 instructions per 16 bytes): plain float compares against the score of the K-th
      // best.  NaN never passes, out-of-range lanes hold -inf; while the list is not full thr_f = -inf
      // sends everything to the exact path below.
      const bool any = (v.x >= thr_f) | (v.y >= thr_f) | (v.z >= thr_f) | (v.w >= thr_f);
      if (!__any_sync(0xffffffffu, any)) continue;
      const int c = t * 128 + lane * 4;
      const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned long long key = topk_key(vs[e], c + e);
        bool pass = (c + e < n_items) && key > thr;
        unsigned bal;
        while ((bal = __ballot_sync(0xffffffffu, pass)) != 0u) {
          const int src = __ffs(bal) - 1;
          const unsigned long long x = __shfl_sync(0xffffffffu, key, src);
          tk_insert<J>(lst, x, lane);
          thr = tk_kth<J>(lst, K);
          thr_f = (thr >> 32) ? topk_key_score(thr) : -INFINITY;
          if (lane == src) pass = false;
          pass = pass && key > thr;
        }
      }
    }
  }
  // per-warp winners (ranks < K, zeros = empty) -> smem, then merge by rank: keys are unique (they embed
  // the item index), so a key's final position is its own rank plus the number of larger keys elsewhere
  int mine = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = lane * J + j;
    sbuf[warp][p] = lst[j];
    mine += (p < K && lst[j] != 0ull) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0) scnt[warp] = mine;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < nw; ++w) total += scnt[w];
  for (int e = threadIdx.x; e < nw * K; e += blockDim.x) {
    const int w = e / K, p = e - w * K;
    if (p >= scnt[w]) continue;
    const unsigned long long key = sbuf[w][p];
    int rank = p;
    for (int w2 = 0; w2 < nw; ++w2) {
      if (w2 == w) continue;
      int lo = 0, hi2 = scnt[w2];                       // first position in list w2 with key' < key
      while (lo < hi2) {
        const int mid = (lo + hi2) >> 1;
        if (sbuf[w2][mid] > key) lo = mid + 1; else hi2 = mid;
      }
      rank += lo;
    }
    if (rank < K) {
      const float sc = topk_key_score(key);
      out_idx[(size_t)blockIdx.x * K + rank] =
          sc != -INFINITY ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : -1;
      out_val[(size_t)blockIdx.x * K + rank] = sc;
    }
  }
  for (int j = total + threadIdx.x; j < K; j += blockDim.x) {
    out_idx[(size_t)blockIdx.x * K + j] = -1;
    out_val[(size_t)blockIdx.x * K + j] = -INFINITY;
  }
}

inline cudaError_t topk_rows(const float* scores, int ld, int n, int n_items, int K, int* out_idx, float* out_val,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int wpr = 8;                                          // warps per row: keep >= ~16 warps per SM
  while (wpr > 1 && (long long)n * (wpr / 2) >= 148LL * 16) wpr >>= 1;
  while (wpr > 1 && (n_items + 127) / 128 < wpr * 4) wpr >>= 1;   // short rows: not worth splitting
  const int threads = 32 * wpr;
  if (K <= 32) topk_rows_kernel<1><<<n, threads, 0, st>>>(scores, ld, n_items, K, out_idx, out_val);
  else if (K <= 64) topk_rows_kernel<2><<<n, threads, 0, st>>>(scores, ld, n_items, K, out_idx, out_val);
  else topk_rows_kernel<4><<<n, threads, 0, st>>>(scores, ld, n_items, K, out_idx, out_val);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------- metrics
enum MetricCol {
  MC_PRECISION = 0, MC_RECALL, MC_PRMD, MC_MAP, MC_NDCG, MC_MRR, MC_ARHR, MC_ROC_AUC, MC_HIT,
  MC_NOVELTY, MC_AVGPOP, MC_COVERED, MC_RMSE, MC_NCOL
};

// numpy pairwise summation for n <= 128 (verified against np.sum): term(i) is evaluated in order.
template <typename T, typename F>
__device__ __forceinline__ T np_sum(int n, F term) {
  if (n < 8) {
    T res = (T)0;
    for (int i = 0; i < n; ++i) res += term(i);
    return res;
  }
  T r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += term(i + j);
  }
  T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += term(i);
  return res;
}

struct EvalTables {
  const int* test_indptr;        // users x items CSR, indices sorted within a row
  const int* test_indices;
  const float* test_gain;        // 2^rating - 1 (float32, numpy-made), aligned with test_indices
  const float* test_gain_desc;   // the same values sorted descending within each row
  const float* logtab;           // float32 ln(j + 2), j < TK_MAXK (numpy-made)
  const double* item_novelty;    // -log2(pop_i / n_interactions) / n_items   (0 where pop_i == 0)
  const unsigned char* item_has_pop;
  const double* item_popnorm;    // pop_i / max pop
};

// One thread per (row, cutoff).  vals[(row * n_cut + ci) * MC_NCOL + col].
__global__ void user_metrics_kernel(const int* __restrict__ topk_idx, int K, const int* __restrict__ users,
                                    int n_rows, const int* __restrict__ cutoffs, int n_cut, EvalTables tb,
                                    double* __restrict__ vals, int* __restrict__ item_counts, int n_items) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_rows * n_cut) return;
  const int r = gid / n_cut, ci = gid % n_cut;
  const int c = cutoffs[ci];
  const int u = users[r];
  const int* lst = topk_idx + (size_t)r * K;
  int L = 0;                                   // valid entries form a prefix
  while (L < c && L < K && lst[L] >= 0) ++L;
  const int ts = tb.test_indptr[u], T = tb.test_indptr[u + 1] - ts;
  const int* tidx = tb.test_indices + ts;
  auto find = [&](int item) -> int {           // position in the user's test row or -1
    int lo = 0, hi = T - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int v = tidx[mid];
      if (v == item) return mid;
      if (v < item) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
  };
  // hit mask (<= 128 bits)
  unsigned hm[4] = {0u, 0u, 0u, 0u};
  int h = 0, first = -1;
  for (int j = 0; j < L; ++j)
    if (find(lst[j]) >= 0) { hm[j >> 5] |= 1u << (j & 31); ++h; if (first < 0) first = j; }
  auto hit = [&](int j) -> bool { return (hm[j >> 5] >> (j & 31)) & 1u; };

  double* o = vals + (size_t)gid * MC_NCOL;
  const int minTL = T < L ? T : L;
  o[MC_PRECISION] = L ? (double)h / (double)L : 0.0;
  o[MC_RECALL] = (double)h / (double)T;
  o[MC_PRMD] = L ? (double)h / (double)minTL : 0.0;
  o[MC_HIT] = (double)h;
  o[MC_MRR] = first >= 0 ? 1.0 / (double)(first + 1) : 0.0;
  o[MC_COVERED] = L > 0 ? 1.0 : 0.0;
  // MAP: is_rel * cumsum(f32) / (1 + arange) -> float64 terms, np.sum, / min(T, L)
  if (L) {
    // np_sum does not evaluate term(i) in increasing i, so the cumulative hit count comes from a popcount of
    // the mask prefix, not from a running counter
    auto cum_at = [&](int j) -> int {
      int s = 0;
      for (int w = 0; w < (j >> 5); ++w) s += __popc(hm[w]);
      s += __popc(hm[j >> 5] & (0xFFFFFFFFu >> (31 - (j & 31))));
      return s;
    };
    const double ap = np_sum<double>(L, [&](int j) -> double {
      return hit(j) ? (double)(float)cum_at(j) / (double)(j + 1) : 0.0;
    });
    o[MC_MAP] = ap / (double)minTL;
  } else {
    o[MC_MAP] = 0.0;
  }
  // NDCG (all float32): dcg over the list, idcg over the sorted test gains cut to L
  {
    const float* g = tb.test_gain + ts;
    const float dcg = np_sum<float>(L, [&](int j) -> float {
      const int p = hit(j) ? find(lst[j]) : -1;
      return (p >= 0 ? g[p] : 0.0f) / tb.logtab[j];
    });
    const float* gd = tb.test_gain_desc + ts;
    const float idcg = np_sum<float>(minTL, [&](int j) -> float { return gd[j] / tb.logtab[j]; });
    o[MC_NDCG] = dcg == 0.0f ? 0.0 : (double)(dcg / idcg);
  }
  // ARHR: sum hit_j / (j + 1) in float64 (reference: BLAS ddot; order-insensitive up to the last bit)
  {
    double a = 0.0;
    for (int j = 0; j < L; ++j) if (hit(j)) a += 1.0 / (double)(j + 1);
    o[MC_ARHR] = a;
  }
  // ROC_AUC: correctly ordered (hit, non-hit) pairs / (n_pos * n_neg)
  {
    const int nneg = L - h;
    if (nneg == 0) o[MC_ROC_AUC] = 1.0;
    else if (h == 0) o[MC_ROC_AUC] = 0.0;
    else {
      long long cnt = 0; int neg_after = nneg;
      for (int j = 0; j < L; ++j) { if (hit(j)) cnt += neg_after; else --neg_after; }
      o[MC_ROC_AUC] = (double)cnt / (double)((long long)h * nneg);
    }
  }
  // NOVELTY (cold items dropped before the sum) and AVERAGE_POPULARITY
  if (L) {
    int np_ = 0;
    unsigned pm[4] = {0u, 0u, 0u, 0u};          // positions with pop != 0
    for (int j = 0; j < L; ++j) if (tb.item_has_pop[lst[j]]) { pm[j >> 5] |= 1u << (j & 31); ++np_; }
    auto nth = [&](int i) -> int {               // i-th set position
      int w = 0, rem = i;
      while (rem >= __popc(pm[w])) { rem -= __popc(pm[w]); ++w; }
      unsigned m = pm[w];
      for (int t = 0; t < rem; ++t) m &= m - 1;
      return (w << 5) + __ffs(m) - 1;
    };
    o[MC_NOVELTY] = np_sum<double>(np_, [&](int i) -> double { return tb.item_novelty[lst[nth(i)]]; });
    o[MC_AVGPOP] = np_sum<double>(L, [&](int j) -> double { return tb.item_popnorm[lst[j]]; }) / (double)L;
    int* cnts = item_counts + (size_t)ci * n_items;
    for (int j = 0; j < L; ++j) atomicAdd(cnts + lst[j], 1);
  } else {
    o[MC_NOVELTY] = 0.0;
    o[MC_AVGPOP] = 0.0;
  }
}

// RMSE over the user's test items with a finite (masked) score: float32 numpy sum / count, sqrt in
// float64 (metrics.py:634-659).  One thread per row; numpy's recursive pairwise split for n > 128.
__device__ inline float np_sum_f32_buf(const float* a, int n) {
  if (n <= 128) return np_sum<float>(n, [&](int i) -> float { return a[i]; });
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_sum_f32_buf(a, n2) + np_sum_f32_buf(a + n2, n - n2);
}
__global__ void user_rmse_kernel(const float* __restrict__ scores, int ld, const int* __restrict__ users,
                                 int n_rows, int n_cut, EvalTables tb, const float* __restrict__ test_data,
                                 float* __restrict__ scratch, double* __restrict__ vals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int u = users[r];
  const int ts = tb.test_indptr[u], T = tb.test_indptr[u + 1] - ts;
  float* e = scratch + ts;                       // per-test-entry scratch (compacted finite errors)
  int n = 0;
  for (int i = 0; i < T; ++i) {
    const float d = scores[(size_t)r * ld + tb.test_indices[ts + i]] - test_data[ts + i];
    const float sq = d * d;
    if (isfinite(sq)) e[n++] = sq;
  }
  const double v = n ? sqrt((double)np_sum_f32_buf(e, n) / (double)n) : nan("");
  for (int ci = 0; ci < n_cut; ++ci) vals[((size_t)r * n_cut + ci) * MC_NCOL + MC_RMSE] = v;
}

// sums[col] += vals[row][col] for rows IN ORDER (bit-exact vs the reference's running sum, which is one
// Python float per metric updated user after user).  Single CTA: the row-major value table is streamed
// through shared memory in contiguous tiles (all threads load, coalesced, next tile in flight in
// registers), and one thread per (cutoff, metric) column performs the strictly sequential adds.
constexpr int OA_THREADS = 256;
constexpr int OA_TILE = 2048;                 // doubles per tile (16 KB), two buffers
__global__ void __launch_bounds__(OA_THREADS)
ordered_accumulate_kernel(const double* __restrict__ vals, int n_rows, int n_cols, double* __restrict__ sums) {
  __shared__ double buf[2][OA_TILE];
  const int rows_per_tile = OA_TILE / n_cols;                 // host guarantees n_cols <= OA_TILE
  const int n_tiles = (n_rows + rows_per_tile - 1) / rows_per_tile;
  const size_t total = (size_t)n_rows * n_cols;
  constexpr int PER = OA_TILE / OA_THREADS;
  double pre[PER];
  auto fetch = [&](int t) {
    const size_t base = (size_t)t * rows_per_tile * n_cols;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const size_t i = base + threadIdx.x + (size_t)j * OA_THREADS;
      pre[j] = (threadIdx.x + j * OA_THREADS < rows_per_tile * n_cols && i < total) ? vals[i] : 0.0;
    }
  };
  auto stash = [&](int b) {
#pragma unroll
    for (int j = 0; j < PER; ++j) buf[b][threadIdx.x + j * OA_THREADS] = pre[j];
  };
  double acc = threadIdx.x < n_cols ? sums[threadIdx.x] : 0.0;
  if (n_tiles > 0) { fetch(0); stash(0); }
  __syncthreads();
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) fetch(t + 1);                        // global loads in flight during the adds
    if (threadIdx.x < n_cols) {
      const int r0 = t * rows_per_tile;
      const int nr = min(rows_per_tile, n_rows - r0);
      const double* b = buf[t & 1] + threadIdx.x;
      int r = 0;
      for (; r + 8 <= nr; r += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = b[(r + j) * n_cols];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
      }
      for (; r < nr; ++r) acc += b[r * n_cols];
    }
    if (t + 1 < n_tiles) stash((t + 1) & 1);
    __syncthreads();
  }
  if (threadIdx.x < n_cols) sums[threadIdx.x] = acc;
}

}  // namespace ganmf
