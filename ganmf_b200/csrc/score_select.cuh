// Fused scoring + seen mask + top-K selection for the evaluator / recommend() (sm_100a):
//   BaseRecommender.recommend (Base/BaseRecommender.py:189-234) on scores = U . V^T (GANMF.py:285-292)
// without ever writing the n x n_items score matrix to HBM.
//
//   1. score_select_kernel   one TF32 tcgen05 pass over all items.  CTA pairs (cta_group::2): a pair owns a
//      block of 256 query rows whose factor tile stays RESIDENT in shared memory (k <= 256: 8 k-blocks of 16 KB
//      per CTA) while the item factors stream through a 4-stage TMA ring (each CTA loads half of every 256-item
//      tile), accumulators double-buffered in TMEM.  The epilogue never stores scores: thread = query row
//      (tcgen05.ld 32x32b), it drops the row's seen items (CSR cursor), compares the 32 scores of a chunk with
//      the row's running KP-th best and inserts the rare survivors into a sorted KP-entry list held in
//      REGISTERS.  Per (row, column half, item segment) one list of KP candidates leaves the kernel.
//   2. rescore_kernel        one warp per row: the candidates that can still reach the top K are re-scored
//      exactly -- fl32(sum_k fp64(p_k * v_k)), the correctly rounded fp32 score -- ranked by (score desc, item
//      asc), and the result is CERTIFIED: every item outside the candidate lists has a TF32 score below its list's
//      final threshold tau, and |TF32 score - exact| <= eps(row), so if the K-th exact score exceeds
//      max(tau) + eps the list equals the top K of the exact score row.  Rows that cannot be certified are
//      handed to the exact fallback (exact_score_rows_kernel + the materialised mask/top-k kernels).
//
// Algorithmic work: 2*I*k FLOP per row on the tensor cores (a third of the split-TF32 scorer), O(I) compares per
// row in the epilogue, no score traffic.  Reported against the 4*I bytes/row a materialised top-k would read
// ("effective" figure).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "eval_kernels.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"

namespace ganmf {

constexpr int SS_ROWS = 256;           // query rows per CTA pair
constexpr int SS_BN = 256;             // items per tile
constexpr int SS_STAGES = 4;           // B ring depth (16 KB per stage per CTA)
constexpr int SS_MAX_KB = 8;           // resident A: up to 8 k-blocks of 32 (k <= 256)
constexpr int SS_MAX_LISTS = 16;       // lists per row = 2 column halves x item segments

struct SelArgs {
  int n_rows, n_items, nkb;
  int tiles_total;                     // item tiles of SS_BN
  int segs, tiles_per_seg;
  int row_blocks;
  uint32_t idesc;
  const int* users;                    // row -> user id (seen CSR row); nullptr: row itself
  const int* seen_indptr;              // nullptr: nothing is masked
  const int* seen_indices;
  float* cand_val;                     // [n_rows][2*segs][KP]
  int* cand_idx;
  unsigned int* row_thr;               // [n_rows] shared rejection threshold of a row (monotone key, 0 = none)
};

// monotone float <-> uint key (larger float = larger key; key 0 is below every float incl. -inf)
__device__ __forceinline__ unsigned int thr_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float thr_val(unsigned int k) {
  if (k == 0u) return -INFINITY;
  return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k);
}

template <int KP>
struct SelSmem {
  static constexpr int A_KB_BYTES = TC_BM * TC_BK * 4;                 // 16 KB: 128 rows x 32 k
  static constexpr int B_BYTES = (SS_BN / 2) * TC_BK * 4;              // 16 KB: 128 items x 32 k
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = SS_MAX_KB * A_KB_BYTES;
  static constexpr int BAR_OFF = B_OFF + SS_STAGES * B_BYTES;
  static constexpr int VAL_OFF = BAR_OFF + (2 * SS_STAGES + 6) * 8 + 16;       // slow-path staging: 32 scores x 256 threads
  static constexpr int TOTAL = VAL_OFF + 32 * 256 * 4 + 1024;
};

// sorted (descending) insertion of (v, idx) into the KP-entry register list; v > sv[KP-1] on entry
template <int KP>
__device__ __forceinline__ void sel_insert(float (&sv)[KP], int (&si)[KP], float v, int idx) {
#pragma unroll
  for (int j = KP - 1; j >= 1; --j) {
    const bool here = v > sv[j];               // the new entry ranks at or above position j
    const bool above = v > sv[j - 1];          // ... and above position j-1: the old j-1 moves down
    sv[j] = here ? (above ? sv[j - 1] : v) : sv[j];
    si[j] = here ? (above ? si[j - 1] : idx) : si[j];
  }
  if (v > sv[0]) { sv[0] = v; si[0] = idx; }
}

// 32 lanes x 32 consecutive fp32 columns into r[OFF .. OFF+32) of a larger register array
template <int OFF, int N>
__device__ __forceinline__ void tmem_ld_32x32_at(uint32_t taddr, uint32_t (&r)[N]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]), "=r"(r[OFF + 17]),
        "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]), "=r"(r[OFF + 22]), "=r"(r[OFF + 23]),
        "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]), "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]),
        "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
      : "r"(taddr)
      : "memory");
}

template <int KP>
__global__ void __launch_bounds__(TC_THREADS, 1)
score_select_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const SelArgs args) {
  using S = SelSmem<KP>;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader: issues the pair's MMAs
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + SS_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SS_STAGES;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint64_t* a_full_bar = tmem_empty_bar + 2;             // resident A tile of the unit has landed
  uint64_t* a_empty_bar = a_full_bar + 1;                // every MMA of the unit has retired: A may be replaced
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = args.row_blocks * args.segs;
  const int n_pairs = (int)gridDim.x / 2;
  const int pair = (int)blockIdx.x / 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < SS_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full_bar[a], 1); ptx::mbar_init(&tmem_empty_bar[a], 16); }
    ptx::mbar_init(a_full_bar, 1);
    ptx::mbar_init(a_empty_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // units are walked statically; consecutive units share the item segment, so the pairs running side by side
  // stream the same item tiles (L2 hits) on different row blocks
  auto unit_rb = [&](int u) { return u % args.row_blocks; };
  auto unit_seg = [&](int u) { return u / args.row_blocks; };
  auto seg_tiles = [&](int seg, int& t0, int& t1) {
    t0 = seg * args.tiles_per_seg;
    t1 = min(t0 + args.tiles_per_seg, args.tiles_total);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t it = 0, ui = 0;
      for (int u = pair; u < n_units; u += n_pairs, ++ui) {
        int t0, t1;
        seg_tiles(unit_seg(u), t0, t1);
        const int m0 = unit_rb(u) * SS_ROWS + (int)rank * TC_BM;
        ptx::mbar_wait(a_empty_bar, (ui & 1) ^ 1);
        const uint32_t afb = ptx::mapa(ptx::smem_u32(a_full_bar), 0u);
        if (rank == 0) ptx::mbar_expect_tx(a_full_bar, 2 * args.nkb * S::A_KB_BYTES);
        for (int kb = 0; kb < args.nkb; ++kb)
          ptx::tma_load_2d_pair(smem + S::A_OFF + kb * S::A_KB_BYTES, &map_a, afb, kb * TC_BK, m0);
        for (int t = t0; t < t1; ++t) {
          const int n0 = t * SS_BN + (int)rank * (SS_BN / 2);
          for (int kb = 0; kb < args.nkb; ++kb, ++it) {
            const int s = it % SS_STAGES;
            ptx::mbar_wait(&empty_bar[s], ((it / SS_STAGES) & 1) ^ 1);
            const uint32_t fb = ptx::mapa(ptx::smem_u32(&full_bar[s]), 0u);
            if (rank == 0) ptx::mbar_expect_tx(&full_bar[s], 2 * S::B_BYTES);
            ptx::tma_load_2d_pair(smem + S::B_OFF + s * S::B_BYTES, &map_b, fb, kb * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && ptx::elect_one()) {
      uint32_t it = 0, ti = 0, ui = 0;
      for (int u = pair; u < n_units; u += n_pairs, ++ui) {
        int t0, t1;
        seg_tiles(unit_seg(u), t0, t1);
        ptx::mbar_wait(a_full_bar, ui & 1);
        for (int t = t0; t < t1; ++t, ++ti) {
          const uint32_t acc = ti & 1;
          ptx::mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * SS_BN;
          for (int kb = 0; kb < args.nkb; ++kb, ++it) {
            const int s = it % SS_STAGES;
            ptx::mbar_wait(&full_bar[s], (it / SS_STAGES) & 1);
            ptx::tc_fence_after();
            const uint64_t da = make_smem_desc(ptx::smem_u32(smem + S::A_OFF + kb * S::A_KB_BYTES), 1, 1024 >> 4, 2);
            const uint64_t db = make_smem_desc(ptx::smem_u32(smem + S::B_OFF + s * S::B_BYTES), 1, 1024 >> 4, 2);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              ptx::mma_tf32_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), args.idesc, (kb | k) ? 1u : 0u);
            ptx::mma_commit_pair(&empty_bar[s], 3);
          }
          ptx::mma_commit_pair(&tmem_full_bar[acc], 3);
        }
        ptx::mma_commit_pair(a_empty_bar, 3);
      }
    }
  } else {
    // ------------------------------------------------------------ selection epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter
    const int half = ew >> 2;                      // column half of every tile
    const int NL = 2 * args.segs;
    uint32_t ti = 0;
    for (int u = pair; u < n_units; u += n_pairs) {
      int t0, t1;
      const int seg = unit_seg(u);
      seg_tiles(seg, t0, t1);
      const int row = unit_rb(u) * SS_ROWS + (int)rank * TC_BM + q * 32 + lane;
      const bool live = row < args.n_rows;
      float sv[KP];
      int si[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) { sv[j] = -INFINITY; si[j] = -1; }
      // Rejection threshold shared by ALL lists of the row (both column halves, every item segment, whichever
      // SM runs them): each list publishes its KP-th best with an atomic max, every list rejects what does not
      // beat the published value.  Any value ever published is the KP-th best of SOME subset of the row's
      // rankable items, so no member of the row's top KP (over all items) is ever rejected, and everything that
      // is rejected or evicted scores at most the final published value -- the tau of the certificate.  A row's
      // insertions drop from (lists x KP ln(n/KP)) to about KP ln(I/KP) in total.
      unsigned int* thr_slot = args.row_thr + (live ? row : 0);
      float thr_ext = live ? thr_val(__ldcg(thr_slot)) : INFINITY;
      float thr_pub = thr_ext;
      // seen-item cursor of this (row, half): entries are consumed in increasing column order, the next one is
      // always already in a register
      int sp = 0, se = 0, ns0 = 0x7fffffff, ns1 = 0x7fffffff;
      if (live && args.seen_indptr) {
        const int uid = args.users ? __ldg(args.users + row) : row;
        sp = __ldg(args.seen_indptr + uid);
        se = __ldg(args.seen_indptr + uid + 1);
        const int c_first = t0 * SS_BN + half * (SS_BN / 2);
        int lo = sp, hi = se;                                   // lower_bound(c_first)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(args.seen_indices + mid) < c_first) lo = mid + 1; else hi = mid;
        }
        sp = lo;
        if (sp < se) ns0 = __ldg(args.seen_indices + sp);
        if (sp + 1 < se) ns1 = __ldg(args.seen_indices + sp + 1);
      }
      // GROUP chunks (32 columns each) are pulled out of TMEM at a time; with GROUP = 4 the whole 128-column share of
      // this warp sits in registers and the accumulator stage is handed back to the MMA warp BEFORE any selection
      // work: the tensor cores never wait for a warp that happens to have insertions to do (ncu: with the release
      // after the last chunk's load only, the MMA warp spent most of its time on tmem_empty -- the slowest of the
      // pair's 16 epilogue warps sets the pace -- and the tensor pipe idled half the time).
      constexpr int GROUP = KP <= 16 ? 4 : 2;
      float* sval = reinterpret_cast<float*>(smem + S::VAL_OFF) + (threadIdx.x - 64);   // [32][256]: slow-path staging
      for (int t = t0; t < t1; ++t, ++ti) {
        const uint32_t acc = ti & 1;
        // what the other lists of this row have published meanwhile (the load is consumed a tile later)
        const unsigned int thr_seen = live ? __ldcg(thr_slot) : 0u;
        ptx::mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int g = 0; g < 4 / GROUP; ++g) {
          uint32_t r[32 * GROUP];
          const uint32_t tbase = tmem_base + acc * SS_BN + (uint32_t)(half * (SS_BN / 2) + g * GROUP * 32) +
                                 ((uint32_t)(q * 32) << 16);
          tmem_ld_32x32_at<0>(tbase, r);
          tmem_ld_32x32_at<32>(tbase + 32, r);
          if (GROUP == 4) {
            tmem_ld_32x32_at<(GROUP == 4 ? 64 : 0)>(tbase + 64, r);
            tmem_ld_32x32_at<(GROUP == 4 ? 96 : 0)>(tbase + 96, r);
          }
          ptx::tmem_ld_wait();
          if (g == 4 / GROUP - 1) {                     // last TMEM read of this warp for this tile
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty_bar[acc]), 0u));
          }
#pragma unroll
          for (int jj = 0; jj < GROUP; ++jj) {
            const int n0 = t * SS_BN + half * (SS_BN / 2) + (g * GROUP + jj) * 32;
            // columns to skip: beyond the matrix (TMA zero-filled them) and the row's seen items
            uint32_t skip = n0 + 32 <= args.n_items ? 0u : (n0 >= args.n_items ? 0xFFFFFFFFu : (0xFFFFFFFFu << (args.n_items - n0)));
            while (ns0 < n0 + 32) {
              if (ns0 >= n0) skip |= 1u << (ns0 - n0);
              ++sp;
              ns0 = ns1;
              ns1 = sp + 1 < se ? __ldg(args.seen_indices + sp + 1) : 0x7fffffff;
            }
            if (!live) skip = 0xFFFFFFFFu;
            const float thr0 = fmaxf(sv[KP - 1], thr_ext);
            // quick reject: nothing of this chunk beats the row's threshold (the common case once the lists have
            // warmed up); masked columns may only cause a false alarm here, they are excluded below
            float mx = __uint_as_float(r[32 * jj]);
#pragma unroll
            for (int e = 1; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[32 * jj + e]));
            if (!__any_sync(0xffffffffu, mx > thr0 && skip != 0xFFFFFFFFu)) continue;
            // slow path: bit e of pm = column e of the chunk beats the threshold and may be ranked
            uint32_t pm = 0u;
#pragma unroll
            for (int e = 0; e < 32; ++e) pm |= (__uint_as_float(r[32 * jj + e]) > thr0 ? 1u : 0u) << e;
            pm &= ~skip;
            if (!__any_sync(0xffffffffu, pm != 0u)) continue;
            // the chunk goes to shared memory so the survivors can be addressed dynamically; ONE copy of the insertion
            // code serves all lanes and all survivors (iterations = the largest survivor count of a lane, usually 1)
#pragma unroll
            for (int e = 0; e < 32; ++e) sval[e * 256] = __uint_as_float(r[32 * jj + e]);
            __syncwarp();
            while (__any_sync(0xffffffffu, pm != 0u)) {
              if (pm) {
                const int e = __ffs(pm) - 1;
                pm &= pm - 1;
                const float v = sval[e * 256];
                if (v > fmaxf(sv[KP - 1], thr_ext)) sel_insert<KP>(sv, si, v, n0 + e);
              }
            }
            __syncwarp();
          }
        }
        if (live && sv[KP - 1] > thr_pub) {             // publish (fire and forget)
          atomicMax(thr_slot, thr_key(sv[KP - 1]));
          thr_pub = sv[KP - 1];
        }
        thr_ext = fmaxf(thr_ext, thr_val(thr_seen));
      }
      if (live) {
        const size_t o = ((size_t)row * NL + (seg * 2 + half)) * KP;
        float4* dv = reinterpret_cast<float4*>(args.cand_val + o);
        int4* di = reinterpret_cast<int4*>(args.cand_idx + o);
#pragma unroll
        for (int j = 0; j < KP / 4; ++j) {
          dv[j] = make_float4(sv[4 * j], sv[4 * j + 1], sv[4 * j + 2], sv[4 * j + 3]);
          di[j] = make_int4(si[4 * j], si[4 * j + 1], si[4 * j + 2], si[4 * j + 3]);
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
struct ScoreSelectCall {
  const float* Q; int ldq;            // query factors [n_rows][k] (gathered rows), 16-byte aligned rows
  const float* V; int ldv;            // ranked factors [n_items][k]
  int n_rows, n_items, k;
  int KP;                             // 16 or 32
  int segs;                           // item segments (lists per row = 2 * segs)
  const int* users; const int* seen_indptr; const int* seen_indices;
  float* cand_val; int* cand_idx;
  unsigned int* row_thr;              // [n_rows], zeroed by the caller
  TmapCache* cache = nullptr;
  int max_ctas = 0;
};

inline int score_select_segments(int n_rows, int n_items, int num_sms) {
  // enough units to fill the CTA pairs evenly; every segment keeps >= 8 item tiles
  static int forced = -1;                  // GANMF_EVAL_SEGS=n: A/B switch
  if (forced < 0) { const char* e = getenv("GANMF_EVAL_SEGS"); forced = e ? atoi(e) : 0; }
  if (forced > 0) return forced > SS_MAX_LISTS / 2 ? SS_MAX_LISTS / 2 : forced;
  const int pairs = num_sms / 2, rb = (n_rows + SS_ROWS - 1) / SS_ROWS;
  const int tiles = (n_items + SS_BN - 1) / SS_BN;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= SS_MAX_LISTS / 2; ++s) {
    if (s > 1 && tiles / s < 8) break;
    const int units = rb * s;
    const double eff = (double)units / (((units + pairs - 1) / pairs) * pairs);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return best;
}

template <int KP>
inline cudaError_t score_select_launch(const ScoreSelectCall& c, const CUtensorMap& ma, const CUtensorMap& mb,
                                       cudaStream_t st) {
  using S = SelSmem<KP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(score_select_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  SelArgs a;
  a.n_rows = c.n_rows; a.n_items = c.n_items; a.nkb = (c.k + TC_BK - 1) / TC_BK;
  a.tiles_total = (c.n_items + SS_BN - 1) / SS_BN;
  a.segs = c.segs;
  a.tiles_per_seg = (a.tiles_total + c.segs - 1) / c.segs;
  a.row_blocks = (c.n_rows + SS_ROWS - 1) / SS_ROWS;
  a.idesc = make_idesc_tf32(SS_BN, 0, 0, 2 * TC_BM);
  a.users = c.users; a.seen_indptr = c.seen_indptr; a.seen_indices = c.seen_indices;
  a.cand_val = c.cand_val; a.cand_idx = c.cand_idx; a.row_thr = c.row_thr;
  const int sms = (c.max_ctas > 0 && c.max_ctas < num_sms) ? c.max_ctas : num_sms;
  const int units = a.row_blocks * a.segs;
  const int pairs = units < sms / 2 ? units : sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, score_select_kernel<KP>, ma, mb, a);
}

inline cudaError_t score_select(const ScoreSelectCall& c, cudaStream_t st) {
  if (c.n_rows <= 0 || c.n_items <= 0 || c.k <= 0 || c.k > SS_MAX_KB * TC_BK) return cudaErrorInvalidValue;
  if ((c.ldq & 3) || (c.ldv & 3) || c.segs < 1 || 2 * c.segs > SS_MAX_LISTS) return cudaErrorInvalidValue;
  CUtensorMap ma, mb;
  const int dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, sw = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  if (make_tmap_2d(&ma, c.Q, c.n_rows, c.k, c.ldq, TC_BK, TC_BM, dt, sw, c.cache)) return cudaErrorUnknown;
  if (make_tmap_2d(&mb, c.V, c.n_items, c.k, c.ldv, TC_BK, SS_BN / 2, dt, sw, c.cache)) return cudaErrorUnknown;
  if (c.KP == 16) return score_select_launch<16>(c, ma, mb, st);
  if (c.KP == 32) return score_select_launch<32>(c, ma, mb, st);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------ exact scores
// fl32( sum_k fp64(q_k * v_k) ): every product is exact in fp64, the sum is accurate to ~k * 2^-53, so the result is
// the correctly rounded fp32 score (independent of summation order except in astronomically rare double-rounding
// cases).  Warp-cooperative: lane l owns elements l, l+32, ...
__device__ __forceinline__ float exact_dot_warp(const float* __restrict__ q, const float* __restrict__ v, int k, int lane) {
  double acc = 0.0;
  for (int i = lane; i < k; i += 32) acc = fma((double)__ldg(q + i), (double)__ldg(v + i), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return (float)acc;
}

// Four exact scores at once (k <= 256: the query row sits in registers as 8 doubles per lane): the 32 loads of a
// group are in flight together and the four reductions interleave, so re-scoring runs at memory latency / 4.
__device__ __forceinline__ void load_query_regs(const float* __restrict__ q, int k, int lane, double (&qd)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) qd[i] = lane + 32 * i < k ? (double)__ldg(q + lane + 32 * i) : 0.0;
}
__device__ __forceinline__ void exact_dot4_warp(const double (&qd)[8], const float* __restrict__ V, int ldv,
                                                const int (&items)[4], int k, int lane, float (&out)[4]) {
  float x[4][8];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float* v = V + (size_t)items[u] * ldv + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[u][i] = lane + 32 * i < k ? __ldg(v + 32 * i) : 0.f;
  }
  double acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    acc[u] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[u] = fma(qd[i], (double)x[u][i], acc[u]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) out[u] = (float)acc[u];
}

// max_j ||V[j, :]||_2 -> *out (float bits, atomicMax on the non-negative pattern); one warp per 4 rows, all four rows'
// 16-byte loads in flight together
__global__ void row_norm_max_kernel(const float* __restrict__ V, int rows, int k, int ld, unsigned int* out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * 4;
  if (r0 >= rows) return;
  const int k4 = k >> 2;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lane; i < k4; i += 32) {
    float4 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      x[j] = r0 + j < rows ? __ldg(reinterpret_cast<const float4*>(V + (size_t)(r0 + j) * ld) + i)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = fmaf(x[j].x, x[j].x, s[j]); s[j] = fmaf(x[j].y, x[j].y, s[j]);
      s[j] = fmaf(x[j].z, x[j].z, s[j]); s[j] = fmaf(x[j].w, x[j].w, s[j]);
    }
  }
  for (int i = (k4 << 2) + lane; i < k; i += 32)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r0 + j < rows) { const float x = V[(size_t)(r0 + j) * ld + i]; s[j] = fmaf(x, x, s[j]); }
  float m = fmaxf(fmaxf(0.f, 0.f), 0.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    m = fmaxf(m, s[j]);
  }
  if (lane == 0) atomicMax(out, __float_as_uint(sqrtf(m) * 1.0001f));
}

// Total order of the evaluator: higher score first, then lower item index (== topk_key's order for finite scores)
__device__ __forceinline__ bool ranks_before(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

// One warp per row.  cand_*: [n_rows][NL][KP] from score_select_kernel.  out_idx/out_val: [n_rows][K].
// gamma: |tf32 score - exact| <= gamma * ||q|| * ||v||  (see DESIGN.md).  Rows that cannot be certified are
// appended to fb_rows (count in fb_count).
template <int KP>
__global__ void __launch_bounds__(128)
rescore_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int NL, int n_rows, int K,
               const float* __restrict__ Q, int ldq, const float* __restrict__ V, int ldv, int k,
               const unsigned int* __restrict__ vmax_bits, const unsigned int* __restrict__ row_thr, float gamma,
               int* __restrict__ out_idx,
               float* __restrict__ out_val, int* __restrict__ fb_count, int* __restrict__ fb_rows) {
  constexpr int CPL = SS_MAX_LISTS * KP / 32;            // candidates per lane (upper bound)
  constexpr int RMAX = 2;                                // re-scored entries per lane (64 per row)
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int C = NL * KP;
  const float* cv = cand_val + (size_t)row * C;
  const int* ci = cand_idx + (size_t)row * C;
  float v[CPL];
  int id[CPL];
  // everything a list of this row rejected or evicted scored (tf32) at most the row's final shared threshold
  const float tau = thr_val(__ldg(row_thr + row));
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int p = j * 32 + lane;
    v[j] = -INFINITY; id[j] = -1;
    if (p < C) { v[j] = cv[p]; id[j] = ci[p]; }
  }
  // a_K: K-th largest tf32 score among the candidates (order (value desc, position asc) makes them distinct)
  float cur_v = INFINITY;
  int cur_p = -1, n_valid = 0;
  for (int t = 0; t < K; ++t) {
    float bv = -INFINITY;
    int bp = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int p = j * 32 + lane;
      const bool after = id[j] >= 0 && (v[j] < cur_v || (v[j] == cur_v && p > cur_p));
      if (after && (v[j] > bv || (v[j] == bv && p < bp))) { bv = v[j]; bp = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
    }
    if (bp == 0x7fffffff) break;
    cur_v = bv; cur_p = bp; ++n_valid;
  }
  const float aK = n_valid == K ? cur_v : -INFINITY;
  // eps(row) = gamma * ||q|| * max ||v||
  const float* q = Q + (size_t)row * ldq;
  float qq = 0.f;
  for (int i = lane; i < k; i += 32) { const float x = __ldg(q + i); qq = fmaf(x, x, qq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  const float eps = gamma * sqrtf(qq) * 1.0001f * __uint_as_float(__ldg(vmax_bits));
  const float keep_from = aK - 2.f * eps;                // members of the exact top K score at least this in tf32
  // exact re-scoring of the survivors, spread over the lanes (entry e lives in lane e % 32, slot e / 32)
  __shared__ int s_items[4][32 * RMAX];
  int* my_items = s_items[threadIdx.x >> 5];
  float rv[RMAX];
  int ri[RMAX];
#pragma unroll
  for (int s = 0; s < RMAX; ++s) { rv[s] = -INFINITY; ri[s] = -1; }
  int n_res = 0;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const bool keep = id[j] >= 0 && v[j] >= keep_from;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int pos = n_res + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < 32 * RMAX) my_items[pos] = id[j];
    n_res += __popc(bal);
  }
  const bool overflow = n_res > 32 * RMAX;
  n_res = min(n_res, 32 * RMAX);
  __syncwarp();
  double qd[8];
  load_query_regs(q, k, lane, qd);
  for (int e0 = 0; e0 < n_res; e0 += 4) {
    int items[4];
    float ex[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) items[u] = my_items[min(e0 + u, n_res - 1)];
    exact_dot4_warp(qd, V, ldv, items, k, lane, ex);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u;
      if (e < n_res && lane == (e & 31)) {
#pragma unroll
        for (int s = 0; s < RMAX; ++s) if (s == (e >> 5)) { rv[s] = ex[u]; ri[s] = items[u]; }
      }
    }
  }
  // rank of every re-scored entry = number of entries that come before it
  int rk[RMAX];
#pragma unroll
  for (int s = 0; s < RMAX; ++s) rk[s] = 0;
  for (int e = 0; e < n_res; ++e) {
    float ov = -INFINITY;
    int oi = -1;
#pragma unroll
    for (int s = 0; s < RMAX; ++s) if (s == (e >> 5)) { ov = rv[s]; oi = ri[s]; }
    ov = __shfl_sync(0xffffffffu, ov, e & 31);
    oi = __shfl_sync(0xffffffffu, oi, e & 31);
#pragma unroll
    for (int s = 0; s < RMAX; ++s)
      if (ri[s] >= 0 && ranks_before(ov, oi, rv[s], ri[s])) ++rk[s];
  }
  float tK = -INFINITY;                                  // K-th exact score (if K entries exist)
#pragma unroll
  for (int s = 0; s < RMAX; ++s) {
    if (ri[s] >= 0 && rk[s] < K) {
      out_idx[(size_t)row * K + rk[s]] = ri[s];
      out_val[(size_t)row * K + rk[s]] = rv[s];
    }
    if (ri[s] >= 0 && rk[s] == K - 1) tK = rv[s];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tK = fmaxf(tK, __shfl_xor_sync(0xffffffffu, tK, o));
  for (int j = n_res + lane; j < K; j += 32) {          // fewer than K rankable items: pad like topk_rows
    out_idx[(size_t)row * K + j] = -1;
    out_val[(size_t)row * K + j] = -INFINITY;
  }
  // certificate: no item outside the lists can reach the K-th place.  tau = -inf: no list ever rejected anything
  // (all rankable items are candidates).
  const bool certified = !overflow && (tau == -INFINITY || (n_res >= K && tK > tau + eps));
  if (!certified && lane == 0) fb_rows[atomicAdd(fb_count, 1)] = row;
}

// Exact score rows for the fallback: out[f][i] = exact score of query row fb_rows[f] against item i; one warp per
// (row, item) pair in turn, grid.y = fallback slot.
__global__ void __launch_bounds__(256)
exact_score_rows_kernel(const int* __restrict__ fb_rows, int f0, int n_fb, const float* __restrict__ Q, int ldq,
                        const float* __restrict__ V, int ldv, int k, int n_items, float* __restrict__ out, int ldo) {
  const int f = blockIdx.y;
  if (f >= n_fb) return;
  const int row = fb_rows[f0 + f];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float* q = Q + (size_t)row * ldq;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n_items; i += gridDim.x * wpb) {
    const float s = exact_dot_warp(q, V + (size_t)i * ldv, k, lane);
    if (lane == 0) out[(size_t)f * ldo + i] = s;
  }
}

// fallback plumbing: user ids of the fallback rows; top-K rows of the fallback block back to their row slots
__global__ void gather_ids_kernel(const int* __restrict__ users, const int* __restrict__ fb_rows, int f0, int n,
                                  int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = users[fb_rows[f0 + i]];
}
__global__ void scatter_topk_kernel(const int* __restrict__ src_idx, const float* __restrict__ src_val,
                                    const int* __restrict__ fb_rows, int f0, int n, int K, int* __restrict__ out_idx,
                                    float* __restrict__ out_val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  const size_t o = (size_t)fb_rows[f0 + i / K] * K + i % K;
  out_idx[o] = src_idx[i];
  out_val[o] = src_val[i];
}

// RMSE column of the evaluator without a score matrix: exact scores of the user's TEST items only
// (metrics.py:634-659; seen items count as -inf and drop out, as in the masked score rows).  One warp per row.
__global__ void __launch_bounds__(128)
user_rmse_exact_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ V, int ldv, int k,
                       const int* __restrict__ users, int n_rows, int n_cut, EvalTables tb,
                       const float* __restrict__ test_data, const int* __restrict__ seen_indptr,
                       const int* __restrict__ seen_indices, float* __restrict__ scratch, double* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const int u = users[r];
  const int ts = tb.test_indptr[u], T = tb.test_indptr[u + 1] - ts;
  const float* q = Q + (size_t)r * ldq;
  float* e = scratch + ts;
  int n = 0;
  double qd[8];
  load_query_regs(q, k, lane, qd);                        // (the fused route guarantees k <= 256)
  for (int i0 = 0; i0 < T; i0 += 32) {
    // lane l looks at test entry i0 + l: is the item seen (masked to -inf, dropped)?
    const int i = i0 + lane;
    bool use = i < T;
    const int item = use ? tb.test_indices[ts + i] : 0;
    if (use && seen_indptr) {
      int lo = seen_indptr[u], hi = seen_indptr[u + 1] - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1, v = seen_indices[mid];
        if (v == item) { use = false; break; }
        if (v < item) lo = mid + 1; else hi = mid - 1;
      }
    }
    unsigned bal = __ballot_sync(0xffffffffu, use);
    while (bal) {                                         // four unseen test items per pass, in entry order
      int src[4], items[4];
      float ex[4];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (bal) { src[j] = __ffs(bal) - 1; bal &= bal - 1; ++cnt; }
        else src[j] = src[0];                             // (padding: re-scores the first item, result unused)
        items[j] = __shfl_sync(0xffffffffu, item, src[j]);
      }
      exact_dot4_warp(qd, V, ldv, items, k, lane, ex);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < cnt) {
          const float d = ex[j] - test_data[ts + i0 + src[j]];
          const float sq = d * d;
          if (isfinite(sq)) { if (lane == 0) e[n] = sq; ++n; }
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    const double v = n ? sqrt((double)np_sum_f32_buf(e, n) / (double)n) : nan("");
    for (int ci = 0; ci < n_cut; ++ci) vals[((size_t)r * n_cut + ci) * MC_NCOL + MC_RMSE] = v;
  }
}

}  // namespace ganmf
