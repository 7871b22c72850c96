// Generator GEMM with a RESIDENT A tile (sm_100a):  out[M, N] = A[M, K] . B[N, K]^T  for K <= 256, both operands
// K-major -- the fake profiles F = P[uids] . V^T of GANRec/GANMF.py:82-84 (M = minibatch rows, N = items, K = k).
//
// Why a second GEMM kernel.  With K = 250 a 256 x 256 output tile costs 8 k-blocks of MMA work (~3.4 us on a CTA
// pair) but the general kernel (tc_gemm.cuh) re-loads BOTH operand tiles for it (2 x 128 KB per CTA) and stores
// 128 KB per CTA: 384 KB through the SM <-> L2 port per tile, which runs at ~80 GB/s per SM -> 4.8 us.  ncu shows
// exactly that: tensor pipe 45 %, the MMA warp waiting on the operand ring, HBM at 62 % although the kernel only has
// to write its output.  Here a CTA pair keeps the A tile of its 256 rows in shared memory (8 k-blocks x 16 KB per
// CTA) for a whole row block and streams only B: 128 KB in + 128 KB out per tile per CTA.  Same skeleton as the
// fused scorer (score_select.cuh: resident query tile, B ring, double-buffered TMEM accumulators); the epilogue
// pulls the warp's 128-column share out of TMEM in two register groups, hands the stage back, and stores through the
// XOR-swizzled slab so that every global store is a full 128-byte row segment.
//   warp 0: TMA producer   warp 1: MMA issuer (leader CTA)   warps 2..9: store epilogue
#pragma once
#include "score_select.cuh"

namespace ganmf {

constexpr int GG_ROWS = 256;           // rows per CTA pair
constexpr int GG_BN = 256;             // columns per tile
constexpr int GG_STAGES = 4;           // B ring depth (16 KB per stage per CTA)
constexpr int GG_MAX_KB = 8;           // resident A: up to 8 k-blocks of 32 (K <= 256)

struct GenArgs {
  int M, N, nkb;
  int tiles_total;                     // column tiles of GG_BN
  int segs, tiles_per_seg;             // a unit = (row block, segment of column tiles)
  int row_blocks;
  uint32_t idesc;
  float* out;
  int ldo;                             // >= roundup(N, 32): a chunk that starts below N lies inside the padded row
};

struct GenSmem {
  static constexpr int A_KB_BYTES = TC_BM * TC_BK * 4;                 // 16 KB: 128 rows x 32 k
  static constexpr int B_BYTES = (GG_BN / 2) * TC_BK * 4;              // 16 KB: 128 columns x 32 k
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = GG_MAX_KB * A_KB_BYTES;
  static constexpr int BAR_OFF = B_OFF + GG_STAGES * B_BYTES;
  static constexpr int SLAB_OFF = BAR_OFF + (2 * GG_STAGES + 6) * 8 + 16;       // 8 warps x 32 rows x 8 float4
  static constexpr int TOTAL = SLAB_OFF + 8 * 32 * 32 * 4 + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
resident_a_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const GenArgs args) {
  using S = GenSmem;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader: issues the pair's MMAs
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + GG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GG_STAGES;       // [2]
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
    for (int s = 0; s < GG_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
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

  // units are walked statically; consecutive units share the column segment, so the pairs running side by side
  // stream the same B tiles (L2 hits) on different row blocks
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
        const int m0 = unit_rb(u) * GG_ROWS + (int)rank * TC_BM;
        ptx::mbar_wait(a_empty_bar, (ui & 1) ^ 1);
        const uint32_t afb = ptx::mapa(ptx::smem_u32(a_full_bar), 0u);
        if (rank == 0) ptx::mbar_expect_tx(a_full_bar, 2 * args.nkb * S::A_KB_BYTES);
        for (int kb = 0; kb < args.nkb; ++kb)
          ptx::tma_load_2d_pair(smem + S::A_OFF + kb * S::A_KB_BYTES, &map_a, afb, kb * TC_BK, m0);
        for (int t = t0; t < t1; ++t) {
          const int n0 = t * GG_BN + (int)rank * (GG_BN / 2);
          for (int kb = 0; kb < args.nkb; ++kb, ++it) {
            const int s = it % GG_STAGES;
            ptx::mbar_wait(&empty_bar[s], ((it / GG_STAGES) & 1) ^ 1);
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
          const uint32_t tmem_d = tmem_base + acc * GG_BN;
          for (int kb = 0; kb < args.nkb; ++kb, ++it) {
            const int s = it % GG_STAGES;
            ptx::mbar_wait(&full_bar[s], (it / GG_STAGES) & 1);
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
    // ------------------------------------------------------------ store epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter: rows q*32 .. q*32+31 of this CTA's 128
    const int half = ew >> 2;                      // column half of every tile
    float4* slab = reinterpret_cast<float4*>(smem + S::SLAB_OFF) + ew * 256;     // 32 rows x 8 float4
    const int sub_r = lane >> 3, sub_g = lane & 7;   // coalesced phase: 4 rows x 8 lanes x float4
    uint32_t ti = 0;
    for (int u = pair; u < n_units; u += n_pairs) {
      int t0, t1;
      seg_tiles(unit_seg(u), t0, t1);
      const int mw = unit_rb(u) * GG_ROWS + (int)rank * TC_BM + q * 32;     // first row of this warp's slab
      for (int t = t0; t < t1; ++t, ++ti) {
        const uint32_t acc = ti & 1;
        ptx::mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
        ptx::tc_fence_after();
        // two groups of 64 columns (64 registers each: the whole 128-column share would spill); the stage goes back
        // to the MMA warp right after the second group has left TMEM, before its stores are issued
        const bool rows_live = mw < args.M;                                   // warp-uniform
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          uint32_t r[64];
          const uint32_t tbase = tmem_base + acc * GG_BN + (uint32_t)(half * (GG_BN / 2) + g2 * 64) +
                                 ((uint32_t)(q * 32) << 16);
          tmem_ld_32x32_at<0>(tbase, r);
          tmem_ld_32x32_at<32>(tbase + 32, r);
          ptx::tmem_ld_wait();
          if (g2 == 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty_bar[acc]), 0u));
          }
          if (!rows_live) continue;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int nb = t * GG_BN + half * (GG_BN / 2) + (g2 * 2 + jj) * 32;
            if (nb >= args.N) continue;                                      // warp-uniform
            // thread = row: 8 float4 stores, column group g lands in slot g ^ (row & 7) (conflict free)
#pragma unroll
            for (int g = 0; g < 8; ++g)
              slab[lane * 8 + (g ^ (lane & 7))] =
                  make_float4(__uint_as_float(r[32 * jj + 4 * g]), __uint_as_float(r[32 * jj + 4 * g + 1]),
                              __uint_as_float(r[32 * jj + 4 * g + 2]), __uint_as_float(r[32 * jj + 4 * g + 3]));
            __syncwarp();
            float* dst = args.out + (size_t)(mw + sub_r) * args.ldo + nb + sub_g * 4;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int row = itr * 4 + sub_r;
              if (mw + row < args.M)
                *reinterpret_cast<float4*>(dst + (size_t)itr * 4 * args.ldo) = slab[row * 8 + (sub_g ^ (row & 7))];
            }
            __syncwarp();                           // slab is reused by the next chunk
          }
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
struct GenGemmCall {
  const float* A; int lda;            // [M][K], K-major
  const float* B; int ldb;            // [N][K], K-major
  int M, N, K;
  float* out; int ldo;
  TmapCache* cache = nullptr;
  int max_ctas = 0;
};

inline bool resident_a_gemm_ok(const GenGemmCall& c) {
  return c.M > 0 && c.N > 0 && c.K > 0 && c.K <= GG_MAX_KB * TC_BK && !(c.lda & 3) && !(c.ldb & 3) && !(c.ldo & 3) &&
         c.ldo >= ((c.N + 31) & ~31) && !(reinterpret_cast<uintptr_t>(c.A) & 15) && !(reinterpret_cast<uintptr_t>(c.B) & 15) &&
         !(reinterpret_cast<uintptr_t>(c.out) & 15);
}

inline cudaError_t resident_a_gemm(const GenGemmCall& c, cudaStream_t st) {
  if (!resident_a_gemm_ok(c)) return cudaErrorInvalidValue;
  using S = GenSmem;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(resident_a_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
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
  CUtensorMap ma, mb;
  const int dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, sw = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  if (make_tmap_2d(&ma, c.A, c.M, c.K, c.lda, TC_BK, TC_BM, dt, sw, c.cache)) return cudaErrorUnknown;
  if (make_tmap_2d(&mb, c.B, c.N, c.K, c.ldb, TC_BK, GG_BN / 2, dt, sw, c.cache)) return cudaErrorUnknown;
  GenArgs a;
  a.M = c.M; a.N = c.N; a.nkb = (c.K + TC_BK - 1) / TC_BK;
  a.tiles_total = (c.N + GG_BN - 1) / GG_BN;
  a.row_blocks = (c.M + GG_ROWS - 1) / GG_ROWS;
  const int sms = (c.max_ctas > 0 && c.max_ctas < num_sms) ? c.max_ctas : num_sms;
  // column segments: enough units to fill the CTA pairs evenly, every segment keeps >= 8 tiles so the resident A
  // tile is amortised
  const int pairs_avail = sms / 2;
  int segs = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 64; ++s) {
    if (s > 1 && a.tiles_total / s < 8) break;
    const int units = a.row_blocks * s;
    const double eff = (double)units / (((units + pairs_avail - 1) / pairs_avail) * pairs_avail);
    if (eff > best_eff + 0.02) { best_eff = eff; segs = s; }
  }
  a.segs = segs;
  a.tiles_per_seg = (a.tiles_total + segs - 1) / segs;
  a.idesc = make_idesc_tf32(GG_BN, 0, 0, 2 * TC_BM);
  a.out = c.out; a.ldo = c.ldo;
  const int units = a.row_blocks * a.segs;
  const int pairs = units < pairs_avail ? units : pairs_avail;
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
  return cudaLaunchKernelEx(&cfg, resident_a_gemm_kernel, ma, mb, a);
}

}  // namespace ganmf
