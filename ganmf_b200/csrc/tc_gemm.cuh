// TF32 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared-memory ring ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld epilogue.
//
//   D[m,n] = epilogue( sum_k A[m,k] * B[n,k] )          A: M x K, B: N x K (logical)
//
// Each operand is an fp32 row-major matrix in HBM and is either
//   K-major  : stored [MN rows][K cols]   (reduction index contiguous), or
//   MN-major : stored [K rows][MN cols]   (M / N index contiguous).
// All four combinations occur on the GANMF path:
// forward GEMMs read weights MN-major, the backward ones read the same weights
// K-major, weight-gradient GEMMs read both activations MN-major.
//
// One CTA computes one 128 x BN output tile (optionally one K-split of it).
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9  : epilogue, two warps per TMEM lane quarter (warp_id % 4), half the columns each
// Stage = 32 fp32 of K (one 128-byte swizzle row), i.e. 4 UMMA (K=8) per stage.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "ptx.cuh"

namespace ganmf {

constexpr int TC_BM = 128;         // output tile rows (UMMA M)
constexpr int TC_BK = 32;          // fp32 elements of K per stage (128 bytes)
constexpr int TC_UMMA_K = 8;       // tf32: 32 bytes of K per instruction
constexpr int TC_THREADS = 320;     // TMA warp, MMA warp, 8 epilogue warps
constexpr int TC_SCHED = 8;         // depth of the work-unit ring between the producer and its consumers
constexpr int TC_SCHED_SLOTS = 64;  // scheduler states cycled over launches (GEMMs in flight at once)

enum Act { ACT_LINEAR = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };
__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case ACT_TANH: return tanhf(z);
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}

// Fused epilogue applied to every accumulator element (m, n):
//   acc' = acc + r1_row[m] * r1_col[n]        (rank-1 term: the id column of DisGANMF's concat input)
//   v = act( alpha * rs(m) * acc' + bias[n] + beta1 * C1[m,n] + beta2 * C2[m,n] )
//   rs(m) = row_scale2 ? row_scale2[m >= row_split] : 1      (device-side scalars:
//           the hinge-gate coefficients are only known on the device)
//   out[m,n] = round_out ? rna_tf32(v) : v
//   sumsq2[m >= row_split] += v*v                            (energy loss, fp64)
//   colpart[m / 32][n]     = sum of v over the 32-row group   (bias gradients without another pass)
struct Epilogue {
  float* out = nullptr;
  int ldo = 0;
  float alpha = 1.f;
  const float* row_scale2 = nullptr;
  int row_split = 0x7fffffff;
  const float* bias = nullptr;
  const float* c1 = nullptr;
  int ldc1 = 0;
  float beta1 = 0.f;
  const float* c2 = nullptr;
  int ldc2 = 0;
  float beta2 = 0.f;
  int round_out = 0;
  double* sumsq2 = nullptr;
  // per-32-row column sums of the stored values: colpart[(m >> 5) * ldcp + n] = sum of out[m & ~31 .. +31][n]
  // (rows >= M count as zero).  Only unsplit launches of the lean tensor-core kernel honour it (the caller checks,
  // tc_lean_epilogue()); the decoder-bias
  // gradient is then a weighted sum of M/32 partial rows instead of a pass over the whole residual.
  float* colpart = nullptr;
  int ldcp = 0;
  const float* r1_row = nullptr;
  const float* r1_col = nullptr;
  int act = ACT_LINEAR;
  // Fused optimiser (single-GPU weight-gradient GEMMs): when adam_m != nullptr the epilogue value is the
  // data gradient g of the parameter matrix `out` and, instead of being stored, drives TF's ApplyAdam in
  // place:  g' = g + reg*theta; m += (g'-m)(1-b1); v += (g'^2-v)(1-b2); theta -= alpha*m/(sqrt(v)+eps);
  // adam_l2 accumulates sum(theta_old^2) (l2 term of the reported loss).  m, v share out's layout.
  float* adam_m = nullptr;
  float* adam_v = nullptr;
  float adam_alpha = 0.f;
  float adam_reg = 0.f;
  double* adam_l2 = nullptr;
};

constexpr float TC_ADAM_B1 = 0.9f, TC_ADAM_B2 = 0.999f, TC_ADAM_EPS = 1e-8f;
// The ONE element update every optimiser path shares (fused_adam_kernel, the GEMM epilogue, the lazy
// user-factor kernels).  Every operation is an explicit intrinsic, so no path depends on the compiler's
// contraction choices and a replayed update is bit-identical to the one it stands for.  The square root and
// the quotient are the hardware approximations (sqrt.approx / div.approx, <= 2 ulp, deterministic): inside the
// weight-gradient GEMM epilogues the IEEE sequences (~20 extra instructions per parameter) made the optimiser
// instruction-issue bound; the 2^-22 relative error sits four orders below the stated parity tolerance.
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void adam_elem(float g, float& th, float& m, float& v, float alpha, float reg) {
  const float ge = __fmaf_rn(reg, th, g);
  m = __fmaf_rn(__fsub_rn(ge, m), 1.f - TC_ADAM_B1, m);
  v = __fmaf_rn(__fmaf_rn(ge, ge, -v), 1.f - TC_ADAM_B2, v);
  th = __fsub_rn(th, __fdividef(__fmul_rn(m, alpha), __fadd_rn(sqrt_approx(v), TC_ADAM_EPS)));
}

struct TcGemmArgs {
  int M, N, K;
  int a_mn, b_mn;          // 1 = MN-major operand
  int kb_per_split;        // k-blocks (of TC_BK) per K split
  int splits;
  int dbg_epi;             // bring-up: 1 = no global stores, 2 = no TMEM loads either
  float* ws;               // split-K partials [splits][M][ws_ld] (splits > 1)
  int ws_ld;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // smem descriptor strides, bytes >> 4
  uint32_t a_layout, b_layout;           // UMMA LayoutType of each operand's smem tile
  uint32_t a_kstep, b_kstep;             // start-address advance per UMMA_K, bytes >> 4
  uint32_t a_mtstep;                     // start-address advance per 128-row sub-tile of A, bytes >> 4
  uint32_t idesc;
  unsigned int* sched;     // {next unit, CTAs done}: dynamic unit scheduler state (self-resetting)
  int pf_planes;           // > 0: map_p describes [planes][M][N] (theta, m, v of the fused optimiser): the producer
                           // prefetches the unit's tile of every plane into L2 before its operand loads
  Epilogue ep;
};

__device__ __forceinline__ float apply_epilogue(const Epilogue& ep, float acc, int m, int n,
                                                float rs) {
  if (ep.r1_row) acc = fmaf(__ldg(ep.r1_row + m), __ldg(ep.r1_col + n), acc);
  float v = ep.alpha * rs * acc;
  if (ep.bias) v += __ldg(ep.bias + n);
  if (ep.c1) v += ep.beta1 * __ldg(ep.c1 + (size_t)m * ep.ldc1 + n);
  if (ep.c2) v += ep.beta2 * __ldg(ep.c2 + (size_t)m * ep.ldc2 + n);
  if (ep.act != ACT_LINEAR) v = act_fwd(ep.act, v);
  if (ep.round_out) v = ptx::round_tf32(v);
  return v;
}

// Generic (scalar) tail of the epilogue: store the value, or run the fused Adam update with it.
// Returns the quantity that feeds the sum-of-squares accumulators (value^2, or theta_old^2 in Adam mode).
__device__ __forceinline__ float finish_element(const Epilogue& ep, float x, int m, int n) {
  const size_t o = (size_t)m * ep.ldo + n;
  if (!ep.adam_m) {
    ep.out[o] = x;
    return x * x;
  }
  float th = ep.out[o], mm = ep.adam_m[o], vv = ep.adam_v[o];
  const float th0 = th;
  adam_elem(x, th, mm, vv, ep.adam_alpha, ep.adam_reg);
  ep.out[o] = th; ep.adam_m[o] = mm; ep.adam_v[o] = vv;
  return th0 * th0;
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo,
                                                   uint32_t layout_type) {
  // cute::UMMA::SmemDescriptor layout (sm_100): start[0,14) lbo[16,30) sbo[32,46)
  // version=1 [46,48) layout_type[61,64): SWIZZLE_128B = 2 (16-byte swizzle atoms, used for
  // K-major tiles), SWIZZLE_128B_BASE32B = 1 (32-byte atoms: the only swizzled layout the
  // hardware accepts for MN-major 32-bit operands).
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo & 0x3FFF) << 16;
  d |= (uint64_t)(sbo & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// MT = number of 128-row MMA tiles stacked along M in one CTA tile (1 or 2).  MT = 2 (a 256 x BN CTA
// tile, two accumulators sharing every B stage) raises the flops per byte fetched from L2 by 1.5x for
// the long-K GEMMs that are L2-bandwidth bound; it fills the whole TMEM with BN = 256, so those tiles
// do not double-buffer the accumulator (fine: their epilogue is a small fraction of a long K loop).
// CG = 2 (a CTA pair = the two SMs of a TPC, tcgen05 cta_group::2): the pair computes ONE 256 x BN tile, each
// CTA holding its 128 accumulator rows in its own TMEM and only HALF of every B stage in its shared memory
// (the tensor cores read the peer's half directly).  Per CTA a stage is 16 KB of A + 16 KB of B for the same
// 128 x 256 x 32 MMA volume that costs 48 KB on a single CTA: the GEMMs with many output tiles are bound by
// the L2 -> SM fill rate (ncu: ~12 TB/s, the LTS cap), so 1.5x fewer bytes per flop is 1.5x more flops, and
// unlike MT = 2 the accumulator still double-buffers (256 of the 512 TMEM columns per stage).
template <int BN, int STAGES, int MT = 1, int CG = 1>
struct TcSmem {
  static constexpr int A_BYTES = MT * TC_BM * TC_BK * 4;
  static constexpr int B_BYTES = (BN / CG) * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;          // 8 warps x 32 rows x 8 float4 (swizzled)
  static constexpr int EPI_BYTES = 8 * 32 * 32 * 4;
  static constexpr int BAR_OFF = EPI_OFF + EPI_BYTES;
  static constexpr int SCHED_OFF = BAR_OFF + (2 * STAGES + 4) * 8 + 16;     // unit ring: full/empty barriers + ids
  static constexpr int TOTAL = SCHED_OFF + TC_SCHED * (8 + 8 + 4) + 16 + 1024;  // + align slack
};

// Work unit = (output tile, K split).  Units are numbered so that consecutive units (which run
// concurrently on neighbouring SMs) share the operand tile of the dimension with FEWER tiles: the
// small operand stays L2 resident and the large one streams from HBM once.
struct TcUnit { int m0, n0, split; };
__device__ __forceinline__ TcUnit tc_unit(int u, int tiles_m, int tiles_n, int m_fastest, int BN,
                                          int BM = TC_BM) {
  const int tiles = tiles_m * tiles_n;
  TcUnit r;
  r.split = u / tiles;
  const int t = u - r.split * tiles;
  int tm, tn;
  if (m_fastest) { tn = t / tiles_m; tm = t - tn * tiles_m; }
  else           { tm = t / tiles_n; tn = t - tm * tiles_n; }
  r.m0 = tm * BM;
  r.n0 = tn * BN;
  return r;
}

// Persistent kernel: grid = min(#units, #SMs).  Units are handed out DYNAMICALLY: the producer lane takes
// the next unit from a global counter (one atomicAdd per unit, issued one unit ahead so its latency hides
// under the loads) and publishes it to the MMA and epilogue warps through a small shared-memory ring.
// A static blockIdx.x + i*gridDim.x walk needs every CTA resident from the start; when a collective
// (NCCL over NVLink) holds some SMs, the CTAs that cannot be placed would run as a second wave and double
// the GEMM's duration.  With the counter, late CTAs simply find less (or no) work.  The last CTA to leave
// resets the counter pair, so launches need no memset.
// Two TMEM accumulator stages let the epilogue of unit i overlap the MMAs of unit i+1.
// EPI = 1 ("lean"): the epilogue without the fused-optimiser, activation and tf32-rounding variants -- the
// kernel the many-tile GEMMs of the step use; its epilogue is half the code (the full one overflows the
// instruction cache: ncu attributes 10-17 % of the epilogue warps' stalls to instruction fetch).
template <int BN, int STAGES, int MT, int CG, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_p, const TcGemmArgs args) {
  static_assert(CG == 1 || (CG == 2 && MT == 1), "a CTA pair holds one 128-row sub-tile per CTA");
  using S = TcSmem<BN, STAGES, MT, CG>;
  constexpr bool LEAN = EPI == 1;
  constexpr int BM = MT * TC_BM * CG;                  // rows of a unit's tile (CTA pair: 256)
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;     // 0 = leader: issues the pair's MMAs
  const int row_off = (int)rank * TC_BM;               // this CTA's rows inside the tile
  const int col_off = (int)rank * (BN / CG);           // this CTA's share of the B tile
  constexpr int ACC = (2 * MT * BN <= 512) ? 2 : 1;    // accumulator stages that fit the 512 TMEM columns
  constexpr int ACC_COLS = MT * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* sched_full = reinterpret_cast<uint64_t*>(smem + S::SCHED_OFF);
  uint64_t* sched_empty = sched_full + TC_SCHED;
  volatile int* sched_unit = reinterpret_cast<volatile int*>(sched_empty + TC_SCHED);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_kb = (args.K + TC_BK - 1) / TC_BK;
  const int tiles_m = (args.M + BM - 1) / BM;
  const int tiles_n = (args.N + BN - 1) / BN;
  const int n_units = tiles_m * tiles_n * args.splits;
  const int m_fastest = tiles_m <= tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 8 * CG);    // one arrival per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < TC_SCHED; ++i) {
      ptx::mbar_init(&sched_full[i], 1);
      ptx::mbar_init(&sched_empty[i], 9);            // MMA lane + one per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { ptx::tmem_alloc_pair(tmem_slot, ACC * ACC_COLS); ptx::tmem_relinquish_pair(); }
    else         { ptx::tmem_alloc(tmem_slot, ACC * ACC_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();                    // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t it = 0, si = 0;
      // first unit = blockIdx.x (no atomic in front of the first load; single-wave GEMMs behave exactly like
      // a static grid), every later unit = gridDim.x + counter
      // (a CTA pair walks units statically: both CTAs must take the same sequence)
      const bool dyn = args.sched != nullptr;
      const int g = (int)gridDim.x / CG;
      int u = (int)blockIdx.x / CG;
      while (true) {
        const int sl = si % TC_SCHED;
        ptx::mbar_wait(&sched_empty[sl], ((si / TC_SCHED) & 1) ^ 1);
        sched_unit[sl] = u < n_units ? u : -1;
        ptx::mbar_arrive(&sched_full[sl]);           // release: the id is visible to the waiters
        ++si;
        if (u >= n_units) break;
        // in flight while this unit's loads are issued
        const int u_next = dyn ? g + (int)atomicAdd(args.sched, 1u) : u + g;
        const TcUnit un = tc_unit(u, tiles_m, tiles_n, m_fastest, BN, BM);
        const int kb_begin = un.split * args.kb_per_split;
        const int kb_end = min(kb_begin + args.kb_per_split, total_kb);
        if (!LEAN && args.pf_planes > 0) {
          // fused optimiser: this CTA's rows of theta / m / v under the unit's tile start their way into L2 now; the
          // epilogue warps read them a whole mainloop later (their LSU loads then cost an L2 hit, not an HBM round trip)
#pragma unroll 1
          for (int pl = 0; pl < args.pf_planes; ++pl)
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_prefetch_l2_3d(&map_p, un.n0, un.m0 + row_off + mt * TC_BM, pl);
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          const int k0 = kb * TC_BK;
          // pair: both CTAs' bytes are counted on the LEADER's full barrier (it alone consumes the stage)
          const uint32_t fb = CG == 2 ? ptx::mapa(ptx::smem_u32(&full_bar[s]), 0u) : 0u;
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[s], CG * S::STAGE_BYTES);
          auto ld = [&](uint8_t* dst, const CUtensorMap* map, int c0, int c1) {
            if (CG == 2) ptx::tma_load_2d_pair(dst, map, fb, c0, c1);
            else ptx::tma_load_2d(dst, map, &full_bar[s], c0, c1);
          };
          const int am = un.m0 + row_off, bn0 = un.n0 + col_off;
          if (!args.a_mn) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)                                 // box {32 k, 128 m}
              ld(sa + mt * (TC_BM * TC_BK * 4), &map_a, k0, am + mt * TC_BM);
          } else {
#pragma unroll
            for (int j = 0; j < MT * TC_BM / 32; ++j)                       // box {32 m, 32 k}
              ld(sa + j * (TC_BK * 128), &map_a, am + 32 * j, k0);
          }
          if (!args.b_mn) {
            ld(sb, &map_b, k0, bn0);                                        // box {32 k, BN / CG n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / CG / 32; ++j)
              ld(sb + j * (TC_BK * 128), &map_b, bn0 + 32 * j, k0);
          }
        }
        u = u_next;
      }
      // every CTA makes its last counter increment before it reports done, so the CTA that sees all others
      // done can hand a clean state to the next launch that uses this slot
      if (dyn) {
        __threadfence();
        if (atomicAdd(args.sched + 1, 1u) == gridDim.x - 1) {
          args.sched[0] = 0u;
          args.sched[1] = 0u;
          __threadfence();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      uint32_t it = 0, ui = 0;
      for (;; ++ui) {
        const int sl = ui % TC_SCHED;
        ptx::mbar_wait(&sched_full[sl], (ui / TC_SCHED) & 1);
        const int u = sched_unit[sl];
        ptx::mbar_arrive(&sched_empty[sl]);
        if (u < 0) break;
        if (CG == 2 && rank != 0) continue;              // the leader issues the pair's MMAs
        const TcUnit un = tc_unit(u, tiles_m, tiles_n, m_fastest, BN, BM);
        const int kb_begin = un.split * args.kb_per_split;
        const int nkb = min(kb_begin + args.kb_per_split, total_kb) - kb_begin;
        const uint32_t acc = ui % ACC;
        ptx::mbar_wait(&tmem_empty_bar[acc], ((ui / ACC) & 1) ^ 1);    // epilogue drained this stage
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
          const uint64_t da = make_smem_desc(sa, args.a_lbo, args.a_sbo, args.a_layout);
          const uint64_t db = make_smem_desc(sb, args.b_lbo, args.b_sbo, args.b_layout);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {          // the 128-row sub-tiles share the B stage
              const uint64_t dak = da + (uint64_t)(mt * args.a_mtstep + k * args.a_kstep);
              const uint64_t dbk = db + (uint64_t)(k * args.b_kstep);
              if (CG == 2) ptx::mma_tf32_ss_pair(tmem_d, dak, dbk, args.idesc, (i | k) ? 1u : 0u);
              else ptx::mma_tf32_ss(tmem_d + mt * BN, dak, dbk, args.idesc, (i | k) ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs of a pair) when those MMAs retire
          if (CG == 2) ptx::mma_commit_pair(&empty_bar[s], 3);
          else ptx::mma_commit(&empty_bar[s]);
        }
        // accumulator of this unit complete (the peer's epilogue warps wait on their own copy of the barrier)
        if (CG == 2) ptx::mma_commit_pair(&tmem_full_bar[acc], 3);
        else ptx::mma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter, each draining half of the tile's columns: the epilogue is
    // instruction-latency bound, so it needs several warps per scheduler to hide under the next
    // tile's MMAs.  Per 32x32 chunk: tcgen05.ld (thread = row) -> XOR-swizzled smem slab ->
    // read back with lanes spanning columns, so every global access is a full 128-byte row segment.
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = ew >> 2;                      // which half of the BN columns
    float4* slab = reinterpret_cast<float4*>(smem + S::EPI_OFF) + ew * 256;     // 32 rows x 8 float4
    const Epilogue& ep = args.ep;
    const bool partial = args.splits > 1;
    const int sub_r = lane >> 3;                   // coalesced phase: 4 rows x 8 lanes x float4
    const int sub_g = lane & 7;
    float rs0 = 1.f, rs1 = 1.f;
    if (!partial && ep.row_scale2) { rs0 = __ldg(ep.row_scale2); rs1 = __ldg(ep.row_scale2 + 1); }
    auto al16 = [](const void* p, int ld) { return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    // fast path: 16-byte aligned everything, at most ONE addend matrix (the second addend and the rank-1
    // term only occur in small / split-K GEMMs, which finish in the generic path or the reduce kernel)
    const bool vec_all = partial ? true
                                 : (al16(ep.out, ep.ldo) && (!ep.c1 || al16(ep.c1, ep.ldc1)) && !ep.c2 &&
                                    (!ep.bias || al16(ep.bias, 4)) && !ep.r1_row &&
                                    (!ep.adam_m || (al16(ep.adam_m, 4) && al16(ep.adam_v, 4))));
    const bool use_c1 = !partial && ep.c1 != nullptr;
    float sq0 = 0.f, sq1 = 0.f;
    for (uint32_t ui = 0;; ++ui) {
      const int sl = ui % TC_SCHED;
      ptx::mbar_wait(&sched_full[sl], (ui / TC_SCHED) & 1);
      const int u = sched_unit[sl];
      __syncwarp();                                   // every lane has read the id before the slot is released
      if (lane == 0) ptx::mbar_arrive(&sched_empty[sl]);
      if (u < 0) break;
      const TcUnit un = tc_unit(u, tiles_m, tiles_n, m_fastest, BN, BM);
      const uint32_t acc = ui % ACC;
      // A ragged last column tile still takes the vector path under the layout convention (ld = roundup(cols, 32),
      // zero padding): chunks are 32 columns wide and start at multiples of 32, so every chunk that starts below N
      // lies inside the padded row (chunks at or beyond N are skipped below); the accumulator columns >= N are
      // exactly zero (TMA zero-fills the operand beyond its true extent), so are the padding columns of addends and
      // biases, and zeros land in padding that is zero anyway.  Without this every GEMM whose N is not a multiple of
      // 256 (I = 200 000 or 25 000 per rank, k = 250) finishes on a column of tiles in the scalar path, ~10x slower.
      // Split-K partials: the workspace rows are padded to 32 floats for the same reason (ws_ld).
      const bool n_pad_ok = partial ? (args.ws_ld & 31) == 0
                                    : (ep.ldo == ((args.N + 31) & ~31) && (!ep.c1 || ep.ldc1 >= ep.ldo) &&
                                       ep.act != ACT_SIGMOID);
      const bool interior = vec_all && (un.m0 + row_off + MT * TC_BM <= args.M) && (un.n0 + BN <= args.N || n_pad_ok);
      // this warp's chunks of the unit: 32 rows x 32 columns each, NCH = MT * BN / 64 of them
      constexpr int CH = BN / 64, NCH = MT * CH;
      auto chunk_mw = [&](int j) { return un.m0 + row_off + (j / CH) * TC_BM + q * 32; };
      auto chunk_c = [&](int j) { return half * (BN / 2) + (j % CH) * 32; };
      // Addend tile (residual / gradient GEMMs): its loads are issued one chunk AHEAD -- the first chunk's
      // before the accumulator is even waited for, chunk j+1's row group right after chunk j's has been
      // consumed -- so the HBM latency (ncu: 30 % of the epilogue warps' time when paid per chunk) hides
      // under the TMEM read, the transpose and the stores of the chunk in between.
      const bool pf = interior && use_c1 && !(ep.adam_m && !LEAN);
      auto chunk_live = [&](int j) { return un.n0 + chunk_c(j) < args.N; };   // (a ragged tile skips its last chunks)
      int t1_for = -1;                              // chunk whose addend rows sit in t1
      float4 t1[8];
      auto load_c1 = [&](int j, int itr) {
        const float* p1 = ep.c1 + (size_t)(chunk_mw(j) + sub_r + itr * 4) * ep.ldc1 + un.n0 + chunk_c(j) + sub_g * 4;
        t1[itr] = __ldg(reinterpret_cast<const float4*>(p1));
      };
      if (pf && chunk_live(0)) {
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) load_c1(0, itr);
        t1_for = 0;
      }
      ptx::mbar_wait(&tmem_full_bar[acc], (ui / ACC) & 1);
      ptx::tc_fence_after();
      for (int j = 0; j < NCH; ++j) {
        const int mt = j / CH, c = chunk_c(j);
        const uint32_t taddr = tmem_base + acc * ACC_COLS + mt * BN + ((uint32_t)(q * 32) << 16);
        const int mw = chunk_mw(j);                   // first row of this warp's 32-row slab
        uint32_t r[32];
        if (args.dbg_epi < 2) {
          ptx::tmem_ld_32x32(taddr + (uint32_t)c, r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) r[jj] = 0u;
        }
        if (j == NCH - 1) {                         // last TMEM read of this warp for this unit
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty_bar[acc]), 0u));
            else ptx::mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        const int nb = un.n0 + c;
        if (nb >= args.N || mw >= args.M || args.dbg_epi >= 1) continue;          // warp-uniform
        // thread = row: 8 float4 stores, column group g lands in slot g ^ (row & 7) (conflict free)
#pragma unroll
        for (int g = 0; g < 8; ++g)
          slab[lane * 8 + (g ^ (lane & 7))] =
              make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                          __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
        __syncwarp();
        const int n = nb + sub_g * 4;
        if (interior) {
          // ---- fast path: whole tile in range, everything 16-byte aligned
          if (partial) {
            float* dst = args.ws + ((size_t)un.split * args.M + mw + sub_r) * args.ws_ld + n;
            const size_t step = (size_t)4 * args.ws_ld;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int row = itr * 4 + sub_r;
              *reinterpret_cast<float4*>(dst) = slab[row * 8 + (sub_g ^ (row & 7))];
              dst += step;
            }
          } else {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
            float* dst = ep.out + (size_t)(mw + sub_r) * ep.ldo + n;
            if (!LEAN && ep.adam_m) {
              // fused Adam on the parameter tile: theta, m, v are read and written in place (24 B/param
              // instead of 28 + the gradient round trip of a separate optimiser kernel).  (Rotating these
              // loads half a chunk ahead, as the addend loads are, measured 12 % SLOWER: the 12 loads of a
              // half chunk issued back to back keep more of HBM busy than loads interleaved with stores.
              // Pulling the next chunk's lines into L2 with prefetch.global.L2 made no measurable difference.)
              const size_t off0 = (size_t)(mw + sub_r) * ep.ldo + n;
              float l2 = 0.f;
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {
                float4 th[4], mm[4], vv[4];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const size_t o = off0 + (size_t)(hb * 4 + i4) * 4 * ep.ldo;
                  th[i4] = *reinterpret_cast<const float4*>(ep.out + o);
                  mm[i4] = *reinterpret_cast<const float4*>(ep.adam_m + o);
                  vv[i4] = *reinterpret_cast<const float4*>(ep.adam_v + o);
                }
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const int itr = hb * 4 + i4;
                  const int row = itr * 4 + sub_r;
                  const float4 a = slab[row * 8 + (sub_g ^ (row & 7))];
                  const float sc = ep.alpha * ((mw + row) >= ep.row_split ? rs1 : rs0);
                  l2 += th[i4].x * th[i4].x + th[i4].y * th[i4].y + th[i4].z * th[i4].z + th[i4].w * th[i4].w;
                  adam_elem(sc * a.x, th[i4].x, mm[i4].x, vv[i4].x, ep.adam_alpha, ep.adam_reg);
                  adam_elem(sc * a.y, th[i4].y, mm[i4].y, vv[i4].y, ep.adam_alpha, ep.adam_reg);
                  adam_elem(sc * a.z, th[i4].z, mm[i4].z, vv[i4].z, ep.adam_alpha, ep.adam_reg);
                  adam_elem(sc * a.w, th[i4].w, mm[i4].w, vv[i4].w, ep.adam_alpha, ep.adam_reg);
                  const size_t o = off0 + (size_t)itr * 4 * ep.ldo;
                  *reinterpret_cast<float4*>(ep.out + o) = th[i4];
                  *reinterpret_cast<float4*>(ep.adam_m + o) = mm[i4];
                  *reinterpret_cast<float4*>(ep.adam_v + o) = vv[i4];
                }
              }
              sq0 += l2;
              __syncwarp();
              continue;
            }
            const bool more = j + 1 < NCH && chunk_live(j + 1);
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_c1 && t1_for != j) {               // (not prefetched: the chunk before it was skipped)
#pragma unroll
              for (int itr = 0; itr < 8; ++itr) load_c1(j, itr);
            }
            if (use_c1) t1_for = more ? j + 1 : -1;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int row = itr * 4 + sub_r;
              const float4 a = slab[row * 8 + (sub_g ^ (row & 7))];
              const bool lower = (mw + row) >= ep.row_split;
              const float sc = ep.alpha * (lower ? rs1 : rs0);
              float4 o = make_float4(fmaf(sc, a.x, b4.x), fmaf(sc, a.y, b4.y), fmaf(sc, a.z, b4.z),
                                     fmaf(sc, a.w, b4.w));
              if (use_c1) {
                o.x = fmaf(ep.beta1, t1[itr].x, o.x); o.y = fmaf(ep.beta1, t1[itr].y, o.y);
                o.z = fmaf(ep.beta1, t1[itr].z, o.z); o.w = fmaf(ep.beta1, t1[itr].w, o.w);
              }
              if (!LEAN && ep.act != ACT_LINEAR) {
                o.x = act_fwd(ep.act, o.x); o.y = act_fwd(ep.act, o.y);
                o.z = act_fwd(ep.act, o.z); o.w = act_fwd(ep.act, o.w);
              }
              if (!LEAN && ep.round_out) {
                o.x = ptx::round_tf32(o.x); o.y = ptx::round_tf32(o.y);
                o.z = ptx::round_tf32(o.z); o.w = ptx::round_tf32(o.w);
              }
              *reinterpret_cast<float4*>(dst + (size_t)itr * 4 * ep.ldo) = o;
              if (use_c1 && more) load_c1(j + 1, itr);         // next chunk's row group, a whole chunk ahead
              const float sq = o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
              if (lower) sq1 += sq; else sq0 += sq;
              if (LEAN) { cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w; }
            }
            if (LEAN && ep.colpart) {
              // this lane summed rows sub_r, sub_r + 4, ...; the four row phases meet through two shuffles
#pragma unroll
              for (int sh = 8; sh <= 16; sh <<= 1) {
                cs.x += __shfl_xor_sync(0xffffffffu, cs.x, sh); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, sh);
                cs.z += __shfl_xor_sync(0xffffffffu, cs.z, sh); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, sh);
              }
              if (sub_r == 0) *reinterpret_cast<float4*>(ep.colpart + (size_t)(mw >> 5) * ep.ldcp + n) = cs;
            }
          }
        } else {
          // ---- generic path: boundary tiles / unaligned views / rank-1 term
          float cs[4] = {0.f, 0.f, 0.f, 0.f};
          for (int itr = 0; itr < 8; ++itr) {
            const int row = itr * 4 + sub_r;
            const int m = mw + row;
            if (m >= args.M || n >= args.N) continue;
            const float4 a4 = slab[row * 8 + (sub_g ^ (row & 7))];
            const float v[4] = {a4.x, a4.y, a4.z, a4.w};
            if (partial) {
              float* dst = args.ws + ((size_t)un.split * args.M + m) * args.ws_ld + n;
              for (int e = 0; e < 4 && n + e < args.N; ++e) dst[e] = v[e];
              continue;
            }
            const float rs = m >= ep.row_split ? rs1 : rs0;
            float sq = 0.f;
            for (int e = 0; e < 4 && n + e < args.N; ++e) {
              const float x = apply_epilogue(ep, v[e], m, n + e, rs);
              if (LEAN) cs[e] += x;
              sq += finish_element(ep, x, m, n + e);
            }
            if (ep.adam_m || m < ep.row_split) sq0 += sq; else sq1 += sq;
          }
          if (LEAN && !partial && ep.colpart) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
            }
            if (sub_r == 0)
              for (int e = 0; e < 4 && n + e < args.N; ++e) ep.colpart[(size_t)(mw >> 5) * ep.ldcp + n + e] = cs[e];
          }
        }
        __syncwarp();                               // slab is reused by the next chunk
      }
    }
    if (!partial && (ep.sumsq2 || ep.adam_l2)) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
        sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
      }
      if (lane == 0) {
        if (ep.adam_m) {
          if (ep.adam_l2 && sq0 != 0.f) atomicAdd(ep.adam_l2, (double)sq0);
        } else {
          if (sq0 != 0.f) atomicAdd(ep.sumsq2, (double)sq0);
          if (sq1 != 0.f) atomicAdd(ep.sumsq2 + 1, (double)sq1);
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();                    // the leader's MMAs read the peer's shared memory until the end
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_pair(tmem_base, ACC * ACC_COLS);
    else ptx::tmem_dealloc(tmem_base, ACC * ACC_COLS);
  }
}

// Sum split-K partials in a fixed order (deterministic) and apply the epilogue.  One thread owns 4 consecutive
// columns of one row; the partials are fetched four splits at a time (independent 16-byte loads in flight:
// with one scalar load per split the kernel ran at 3.0 of the 6.5 TB/s) and added in split order.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                            int ws_ld, Epilogue ep) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int m = blockIdx.y;
  float sq = 0.f;
  if (n < N) {
    const size_t split_stride = (size_t)M * ws_ld;
    const float* p = ws + (size_t)m * ws_ld + n;              // ws_ld is a multiple of 4: 16-byte aligned
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(p + (size_t)(s + j) * split_stride));
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    for (; s < splits; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)s * split_stride));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float rs = ep.row_scale2 ? __ldg(ep.row_scale2 + (m >= ep.row_split ? 1 : 0)) : 1.f;
    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n + e < N) sq += finish_element(ep, apply_epilogue(ep, a4[e], m, n + e, rs), m, n + e);   // value^2 or theta_old^2
  }
  if (ep.sumsq2 || (ep.adam_m && ep.adam_l2)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0 && sq != 0.f) {
      if (ep.adam_m) atomicAdd(ep.adam_l2, (double)sq);
      else atomicAdd(ep.sumsq2 + (m >= ep.row_split ? 1 : 0), (double)sq);
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Tensor maps are pure functions of (pointer, extents, pitch, box, swizzle, dtype) and the operands of
// the training step live in persistent buffers, so the encoded descriptors are memoised: small-batch
// configurations are launch-bound and would otherwise re-encode ~26 maps per step on the host.
struct TmapKey {
  const void* ptr; int rows, cols, ld, box_cols, box_rows, dtype, swizzle;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows && dtype == o.dtype && swizzle == o.swizzle;
  }
};
struct TmapCache {
  static constexpr int N = 256;
  TmapKey keys[N];
  CUtensorMap maps[N];
  bool used[N] = {};
  static unsigned hash(const TmapKey& k) {
    unsigned long long h = (unsigned long long)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)k.rows * 0xBF58476D1CE4E5B9ull + (unsigned long long)k.cols * 0x94D049BB133111EBull +
         (unsigned)k.box_rows * 31u + (unsigned)k.swizzle;
    return (unsigned)(h >> 40);
  }
  const CUtensorMap* find(const TmapKey& k) const {
    const unsigned i = hash(k) % N;
    return used[i] && keys[i] == k ? &maps[i] : nullptr;
  }
  void put(const TmapKey& k, const CUtensorMap& m) {
    const unsigned i = hash(k) % N;          // direct-mapped: a collision simply evicts
    keys[i] = k; maps[i] = m; used[i] = true;
  }
};

// fp32 row-major matrix [rows][cols] with leading dimension ld (elements);
// box = {box_cols (contiguous), box_rows}; 128B swizzle; OOB reads give zeros.
inline int make_tmap_2d(CUtensorMap* map, const float* ptr, int rows, int cols, int ld,
                        int box_cols, int box_rows, int tmap_dtype,
                        int swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B, TmapCache* cache = nullptr) {
  const TmapKey key{ptr, rows, cols, ld, box_cols, box_rows, tmap_dtype, swizzle};
  if (cache) {
    if (const CUtensorMap* hit = cache->find(key)) { *map = *hit; return 0; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, (CUtensorMapDataType)tmap_dtype, 2, const_cast<float*>(ptr), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS && cache) cache->put(key, *map);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// [planes][rows][cols] fp32 (planes plane_bytes apart), box = {box_cols, box_rows, 1}, no swizzle: L2 prefetch only
inline int make_tmap_planes(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, size_t plane_bytes,
                            int planes, int box_cols, int box_rows, TmapCache* cache = nullptr) {
  const TmapKey key{ptr, rows, cols, ld, box_cols, box_rows, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT32, -(int)(plane_bytes >> 4) - planes};
  if (cache) {
    if (const CUtensorMap* hit = cache->find(key)) { *map = *hit; return 0; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return 1;
  if ((plane_bytes & 15) || (reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 3)) return 3;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)plane_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS && cache) cache->put(key, *map);
  return r == CUDA_SUCCESS ? 0 : 2;
}

struct TcGemmCall {
  const float* A; int lda; int a_mn;
  const float* B; int ldb; int b_mn;
  int M, N, K;
  Epilogue ep;
  int splits = 1;          // > 1 needs ws
  float* ws = nullptr;     // >= splits * M * roundup(N,32) floats
  int bn = 128;            // 128 or 256
  int mt = 1;              // 1: 128-row CTA tiles; 2: 256-row CTA tiles (needs bn == 256)
  int cg = 1;              // 2: CTA pairs (cta_group::2) on 256 x 256 tiles (needs bn == 256, mt == 1)
  // TFLOAT32 makes TMA round fp32 -> tf32 to nearest on the way into shared memory
  // (measured: FLOAT32 maps leave the bits alone and the MMA then truncates).
  int tmap_dtype = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  // bring-up overrides for the MN-major tile encoding (0 = use the defaults below)
  int dbg_mn_layout = 0, dbg_mn_sbo = 0, dbg_mn_lbo = 0, dbg_mn_swizzle = 0, dbg_epi = 0;
  TmapCache* cache = nullptr;
  int max_ctas = 0;        // > 0: cap on the persistent grid (SMs left free for a concurrent collective)
};

inline uint32_t make_idesc_tf32(int bn, int a_mn, int b_mn, int m = TC_BM) {
  // cute::UMMA::InstrDescriptor: c_format[4,6)=1 (F32) a_format[7,10)=2 (TF32)
  // b_format[10,13)=2 a_major bit15 b_major bit16 n_dim[17,23)=N>>3 m_dim[24,29)=M>>4
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(bn >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// Scheduler state for the next launch: TC_SCHED_SLOTS {next, done} pairs per device, used round-robin so
// GEMMs that overlap on different streams never share a pair; each pair is zero again when its kernel ends.
inline cudaError_t tc_sched_slot(unsigned int** out) {
  static unsigned int* base[64] = {};
  static unsigned int seq = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!base[dev]) {
    e = cudaMalloc((void**)&base[dev], TC_SCHED_SLOTS * 2 * sizeof(unsigned int));
    if (e != cudaSuccess) return e;
    e = cudaMemset(base[dev], 0, TC_SCHED_SLOTS * 2 * sizeof(unsigned int));
    if (e != cudaSuccess) return e;
  }
  *out = base[dev] + 2 * (seq++ % TC_SCHED_SLOTS);
  return cudaSuccess;
}

template <int BN, int STAGES, int MT, int CG, int EPI>
inline cudaError_t tc_gemm_launch_t(const TcGemmCall& c, TcGemmArgs& args, const CUtensorMap& ma,
                                    const CUtensorMap& mb, const CUtensorMap& mp, int splits, cudaStream_t stream) {
  using S = TcSmem<BN, STAGES, MT, CG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, MT, CG, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
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
  const int units = ((c.N + BN - 1) / BN) * ((c.M + CG * MT * TC_BM - 1) / (CG * MT * TC_BM)) * splits;
  const int sms = (c.max_ctas > 0 && c.max_ctas < num_sms) ? c.max_ctas : num_sms;
  if (CG == 2) {
    // one cluster of two CTAs per unit in flight; units are walked statically (cluster id + i * #clusters)
    const int pairs = units < sms / 2 ? units : sms / 2;
    args.sched = nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, STAGES, MT, CG, EPI>, ma, mb, mp, args);
  }
  const int grid = units < sms ? units : sms;
  static int static_sched = -1;          // GANMF_STATIC_SCHED=1: A/B switch, units blockIdx.x + i*gridDim.x
  if (static_sched < 0) {
    const char* e = getenv("GANMF_STATIC_SCHED");
    static_sched = (e && e[0] == '1') ? 1 : 0;
  }
  args.sched = nullptr;
  if (!static_sched && units > grid) {
    cudaError_t es = tc_sched_slot(&args.sched);
    if (es != cudaSuccess) return es;
  }
  tc_gemm_kernel<BN, STAGES, MT, CG, EPI><<<grid, TC_THREADS, S::TOTAL, stream>>>(ma, mb, mp, args);
  return cudaGetLastError();
}

// EPI = 1 kernels serve every epilogue without the fused optimiser, an activation or tf32 rounding of the output
inline bool tc_lean_epilogue(const Epilogue& ep) { return !ep.adam_m && ep.act == ACT_LINEAR && !ep.round_out; }

// Returns cudaSuccess or an error; never falls back to another implementation.
inline cudaError_t tc_gemm(const TcGemmCall& c, cudaStream_t stream) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) return cudaErrorInvalidValue;
  if ((c.lda & 3) || (c.ldb & 3)) return cudaErrorInvalidValue;   // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(c.A) & 15) || (reinterpret_cast<uintptr_t>(c.B) & 15))
    return cudaErrorInvalidValue;
  const int bn = c.bn == 256 ? 256 : 128;
  const int cg = (c.cg == 2 && bn == 256 && c.mt == 1) ? 2 : 1;
  CUtensorMap ma, mb;
  int rc;
  const int mn_swz = c.dbg_mn_swizzle ? c.dbg_mn_swizzle : (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const int k_swz = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  if (!c.a_mn) rc = make_tmap_2d(&ma, c.A, c.M, c.K, c.lda, TC_BK, TC_BM, c.tmap_dtype, k_swz, c.cache);
  else         rc = make_tmap_2d(&ma, c.A, c.K, c.M, c.lda, 32, TC_BK, c.tmap_dtype, mn_swz, c.cache);
  if (rc) return cudaErrorUnknown;
  if (!c.b_mn) rc = make_tmap_2d(&mb, c.B, c.N, c.K, c.ldb, TC_BK, bn / cg, c.tmap_dtype, k_swz, c.cache);
  else         rc = make_tmap_2d(&mb, c.B, c.K, c.N, c.ldb, 32, TC_BK, c.tmap_dtype, mn_swz, c.cache);
  if (rc) return cudaErrorUnknown;

  TcGemmArgs args;
  args.M = c.M; args.N = c.N; args.K = c.K;
  args.a_mn = c.a_mn; args.b_mn = c.b_mn;
  const int total_kb = (c.K + TC_BK - 1) / TC_BK;
  int splits = c.splits < 1 ? 1 : c.splits;
  if (splits > total_kb) splits = total_kb;
  int kbps = (total_kb + splits - 1) / splits;
  splits = (total_kb + kbps - 1) / kbps;          // no empty split is ever launched
  if (splits > 1 && !c.ws) return cudaErrorInvalidValue;
  args.kb_per_split = kbps;
  args.splits = splits;
  args.dbg_epi = c.dbg_epi;
  args.ws = c.ws;
  args.ws_ld = (c.N + 31) & ~31;       // rows padded like every matrix: ragged tiles store whole 32-column chunks
  // shared-memory matrix descriptors (bytes >> 4):
  //  K-major : SWIZZLE_128B; rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (=1).
  //  MN-major: SWIZZLE_128B_BASE32B (TMA 128B_ATOM_32B); each TMA box is 32 k-rows of 128 B
  //            (= 32 MN elements); 4-k groups 512 B apart (SBO); 32-element MN chunks one
  //            box = TC_BK*128 B apart (LBO); one UMMA (K=8) consumes 1024 B of a box.
  const uint32_t mn_layout = c.dbg_mn_layout ? c.dbg_mn_layout : 1;
  const uint32_t mn_sbo = (c.dbg_mn_sbo ? c.dbg_mn_sbo : 512) >> 4;
  const uint32_t mn_lbo = (c.dbg_mn_lbo ? c.dbg_mn_lbo : TC_BK * 128) >> 4;
  args.a_layout = c.a_mn ? mn_layout : 2;
  args.b_layout = c.b_mn ? mn_layout : 2;
  args.a_lbo = c.a_mn ? mn_lbo : 1;
  args.a_sbo = c.a_mn ? mn_sbo : 1024 >> 4;
  args.b_lbo = c.b_mn ? mn_lbo : 1;
  args.b_sbo = c.b_mn ? mn_sbo : 1024 >> 4;
  // sub-tile mt of A starts TC_BM rows further: K-major rows are 128 B apart, MN-major 32-row chunks
  // are one TMA box (TC_BK*128 B) apart
  args.a_mtstep = c.a_mn ? ((TC_BM / 32) * TC_BK * 128) >> 4 : (TC_BM * TC_BK * 4) >> 4;
  args.a_kstep = c.a_mn ? (1024 >> 4) : (TC_UMMA_K * 4) >> 4;
  args.b_kstep = c.b_mn ? (1024 >> 4) : (TC_UMMA_K * 4) >> 4;
  args.idesc = make_idesc_tf32(bn, c.a_mn, c.b_mn, cg * TC_BM);
  args.ep = c.ep;

  // fused optimiser: theta, m, v as the planes of ONE 3D tensor map when they sit a constant stride apart (they do:
  // [theta | m | v] slabs), box = {BN columns, 128 rows, 1 plane}; prefetched into L2 by the producer, unit by unit.
  // OFF by default: measured at cfg5 it makes both weight-gradient GEMMs 0.24 ms SLOWER (1.56 -> 1.80, 1.50 -> 1.74 ms,
  // profiles/r02b_ab_adam_prefetch.txt) -- 57 MB of prefetched state in flight push the operand tiles out of L2.
  // GANMF_ADAM_PREFETCH=1 turns it on (A/B switch).
  CUtensorMap mp = ma;
  args.pf_planes = 0;
  static int adam_pf = -1;
  if (adam_pf < 0) { const char* ev = getenv("GANMF_ADAM_PREFETCH"); adam_pf = (ev && ev[0] == '1') ? 1 : 0; }
  if (adam_pf && c.ep.adam_m && c.ep.adam_v && splits == 1 && c.ep.adam_m > c.ep.out &&
      (c.ep.adam_v - c.ep.adam_m) == (c.ep.adam_m - c.ep.out)) {
    const size_t plane = (size_t)(c.ep.adam_m - c.ep.out) * 4;
    if (make_tmap_planes(&mp, c.ep.out, c.M, c.N, c.ep.ldo, plane, 3, bn, TC_BM, c.cache) == 0) args.pf_planes = 3;
    else mp = ma;
  }
  cudaError_t e;
  const bool lean = tc_lean_epilogue(c.ep);
#define TC_LAUNCH(BN_, ST_, MT_, CG_)                                                            \
  (lean ? tc_gemm_launch_t<BN_, ST_, MT_, CG_, 1>(c, args, ma, mb, mp, splits, stream)           \
        : tc_gemm_launch_t<BN_, ST_, MT_, CG_, 0>(c, args, ma, mb, mp, splits, stream))
  if (cg == 2)                     e = TC_LAUNCH(256, 6, 1, 2);
  else if (bn == 256 && c.mt == 2) e = TC_LAUNCH(256, 3, 2, 1);
  else if (bn == 256)              e = TC_LAUNCH(256, 4, 1, 1);
  else                             e = TC_LAUNCH(128, 6, 1, 1);
#undef TC_LAUNCH
  if (e != cudaSuccess) return e;
  if (splits > 1) {
    const int rt = c.N >= 1024 ? 256 : (c.N >= 512 ? 128 : 64);      // threads per block, 4 columns each
    dim3 rb(rt), rg((c.N + 4 * rt - 1) / (4 * rt), c.M);
    splitk_reduce_kernel<<<rg, rb, 0, stream>>>(c.ws, splits, c.M, c.N, args.ws_ld, c.ep);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace ganmf
