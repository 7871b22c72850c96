// HBM-bound and small-shape kernels of the GANMF / DisGANMF path (sm_100a):
//   K1  csr_gather_dense      CSR rows -> dense fp32 profiles        (GANMF.py:184 URM[uids].toarray())
//   --  gather_rows           P[uids]                               (GANMF.py:82 embedding_lookup)
//   K6  fused_adam            one launch over all tensors of an optimiser (GANMF.py:104-105,138-139)
//   --  colsum / scale_rows / sqdiff / hinge_gate / act_grad        bias grads, loss scalars
//   --  simt_gemm             exact-fp32 GEMM for skinny shapes (DisGANMF d_nodes=4, N=1) and checks
// Layout convention: every matrix is row-major fp32 with leading dimension ld = roundup(cols, 32);
// padding columns are kept at zero by construction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "tc_gemm.cuh"

namespace ganmf {

constexpr float ADAM_B1 = 0.9f, ADAM_B2 = 0.999f, ADAM_EPS = 1e-8f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------- K1
// grid = (col_chunks, B). Each CTA zero-fills its column range of one output row with 16-byte
// stores, then scatters the row's non-zeros that fall into the range.  Algorithmic bytes:
// 4*B*ld written + 8*nnz(batch) read.
__global__ void csr_gather_dense_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                        const float* __restrict__ data, const int* __restrict__ row_ids,
                                        float* __restrict__ out, int ld, int col_off, int chunk) {
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * chunk;
  int c1 = c0 + chunk;
  if (c1 > ld) c1 = ld;
  float* row = out + (size_t)b * ld;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = c0 + threadIdx.x * 4; c < c1; c += blockDim.x * 4) *reinterpret_cast<float4*>(row + c) = z;
  __syncthreads();
  const int r = row_ids[b];
  const int s = indptr[r], e = indptr[r + 1];
  for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
    const int c = indices[i] + col_off;
    if (c >= c0 && c < c1) row[c] = data ? data[i] : 1.0f;
  }
}

inline cudaError_t csr_gather_dense(const int* indptr, const int* indices, const float* data,
                                    const int* row_ids, int B, float* out, int ld, int col_off,
                                    cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  // aim for >= 4 CTAs per SM; chunk is a multiple of 1024 floats (4 KB)
  int chunks = (148 * 4 + B - 1) / B;
  int chunk = ((ld + chunks - 1) / chunks + 1023) / 1024 * 1024;
  chunks = (ld + chunk - 1) / chunk;
  csr_gather_dense_kernel<<<dim3(chunks, B), 256, 0, st>>>(indptr, indices, data, row_ids, out, ld,
                                                          col_off, chunk);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------- K1b
// Sparse real-profile encode: out[b, :] = bias + sum_{j in CSR row ids[b]} data[j] * W[indices[j], :], i.e. the real
// half of H = X . We + be (GANMF.py:64) as a gather-sum over the row's interactions instead of a dense
// [B, I] x [I, E] product: at 0.1 % density a row touches 200 of the 200 000 rows of We (800 KB read against
// 0.41 GFLOP of dense MMA work).  Exact fp32 FMAs in CSR order (deterministic).  grid = (column chunks, B); one
// thread owns 4 consecutive columns; the row's (index, value) pairs are staged through shared memory and the
// weight rows are fetched eight at a time (independent 16-byte loads in flight).  HBM-bound:
// algorithmic bytes = 4 * E * nnz(batch) read + 4 * B * E written.
constexpr int ENC_THREADS = 256, ENC_UNROLL = 8;
__global__ void __launch_bounds__(ENC_THREADS)
csr_encode_rows_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
                       const int* __restrict__ row_ids, const float* __restrict__ W, int ldw, int E4,
                       const float* __restrict__ bias, float* __restrict__ out, int ldo) {
  __shared__ int s_idx[ENC_THREADS];
  __shared__ float s_val[ENC_THREADS];
  const int b = blockIdx.y;
  const int c4 = blockIdx.x * ENC_THREADS + threadIdx.x;
  const bool live = c4 < E4;
  const int r = row_ids ? row_ids[b] : b;
  const int s = indptr[r], e = indptr[r + 1];
  const float4* w4 = reinterpret_cast<const float4*>(W) + (live ? c4 : 0);
  const size_t ldw4 = (size_t)(ldw >> 2);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias && live) acc = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  for (int base = s; base < e; base += ENC_THREADS) {
    const int cnt = min(ENC_THREADS, e - base);
    __syncthreads();                                   // the previous batch of pairs has been consumed
    if ((int)threadIdx.x < cnt) {
      s_idx[threadIdx.x] = indices[base + threadIdx.x];
      s_val[threadIdx.x] = data ? data[base + threadIdx.x] : 1.0f;
    }
    __syncthreads();
    if (!live) continue;
    int j = 0;
    for (; j + ENC_UNROLL <= cnt; j += ENC_UNROLL) {
      float4 w[ENC_UNROLL];
#pragma unroll
      for (int u = 0; u < ENC_UNROLL; ++u) w[u] = __ldg(w4 + (size_t)s_idx[j + u] * ldw4);
#pragma unroll
      for (int u = 0; u < ENC_UNROLL; ++u) {
        const float v = s_val[j + u];
        acc.x = fmaf(v, w[u].x, acc.x); acc.y = fmaf(v, w[u].y, acc.y);
        acc.z = fmaf(v, w[u].z, acc.z); acc.w = fmaf(v, w[u].w, acc.w);
      }
    }
    for (; j < cnt; ++j) {
      const float4 w = __ldg(w4 + (size_t)s_idx[j] * ldw4);
      const float v = s_val[j];
      acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
      acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
    }
  }
  if (live) *(reinterpret_cast<float4*>(out + (size_t)b * ldo) + c4) = acc;
}

// W: [items][ldw] (ldw and ldo multiples of 4, 16-byte aligned bases); E_pad = columns to produce (multiple of 4;
// padding columns of W and bias are zero, so the padding of out stays zero)
inline cudaError_t csr_encode_rows(const int* indptr, const int* indices, const float* data, const int* row_ids,
                                   int B, const float* W, int ldw, int E_pad, const float* bias, float* out, int ldo,
                                   cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const int E4 = E_pad >> 2;
  csr_encode_rows_kernel<<<dim3((E4 + ENC_THREADS - 1) / ENC_THREADS, B), ENC_THREADS, 0, st>>>(
      indptr, indices, data, row_ids, W, ldw, E4, bias, out, ldo);
  return cudaGetLastError();
}

// out[b, :] = src[ids[b], :]   (ld multiple of 4)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ ids,
                                   float* __restrict__ out, int ld) {
  const int b = blockIdx.x;
  const float4* s = reinterpret_cast<const float4*>(src + (size_t)ids[b] * ld);
  float4* d = reinterpret_cast<float4*>(out + (size_t)b * ld);
  for (int c = threadIdx.x; c < ld / 4; c += blockDim.x) d[c] = s[c];
}

// Split-TF32 operands for fp32-accurate scoring on the tensor cores: x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi).  A rows become [hi | hi | lo], B rows [hi | lo | hi] (3k columns), so one GEMM
// over K' = 3k yields hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-22 relative).
// src = base[ids[r]] when ids != nullptr (gather fused), else base[r].
__global__ void split3_rows_kernel(const float* __restrict__ src, int ld_src, const int* __restrict__ ids,
                                   float* __restrict__ dst, int ld_dst, int k, int b_layout) {
  const int r = blockIdx.x;
  const float* s = src + (size_t)(ids ? ids[r] : r) * ld_src;
  float* d = dst + (size_t)r * ld_dst;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float x = s[j];
    const float hi = ptx::round_tf32(x);
    const float lo = ptx::round_tf32(x - hi);
    d[j] = hi;
    d[k + j] = b_layout ? lo : hi;
    d[2 * k + j] = b_layout ? hi : lo;
  }
}

// The same split for ANY operand of a GEMM (split-TF32 "precise" mode of the training GEMMs): src is the logical
// [MN][K] operand, stored K-major ([MN][K], ld >= K) or MN-major ([K][MN], ld >= MN); dst is K-major
// [MN][3K] = [hi | hi | lo] (b_layout = 0) or [hi | lo | hi] (b_layout = 1).  32 x 32 tiles through shared
// memory so both the read and the write are coalesced whatever the source major-ness.
__global__ void __launch_bounds__(256)
split3_any_kernel(const float* __restrict__ src, int ld_src, int mn_major, int MN, int K, float* __restrict__ dst,
                  int ld_dst, int b_layout) {
  __shared__ float tile[32][33];
  const int mn0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    float x = 0.f;
    if (mn_major) {                                                   // rows of src = k, contiguous = mn
      const int kk = k0 + r, mn = mn0 + tx;
      if (kk < K && mn < MN) x = src[(size_t)kk * ld_src + mn];
      tile[tx][r] = x;                                                // tile[mn][k]
    } else {                                                          // rows of src = mn, contiguous = k
      const int mn = mn0 + r, kk = k0 + tx;
      if (kk < K && mn < MN) x = src[(size_t)mn * ld_src + kk];
      tile[r][tx] = x;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int mn = mn0 + r, kk = k0 + tx;
    if (mn >= MN || kk >= K) continue;
    const float x = tile[r][tx];
    const float hi = ptx::round_tf32(x);
    const float lo = ptx::round_tf32(x - hi);
    float* d = dst + (size_t)mn * ld_dst;
    d[kk] = hi;
    d[K + kk] = b_layout ? lo : hi;
    d[2 * K + kk] = b_layout ? hi : lo;
  }
}

// first column of the DisGANMF discriminator input: float(row id)  (DisGANMF.py:110-111)
// (ids are LOCAL rows of this GPU's shard; id_offset = global id of local row 0)
__global__ void ids_to_float_kernel(const int* __restrict__ ids, float* __restrict__ out, int B, int id_offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = (float)(ids[i] + id_offset);
}

__global__ void set_slots_kernel(int* __restrict__ slot, const int* __restrict__ ids, int B, int reset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) slot[ids[i]] = reset ? -1 : i;
}

__global__ void fill_int_kernel(int* p, int v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ----------------------------------------------------------------------------- K6
// TF-1.12 ApplyAdam, dense, one launch for all tensors of one optimiser:
//   g' = g + reg*theta; m += (g'-m)(1-b1); v += (g'^2-v)(1-b2); theta -= alpha*m/(sqrt(v)+eps)
// A segment may take its gradient from a compact [B, ld] block through a row->slot map (the
// user-factor matrix: rows outside the minibatch have zero data gradient but are still updated,
// as TF's dense update does).  Also accumulates sum(theta_old^2) for the l2 term of the loss.
// Algorithmic bytes: 28 per parameter (read theta,m,v,g; write theta,m,v).
constexpr int ADAM_MAX_SEG = 8;
struct AdamSeg {
  float* theta; float* m; float* v;
  const float* g;          // dense gradient, or compact rows when slot != nullptr
  const int* slot;         // row -> compact row (or -1)
  int ld;                  // row length (slot segments)
  unsigned long long n4;   // float4 count
};
struct AdamArgs {
  AdamSeg seg[ADAM_MAX_SEG];
  unsigned long long blk_begin[ADAM_MAX_SEG + 1];   // first block of each segment
  int nseg;
  float alpha, reg;
  double* l2_out;          // += sum theta_old^2 (nullable)
  double* l2_shard_out;    // same, for segments that carry a slot map (row-sharded under DP)
};
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_THREAD = 4;     // float4s per thread

__global__ void __launch_bounds__(ADAM_THREADS) fused_adam_kernel(const AdamArgs a) {
  int s = 0;
  while (s + 1 < a.nseg && blockIdx.x >= a.blk_begin[s + 1]) ++s;
  const AdamSeg sg = a.seg[s];
  const unsigned long long base =
      (blockIdx.x - a.blk_begin[s]) * (unsigned long long)(ADAM_THREADS * ADAM_VEC_PER_THREAD);
  float4* th = reinterpret_cast<float4*>(sg.theta);
  float4* mm = reinterpret_cast<float4*>(sg.m);
  float4* vv = reinterpret_cast<float4*>(sg.v);
  const float4* gg = reinterpret_cast<const float4*>(sg.g);
  const int ld4 = sg.ld >> 2;
  float sq = 0.f;
  float4 t[ADAM_VEC_PER_THREAD], m[ADAM_VEC_PER_THREAD], v[ADAM_VEC_PER_THREAD], g[ADAM_VEC_PER_THREAD];
#pragma unroll
  for (int j = 0; j < ADAM_VEC_PER_THREAD; ++j) {          // all loads first (MLP), then math
    const unsigned long long i = base + j * ADAM_THREADS + threadIdx.x;
    if (i < sg.n4) {
      t[j] = th[i]; m[j] = mm[i]; v[j] = vv[i];
      if (sg.slot) {
        // row of the factor matrix this float4 belongs to (32-bit division whenever the segment allows it)
        const unsigned long long row = sg.n4 < 0xFFFFFFFFull ? (unsigned long long)((unsigned)i / (unsigned)ld4)
                                                             : i / (unsigned long long)ld4;
        const int sl = __ldg(sg.slot + row);
        g[j] = sl >= 0 ? gg[(unsigned long long)sl * ld4 + (i - row * ld4)] : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        g[j] = gg[i];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < ADAM_VEC_PER_THREAD; ++j) {
    const unsigned long long i = base + j * ADAM_THREADS + threadIdx.x;
    if (i < sg.n4) {
      float* tp = reinterpret_cast<float*>(&t[j]);
      float* mp = reinterpret_cast<float*>(&m[j]);
      float* vp = reinterpret_cast<float*>(&v[j]);
      const float* gp = reinterpret_cast<const float*>(&g[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sq += tp[e] * tp[e];
        adam_elem(gp[e], tp[e], mp[e], vp[e], a.alpha, a.reg);
      }
      th[i] = t[j]; mm[i] = m[j]; vv[i] = v[j];
    }
  }
  if (a.l2_out) {
    __shared__ float red[ADAM_THREADS / 32];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x < 32) {
      float x = threadIdx.x < ADAM_THREADS / 32 ? red[threadIdx.x] : 0.f;
      x = warp_sum(x);
      if (threadIdx.x == 0) atomicAdd((sg.slot && a.l2_shard_out) ? a.l2_shard_out : a.l2_out, (double)x);
    }
  }
}

inline cudaError_t fused_adam(AdamArgs& a, cudaStream_t st) {
  unsigned long long blk = 0;
  for (int s = 0; s < a.nseg; ++s) {
    a.blk_begin[s] = blk;
    const unsigned long long per = ADAM_THREADS * ADAM_VEC_PER_THREAD;
    blk += (a.seg[s].n4 + per - 1) / per;
  }
  a.blk_begin[a.nseg] = blk;
  if (blk == 0) return cudaSuccess;
  fused_adam_kernel<<<(unsigned)blk, ADAM_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------- K6b lazy user factors
// TF's Adam is dense over the user-factor matrix P: with g_reg = 0 a row outside the minibatch still takes
// the zero-gradient update  m *= b1-ish, v *= b2-ish, theta -= alpha_t * m / (sqrt(v) + eps)  every G
// step (28 B/parameter of HBM traffic for rows whose values nobody reads until they are sampled again).
// These updates depend only on the row's own (theta, m, v) and the step sizes alpha_t, so they are
// DEFERRED: last[row] records the G step the row is current at, alpha_log keeps alpha_t, and a row replays
// the steps it missed -- the same adam_elem() calls in the same order, in registers -- right before it is
// read (minibatch gather, scoring, snapshot, export).  Bit-identical to the dense sweep by construction.
__global__ void __launch_bounds__(64) p_catchup_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                       float* __restrict__ v, int ld, const int* __restrict__ ids,
                                                       int n, int* __restrict__ last,
                                                       const float* __restrict__ alpha_log, int T, int log_base) {
  const int ld4 = ld >> 2;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    const int row = ids ? ids[b] : b;
    const int t0 = last[row];
    __syncthreads();                                  // everyone has read last[row] before it is advanced
    if (t0 >= T) continue;                            // block-uniform
    float4* th4 = reinterpret_cast<float4*>(theta + (size_t)row * ld);
    float4* m4 = reinterpret_cast<float4*>(m + (size_t)row * ld);
    float4* v4 = reinterpret_cast<float4*>(v + (size_t)row * ld);
    for (int c = threadIdx.x; c < ld4; c += blockDim.x) {
      float4 mm = m4[c], vv = v4[c];
      const bool idle = mm.x == 0.f && mm.y == 0.f && mm.z == 0.f && mm.w == 0.f && vv.x == 0.f && vv.y == 0.f &&
                        vv.z == 0.f && vv.w == 0.f;
      if (idle) continue;                             // never-sampled row: m = v = 0 is a fixed point
      float4 tt = th4[c];
      for (int t = t0; t < T; ++t) {
        const float a = __ldg(alpha_log + (t - log_base));
        adam_elem(0.f, tt.x, mm.x, vv.x, a, 0.f);
        adam_elem(0.f, tt.y, mm.y, vv.y, a, 0.f);
        adam_elem(0.f, tt.z, mm.z, vv.z, a, 0.f);
        adam_elem(0.f, tt.w, mm.w, vv.w, a, 0.f);
      }
      th4[c] = tt; m4[c] = mm; v4[c] = vv;
    }
    if (threadIdx.x == 0) last[row] = T;
  }
}

// G step T+1 on the (current) minibatch rows with their data gradient g[b, :]; logs alpha_{T+1}.
__global__ void __launch_bounds__(64) p_batch_adam_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                          float* __restrict__ v, int ld, const int* __restrict__ ids,
                                                          const float* __restrict__ g, int ldg, float alpha,
                                                          int* __restrict__ last, float* __restrict__ alpha_log,
                                                          int T, int log_base) {
  const int row = ids[blockIdx.x];
  const int ld4 = ld >> 2;
  float4* th4 = reinterpret_cast<float4*>(theta + (size_t)row * ld);
  float4* m4 = reinterpret_cast<float4*>(m + (size_t)row * ld);
  float4* v4 = reinterpret_cast<float4*>(v + (size_t)row * ld);
  const float4* g4 = reinterpret_cast<const float4*>(g + (size_t)blockIdx.x * ldg);
  for (int c = threadIdx.x; c < ld4; c += blockDim.x) {
    float4 tt = th4[c], mm = m4[c], vv = v4[c];
    const float4 gg = g4[c];
    adam_elem(gg.x, tt.x, mm.x, vv.x, alpha, 0.f);
    adam_elem(gg.y, tt.y, mm.y, vv.y, alpha, 0.f);
    adam_elem(gg.z, tt.z, mm.z, vv.z, alpha, 0.f);
    adam_elem(gg.w, tt.w, mm.w, vv.w, alpha, 0.f);
    th4[c] = tt; m4[c] = mm; v4[c] = vv;
  }
  if (threadIdx.x == 0) {
    last[row] = T + 1;
    if (blockIdx.x == 0) alpha_log[T - log_base] = alpha;
  }
}

// ----------------------------------------------------------------------------- reductions
// out[n] = sum_m w(m) * X[m, n],  w(m) = row_w ? row_w[m] : 1, times row_scale2[m >= row_split].
// blockDim = (32, 8): 32 consecutive columns, 8 row lanes; deterministic order.
__global__ void colsum_kernel(const float* __restrict__ X, int M, int N, int ld,
                              const float* __restrict__ row_scale2, int row_split,
                              const float* __restrict__ row_w, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N) {
    const float s0 = row_scale2 ? row_scale2[0] : 1.f, s1 = row_scale2 ? row_scale2[1] : 1.f;
    for (int m = threadIdx.y; m < M; m += 8) {
      float w = m >= row_split ? s1 : s0;
      if (row_w) w *= row_w[m];
      acc += w * X[(size_t)m * ld + n];
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    out[n] = t;
  }
}

// The same weighted column sums from the per-32-row partial sums a GEMM epilogue left behind
// (Epilogue::colpart: part[p][n] = sum of out[32p .. 32p+31][n]): out[n] = rs[0] * sum_{p < split_part} part[p][n] +
// rs[1] * sum_{p >= split_part} part[p][n].  Reads M/32 rows instead of M (the decoder-bias gradient no longer
// costs a pass over the [2B, I] residual); fixed summation order.
__global__ void __launch_bounds__(256)
colsum_parts_kernel(const float* __restrict__ part, int nparts, int split_part, int N, int ld,
                    const float* __restrict__ row_scale2, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s0 = 0.f, s1 = 0.f;
  for (int p = 0; p < split_part && p < nparts; ++p) s0 += part[(size_t)p * ld + n];
  for (int p = split_part; p < nparts; ++p) s1 += part[(size_t)p * ld + n];
  out[n] = row_scale2[0] * s0 + row_scale2[1] * s1;
}

// out[e] = sum_i W[e, i] * x[i]   (fp32; ld multiple of 4).
// Bias gradient of the encoder without tensor-core rounding: dbe = dbd . Wd^T, because
// colsum(rs * Res2 . Wd^T) = (sum_m rs(m) Res2[m, :]) . Wd^T  and the bracket IS dbd (GANMF.py:64-68).
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ W, int rows, int cols, int ld, const float* __restrict__ x,
              float* __restrict__ out) {
  // one CTA per row of W: 8 warps stream the row with 16-byte loads, fixed-order block reduction
  const int e = blockIdx.x;
  const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)e * ld);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float acc = 0.f;
  const int n4 = cols >> 2;
  int i = threadIdx.x;
  for (; i + 3 * 256 < n4; i += 4 * 256) {          // four independent 16-byte loads of W in flight per thread
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = w4[i + u * 256]; b[u] = __ldg(x4 + i + u * 256); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc = fmaf(a[u].x, b[u].x, acc); acc = fmaf(a[u].y, b[u].y, acc);
      acc = fmaf(a[u].z, b[u].z, acc); acc = fmaf(a[u].w, b[u].w, acc);
    }
  }
  for (; i < n4; i += 256) {
    const float4 a = w4[i], b = __ldg(x4 + i);
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < cols; i += 256) acc = fmaf(W[(size_t)e * ld + i], x[i], acc);
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j];
    out[e] = t;
  }
}

// Y[m, :] = X[m, :] * row_scale2[m >= row_split]
__global__ void scale_rows_kernel(const float* __restrict__ X, float* __restrict__ Y, int ld4,
                                  const float* __restrict__ row_scale2, int row_split) {
  const int m = blockIdx.y;
  const float s = row_scale2[m >= row_split ? 1 : 0];
  const float4* x = reinterpret_cast<const float4*>(X) + (size_t)m * ld4;
  float4* y = reinterpret_cast<float4*>(Y) + (size_t)m * ld4;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld4; c += gridDim.x * blockDim.x) {
    float4 v = x[c];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    y[c] = v;
  }
}

// *out += sum (A - B)^2 over [M, N]  (feature-matching loss, GANMF.py:134)
__global__ void sqdiff_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int M, int N,
                              int ld, double* out) {
  float acc = 0.f;
  for (int m = blockIdx.x; m < M; m += gridDim.x)
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float d = A[(size_t)m * ld + n] - Bm[(size_t)m * ld + n];
      acc += d * d;
    }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) t += red[j];
    atomicAdd(out, (double)t);
  }
}

// Device-side scalars of one step (no host round trip between forward and backward).
struct StepScalars {
  double sumsq[2];     // sum (Dr-R)^2, sum (Df-F)^2          (GEMM epilogue accumulates)
  double fm;           // sum (Hr-Hf)^2 / sum (feat_r-feat_f)^2
  double l2;           // sum theta^2 over the optimiser's tensors (Adam kernel accumulates)
  double bce[2];       // DisGANMF: sum softplus(-out_r), sum softplus(out_f)
  double l2_shard;     // sum theta^2 of row-sharded tensors (user factors): summed over ranks under DP
  float row_scale[2];  // GANMF D-step: cr' = (1+g*m)*2/N, cf' = -g*2/N
  float loss_main;     // loss without the l2 term
  float pad;
};

// dloss pieces + hinge gate (GANMF.py:131-132): gate = 1[m*Lr - Lf > 0] (strict).
__global__ void hinge_gate_kernel(StepScalars* s, float m_hinge, double n_elems) {
  const double Lr = s->sumsq[0] / n_elems, Lf = s->sumsq[1] / n_elems;
  const float Lr32 = (float)Lr, Lf32 = (float)Lf;
  const float hinge = m_hinge * Lr32 - Lf32;
  const float gate = hinge > 0.f ? 1.f : 0.f;
  s->row_scale[0] = (float)((1.0 + gate * m_hinge) * 2.0 / n_elems);
  s->row_scale[1] = (float)(-gate * 2.0 / n_elems);
  s->loss_main = Lr32 + fmaxf(0.f, hinge);
}

// gloss pieces (GANMF.py:133-135): (1-a)*Lf + a*mean((Hr-Hf)^2).  w_fm = a, except on the ranks > 0 of an
// item-sharded group, where it is 0: there sumsq[0] holds this rank's item slice of the reconstruction sum, the
// loss is linear in it, and the per-rank values are summed by the caller (the feature-matching term, identical
// on every rank, must enter that sum once).
__global__ void gloss_kernel(StepScalars* s, float alpha, float w_fm, double n_elems, double m_elems) {
  const float Lf = (float)(s->sumsq[0] / n_elems);
  const float fm = (float)(s->fm / m_elems);
  s->loss_main = (1.f - alpha) * Lf + w_fm * fm;
}

// losses[slot] = loss_main + reg * l2 / 2     (l2 accumulated by the Adam kernel, pre-update).
// Item-sharded groups log per-rank PARTIAL losses (summed over ranks once per epoch): main_w / shard_w are 0
// where loss_main / the l2 of a replicated tensor is already counted by another rank.
__global__ void finalize_loss_kernel(const StepScalars* s, float reg, float* losses, int slot, float main_w = 1.f,
                                     float shard_w = 1.f) {
  losses[slot] = main_w * s->loss_main + (float)(reg * 0.5 * (s->l2 + shard_w * s->l2_shard));
}

// ----------------------------------------------------------------------------- activations
__device__ __forceinline__ float act_bwd_from_out(int act, float h) {
  switch (act) {
    case ACT_TANH: return 1.f - h * h;
    case ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case ACT_SIGMOID: return h * (1.f - h);
    default: return 1.f;
  }
}
// X = act(X) in place over [M, N]
__global__ void act_fwd_kernel(float* X, int M, int N, int ld, int act) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n < N) X[(size_t)m * ld + n] = act_fwd(act, X[(size_t)m * ld + n]);
}
// dZ = dH * act'(H)
__global__ void act_bwd_kernel(const float* dH, const float* H, float* dZ, int M, int N, int ld, int act) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n < N) dZ[(size_t)m * ld + n] = dH[(size_t)m * ld + n] * act_bwd_from_out(act, H[(size_t)m * ld + n]);
}

// Z = a*X + b*Y over [M, N] (same ld)
__global__ void axpby_kernel(const float* X, const float* Y, float* Z, int M, int N, int ld, float a, float b) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n < N) Z[(size_t)m * ld + n] = a * X[(size_t)m * ld + n] + b * Y[(size_t)m * ld + n];
}

// DisGANMF output losses and their gradients (DisGANMF.py:114-117).
//   out2 = [out_r (B) ; out_f (B)] logits.  mode 0 (D step): d_r = -sigmoid(-o)/Bg, d_f = sigmoid(o)/Bg
//   mode 1 (G step): only the fake half gets d_f = sigmoid(o)/Bg (real half set to 0).
//   Bg = rows of the whole (all-GPU) minibatch: the losses are means over it.
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__global__ void bce_kernel(const float* __restrict__ out2, float* __restrict__ dout2, int B, int mode,
                           StepScalars* s, float Bg) {
  float lr = 0.f, lf = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * B; i += gridDim.x * blockDim.x) {
    const float o = out2[i];
    if (i < B) {
      lr += softplusf(-o);
      dout2[i] = mode == 0 ? -(1.f / (1.f + expf(o))) / Bg : 0.f;
    } else {
      lf += softplusf(o);
      dout2[i] = (1.f / (1.f + expf(-o))) / Bg;
    }
  }
  lr = warp_sum(lr);
  lf = warp_sum(lf);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s->bce[0], (double)lr);
    atomicAdd(&s->bce[1], (double)lf);
  }
}
__global__ void dis_loss_kernel(StepScalars* s, int mode, float alpha, double B, double m_elems) {
  const float lr = (float)(s->bce[0] / B), lf = (float)(s->bce[1] / B);
  s->loss_main = mode == 0 ? lr + lf : lf + alpha * (float)(s->fm / m_elems);
}

// ----------------------------------------------------------------------------- SIMT GEMM
// Exact-fp32 (FMA) GEMM with the same epilogue as the tensor-core kernel, for shapes where a
// 128-wide MMA tile would be almost empty (DisGANMF layers with 1..64 units, out layer N=1).
// A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                 long long sbn, long long sbk, int M, int N, int K, Epilogue ep) {
  __shared__ float sA[SG_K][SG_T + 1], sB[SG_K][SG_T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
      // index so that the contiguous global dimension is the fast thread index
      int mm, kk;
      if (sak == 1) { kk = i % SG_K; mm = i / SG_K; } else { mm = i % SG_T; kk = i / SG_T; }
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
      int nn, kb;
      if (sbk == 1) { kb = i % SG_K; nn = i / SG_K; } else { nn = i % SG_T; kb = i / SG_T; }
      const int n = n0 + nn, k2 = k0 + kb;
      sB[kb][nn] = (n < N && k2 < K) ? B[n * sbn + k2 * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float rs = ep.row_scale2 ? ep.row_scale2[m >= ep.row_split ? 1 : 0] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v2 = finish_element(ep, apply_epilogue(ep, acc[i][j], m, n, rs), m, n);
      if (ep.adam_m || m < ep.row_split) sq0 += v2; else sq1 += v2;
    }
  }
  if (ep.sumsq2 || (ep.adam_m && ep.adam_l2)) {
    sq0 = warp_sum(sq0);
    sq1 = warp_sum(sq1);
    if ((threadIdx.x & 31) == 0) {
      if (ep.adam_m) {
        if (sq0 != 0.f) atomicAdd(ep.adam_l2, (double)sq0);
      } else {
        if (sq0 != 0.f) atomicAdd(ep.sumsq2, (double)sq0);
        if (sq1 != 0.f) atomicAdd(ep.sumsq2 + 1, (double)sq1);
      }
    }
  }
}

inline cudaError_t simt_gemm(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M,
                             int N, int K, const Epilogue& ep, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T);
  simt_gemm_kernel<<<grid, 256, 0, st>>>(A, a_mn ? 1 : lda, a_mn ? lda : 1, B, b_mn ? 1 : ldb,
                                         b_mn ? ldb : 1, M, N, K, ep);
  return cudaGetLastError();
}

}  // namespace ganmf
