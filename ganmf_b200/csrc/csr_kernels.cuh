// CSR transposition on the device: the reference transposes the training matrix on the host for
// item mode (GANRec/GANMF.py:32-33, `URM_train.T.tocsr()`); here the matrix is uploaded once and turned on the GPU.
//   1. csr_col_count_kernel     counts per column (integer atomics: the result does not depend on their order)
//   2. exclusive_scan_kernel    column counts -> indptr of the transpose (one CTA, running carry)
//   3. csr_scatter_kernel       every entry claims a slot of its column (order inside a column arbitrary)
//   4. csr_sort_segments_kernel one CTA per column sorts its (row, value) pairs by row: bitonic network in shared
//                               memory up to TR_CAP entries, in global memory (padded scratch) beyond
// The output is canonical (indices of every row of the transpose ascending), i.e. bit-identical to scipy's.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace ganmf {

constexpr int TR_CAP = 4096;          // pairs sorted in shared memory (32 KB)
constexpr int TR_THREADS = 512;

__global__ void csr_col_count_kernel(const int* __restrict__ indices, long long nnz, int* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) atomicAdd(cnt + indices[i], 1);
}

// out[0] = 0, out[i + 1] = in[0] + ... + in[i]   (n up to a few million: one CTA, fixed order)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* __restrict__ in, int n, int* __restrict__ out) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) { carry_s = 0; out[0] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int v = i < n ? in[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = v + (w ? warp_tot[w - 1] : 0) + carry;
    if (i < n) out[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

// one warp per source row: entry (r, c, v) -> slot cursor[c]++ of column c
__global__ void csr_scatter_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                   const float* __restrict__ data, int n_rows, const int* __restrict__ indptr_t,
                                   int* __restrict__ cursor, int* __restrict__ indices_t, float* __restrict__ data_t) {
  const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (r >= n_rows) return;
  const int lane = threadIdx.x & 31;
  for (int p = indptr[r] + lane; p < indptr[r + 1]; p += 32) {
    const int c = indices[p];
    const int q = indptr_t[c] + atomicAdd(cursor + c, 1);
    indices_t[q] = r;
    if (data_t) data_t[q] = data ? data[p] : 1.0f;
  }
}

__device__ __forceinline__ void bitonic_pairs(int* keys, float* vals, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int x = i ^ j;
        if (x > i) {
          const int a = keys[i], b = keys[x];
          if ((a > b) == ((i & k) == 0)) {
            keys[i] = b; keys[x] = a;
            if (vals) { const float t = vals[i]; vals[i] = vals[x]; vals[x] = t; }
          }
        }
      }
      __syncthreads();
    }
}

// grid = columns of the source (rows of the transpose).  Segments longer than TR_CAP are left to the long kernel.
__global__ void __launch_bounds__(TR_THREADS)
csr_sort_segments_kernel(const int* __restrict__ indptr_t, int* __restrict__ indices_t, float* __restrict__ data_t) {
  __shared__ int keys[TR_CAP];
  __shared__ float vals[TR_CAP];
  const int s = indptr_t[blockIdx.x], len = indptr_t[blockIdx.x + 1] - s;
  if (len <= 1 || len > TR_CAP) return;
  int n2 = 2;
  while (n2 < len) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    keys[i] = i < len ? indices_t[s + i] : INT_MAX;
    vals[i] = (i < len && data_t) ? data_t[s + i] : 0.f;
  }
  __syncthreads();
  bitonic_pairs(keys, data_t ? vals : nullptr, n2);
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    indices_t[s + i] = keys[i];
    if (data_t) data_t[s + i] = vals[i];
  }
}

// one CTA per LONG segment (list on the device), sorted in a padded global scratch [n_long][cap2]
__global__ void __launch_bounds__(TR_THREADS)
csr_sort_long_segments_kernel(const int* __restrict__ seg_ids, const int* __restrict__ indptr_t,
                              int* __restrict__ indices_t, float* __restrict__ data_t, int* __restrict__ scratch_k,
                              float* __restrict__ scratch_v, int cap2) {
  const int c = seg_ids[blockIdx.x];
  const int s = indptr_t[c], len = indptr_t[c + 1] - s;
  int* keys = scratch_k + (size_t)blockIdx.x * cap2;
  float* vals = data_t ? scratch_v + (size_t)blockIdx.x * cap2 : nullptr;
  int n2 = 2;
  while (n2 < len) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    keys[i] = i < len ? indices_t[s + i] : INT_MAX;
    if (vals) vals[i] = i < len ? data_t[s + i] : 0.f;
  }
  __syncthreads();
  bitonic_pairs(keys, vals, n2);
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    indices_t[s + i] = keys[i];
    if (vals) data_t[s + i] = vals[i];
  }
}

}  // namespace ganmf
