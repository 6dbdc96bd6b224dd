// sfxm_reduce.cuh - deterministic per-member sums over the rows of a member-fastest array.
//
// out[m] = sum over rows q in [lo, hi) of term(q, m), where term() is a functor.  Two stages:
//   1. a CTA owns a tile of TM members x ROWS_PER_CTA rows; thread (mx, ry) adds its rows ry, ry + TR, ... in
//      increasing order, the TR partial sums of a member are then added in increasing ry by one thread
//      -> partial[tile][m];
//   2. the tile partials of a member are added in increasing tile order.
// The summation order depends only on (lo, hi, batch), never on scheduling: results are bit-reproducible (the Newton
// criterion and the RK45 step controller branch on these numbers).  Loads are coalesced for every batch size:
// TM = 32 (members contiguous) when batch >= 32, TM = 1 (rows contiguous) otherwise.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sfxm {

static constexpr int kReduceThreads = 256;
static constexpr int kReduceRowsPerThread = 8;

__host__ __device__ inline int64_t reduce_rows_per_cta(int64_t batch) {
  return (int64_t)(batch >= 32 ? kReduceThreads / 32 : kReduceThreads) * kReduceRowsPerThread;
}
inline int64_t reduce_tiles(int64_t n, int64_t batch) {
  const int64_t r = reduce_rows_per_cta(batch);
  return (n + r - 1) / r;
}

template <class Term>
__global__ void __launch_bounds__(kReduceThreads)
reduce_partial_kernel(Term term, int64_t lo, int64_t hi, int64_t B, double *__restrict__ partial) {
  __shared__ double sh[kReduceThreads];
  const int TM = blockDim.x, TR = blockDim.y;
  const int64_t m = (int64_t)blockIdx.y * TM + threadIdx.x;
  const int64_t q0 = lo + (int64_t)blockIdx.x * TR * kReduceRowsPerThread + threadIdx.y;
  double acc = 0.0;
  if (m < B) {
#pragma unroll
    for (int k = 0; k < kReduceRowsPerThread; ++k) {
      const int64_t q = q0 + (int64_t)k * TR;
      if (q < hi) acc += term(q, m);
    }
  }
  sh[threadIdx.y * TM + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && m < B) {
    double tot = sh[threadIdx.x];
    for (int r = 1; r < TR; ++r) tot += sh[r * TM + threadIdx.x];
    partial[(int64_t)blockIdx.x * B + m] = tot;
  }
}

static __global__ void reduce_final_kernel(const double *__restrict__ partial, int64_t ntiles, int64_t B,
                                    double *__restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
  double acc = 0.0;
  for (int64_t k = 0; k < ntiles; ++k) acc += partial[k * B + m];
  out[m] = acc;
}

// Final stage for few members and many tiles (one large mesh: 1465 tiles at nx = 1e6, which the one-thread-per-member
// loop above adds serially in ~0.1 ms): one CTA per member; thread t adds tiles t, t + 256, ... in increasing order,
// then the 256 partial sums are folded by a fixed binary tree.  The order depends on (ntiles) only: bit-reproducible.
static __global__ void __launch_bounds__(256)
reduce_final_tree_kernel(const double *__restrict__ partial, int64_t ntiles, int64_t B, double *__restrict__ out) {
  __shared__ double sh[256];
  const int64_t m = blockIdx.x;
  double acc = 0.0;
  for (int64_t k = threadIdx.x; k < ntiles; k += 256) acc += partial[k * B + m];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[m] = sh[0];
}

static inline void reduce_final(const double *partial, int64_t ntiles, int64_t B, double *out, cudaStream_t st) {
  if (B <= 64 && ntiles > 64) reduce_final_tree_kernel<<<(unsigned)B, 256, 0, st>>>(partial, ntiles, B, out);
  else reduce_final_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(partial, ntiles, B, out);
}

// two sums per member from one pass: term(q, m) returns a double2; partial[(tile * 2 + w) * B + m], out2[w * B + m].
// Same tiles and the same order of additions as reduce_partial_kernel, so each sum has the bits the one-sum kernel gives.
template <class Term>
__global__ void __launch_bounds__(kReduceThreads)
reduce_partial2_kernel(Term term, int64_t lo, int64_t hi, int64_t B, double *__restrict__ partial) {
  __shared__ double sh[2][kReduceThreads];
  const int TM = blockDim.x, TR = blockDim.y;
  const int64_t m = (int64_t)blockIdx.y * TM + threadIdx.x;
  const int64_t q0 = lo + (int64_t)blockIdx.x * TR * kReduceRowsPerThread + threadIdx.y;
  double a0 = 0.0, a1 = 0.0;
  if (m < B) {
#pragma unroll
    for (int k = 0; k < kReduceRowsPerThread; ++k) {
      const int64_t q = q0 + (int64_t)k * TR;
      if (q < hi) {
        const double2 t = term(q, m);
        a0 += t.x;
        a1 += t.y;
      }
    }
  }
  sh[0][threadIdx.y * TM + threadIdx.x] = a0;
  sh[1][threadIdx.y * TM + threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && m < B) {
    double t0 = sh[0][threadIdx.x], t1 = sh[1][threadIdx.x];
    for (int r = 1; r < TR; ++r) {
      t0 += sh[0][r * TM + threadIdx.x];
      t1 += sh[1][r * TM + threadIdx.x];
    }
    partial[((int64_t)blockIdx.x * 2) * B + m] = t0;
    partial[((int64_t)blockIdx.x * 2 + 1) * B + m] = t1;
  }
}

// work must hold 2 * reduce_tiles(hi - lo, batch) * batch doubles; out2 holds 2 * batch
template <class Term>
static void reduce_rows2(Term term, int64_t lo, int64_t hi, int64_t batch, double *out2, double *work, cudaStream_t st) {
  const int64_t ntiles = reduce_tiles(hi - lo, batch);
  if (ntiles > 0) {
    dim3 b = batch >= 32 ? dim3(32, kReduceThreads / 32) : dim3(1, kReduceThreads);
    dim3 g((unsigned)ntiles, (unsigned)((batch + b.x - 1) / b.x));
    reduce_partial2_kernel<Term><<<g, b, 0, st>>>(term, lo, hi, batch, work);
  }
  reduce_final(work, ntiles, 2 * batch, out2, st);
}

// work must hold reduce_tiles(hi - lo, batch) * batch doubles
template <class Term>
static void reduce_rows(Term term, int64_t lo, int64_t hi, int64_t batch, double *out, double *work, cudaStream_t st) {
  const int64_t ntiles = reduce_tiles(hi - lo, batch);
  if (ntiles > 0) {
    dim3 b = batch >= 32 ? dim3(32, kReduceThreads / 32) : dim3(1, kReduceThreads);
    dim3 g((unsigned)ntiles, (unsigned)((batch + b.x - 1) / b.x));
    reduce_partial_kernel<Term><<<g, b, 0, st>>>(term, lo, hi, batch, work);
  }
  reduce_final(work, ntiles, batch, out, st);
}

}  // namespace sfxm
