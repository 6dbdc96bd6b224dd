// sfxm_pcr.cuh - shared pieces of the block parallel-cyclic-reduction solvers (sfxm_pcr.cu: 3-cell stencils,
// sfxm_pcr4.cu: 4-cell stencils folded into two-cell super-blocks).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {

template <int S, int NC>
__device__ __forceinline__ bool gj_solve_pcr(double (&A)[S][S], double (&R)[S][NC]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    int piv = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int r = k + 1; r < S; ++r) {
      const double a = fabs(A[r][k]);
      if (a > best) { best = a; piv = r; }
    }
    if (best == 0.0) ok = false;
#pragma unroll
    for (int r = k + 1; r < S; ++r) {
      if (r == piv) {
#pragma unroll
        for (int c = 0; c < S; ++c) { const double t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; }
#pragma unroll
        for (int c = 0; c < NC; ++c) { const double t = R[k][c]; R[k][c] = R[r][c]; R[r][c] = t; }
      }
    }
    const double inv = 1.0 / A[k][k];
#pragma unroll
    for (int c = 0; c < S; ++c) A[k][c] *= inv;
#pragma unroll
    for (int c = 0; c < NC; ++c) R[k][c] *= inv;
#pragma unroll
    for (int r = 0; r < S; ++r) {
      if (r == k) continue;
      const double f = A[r][k];
#pragma unroll
      for (int c = 0; c < S; ++c) A[r][c] -= f * A[k][c];
#pragma unroll
      for (int c = 0; c < NC; ++c) R[r][c] -= f * R[k][c];
    }
  }
  return ok;
}

// row storage: ROW = 2 S^2 + S NR doubles per (cell, member): [Abar | Cbar | Dbar], element e at (i*ROW + e)*B + m
template <int S, int NR>
struct Row {
  static constexpr int kRow = 2 * S * S + S * NR;
  double R[S][2 * S + NR];   // columns: Abar (S) | Cbar (S) | Dbar (NR)
  __device__ __forceinline__ void load(const double *__restrict__ w, int64_t i, int64_t B, int64_t m) {
    const double *p = w + (i * kRow) * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < 2 * S + NR; ++c) R[r][c] = __ldg(p + (int64_t)(r * (2 * S + NR) + c) * B);
  }
  // shared-memory copy of the rows of ONE member: row i at sm + i * pitch (pitch odd: conflict-free row-per-thread access)
  __device__ __forceinline__ void load_sm(const double *sm, int i, int pitch) {
    const double *p = sm + i * pitch;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < 2 * S + NR; ++c) R[r][c] = p[r * (2 * S + NR) + c];
  }
  __device__ __forceinline__ void store_sm(double *sm, int i, int pitch) const {
    double *p = sm + i * pitch;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < 2 * S + NR; ++c) p[r * (2 * S + NR) + c] = R[r][c];
  }
  __device__ __forceinline__ void store(double *__restrict__ w, int64_t i, int64_t B, int64_t m) const {
    double *p = w + (i * kRow) * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < 2 * S + NR; ++c) p[(int64_t)(r * (2 * S + NR) + c) * B] = R[r][c];
  }
};

__device__ __forceinline__ void pcr_coords(int64_t &m, int64_t &i) {
  m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
}

// one PCR pass for one row: [Abar | Cbar | Dbar] of row i from its own row and the rows h away (has_lo / has_hi: such a row exists)
template <int S, int NR>
__device__ __forceinline__ bool pcr_combine(const Row<S, NR> &me, bool has_lo, const Row<S, NR> &lo, bool has_hi,
                                            const Row<S, NR> &hi, double (&out)[S][2 * S + NR]) {
  double Bd[S][S];
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      Bd[r][c] = (r == c) ? 1.0 : 0.0;
      out[r][c] = 0.0;
      out[r][S + c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) out[r][2 * S + c] = me.R[r][2 * S + c];
  }
  if (has_lo) {
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double a = me.R[r][k];                        // Abar_i
#pragma unroll
        for (int c = 0; c < S; ++c) {
          out[r][c] -= a * lo.R[k][c];                     // A' = -Abar_i Abar_{i-h}
          Bd[r][c] -= a * lo.R[k][S + c];                  // B' -= Abar_i Cbar_{i-h}
        }
#pragma unroll
        for (int c = 0; c < NR; ++c) out[r][2 * S + c] -= a * lo.R[k][2 * S + c];
      }
  }
  if (has_hi) {
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double g = me.R[r][S + k];                    // Cbar_i
#pragma unroll
        for (int c = 0; c < S; ++c) {
          out[r][S + c] -= g * hi.R[k][S + c];             // C' = -Cbar_i Cbar_{i+h}
          Bd[r][c] -= g * hi.R[k][c];                      // B' -= Cbar_i Abar_{i+h}
        }
#pragma unroll
        for (int c = 0; c < NR; ++c) out[r][2 * S + c] -= g * hi.R[k][2 * S + c];
      }
  }
  return gj_solve_pcr<S, 2 * S + NR>(Bd, out);
}

template <int S, int NR>
__global__ void pcr_step_kernel(int64_t ne, int64_t B, int64_t h, const double *__restrict__ win,
                                double *__restrict__ wout, int32_t *__restrict__ status) {
  int64_t m, i;
  pcr_coords(m, i);
  if (m >= B || i >= ne) return;
  Row<S, NR> me, lo, hi;
  me.load(win, i, B, m);
  const bool has_lo = i - h >= 0, has_hi = i + h < ne;
  if (has_lo) lo.load(win, i - h, B, m);
  if (has_hi) hi.load(win, i + h, B, m);
  double out[S][2 * S + NR];
  if (!pcr_combine<S, NR>(me, has_lo, lo, has_hi, hi, out) && status) status[m] = 1;
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < 2 * S + NR; ++c) me.R[r][c] = out[r][c];
  me.store(wout, i, B, m);
}

static dim3 pcr_block(int64_t B) { return B >= 32 ? dim3(32, 4) : dim3(1, 128); }
static dim3 pcr_grid(int64_t B, int64_t n, dim3 b) {
  return dim3((unsigned)((n + b.y - 1) / b.y), (unsigned)((B + b.x - 1) / b.x));
}

}  // namespace sfxm
