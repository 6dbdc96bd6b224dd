// sfxm_pcr.cuh - shared pieces of the block parallel-cyclic-reduction solvers (sfxm_pcr.cu: 3-cell stencils,
// sfxm_pcr4.cu: 4-cell stencils folded into two-cell super-blocks).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------------
// Large meshes: one level of nested dissection around the shared-memory PCR.  Plain PCR streams 4 rows per cell
// through HBM ceil(log2 n) times (nx = 1e6: 20 passes, ~13 KB per cell for 264 algorithmic bytes).  Instead every
// kL-th cell is a SEPARATOR; the kL - 1 cells between two separators form a chunk that one CTA reduces entirely in
// shared memory, carrying 1 + 2 s right-hand sides [ -f | coupling to the left separator | to the right separator ]:
//       x_i = q_i - V_i x_sepL - W_i x_sepR                          (chunk_reduce_kernel: J read once, rows written once)
// Substituting the two cells next to a separator into its own block row gives a block-tridiagonal system of the
// SAME form in the separator unknowns only (separator_kernel), cyclic exactly when the original is - it is written in
// the banded layout of sfxm_jacobian and handed back to this solver (n / kL rows: shared memory or a few short
// passes).  chunk_fold_kernel then evaluates the chunk cells.  O(n) traffic: ~75 doubles per cell for s = 3.
// ------------------------------------------------------------------------------------------------
constexpr int kChunkPeriod = 256;   // separators at cells 0, 256, 512, ...; chunks of <= 255 cells, one thread each

// Separator k sits at cell k * kChunkPeriod; with force_last the last cell of the mesh is a separator too (the slab
// solve of sfxm_spike.cu needs both ends of the slab among the separators).  Chunk k = the cells between separators k
// and k + 1; the chunk after the last separator runs to the end of the mesh and its right separator is separator 0
// (through the wrap slot of cell n - 1).
struct SepMap {
  int64_t n, np;
  int force_last;
  __host__ __device__ int64_t pos(int64_t k) const { return (force_last && k == np - 1) ? n - 1 : k * kChunkPeriod; }
  __host__ __device__ bool is_sep(int64_t i) const { return (i % kChunkPeriod) == 0 || (force_last && i == n - 1); }
  __host__ __device__ int64_t chunk_end(int64_t k) const { return k + 1 < np ? pos(k + 1) : n; }
  __host__ __device__ int64_t sep_index(int64_t i) const { return (force_last && i == n - 1) ? np - 1 : i / kChunkPeriod; }
};
static inline SepMap make_sepmap(int64_t n, bool want_last) {
  const int64_t nreg = (n + kChunkPeriod - 1) / kChunkPeriod;
  const bool forced = want_last && ((n - 1) % kChunkPeriod != 0);
  return SepMap{n, nreg + (forced ? 1 : 0), forced ? 1 : 0};
}

// rows[((cell * S + r) * (1 + 2 S) + c) * B + m]: (q | V | W) of a chunk cell
// Register cap per block size (measured at nx = 1e6, gpurun_out r2v): the 2 x 2 kernel needs 150 registers uncapped
// (one 256-thread CTA per SM); capped at 128 (two CTAs, 74 bytes spilled) a split [1] Newton iteration takes 0.557
// instead of 0.596 ms.  The 3 x 3 kernel spills 1.1 KB under the same cap and doubles its time (1.09 -> 2.07 ms): uncapped.
#ifndef SFXM_CHUNK_MIN_CTAS
#define SFXM_CHUNK_MIN_CTAS(S) ((S) == 2 ? 2 : 1)
#endif
template <int NV, int S>
__global__ void __launch_bounds__(kChunkPeriod, SFXM_CHUNK_MIN_CTAS(S))
chunk_reduce_kernel(const __grid_constant__ sfxm_problem_t P, const SepMap sp, int o, const double *__restrict__ J,
                    const double *__restrict__ f, const double *__restrict__ shift, double *__restrict__ rows,
                    int32_t *__restrict__ status) {
  constexpr int NR = 1 + 2 * S;
  constexpr int NC = 2 * S + NR;
  constexpr int kPitch = (S * NC) | 1;
  extern __shared__ double sm[];
  const int64_t k = blockIdx.x, m = blockIdx.y, B = P.batch;
  const int64_t first = sp.pos(k) + 1;                                // first chunk cell
  const int nloc = (int)(sp.chunk_end(k) - first);
  if (nloc <= 0) return;
  const int i = threadIdx.x;
  const bool mine = i < nloc;
  const int64_t cell = first + i;
  const double sh = shift ? shift[m] : 0.0;
  bool ok = true;
  Row<S, NR> me;
  if (mine) {
    double Bd[S][S];
    const double *p = J + (cell * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
        const double a = __ldg(p + e), b = __ldg(p + (int64_t)(NV * NV) * B + e), cc = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
        Bd[r][c] = b + (r == c ? sh : 0.0);
        me.R[r][c] = (i == 0) ? 0.0 : a;
        me.R[r][S + c] = (i == nloc - 1) ? 0.0 : cc;
        me.R[r][2 * S + 1 + c] = (i == 0) ? a : 0.0;                  // coupling to the separator on the left
        me.R[r][2 * S + 1 + S + c] = (i == nloc - 1) ? cc : 0.0;      // ... on the right (cell n-1: the wrap block)
      }
      const int64_t q = (P.var_base[o + r] + cell * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
      me.R[r][2 * S] = -__ldg(f + q);
    }
    ok = gj_solve_pcr<S, NC>(Bd, me.R);
    me.store_sm(sm, i, kPitch);
  }
  __syncthreads();
  for (int h = 1; h < nloc; h *= 2) {
    double out[S][NC];
    if (mine) {
      Row<S, NR> lo, hi;
      const bool has_lo = i - h >= 0, has_hi = i + h < nloc;
      if (has_lo) lo.load_sm(sm, i - h, kPitch);
      if (has_hi) hi.load_sm(sm, i + h, kPitch);
      ok = pcr_combine<S, NR>(me, has_lo, lo, has_hi, hi, out) && ok;
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) me.R[r][c] = out[r][c];
    }
    __syncthreads();
    if (mine) me.store_sm(sm, i, kPitch);
    __syncthreads();
  }
  if (mine) {
    double *w = rows + (cell * S) * (int64_t)NR * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < NR; ++c) w[(int64_t)(r * NR + c) * B] = me.R[r][2 * S + c];
  }
  if (!ok && status) status[m] = 1;
}

// block row of separator k in the reduced system, in the banded layout [(k * 3 + d) * NV * NV + r * NV + c] * B + m,
// and fr = -(right-hand side) in the reduced flat layout (Pr)
template <int NV, int S>
__global__ void separator_kernel(const __grid_constant__ sfxm_problem_t P, const __grid_constant__ sfxm_problem_t Pr,
                                 const SepMap sp, int o,
                                 const double *__restrict__ J, const double *__restrict__ f,
                                 const double *__restrict__ shift, const double *__restrict__ rows,
                                 double *__restrict__ Jr, double *__restrict__ fr) {
  constexpr int NR = 1 + 2 * S;
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t k = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int64_t B = P.batch, n = P.nx, np = Pr.nx;
  if (m >= B || k >= np) return;
  const int64_t cell = sp.pos(k);
  const double sh = shift ? shift[m] : 0.0;
  double A[S][S], Bd[S][S], C[S][S], Ar[S][S], Cr[S][S], rhs[S];
  const double *p = J + (cell * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
      A[r][c] = __ldg(p + e);
      Bd[r][c] = __ldg(p + (int64_t)(NV * NV) * B + e) + (r == c ? sh : 0.0);
      C[r][c] = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
      Ar[r][c] = 0.0;
      Cr[r][c] = 0.0;
    }
    rhs[r] = -__ldg(f + (P.var_base[o + r] + cell * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m);
  }
  const int64_t left = (cell + n - 1) % n, right = (cell + 1) % n;
  if (sp.is_sep(left)) {                     // the neighbour is itself a separator
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < S; ++c) Ar[r][c] = A[r][c];
  } else {                                   // x_left = q - V x_sep(k-1) - W x_sep(k)
    const double *w = rows + (left * S) * (int64_t)NR * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int t = 0; t < S; ++t) {
        const double a = A[r][t];
        rhs[r] -= a * w[(int64_t)(t * NR) * B];
#pragma unroll
        for (int c = 0; c < S; ++c) {
          Ar[r][c] -= a * w[(int64_t)(t * NR + 1 + c) * B];
          Bd[r][c] -= a * w[(int64_t)(t * NR + 1 + S + c) * B];
        }
      }
  }
  if (sp.is_sep(right)) {
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < S; ++c) Cr[r][c] = C[r][c];
  } else {                                   // x_right = q - V x_sep(k) - W x_sep(k+1)
    const double *w = rows + (right * S) * (int64_t)NR * B + m;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int t = 0; t < S; ++t) {
        const double g = C[r][t];
        rhs[r] -= g * w[(int64_t)(t * NR) * B];
#pragma unroll
        for (int c = 0; c < S; ++c) {
          Bd[r][c] -= g * w[(int64_t)(t * NR + 1 + c) * B];
          Cr[r][c] -= g * w[(int64_t)(t * NR + 1 + S + c) * B];
        }
      }
  }
  double *q = Jr + (k * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
      q[e] = Ar[r][c];
      q[(int64_t)(NV * NV) * B + e] = Bd[r][c];
      q[(int64_t)(2 * NV * NV) * B + e] = Cr[r][c];
    }
    fr[(Pr.var_base[o + r] + k * (int64_t)Pr.var_stride[o + r] + Pr.var_off[o + r]) * B + m] = -rhs[r];
  }
}

template <int S>
__global__ void chunk_fold_kernel(const __grid_constant__ sfxm_problem_t P, const __grid_constant__ sfxm_problem_t Pr,
                                  const SepMap sp, int o,
                                  const double *__restrict__ rows, const double *__restrict__ xr,
                                  const int32_t *__restrict__ active, double *__restrict__ s_out) {
  constexpr int NR = 1 + 2 * S;
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch, np = Pr.nx;
  if (m >= B || i >= P.nx) return;
  if (active && !active[m]) return;
  const int64_t k = sp.sep_index(i);         // a chunk cell's left separator, or the separator's own index
  auto sep = [&](int64_t kk, int r) {
    return __ldg(xr + (Pr.var_base[o + r] + kk * (int64_t)Pr.var_stride[o + r] + Pr.var_off[o + r]) * B + m);
  };
  double xl[S], xh[S];
#pragma unroll
  for (int r = 0; r < S; ++r) {
    xl[r] = sep(k, r);
    xh[r] = sep((k + 1) % np, r);
  }
  const bool is_sep = sp.is_sep(i);
  const double *w = rows + (i * S) * (int64_t)NR * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    double xv = xl[r];
    if (!is_sep) {
      xv = w[(int64_t)(r * NR) * B];
#pragma unroll
      for (int c = 0; c < S; ++c) xv -= w[(int64_t)(r * NR + 1 + c) * B] * xl[c] + w[(int64_t)(r * NR + 1 + S + c) * B] * xh[c];
    }
    s_out[(P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m] = xv;
  }
}

static inline bool chunked_pcr_enabled() {    // SFXM_PCR_CHUNKED=0: plain multi-pass PCR for every size (tests compare the two)
  const char *e = getenv("SFXM_PCR_CHUNKED");
  return !(e && e[0] == '0');
}


// workspace layout shared by the solve (sfxm_pcr.cu) and the slab reduce / fold (sfxm_spike.cu)
struct ChunkLayout {
  sfxm_problem_t Pr;      // the separator system: same variables, segments and BC types, np "cells"
  SepMap sp;
  double *rows, *Jr, *fr, *xr, *work_r;
};
static inline ChunkLayout chunk_layout(const sfxm_problem_t &P, int S, bool want_last, double *work) {
  ChunkLayout L;
  L.sp = make_sepmap(P.nx, want_last);
  L.Pr = P;
  L.Pr.nx = L.sp.np;
  for (int j = 0; j < P.nv; ++j) L.Pr.var_base[j] = (P.var_base[j] / P.nx) * L.sp.np;     // var_base = (segment offset) * nx
  const int64_t B = P.batch, nv = P.nv;
  L.rows = work;
  L.Jr = L.rows + P.nx * (int64_t)(S * (1 + 2 * S)) * B;
  L.fr = L.Jr + L.sp.np * 3 * nv * nv * B;
  L.xr = L.fr + L.sp.np * nv * B;
  L.work_r = L.xr + L.sp.np * nv * B;
  return L;
}

// chunk reduction + separator rows: fills L.rows, L.Jr, L.fr
template <int NV, int S>
static int chunk_reduce_launch(const sfxm_problem_t &P, const ChunkLayout &L, int o, const double *J, const double *f,
                               const double *shift, int32_t *status, cudaStream_t st) {
  constexpr int NR = 1 + 2 * S;
  constexpr size_t smem = (size_t)(kChunkPeriod - 1) * ((S * (2 * S + NR)) | 1) * sizeof(double);
  const int64_t B = P.batch;
  auto kern = chunk_reduce_kernel<NV, S>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return set_error("chunked PCR: cannot raise the dynamic shared memory limit");
    configured = true;
  }
  if (B > 65535) return set_error("batch too large for the PCR solver (use the Thomas solver)");
  kern<<<dim3((unsigned)L.sp.np, (unsigned)B), kChunkPeriod, smem, st>>>(P, L.sp, o, J, f, shift, L.rows, status);
  dim3 b = pcr_block(B);
  separator_kernel<NV, S><<<pcr_grid(B, L.sp.np, b), b, 0, st>>>(P, L.Pr, L.sp, o, J, f, shift, L.rows, L.Jr, L.fr);
  return check_launch("chunked PCR reduce");
}

template <int S>
static int chunk_fold_launch(const sfxm_problem_t &P, const ChunkLayout &L, int o, const int32_t *active, double *s,
                             cudaStream_t st) {
  dim3 b = pcr_block(P.batch);
  chunk_fold_kernel<S><<<pcr_grid(P.batch, P.nx, b), b, 0, st>>>(P, L.Pr, L.sp, o, L.rows, L.xr, active, s);
  return check_launch("chunked PCR fold");
}

static inline bool chunked_applies(const sfxm_problem_t &P) { return P.nx >= 8 * kChunkPeriod && chunked_pcr_enabled(); }

}  // namespace sfxm
