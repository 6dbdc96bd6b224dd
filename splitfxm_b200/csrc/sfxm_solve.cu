// sfxm_solve.cu - batched block-tridiagonal (optionally cyclic) solves for Newton / Split-Newton, plus the
// small per-member vector kernels the Newton driver needs.  FP64, no tensor cores: the blocks are nv x nv
// (nv <= 4), far too small for an MMA tile, and the sweep is bound by streaming J once through HBM.
//
// One thread owns one ensemble member; memory is member-fastest, so a warp's 32 members read 32
// consecutive doubles for every matrix entry of the sweep (fully coalesced 256-byte rows).
//
// The solver the reference calls (splitnewton.newton / split_newton at simulation.py:515-523) is a
// third-party package that is not vendored; the algorithm here is specified in oracle/newton.py.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"
#include "sfxm_reduce.cuh"

namespace sfxm {

// R <- A^{-1} R by Gauss-Jordan elimination with partial (row) pivoting; returns false on a zero pivot
template <int S, int NC>
__device__ __forceinline__ bool gj_solve(double (&A)[S][S], double (&R)[S][NC]) {
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

// J block (i, slot d) restricted to rows/cols [o, o+S)
template <int NV, int S>
__device__ __forceinline__ void load_block(const double *__restrict__ J, int64_t i, int d, int o, int64_t B, int64_t m,
                                           double (&blk)[S][S]) {
  const double *p = J + ((i * 3 + d) * (int64_t)(NV * NV)) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < S; ++c) blk[r][c] = __ldg(p + (int64_t)((o + r) * NV + (o + c)) * B);
}

// Block Thomas.  CYCLIC: rows 0..n-2 form an acyclic system with 1 + S right-hand sides (the extra S columns
// carry the coupling to the last block unknown), the last block row is then solved for x_{n-1} and folded back.
//   workC[(i*NV*NV + (o+r)*NV + (o+c))*B + m] : C'_i           workR[((i*NV + (o+r))*(1+NV) + k)*B + m] : rhs'_i / X_i
template <int NV, int S, bool CYCLIC>
__global__ void __launch_bounds__(128)
block_thomas_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                    const double *__restrict__ f, const double *__restrict__ shift,
                    const int32_t *__restrict__ active, double *__restrict__ s_out, double *__restrict__ workC,
                    double *__restrict__ workR, int32_t *__restrict__ status) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch;
  if (m >= B) return;
  if (active && !active[m]) return;
  constexpr int NR = CYCLIC ? 1 + S : 1;
  constexpr int NC = S + NR;
  const double sh = shift ? shift[m] : 0.0;
  const int64_t n = P.nx;
  const int64_t ne = CYCLIC ? n - 1 : n;
  bool ok = true;
  double Cp[S][S], Rp[S][NR];   // C'_{i-1}, rhs'_{i-1}
  auto fidx = [&](int64_t i, int r) { return (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m; };
  for (int64_t i = 0; i < ne; ++i) {
    double A[S][S], Bd[S][S], R[S][NC];
    load_block<NV, S>(J, i, 1, o, B, m, Bd);
#pragma unroll
    for (int r = 0; r < S; ++r) Bd[r][r] += sh;
    {
      double C[S][S];
      load_block<NV, S>(J, i, 2, o, B, m, C);
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int c = 0; c < S; ++c) R[r][c] = (CYCLIC && i == ne - 1) ? 0.0 : C[r][c];
        R[r][S] = -__ldg(f + fidx(i, r));
        if constexpr (CYCLIC) {
#pragma unroll
          for (int c = 0; c < S; ++c) R[r][S + 1 + c] = (i == ne - 1) ? C[r][c] : 0.0;
        }
      }
    }
    load_block<NV, S>(J, i, 0, o, B, m, A);
    if (i == 0) {
      if constexpr (CYCLIC) {
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
          for (int c = 0; c < S; ++c) R[r][S + 1 + c] += A[r][c];   // wrap block couples row 0 to x_{n-1}
      }
    } else {
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const double a = A[r][k];
#pragma unroll
          for (int c = 0; c < S; ++c) Bd[r][c] -= a * Cp[k][c];
#pragma unroll
          for (int c = 0; c < NR; ++c) R[r][S + c] -= a * Rp[k][c];
        }
      }
    }
    ok = gj_solve<S, NC>(Bd, R) && ok;
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        Cp[r][c] = R[r][c];
        workC[(i * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m] = R[r][c];
      }
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        Rp[r][c] = R[r][S + c];
        if (CYCLIC || i < ne - 1) workR[((i * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m] = R[r][S + c];
      }
    }
  }
  // back substitution: X_i = rhs'_i - C'_i X_{i+1}
  double X[S][NR];
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < NR; ++c) X[r][c] = Rp[r][c];
  if constexpr (!CYCLIC) {
#pragma unroll
    for (int r = 0; r < S; ++r) s_out[fidx(ne - 1, r)] = X[r][0];
  } else {
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < NR; ++c) workR[(((ne - 1) * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m] = X[r][c];
  }
  double Xfirst[S][NR];   // X_0 (needed by the last row in the cyclic case)
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < NR; ++c) Xfirst[r][c] = X[r][c];
  double Xlast[S][NR];    // X_{n-2}
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < NR; ++c) Xlast[r][c] = X[r][c];
  for (int64_t i = ne - 2; i >= 0; --i) {
    double C[S][S], Rn[S][NR];
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int c = 0; c < S; ++c) C[r][c] = workC[(i * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m];
#pragma unroll
      for (int c = 0; c < NR; ++c) Rn[r][c] = workR[((i * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m];
    }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double a = C[r][k];
#pragma unroll
        for (int c = 0; c < NR; ++c) Rn[r][c] -= a * X[k][c];
      }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < NR; ++c) X[r][c] = Rn[r][c];
    if constexpr (!CYCLIC) {
#pragma unroll
      for (int r = 0; r < S; ++r) s_out[fidx(i, r)] = X[r][0];
    } else {
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) workR[((i * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m] = X[r][c];
      if (i == 0) {
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
          for (int c = 0; c < NR; ++c) Xfirst[r][c] = X[r][c];
      }
    }
  }
  if constexpr (CYCLIC) {
    // last block row: C_{n-1} x_0 + A_{n-1} x_{n-2} + B_{n-1} x_{n-1} = -f_{n-1},  x_i = q_i - Q_i x_{n-1}
    if (ne == 1) {
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) Xfirst[r][c] = Xlast[r][c];
    }
    double A[S][S], Bd[S][S], C[S][S], rhs[S][1];
    load_block<NV, S>(J, n - 1, 0, o, B, m, A);
    load_block<NV, S>(J, n - 1, 1, o, B, m, Bd);
    load_block<NV, S>(J, n - 1, 2, o, B, m, C);
#pragma unroll
    for (int r = 0; r < S; ++r) {
      Bd[r][r] += sh;
      rhs[r][0] = -__ldg(f + fidx(n - 1, r));
    }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int k = 0; k < S; ++k) {
        rhs[r][0] -= C[r][k] * Xfirst[k][0] + A[r][k] * Xlast[k][0];
#pragma unroll
        for (int c = 0; c < S; ++c) Bd[r][c] -= C[r][k] * Xfirst[k][1 + c] + A[r][k] * Xlast[k][1 + c];
      }
    ok = gj_solve<S, 1>(Bd, rhs) && ok;
#pragma unroll
    for (int r = 0; r < S; ++r) s_out[fidx(n - 1, r)] = rhs[r][0];
    for (int64_t i = 0; i < ne; ++i) {
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const double *w = workR + ((i * NV + (o + r)) * (int64_t)(1 + NV)) * B + m;
        double xv = w[0];
#pragma unroll
        for (int c = 0; c < S; ++c) xv -= w[(int64_t)(1 + c) * B] * rhs[c][0];
        s_out[fidx(i, r)] = xv;
      }
    }
  }
  if (status && !ok) status[m] = 1;
}

// Right-periodic only (lower-left corner block F = C_{n-1} at (n-1, 0), no upper-right one): plain Thomas on rows
// 0..n-2 while the fill that eliminating x_0, x_1, ... creates in the LAST row is carried in registers:
//   G_0 = F ;  (G_{n-2} += A_{n-1}) ;  d_last -= G_i rhs'_i ;  G_{i+1} = -G_i C'_i
// then (B_{n-1} + G_{n-1}) x_{n-1} = d_last and the usual back substitution.  One right-hand side, no extra
// workspace columns: ~40 % less traffic than the condensation the fully cyclic case needs.
template <int NV, int S>
__global__ void __launch_bounds__(128)
block_thomas_ll_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                       const double *__restrict__ f, const double *__restrict__ shift,
                       const int32_t *__restrict__ active, double *__restrict__ s_out, double *__restrict__ workC,
                       double *__restrict__ workR, int32_t *__restrict__ status) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch;
  if (m >= B) return;
  if (active && !active[m]) return;
  const double sh = shift ? shift[m] : 0.0;
  const int64_t n = P.nx;
  bool ok = true;
  auto fidx = [&](int64_t i, int r) { return (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m; };
  double G[S][S], Alast[S][S], Blast[S][S], dl[S];
  load_block<NV, S>(J, n - 1, 0, o, B, m, Alast);
  load_block<NV, S>(J, n - 1, 1, o, B, m, Blast);
  load_block<NV, S>(J, n - 1, 2, o, B, m, G);             // F = wrap block of the last row
#pragma unroll
  for (int r = 0; r < S; ++r) {
    Blast[r][r] += sh;
    dl[r] = -__ldg(f + fidx(n - 1, r));
  }
  double Cp[S][S], Rp[S];
  for (int64_t i = 0; i < n - 1; ++i) {
    double A[S][S], Bd[S][S], R[S][S + 1];
    load_block<NV, S>(J, i, 1, o, B, m, Bd);
    {
      double C[S][S];
      load_block<NV, S>(J, i, 2, o, B, m, C);
#pragma unroll
      for (int r = 0; r < S; ++r) {
        Bd[r][r] += sh;
#pragma unroll
        for (int c = 0; c < S; ++c) R[r][c] = C[r][c];
        R[r][S] = -__ldg(f + fidx(i, r));
      }
    }
    if (i > 0) {
      load_block<NV, S>(J, i, 0, o, B, m, A);
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const double a = A[r][k];
#pragma unroll
          for (int c = 0; c < S; ++c) Bd[r][c] -= a * Cp[k][c];
          R[r][S] -= a * Rp[k];
        }
    }
    ok = gj_solve<S, S + 1>(Bd, R) && ok;
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        Cp[r][c] = R[r][c];
        workC[(i * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m] = R[r][c];
      }
      Rp[r] = R[r][S];
      workR[((i * NV + (o + r)) * (int64_t)(1 + NV)) * B + m] = R[r][S];
    }
    // carry the last row's fill
    if (i == n - 2) {
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < S; ++c) G[r][c] += Alast[r][c];
    }
    double Gn[S][S];
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int c = 0; c < S; ++c) Gn[r][c] = 0.0;
#pragma unroll
      for (int k = 0; k < S; ++k) {
        dl[r] -= G[r][k] * Rp[k];
#pragma unroll
        for (int c = 0; c < S; ++c) Gn[r][c] -= G[r][k] * Cp[k][c];
      }
    }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < S; ++c) G[r][c] = Gn[r][c];
  }
  double X[S][1];
#pragma unroll
  for (int r = 0; r < S; ++r) {
    X[r][0] = dl[r];
#pragma unroll
    for (int c = 0; c < S; ++c) Blast[r][c] += G[r][c];
  }
  ok = gj_solve<S, 1>(Blast, X) && ok;
#pragma unroll
  for (int r = 0; r < S; ++r) s_out[fidx(n - 1, r)] = X[r][0];
  for (int64_t i = n - 2; i >= 0; --i) {
    double Xn[S];
#pragma unroll
    for (int r = 0; r < S; ++r) {
      double v = workR[((i * NV + (o + r)) * (int64_t)(1 + NV)) * B + m];
#pragma unroll
      for (int c = 0; c < S; ++c) v -= workC[(i * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m] * X[c][0];
      Xn[r] = v;
    }
#pragma unroll
    for (int r = 0; r < S; ++r) {
      X[r][0] = Xn[r];
      s_out[fidx(i, r)] = Xn[r];
    }
  }
  if (status && !ok) status[m] = 1;
}

// Compact copy of the same-segment entries of a split Jacobian (the only ones that exist, simulation.py:373-392):
// dst[((i*W + d)*L + e) * B + m] = src[((i*W + d)*nv*nv + rc[e]) * B + m],  rc = the L live (r, c) pairs in
// row-major order.  Pure row gather (coalesced 8-byte per lane reads/writes): what the host transfer moves.
struct PackMap {
  int32_t rc[SFXM_MAX_NV * SFXM_MAX_NV];
  int32_t L;
};
__global__ void pack_jacobian_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t rows, int nvnv,
                                     const __grid_constant__ PackMap map, int64_t B) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;   // (row, e)
  if (m >= B || t >= rows * map.L) return;
  const int64_t row = t / map.L;
  const int e = (int)(t - row * map.L);
  dst[t * B + m] = __ldg(src + (row * nvnv + map.rc[e]) * B + m);
}

struct SquareTerm {   // the norms of the Newton convergence criterion
  const double *v;
  int64_t B;
  __device__ __forceinline__ double operator()(int64_t q, int64_t m) const {
    const double a = __ldg(v + q * B + m);
    return a * a;
  }
};

__global__ void axpy_masked_kernel(double *__restrict__ y, const double *__restrict__ s, const double *__restrict__ alpha,
                                   const int32_t *__restrict__ active, int64_t lo, int64_t hi, int64_t B) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t q = lo + (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (m >= B || q >= hi) return;
  if (active && !active[m]) return;
  const double a = alpha ? alpha[m] : 1.0;
  y[q * B + m] = y[q * B + m] + a * s[q * B + m];
}

template <int NV, int S>
static int launch_thomas(const sfxm_problem_t &P, int o, int cyclic, const double *J, const double *f,
                         const double *shift, const int32_t *active, double *s, double *work, int32_t *status,
                         cudaStream_t st) {
  double *workC = work;
  double *workR = work + P.nx * (int64_t)(NV * NV) * P.batch;
  const int threads = P.batch >= 128 ? 128 : 32;
  const unsigned blocks = (unsigned)((P.batch + threads - 1) / threads);
  if (cyclic == 1)   // lower-left corner only
    block_thomas_ll_kernel<NV, S><<<blocks, threads, 0, st>>>(P, o, J, f, shift, active, s, workC, workR, status);
  else if (cyclic)
    block_thomas_kernel<NV, S, true><<<blocks, threads, 0, st>>>(P, o, J, f, shift, active, s, workC, workR, status);
  else
    block_thomas_kernel<NV, S, false><<<blocks, threads, 0, st>>>(P, o, J, f, shift, active, s, workC, workR, status);
  return check_launch("sfxm_block_tridiag_solve");
}

template <int NV>
static int solve_nv(const sfxm_problem_t &P, int o, int S, int cyclic, const double *J, const double *f,
                    const double *shift, const int32_t *active, double *s, double *work, int32_t *status,
                    cudaStream_t st) {
  if (S > NV) return set_error("segment larger than nv");
  switch (S) {
    case 1: return launch_thomas<NV, 1>(P, o, cyclic, J, f, shift, active, s, work, status, st);
    case 2: if constexpr (NV >= 2) return launch_thomas<NV, 2>(P, o, cyclic, J, f, shift, active, s, work, status, st); break;
    case 3: if constexpr (NV >= 3) return launch_thomas<NV, 3>(P, o, cyclic, J, f, shift, active, s, work, status, st); break;
    case 4: if constexpr (NV >= 4) return launch_thomas<NV, 4>(P, o, cyclic, J, f, shift, active, s, work, status, st); break;
  }
  return set_error("bad segment size");
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int64_t sfxm_solve_workspace_bytes(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const int64_t nv = prob->nv;
  return prob->nx * prob->batch * (nv * nv + nv * (1 + nv)) * (int64_t)sizeof(double);
}

extern "C" int sfxm_block_tridiag_solve(const sfxm_problem_t *prob, const double *J, const double *f,
                                        const double *shift, const int32_t *active, int32_t seg_lo, int32_t seg_hi,
                                        double *s, double *work, int32_t *status, void *stream) {
  if (!prob || !J || !f || !s || !work) return set_error("sfxm_block_tridiag_solve: null argument");
  const sfxm_problem_t &P = *prob;
  if (P.nbl != 1 || P.nbr != 1) return set_error("sfxm_block_tridiag_solve: needs a 3-cell stencil (nb 1/1)");
  if (P.nx < 3) return set_error("sfxm_block_tridiag_solve: nx < 3");
  cudaStream_t st = (cudaStream_t)stream;
  // segments are runs of equal var_seg
  for (int seg = seg_lo; seg < seg_hi; ++seg) {
    int o = -1, S = 0;
    bool left_per = false, right_per = false;
    for (int j = 0; j < P.nv; ++j)
      if (P.var_seg[j] == seg) {
        if (o < 0) o = j;
        ++S;
        left_per = left_per || P.bc_ltype[j] == SFXM_BC_PERIODIC;     // wrap block in row 0 (upper-right corner)
        right_per = right_per || P.bc_rtype[j] == SFXM_BC_PERIODIC;   // wrap block in row n-1 (lower-left corner)
      }
    if (o < 0) continue;   // empty segment
    const int cyclic = left_per ? 2 : (right_per ? 1 : 0);            // 2: condensation, 1: last-row fill, 0: plain
    int e = 0;
    switch (P.nv) {
      case 1: e = solve_nv<1>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 2: e = solve_nv<2>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 3: e = solve_nv<3>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 4: e = solve_nv<4>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      default: return set_error("nv must be 1..4");
    }
    if (e) return e;
  }
  return 0;
}

extern "C" int sfxm_pack_jacobian(const sfxm_problem_t *prob, const double *J, double *J_packed, int32_t *n_live_host,
                                  void *stream) {
  if (!prob || !n_live_host) return set_error("sfxm_pack_jacobian: null argument");
  const sfxm_problem_t &P = *prob;
  PackMap map;
  map.L = 0;
  for (int r = 0; r < P.nv; ++r)
    for (int c = 0; c < P.nv; ++c)
      if (P.var_seg[r] == P.var_seg[c]) map.rc[map.L++] = r * P.nv + c;
  *n_live_host = map.L;
  if (!J || !J_packed) return 0;     // query only
  const int64_t rows = P.nx * (P.nbl + P.nbr + 1);
  dim3 b = P.batch >= 32 ? dim3(32, 8) : dim3(1, 256);
  dim3 g((unsigned)((rows * map.L + b.y - 1) / b.y), (unsigned)((P.batch + b.x - 1) / b.x));
  if (g.y > 65535) return set_error("sfxm_pack_jacobian: batch too large");
  pack_jacobian_kernel<<<g, b, 0, (cudaStream_t)stream>>>(J, J_packed, rows, P.nv * P.nv, map, P.batch);
  return check_launch("sfxm_pack_jacobian");
}

extern "C" int64_t sfxm_sumsq_workspace_bytes(int64_t n, int64_t batch) {
  return (reduce_tiles(n, batch) + 1) * batch * (int64_t)sizeof(double);
}

extern "C" int sfxm_sumsq(const double *v, int64_t lo, int64_t hi, int64_t batch, double *out, double *work,
                          void *stream) {
  if (!v || !out || !work || hi < lo) return set_error("sfxm_sumsq: bad argument");
  if (batch > 65535LL * 32) return set_error("sfxm_sumsq: batch too large");
  reduce_rows(SquareTerm{v, batch}, lo, hi, batch, out, work, (cudaStream_t)stream);
  return check_launch("sfxm_sumsq");
}

extern "C" int sfxm_axpy_masked(double *y, const double *s, const double *alpha, const int32_t *active, int64_t lo,
                                int64_t hi, int64_t batch, void *stream) {
  if (!y || !s || hi < lo) return set_error("sfxm_axpy_masked: bad argument");
  if (hi == lo) return 0;
  dim3 b = batch >= 32 ? dim3(32, 8) : dim3(1, 256);
  dim3 g((unsigned)((hi - lo + b.y - 1) / b.y), (unsigned)((batch + b.x - 1) / b.x));
  axpy_masked_kernel<<<g, b, 0, (cudaStream_t)stream>>>(y, s, alpha, active, lo, hi, batch);
  return check_launch("sfxm_axpy_masked");
}
