// sfxm_thomas.cuh - batched block-Thomas kernels (one thread per ensemble member), templated on the source of the
// block rows: an assembled Jacobian in HBM (sfxm_solve.cu) or the cell's finite-difference block row evaluated on the
// fly (sfxm_fused.cuh), so that the Newton linear solve never writes or reads J.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"
#include "sfxm_cpasync.cuh"

namespace sfxm {

// The sweeps are chains of dependent iterations with one thread per member, so nothing overlaps a DRAM round trip unless
// the rows a later iteration needs are requested early: every iteration prefetches the rows kPrefetch<S> iterations
// ahead into L1 (profile r1h: without it each iteration costs ~1500 cycles, one exposed HBM latency).
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// (Round 2, second half: the addresses inside the sweeps were rewritten as base + i * row with 32-bit row lengths - one
// IMAD.WIDE.U32 per address, 1575 -> 1147 and 2017 -> 1476 static instructions for the fused s = 1 / s = 2 kernels.  Times:
// fused sweeps 0.42 + 0.65 ms either way, assembled-J sweeps (MemSrc) 0.944 -> 1.003 ms: the sweeps are bound by their
// dependent chains at 8-14 resident warps per SM, not by instruction count, and the extra live registers cost more than
// the integer work they save; removed.  The fused row source keeps its 32-bit row strides, sfxm_fused.cuh.)
inline bool thomas_rows_fit(const sfxm_problem_t &P) {
  return P.nx < ((int64_t)1 << 32) && P.batch * (int64_t)(SFXM_MAX_NV * (1 + SFXM_MAX_NV) * 8) < ((int64_t)1 << 32);
}
// The back substitution re-reads the C' / rhs' rows the forward sweep wrote; its iterations are short (s^2 FMAs), so the
// rows are fetched kRingDepth iterations ahead with cp.async into a per-thread ring in shared memory (thread-private
// columns: no barrier needed, cp.async.wait_group orders the thread's own copies).  One-warp CTAs only.
constexpr int kThomasThreads = 32;
template <int K>   // K doubles per row and thread
struct RingDepth {
  static constexpr int value = K <= 2 ? 8 : (K <= 12 ? 4 : 2);   // power of two; 32 threads x K x depth x 8 B <= 11 KB
};

template <int S>
struct PrefetchDist {
  static constexpr int value = S == 1 ? 6 : (S == 2 ? 4 : 2);   // rows in flight x row size stays inside L1
};

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

// Row source that reads an assembled banded Jacobian and residual from memory (sfxm_jacobian's layout).
template <int NV>
struct MemSrc {
  static constexpr int kMaxThreads = kThomasThreads, kMinCtas = 1;
  struct Args {
    const double *J, *f;
  };
  const double *J, *f;
  __device__ __forceinline__ MemSrc(const Args &a, const sfxm_problem_t &, int64_t) : J(a.J), f(a.f) {}
  __device__ __forceinline__ void prepare(const sfxm_problem_t &, int64_t, int64_t) {}
  // request the blocks and the right-hand side of block row i (a row the sweep reaches a few iterations later)
  template <int S>
  __device__ __forceinline__ void prefetch(const sfxm_problem_t &P, int64_t i, int o, int64_t m) const {
    const double *p = J + (i * 3) * (int64_t)(NV * NV) * P.batch + m;
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < S; ++c) prefetch_l1(p + (int64_t)((d * NV + o + r) * NV + (o + c)) * P.batch);
#pragma unroll
    for (int r = 0; r < S; ++r)
      prefetch_l1(f + (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * P.batch + m);
  }
  template <int S>
  __device__ __forceinline__ void block(const sfxm_problem_t &P, int64_t i, int d, int o, int64_t m, double (&blk)[S][S]) const {
    load_block<NV, S>(J, i, d, o, P.batch, m, blk);
  }
  __device__ __forceinline__ double rhs(const sfxm_problem_t &P, int64_t i, int o, int r, int64_t m) const {
    return __ldg(f + (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * P.batch + m);
  }
};

// The Newton update folded into the back substitution: the thread that computes a member's step also applies it,
// y += s, and accumulates the two sums of the convergence criterion (sum s^2, sum y_new^2) on the way - one thread owns
// one member for the whole sweep, so the per-member norms need no reduction at all and the separate update + norms pass
// over (y, s) (sfxm_update_norms: two launches, ~0.11 ms per segment of the bench ensemble) disappears.  Every read of y
// by the row source precedes the first write of a step (the forward sweep and the last-row set-up come first).
struct StepUpd {
  double *y;        // state to update in place (nullptr: the step is only stored)
  double *norms;    // norms[m] = sum s^2, norms[batch + m] = sum y_new^2 over the segment (nullptr with y == nullptr)
};

struct StepSink {
  double *s_out, *y;
  double ss, yy;
  __device__ __forceinline__ StepSink(double *s, const StepUpd &u) : s_out(s), y(u.y), ss(0.0), yy(0.0) {}
  __device__ __forceinline__ void put(int64_t idx, double v) {     // y[idx] read here: one exposed load
    s_out[idx] = v;
    if (y) put_y(idx, v, y[idx]);
  }
  // yold = y[idx] fetched ahead of time (the back substitution is a chain of dependent iterations: a load issued inside
  // an iteration would add one HBM round trip to each of them - measured 16 ms instead of 5.5 ms per steady state)
  __device__ __forceinline__ void put(int64_t idx, double v, double yold) {
    s_out[idx] = v;
    if (y) put_y(idx, v, yold);
  }
  __device__ __forceinline__ void put_y(int64_t idx, double v, double yold) {
    const double yn = yold + v;
    y[idx] = yn;
    ss += v * v;
    yy += yn * yn;
  }
  __device__ __forceinline__ void finish(const StepUpd &u, int64_t B, int64_t m) const {
    if (u.y && u.norms) {
      u.norms[m] = ss;
      u.norms[B + m] = yy;
    }
  }
};

// The kernels below take their block rows from a row source `Src`: MemSrc (an assembled Jacobian in HBM) or the fused
// source of sfxm_fused.cuh, which evaluates the block row of a cell on the fly so that J never exists in memory.
// Protocol per block row i: src.prepare(P, i, m) once, then src.block<S>(P, i, d, o, m, blk) / src.rhs(P, i, o, r, m).
//
// Block Thomas.  CYCLIC: rows 0..n-2 form an acyclic system with 1 + S right-hand sides (the extra S columns
// carry the coupling to the last block unknown), the last block row is then solved for x_{n-1} and folded back.
//   workC[(i*NV*NV + (o+r)*NV + (o+c))*B + m] : C'_i           workR[((i*NV + (o+r))*(1+NV) + k)*B + m] : rhs'_i / X_i
template <class Src, int NV, int S, bool CYCLIC>
__global__ void __launch_bounds__(Src::kMaxThreads, Src::kMinCtas)
block_thomas_kernel(const __grid_constant__ sfxm_problem_t P, int o, const __grid_constant__ typename Src::Args sargs,
                    const double *__restrict__ shift,
                    const int32_t *__restrict__ active, double *__restrict__ s_out, double *__restrict__ workC,
                    double *__restrict__ workR, int32_t *__restrict__ status, const StepUpd upd) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch;
  if (m >= B) return;
  if (active && !active[m]) return;
  Src src(sargs, P, m);
  constexpr int NR = CYCLIC ? 1 + S : 1;
  constexpr int NC = S + NR;
  const double sh = shift ? shift[m] : 0.0;
  const int64_t n = P.nx;
  const int64_t ne = CYCLIC ? n - 1 : n;
  bool ok = true;
  StepSink sink(s_out, upd);
  double Cp[S][S], Rp[S][NR];   // C'_{i-1}, rhs'_{i-1}
  auto fidx = [&](int64_t i, int r) { return (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m; };
  for (int64_t i = 0; i < ne; ++i) {
    double A[S][S], Bd[S][S], R[S][NC];
    if (i + PrefetchDist<S>::value < n) src.template prefetch<S>(P, i + PrefetchDist<S>::value, o, m);
    src.prepare(P, i, m);
    src.template block<S>(P, i, 1, o, m, Bd);
#pragma unroll
    for (int r = 0; r < S; ++r) Bd[r][r] += sh;
    {
      double C[S][S];
      src.template block<S>(P, i, 2, o, m, C);
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int c = 0; c < S; ++c) R[r][c] = (CYCLIC && i == ne - 1) ? 0.0 : C[r][c];
        R[r][S] = -src.rhs(P, i, o, r, m);
        if constexpr (CYCLIC) {
#pragma unroll
          for (int c = 0; c < S; ++c) R[r][S + 1 + c] = (i == ne - 1) ? C[r][c] : 0.0;
        }
      }
    }
    src.template block<S>(P, i, 0, o, m, A);
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
    for (int r = 0; r < S; ++r) sink.put(fidx(ne - 1, r), X[r][0]);
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
  constexpr int KB = S * S + S * NR;
  constexpr int DB = RingDepth<KB>::value;
  __shared__ double ring[DB][KB + S][kThomasThreads];   // + S: the state values the step is added to (StepUpd)
  const int tx = threadIdx.x;
  const bool upd_y = !CYCLIC && upd.y != nullptr;       // (cyclic: the steps are only known after the last row)
  auto issue = [&](int64_t row) {     // one commit group per call, empty past the first row: uniform group count
    if (row >= 0) {
      const int slot = (int)(row & (DB - 1));
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int c = 0; c < S; ++c)
          cp_async8(&ring[slot][r * S + c][tx], workC + (row * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m);
#pragma unroll
        for (int c = 0; c < NR; ++c)
          cp_async8(&ring[slot][S * S + r * NR + c][tx], workR + ((row * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m);
        if (upd_y) cp_async8(&ring[slot][KB + r][tx], upd.y + fidx(row, r));
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < DB; ++k) issue(ne - 2 - k);
  for (int64_t i = ne - 2; i >= 0; --i) {
    double C[S][S], Rn[S][NR], yold[S];
    cp_async_wait<DB - 1>();          // the oldest group in flight (row i) has landed
    {
      const int slot = (int)(i & (DB - 1));
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int c = 0; c < S; ++c) C[r][c] = ring[slot][r * S + c][tx];
#pragma unroll
        for (int c = 0; c < NR; ++c) Rn[r][c] = ring[slot][S * S + r * NR + c][tx];
        yold[r] = upd_y ? ring[slot][KB + r][tx] : 0.0;
      }
    }
    issue(i - DB);                    // refills the slot just read
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
      for (int r = 0; r < S; ++r) sink.put(fidx(i, r), X[r][0], yold[r]);
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
    src.prepare(P, n - 1, m);
    src.template block<S>(P, n - 1, 0, o, m, A);
    src.template block<S>(P, n - 1, 1, o, m, Bd);
    src.template block<S>(P, n - 1, 2, o, m, C);
#pragma unroll
    for (int r = 0; r < S; ++r) {
      Bd[r][r] += sh;
      rhs[r][0] = -src.rhs(P, n - 1, o, r, m);
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
    for (int r = 0; r < S; ++r) sink.put(fidx(n - 1, r), rhs[r][0]);
    for (int64_t i = 0; i < ne; ++i) {
      if (i + PrefetchDist<S>::value < ne) {
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
          for (int c = 0; c < NR; ++c)
            prefetch_l1(workR + (((i + PrefetchDist<S>::value) * NV + (o + r)) * (int64_t)(1 + NV) + c) * B + m);
        if (upd.y) {
#pragma unroll
          for (int r = 0; r < S; ++r) prefetch_l1(upd.y + fidx(i + PrefetchDist<S>::value, r));
        }
      }
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const double *w = workR + ((i * NV + (o + r)) * (int64_t)(1 + NV)) * B + m;
        double xv = w[0];
#pragma unroll
        for (int c = 0; c < S; ++c) xv -= w[(int64_t)(1 + c) * B] * rhs[c][0];
        sink.put(fidx(i, r), xv);
      }
    }
  }
  sink.finish(upd, B, m);
  if (status && !ok) status[m] = 1;
}

// Right-periodic only (lower-left corner block F = C_{n-1} at (n-1, 0), no upper-right one): plain Thomas on rows
// 0..n-2 while the fill that eliminating x_0, x_1, ... creates in the LAST row is carried in registers:
//   G_0 = F ;  (G_{n-2} += A_{n-1}) ;  d_last -= G_i rhs'_i ;  G_{i+1} = -G_i C'_i
// then (B_{n-1} + G_{n-1}) x_{n-1} = d_last and the usual back substitution.  One right-hand side, no extra
// workspace columns: ~40 % less traffic than the condensation the fully cyclic case needs.
template <class Src, int NV, int S>
__global__ void __launch_bounds__(Src::kMaxThreads, Src::kMinCtas)
block_thomas_ll_kernel(const __grid_constant__ sfxm_problem_t P, int o, const __grid_constant__ typename Src::Args sargs,
                       const double *__restrict__ shift,
                       const int32_t *__restrict__ active, double *__restrict__ s_out, double *__restrict__ workC,
                       double *__restrict__ workR, int32_t *__restrict__ status, const StepUpd upd) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch;
  if (m >= B) return;
  if (active && !active[m]) return;
  Src src(sargs, P, m);
  const double sh = shift ? shift[m] : 0.0;
  const int64_t n = P.nx;
  bool ok = true;
  StepSink sink(s_out, upd);
  auto fidx = [&](int64_t i, int r) { return (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m; };
  double G[S][S], Alast[S][S], Blast[S][S], dl[S];
  src.prepare(P, n - 1, m);
  src.template block<S>(P, n - 1, 0, o, m, Alast);
  src.template block<S>(P, n - 1, 1, o, m, Blast);
  src.template block<S>(P, n - 1, 2, o, m, G);            // F = wrap block of the last row
#pragma unroll
  for (int r = 0; r < S; ++r) {
    Blast[r][r] += sh;
    dl[r] = -src.rhs(P, n - 1, o, r, m);
  }
  double Cp[S][S], Rp[S];
  for (int64_t i = 0; i < n - 1; ++i) {
    double A[S][S], Bd[S][S], R[S][S + 1];
    if (i + PrefetchDist<S>::value < n) src.template prefetch<S>(P, i + PrefetchDist<S>::value, o, m);
    src.prepare(P, i, m);
    src.template block<S>(P, i, 1, o, m, Bd);
    {
      double C[S][S];
      src.template block<S>(P, i, 2, o, m, C);
#pragma unroll
      for (int r = 0; r < S; ++r) {
        Bd[r][r] += sh;
#pragma unroll
        for (int c = 0; c < S; ++c) R[r][c] = C[r][c];
        R[r][S] = -src.rhs(P, i, o, r, m);
      }
    }
    if (i > 0) {
      src.template block<S>(P, i, 0, o, m, A);
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
  for (int r = 0; r < S; ++r) sink.put(fidx(n - 1, r), X[r][0]);
  constexpr int KB = S * S + S;
  constexpr int DB = RingDepth<KB>::value;
  __shared__ double ring[DB][KB + S][kThomasThreads];   // + S: the state values the step is added to (StepUpd)
  const int tx = threadIdx.x;
  const bool upd_y = upd.y != nullptr;
  auto issue = [&](int64_t row) {
    if (row >= 0) {
      const int slot = (int)(row & (DB - 1));
#pragma unroll
      for (int r = 0; r < S; ++r) {
        if (upd_y) cp_async8(&ring[slot][KB + r][tx], upd.y + fidx(row, r));
        cp_async8(&ring[slot][S * S + r][tx], workR + ((row * NV + (o + r)) * (int64_t)(1 + NV)) * B + m);
#pragma unroll
        for (int c = 0; c < S; ++c)
          cp_async8(&ring[slot][r * S + c][tx], workC + (row * (int64_t)(NV * NV) + (o + r) * NV + (o + c)) * B + m);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < DB; ++k) issue(n - 2 - k);
  for (int64_t i = n - 2; i >= 0; --i) {
    double Xn[S], yold[S];
    cp_async_wait<DB - 1>();
    const int slot = (int)(i & (DB - 1));
#pragma unroll
    for (int r = 0; r < S; ++r) {
      double v = ring[slot][S * S + r][tx];
#pragma unroll
      for (int c = 0; c < S; ++c) v -= ring[slot][r * S + c][tx] * X[c][0];
      Xn[r] = v;
      yold[r] = upd_y ? ring[slot][KB + r][tx] : 0.0;
    }
    issue(i - DB);
#pragma unroll
    for (int r = 0; r < S; ++r) {
      X[r][0] = Xn[r];
      sink.put(fidx(i, r), Xn[r], yold[r]);
    }
  }
  sink.finish(upd, B, m);
  if (status && !ok) status[m] = 1;
}


}  // namespace sfxm
