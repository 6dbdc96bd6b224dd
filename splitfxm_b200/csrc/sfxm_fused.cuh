// sfxm_fused.cuh - Newton linear solve for ensembles with the Jacobian assembled INSIDE the block-Thomas sweep.
//
// The unfused path writes the banded Jacobian to HBM (sfxm_jacobian: 72-216 B per cell-member) and the solver reads it
// back, on top of the solver's own C'/rhs' round trip.  The thread that sweeps a member's cells can evaluate the
// finite-difference block row of the cell it stands on (the same jacobian_cell arithmetic, simulation.py:320-354) into
// registers instead: J never exists in memory, and the independent perturbation evaluations fill the issue slots the
// dependent elimination chain leaves empty.  Same kernels (sfxm_thomas.cuh), same operation order: the solution is
// bit-identical to sfxm_jacobian + sfxm_block_tridiag_solve.
//
// Traffic per cell-member (nv = 3): read U 24 (once; the two neighbouring cells hit L1), write + read C' (s^2) and rhs'
// (s), write s:  s = 1: 24 + 40 = 64 B   s = 2: 24 + 112 = 136 B   s = 3: 24 + 216 = 240 B
// against 104 + 24 / 336 + 24 / 696 + 24 B for assembly + solve.
#pragma once
#include "sfxm_kernels.cuh"
#include "sfxm_thomas.cuh"

namespace sfxm {

struct FusedArgs {
  const double *x_int, *params;
  int64_t pstride;
  const double *y;
  double eps;
  const double *geom_tab;
  uint32_t mask;      // entry mask of the segment being solved (bit r * NV + c)
  double *r0_out;     // optional: the segment's rows of the residual f(y), flat layout (what Newton's norms need)
};

// general path of one block row (cells whose stencil holds ghosts, inputs outside the fast-division range): out of
// line, result through memory - keeps the hot path's block row in registers
template <class Model, int SCHEME, int NV>
__device__ __noinline__ void fused_row_general(const sfxm_problem_t &P, const FusedArgs &a, int64_t i, int64_t m,
                                               double *__restrict__ Jl, double *__restrict__ r0l) {
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(a.params + (a.pstride ? k * a.pstride + m : k));
  View v{a.y, a.x_int, m, -1, 0, 0.0};
  double u[3][NV], xs[3], r0[NV];
  int64_t off[NV];
  const bool interior = gather_stencil<NV, 3>(P, v, i, a.geom_tab == nullptr, u, xs, off);
  Geom<SCHEME, 3> geom = a.geom_tab ? Geom<SCHEME, 3>(a.geom_tab, P.nx, i) : Geom<SCHEME, 3>(xs);
  geom.and_inputs(u, prm);
  geom.fast = geom.fast && sane(a.eps, 120);
  for (int e = 0; e < 3 * NV * NV; ++e) Jl[e] = 0.0;
  if (geom.fast)
    jacobian_cell<Model, SCHEME, NV, 3, true, true>(P, a.x_int, a.y, prm, i, m, interior, off, u, geom, a.eps, Jl, a.r0_out, a.mask, 1, r0);
  else
    jacobian_cell<Model, SCHEME, NV, 3, false, true>(P, a.x_int, a.y, prm, i, m, interior, off, u, geom, a.eps, Jl, a.r0_out, a.mask, 1, r0);
#pragma unroll
  for (int j = 0; j < NV; ++j) r0l[j] = r0[j];
}

// Resident one-warp CTAs per SM the register allocation must allow.  Measured on the 65,536 x nx 250 ensemble
// (gpurun_out/fused_time_variants.log): scalar segments run 0.64 -> 0.49 ms when capped at 128 registers (14 CTAs per SM:
// all 2048 warps resident in one wave, the spills are cheaper than a second wave); 2x2 and 3x3 segments spill too much
// under the cap and are left at 255 registers (8 CTAs per SM).  Re-measured in round 2 after the update was folded into the
// back substitution (gpurun_out r2x, 2x2 segment of the bench ensemble): uncapped 0.756 ms; 10 or 12 CTAs (168 registers,
// ~220 bytes spilled) 0.93 / 0.90 ms; 14 CTAs (128 registers) 1.05 ms.
#ifndef SFXM_FUSED_MIN_CTAS
#define SFXM_FUSED_MIN_CTAS(S) ((S) == 1 ? 14 : 1)
#endif

// select element `idx` (a run-time value) of a small register array without indexing it dynamically
template <int N>
__device__ __forceinline__ double pick(const double (&v)[N], int idx) {
  double out = v[0];
#pragma unroll
  for (int j = 1; j < N; ++j) out = (j == idx) ? v[j] : out;
  return out;
}

// Row source for segment size S.  The segment's block row lives in Jseg[d][r][c] / rseg[r] (compile-time indices only:
// registers); the segment offset o is a run-time value that only enters addresses and selects.
template <class Model, int SCHEME, int NV, int S>
struct FusedSrc {
  static constexpr int kMaxThreads = kThomasThreads, kMinCtas = SFXM_FUSED_MIN_CTAS(S);   // one-warp CTAs (fused_launch_s)
  using Args = FusedArgs;
  using M1 = typename Model::Scalar;
  const FusedArgs &a;
  const int o;
  double prm[Model::kParams];
  bool prm_ok;
  double reps;              // RN(1 / eps)
  int64_t seq_next;         // the cell the forward ring expects next (-1: ring not primed)
  double Jseg[3][S][S];
  double rseg[S];
  static constexpr int KV = Model::kDecoupled ? S : NV;   // state variables this source reads per cell
  int64_t ybase[KV];        // flat index of (cell 0, variable k, this member); + cell * ystride[k] (32-bit row length)
  uint32_t ystride[KV];

  __device__ __forceinline__ FusedSrc(const FusedArgs &args, const sfxm_problem_t &P, int64_t m) : a(args), o(first_var(P, args.mask)) {
    prm_ok = sane(a.eps, 120);
    reps = 1.0 / a.eps;
    seq_next = -1;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int j = Model::kDecoupled ? o + k : k;
      ybase[k] = (P.var_base[j] + P.var_off[j]) * P.batch + m;
      ystride[k] = (uint32_t)(P.var_stride[j] * P.batch);
    }
#pragma unroll
    for (int k = 0; k < Model::kParams; ++k) {
      prm[k] = __ldg(a.params + (a.pstride ? k * a.pstride + m : k));
      prm_ok = prm_ok & sane(prm[k], 120);
    }
  }
  __device__ __forceinline__ static int first_var(const sfxm_problem_t &, uint32_t mask) {
    int f = 0;
#pragma unroll
    for (int j = NV - 1; j >= 0; --j)
      if (mask & (1u << (j * NV + j))) f = j;
    return f;
  }

  template <int SS>
  __device__ __forceinline__ void prefetch(const sfxm_problem_t &, int64_t, int, int64_t) const {}   // see window()

  // The state of cells i-1, i, i+1 (interior cell i) for the variables this source needs: the segment's own for
  // decoupled models, all of them otherwise.  The sweep visits consecutive cells, so every cell is copied from HBM
  // ONCE, kDF - 2 cells ahead of its first use, by cp.async into a per-thread ring in shared memory (slot = cell mod
  // kDF); a jump in the cell sequence drains and re-primes the ring.
  static constexpr int kDF = 8;
  __device__ __forceinline__ int64_t cell_idx(int64_t cell, int k) const {
    return ybase[k] + (int64_t)((uint64_t)(uint32_t)cell * ystride[k]);
  }
  __device__ __forceinline__ const double *cell_ptr(const sfxm_problem_t &, int64_t cell, int k, int64_t) const {
    return a.y + cell_idx(cell, k);
  }
  __device__ __forceinline__ void window(const sfxm_problem_t &P, int64_t i, int64_t m, double (&w)[3][KV]) {
    __shared__ double ring[kDF][KV][kThomasThreads];
    const int tx = threadIdx.x;
    auto issue = [&](int64_t cell) {        // one commit group per cell (empty beyond the mesh)
      if (cell >= 0 && cell < P.nx) {
        const int slot = (int)(cell & (kDF - 1));
#pragma unroll
        for (int k = 0; k < KV; ++k) cp_async8(&ring[slot][k][tx], cell_ptr(P, cell, k, m));
      }
      cp_async_commit();
    };
    if (i != seq_next) {
      cp_async_wait<0>();                   // nothing of an earlier sequence may land after the re-prime
#pragma unroll
      for (int k = -1; k <= kDF - 2; ++k) issue(i + k);
    }
    cp_async_wait<kDF - 3>();               // cells <= i + 1 have landed
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
      const int slot = (int)((i - 1 + sl) & (kDF - 1));
#pragma unroll
      for (int k = 0; k < KV; ++k) w[sl][k] = ring[slot][k][tx];
    }
    issue(i + kDF - 1);                     // takes the slot of cell i - 1, read above
    seq_next = i + 1;
  }

  __device__ __forceinline__ void prepare(const sfxm_problem_t &P, int64_t i, int64_t m) {
    const int64_t B = P.batch;
    bool ok = prm_ok && a.geom_tab != nullptr && i >= 1 && i + 1 < P.nx;
    if (ok) {
      const Geom<SCHEME, 3> geom(a.geom_tab, P.nx, i);
      ok = geom.fast;
      if constexpr (Model::kDecoupled) {
        // one scalar equation per variable (jac_column's decoupled branch): only the segment's own variables enter
        double us[S][3][1];
        int64_t off[S];
        double w[3][KV];
        window(P, i, m, w);
#pragma unroll
        for (int cc = 0; cc < S; ++cc) {
          off[cc] = cell_idx(i, cc);
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) {
            us[cc][sl][0] = w[sl][cc];
            ok = ok & sane(us[cc][sl][0], 120);
          }
        }
        if (ok) {
#pragma unroll
          for (int cc = 0; cc < S; ++cc) {
            double r0[1];
            cell_residual<M1, SCHEME, 1, 3, true>(prm, us[cc], geom, r0);                 // simulation.py:320
            rseg[cc] = r0[0];
            if (a.r0_out) a.r0_out[off[cc]] = r0[0];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double up[3][1], r1[1];
#pragma unroll
              for (int sl = 0; sl < 3; ++sl) up[sl][0] = us[cc][sl][0];
              up[d][0] = us[cc][d][0] + a.eps;                                           // simulation.py:337
              cell_residual<M1, SCHEME, 1, 3, true>(prm, up, geom, r1);                  // simulation.py:345
#pragma unroll
              for (int rr = 0; rr < S; ++rr) Jseg[d][rr][cc] = 0.0;
              Jseg[d][cc][cc] = div_known<true>(r1[0] - r0[0], a.eps, reps);              // simulation.py:354
            }
          }
          return;
        }
      } else {
        double u[3][NV];
        int64_t off[NV];
        window(P, i, m, u);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          off[j] = cell_idx(i, j);
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) ok = ok & sane(u[sl][j], 120);
        }
        if (ok) {
          double r0[NV];
          cell_residual<Model, SCHEME, NV, 3, true>(prm, u, geom, r0);                    // simulation.py:320
#pragma unroll
          for (int rr = 0; rr < S; ++rr) {
            rseg[rr] = pick<NV>(r0, o + rr);
            if (a.r0_out) {
              int64_t q = off[0];
#pragma unroll
              for (int j = 1; j < NV; ++j) q = (j == o + rr) ? off[j] : q;
              a.r0_out[q] = rseg[rr];
            }
          }
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int cc = 0; cc < S; ++cc) {
              const int c = o + cc;
              double up[3][NV], r1[NV];
#pragma unroll
              for (int sl = 0; sl < 3; ++sl)
#pragma unroll
                for (int j = 0; j < NV; ++j) up[sl][j] = u[sl][j];
#pragma unroll
              for (int j = 0; j < NV; ++j) up[d][j] = (j == c) ? u[d][j] + a.eps : u[d][j];   // simulation.py:337
              cell_residual<Model, SCHEME, NV, 3, true>(prm, up, geom, r1);                   // simulation.py:345
#pragma unroll
              for (int rr = 0; rr < S; ++rr)
                Jseg[d][rr][cc] = div_known<true>(pick<NV>(r1, o + rr) - rseg[rr], a.eps, reps);   // simulation.py:354
            }
          return;
        }
      }
    }
    // general path: ghost cells in the stencil, or inputs outside the range of the fast division
    double Jt[3 * NV * NV], rt[NV];
    fused_row_general<Model, SCHEME, NV>(P, a, i, m, Jt, rt);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int rr = 0; rr < S; ++rr)
#pragma unroll
        for (int cc = 0; cc < S; ++cc) Jseg[d][rr][cc] = Jt[(d * NV + o + rr) * NV + o + cc];
#pragma unroll
    for (int rr = 0; rr < S; ++rr) rseg[rr] = rt[o + rr];
  }

  template <int SS>
  __device__ __forceinline__ void block(const sfxm_problem_t &, int64_t, int d, int, int64_t, double (&blk)[SS][SS]) const {
    static_assert(SS == S, "row source instantiated for another segment size");
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < S; ++c) blk[r][c] = d == 0 ? Jseg[0][r][c] : (d == 1 ? Jseg[1][r][c] : Jseg[2][r][c]);
  }
  __device__ __forceinline__ double rhs(const sfxm_problem_t &, int64_t, int, int r, int64_t) const {
    double v = rseg[0];
#pragma unroll
    for (int j = 1; j < S; ++j) v = (j == r) ? rseg[j] : v;
    return v;
  }
};

struct FusedLaunch {
  const sfxm_problem_t *P;
  FusedArgs args;
  int seg;
  const double *shift;
  const int32_t *active;
  double *s, *work;
  int32_t *status;
  cudaStream_t st;
  StepUpd upd;      // {y, norms}: apply the step and return the criterion's sums from the same sweep (nullptr: do not)
};

template <class Model, int SCHEME, int NV, int S>
static int fused_launch_s(const FusedLaunch &L, int o, int cyclic) {
  const sfxm_problem_t &P = *L.P;
  using Src = FusedSrc<Model, SCHEME, NV, S>;
  double *workC = L.work;
  double *workR = L.work + P.nx * (int64_t)(NV * NV) * P.batch;
  // 32-thread CTAs: 65,536 members are 2048 warps = 13.8 per SM; small CTAs spread them evenly over the 148 SMs
  const int threads = kThomasThreads;
  const unsigned blocks = (unsigned)((P.batch + threads - 1) / threads);
  if (cyclic == 1)
    block_thomas_ll_kernel<Src, NV, S><<<blocks, threads, 0, L.st>>>(P, o, L.args, L.shift, L.active, L.s, workC, workR, L.status, L.upd);
  else if (cyclic)
    block_thomas_kernel<Src, NV, S, true><<<blocks, threads, 0, L.st>>>(P, o, L.args, L.shift, L.active, L.s, workC, workR, L.status, L.upd);
  else
    block_thomas_kernel<Src, NV, S, false><<<blocks, threads, 0, L.st>>>(P, o, L.args, L.shift, L.active, L.s, workC, workR, L.status, L.upd);
  return check_launch("sfxm_newton_solve_fused");
}

template <class Model, int SCHEME, int NV>
static int fused_launch_scheme(const FusedLaunch &L0) {
  FusedLaunch L = L0;
  const sfxm_problem_t &P = *L.P;
  int o = -1, S = 0;
  bool left_per = false, right_per = false;
  uint32_t mask = 0;
  for (int j = 0; j < NV; ++j)
    if (P.var_seg[j] == L.seg) {
      if (o < 0) o = j;
      ++S;
      left_per = left_per || P.bc_ltype[j] == SFXM_BC_PERIODIC;
      right_per = right_per || P.bc_rtype[j] == SFXM_BC_PERIODIC;
    }
  if (o < 0) return set_error("sfxm_newton_solve_fused: empty segment");
  for (int r = o; r < o + S; ++r)
    for (int c = o; c < o + S; ++c) mask |= 1u << (r * NV + c);
  L.args.mask = mask;
  const int cyclic = left_per ? 2 : (right_per ? 1 : 0);
  switch (S) {
    case 1: return fused_launch_s<Model, SCHEME, NV, 1>(L, o, cyclic);
    case 2: if constexpr (NV >= 2) return fused_launch_s<Model, SCHEME, NV, 2>(L, o, cyclic); break;
    case 3: if constexpr (NV >= 3) return fused_launch_s<Model, SCHEME, NV, 3>(L, o, cyclic); break;
    case 4: if constexpr (NV >= 4) return fused_launch_s<Model, SCHEME, NV, 4>(L, o, cyclic); break;
  }
  return set_error("sfxm_newton_solve_fused: bad segment size");
}

}  // namespace sfxm
