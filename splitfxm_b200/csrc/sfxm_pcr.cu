// sfxm_pcr.cu - block-tridiagonal solves by parallel cyclic reduction (PCR), for one large mesh or a small batch.
//
// sfxm_block_tridiag_solve (sfxm_solve.cu) gives one thread per ensemble member: with batch = 1 that is a
// single thread walking the whole mesh.  PCR exposes one thread per (cell, member) instead: log2(n) passes, each
// eliminating the couplings to the rows +-h away (h = 1, 2, 4, ...) until every row stands alone.
//
// Rows are kept normalised (B_i = I):   Abar_i x_{i-h} + x_i + Cbar_i x_{i+h} = Dbar_i .   One pass:
//   A'_i = -Abar_i Abar_{i-h}          C'_i = -Cbar_i Cbar_{i+h}
//   B'_i = I - Abar_i Cbar_{i-h} - Cbar_i Abar_{i+h}
//   D'_i = Dbar_i - Abar_i Dbar_{i-h} - Cbar_i Dbar_{i+h}        then  [A' C' D'] <- B'^{-1} [A' C' D']
// (Gauss-Jordan with partial pivoting inside the s x s block, s <= 4).  Rows outside [0, n) do not exist: their
// couplings are zero.  After ceil(log2 n) passes x_i = Dbar_i.  Periodic variables make the system cyclic; as in
// the Thomas kernel it is condensed on the last block unknown: rows 0..n-2 are solved with 1 + s right-hand sides,
// then an s x s solve gives x_{n-1} and the rest is folded back.
// Traffic per pass and cell-member: 4 rows of (2 s^2 + s NR) doubles - HBM-bound streaming.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"
#include "sfxm_pcr.cuh"

namespace sfxm {

// J (banded, W = 3) + f -> normalised rows of the (condensed) system of ne = n or n-1 block rows
template <int NV, int S, bool CYCLIC>
__device__ __forceinline__ bool pcr_init_row(const sfxm_problem_t &P, int o, const double *__restrict__ J,
                                             const double *__restrict__ f, double sh, int64_t i, int64_t m,
                                             Row<S, CYCLIC ? 1 + S : 1> &row) {
  constexpr int NR = CYCLIC ? 1 + S : 1;
  const int64_t B = P.batch, ne = CYCLIC ? P.nx - 1 : P.nx;
  double Bd[S][S];
  const double *p = J + (i * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
      const double a = __ldg(p + e), b = __ldg(p + (int64_t)(NV * NV) * B + e), cc = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
      Bd[r][c] = b + (r == c ? sh : 0.0);
      row.R[r][c] = (i == 0) ? 0.0 : a;                    // slot 0 of row 0 is the wrap block, not a band entry
      row.R[r][S + c] = (i == ne - 1) ? 0.0 : cc;          // last row of the (condensed) system has no right neighbour
      if constexpr (CYCLIC) row.R[r][2 * S + 1 + c] = (i == 0 ? a : 0.0) + (i == ne - 1 ? cc : 0.0);   // coupling to x_{n-1}
    }
    const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    row.R[r][2 * S] = -__ldg(f + q);
  }
  return gj_solve_pcr<S, 2 * S + NR>(Bd, row.R);
}

template <int NV, int S, bool CYCLIC>
__global__ void pcr_init_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                                const double *__restrict__ f, const double *__restrict__ shift,
                                double *__restrict__ w, int32_t *__restrict__ status) {
  constexpr int NR = CYCLIC ? 1 + S : 1;
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch, ne = CYCLIC ? P.nx - 1 : P.nx;
  if (m >= B || i >= ne) return;
  Row<S, NR> row;
  if (!pcr_init_row<NV, S, CYCLIC>(P, o, J, f, shift ? shift[m] : 0.0, i, m, row) && status) status[m] = 1;
  row.store(w, i, B, m);
}

// ------------------------------------------------------------------------------------------------
// Small meshes: the whole reduction of one member in ONE launch, rows in shared memory.  The multi-launch path above
// costs 2 + log2(n) launches per segment, which is all a Newton iteration on a few hundred cells consists of (launch-
// bound: 0.13 ms per iteration at nx = 250).  One CTA per member, one thread per block row, the same row arithmetic
// (pcr_init_row / pcr_combine), a barrier between reading the neighbours' rows and overwriting one's own.
// ------------------------------------------------------------------------------------------------
template <int NV, int S, bool CYCLIC>
__global__ void __launch_bounds__(S <= 2 ? 1024 : 512)      // 64 / 128 registers: 3 x 3 rows spill badly at 64
pcr_smem_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                const double *__restrict__ f, const double *__restrict__ shift, const int32_t *__restrict__ active,
                double *__restrict__ s_out, int32_t *__restrict__ status) {
  constexpr int NR = CYCLIC ? 1 + S : 1;
  constexpr int kRow = 2 * S * S + S * NR;
  constexpr int kPitch = kRow | 1;
  extern __shared__ double sm[];
  const int64_t m = blockIdx.x, B = P.batch;
  const int n = (int)P.nx, ne = CYCLIC ? n - 1 : n;
  const int i = threadIdx.x;
  const bool mine = i < ne;
  const double sh = shift ? shift[m] : 0.0;
  bool ok = true;
  Row<S, NR> me;
  if (mine) {
    ok = pcr_init_row<NV, S, CYCLIC>(P, o, J, f, sh, i, m, me);
    me.store_sm(sm, i, kPitch);
  }
  __syncthreads();
  for (int h = 1; h < ne; h *= 2) {
    double out[S][2 * S + NR];
    if (mine) {
      Row<S, NR> lo, hi;
      const bool has_lo = i - h >= 0, has_hi = i + h < ne;
      if (has_lo) lo.load_sm(sm, i - h, kPitch);
      if (has_hi) hi.load_sm(sm, i + h, kPitch);
      ok = pcr_combine<S, NR>(me, has_lo, lo, has_hi, hi, out) && ok;
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int c = 0; c < 2 * S + NR; ++c) me.R[r][c] = out[r][c];
    }
    __syncthreads();                       // every thread has read its neighbours' rows of this pass
    if (mine) me.store_sm(sm, i, kPitch);
    __syncthreads();
  }
  const bool write = !active || active[m];
  auto q_of = [&](int64_t cell, int r) {
    return (P.var_base[o + r] + cell * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
  };
  if constexpr (!CYCLIC) {
    if (mine && write) {
#pragma unroll
      for (int r = 0; r < S; ++r) s_out[q_of(i, r)] = me.R[r][2 * S];
    }
  } else {
    // last block row: C_{n-1} x_0 + A_{n-1} x_{n-2} + B_{n-1} x_{n-1} = -f_{n-1},  x_i = q_i - Q_i x_{n-1}
    double *xl = sm + (size_t)ne * kPitch;      // x_{n-1}, S doubles
    if (i == 0) {
      Row<S, NR> r0, rl;
      r0.load_sm(sm, 0, kPitch);
      rl.load_sm(sm, ne - 1, kPitch);
      double A[S][S], Bd[S][S], C[S][S], rhs[S][1];
      const double *p = J + ((int64_t)(n - 1) * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
      for (int r = 0; r < S; ++r) {
#pragma unroll
        for (int c = 0; c < S; ++c) {
          const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
          A[r][c] = __ldg(p + e);
          Bd[r][c] = __ldg(p + (int64_t)(NV * NV) * B + e) + (r == c ? sh : 0.0);
          C[r][c] = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
        }
        rhs[r][0] = -__ldg(f + q_of(n - 1, r));
      }
#pragma unroll
      for (int r = 0; r < S; ++r)
#pragma unroll
        for (int k = 0; k < S; ++k) {
          rhs[r][0] -= C[r][k] * r0.R[k][2 * S] + A[r][k] * rl.R[k][2 * S];
#pragma unroll
          for (int c = 0; c < S; ++c) Bd[r][c] -= C[r][k] * r0.R[k][2 * S + 1 + c] + A[r][k] * rl.R[k][2 * S + 1 + c];
        }
      ok = gj_solve_pcr<S, 1>(Bd, rhs) && ok;
#pragma unroll
      for (int r = 0; r < S; ++r) {
        xl[r] = rhs[r][0];
        if (write) s_out[q_of(n - 1, r)] = rhs[r][0];
      }
    }
    __syncthreads();
    if (mine && write) {
#pragma unroll
      for (int r = 0; r < S; ++r) {
        double xv = me.R[r][2 * S];
#pragma unroll
        for (int c = 0; c < S; ++c) xv -= me.R[r][2 * S + 1 + c] * xl[c];
        s_out[q_of(i, r)] = xv;
      }
    }
  }
  if (!ok && status) status[m] = 1;
}

// non-cyclic: s = Dbar
template <int S>
__global__ void pcr_write_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ w,
                                 const int32_t *__restrict__ active, double *__restrict__ s_out) {
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch;
  if (m >= B || i >= P.nx) return;
  if (active && !active[m]) return;
  constexpr int kRow = 2 * S * S + S;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    s_out[q] = w[((i * kRow) + r * (2 * S + 1) + 2 * S) * B + m];
  }
}

// cyclic: last block row  C_{n-1} x_0 + A_{n-1} x_{n-2} + B_{n-1} x_{n-1} = -f_{n-1},  x_i = q_i - Q_i x_{n-1}
template <int NV, int S>
__global__ void pcr_cyclic_last_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                                       const double *__restrict__ f, const double *__restrict__ shift,
                                       const double *__restrict__ w, const int32_t *__restrict__ active,
                                       double *__restrict__ s_out, double *__restrict__ xlast, int32_t *__restrict__ status) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch, n = P.nx, ne = n - 1;
  if (m >= B) return;
  constexpr int NR = 1 + S;
  constexpr int kRow = 2 * S * S + S * NR;
  const double sh = shift ? shift[m] : 0.0;
  double A[S][S], Bd[S][S], C[S][S], rhs[S][1], X0[S][NR], Xl[S][NR];
  const double *p = J + ((n - 1) * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
      A[r][c] = __ldg(p + e);
      Bd[r][c] = __ldg(p + (int64_t)(NV * NV) * B + e) + (r == c ? sh : 0.0);
      C[r][c] = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      X0[r][c] = w[((0 * kRow) + r * (2 * S + NR) + 2 * S + c) * B + m];
      Xl[r][c] = w[(((ne - 1) * kRow) + r * (2 * S + NR) + 2 * S + c) * B + m];
    }
    const int64_t q = (P.var_base[o + r] + (n - 1) * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    rhs[r][0] = -__ldg(f + q);
  }
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int k = 0; k < S; ++k) {
      rhs[r][0] -= C[r][k] * X0[k][0] + A[r][k] * Xl[k][0];
#pragma unroll
      for (int c = 0; c < S; ++c) Bd[r][c] -= C[r][k] * X0[k][1 + c] + A[r][k] * Xl[k][1 + c];
    }
  if (!gj_solve_pcr<S, 1>(Bd, rhs) && status) status[m] = 1;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    xlast[r * B + m] = rhs[r][0];
    if (!active || active[m]) {
      const int64_t q = (P.var_base[o + r] + (n - 1) * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
      s_out[q] = rhs[r][0];
    }
  }
}

template <int S>
__global__ void pcr_cyclic_fold_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ w,
                                       const double *__restrict__ xlast, const int32_t *__restrict__ active,
                                       double *__restrict__ s_out) {
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch, ne = P.nx - 1;
  if (m >= B || i >= ne) return;
  if (active && !active[m]) return;
  constexpr int NR = 1 + S;
  constexpr int kRow = 2 * S * S + S * NR;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    const double *d = w + ((i * kRow) + r * (2 * S + NR) + 2 * S) * B + m;
    double xv = d[0];
#pragma unroll
    for (int c = 0; c < S; ++c) xv -= d[(int64_t)(1 + c) * B] * xlast[c * B + m];
    const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    s_out[q] = xv;
  }
}


template <int NV, int S>
static int chunked_solve_t(const sfxm_problem_t &P, int o, int seg, const double *J, const double *f, const double *shift,
                           const int32_t *active, double *s, double *work, int32_t *status, cudaStream_t st) {
  const ChunkLayout L = chunk_layout(P, S, false, work);
  if (int e = chunk_reduce_launch<NV, S>(P, L, o, J, f, shift, status, st)) return e;
  // the separator system has the form of the original one: solve it with this very solver (the shift is already in it)
  if (int e = sfxm_block_tridiag_solve_pcr(&L.Pr, L.Jr, L.fr, nullptr, nullptr, seg, seg + 1, L.xr, L.work_r, status, (void *)st))
    return e;
  return chunk_fold_launch<S>(P, L, o, active, s, st);
}

static bool smem_pcr_enabled() {     // SFXM_PCR_SMEM=0: always take the multi-launch path (tests compare the two)
  const char *e = getenv("SFXM_PCR_SMEM");
  return !(e && e[0] == '0');
}

template <int NV, int S, bool CYCLIC>
static int pcr_solve_t(const sfxm_problem_t &P, int o, const double *J, const double *f, const double *shift,
                       const int32_t *active, double *s, double *work, int32_t *status, cudaStream_t st) {
  constexpr int NR = CYCLIC ? 1 + S : 1;
  constexpr int kRow = 2 * S * S + S * NR;
  const int64_t B = P.batch, ne = CYCLIC ? P.nx - 1 : P.nx;
  // small meshes: one launch, rows in shared memory (one CTA per member)
  const size_t smem = ((size_t)ne * (kRow | 1) + S) * sizeof(double);
  if (ne <= (S <= 2 ? 1024 : 512) && smem <= 200 * 1024 && smem_pcr_enabled()) {
    auto kern = pcr_smem_kernel<NV, S, CYCLIC>;
    static bool configured = false;      // per instantiation
    if (!configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return set_error("sfxm_block_tridiag_solve_pcr: cannot raise the dynamic shared memory limit");
      configured = true;
    }
    const int threads = (int)((ne + 31) / 32 * 32);
    kern<<<(unsigned)B, threads, smem, st>>>(P, o, J, f, shift, active, s, status);
    return check_launch("sfxm_block_tridiag_solve_pcr");
  }
  double *w0 = work, *w1 = work + ne * kRow * B, *xlast = w1 + ne * kRow * B;
  dim3 b = pcr_block(B), g = pcr_grid(B, ne, b);
  if (g.y > 65535) return set_error("batch too large for the PCR solver (use the Thomas solver)");
  pcr_init_kernel<NV, S, CYCLIC><<<g, b, 0, st>>>(P, o, J, f, shift, w0, status);
  for (int64_t h = 1; h < ne; h *= 2) {
    pcr_step_kernel<S, NR><<<g, b, 0, st>>>(ne, B, h, w0, w1, status);
    double *t = w0; w0 = w1; w1 = t;
  }
  if constexpr (!CYCLIC) {
    pcr_write_kernel<S><<<pcr_grid(B, P.nx, b), b, 0, st>>>(P, o, w0, active, s);
  } else {
    pcr_cyclic_last_kernel<NV, S><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(P, o, J, f, shift, w0, active, s, xlast, status);
    pcr_cyclic_fold_kernel<S><<<g, b, 0, st>>>(P, o, w0, xlast, active, s);
  }
  return check_launch("sfxm_block_tridiag_solve_pcr");
}

template <int NV>
static int pcr_nv(const sfxm_problem_t &P, int o, int seg, int S, bool cyc, const double *J, const double *f,
                  const double *shift, const int32_t *active, double *s, double *work, int32_t *status, cudaStream_t st) {
  // large meshes: one level of nested dissection (the separator system comes back here with nx / 256 rows)
  if (chunked_applies(P)) {
#define SFXM_CHUNK_CASE(SS)                                                                                         \
  if constexpr (NV >= SS) {                                                                                         \
    if (S == SS) return chunked_solve_t<NV, SS>(P, o, seg, J, f, shift, active, s, work, status, st);               \
  }
    SFXM_CHUNK_CASE(1)
    SFXM_CHUNK_CASE(2)
    SFXM_CHUNK_CASE(3)
    SFXM_CHUNK_CASE(4)
#undef SFXM_CHUNK_CASE
  }
#define SFXM_PCR_CASE(SS)                                                                                           \
  if constexpr (NV >= SS) {                                                                                         \
    if (S == SS) return cyc ? pcr_solve_t<NV, SS, true>(P, o, J, f, shift, active, s, work, status, st)            \
                            : pcr_solve_t<NV, SS, false>(P, o, J, f, shift, active, s, work, status, st);          \
  }
  SFXM_PCR_CASE(1)
  SFXM_PCR_CASE(2)
  SFXM_PCR_CASE(3)
  SFXM_PCR_CASE(4)
#undef SFXM_PCR_CASE
  return set_error("bad segment size");
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int64_t sfxm_pcr_workspace_bytes(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const int64_t s = prob->nv;
  const int64_t row = 2 * s * s + s * (1 + s);
  return (2 * prob->nx * row + s) * prob->batch * (int64_t)sizeof(double);
}

extern "C" int sfxm_block_tridiag_solve_pcr(const sfxm_problem_t *prob, const double *J, const double *f,
                                            const double *shift, const int32_t *active, int32_t seg_lo,
                                            int32_t seg_hi, double *s, double *work, int32_t *status, void *stream) {
  if (!prob || !J || !f || !s || !work) return set_error("sfxm_block_tridiag_solve_pcr: null argument");
  const sfxm_problem_t &P = *prob;
  if (P.nbl != 1 || P.nbr != 1) return set_error("sfxm_block_tridiag_solve_pcr: needs a 3-cell stencil (nb 1/1)");
  if (P.nx < 3) return set_error("sfxm_block_tridiag_solve_pcr: nx < 3");
  cudaStream_t st = (cudaStream_t)stream;
  for (int seg = seg_lo; seg < seg_hi; ++seg) {
    int o = -1, S = 0;
    bool cyclic = false;
    for (int j = 0; j < P.nv; ++j)
      if (P.var_seg[j] == seg) {
        if (o < 0) o = j;
        ++S;
        if (P.bc_ltype[j] == SFXM_BC_PERIODIC || P.bc_rtype[j] == SFXM_BC_PERIODIC) cyclic = true;
      }
    if (o < 0) continue;
    int e = 0;
    switch (P.nv) {
      case 1: e = pcr_nv<1>(P, o, seg, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 2: e = pcr_nv<2>(P, o, seg, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 3: e = pcr_nv<3>(P, o, seg, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 4: e = pcr_nv<4>(P, o, seg, S, cyclic, J, f, shift, active, s, work, status, st); break;
      default: return set_error("nv must be 1..4");
    }
    if (e) return e;
  }
  return 0;
}
