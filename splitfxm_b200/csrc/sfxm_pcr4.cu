// sfxm_pcr4.cu - Newton linear solves for 4-cell stencils (FD RIGHT_BIAS: block band (nbl, nbr) = (1,2) or (2,1)).
//
// Cells are folded in pairs: super-cell k = cells (2k, 2k+1).  A cell couples to cells i-nbl .. i+nbr, i.e. to
// super-cells k-1, k, k+1 only, so the matrix is block tridiagonal in 2s x 2s super-blocks and the PCR passes of
// sfxm_pcr.cuh apply unchanged.  Periodic variables add corner super-blocks (cyclic system): condensation on the
// last super-cell, as for 3-cell stencils.  An odd mesh gets a dummy trailing cell (identity row); that is only
// possible without periodic coupling (a wrap from the second-to-last super-cell would skip one).
#include "sfxm_pcr.cuh"

namespace sfxm {

template <int NV, int S>
struct SuperRows {   // the three super-blocks and right-hand side of super-cell k, assembled from the banded J (W = 4)
  static constexpr int S2 = 2 * S;
  double A[S2][S2], Bd[S2][S2], C[S2][S2], Wl[S2][S2], Wr[S2][S2], d[S2];
  // Wl: wrap coupling of super-cell 0 to super-cell ns-1 (upper-right corner); Wr: of super-cell ns-1 to super-cell 0
  __device__ void build(const sfxm_problem_t &P, int o, const double *__restrict__ J, const double *__restrict__ f,
                        double sh, int64_t k, int64_t m) {
    const int64_t B = P.batch, n = P.nx;
#pragma unroll
    for (int r = 0; r < S2; ++r) {
      d[r] = 0.0;
#pragma unroll
      for (int c = 0; c < S2; ++c) A[r][c] = Bd[r][c] = C[r][c] = Wl[r][c] = Wr[r][c] = 0.0;
    }
    for (int a = 0; a < 2; ++a) {
      const int64_t i = 2 * k + a;
      if (i >= n) {                                  // dummy cell of an odd mesh: x = 0
        for (int r = 0; r < S; ++r) Bd[a * S + r][a * S + r] = 1.0;
        continue;
      }
      for (int r = 0; r < S; ++r) {
        const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
        d[a * S + r] = -__ldg(f + q);
      }
      for (int dd = 0; dd < 4; ++dd) {
        int64_t t = i - P.nbl + dd;
        bool wrap = false;
        if (t < 0) { t += n; wrap = true; }
        else if (t >= n) { t -= n; wrap = true; }
        const int64_t kt = t >> 1;
        const int b = (int)(t & 1);
        const double *p = J + ((i * 4 + dd) * (int64_t)(NV * NV)) * B + m;
        for (int r = 0; r < S; ++r)
          for (int c = 0; c < S; ++c) {
            double v = __ldg(p + (int64_t)((o + r) * NV + (o + c)) * B);
            if (!wrap && t == i && r == c) v += sh;
            double(*dst)[S2];
            if (wrap) dst = (k == 0) ? Wl : Wr;
            else if (kt == k) dst = Bd;
            else if (kt == k - 1) dst = A;
            else dst = C;
            dst[a * S + r][b * S + c] += v;
          }
      }
    }
  }
};

template <int NV, int S, bool CYCLIC>
__global__ void pcr4_init_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                                 const double *__restrict__ f, const double *__restrict__ shift, int64_t ns,
                                 double *__restrict__ w, int32_t *__restrict__ status) {
  constexpr int S2 = 2 * S;
  constexpr int NR = CYCLIC ? 1 + S2 : 1;
  int64_t m, k;
  pcr_coords(m, k);
  const int64_t B = P.batch, ne = CYCLIC ? ns - 1 : ns;
  if (m >= B || k >= ne) return;
  SuperRows<NV, S> sr;
  sr.build(P, o, J, f, shift ? shift[m] : 0.0, k, m);
  Row<S2, NR> row;
#pragma unroll
  for (int r = 0; r < S2; ++r) {
#pragma unroll
    for (int c = 0; c < S2; ++c) {
      row.R[r][c] = (k == 0) ? 0.0 : sr.A[r][c];
      row.R[r][S2 + c] = (k == ne - 1) ? 0.0 : sr.C[r][c];
      if constexpr (CYCLIC) row.R[r][2 * S2 + 1 + c] = (k == 0 ? sr.Wl[r][c] : 0.0) + (k == ne - 1 ? sr.C[r][c] : 0.0);
    }
    row.R[r][2 * S2] = sr.d[r];
  }
  if (!gj_solve_pcr<S2, 2 * S2 + NR>(sr.Bd, row.R) && status) status[m] = 1;
  row.store(w, k, B, m);
}

// x of the (condensed) system -> s_out.  CYCLIC: x_k = q_k - Q_k x_last, x_last from pcr4_last_kernel.
template <int S, bool CYCLIC>
__global__ void pcr4_write_kernel(const __grid_constant__ sfxm_problem_t P, int o, int64_t ns, const double *__restrict__ w,
                                  const double *__restrict__ xlast, const int32_t *__restrict__ active,
                                  double *__restrict__ s_out) {
  constexpr int S2 = 2 * S;
  constexpr int NR = CYCLIC ? 1 + S2 : 1;
  constexpr int kRow = 2 * S2 * S2 + S2 * NR;
  int64_t m, k;
  pcr_coords(m, k);
  const int64_t B = P.batch, ne = CYCLIC ? ns - 1 : ns;
  if (m >= B || k >= ne) return;
  if (active && !active[m]) return;
  for (int a = 0; a < 2; ++a) {
    const int64_t i = 2 * k + a;
    if (i >= P.nx) continue;
    for (int r = 0; r < S; ++r) {
      const double *dd = w + ((k * kRow) + (a * S + r) * (2 * S2 + NR) + 2 * S2) * B + m;
      double xv = dd[0];
      if constexpr (CYCLIC) {
        for (int c = 0; c < S2; ++c) xv -= dd[(int64_t)(1 + c) * B] * xlast[c * B + m];
      }
      s_out[(P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m] = xv;
    }
  }
}

// last super-row of a cyclic system:  Wr x_0 + A x_{ns-2} + B x_{ns-1} = d
template <int NV, int S>
__global__ void pcr4_last_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                                 const double *__restrict__ f, const double *__restrict__ shift, int64_t ns,
                                 const double *__restrict__ w, const int32_t *__restrict__ active,
                                 double *__restrict__ s_out, double *__restrict__ xlast, int32_t *__restrict__ status) {
  constexpr int S2 = 2 * S;
  constexpr int NR = 1 + S2;
  constexpr int kRow = 2 * S2 * S2 + S2 * NR;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t B = P.batch, ne = ns - 1;
  if (m >= B) return;
  SuperRows<NV, S> sr;
  sr.build(P, o, J, f, shift ? shift[m] : 0.0, ns - 1, m);
  double rhs[S2][1];
  for (int r = 0; r < S2; ++r) rhs[r][0] = sr.d[r];
  for (int kk = 0; kk < S2; ++kk) {
    double X0[NR], Xl[NR];
    for (int c = 0; c < NR; ++c) {
      X0[c] = w[((0 * kRow) + kk * (2 * S2 + NR) + 2 * S2 + c) * B + m];
      Xl[c] = w[(((ne - 1) * kRow) + kk * (2 * S2 + NR) + 2 * S2 + c) * B + m];
    }
    for (int r = 0; r < S2; ++r) {
      rhs[r][0] -= sr.Wr[r][kk] * X0[0] + sr.A[r][kk] * Xl[0];
      for (int c = 0; c < S2; ++c) sr.Bd[r][c] -= sr.Wr[r][kk] * X0[1 + c] + sr.A[r][kk] * Xl[1 + c];
    }
  }
  if (!gj_solve_pcr<S2, 1>(sr.Bd, rhs) && status) status[m] = 1;
  for (int r = 0; r < S2; ++r) xlast[r * B + m] = rhs[r][0];
  if (active && !active[m]) return;
  for (int a = 0; a < 2; ++a) {
    const int64_t i = 2 * (ns - 1) + a;
    if (i >= P.nx) continue;
    for (int r = 0; r < S; ++r)
      s_out[(P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m] = rhs[a * S + r][0];
  }
}

template <int NV, int S, bool CYCLIC>
static int pcr4_solve_t(const sfxm_problem_t &P, int o, const double *J, const double *f, const double *shift,
                        const int32_t *active, double *s, double *work, int32_t *status, cudaStream_t st) {
  constexpr int S2 = 2 * S;
  constexpr int NR = CYCLIC ? 1 + S2 : 1;
  constexpr int kRow = 2 * S2 * S2 + S2 * NR;
  const int64_t B = P.batch, ns = (P.nx + 1) / 2, ne = CYCLIC ? ns - 1 : ns;
  double *w0 = work, *w1 = work + ne * kRow * B, *xlast = w1 + ne * kRow * B;
  dim3 b = pcr_block(B), g = pcr_grid(B, ne, b);
  if (g.y > 65535) return set_error("batch too large for the PCR solver");
  pcr4_init_kernel<NV, S, CYCLIC><<<g, b, 0, st>>>(P, o, J, f, shift, ns, w0, status);
  for (int64_t h = 1; h < ne; h *= 2) {
    pcr_step_kernel<S2, NR><<<g, b, 0, st>>>(ne, B, h, w0, w1, status);
    double *t = w0; w0 = w1; w1 = t;
  }
  if constexpr (CYCLIC)
    pcr4_last_kernel<NV, S><<<(unsigned)((B + 63) / 64), 64, 0, st>>>(P, o, J, f, shift, ns, w0, active, s, xlast, status);
  pcr4_write_kernel<S, CYCLIC><<<g, b, 0, st>>>(P, o, ns, w0, xlast, active, s);
  return check_launch("sfxm_block_band4_solve_pcr");
}

template <int NV>
static int pcr4_nv(const sfxm_problem_t &P, int o, int S, bool cyc, const double *J, const double *f, const double *shift,
                   const int32_t *active, double *s, double *work, int32_t *status, cudaStream_t st) {
#define SFXM_PCR4_CASE(SS)                                                                                          \
  if constexpr (NV >= SS) {                                                                                         \
    if (S == SS) return cyc ? pcr4_solve_t<NV, SS, true>(P, o, J, f, shift, active, s, work, status, st)           \
                            : pcr4_solve_t<NV, SS, false>(P, o, J, f, shift, active, s, work, status, st);         \
  }
  SFXM_PCR4_CASE(1)
  SFXM_PCR4_CASE(2)
  SFXM_PCR4_CASE(3)
  SFXM_PCR4_CASE(4)
#undef SFXM_PCR4_CASE
  return set_error("bad segment size");
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int64_t sfxm_band4_workspace_bytes(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const int64_t s2 = 2 * (int64_t)prob->nv;
  const int64_t row = 2 * s2 * s2 + s2 * (1 + s2);
  return (2 * ((prob->nx + 1) / 2) * row + s2) * prob->batch * (int64_t)sizeof(double);
}

extern "C" int sfxm_block_band4_solve_pcr(const sfxm_problem_t *prob, const double *J, const double *f,
                                          const double *shift, const int32_t *active, int32_t seg_lo, int32_t seg_hi,
                                          double *s, double *work, int32_t *status, void *stream) {
  if (!prob || !J || !f || !s || !work) return set_error("sfxm_block_band4_solve_pcr: null argument");
  const sfxm_problem_t &P = *prob;
  if (P.nbl + P.nbr != 3) return set_error("sfxm_block_band4_solve_pcr: needs a 4-cell stencil");
  if (P.nx < 8) return set_error("sfxm_block_band4_solve_pcr: nx < 8");
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
    if (cyclic && (P.nx & 1)) return set_error("4-cell stencil with periodic variables needs an even number of cells");
    int e = 0;
    switch (P.nv) {
      case 1: e = pcr4_nv<1>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 2: e = pcr4_nv<2>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 3: e = pcr4_nv<3>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      case 4: e = pcr4_nv<4>(P, o, S, cyclic, J, f, shift, active, s, work, status, st); break;
      default: return set_error("nv must be 1..4");
    }
    if (e) return e;
  }
  return 0;
}
