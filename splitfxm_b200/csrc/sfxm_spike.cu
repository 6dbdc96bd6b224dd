// sfxm_spike.cu - the Newton linear solve of ONE mesh that is slab-decomposed over several GPUs.
//
// The reference solves the whole-mesh system in one process (splitnewton at simulation.py:515-523).  Here rank k owns
// the block rows of its slab: a block-tridiagonal T_k plus two coupling blocks, L_k = d r_first / d x_(last cell of the
// left neighbour) and R_k = d r_last / d x_(first cell of the right neighbour) - the wrap slots (cell 0, slot 0) and
// (cell nx-1, slot 2) of the slab Jacobian, which sfxm_jacobian fills for SFXM_BC_HALO ghosts.  Partitioned solve:
//
//   1. reduce (local, this file): PCR on T_k with 1 + 2 s right-hand sides  [ -f | e_first L_k | e_last R_k ]  gives
//          x_i = q_i - V_i xl_prev - W_i xf_next            for every local cell i
//      and the first/last rows of (q, V, W) are copied out as the slab's interface rows (2 s (1 + 2 s) doubles);
//   2. the interface rows of all ranks are all-gathered (a few hundred bytes per rank) and every rank solves the same
//      2 G s x 2 G s system for the interface unknowns (host side, splitfxm_b200/spike.py);
//   3. fold (local, this file): x_i = q_i - V_i xl_prev - W_i xf_next.
//
// No global sweep crosses a slab boundary, so the only data-path collective per solve is that one small all-gather.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"
#include "sfxm_pcr.cuh"

namespace sfxm {

template <int NV, int S>
__global__ void spike_init_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ J,
                                  const double *__restrict__ f, const double *__restrict__ shift,
                                  double *__restrict__ w, int32_t *__restrict__ status) {
  constexpr int NR = 1 + 2 * S;
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch, n = P.nx;
  if (m >= B || i >= n) return;
  const double sh = shift ? shift[m] : 0.0;
  double Bd[S][S];
  Row<S, NR> row;
  const double *p = J + (i * 3) * (int64_t)(NV * NV) * B + m;
#pragma unroll
  for (int r = 0; r < S; ++r) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int64_t e = (int64_t)((o + r) * NV + (o + c)) * B;
      const double a = __ldg(p + e), b = __ldg(p + (int64_t)(NV * NV) * B + e), cc = __ldg(p + (int64_t)(2 * NV * NV) * B + e);
      Bd[r][c] = b + (r == c ? sh : 0.0);
      row.R[r][c] = (i == 0) ? 0.0 : a;
      row.R[r][S + c] = (i == n - 1) ? 0.0 : cc;
      row.R[r][2 * S + 1 + c] = (i == 0) ? a : 0.0;             // L_k: coupling to the left neighbour's last cell
      row.R[r][2 * S + 1 + S + c] = (i == n - 1) ? cc : 0.0;    // R_k: coupling to the right neighbour's first cell
    }
    const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    row.R[r][2 * S] = -__ldg(f + q);
  }
  if (!gj_solve_pcr<S, 2 * S + NR>(Bd, row.R) && status) status[m] = 1;
  row.store(w, i, B, m);
}

// iface[((e * S + r) * NR + c) * B + m]: row r of (q | V | W) of the first (e = 0) and last (e = 1) local cell
template <int S>
__global__ void spike_iface_kernel(int64_t n, int64_t B, const double *__restrict__ w, double *__restrict__ iface) {
  constexpr int NR = 1 + 2 * S;
  constexpr int kRow = 2 * S * S + S * NR;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int64_t i = e ? n - 1 : 0;
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < NR; ++c)
        iface[((int64_t)((e * S + r) * NR + c)) * B + m] = w[((i * kRow) + r * (2 * S + NR) + 2 * S + c) * B + m];
  }
}

// nbr[(e * S + c) * B + m]: e = 0 the left neighbour's last block unknown, e = 1 the right neighbour's first
template <int S>
__global__ void spike_fold_kernel(const __grid_constant__ sfxm_problem_t P, int o, const double *__restrict__ w,
                                  const double *__restrict__ nbr, const int32_t *__restrict__ active,
                                  double *__restrict__ s_out) {
  constexpr int NR = 1 + 2 * S;
  constexpr int kRow = 2 * S * S + S * NR;
  int64_t m, i;
  pcr_coords(m, i);
  const int64_t B = P.batch;
  if (m >= B || i >= P.nx) return;
  if (active && !active[m]) return;
  double z[2 * S];
#pragma unroll
  for (int c = 0; c < 2 * S; ++c) z[c] = __ldg(nbr + (int64_t)c * B + m);
#pragma unroll
  for (int r = 0; r < S; ++r) {
    const double *d = w + ((i * kRow) + r * (2 * S + NR) + 2 * S) * B + m;
    double xv = d[0];
#pragma unroll
    for (int c = 0; c < 2 * S; ++c) xv -= d[(int64_t)(1 + c) * B] * z[c];
    const int64_t q = (P.var_base[o + r] + i * (int64_t)P.var_stride[o + r] + P.var_off[o + r]) * B + m;
    s_out[q] = xv;
  }
}

// Dense solve of the small interface system: one CTA per member, one thread per row, the augmented matrix
// [N][N + 1] in shared memory, Gaussian elimination with partial pivoting.  N = 2 G s <= 76 (8 GPUs x s = 4 -> 64).
__global__ void dense_solve_kernel(const double *__restrict__ aug, int N, double *__restrict__ z,
                                   int32_t *__restrict__ status) {
  extern __shared__ double A[];          // [N][N + 1]
  __shared__ int piv_s;
  const int64_t m = blockIdx.x;
  const int t = threadIdx.x, ld = N + 1;
  const double *src = aug + m * (int64_t)N * ld;
  for (int e = t; e < N * ld; e += blockDim.x) A[e] = src[e];
  __syncthreads();
  bool ok = true;
  for (int k = 0; k < N; ++k) {
    if (t == 0) {
      int piv = k;
      double best = fabs(A[k * ld + k]);
      for (int r = k + 1; r < N; ++r) {
        const double a = fabs(A[r * ld + k]);
        if (a > best) { best = a; piv = r; }
      }
      piv_s = piv;
      if (best == 0.0) ok = false;
    }
    __syncthreads();
    const int piv = piv_s;
    if (piv != k)
      for (int c = t; c < ld; c += blockDim.x) {
        const double a = A[k * ld + c];
        A[k * ld + c] = A[piv * ld + c];
        A[piv * ld + c] = a;
      }
    __syncthreads();
    if (t > k && t < N) {
      const double f = A[t * ld + k] / A[k * ld + k];
      for (int c = k; c < ld; ++c) A[t * ld + c] -= f * A[k * ld + c];
    }
    __syncthreads();
  }
  if (t == 0) {
    for (int r = N - 1; r >= 0; --r) {
      double v = A[r * ld + N];
      for (int c = r + 1; c < N; ++c) v -= A[r * ld + c] * A[c * ld + N];
      A[r * ld + N] = v / A[r * ld + r];
    }
    if (!ok && status) status[m] = 1;
  }
  __syncthreads();
  for (int r = t; r < N; r += blockDim.x) z[m * (int64_t)N + r] = A[r * ld + N];
}

// the buffer that holds the final rows after ceil(log2 n) ping-pong passes
static double *spike_final(double *work, int64_t n, int64_t row, int64_t B) {
  int passes = 0;
  for (int64_t h = 1; h < n; h *= 2) ++passes;
  return (passes & 1) ? work + n * row * B : work;
}

// Large slabs: nested dissection as in sfxm_pcr.cu, with the slab's first AND last cell among the separators.  The
// chunk reduction carries nothing extra: the coupling blocks to the neighbouring slabs sit in the wrap slots of cells 0
// and n - 1, which are separators, so they turn up as the wrap slots of the separator system - a system of the same
// form as the slab's own, np = n / 256 rows, that is handed back to sfxm_spike_reduce.  Its interface rows ARE the
// slab's interface rows; sfxm_spike_fold folds the separator values, then the chunk cells.
template <int NV, int S>
static int spike_reduce_chunked(const sfxm_problem_t &P, int o, int seg, const double *J, const double *f,
                                const double *shift, double *work, double *iface, int32_t *status, cudaStream_t st) {
  const ChunkLayout L = chunk_layout(P, S, true, work);
  if (int e = chunk_reduce_launch<NV, S>(P, L, o, J, f, shift, status, st)) return e;
  return sfxm_spike_reduce(&L.Pr, L.Jr, L.fr, nullptr, seg, L.work_r, iface, status, (void *)st);
}

template <int S>
static int spike_fold_chunked(const sfxm_problem_t &P, int o, int seg, double *work, const double *nbr,
                              const int32_t *active, double *s, cudaStream_t st) {
  const ChunkLayout L = chunk_layout(P, S, true, work);
  if (int e = sfxm_spike_fold(&L.Pr, seg, L.work_r, nbr, nullptr, L.xr, (void *)st)) return e;
  return chunk_fold_launch<S>(P, L, o, active, s, st);
}

template <int NV, int S>
static int spike_reduce_t(const sfxm_problem_t &P, int o, const double *J, const double *f, const double *shift,
                          double *work, double *iface, int32_t *status, cudaStream_t st) {
  constexpr int NR = 1 + 2 * S;
  constexpr int kRow = 2 * S * S + S * NR;
  const int64_t B = P.batch, n = P.nx;
  double *w0 = work, *w1 = work + n * kRow * B;
  dim3 b = pcr_block(B), g = pcr_grid(B, n, b);
  if (g.y > 65535) return set_error("sfxm_spike_reduce: batch too large");
  spike_init_kernel<NV, S><<<g, b, 0, st>>>(P, o, J, f, shift, w0, status);
  for (int64_t h = 1; h < n; h *= 2) {
    pcr_step_kernel<S, NR><<<g, b, 0, st>>>(n, B, h, w0, w1, status);
    double *t = w0; w0 = w1; w1 = t;
  }
  spike_iface_kernel<S><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(n, B, w0, iface);
  return check_launch("sfxm_spike_reduce");
}

template <int S>
static int spike_fold_t(const sfxm_problem_t &P, int o, double *work, const double *nbr, const int32_t *active,
                        double *s, cudaStream_t st) {
  constexpr int kRow = 2 * S * S + S * (1 + 2 * S);
  const int64_t B = P.batch;
  dim3 b = pcr_block(B), g = pcr_grid(B, P.nx, b);
  if (g.y > 65535) return set_error("sfxm_spike_fold: batch too large");
  spike_fold_kernel<S><<<g, b, 0, st>>>(P, o, spike_final(work, P.nx, kRow, B), nbr, active, s);
  return check_launch("sfxm_spike_fold");
}

static bool segment_of(const sfxm_problem_t &P, int seg, int &o, int &S) {
  o = -1;
  S = 0;
  for (int j = 0; j < P.nv; ++j)
    if (P.var_seg[j] == seg) {
      if (o < 0) o = j;
      ++S;
    }
  return o >= 0;
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int64_t sfxm_spike_workspace_bytes(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const int64_t s = prob->nv;
  const int64_t row = 2 * s * s + s * (1 + 2 * s);
  return 2 * prob->nx * row * prob->batch * (int64_t)sizeof(double);
}

extern "C" int sfxm_spike_reduce(const sfxm_problem_t *prob, const double *J, const double *f, const double *shift,
                                 int32_t seg, double *work, double *iface, int32_t *status, void *stream) {
  if (!prob || !J || !f || !work || !iface) return set_error("sfxm_spike_reduce: null argument");
  const sfxm_problem_t &P = *prob;
  if (P.nbl != 1 || P.nbr != 1) return set_error("sfxm_spike_reduce: needs a 3-cell stencil (nb 1/1)");
  if (P.nx < 3) return set_error("sfxm_spike_reduce: slab with fewer than 3 cells");
  int o, S;
  if (!segment_of(P, seg, o, S)) return set_error("sfxm_spike_reduce: empty segment");
  cudaStream_t st = (cudaStream_t)stream;
#define SFXM_SPIKE_CASE(NVV, SS)                                                                                    \
  if (P.nv == NVV && S == SS)                                                                                       \
    return chunked_applies(P) ? spike_reduce_chunked<NVV, SS>(P, o, seg, J, f, shift, work, iface, status, st)     \
                              : spike_reduce_t<NVV, SS>(P, o, J, f, shift, work, iface, status, st);
  SFXM_SPIKE_CASE(1, 1)
  SFXM_SPIKE_CASE(2, 1) SFXM_SPIKE_CASE(2, 2)
  SFXM_SPIKE_CASE(3, 1) SFXM_SPIKE_CASE(3, 2) SFXM_SPIKE_CASE(3, 3)
  SFXM_SPIKE_CASE(4, 1) SFXM_SPIKE_CASE(4, 2) SFXM_SPIKE_CASE(4, 3) SFXM_SPIKE_CASE(4, 4)
#undef SFXM_SPIKE_CASE
  return set_error("sfxm_spike_reduce: nv must be 1..4");
}

extern "C" int sfxm_spike_fold(const sfxm_problem_t *prob, int32_t seg, double *work, const double *nbr,
                               const int32_t *active, double *s, void *stream) {
  if (!prob || !work || !nbr || !s) return set_error("sfxm_spike_fold: null argument");
  const sfxm_problem_t &P = *prob;
  int o, S;
  if (!segment_of(P, seg, o, S)) return set_error("sfxm_spike_fold: empty segment");
  cudaStream_t st = (cudaStream_t)stream;
  const bool chunked = chunked_applies(P);
  switch (S) {
    case 1: return chunked ? spike_fold_chunked<1>(P, o, seg, work, nbr, active, s, st) : spike_fold_t<1>(P, o, work, nbr, active, s, st);
    case 2: return chunked ? spike_fold_chunked<2>(P, o, seg, work, nbr, active, s, st) : spike_fold_t<2>(P, o, work, nbr, active, s, st);
    case 3: return chunked ? spike_fold_chunked<3>(P, o, seg, work, nbr, active, s, st) : spike_fold_t<3>(P, o, work, nbr, active, s, st);
    case 4: return chunked ? spike_fold_chunked<4>(P, o, seg, work, nbr, active, s, st) : spike_fold_t<4>(P, o, work, nbr, active, s, st);
  }
  return set_error("sfxm_spike_fold: bad segment size");
}

extern "C" int sfxm_dense_solve(const double *aug, int32_t n, int64_t batch, double *z, int32_t *status, void *stream) {
  if (!aug || !z || n < 1 || batch < 1) return set_error("sfxm_dense_solve: bad argument");
  if (n > 76) return set_error("sfxm_dense_solve: system larger than 76 unknowns (2 x ranks x segment size)");
  const int threads = ((n + 31) / 32) * 32;
  dense_solve_kernel<<<(unsigned)batch, threads, (size_t)n * (n + 1) * sizeof(double), (cudaStream_t)stream>>>(aug, n, z, status);
  return check_launch("sfxm_dense_solve");
}
