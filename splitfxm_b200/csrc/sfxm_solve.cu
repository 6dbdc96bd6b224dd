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
#include "sfxm_thomas.cuh"

namespace sfxm {

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

// Newton update and the two norms of its convergence criterion in one pass over (y, s):
//   y += alpha s for the active members (update != 0), then  sum s^2  and  sum y_new^2  per member
struct UpdateNormsTerm {
  double *y;
  const double *s, *alpha;
  const int32_t *active;
  int64_t B;
  int update;
  __device__ __forceinline__ double2 operator()(int64_t q, int64_t m) const {
    const double sv = __ldg(s + q * B + m);
    double yv = y[q * B + m];
    if (update && (!active || active[m])) {
      yv = yv + (alpha ? alpha[m] : 1.0) * sv;
      y[q * B + m] = yv;
    }
    return make_double2(sv * sv, yv * yv);
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
  const int threads = kThomasThreads;   // one-warp CTAs: even spread over the SMs, thread-private cp.async rings
  const typename MemSrc<NV>::Args src{J, f};
  const unsigned blocks = (unsigned)((P.batch + threads - 1) / threads);
  if (cyclic == 1)   // lower-left corner only
    block_thomas_ll_kernel<MemSrc<NV>, NV, S><<<blocks, threads, 0, st>>>(P, o, src, shift, active, s, workC, workR, status, StepUpd{nullptr, nullptr});
  else if (cyclic)
    block_thomas_kernel<MemSrc<NV>, NV, S, true><<<blocks, threads, 0, st>>>(P, o, src, shift, active, s, workC, workR, status, StepUpd{nullptr, nullptr});
  else
    block_thomas_kernel<MemSrc<NV>, NV, S, false><<<blocks, threads, 0, st>>>(P, o, src, shift, active, s, workC, workR, status, StepUpd{nullptr, nullptr});
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
  if (!thomas_rows_fit(P)) return set_error("sfxm_block_tridiag_solve: batch too large for 32-bit row strides (> 2^24 members)");
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
  // advection-diffusion: one scalar equation per variable, so the off-diagonal entries of a block are exact zeros -
  // which the reference's lil_matrix never stores (simulation.py:357-392 assigns them, scipy drops zero assignments)
  const bool decoupled = P.model == SFXM_MODEL_AD;
  for (int r = 0; r < P.nv; ++r)
    for (int c = 0; c < P.nv; ++c)
      if (P.var_seg[r] == P.var_seg[c] && (!decoupled || r == c)) map.rc[map.L++] = r * P.nv + c;
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

extern "C" int sfxm_update_norms(double *y, const double *s, const double *alpha, const int32_t *active, int32_t update,
                                 int64_t lo, int64_t hi, int64_t batch, double *out2, double *work, void *stream) {
  if (!y || !s || !out2 || !work || hi < lo) return set_error("sfxm_update_norms: bad argument");
  if (batch > 65535LL * 32) return set_error("sfxm_update_norms: batch too large");
  reduce_rows2(UpdateNormsTerm{y, s, alpha, active, batch, update}, lo, hi, batch, out2, work, (cudaStream_t)stream);
  return check_launch("sfxm_update_norms");
}
