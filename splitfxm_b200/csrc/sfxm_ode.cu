// sfxm_ode.cu - explicit Runge-Kutta building blocks for Simulation.evolve (simulation.py:396-472).
//
// The reference hands the residual closure to scipy.integrate.solve_ivp (RK45 by default, :468-469); these
// kernels are the vector arithmetic of scipy/integrate/_ivp/rk.py (rk_step, _estimate_error_norm,
// RkDenseOutput._call_impl) and common.py (select_initial_step, norm), kept in the same operation order,
// with the state resident in HBM.  All are pure streaming kernels (HBM-bound, no reuse).
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {

struct Coefs {
  double c[SFXM_RK_MAX_STAGES];
  int n;
};
struct DenseCoefs {
  double p[SFXM_RK_MAX_STAGES][4];
  int n;
};

// out = y + (sum_s K_s a_s) * h        rk.py: dy = np.dot(K[:s].T, a[:s]) * h ; y + dy
__global__ void lincomb_kernel(const double *__restrict__ y, const double *__restrict__ K, int64_t stage_stride,
                               const __grid_constant__ Coefs a, double h, double *__restrict__ out, int64_t total) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  double acc = __ldg(K + q) * a.c[0];
  for (int s = 1; s < a.n; ++s) acc = acc + __ldg(K + s * stage_stride + q) * a.c[s];
  out[q] = __ldg(y + q) + acc * h;
}

// out = y_old + h * ((((Q0 + Q1) + Q2) + Q3)),  Q_k = sum_s K_s P[s][k]   (dense output at x = 1,
// rk.py: RkDenseOutput._call_impl with Q = K.T.dot(P))
__global__ void dense_end_kernel(const double *__restrict__ y_old, const double *__restrict__ K, int64_t stage_stride,
                                 const __grid_constant__ DenseCoefs P, double h, double *__restrict__ out,
                                 int64_t total) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  double Q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) Q[k] = __ldg(K + q) * P.p[0][k];
  for (int s = 1; s < P.n; ++s) {
    const double ks = __ldg(K + s * stage_stride + q);
#pragma unroll
    for (int k = 0; k < 4; ++k) Q[k] = Q[k] + ks * P.p[s][k];
  }
  const double dot = ((Q[0] + Q[1]) + Q[2]) + Q[3];
  out[q] = h * dot + __ldg(y_old + q);
}

// Per-member sum over q in [0, n) of (num / scale)^2 in fixed order, two stages (see sumsq in sfxm_solve.cu).
//   MODE 0: num = a[q] - (a2 ? a2[q] : 0)            scale = atol + |b[q]| * rtol            (select_initial_step)
//   MODE 1: num = (sum_s K_s E_s) * h                 scale = atol + max(|b|, |b2|) * rtol    (_estimate_error_norm)
template <int MODE>
__global__ void scaled_sumsq_partial_kernel(const double *__restrict__ a, const double *__restrict__ a2,
                                            const double *__restrict__ b, const double *__restrict__ b2,
                                            int64_t stage_stride, const __grid_constant__ Coefs E, double h,
                                            double atol, double rtol, int64_t n, int64_t B, int64_t chunk,
                                            double *__restrict__ partial) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t k = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int64_t q0 = k * chunk;
  if (m >= B || q0 >= n) return;
  const int64_t q1 = q0 + chunk < n ? q0 + chunk : n;
  double acc = 0.0;
  for (int64_t q = q0; q < q1; ++q) {
    const int64_t e = q * B + m;
    double num, scale;
    if (MODE == 0) {
      num = __ldg(a + e);
      if (a2) num = num - __ldg(a2 + e);
      scale = atol + fabs(__ldg(b + e)) * rtol;
    } else {
      double d = __ldg(a + e) * E.c[0];
      for (int s = 1; s < E.n; ++s) d = d + __ldg(a + s * stage_stride + e) * E.c[s];
      num = d * h;
      scale = atol + fmax(fabs(__ldg(b + e)), fabs(__ldg(b2 + e))) * rtol;
    }
    const double t = num / scale;
    acc += t * t;
  }
  partial[k * B + m] = acc;
}

__global__ void sum_partials_kernel(const double *__restrict__ partial, int64_t nchunk, int64_t B,
                                    double *__restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
  double acc = 0.0;
  for (int64_t k = 0; k < nchunk; ++k) acc += partial[k * B + m];
  out[m] = acc;
}

static const int64_t kChunk = 256;

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_rk_lincomb(const double *y, const double *K, int64_t stage_stride, const double *coef_host,
                               int32_t nstage, double h, double *out, int64_t total, void *stream) {
  if (!y || !K || !coef_host || !out || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_lincomb: bad argument");
  Coefs a;
  a.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s) a.c[s] = s < nstage ? coef_host[s] : 0.0;
  if (total == 0) return 0;
  lincomb_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, K, stage_stride, a, h, out, total);
  return check_launch("sfxm_rk_lincomb");
}

extern "C" int sfxm_rk_dense_end(const double *y_old, const double *K, int64_t stage_stride, const double *P_host,
                                 int32_t nstage, double h, double *out, int64_t total, void *stream) {
  if (!y_old || !K || !P_host || !out || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_dense_end: bad argument");
  DenseCoefs P;
  P.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s)
    for (int k = 0; k < 4; ++k) P.p[s][k] = s < nstage ? P_host[s * 4 + k] : 0.0;
  if (total == 0) return 0;
  dense_end_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y_old, K, stage_stride, P, h, out, total);
  return check_launch("sfxm_rk_dense_end");
}

extern "C" int64_t sfxm_scaled_sumsq_workspace_bytes(int64_t n, int64_t batch) {
  return ((n + kChunk - 1) / kChunk) * batch * (int64_t)sizeof(double);
}

static int scaled_launch(int mode, const double *a, const double *a2, const double *b, const double *b2,
                         int64_t stage_stride, const Coefs &E, double h, double atol, double rtol, int64_t n,
                         int64_t batch, double *out, double *work, cudaStream_t st) {
  const int64_t nchunk = (n + kChunk - 1) / kChunk;
  if (nchunk > 0) {
    dim3 blk = batch >= 32 ? dim3(32, 8) : dim3(1, 256);
    dim3 g((unsigned)((nchunk + blk.y - 1) / blk.y), (unsigned)((batch + blk.x - 1) / blk.x));
    if (mode == 0)
      scaled_sumsq_partial_kernel<0><<<g, blk, 0, st>>>(a, a2, b, b2, stage_stride, E, h, atol, rtol, n, batch, kChunk, work);
    else
      scaled_sumsq_partial_kernel<1><<<g, blk, 0, st>>>(a, a2, b, b2, stage_stride, E, h, atol, rtol, n, batch, kChunk, work);
  }
  sum_partials_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(work, nchunk, batch, out);
  return check_launch("sfxm scaled sumsq");
}

extern "C" int sfxm_scaled_diff_sumsq(const double *a, const double *a2, const double *b, double atol, double rtol,
                                      int64_t n, int64_t batch, double *out, double *work, void *stream) {
  if (!a || !b || !out || !work) return set_error("sfxm_scaled_diff_sumsq: bad argument");
  Coefs E{};
  return scaled_launch(0, a, a2, b, nullptr, 0, E, 0.0, atol, rtol, n, batch, out, work, (cudaStream_t)stream);
}

extern "C" int sfxm_rk_error_sumsq(const double *K, int64_t stage_stride, const double *E_host, int32_t nstage,
                                   double h, const double *y, const double *y_new, double atol, double rtol,
                                   int64_t n, int64_t batch, double *out, double *work, void *stream) {
  if (!K || !E_host || !y || !y_new || !out || !work || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_error_sumsq: bad argument");
  Coefs E;
  E.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s) E.c[s] = s < nstage ? E_host[s] : 0.0;
  return scaled_launch(1, K, nullptr, y, y_new, stage_stride, E, h, atol, rtol, n, batch, out, work, (cudaStream_t)stream);
}
