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
#include "sfxm_reduce.cuh"

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

// Terms of the deterministic row reduction (sfxm_reduce.cuh):
//   ScaledDiffTerm: ((a - a2) / (atol + |b| rtol))^2                           (select_initial_step's d0, d1, d2)
//   RkErrorTerm:    (((sum_s K_s E_s) * h) / (atol + max(|y|, |y_new|) rtol))^2  (_estimate_error_norm)
struct ScaledDiffTerm {
  const double *a, *a2, *b;
  double atol, rtol;
  int64_t B;
  __device__ __forceinline__ double operator()(int64_t q, int64_t m) const {
    const int64_t e = q * B + m;
    double num = __ldg(a + e);
    if (a2) num = num - __ldg(a2 + e);
    const double t = num / (atol + fabs(__ldg(b + e)) * rtol);
    return t * t;
  }
};
struct RkErrorTerm {
  const double *K, *y, *y_new;
  int64_t stage_stride;
  Coefs E;
  double h, atol, rtol;
  int64_t B;
  __device__ __forceinline__ double operator()(int64_t q, int64_t m) const {
    const int64_t e = q * B + m;
    double d = __ldg(K + e) * E.c[0];
    for (int s = 1; s < E.n; ++s) d = d + __ldg(K + s * stage_stride + e) * E.c[s];
    const double t = (d * h) / (atol + fmax(fabs(__ldg(y + e)), fabs(__ldg(y_new + e))) * rtol);
    return t * t;
  }
};

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
  return (reduce_tiles(n, batch) + 1) * batch * (int64_t)sizeof(double);
}

extern "C" int sfxm_scaled_diff_sumsq(const double *a, const double *a2, const double *b, double atol, double rtol,
                                      int64_t n, int64_t batch, double *out, double *work, void *stream) {
  if (!a || !b || !out || !work) return set_error("sfxm_scaled_diff_sumsq: bad argument");
  reduce_rows(ScaledDiffTerm{a, a2, b, atol, rtol, batch}, 0, n, batch, out, work, (cudaStream_t)stream);
  return check_launch("sfxm_scaled_diff_sumsq");
}

extern "C" int sfxm_rk_error_sumsq(const double *K, int64_t stage_stride, const double *E_host, int32_t nstage,
                                   double h, const double *y, const double *y_new, double atol, double rtol,
                                   int64_t n, int64_t batch, double *out, double *work, void *stream) {
  if (!K || !E_host || !y || !y_new || !out || !work || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_error_sumsq: bad argument");
  Coefs E;
  E.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s) E.c[s] = s < nstage ? E_host[s] : 0.0;
  reduce_rows(RkErrorTerm{K, y, y_new, stage_stride, E, h, atol, rtol, batch}, 0, n, batch, out, work, (cudaStream_t)stream);
  return check_launch("sfxm_rk_error_sumsq");
}
