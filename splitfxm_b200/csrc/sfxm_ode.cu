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

// ---------------------------------------------------------------------------------------------------------------------
// RK45 with the step-size controller on the device (round 2).  scipy's RK45._step_impl (rk.py:124-187) restated as two
// one-thread kernels around the stage kernels, so that one ATTEMPTED step is a fixed sequence of launches with no host
// decision inside - captured once in a CUDA graph and replayed (splitfxm_b200/ode.py RK45Stepper.integrate_device).
//   ctl[0] t  [1] h_abs  [2] h (this attempt, signed)  [3] t_new  [4] t_bound  [5] max_step  [6] sqrt(n_global)
//   flag[0] done  [1] step_rejected  [2] need_begin  [3] accept  [4] final  [5] failed  [6] nfev  [7] accepted  [8] attempts
// Same IEEE operations as the host controller in ode.py, except x ** -0.2 (CUDA pow vs the host libm's: <= 2 ulp apart).
// ---------------------------------------------------------------------------------------------------------------------
enum { RK_T = 0, RK_HABS = 1, RK_H = 2, RK_TNEW = 3, RK_TBOUND = 4, RK_MAXSTEP = 5, RK_SQRTN = 6 };
enum { RF_DONE = 0, RF_REJECTED = 1, RF_NEED_BEGIN = 2, RF_ACCEPT = 3, RF_FINAL = 4, RF_FAILED = 5, RF_NFEV = 6, RF_ACCEPTED = 7,
       RF_ATTEMPTS = 8 };

__global__ void rk45_begin_kernel(double *__restrict__ ctl, int32_t *__restrict__ flag) {
  if (flag[RF_DONE]) return;
  const double t = ctl[RK_T];
  double h_abs = ctl[RK_HABS];
  const double min_step = 10.0 * fabs(nextafter(t, __longlong_as_double(0x7ff0000000000000LL)) - t);   // rk.py:128
  if (flag[RF_NEED_BEGIN]) {                                                // a new step: rk.py:130-133
    if (h_abs > ctl[RK_MAXSTEP]) h_abs = ctl[RK_MAXSTEP];
    else if (h_abs < min_step) h_abs = min_step;
    flag[RF_REJECTED] = 0;
    flag[RF_NEED_BEGIN] = 0;
  }
  if (h_abs < min_step) {                                                   // rk.py:141-142
    flag[RF_FAILED] = 1;
    flag[RF_DONE] = 1;
    return;
  }
  double h = h_abs;                                                         // direction = +1
  double t_new = t + h;
  if (t_new - ctl[RK_TBOUND] > 0.0) t_new = ctl[RK_TBOUND];                 // rk.py:147-148
  h = t_new - t;
  ctl[RK_H] = h;
  ctl[RK_HABS] = fabs(h);
  ctl[RK_TNEW] = t_new;
  flag[RF_ATTEMPTS] += 1;
}

__global__ void rk45_control_kernel(double *__restrict__ ctl, int32_t *__restrict__ flag, const double *__restrict__ err_sumsq) {
  if (flag[RF_DONE]) {
    flag[RF_ACCEPT] = 0;
    return;
  }
  flag[RF_NFEV] += 6;
  const double error_norm = sqrt(err_sumsq[0]) / ctl[RK_SQRTN];              // common.py norm(): |x|_2 / sqrt(n)
  double h_abs = ctl[RK_HABS];
  if (error_norm < 1.0) {                                                   // rk.py:157-168
    double factor = error_norm == 0.0 ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, -0.2));
    if (flag[RF_REJECTED]) factor = fmin(1.0, factor);
    ctl[RK_HABS] = h_abs * factor;
    flag[RF_ACCEPT] = 1;
    flag[RF_ACCEPTED] += 1;
    flag[RF_FINAL] = (ctl[RK_TNEW] == ctl[RK_TBOUND]) ? 1 : 0;
    ctl[RK_T] = ctl[RK_TNEW];
    flag[RF_NEED_BEGIN] = 1;
    if (flag[RF_FINAL]) flag[RF_DONE] = 1;
  } else {                                                                  // rk.py:169-172
    ctl[RK_HABS] = h_abs * fmax(0.2, 0.9 * pow(error_norm, -0.2));
    flag[RF_REJECTED] = 1;
    flag[RF_ACCEPT] = 0;
  }
}

// lincomb_kernel with h read from the control block; skipped once the integration is done
__global__ void lincomb_dev_kernel(const double *__restrict__ y, const double *__restrict__ K, int64_t stage_stride,
                                   const __grid_constant__ Coefs a, const double *__restrict__ ctl,
                                   const int32_t *__restrict__ flag, double *__restrict__ out, int64_t total) {
  if (flag[RF_DONE]) return;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const double h = ctl[RK_H];
  double acc = __ldg(K + q) * a.c[0];
  for (int s = 1; s < a.n; ++s) acc = acc + __ldg(K + s * stage_stride + q) * a.c[s];
  out[q] = __ldg(y + q) + acc * h;
}

struct RkErrorDevTerm {
  const double *K, *y, *y_new, *ctl;
  int64_t stage_stride;
  Coefs E;
  double atol, rtol;
  int64_t B;
  __device__ __forceinline__ double operator()(int64_t q, int64_t m) const {
    const int64_t e = q * B + m;
    double d = __ldg(K + e) * E.c[0];
    for (int s = 1; s < E.n; ++s) d = d + __ldg(K + s * stage_stride + e) * E.c[s];
    const double t = (d * ctl[RK_H]) / (atol + fmax(fabs(__ldg(y + e)), fabs(__ldg(y_new + e))) * rtol);
    return t * t;
  }
};

// an accepted step: (the last one: result = dense output at its end point, rk.py RkDenseOutput) then y <- y_new,
// K_0 <- K_last (FSAL).  One pass; nothing happens after a rejected attempt.
__global__ void rk45_commit_kernel(double *__restrict__ y, const double *__restrict__ y_new, double *__restrict__ K,
                                   int64_t stage_stride, const __grid_constant__ DenseCoefs P, const double *__restrict__ ctl,
                                   const int32_t *__restrict__ flag, double *__restrict__ result, int64_t total) {
  if (!flag[RF_ACCEPT]) return;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  if (flag[RF_FINAL]) {
    double Q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) Q[k] = K[q] * P.p[0][k];
    for (int s = 1; s < P.n; ++s) {
      const double ks = K[s * stage_stride + q];
#pragma unroll
      for (int k = 0; k < 4; ++k) Q[k] = Q[k] + ks * P.p[s][k];
    }
    const double dot = ((Q[0] + Q[1]) + Q[2]) + Q[3];
    result[q] = ctl[RK_H] * dot + y[q];
  }
  y[q] = y_new[q];
  K[q] = K[(int64_t)(P.n - 1) * stage_stride + q];
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_rk45_begin(double *ctl, int32_t *flag, void *stream) {
  if (!ctl || !flag) return set_error("sfxm_rk45_begin: null argument");
  rk45_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, flag);
  return check_launch("sfxm_rk45_begin");
}

extern "C" int sfxm_rk45_control(double *ctl, int32_t *flag, const double *err_sumsq, void *stream) {
  if (!ctl || !flag || !err_sumsq) return set_error("sfxm_rk45_control: null argument");
  rk45_control_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, flag, err_sumsq);
  return check_launch("sfxm_rk45_control");
}

extern "C" int sfxm_rk_lincomb_dev(const double *y, const double *K, int64_t stage_stride, const double *coef_host,
                                   int32_t nstage, const double *ctl, const int32_t *flag, double *out, int64_t total,
                                   void *stream) {
  if (!y || !K || !coef_host || !ctl || !flag || !out || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_lincomb_dev: bad argument");
  Coefs a;
  a.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s) a.c[s] = s < nstage ? coef_host[s] : 0.0;
  if (total == 0) return 0;
  lincomb_dev_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, K, stage_stride, a, ctl, flag, out, total);
  return check_launch("sfxm_rk_lincomb_dev");
}

extern "C" int sfxm_rk_error_sumsq_dev(const double *K, int64_t stage_stride, const double *E_host, int32_t nstage,
                                       const double *ctl, const double *y, const double *y_new, double atol, double rtol,
                                       int64_t n, int64_t batch, double *out, double *work, void *stream) {
  if (!K || !E_host || !ctl || !y || !y_new || !out || !work || nstage < 1 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk_error_sumsq_dev: bad argument");
  Coefs E;
  E.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s) E.c[s] = s < nstage ? E_host[s] : 0.0;
  reduce_rows(RkErrorDevTerm{K, y, y_new, ctl, stage_stride, E, atol, rtol, batch}, 0, n, batch, out, work, (cudaStream_t)stream);
  return check_launch("sfxm_rk_error_sumsq_dev");
}

extern "C" int sfxm_rk45_commit(double *y, const double *y_new, double *K, int64_t stage_stride, const double *P_host,
                                int32_t nstage, const double *ctl, const int32_t *flag, double *result, int64_t total,
                                void *stream) {
  if (!y || !y_new || !K || !P_host || !ctl || !flag || !result || nstage < 2 || nstage > SFXM_RK_MAX_STAGES)
    return set_error("sfxm_rk45_commit: bad argument");
  DenseCoefs P;
  P.n = nstage;
  for (int s = 0; s < SFXM_RK_MAX_STAGES; ++s)
    for (int k = 0; k < 4; ++k) P.p[s][k] = s < nstage ? P_host[s * 4 + k] : 0.0;
  if (total == 0) return 0;
  rk45_commit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, y_new, K, stage_stride, P, ctl, flag, result,
                                                                                     total);
  return check_launch("sfxm_rk45_commit");
}

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
