// sfxm_dispatch.cu - C ABI of the residual / Jacobian / domain-fill kernels and the model dispatch.
#include <cuda_runtime.h>
#include "sfxm_kernels.cuh"
#include "sfxm_fused.cuh"

namespace sfxm {

// fused Jacobian + block-Thomas kernels (sfxm_fused_*.cu); -2 = combination not instantiated
int fused_ad3(const FusedLaunch &L);
int fused_euler_cons(const FusedLaunch &L);
int fused_euler_prim(const FusedLaunch &L);

// one translation unit per model (sfxm_inst_*.cu)
int launch_ad1(const Launch &L);
int launch_ad2(const Launch &L);
int launch_ad3(const Launch &L);
int launch_ad4(const Launch &L);
int launch_euler_prim(const Launch &L);
int launch_euler_cons(const Launch &L);

// user models: plugins built from user CUDA source (splitfxm_b200/custom.py) register a launcher per slot
typedef int (*user_launch_fn)(const void *launch);
static user_launch_fn g_user[SFXM_MAX_USER_MODELS] = {nullptr};

static int validate(const sfxm_problem_t &P) {
  if (P.nv < 1 || P.nv > SFXM_MAX_NV) return set_error("nv must be 1..4");
  if (P.nbl < 1 || P.nbr < 1 || P.nbl > SFXM_MAX_NB || P.nbr > SFXM_MAX_NB) return set_error("nb_left/nb_right must be 1..2");
  // the reference has no minimum (tests/test_simulation.py drives Domain.from_size(5, 1, 1, ...)); the kernels need the
  // periodic wrap columns of a block row to be distinct from its in-band columns: nx >= stencil width
  if (P.nx < P.nbl + P.nbr + 1) return set_error("mesh smaller than the stencil");
  if (P.batch < 1) return set_error("batch must be >= 1");
  return 0;
}

static int dispatch(const Launch &L) {
  const sfxm_problem_t &P = *L.P;
  if (int e = validate(P)) return e;
  switch (P.model) {
    case SFXM_MODEL_AD:
      switch (P.nv) {
        case 1: return launch_ad1(L);
        case 2: return launch_ad2(L);
        case 3: return launch_ad3(L);
        case 4: return launch_ad4(L);
      }
      break;
    case SFXM_MODEL_EULER_PRIM:
      if (P.nv != 3) return set_error("Euler1D has 3 variables");
      return launch_euler_prim(L);
    case SFXM_MODEL_EULER_CONS:
      if (P.nv != 3) return set_error("Euler1DConservative has 3 variables");
      return launch_euler_cons(L);
  }
  if (P.model >= SFXM_MODEL_USER0 && P.model < SFXM_MODEL_USER0 + SFXM_MAX_USER_MODELS) {
    user_launch_fn fn = g_user[P.model - SFXM_MODEL_USER0];
    if (!fn) return set_error("user model slot is not registered");
    return fn(&L);
  }
  return set_error("unknown model");
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_residual(const sfxm_problem_t *prob, const double *x_int, const double *params,
                             int64_t param_stride, const double *geom, const double *y, double *r, double *y_out,
                             double dt, void *stream) {
  if (!prob || !x_int || !params || !y || (!r && !y_out)) return set_error("sfxm_residual: null argument");
  Launch L{OP_RESIDUAL, prob, x_int, params, param_stride, y, r, y_out, dt, (cudaStream_t)stream, geom, 0, SFXM_MAX_NV + 1};
  return dispatch(L);
}

extern "C" int64_t sfxm_residual_warp_count(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const int per_warp = 32 - (prob->nbl + prob->nbr);
  return (prob->nx + per_warp - 1) / per_warp;
}

extern "C" int sfxm_residual_warps(const sfxm_problem_t *prob, const double *x_int, const double *params,
                                   int64_t param_stride, const double *y, double *r, double *y_out, double dt,
                                   int64_t warp_lo, int64_t warp_hi, void *stream) {
  if (!prob || !x_int || !params || !y || (!r && !y_out)) return set_error("sfxm_residual_warps: null argument");
  if (warp_lo < -1 || warp_hi <= warp_lo || warp_hi < 1) return set_error("sfxm_residual_warps: empty warp range");
  Launch L{OP_RESIDUAL, prob, x_int, params, param_stride, y, r, y_out, dt, (cudaStream_t)stream, nullptr, 0, SFXM_MAX_NV + 1,
           warp_lo, warp_hi};
  return dispatch(L);
}

extern "C" int sfxm_jacobian(const sfxm_problem_t *prob, const double *x_int, const double *params,
                             int64_t param_stride, const double *geom, const double *y, double eps, double *J,
                             double *r0, void *stream) {
  if (!prob || !x_int || !params || !y || !J) return set_error("sfxm_jacobian: null argument");
  if (!(eps > 0.0)) return set_error("sfxm_jacobian: epsilon must be positive");
  Launch L{OP_JACOBIAN, prob, x_int, params, param_stride, y, J, r0, eps, (cudaStream_t)stream, geom, 0, SFXM_MAX_NV + 1};
  return dispatch(L);
}

extern "C" int sfxm_jacobian_segments(const sfxm_problem_t *prob, const double *x_int, const double *params,
                                      int64_t param_stride, const double *geom, const double *y, double eps,
                                      int32_t seg_lo, int32_t seg_hi, double *J, double *r0, void *stream) {
  if (!prob || !x_int || !params || !y || !J) return set_error("sfxm_jacobian_segments: null argument");
  if (!(eps > 0.0)) return set_error("sfxm_jacobian_segments: epsilon must be positive");
  if (seg_lo < 0 || seg_hi <= seg_lo) return set_error("sfxm_jacobian_segments: empty segment range");
  Launch L{OP_JACOBIAN, prob, x_int, params, param_stride, y, J, r0, eps, (cudaStream_t)stream, geom, seg_lo, seg_hi};
  return dispatch(L);
}

extern "C" int sfxm_fill_domain(const sfxm_problem_t *prob, const double *x_int, const double *y, double *x_all,
                                double *U_all, void *stream) {
  if (!prob || !x_int || !y || !U_all) return set_error("sfxm_fill_domain: null argument");
  const sfxm_problem_t &P = *prob;
  if (int e = validate(P)) return e;
  dim3 b = block_for(P.batch), g = grid_for(P.batch, P.nx + P.nbl + P.nbr, b);
  if (g.z > 65535) return set_error("mesh too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  switch (P.nv) {
    case 1: fill_domain_kernel<1><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 2: fill_domain_kernel<2><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 3: fill_domain_kernel<3><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 4: fill_domain_kernel<4><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
  }
  return check_launch("sfxm_fill_domain");
}

extern "C" int sfxm_geometry(const sfxm_problem_t *prob, const double *x_int, double *geom, void *stream) {
  if (!prob || !x_int || !geom) return set_error("sfxm_geometry: null argument");
  const sfxm_problem_t &P = *prob;
  if (int e = validate(P)) return e;
  const int W = P.nbl + P.nbr + 1;
  const unsigned blocks = (unsigned)((P.nx + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (P.scheme) {
    case SFXM_FV_LAX_FRIEDRICHS:
    case SFXM_FV_CENTRAL:
    case SFXM_FV_UPWIND:
      if (W != 3) return set_error("Stencil size does not match boundary conditions");
      geometry_kernel<SFXM_FV_LAX_FRIEDRICHS, 3><<<blocks, 256, 0, st>>>(P, x_int, geom);   // same geometry for the 3 FV schemes
      break;
    case SFXM_FD_CENTRAL:
      if (W != 3) return set_error("Stencil size does not match boundary conditions");
      geometry_kernel<SFXM_FD_CENTRAL, 3><<<blocks, 256, 0, st>>>(P, x_int, geom);
      break;
    case SFXM_FD_RIGHT_BIAS:
      if (W != 4) return set_error("Stencil size does not match boundary conditions");
      geometry_kernel<SFXM_FD_RIGHT_BIAS, 4><<<blocks, 256, 0, st>>>(P, x_int, geom);
      break;
    default:
      return set_error("unknown scheme");
  }
  return check_launch("sfxm_geometry");
}

extern "C" int sfxm_register_user_model(int32_t slot, void *launch_fn) {
  if (slot < 0 || slot >= SFXM_MAX_USER_MODELS || !launch_fn) return set_error("sfxm_register_user_model: bad argument");
  g_user[slot] = (user_launch_fn)launch_fn;
  return 0;
}

// 1 if sfxm_newton_solve_fused has kernels for this problem (model / scheme / stencil), else 0
extern "C" int sfxm_newton_fused_supported(const sfxm_problem_t *prob) {
  if (!prob) return 0;
  const sfxm_problem_t &P = *prob;
  if (P.nbl != 1 || P.nbr != 1 || P.nx < 3) return 0;
  if (P.model == SFXM_MODEL_AD && P.nv == 3)
    return P.scheme == SFXM_FV_LAX_FRIEDRICHS || P.scheme == SFXM_FV_CENTRAL || P.scheme == SFXM_FD_CENTRAL;
  if (P.model == SFXM_MODEL_EULER_CONS || P.model == SFXM_MODEL_EULER_PRIM) return P.nv == 3 && P.scheme == SFXM_FV_LAX_FRIEDRICHS;
  return 0;
}

static int newton_fused_common(const sfxm_problem_t *prob, const double *x_int, const double *params, int64_t param_stride,
                               const double *geom, const double *y, double eps, const double *shift, const int32_t *active,
                               int32_t seg, double *s, double *r0, double *work, int32_t *status, double *y_upd,
                               double *norms, void *stream) {
  if (!prob || !x_int || !params || !y || !s || !work) return set_error("sfxm_newton_solve_fused: null argument");
  if (!(eps > 0.0)) return set_error("sfxm_newton_solve_fused: epsilon must be positive");
  if (int e = validate(*prob)) return e;
  if (!sfxm_newton_fused_supported(prob)) return set_error("sfxm_newton_solve_fused: no fused kernel for this model / scheme");
  if (prob->nx >= ((int64_t)1 << 32) || prob->batch * (int64_t)(SFXM_MAX_NV * (1 + SFXM_MAX_NV) * 8) >= ((int64_t)1 << 32))
    return set_error("sfxm_newton_solve_fused: batch too large for 32-bit row strides (> 2^24 members)");
  FusedLaunch L{prob, FusedArgs{x_int, params, param_stride, y, eps, geom, 0u, r0}, seg, shift, active, s, work, status,
                (cudaStream_t)stream, StepUpd{y_upd, norms}};
  int e = -2;
  switch (prob->model) {
    case SFXM_MODEL_AD: e = fused_ad3(L); break;
    case SFXM_MODEL_EULER_CONS: e = fused_euler_cons(L); break;
    case SFXM_MODEL_EULER_PRIM: e = fused_euler_prim(L); break;
  }
  if (e == -2) return set_error("sfxm_newton_solve_fused: no fused kernel for this model / scheme");
  return e;
}

extern "C" int sfxm_newton_solve_fused(const sfxm_problem_t *prob, const double *x_int, const double *params,
                                       int64_t param_stride, const double *geom, const double *y, double eps,
                                       const double *shift, const int32_t *active, int32_t seg, double *s, double *r0,
                                       double *work, int32_t *status, void *stream) {
  return newton_fused_common(prob, x_int, params, param_stride, geom, y, eps, shift, active, seg, s, r0, work, status, nullptr,
                             nullptr, stream);
}

extern "C" int sfxm_newton_step_fused(const sfxm_problem_t *prob, const double *x_int, const double *params,
                                      int64_t param_stride, const double *geom, double *y, double eps, const double *shift,
                                      const int32_t *active, int32_t seg, double *s, double *r0, double *work, int32_t *status,
                                      double *norms2, void *stream) {
  if (!norms2) return set_error("sfxm_newton_step_fused: null argument");
  return newton_fused_common(prob, x_int, params, param_stride, geom, y, eps, shift, active, seg, s, r0, work, status, y, norms2,
                             stream);
}
