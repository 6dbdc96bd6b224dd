// sfxm_kernels.cu - residual / ghost fill / finite-difference Jacobian kernels + their C ABI (sm_100a).
//
// Thread mapping (all kernels): threadIdx.x walks ensemble members (fastest-varying in memory, so a
// warp's accesses are one contiguous line per (cell, variable) row), threadIdx.y walks cells.  With
// batch == 1 the block is (1, 256): consecutive threads own consecutive cells of the cell-major vector.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "sfxm_device.cuh"
#include "sfxm_host.h"

namespace sfxm {

// ------------------------------------------------------------------------------------------------
// residual (+ optional fused explicit-Euler update)
// ------------------------------------------------------------------------------------------------
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256)
residual_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                double *__restrict__ r_out, double *__restrict__ y_out, double dt) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (m >= P.batch || i >= P.nx) return;
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + k * param_stride + (param_stride ? m : 0));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][NV], xs[W], r[NV];
#pragma unroll
  for (int s = 0; s < W; ++s) load_slot<NV>(P, v, i - P.nbl + s, xs[s], u[s]);
  cell_residual<Model, SCHEME, NV, W>(prm, u, xs, r);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int64_t q = flat_index(P, i, j) * P.batch + m;
    if (r_out) r_out[q] = r[j];
    if (y_out) y_out[q] = u[P.nbl][j] + dt * r[j];          // sol.y += delta_t * f, simulation.py:459
  }
}

// ------------------------------------------------------------------------------------------------
// Domain view: interior + ghosts, cell-major unsplit (initialize_from_list + apply_BC)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void fill_domain_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                                   const double *__restrict__ y, double *__restrict__ x_all,
                                   double *__restrict__ U_all) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;   // index into the padded mesh
  const int64_t ntot = P.nx + P.nbl + P.nbr;
  if (m >= P.batch || c >= ntot) return;
  View v{y, x_int, m, -1, 0, 0.0};
  double xs, u[NV];
  load_slot<NV>(P, v, c - P.nbl, xs, u);
  if (m == 0 && x_all) x_all[c] = xs;
#pragma unroll
  for (int j = 0; j < NV; ++j) U_all[(c * NV + j) * P.batch + m] = u[j];
}

// ------------------------------------------------------------------------------------------------
// finite-difference Jacobian, all perturbations in one pass (simulation.py:270-394)
// ------------------------------------------------------------------------------------------------
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256)
jacobian_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                double eps, double *__restrict__ J, double *__restrict__ r0_out) {
  const int64_t m = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (m >= P.batch || i >= P.nx) return;
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + k * param_stride + (param_stride ? m : 0));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][NV], xs[W], r0[NV];
#pragma unroll
  for (int s = 0; s < W; ++s) load_slot<NV>(P, v, i - P.nbl + s, xs[s], u[s]);
  cell_residual<Model, SCHEME, NV, W>(prm, u, xs, r0);      // unperturbed rhs, simulation.py:320
  if (r0_out) {
#pragma unroll
    for (int j = 0; j < NV; ++j) r0_out[flat_index(P, i, j) * P.batch + m] = r0[j];
  }
  // does this cell's stencil contain ghost cells?  (then BCs are re-applied after every perturbation,
  // simulation.py:340-342)
  const bool near_left = i - P.nbl < 0;
  const bool near_right = i - P.nbl + (W - 1) >= P.nx;
  double *Jrow = J + (i * W) * (int64_t)(NV * NV) * P.batch + m;
#pragma unroll
  for (int d = 0; d < W; ++d) {
    const int64_t pos = i - P.nbl + d;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      // column cell: in-band interior cell, or the periodic image (extend_band, bc.py:195-240)
      int64_t pc = pos;
      bool live = true;
      if (pos < 0) {
        live = (P.bc_ltype[c] == SFXM_BC_PERIODIC);
        pc = pos + P.nx;
      } else if (pos >= P.nx) {
        live = (P.bc_rtype[c] == SFXM_BC_PERIODIC);
        pc = pos - P.nx;
      }
      double col[NV];
      if (live) {
        double up[W][NV];
#pragma unroll
        for (int s = 0; s < W; ++s)
#pragma unroll
          for (int j = 0; j < NV; ++j) up[s][j] = u[s][j];
        if (!near_left && !near_right) {
          up[d][c] = u[d][c] + eps;                         // simulation.py:337
        } else {
          View vp{y, x_int, m, pc, c, eps};
#pragma unroll
          for (int s = 0; s < W; ++s) {
            const int64_t ps = i - P.nbl + s;
            if (ps < 0) up[s][c] = ghost_u_left(P, vp, (int)(-ps - 1), c);
            else if (ps >= P.nx) up[s][c] = ghost_u_right(P, vp, (int)(ps - P.nx), c);
            else if (ps == pc) up[s][c] = u[s][c] + eps;
          }
        }
        double r1[NV];
        cell_residual<Model, SCHEME, NV, W>(prm, up, xs, r1);   // simulation.py:345
#pragma unroll
        for (int r = 0; r < NV; ++r) col[r] = (r1[r] - r0[r]) / eps;   // simulation.py:354
      } else {
#pragma unroll
        for (int r = 0; r < NV; ++r) col[r] = 0.0;
      }
#pragma unroll
      for (int r = 0; r < NV; ++r)
        if (P.var_seg[r] == P.var_seg[c])                   // split: same-segment rows only, :373-392
          Jrow[((int64_t)(d * NV + r) * NV + c) * P.batch] = col[r];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
static dim3 block_for(int64_t batch) { return batch >= 32 ? dim3(32, 8) : dim3(1, 256); }
static dim3 grid_for(int64_t batch, int64_t ncells, dim3 b) {
  return dim3((unsigned)((ncells + b.y - 1) / b.y), (unsigned)((batch + b.x - 1) / b.x));   // x: cells, y: members
}

enum Op { OP_RESIDUAL, OP_JACOBIAN };

struct Launch {
  Op op;
  const sfxm_problem_t *P;
  const double *x, *params;
  int64_t pstride;
  const double *y;
  double *out0, *out1;
  double scalar;
  cudaStream_t st;
};

template <class Model, int SCHEME, int NV, int W>
static int launch_t(const Launch &L) {
  const sfxm_problem_t &P = *L.P;
  dim3 b = block_for(P.batch), g = grid_for(P.batch, P.nx, b);
  if (g.y > 65535) return set_error("batch too large for one launch (> 65535 * 32 members)");
  if (L.op == OP_RESIDUAL)
    residual_kernel<Model, SCHEME, NV, W><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar);
  else
    jacobian_kernel<Model, SCHEME, NV, W><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1);
  return check_launch("sfxm kernel launch");
}

template <class Model, int NV>
static int launch_scheme(const Launch &L) {
  const sfxm_problem_t &P = *L.P;
  const int W = P.nbl + P.nbr + 1;
  switch (P.scheme) {
    case SFXM_FV_LAX_FRIEDRICHS:
      if (W != 3) break;
      return launch_t<Model, SFXM_FV_LAX_FRIEDRICHS, NV, 3>(L);
    case SFXM_FV_CENTRAL:
      if (W != 3) break;
      return launch_t<Model, SFXM_FV_CENTRAL, NV, 3>(L);
    case SFXM_FV_UPWIND:
      if (W != 3) break;
      if constexpr (Model::kDiagonalDFDU) return launch_t<Model, SFXM_FV_UPWIND, NV, 3>(L);
      else return set_error("UPWIND needs a diagonal dFdU (the reference's eigenvector order is LAPACK-defined)");
    case SFXM_FD_CENTRAL:
      if (W != 3) break;
      return launch_t<Model, SFXM_FD_CENTRAL, NV, 3>(L);
    case SFXM_FD_RIGHT_BIAS:
      if (W != 4) break;
      return launch_t<Model, SFXM_FD_RIGHT_BIAS, NV, 4>(L);
    default:
      return set_error("unknown scheme");
  }
  return set_error("Stencil size does not match boundary conditions");
}

static int validate(const sfxm_problem_t &P) {
  if (P.nv < 1 || P.nv > SFXM_MAX_NV) return set_error("nv must be 1..4");
  if (P.nbl < 1 || P.nbr < 1 || P.nbl > SFXM_MAX_NB || P.nbr > SFXM_MAX_NB) return set_error("nb_left/nb_right must be 1..2");
  if (P.nx < 2 * (P.nbl + P.nbr + 1)) return set_error("mesh smaller than twice the stencil");
  if (P.batch < 1) return set_error("batch must be >= 1");
  return 0;
}

static int dispatch(const Launch &L) {
  const sfxm_problem_t &P = *L.P;
  if (int e = validate(P)) return e;
  switch (P.model) {
    case SFXM_MODEL_AD:
      switch (P.nv) {
        case 1: return launch_scheme<ModelAD<1>, 1>(L);
        case 2: return launch_scheme<ModelAD<2>, 2>(L);
        case 3: return launch_scheme<ModelAD<3>, 3>(L);
        case 4: return launch_scheme<ModelAD<4>, 4>(L);
      }
      break;
    case SFXM_MODEL_EULER_PRIM:
      if (P.nv != 3) return set_error("Euler1D has 3 variables");
      return launch_scheme<ModelEulerPrim, 3>(L);
    case SFXM_MODEL_EULER_CONS:
      if (P.nv != 3) return set_error("Euler1DConservative has 3 variables");
      return launch_scheme<ModelEulerCons, 3>(L);
  }
  return set_error("unknown model");
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_residual(const sfxm_problem_t *prob, const double *x_int, const double *params,
                             int64_t param_stride, const double *y, double *r, double *y_out, double dt,
                             void *stream) {
  if (!prob || !x_int || !params || !y || (!r && !y_out)) return set_error("sfxm_residual: null argument");
  Launch L{OP_RESIDUAL, prob, x_int, params, param_stride, y, r, y_out, dt, (cudaStream_t)stream};
  return dispatch(L);
}

extern "C" int sfxm_jacobian(const sfxm_problem_t *prob, const double *x_int, const double *params,
                             int64_t param_stride, const double *y, double eps, double *J, double *r0,
                             void *stream) {
  if (!prob || !x_int || !params || !y || !J) return set_error("sfxm_jacobian: null argument");
  if (!(eps > 0.0)) return set_error("sfxm_jacobian: epsilon must be positive");
  Launch L{OP_JACOBIAN, prob, x_int, params, param_stride, y, J, r0, eps, (cudaStream_t)stream};
  return dispatch(L);
}

extern "C" int sfxm_fill_domain(const sfxm_problem_t *prob, const double *x_int, const double *y, double *x_all,
                                double *U_all, void *stream) {
  if (!prob || !x_int || !y || !U_all) return set_error("sfxm_fill_domain: null argument");
  const sfxm_problem_t &P = *prob;
  if (int e = validate(P)) return e;
  dim3 b = block_for(P.batch), g = grid_for(P.batch, P.nx + P.nbl + P.nbr, b);
  if (g.y > 65535) return set_error("batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  switch (P.nv) {
    case 1: fill_domain_kernel<1><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 2: fill_domain_kernel<2><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 3: fill_domain_kernel<3><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
    case 4: fill_domain_kernel<4><<<g, b, 0, st>>>(P, x_int, y, x_all, U_all); break;
  }
  return check_launch("sfxm_fill_domain");
}
