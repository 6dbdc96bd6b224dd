// sfxm_stencil_api.cu - the reference's kernel-level entry points on ONE stencil, on the device.
//
// flux.pyx (lax_friedrichs :129-169, upwind :174-224, central :229-257, diff_central :263-299) and derivatives.pyx
// (derivative_callable :26-125, derivative_values :131-192) are `cdef` workers behind the Python wrappers fluxes /
// diffusion_fluxes / Dx / D2x / dx / d2x, which evaluate the user's Python callables F, D, dFdU on the stencil cells
// first.  The wrappers in splitfxm_b200/flux.py and derivatives.py do the same (a callable can only run on the
// host) and hand the VALUES to these kernels; the arithmetic is the one the mesh kernels use (same operation order,
// -fmad=false).  One thread per variable; these are API-parity entry points, not a throughput path.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {

// F, u: [3][nv] row-major; eig: [nv] (UPWIND: eigenvalues of a diagonal dFdU at the centre cell)
__global__ void fv_fluxes_kernel(int scheme, int nv, const double *__restrict__ F, const double *__restrict__ u,
                                 double sigma, const double *__restrict__ eig, double *__restrict__ Fw,
                                 double *__restrict__ Fe) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  const double Fl = F[j], Fc = F[nv + j], Fr = F[2 * nv + j];
  if (scheme == SFXM_FV_LAX_FRIEDRICHS) {
    const double hs = 0.5 * sigma;                                   // flux.pyx:154-167
    Fw[j] = 0.5 * (Fl + Fc) - hs * (u[nv + j] - u[j]);
    Fe[j] = 0.5 * (Fc + Fr) - hs * (u[2 * nv + j] - u[nv + j]);
  } else if (scheme == SFXM_FV_CENTRAL) {
    Fw[j] = 0.5 * (Fl + Fc);                                         // flux.pyx:253-255
    Fe[j] = 0.5 * (Fc + Fr);
  } else {
    const bool pos = eig[j] > 0.0;                                   // flux.pyx:209-218 (R = I)
    Fw[j] = pos ? Fl : Fc;
    Fe[j] = pos ? Fc : Fr;
  }
}

// D: [3][nv], x: [3]
__global__ void fv_diffusion_kernel(int nv, const double *__restrict__ D, const double *__restrict__ x,
                                    double *__restrict__ Dw, double *__restrict__ De) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  const double dxw = 0.5 * (x[1] - x[0]);                            // flux.pyx:285 (note the 0.5)
  const double dxe = 0.5 * (x[2] - x[1]);                            // flux.pyx:294
  Dw[j] = (D[nv + j] - D[j]) / dxw;
  De[j] = (D[2 * nv + j] - D[nv + j]) / dxe;
}

// F: [W][nv], x: [W]; scheme SFXM_FD_CENTRAL (W = 3) or SFXM_FD_RIGHT_BIAS (W = 4); order 1 or 2
__global__ void fd_derivative_kernel(int scheme, int order, int nv, const double *__restrict__ F,
                                     const double *__restrict__ x, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  if (scheme == SFXM_FD_CENTRAL) {
    if (order == 1) {
      out[j] = (F[2 * nv + j] - F[j]) / (x[2] - x[0]);               // derivatives.pyx:75-81
    } else {
      const double dxw = x[1] - x[0], dxe = x[2] - x[1];
      const double Fw = (F[nv + j] - F[j]) / dxw;                    // derivatives.pyx:96-112
      const double Fe = (F[2 * nv + j] - F[nv + j]) / dxe;
      out[j] = (Fe - Fw) / ((dxw + dxe) / 2.0);
    }
  } else {
    const double h = x[1] - x[0];
    const double F0 = F[j], F1 = F[nv + j], F2 = F[2 * nv + j], F3 = F[3 * nv + j];
    if (order == 1) out[j] = (((-F0 - F1) + F2) + F3) / (4.0 * h);   // derivatives.pyx:83-90
    else out[j] = (((F0 - F1) - F2) + F3) / (2.0 * (h * h));         // derivatives.pyx:114-121
  }
}

// One cell's finite-volume residual from the stencil VALUES (fv_transport.py:49-58): the fluxes above, then
// rhs = (S - (1/dx)(Fe - Fw)) + (1/dx)(DFe - DFw) with dx = 0.5 (dxw + dxe).  F, D, u: [3][nv]; x: [3]; S: [nv].
__global__ void fv_cell_residual_kernel(int scheme, int nv, const double *__restrict__ F, const double *__restrict__ D,
                                        const double *__restrict__ u, const double *__restrict__ x,
                                        const double *__restrict__ S, double sigma, const double *__restrict__ eig,
                                        double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  const double Fl = F[j], Fc = F[nv + j], Fr = F[2 * nv + j];
  double Fw, Fe;
  if (scheme == SFXM_FV_LAX_FRIEDRICHS) {
    const double hs = 0.5 * sigma;
    Fw = 0.5 * (Fl + Fc) - hs * (u[nv + j] - u[j]);
    Fe = 0.5 * (Fc + Fr) - hs * (u[2 * nv + j] - u[nv + j]);
  } else if (scheme == SFXM_FV_CENTRAL) {
    Fw = 0.5 * (Fl + Fc);
    Fe = 0.5 * (Fc + Fr);
  } else {
    const bool pos = eig[j] > 0.0;
    Fw = pos ? Fl : Fc;
    Fe = pos ? Fc : Fr;
  }
  const double dxw = x[1] - x[0], dxe = x[2] - x[1];
  const double DFw = (D[nv + j] - D[j]) / (0.5 * dxw);
  const double DFe = (D[2 * nv + j] - D[nv + j]) / (0.5 * dxe);
  const double dx = 0.5 * (dxw + dxe);
  const double inv = 1.0 / dx;
  out[j] = (S[j] - inv * (Fe - Fw)) + inv * (DFe - DFw);
}

// One cell's finite-difference residual (fd_transport.py:46-50): rhs = (S - Dx(F)) + D2x(D).  F, D: [W][nv]; x: [W].
__global__ void fd_cell_residual_kernel(int scheme, int nv, const double *__restrict__ F, const double *__restrict__ D,
                                        const double *__restrict__ x, const double *__restrict__ S,
                                        double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  double d1, d2;
  if (scheme == SFXM_FD_CENTRAL) {
    d1 = (F[2 * nv + j] - F[j]) / (x[2] - x[0]);
    const double dxw = x[1] - x[0], dxe = x[2] - x[1];
    const double Dw = (D[nv + j] - D[j]) / dxw;
    const double De = (D[2 * nv + j] - D[nv + j]) / dxe;
    d2 = (De - Dw) / ((dxw + dxe) / 2.0);
  } else {
    const double h = x[1] - x[0];
    d1 = (((-F[j] - F[nv + j]) + F[2 * nv + j]) + F[3 * nv + j]) / (4.0 * h);
    d2 = (((D[j] - D[nv + j]) - D[2 * nv + j]) + D[3 * nv + j]) / (2.0 * (h * h));
  }
  out[j] = (S[j] - d1) + d2;
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_fv_cell_residual(int32_t scheme, int32_t nv, const double *F, const double *D, const double *u,
                                     const double *x, const double *S, double sigma, const double *eig, double *out,
                                     void *stream) {
  if (nv < 1 || !F || !D || !u || !x || !S || !out) return set_error("sfxm_fv_cell_residual: bad argument");
  if (scheme != SFXM_FV_LAX_FRIEDRICHS && scheme != SFXM_FV_CENTRAL && scheme != SFXM_FV_UPWIND)
    return set_error("Unsupported scheme");
  if (scheme == SFXM_FV_UPWIND && !eig) return set_error("dFdU is required for determining upwind direction.");
  fv_cell_residual_kernel<<<(nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(scheme, nv, F, D, u, x, S, sigma, eig, out);
  return check_launch("sfxm_fv_cell_residual");
}

extern "C" int sfxm_fd_cell_residual(int32_t scheme, int32_t nv, const double *F, const double *D, const double *x,
                                     const double *S, double *out, void *stream) {
  if (nv < 1 || !F || !D || !x || !S || !out) return set_error("sfxm_fd_cell_residual: bad argument");
  if (scheme != SFXM_FD_CENTRAL && scheme != SFXM_FD_RIGHT_BIAS) return set_error("Unsupported scheme");
  fd_cell_residual_kernel<<<(nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(scheme, nv, F, D, x, S, out);
  return check_launch("sfxm_fd_cell_residual");
}

extern "C" int sfxm_fv_fluxes(int32_t scheme, int32_t nv, const double *F, const double *u, double sigma,
                              const double *eig, double *Fw, double *Fe, void *stream) {
  if (nv < 1 || !F || !u || !Fw || !Fe) return set_error("sfxm_fv_fluxes: bad argument");
  if (scheme != SFXM_FV_LAX_FRIEDRICHS && scheme != SFXM_FV_CENTRAL && scheme != SFXM_FV_UPWIND)
    return set_error("Unsupported scheme");
  if (scheme == SFXM_FV_UPWIND && !eig) return set_error("dFdU is required for determining upwind direction.");
  fv_fluxes_kernel<<<(nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(scheme, nv, F, u, sigma, eig, Fw, Fe);
  return check_launch("sfxm_fv_fluxes");
}

extern "C" int sfxm_fv_diffusion_fluxes(int32_t nv, const double *D, const double *x, double *Dw, double *De,
                                        void *stream) {
  if (nv < 1 || !D || !x || !Dw || !De) return set_error("sfxm_fv_diffusion_fluxes: bad argument");
  fv_diffusion_kernel<<<(nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(nv, D, x, Dw, De);
  return check_launch("sfxm_fv_diffusion_fluxes");
}

extern "C" int sfxm_fd_derivative(int32_t scheme, int32_t order, int32_t nv, const double *F, const double *x,
                                  double *out, void *stream) {
  if (nv < 1 || !F || !x || !out) return set_error("sfxm_fd_derivative: bad argument");
  if (scheme != SFXM_FD_CENTRAL && scheme != SFXM_FD_RIGHT_BIAS) return set_error("Unsupported scheme");
  if (order != 1 && order != 2) return set_error("Unsupported order");
  fd_derivative_kernel<<<(nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(scheme, order, nv, F, x, out);
  return check_launch("sfxm_fd_derivative");
}
