// sfxm_device.cuh - device-side building blocks of the SplitFXM hot path (sm_100a, FP64).
//
// Everything here is arithmetic that must reproduce the reference's IEEE operation order
// (SURVEY.md 8a); the translation unit is compiled with -fmad=false so no multiply-add is fused.
// Citations are to the reference tree (gpavanb1/SplitFXM v0.5.2).
#pragma once
#include <cstdint>
#include "../../include/sfxm_b200.h"

namespace sfxm {

// ------------------------------------------------------------------------------------------------
// State access: flat vector, member-fastest (include/sfxm_b200.h "Data layout")
// ------------------------------------------------------------------------------------------------
struct View {
  const double *__restrict__ y;   // flat state
  const double *__restrict__ x;   // interior positions
  int64_t m;                      // member
  int64_t pc;                     // perturbed cell (-1: none)   simulation.py:337
  int pv;                         // perturbed variable
  double eps;
};

__device__ __forceinline__ int64_t flat_index(const sfxm_problem_t &P, int64_t cell, int var) {
  return P.var_base[var] + cell * (int64_t)P.var_stride[var] + P.var_off[var];
}

__device__ __forceinline__ double ld_u(const sfxm_problem_t &P, const View &v, int64_t cell, int var) {
  double val = __ldg(v.y + flat_index(P, cell, var) * P.batch + v.m);
  if (cell == v.pc && var == v.pv) val = val + v.eps;   // current_value + epsilon, simulation.py:337
  return val;
}

// ------------------------------------------------------------------------------------------------
// Ghost cells (bc.py:6-159).  Ghost g = 0 is the one nearest the interior (domain.py:167-177).
// ------------------------------------------------------------------------------------------------
// position an outflow/neumann/dirichlet variable writes before it computes its value
__device__ __forceinline__ double mirror_x_left(const sfxm_problem_t &P, const double *x, int g) {
  return P.xmin - (__ldg(x + g + 1) - P.xmin);                       // bc.py:72-73, 113-114, 140-141
}
__device__ __forceinline__ double mirror_x_right(const sfxm_problem_t &P, const double *x, int g) {
  return P.xmax + (P.xmax - __ldg(x + P.nx - 2 - g));                // bc.py:86-87, 126-127, 150-151
}
// position finally stored in the (shared) ghost cell: the last variable in `bcs` order wins
__device__ __forceinline__ double ghost_x_left(const sfxm_problem_t &P, const double *x, int g) {
  if (P.gx_left == SFXM_GX_PERIODIC) return P.xmin - (P.xmax - __ldg(x + P.nx - 2 - g));   // bc.py:55-56
  if (P.gx_left == SFXM_GX_MIRROR) return mirror_x_left(P, x, g);
  return P.gx_keep_left[g];
}
__device__ __forceinline__ double ghost_x_right(const sfxm_problem_t &P, const double *x, int g) {
  if (P.gx_right == SFXM_GX_PERIODIC) return P.xmax + (__ldg(x + g + 1) - P.xmin);         // bc.py:63-64
  if (P.gx_right == SFXM_GX_MIRROR) return mirror_x_right(P, x, g);
  return P.gx_keep_right[g];
}

__device__ inline double ghost_u_left(const sfxm_problem_t &P, const View &v, int g, int j) {
  const int t = P.bc_ltype[j];
  if (t == SFXM_BC_HALO) return __ldg(P.halo_left + (int64_t)(g * P.nv + j) * P.batch + v.m);   // neighbouring slab
  if (t == SFXM_BC_PERIODIC) return ld_u(P, v, P.nx - 1 - g, j);            // bc.py:57
  if (t == SFXM_BC_DIRICHLET || t == SFXM_BC_KEEP) return P.bc_lval[j][g];  // bc.py:143
  const double u0 = ld_u(P, v, 0, j);
  double slope;
  if (t == SFXM_BC_NEUMANN) {
    slope = P.bc_lval[j][0];                                                // bc.py:110
  } else {                                                                  // outflow, bc.py:75-76
    const double dy = ld_u(P, v, 1, j) - u0;
    const double dx = __ldg(v.x + 1) - __ldg(v.x);
    slope = dy / dx;
  }
  double prev = u0, px = __ldg(v.x), val = u0;
  for (int i = 0; i <= g; ++i) {                                            // bc.py:77-80, 116-119
    const double xg = mirror_x_left(P, v.x, i);
    val = prev - slope * (px - xg);
    prev = val;
    px = xg;
  }
  return val;
}

__device__ inline double ghost_u_right(const sfxm_problem_t &P, const View &v, int g, int j) {
  const int t = P.bc_rtype[j];
  if (t == SFXM_BC_HALO) return __ldg(P.halo_right + (int64_t)(g * P.nv + j) * P.batch + v.m);  // neighbouring slab
  if (t == SFXM_BC_PERIODIC) return ld_u(P, v, g, j);                       // bc.py:65
  if (t == SFXM_BC_DIRICHLET || t == SFXM_BC_KEEP) return P.bc_rval[j][g];  // bc.py:153
  const int64_t ihi = P.nx - 1;
  const double uh = ld_u(P, v, ihi, j);
  double slope;
  if (t == SFXM_BC_NEUMANN) {
    slope = P.bc_rval[j][0];                                                // bc.py:123
  } else {                                                                  // outflow, bc.py:89-90
    const double dy = uh - ld_u(P, v, ihi - 1, j);
    const double dx = __ldg(v.x + ihi) - __ldg(v.x + ihi - 1);
    slope = dy / dx;
  }
  double prev = uh, px = __ldg(v.x + ihi), val = uh;
  for (int i = 0; i <= g; ++i) {                                            // bc.py:91-94, 129-132
    const double xg = mirror_x_right(P, v.x, i);
    val = prev + slope * (xg - px);
    prev = val;
    px = xg;
  }
  return val;
}

// fused halo exchange: the updated value of a slab-boundary cell also goes into the neighbour's ghost buffer
// (peer-mapped memory of another GPU; plain global stores that travel over NVLink)
__device__ __forceinline__ void push_halo(const sfxm_problem_t &P, int64_t i, int j, int64_t m, double val) {
  if (P.peer_left_dst != nullptr && i < P.nbr)            // my first cells are the left neighbour's right ghosts
    P.peer_left_dst[(int64_t)(i * P.nv + j) * P.batch + m] = val;
  if (P.peer_right_dst != nullptr && i >= P.nx - P.nbl)   // my last cells are the right neighbour's left ghosts
    P.peer_right_dst[(int64_t)((P.nx - 1 - i) * P.nv + j) * P.batch + m] = val;
}

// stencil slot s of cell i -> position and values (system.py:58-61: cells[i - nbl .. i + nbr])
// (j0: first variable of the group this thread handles - 0 for the coupled kernels, the thread's own variable for
// the one-variable-per-thread kernels of decoupled models)
template <int NV>
__device__ __forceinline__ void load_slot(const sfxm_problem_t &P, const View &v, int64_t pos, double &xs,
                                          double (&u)[NV], int j0 = 0) {
  if (pos < 0) {
    const int g = (int)(-pos - 1);
    xs = ghost_x_left(P, v.x, g);
#pragma unroll
    for (int j = 0; j < NV; ++j) u[j] = ghost_u_left(P, v, g, j0 + j);
  } else if (pos >= P.nx) {
    const int g = (int)(pos - P.nx);
    xs = ghost_x_right(P, v.x, g);
#pragma unroll
    for (int j = 0; j < NV; ++j) u[j] = ghost_u_right(P, v, g, j0 + j);
  } else {
    xs = __ldg(v.x + pos);
#pragma unroll
    for (int j = 0; j < NV; ++j) u[j] = ld_u(P, v, pos, j0 + j);
  }
}

// Whole stencil of cell i for member m.  off[j] = element offset of unknown (i, j) in the flat vectors (the same
// for y, r and y_out).  Interior cells (no ghost in the stencil) take straight strided loads: one 64-bit multiply
// per variable, then pointer steps - the general per-slot path costs ~40 integer instructions per slot.
template <int NV, int W>
__device__ __forceinline__ bool gather_stencil(const sfxm_problem_t &P, const View &v, int64_t i, bool need_x,
                                               double (&u)[W][NV], double (&xs)[W], int64_t (&off)[NV], int j0 = 0) {
  const int64_t B = P.batch;
  const int nbl = (W == 3) ? 1 : P.nbl;   // a 3-cell stencil is always 1 + 1 + 1 (validate())
  const bool interior = (i >= nbl) && (i + (W - 1 - nbl) < P.nx);
  int64_t rs[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    rs[j] = (int64_t)P.var_stride[j0 + j] * B;
    off[j] = (P.var_base[j0 + j] + P.var_off[j0 + j]) * B + v.m + i * rs[j];
  }
  if (interior) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double *p = v.y + (off[j] - nbl * rs[j]);
#pragma unroll
      for (int s = 0; s < W; ++s) {
        u[s][j] = __ldg(p);
        p += rs[j];
      }
    }
    if (need_x) {
#pragma unroll
      for (int s = 0; s < W; ++s) xs[s] = __ldg(v.x + (i - nbl + s));
    }
  } else {
#pragma unroll
    for (int s = 0; s < W; ++s) load_slot<NV>(P, v, i - nbl + s, xs[s], u[s], j0);
  }
  return interior;
}

// ------------------------------------------------------------------------------------------------
// Models: F (advective flux), D (diffusive quantity), spectral radius of dFdU.  S = 0 for all three.
// ------------------------------------------------------------------------------------------------
template <int NV>
struct ModelAD {   // models/advection_diffusion.py:28-31 ; params = (c, nu)
  static constexpr int kParams = 2;
  static constexpr bool kDiagonalDFDU = true;
  static constexpr bool kHasDiffusion = true;
  // every variable obeys the same scalar equation: row r of the Jacobian depends on variable r only, and
  // perturbing variable c leaves the other rows bit-identical (exact zeros in the reference as well)
  static constexpr bool kDecoupled = true;
  using Scalar = ModelAD<1>;
  __device__ __forceinline__ static void F(const double *p, const double (&u)[NV], double (&f)[NV]) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = p[0] * u[j];
  }
  __device__ __forceinline__ static void D(const double *p, const double (&u)[NV], double (&d)[NV]) {
#pragma unroll
    for (int j = 0; j < NV; ++j) d[j] = p[1] * u[j];
  }
  // max |eigvals(diag(c))| = |c| exactly (flux.pyx:98-99)
  __device__ __forceinline__ static double sigma(const double *p, const double (&)[NV]) { return fabs(p[0]); }
  __device__ __forceinline__ static double eig(const double *p, const double (&)[NV], int) { return p[0]; }
};

// one real root of l^3 - t l^2 + m l - d, polished; then deflation -> max modulus of the three roots
__device__ inline double spectral_radius_cubic(double t, double m, double d) {
  const double third = 1.0 / 3.0;
  const double sh = t * third;
  const double p = m - t * t * third;
  const double q = -2.0 * sh * sh * sh + sh * m - d;
  const double disc = 0.25 * q * q + p * p * p / 27.0;
  double z;
  if (disc > 0.0) {
    const double sq = sqrt(disc);
    z = cbrt(-0.5 * q + sq) + cbrt(-0.5 * q - sq);
  } else if (p < 0.0) {
    const double rr = sqrt(-p * third);
    double arg = 1.5 * q / (p * rr);
    arg = fmin(1.0, fmax(-1.0, arg));
    z = 2.0 * rr * cos(acos(arg) * third);
  } else {
    z = 0.0;
  }
  double l1 = z + sh;
  for (int it = 0; it < 3; ++it) {   // Newton polish on the characteristic polynomial
    const double f = ((l1 - t) * l1 + m) * l1 - d;
    const double fp = (3.0 * l1 - 2.0 * t) * l1 + m;
    if (fp != 0.0) l1 -= f / fp;
  }
  const double b = l1 - t;          // (l - l1)(l^2 + b l + c)
  const double c = m + l1 * b;
  const double dq = b * b - 4.0 * c;
  double other;
  if (dq < 0.0) {
    other = sqrt(fmax(c, 0.0));     // complex pair: |l|^2 = c
  } else {
    const double s = sqrt(dq);
    other = fmax(fabs(0.5 * (-b + s)), fabs(0.5 * (-b - s)));
  }
  return fmax(fabs(l1), other);
}

struct ModelEulerPrim {   // tutorials/sod_shock/euler1d.py:27-50 ; params = (gamma); W = (rho, u, p)
  static constexpr int kParams = 1;
  static constexpr bool kDiagonalDFDU = false;
  static constexpr bool kHasDiffusion = false;
  static constexpr bool kDecoupled = false;
  using Scalar = ModelEulerPrim;
  __device__ __forceinline__ static void F(const double *pr, const double (&w)[3], double (&f)[3]) {
    const double g = pr[0], rho = w[0], u = w[1], p = w[2];
    const double u2 = u * u;
    const double E = p / (g - 1.0) + 0.5 * rho * u2;
    f[0] = rho * u;
    f[1] = rho * u2 + p;
    f[2] = u * (E + p);
  }
  __device__ __forceinline__ static void D(const double *, const double (&)[3], double (&d)[3]) {
    d[0] = d[1] = d[2] = 0.0;
  }
  // max |eigvals(dF/dW)| - LAPACK dgeev in the reference (flux.pyx:98-99); closed-form cubic here
  __device__ __forceinline__ static double sigma(const double *pr, const double (&w)[3]) {
    const double g = pr[0], rho = w[0], u = w[1], p = w[2];
    const double u2 = u * u;
    const double a00 = u, a01 = rho;                                   // a02 = 0
    const double a10 = u2, a11 = 2.0 * rho * u, a12 = 1.0;
    const double a20 = 0.5 * (u2 * u), a21 = 1.5 * rho * u2 + g * p / (g - 1.0), a22 = g * u / (g - 1.0);
    const double t = a00 + a11 + a22;
    const double m = (a00 * a11 - a01 * a10) + (a00 * a22) + (a11 * a22 - a12 * a21);
    const double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20);
    return spectral_radius_cubic(t, m, d);
  }
  __device__ __forceinline__ static double eig(const double *, const double (&)[3], int) { return 0.0; }
};

struct ModelEulerCons {   // tutorials/conservative_sod_shock/euler1d.py:28-63 ; params = (gamma)
  static constexpr int kParams = 1;
  static constexpr bool kDiagonalDFDU = false;
  static constexpr bool kHasDiffusion = false;
  static constexpr bool kDecoupled = false;
  using Scalar = ModelEulerCons;
  __device__ __forceinline__ static void F(const double *pr, const double (&U)[3], double (&f)[3]) {
    const double g = pr[0], U1 = U[0], U2 = U[1], U3 = U[2];
    f[0] = U2;
    f[1] = 0.5 * (3.0 - g) * (U2 * U2) / U1 + (g - 1.0) * U3;
    f[2] = g * U2 * U3 / U1 - 0.5 * (g - 1.0) * ((U2 * U2 * U2) / (U1 * U1));
  }
  __device__ __forceinline__ static void D(const double *, const double (&)[3], double (&d)[3]) {
    d[0] = d[1] = d[2] = 0.0;
  }
  // eigenvalues of dF/dU are u, u +- a: spectral radius |u| + a (LAPACK in the reference, flux.pyx:98-99;
  // agreement ~1e-15 relative)
  __device__ __forceinline__ static double sigma(const double *pr, const double (&U)[3]) {
    const double g = pr[0];
    const double u = U[1] / U[0];
    const double p = (g - 1.0) * (U[2] - 0.5 * U[1] * u);
    return fabs(u) + sqrt(g * p / U[0]);
  }
  __device__ __forceinline__ static double eig(const double *, const double (&)[3], int) { return 0.0; }
};

// ------------------------------------------------------------------------------------------------
// Exact division by a divisor whose reciprocal is already known.
//
// The reference divides (flux.pyx:288,297; derivatives.pyx:81,90,102,109,112,121; simulation.py:354), so the
// parity target is the correctly rounded IEEE quotient.  nvcc expands a/b into ~35 FP64-pipe instructions
// (reciprocal seed + Newton iterations + fix-up); the divisors here are mesh spacings and the FD epsilon,
// shared by every perturbation of a cell, so their correctly rounded reciprocals rb = RN(1/b) are computed
// once and each quotient costs one multiplication and two FMAs:
//      q0 = RN(a rb),   r = RN(a - q0 b) (exact),   q = RN(q0 + r rb)
// Why one correction step is enough (round 2; rounds 1-2a ran two).  Scale so that b, a/b are in [1, 2)
// (u = 2^-52).  With e = 1 - b rb, |e| < b 2^-54, the FMA's argument is exactly  q0 + r rb = a/b - (a/b - q0) e.
//  * a/b within u/4 of a rounding boundary m (midpoint of two doubles): a rb = (a/b)(1 - e) lies within
//    (a/b) b 2^-54 < u of a/b, so q0 is one of m +- u/2, |a/b - q0| <= u/2 + |a/b - m| and r (a multiple of
//    2^-104 below 1.5 2^-52) is exact.  a/b cannot BE a midpoint: a - m b is a non-zero multiple of 2^-105, so
//    |a/b - m| >= 2^-105 / b, while the perturbation is below (u/2) b 2^-54 (1 + 2^-52) = b 2^-107 (1 + ..),
//    smaller because b^2 < 4 (1 - 2^-53): the argument stays on a/b's side of m and rounds like a/b.
//  * elsewhere |a/b - q0| <= 1.5 u, the perturbation (with a possibly inexact r) is below 2^-51 u << u/4.
// This holds while nothing under- or overflows, so each thread checks ONCE that its inputs are "sane" -
// stencil values and parameters zero or within 2^+-120, divisors within 2^+-120: then every numerator (a
// difference of such values: zero or >= 2^-173) and every quotient stays far inside the normal range - and
// otherwise (FAST = false instantiation) all its quotients use the IEEE division.  sfxm_selftest_div compares
// the two on random, adversarial and nearest-to-midpoint operands (tests/test_gpu_div.py); tests/test_div_known.py
// checks the same three operations in exact rational arithmetic on the CPU.
// ------------------------------------------------------------------------------------------------
// out of line on purpose: inlined, the compiler if-converts the rare path and executes the full IEEE division
// (MUFU.RCP64H + ~10 DFMA + fix-up) for every quotient, which is what div_known exists to avoid
static __device__ __noinline__ double div_ieee(double a, double b) { return a / b; }

template <bool FAST>
__device__ __forceinline__ double div_known(double a, double b, double rb) {
  if constexpr (!FAST) return div_ieee(a, b);
#ifdef SFXM_DIV_TWO_CORRECTIONS
  const double q0 = a * rb;
  const double q1 = __fma_rn(__fma_rn(-q0, b, a), rb, q0);
  return __fma_rn(__fma_rn(-q1, b, a), rb, q1);
#else
  const double q0 = a * rb;
  return __fma_rn(__fma_rn(-q0, b, a), rb, q0);
#endif
}

// exponent of v within 2^+-span, or v == 0 (four integer instructions: mask, subtract, or-with-low-word, compare)
__device__ __forceinline__ bool sane(double v, unsigned span) {
  const unsigned a = (unsigned)__double2hiint(v) & 0x7fffffffu;
  // bitwise, not short-circuit: keeps the check branch-free (integer pipe only)
  return ((a - ((1023u - span) << 20)) < ((2u * span) << 20)) | ((a | (unsigned)__double2loint(v)) == 0u);
}

// ------------------------------------------------------------------------------------------------
// Per-cell residual (fv_transport.py:49-58, fd_transport.py:46-50, flux.pyx, derivatives.pyx)
// u[s][j] : stencil slots s = 0..W-1 (slot 1 is the centre for the FV schemes, fv_transport.py:49)
// Geom    : everything that depends on the positions only (shared by all perturbations of a cell)
// ------------------------------------------------------------------------------------------------
template <int SCHEME, int W>
struct Geom {
  double d[4];    // divisors, in the order the reference forms them
  double rd[4];   // RN(1 / d[k])
  double inv;     // FV: 1 / dx (the reference multiplies by it, fv_transport.py:58)
  bool fast;      // divisors sane (and, once and_inputs() was called, the stencil values too): use div_known's FMAs
  __device__ __forceinline__ explicit Geom(const double (&xs)[W]) {
    d[0] = d[1] = d[2] = d[3] = 1.0;
    inv = 0.0;
    if constexpr (SCHEME == SFXM_FV_LAX_FRIEDRICHS || SCHEME == SFXM_FV_CENTRAL || SCHEME == SFXM_FV_UPWIND) {
      const double dxw = xs[1] - xs[0];
      const double dxe = xs[2] - xs[1];
      const double dx = 0.5 * (dxw + dxe);                 // fv_transport.py:50-52
      inv = 1.0 / dx;
      d[0] = 0.5 * (xs[1] - xs[0]);                        // flux.pyx:285 (note the 0.5)
      d[1] = 0.5 * (xs[2] - xs[1]);                        // flux.pyx:294
    } else if constexpr (SCHEME == SFXM_FD_CENTRAL) {
      d[0] = xs[2] - xs[0];                                // derivatives.pyx:78
      d[1] = xs[1] - xs[0];                                // derivatives.pyx:99
      d[2] = xs[2] - xs[1];                                // derivatives.pyx:106
      d[3] = (d[1] + d[2]) / 2.0;                          // derivatives.pyx:111
    } else {
      const double h = xs[1] - xs[0];                      // only the first spacing, derivatives.pyx:89,120
      d[0] = 4.0 * h;
      d[1] = 2.0 * (h * h);
    }
    fast = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      rd[k] = 1.0 / d[k];
      fast = fast & sane(d[k], 120) & (d[k] != 0.0);
    }
  }
  // from the per-cell table written by geometry_kernel: one 80-byte record per cell, tab[i * 10 + k], k = d0..d3,
  // rd0..rd3, inv, fast - five 16-byte loads from one address (the table must be 16-byte aligned)
  __device__ __forceinline__ Geom(const double *__restrict__ tab, int64_t, int64_t i) {
    const double2 *__restrict__ rec = reinterpret_cast<const double2 *>(tab + i * 10);
    const double2 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2), e = __ldg(rec + 3), f = __ldg(rec + 4);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
    rd[0] = c.x; rd[1] = c.y; rd[2] = e.x; rd[3] = e.y;
    inv = f.x;
    fast = f.y != 0.0;
  }
  __device__ __forceinline__ void store(double *__restrict__ tab, int64_t, int64_t i) const {
    double2 *__restrict__ rec = reinterpret_cast<double2 *>(tab + i * 10);
    rec[0] = make_double2(d[0], d[1]);
    rec[1] = make_double2(d[2], d[3]);
    rec[2] = make_double2(rd[0], rd[1]);
    rec[3] = make_double2(rd[2], rd[3]);
    rec[4] = make_double2(inv, fast ? 1.0 : 0.0);
  }
  template <int NV, int NP>
  __device__ __forceinline__ void and_inputs(const double (&u)[W][NV], const double (&prm)[NP]) {
#pragma unroll
    for (int s = 0; s < W; ++s)
#pragma unroll
      for (int j = 0; j < NV; ++j) fast = fast & sane(u[s][j], 120);
#pragma unroll
    for (int k = 0; k < NP; ++k) fast = fast & sane(prm[k], 120);
  }
};

// FAST: this thread's inputs passed the sanity check, quotients use div_known's FMAs (else the IEEE division)
//
// cell_combine: the stencil arithmetic on flux values Fv[s] = F(u[s]), Dv[s] = D(u[s]) already evaluated per slot
// (sigma_c = spectral radius at the centre slot, Lax-Friedrichs only).  cell_residual evaluates them itself; the
// warp-cooperative kernel evaluates each cell's fluxes once and passes them between lanes by shuffle.
template <class Model, int SCHEME, int NV, int W, bool FAST>
__device__ __forceinline__ void cell_combine(const double *prm, const double (&u)[W][NV], const double (&Fv)[W][NV],
                                             const double (&Dv)[W][NV], double sigma_c, const Geom<SCHEME, W> &g,
                                             double (&r)[NV]) {
  if constexpr (SCHEME == SFXM_FV_LAX_FRIEDRICHS || SCHEME == SFXM_FV_CENTRAL || SCHEME == SFXM_FV_UPWIND) {
    static_assert(W == 3 || SCHEME > SFXM_FV_UPWIND, "FV schemes use a 3-cell stencil");
    double Fw[NV], Fe[NV];
    if constexpr (SCHEME == SFXM_FV_LAX_FRIEDRICHS) {
      const double hs = 0.5 * sigma_c;                      // flux.pyx:98-99, "0.5 * sigma * u_diff"
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        Fw[j] = 0.5 * (Fv[0][j] + Fv[1][j]) - hs * (u[1][j] - u[0][j]);   // flux.pyx:154-160
        Fe[j] = 0.5 * (Fv[1][j] + Fv[2][j]) - hs * (u[2][j] - u[1][j]);   // flux.pyx:163-167
      }
    } else if constexpr (SCHEME == SFXM_FV_CENTRAL) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        Fw[j] = 0.5 * (Fv[0][j] + Fv[1][j]);                // flux.pyx:253-255
        Fe[j] = 0.5 * (Fv[1][j] + Fv[2][j]);
      }
    } else {                                                // UPWIND, diagonal dFdU only (R = I)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const bool pos = Model::eig(prm, u[1], j) > 0.0;    // flux.pyx:209-218
        Fw[j] = pos ? Fv[0][j] : Fv[1][j];
        Fe[j] = pos ? Fv[1][j] : Fv[2][j];
      }
    }
    if constexpr (Model::kHasDiffusion) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double DFw = div_known<FAST>(Dv[1][j] - Dv[0][j], g.d[0], g.rd[0]);      // flux.pyx:288
        const double DFe = div_known<FAST>(Dv[2][j] - Dv[1][j], g.d[1], g.rd[1]);      // flux.pyx:297
        r[j] = (0.0 - g.inv * (Fe[j] - Fw[j])) + g.inv * (DFe - DFw);                  // fv_transport.py:58
      }
    } else {
      // D == 0: (Dc-Dl)/hw = +0, DFe-DFw = +0, inv*0 = 0, and a + 0 = a
#pragma unroll
      for (int j = 0; j < NV; ++j) r[j] = (0.0 - g.inv * (Fe[j] - Fw[j])) + 0.0;
    }
  } else if constexpr (SCHEME == SFXM_FD_CENTRAL) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double dF = div_known<FAST>(Fv[2][j] - Fv[0][j], g.d[0], g.rd[0]);         // derivatives.pyx:81
      double d2 = 0.0;
      if constexpr (Model::kHasDiffusion) {
        const double Dw = div_known<FAST>(Dv[1][j] - Dv[0][j], g.d[1], g.rd[1]);       // derivatives.pyx:102
        const double De = div_known<FAST>(Dv[2][j] - Dv[1][j], g.d[2], g.rd[2]);       // derivatives.pyx:109
        d2 = div_known<FAST>(De - Dw, g.d[3], g.rd[3]);                                // derivatives.pyx:112
      }
      r[j] = (0.0 - dF) + d2;                               // fd_transport.py:46-50
    }
  } else {                                                  // FD RIGHT_BIAS, 4 slots
    static_assert(W == 4 || SCHEME != SFXM_FD_RIGHT_BIAS, "RIGHT_BIAS uses a 4-cell stencil");
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double dF = div_known<FAST>(((-Fv[0][j] - Fv[1][j]) + Fv[2][j]) + Fv[W - 1][j], g.d[0], g.rd[0]);   // derivatives.pyx:88-90
      double d2 = 0.0;
      if constexpr (Model::kHasDiffusion)
        d2 = div_known<FAST>(((Dv[0][j] - Dv[1][j]) - Dv[2][j]) + Dv[W - 1][j], g.d[1], g.rd[1]);               // :119-121
      r[j] = (0.0 - dF) + d2;
    }
  }
}

template <class Model, int SCHEME, int NV, int W, bool FAST>
__device__ __forceinline__ void cell_residual(const double *prm, const double (&u)[W][NV], const Geom<SCHEME, W> &g,
                                              double (&r)[NV]) {
  // fluxes per slot (slots a scheme does not use are dead code: FD CENTRAL never reads Fv[1])
  double Fv[W][NV], Dv[W][NV];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    Model::F(prm, u[s], Fv[s]);
    if constexpr (Model::kHasDiffusion) {
      Model::D(prm, u[s], Dv[s]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) Dv[s][j] = 0.0;
    }
  }
  double sigma_c = 0.0;
  if constexpr (SCHEME == SFXM_FV_LAX_FRIEDRICHS) sigma_c = Model::sigma(prm, u[1]);
  cell_combine<Model, SCHEME, NV, W, FAST>(prm, u, Fv, Dv, sigma_c, g, r);
}

}  // namespace sfxm
