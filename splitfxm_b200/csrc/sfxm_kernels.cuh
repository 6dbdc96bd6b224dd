// sfxm_kernels.cuh - residual / ghost fill / finite-difference Jacobian kernel templates (sm_100a).
// Instantiated per model in sfxm_inst_*.cu (separate translation units so they compile in parallel);
// the C ABI and the model dispatch are in sfxm_dispatch.cu.
//
// Thread mapping (all kernels): threadIdx.x walks ensemble members (fastest-varying in memory, so a
// warp's accesses are one contiguous line per (cell, variable) row), threadIdx.y walks cells.  With
// batch == 1 the block is (1, 256): consecutive threads own consecutive cells of the cell-major vector.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sfxm_device.cuh"
#include "sfxm_host.h"
#include "sfxm_cpasync.cuh"

#ifndef SFXM_PAIR_MIN_CTAS
#define SFXM_PAIR_MIN_CTAS 16       // Jacobian pair kernel: 128 registers (20 / 24 CTAs spill and are slower; re-measured on the
                                    // round-2 kernel, bench shape: 12 CTAs 0.387 ms, 16 0.339, 20 0.357, 24 0.391)
#endif
#ifndef SFXM_RES_PAIR_MIN_CTAS
#define SFXM_RES_PAIR_MIN_CTAS 16   // residual pair kernel: 128 registers (24 CTAs spill: 0.39 ms; 12: 0.28; 16: 0.26)
#endif

namespace sfxm {

// (member, cell) of this thread.  Ensembles (blockDim.x == 32): blockIdx.x walks member tiles, so the CTAs that are
// resident together write long contiguous pieces of the same (cell, variable) rows; blockIdx.y/z walk cell tiles.
// Small batches (blockDim.x == 1): blockIdx.x walks cell tiles, blockIdx.y members.
__device__ __forceinline__ void thread_coords(int64_t &m, int64_t &i) {
  if (blockDim.x == 32) {
    m = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (blockDim.y == 1) i = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;   // one warp per CTA: the cell index is warp-uniform
    else i = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * blockDim.y + threadIdx.y;
  } else {
    m = blockIdx.y;
    i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  }
}

// ------------------------------------------------------------------------------------------------
// residual (+ optional fused explicit-Euler update)
// ------------------------------------------------------------------------------------------------
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256)
residual_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                double *__restrict__ r_out, double *__restrict__ y_out, double dt,
                const double *__restrict__ geom_tab, int strip) {
  int64_t m, i0;
  thread_coords(m, i0);
  i0 *= strip;
  if (m >= P.batch) return;
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  View v{y, x_int, m, -1, 0, 0.0};
  // strip > 1 only for one-warp CTAs (ensembles): the warp walks `strip` consecutive cells, so parameters and
  // address bases are set up once and the neighbours' rows are still in L1 when the next cell needs them
  for (int64_t i = i0; i < i0 + strip && i < P.nx; ++i) {
  double u[W][NV], xs[W], r[NV];
  int64_t off[NV];
  gather_stencil<NV, W>(P, v, i, geom_tab == nullptr, u, xs, off);
  // geometry depends on the mesh only: with many members per cell it comes from a table (sfxm_geometry)
  Geom<SCHEME, W> geom = geom_tab ? Geom<SCHEME, W>(geom_tab, P.nx, i) : Geom<SCHEME, W>(xs);
  geom.and_inputs(u, prm);
  if (geom.fast) cell_residual<Model, SCHEME, NV, W, true>(prm, u, geom, r);
  else cell_residual<Model, SCHEME, NV, W, false>(prm, u, geom, r);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (r_out) r_out[off[j]] = r[j];
    // the cell's own value is slot nbl (selected, not indexed: a dynamic index would push u[][] to local memory)
    const double uc = (W == 3 || P.nbl == 1) ? u[1][j] : u[W == 3 ? 1 : 2][j];
    if (y_out) { const double yn = uc + dt * r[j]; y_out[off[j]] = yn; push_halo(P, i, j, m, yn); }              // sol.y += delta_t * f, simulation.py:459
  }
  }
}

// Ensemble twin of residual_kernel: one warp per CTA walking `strip` cells, two adjacent members per lane
// (16-byte loads/stores), warp-uniform cell index.  Boundary cells and insane inputs fall back per member.
template <class Model, int SCHEME, int NV, int W>
__device__ __noinline__ void residual_member_noinline(const sfxm_problem_t &P, const double *__restrict__ x_int,
                                                      const double *__restrict__ params, int64_t param_stride,
                                                      const double *__restrict__ y, double *__restrict__ r_out,
                                                      double *__restrict__ y_out, double dt,
                                                      const double *__restrict__ geom_tab, int64_t i, int64_t m) {
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][NV], xs[W], r[NV];
  int64_t off[NV];
  gather_stencil<NV, W>(P, v, i, geom_tab == nullptr, u, xs, off);
  Geom<SCHEME, W> geom = geom_tab ? Geom<SCHEME, W>(geom_tab, P.nx, i) : Geom<SCHEME, W>(xs);
  geom.and_inputs(u, prm);
  if (geom.fast) cell_residual<Model, SCHEME, NV, W, true>(prm, u, geom, r);
  else cell_residual<Model, SCHEME, NV, W, false>(prm, u, geom, r);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (r_out) r_out[off[j]] = r[j];
    const double uc = (W == 3 || P.nbl == 1) ? u[1][j] : u[W == 3 ? 1 : 2][j];
    if (y_out) { const double yn = uc + dt * r[j]; y_out[off[j]] = yn; push_halo(P, i, j, m, yn); }
  }
}

// UPDATE: also write y_out = y + dt f(y) (and feed the neighbours' halos); a separate instantiation keeps the
// plain residual free of the update's registers
template <class Model, int SCHEME, int NV, int W, bool UPDATE>
__global__ void __launch_bounds__(32, SFXM_RES_PAIR_MIN_CTAS)
residual_pair_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                     const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                     double *__restrict__ r_out, double *__restrict__ y_out, double dt,
                     const double *__restrict__ geom_tab, int strip) {
  const int64_t i0 = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * strip;
  const int64_t m0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 2;
  if (m0 >= P.batch) return;
  const int64_t B = P.batch;
  const int nbl = (W == 3) ? 1 : P.nbl;
  double prm[2][Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) {
    if (param_stride) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(params + k * param_stride + m0));
      prm[0][k] = t.x;
      prm[1][k] = t.y;
    } else {
      prm[0][k] = prm[1][k] = __ldg(params + k);
    }
  }
  bool prm_ok = true;
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm_ok = prm_ok & sane(prm[0][k], 120) & sane(prm[1][k], 120);
  for (int64_t i = i0; i < i0 + strip && i < P.nx; ++i) {
    const bool interior = (i >= nbl) && (i + (W - 1 - nbl) < P.nx);
    bool ok = interior && prm_ok;
    double u[2][W][NV];
    int64_t off[NV];
    if (ok) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t rs = (int64_t)P.var_stride[j] * B;
        off[j] = (P.var_base[j] + P.var_off[j]) * B + m0 + i * rs;
        const double *q = y + (off[j] - nbl * rs);
#pragma unroll
        for (int s = 0; s < W; ++s) {
          const double2 t = __ldg(reinterpret_cast<const double2 *>(q));
          u[0][s][j] = t.x;
          u[1][s][j] = t.y;
          ok = ok & sane(t.x, 120) & sane(t.y, 120);
          q += rs;
        }
      }
    }
    const Geom<SCHEME, W> geom(geom_tab, P.nx, i);
    if (!(ok && geom.fast)) {
      residual_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, r_out, y_out, dt, geom_tab, i, m0);
      residual_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, r_out, y_out, dt, geom_tab, i, m0 + 1);
      continue;
    }
    double r[2][NV];
#pragma unroll
    for (int v = 0; v < 2; ++v) cell_residual<Model, SCHEME, NV, W, true>(prm[v], u[v], geom, r[v]);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (r_out) *reinterpret_cast<double2 *>(r_out + off[j]) = make_double2(r[0][j], r[1][j]);
      if constexpr (UPDATE) {
        const double c0 = (W == 3 || P.nbl == 1) ? u[0][1][j] : u[0][W == 3 ? 1 : 2][j];
        const double c1 = (W == 3 || P.nbl == 1) ? u[1][1][j] : u[1][W == 3 ? 1 : 2][j];
        const double2 yn = make_double2(c0 + dt * r[0][j], c1 + dt * r[1][j]);                                // simulation.py:459
        *reinterpret_cast<double2 *>(y_out + off[j]) = yn;
        push_halo(P, i, j, m0, yn.x);
        push_halo(P, i, j, m0 + 1, yn.y);
      }
    }
  }
}

// 3-cell stencils: the same kernel with the state streamed ONCE.  residual_pair_kernel gathers the three cells of every
// stencil (two of them L1 hits) and can only start a cell's loads when the previous cell is done, so a one-warp CTA has
// one cell's rows in flight.  Here the warp keeps the stencil as a sliding register window; the rows of the cell that
// enters the window are copied cp.async into a per-lane ring in shared memory kRingCells cells ahead of their use.
#ifndef SFXM_RES_RING_CELLS
#define SFXM_RES_RING_CELLS 4
#endif
template <class Model, int SCHEME, int NV, bool UPDATE>
__global__ void __launch_bounds__(32, SFXM_RES_PAIR_MIN_CTAS)
residual_ring_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                     const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                     double *__restrict__ r_out, double *__restrict__ y_out, double dt,
                     const double *__restrict__ geom_tab, int strip) {
  constexpr int W = 3, D = SFXM_RES_RING_CELLS;
  __shared__ double2 ring[D][NV][32];
  const int64_t i0 = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * strip;
  const int64_t m0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 2;
  if (m0 >= P.batch || i0 >= P.nx) return;
  const int64_t B = P.batch, nx = P.nx;
  const int lane = threadIdx.x;
  const int64_t i1 = i0 + strip < nx ? i0 + strip : nx;
  double prm[2][Model::kParams];
  bool prm_ok = true;
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) {
    if (param_stride) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(params + k * param_stride + m0));
      prm[0][k] = t.x;
      prm[1][k] = t.y;
    } else {
      prm[0][k] = prm[1][k] = __ldg(params + k);
    }
    prm_ok = prm_ok & sane(prm[0][k], 120) & sane(prm[1][k], 120);
  }
  int64_t base[NV], rs[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    rs[j] = (int64_t)P.var_stride[j] * B;
    base[j] = (P.var_base[j] + P.var_off[j]) * B + m0;
  }
  auto issue = [&](int64_t cell) {           // one commit group per cell (empty beyond the mesh)
    if (cell < nx) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        cp_async16(&ring[cell & (D - 1)][j][lane], reinterpret_cast<const double2 *>(y + base[j] + cell * rs[j]));
    }
    cp_async_commit();
  };
  // window: slot 1 = cell i - 1, slot 2 = cell i at loop entry (shifted down at the top of every iteration)
  double u[2][W][NV];
  bool ok_w[W];
  {
    const int64_t cm = i0 > 0 ? i0 - 1 : 0;
    ok_w[0] = ok_w[1] = ok_w[2] = true;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double2 a = __ldg(reinterpret_cast<const double2 *>(y + base[j] + cm * rs[j]));
      const double2 b = __ldg(reinterpret_cast<const double2 *>(y + base[j] + i0 * rs[j]));
      u[0][1][j] = a.x; u[1][1][j] = a.y;
      u[0][2][j] = b.x; u[1][2][j] = b.y;
      u[0][0][j] = u[1][0][j] = 0.0;
      ok_w[1] = ok_w[1] & sane(a.x, 120) & sane(a.y, 120);
      ok_w[2] = ok_w[2] & sane(b.x, 120) & sane(b.y, 120);
    }
  }
#pragma unroll
  for (int k = 0; k < D; ++k) issue(i0 + 1 + k);
  for (int64_t i = i0; i < i1; ++i) {
    ok_w[0] = ok_w[1];
    ok_w[1] = ok_w[2];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      u[0][0][j] = u[0][1][j]; u[1][0][j] = u[1][1][j];
      u[0][1][j] = u[0][2][j]; u[1][1][j] = u[1][2][j];
    }
    cp_async_wait<D - 1>();                  // the oldest group in flight (cell i + 1) has landed
    if (i + 1 < nx) {
      ok_w[2] = true;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double2 t = ring[(i + 1) & (D - 1)][j][lane];
        u[0][2][j] = t.x; u[1][2][j] = t.y;
        ok_w[2] = ok_w[2] & sane(t.x, 120) & sane(t.y, 120);
      }
    }
    issue(i + 1 + D);                        // refills the slot just read
    const bool interior = (i >= 1) && (i + 1 < nx);
    const Geom<SCHEME, W> geom(geom_tab, nx, i);
    if (!(interior && prm_ok && ok_w[0] && ok_w[1] && ok_w[2] && geom.fast)) {
      residual_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, r_out, y_out, dt, geom_tab, i, m0);
      residual_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, r_out, y_out, dt, geom_tab, i, m0 + 1);
      continue;
    }
    double r[2][NV];
#pragma unroll
    for (int v = 0; v < 2; ++v) cell_residual<Model, SCHEME, NV, W, true>(prm[v], u[v], geom, r[v]);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int64_t off = base[j] + i * rs[j];
      if (r_out) *reinterpret_cast<double2 *>(r_out + off) = make_double2(r[0][j], r[1][j]);
      if constexpr (UPDATE) {
        const double2 yn = make_double2(u[0][1][j] + dt * r[0][j], u[1][1][j] + dt * r[1][j]);               // simulation.py:459
        *reinterpret_cast<double2 *>(y_out + off) = yn;
        push_halo(P, i, j, m0, yn.x);
        push_halo(P, i, j, m0 + 1, yn.y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Coupled models on one mesh (batch < 32): warp-cooperative residual.  Consecutive lanes own consecutive cells; each
// lane loads ITS cell and evaluates F, D (and the spectral radius) ONCE, the stencil neighbours' values and fluxes
// arrive by warp shuffle - a third of the loads and of the flux evaluations (the Euler fluxes are three IEEE
// divisions each) of the one-thread-does-everything kernel.  Lanes [nbl, 32 - nbr) of a warp produce residuals, the
// outer lanes only feed their neighbours.  Warps that touch a domain end use the general per-thread path.
// ------------------------------------------------------------------------------------------------
// resident 256-thread CTAs per SM the compiler must allow (register cap).  Conservative Sod, nx = 1e7, explicit Euler step
// (gpurun_out r2w): uncapped (80 registers, 3 CTAs) 0.420 ms; 4 CTAs (64 registers, 16 bytes spilled) 0.314 ms; 5 (48
// registers) 0.348 ms; 6 (40 registers) 0.412 ms.
#ifndef SFXM_WARP_RES_MIN_CTAS
#define SFXM_WARP_RES_MIN_CTAS 4
#endif
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256, SFXM_WARP_RES_MIN_CTAS)
residual_warp_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                     const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                     double *__restrict__ r_out, double *__restrict__ y_out, double dt, int64_t warp0, int64_t warp_end) {
  // warps [warp0, warp_end) of the mesh's ceil(nx / per_warp): the whole mesh, or a piece of it (sfxm_residual_warps: a
  // slab's interior goes out before its neighbours' halos have arrived, the two end warps after)
  const int64_t m = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int nbl = (W == 3) ? 1 : P.nbl;
  const int nbr = W - 1 - nbl;
  const int per_warp = 32 - (W - 1);
  // warp0 < 0: the two END warps of the mesh in one launch (block 0: warp 0, block 1: warp warp_end - 1; one warp each)
  if (warp0 < 0 && (threadIdx.x >> 5) != 0) return;
  const int64_t warp = warp0 < 0 ? (blockIdx.x == 0 ? 0 : warp_end - 1)
                                 : warp0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t first = warp * per_warp;                  // first cell this warp produces
  if (warp >= warp_end || first >= P.nx) return;
  const int64_t i = first - nbl + lane;                   // this lane's cell
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  const bool warp_interior = (first - nbl >= 0) && (first - nbl + 31 < P.nx);
  if (!warp_interior) {
    // general path: every producing lane gathers its own stencil (ghost cells, mesh ends)
    if (lane < nbl || lane >= 32 - nbr || i >= P.nx) return;
    residual_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, r_out, y_out, dt, nullptr, i, m);
    return;
  }
  const int64_t B = P.batch;
  int64_t off[NV];
  double uc[NV], Fc[NV], Dc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    // (32-bit row lengths: one multiply-add per address; nx * nv * batch < 2^32 is checked by the launcher)
    off[j] = (P.var_base[j] + P.var_off[j]) * B + m + (int64_t)((uint64_t)(uint32_t)i * (uint32_t)(P.var_stride[j] * (int32_t)B));
    uc[j] = __ldg(y + off[j]);
  }
  const double xc = __ldg(x_int + i);
  Model::F(prm, uc, Fc);
  if constexpr (Model::kHasDiffusion) {
    Model::D(prm, uc, Dc);
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) Dc[j] = 0.0;
  }
  double sigma_c = 0.0;
  if constexpr (SCHEME == SFXM_FV_LAX_FRIEDRICHS) sigma_c = Model::sigma(prm, uc);
  // stencil slot s of lane l is lane l - nbl + s; the lane's own slot needs no exchange
  double u[W][NV], Fv[W][NV], Dv[W][NV], xs[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    const int src = lane - nbl + s;
    const bool own = (W == 3) && s == 1;                  // (4-cell stencils: the own slot is nbl, a run-time value)
    xs[s] = own ? xc : __shfl_sync(0xffffffffu, xc, src);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      u[s][j] = own ? uc[j] : __shfl_sync(0xffffffffu, uc[j], src);
      Fv[s][j] = own ? Fc[j] : __shfl_sync(0xffffffffu, Fc[j], src);
      Dv[s][j] = Model::kHasDiffusion ? (own ? Dc[j] : __shfl_sync(0xffffffffu, Dc[j], src)) : 0.0;
    }
  }
  if (lane < nbl || lane >= 32 - nbr) return;             // feeder lanes are done
  Geom<SCHEME, W> geom(xs);
  double r[NV];
  // the range check guards div_known; a finite-volume scheme without diffusion has no quotient by a known divisor at all
  // (fv_transport.py:58 multiplies by 1 / dx), so both instantiations are the same arithmetic
  constexpr bool kNoKnownDivision = !Model::kHasDiffusion && (SCHEME == SFXM_FV_LAX_FRIEDRICHS || SCHEME == SFXM_FV_CENTRAL ||
                                                              SCHEME == SFXM_FV_UPWIND);
  if constexpr (kNoKnownDivision) {
    cell_combine<Model, SCHEME, NV, W, false>(prm, u, Fv, Dv, sigma_c, geom, r);
  } else {
    geom.and_inputs(u, prm);
    if (geom.fast) cell_combine<Model, SCHEME, NV, W, true>(prm, u, Fv, Dv, sigma_c, geom, r);
    else cell_combine<Model, SCHEME, NV, W, false>(prm, u, Fv, Dv, sigma_c, geom, r);
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (r_out) r_out[off[j]] = r[j];
    if (y_out) { const double yn = uc[j] + dt * r[j]; y_out[off[j]] = yn; push_halo(P, i, j, m, yn); }           // sol.y += delta_t * f, simulation.py:459
  }
}

// ------------------------------------------------------------------------------------------------
// per-cell geometry table (divisors, their reciprocals, 1/dx): mesh-only data shared by all ensemble members
// ------------------------------------------------------------------------------------------------
template <int SCHEME, int W>
__global__ void geometry_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                                double *__restrict__ tab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nx) return;
  double xs[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    const int64_t pos = i - P.nbl + s;
    xs[s] = pos < 0 ? ghost_x_left(P, x_int, (int)(-pos - 1))
                    : (pos >= P.nx ? ghost_x_right(P, x_int, (int)(pos - P.nx)) : __ldg(x_int + pos));
  }
  const Geom<SCHEME, W> g(xs);
  g.store(tab, P.nx, i);
}

// ------------------------------------------------------------------------------------------------
// Domain view: interior + ghosts, cell-major unsplit (initialize_from_list + apply_BC)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void fill_domain_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                                   const double *__restrict__ y, double *__restrict__ x_all,
                                   double *__restrict__ U_all) {
  int64_t m, c;   // c: index into the padded mesh
  thread_coords(m, c);
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
// bits of entry_mask that belong to column c (rows r = 0..NV-1)
template <int NV>
__device__ __forceinline__ constexpr uint32_t column_bits(int c) {
  uint32_t b = 0;
  for (int r = 0; r < NV; ++r) b |= 1u << (r * NV + c);
  return b;
}

// Entries of a block that are ever written.  A decoupled model (one scalar equation per variable) has exact zeros off
// the diagonal of every block - the reference computes them and its lil_matrix drops them (simulation.py:354-392) - so
// they are structural: J buffers are zero-filled once by their owner (include/sfxm_b200.h) and these entries are never
// touched.  That is 40 % of the stores of the bench shape, and the store path out of the SMs is what binds the assembly
// kernel (profile r1g: l1tex->xbar write bytes at 74 % of peak with DRAM at 51 %).
template <class Model>
__device__ __forceinline__ constexpr bool stored_entry(int r, int c) {
  return !Model::kDecoupled || r == c;
}

// One perturbed evaluation: column c of the stencil is `upc`, everything else is the base state.
// Returns (r(u + eps e) - r0) / eps for every row (decoupled models: only row c can be non-zero).
template <class Model, int SCHEME, int NV, int W, bool FAST>
__device__ __forceinline__ void jac_column(const double *prm, const double (&u)[W][NV], const double (&upc)[W], int c,
                                           const Geom<SCHEME, W> &geom, const double (&r0)[NV], double eps, double reps,
                                           double (&col)[NV]) {
  if constexpr (Model::kDecoupled) {
    // scalar equation per variable: only row c can change (the other rows are exact zeros in the reference too)
    double us[W][1], r1[1];
#pragma unroll
    for (int s = 0; s < W; ++s) us[s][0] = upc[s];
    cell_residual<typename Model::Scalar, SCHEME, 1, W, FAST>(prm, us, geom, r1);    // simulation.py:345
#pragma unroll
    for (int r = 0; r < NV; ++r) col[r] = 0.0;
    col[c] = div_known<FAST>(r1[0] - r0[c], eps, reps);                              // simulation.py:354
  } else {
    double up[W][NV], r1[NV];
#pragma unroll
    for (int s = 0; s < W; ++s)
#pragma unroll
      for (int j = 0; j < NV; ++j) up[s][j] = (j == c) ? upc[s] : u[s][j];
    cell_residual<Model, SCHEME, NV, W, FAST>(prm, up, geom, r1);                    // simulation.py:345
#pragma unroll
    for (int r = 0; r < NV; ++r) col[r] = div_known<FAST>(r1[r] - r0[r], eps, reps); // simulation.py:354
  }
}

// Block row of one cell.  entry_mask bit (r * NV + c): rows r and columns c of the same split segment
// (simulation.py:373-392).  J entry (d, r, c) of this cell/member is at p + ((d*NV + r)*NV + c) * B.
template <class Model, int SCHEME, int NV, int W, bool FAST, bool RESTRICT>
__device__ __forceinline__ void jacobian_cell(const sfxm_problem_t &P, const double *__restrict__ x_int,
                                              const double *__restrict__ y, const double *prm, int64_t i, int64_t m,
                                              bool interior, const int64_t (&off)[NV], const double (&u)[W][NV],
                                              const Geom<SCHEME, W> &geom, double eps,
                                              double *__restrict__ p, double *__restrict__ r0_out, uint32_t entry_mask,
                                              const int64_t B, double (&r0)[NV]) {
  // B: distance between consecutive entries at p (the batch size for the banded layout in HBM; 1 when the block row
  // goes to a thread-local array, sfxm_fused.cuh).  r0: the unperturbed residual, also returned to the caller.
  const double reps = 1.0 / eps;
  cell_residual<Model, SCHEME, NV, W, FAST>(prm, u, geom, r0);   // unperturbed rhs, simulation.py:320
  if (r0_out) {
#pragma unroll
    for (int j = 0; j < NV; ++j) r0_out[off[j]] = r0[j];
  }
  // interior == false: the stencil contains ghost cells (then BCs are re-applied after every perturbation,
  // simulation.py:340-342, and wrap columns exist for periodic variables, bc.py:195-240)
  if (interior) {
#pragma unroll
    for (int d = 0; d < W; ++d) {
      double cols[NV][NV];   // cols[c][r]
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if constexpr (RESTRICT) {
          if (!(entry_mask & column_bits<NV>(c))) continue;   // variable outside the requested segments: nothing stored
        }
        double upc[W];
#pragma unroll
        for (int s = 0; s < W; ++s) upc[s] = u[s][c];
        upc[d] = u[d][c] + eps;                             // simulation.py:337
        jac_column<Model, SCHEME, NV, W, FAST>(prm, u, upc, c, geom, r0, eps, reps, cols[c]);
      }
#pragma unroll
      for (int r = 0; r < NV; ++r)
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          if (stored_entry<Model>(r, c) && (entry_mask & (1u << (r * NV + c)))) *p = cols[c][r];
          p += B;
        }
    }
    return;
  }
#pragma unroll 1
  for (int d = 0; d < W; ++d) {
    const int64_t pos = i - P.nbl + d;
    double cols[NV][NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      // column cell: in-band interior cell, or the periodic image (extend_band, bc.py:195-240)
      // (a HALO ghost belongs to the neighbouring slab: its own value is the unknown that is perturbed, and the
      // slot holds the coupling block to that slab)
      int64_t pc = pos;
      bool live = true, halo = false;
      if (pos < 0) {
        halo = (P.bc_ltype[c] == SFXM_BC_HALO);
        live = halo || (P.bc_ltype[c] == SFXM_BC_PERIODIC);
        pc = halo ? -1 : pos + P.nx;
      } else if (pos >= P.nx) {
        halo = (P.bc_rtype[c] == SFXM_BC_HALO);
        live = halo || (P.bc_rtype[c] == SFXM_BC_PERIODIC);
        pc = halo ? -1 : pos - P.nx;
      }
#pragma unroll
      for (int r = 0; r < NV; ++r) cols[c][r] = 0.0;
      if (live && (entry_mask & column_bits<NV>(c))) {
        View vp{y, x_int, m, pc, c, eps};
        double upc[W];
#pragma unroll
        for (int s = 0; s < W; ++s) {
          const int64_t ps = i - P.nbl + s;
          if (ps < 0) upc[s] = ghost_u_left(P, vp, (int)(-ps - 1), c);
          else if (ps >= P.nx) upc[s] = ghost_u_right(P, vp, (int)(ps - P.nx), c);
          else upc[s] = (ps == pc) ? u[s][c] + eps : u[s][c];
          if (halo && ps == pos) upc[s] = upc[s] + eps;
        }
        jac_column<Model, SCHEME, NV, W, FAST>(prm, u, upc, c, geom, r0, eps, reps, cols[c]);
      }
    }
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (stored_entry<Model>(r, c) && (entry_mask & (1u << (r * NV + c)))) *p = cols[c][r];
        p += B;
      }
  }
}

// Everything one (cell, member) does for the Jacobian: gather, sanity check, block row (general path).
template <class Model, int SCHEME, int NV, int W, bool RESTRICT>
__device__ __forceinline__ void jacobian_member(const sfxm_problem_t &P, const double *__restrict__ x_int,
                                                const double *__restrict__ params, int64_t param_stride,
                                                const double *__restrict__ y, double eps, double *__restrict__ J,
                                                double *__restrict__ r0_out, uint32_t entry_mask,
                                                const double *__restrict__ geom_tab, int64_t i, int64_t m) {
  double prm[Model::kParams];
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][NV], xs[W];
  int64_t off[NV];
  const bool interior = gather_stencil<NV, W>(P, v, i, geom_tab == nullptr, u, xs, off);
  Geom<SCHEME, W> geom = geom_tab ? Geom<SCHEME, W>(geom_tab, P.nx, i) : Geom<SCHEME, W>(xs);
  geom.and_inputs(u, prm);
  geom.fast = geom.fast && sane(eps, 120);
  double *p = J + (i * W) * (int64_t)(NV * NV) * P.batch + m;
  double r0[NV];
  if (geom.fast)
    jacobian_cell<Model, SCHEME, NV, W, true, RESTRICT>(P, x_int, y, prm, i, m, interior, off, u, geom, eps, p, r0_out, entry_mask, P.batch, r0);
  else
    jacobian_cell<Model, SCHEME, NV, W, false, RESTRICT>(P, x_int, y, prm, i, m, interior, off, u, geom, eps, p, r0_out, entry_mask, P.batch, r0);
}

// out-of-line copy for the rare paths of the pair kernel (keeps its straight-line path free of their registers)
template <class Model, int SCHEME, int NV, int W>
__device__ __noinline__ void jacobian_member_noinline(const sfxm_problem_t &P, const double *__restrict__ x_int,
                                                      const double *__restrict__ params, int64_t param_stride,
                                                      const double *__restrict__ y, double eps, double *__restrict__ J,
                                                      double *__restrict__ r0_out, uint32_t entry_mask,
                                                      const double *__restrict__ geom_tab, int64_t i, int64_t m) {
  jacobian_member<Model, SCHEME, NV, W, true>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m);
}

template <class Model, int SCHEME, int NV, int W, bool RESTRICT>
__global__ void __launch_bounds__(256)
jacobian_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                double eps, double *__restrict__ J, double *__restrict__ r0_out, uint32_t entry_mask,
                const double *__restrict__ geom_tab) {
  // (no strip loop here: wrapping this body in a loop costs registers - 128 + spills - and measured slower)
  int64_t m, i;
  thread_coords(m, i);
  if (m >= P.batch || i >= P.nx) return;
  jacobian_member<Model, SCHEME, NV, W, RESTRICT>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m);
}

// ------------------------------------------------------------------------------------------------
// Ensembles: one warp per CTA, the cell index comes from blockIdx (warp-uniform address arithmetic) and every
// lane owns TWO adjacent members: 16-byte loads/stores halve the memory instructions and their 64-bit address
// arithmetic per member, and the two members' division chains interleave (the r1f profile is latency-bound on
// those chains at 21 % occupancy).  Needs an even batch; boundary cells and threads whose inputs fail the
// sanity check take the general path member by member.
// ------------------------------------------------------------------------------------------------
// RESTRICT: only the columns of the requested segments are evaluated (sfxm_jacobian_segments).  A separate
// instantiation, because the per-column branch ends the straight-line code in which the compiler interleaves the
// independent column evaluations (measured: 0.57 -> 0.70 ms when the unrestricted kernel carried the test).
__device__ __forceinline__ void prefetch_l1_row(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Host-prepared arguments (round 2, second half): everything that is the same for every CTA is formed once on the
// host - row bases with the split permutation folded in, 32-bit row lengths - so that a CTA's address arithmetic is a
// handful of 32 x 32 -> 64-bit multiply-adds instead of ~25 64-bit products of sfxm_problem_t fields (SASS of the
// previous kernel: 177 IMAD + 48 LEA + 41 IADD3 of 1199 instructions per warp, more than a third of its non-FP64 work).
template <int NV>
struct PairArgs {
  const double *yv[NV];   // y + (var_base[j] + var_off[j]) * batch: variable j of cell 0, member 0
  double *r0v[NV];        // the same rows of r0_out (null: not wanted)
  uint32_t row[NV];       // var_stride[j] * batch: elements from one cell to the next
  uint32_t batch, nx, nbl;
  uint32_t pstride;       // parameter k of member m at params[k * pstride + m] (0: shared by all members)
  uint32_t eps_sane;      // the finite-difference step passes the range check of div_known (sfxm_device.cuh)
  double reps;            // RN(1 / eps), formed on the host (IEEE division: the same bits as on the device)
};

#ifndef SFXM_PAIR_CELLS
#define SFXM_PAIR_CELLS 1           // consecutive cells per one-warp CTA of the ensemble Jacobian kernel (2 / 4 cells with the
                                    // next cell's new row prefetched: 0.343 / 0.339 ms against 0.341 - the load phase is not what
                                    // binds the kernel; kept at 1)
#endif

template <class Model, int SCHEME, int NV, int W, bool RESTRICT>
__device__ __forceinline__ void jacobian_pair_cell(const sfxm_problem_t &P, const PairArgs<NV> &A,
                     const double *__restrict__ x_int, const double *__restrict__ params, int64_t param_stride,
                     const double *__restrict__ y, double eps, double *__restrict__ J, double *__restrict__ r0_out,
                     uint32_t entry_mask, const double *__restrict__ geom_tab, const uint32_t i, const uint32_t m0) {
  const uint32_t nbl = (W == 3) ? 1u : A.nbl;
  const bool interior = (i >= nbl) && (i + (W - 1 - nbl) < A.nx);
  if (!interior) {   // warp-uniform: boundary cells take the general path
    jacobian_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0);
    jacobian_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0 + 1);
    return;
  }
  const uint32_t B = A.batch;
  double prm[2][Model::kParams];
  double u[2][W][NV];
  const Geom<SCHEME, W> geom(geom_tab, A.nx, i);
#pragma unroll
  for (int k = 0; k < Model::kParams; ++k) {
    const double2 t = A.pstride ? __ldg(reinterpret_cast<const double2 *>(params + ((uint64_t)((uint32_t)k * A.pstride) + m0)))
                                : make_double2(__ldg(params + k), __ldg(params + k));
    prm[0][k] = t.x;
    prm[1][k] = t.y;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double *q = A.yv[j] + ((uint64_t)(i - nbl) * A.row[j] + m0);
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(q + (uint64_t)((uint32_t)s * A.row[j])));
      u[0][s][j] = t.x;
      u[1][s][j] = t.y;
    }
  }
  bool vector_ok = geom.fast & (A.eps_sane != 0u);
#pragma unroll
  for (int v = 0; v < 2; ++v) {
#pragma unroll
    for (int s = 0; s < W; ++s)
#pragma unroll
      for (int j = 0; j < NV; ++j) vector_ok = vector_ok & sane(u[v][s][j], 120);
#pragma unroll
    for (int k = 0; k < Model::kParams; ++k) vector_ok = vector_ok & sane(prm[v][k], 120);
  }
  if (!vector_ok) {
    jacobian_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0);
    jacobian_member_noinline<Model, SCHEME, NV, W>(P, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0 + 1);
    return;
  }
  const double reps = A.reps;
  double *const Jc = J + ((uint64_t)i * ((uint32_t)(W * NV * NV) * B) + m0);   // entry (d, r, c) at Jc + ((d NV + r) NV + c) B
  if constexpr (Model::kDecoupled) {
    // one scalar equation per variable (only row c of column c can be non-zero, see stored_entry): variable by variable,
    // so that a variable's stencil values and residual die as soon as its W entries are stored
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if constexpr (RESTRICT) {
        if (!(entry_mask & column_bits<NV>(c))) continue;   // variable outside the requested segments
      }
      double us[2][W][1], r0[2][1];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int s = 0; s < W; ++s) us[v][s][0] = u[v][s][c];
        cell_residual<typename Model::Scalar, SCHEME, 1, W, true>(prm[v], us[v], geom, r0[v]);   // simulation.py:320
      }
      if (A.r0v[c])
        *reinterpret_cast<double2 *>(A.r0v[c] + ((uint64_t)i * A.row[c] + m0)) = make_double2(r0[0][0], r0[1][0]);
#pragma unroll
      for (int d = 0; d < W; ++d) {
        double e[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          double up[W][1], r1[1];
#pragma unroll
          for (int s = 0; s < W; ++s) up[s][0] = us[v][s][0];
          up[d][0] = us[v][d][0] + eps;                                                          // simulation.py:337
          cell_residual<typename Model::Scalar, SCHEME, 1, W, true>(prm[v], up, geom, r1);       // simulation.py:345
          e[v] = div_known<true>(r1[0] - r0[v][0], eps, reps);                                   // simulation.py:354
        }
        if (entry_mask & (1u << (c * NV + c)))
          *reinterpret_cast<double2 *>(Jc + (uint64_t)((uint32_t)((d * NV + c) * NV + c) * B)) = make_double2(e[0], e[1]);
      }
    }
  } else {
  double r0[2][NV];
#pragma unroll
  for (int v = 0; v < 2; ++v) cell_residual<Model, SCHEME, NV, W, true>(prm[v], u[v], geom, r0[v]);   // simulation.py:320
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (A.r0v[j])
      *reinterpret_cast<double2 *>(A.r0v[j] + ((uint64_t)i * A.row[j] + m0)) = make_double2(r0[0][j], r0[1][j]);
#pragma unroll
  for (int d = 0; d < W; ++d) {
    double cols[2][NV][NV];   // [member][c][r]
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if constexpr (RESTRICT) {
          if (!(entry_mask & column_bits<NV>(c))) continue;   // variable outside the requested segments
        }
        double upc[W];
#pragma unroll
        for (int s = 0; s < W; ++s) upc[s] = u[v][s][c];
        upc[d] = u[v][d][c] + eps;                          // simulation.py:337
        jac_column<Model, SCHEME, NV, W, true>(prm[v], u[v], upc, c, geom, r0[v], eps, reps, cols[v][c]);
      }
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if (stored_entry<Model>(r, c) && (entry_mask & (1u << (r * NV + c))))
          *reinterpret_cast<double2 *>(Jc + (uint64_t)((uint32_t)((d * NV + r) * NV + c) * B)) =
              make_double2(cols[0][c][r], cols[1][c][r]);
  }
  }
}

template <class Model, int SCHEME, int NV, int W, bool RESTRICT>
__global__ void __launch_bounds__(32, SFXM_PAIR_MIN_CTAS)
jacobian_pair_kernel(const __grid_constant__ sfxm_problem_t P, const __grid_constant__ PairArgs<NV> A,
                     const double *__restrict__ x_int, const double *__restrict__ params, int64_t param_stride,
                     const double *__restrict__ y, double eps, double *__restrict__ J, double *__restrict__ r0_out,
                     uint32_t entry_mask, const double *__restrict__ geom_tab) {
  const uint32_t m0 = (blockIdx.x * 32 + threadIdx.x) * 2;
  if (m0 >= A.batch) return;
  if constexpr (SFXM_PAIR_CELLS == 1) {
    const uint32_t i = blockIdx.z * gridDim.y + blockIdx.y;
    if (i >= A.nx) return;
    jacobian_pair_cell<Model, SCHEME, NV, W, RESTRICT>(P, A, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0);
  } else {
    // consecutive cells share W - 1 of their W stencil rows (L1 hits for the later cell); the one new row of the next
    // cell is requested before this cell's arithmetic starts
    const uint32_t i0 = (blockIdx.z * gridDim.y + blockIdx.y) * SFXM_PAIR_CELLS;
#pragma unroll 1
    for (uint32_t i = i0; i < i0 + SFXM_PAIR_CELLS && i < A.nx; ++i) {
      const uint32_t nbl = (W == 3) ? 1u : A.nbl;
      const uint32_t inext = i + 1 - nbl + (W - 1);
      if (i + 1 < i0 + SFXM_PAIR_CELLS && inext < A.nx) {
#pragma unroll
        for (int j = 0; j < NV; ++j) prefetch_l1_row(A.yv[j] + ((uint64_t)inext * A.row[j] + m0));
      }
      jacobian_pair_cell<Model, SCHEME, NV, W, RESTRICT>(P, A, x_int, params, param_stride, y, eps, J, r0_out, entry_mask, geom_tab, i, m0);
    }
  }
}

// (Round 2: TMA bulk stores - the W * NV stored 512-byte rows of a 64-member tile staged in shared memory and handed to
// the copy engine by one lane, cp.async.bulk.global.shared::cta / SASS UBLKCP, instead of 16-byte STG per lane, because
// ncu r1j showed the l1tex -> xbar store path as the busiest unit after the FP64 pipe: 0.543 ms against 0.506 ms on the
// bench shape, both of two runs (gpurun_out r2q) - the staging, the proxy fence and the wait before the one-cell CTA
// exits cost more than the store path gives back; removed.)
// (The ring / sliding-window structure of residual_ring_kernel was tried here too - one warp walking 4-32 cells, the
// entering cell's rows prefetched by cp.async: the loop around the unrolled perturbations spills 500-800 bytes at 128
// registers (276 at 168) and runs at 0.73-0.77 ms against 0.50 ms for one cell per CTA; removed.  The assembly kernel is
// bound by its FP64 instruction stream, not by load latency.)
// (A per-variable variant for decoupled models - one warp per (cell, variable), two members per lane, 64-96 registers,
// 20-32 one-warp CTAs per SM instead of 16 - was tried to get more resident warps: bit-identical, but 0.63-0.67 ms
// against 0.57 ms on the bench shape at every register setting (the per-warp set-up - geometry row, parameters,
// addresses, sanity check - is paid three times and more warps did not buy overlap); removed.)
// (A strip variant - the warp walks several cells with the stencil as a sliding register window and prefetches the
// next cell's slot before evaluating the current one - was tried to overlap the load, compute and store phases: the
// loop body needs > 255 registers, spills, and runs at 0.77 ms against 0.57 ms for one cell per CTA; removed.)

// ------------------------------------------------------------------------------------------------
// Decoupled models (every variable obeys the same scalar equation): ONE THREAD PER (cell, variable, member).
// Three times the warps of the cell-per-thread kernels at a third of the registers (latency hiding instead of
// 21 % occupancy, profiles/README.md r1f), and with batch = 1 consecutive threads touch consecutive unknowns of
// the cell-major vector.  threadIdx.y / blockIdx walk t = cell * NV + variable.
// ------------------------------------------------------------------------------------------------
// (register cap, as above) one mesh, nx = 1e6: uncapped (60 registers, 4 CTAs) 0.0417 ms; 5 CTAs (48 registers) 0.0356 ms;
// 6 (40 registers) 0.0393 ms; 8 (32 registers) 0.0436 ms.
#ifndef SFXM_SCALAR_RES_MIN_CTAS
#define SFXM_SCALAR_RES_MIN_CTAS 5
#endif
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256, SFXM_SCALAR_RES_MIN_CTAS)
residual_scalar_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                       const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                       double *__restrict__ r_out, double *__restrict__ y_out, double dt,
                       const double *__restrict__ geom_tab) {
  using M1 = typename Model::Scalar;
  int64_t m, t;
  thread_coords(m, t);
  if (m >= P.batch || t >= P.nx * NV) return;
  const int64_t i = t / NV;
  const int j = (int)(t - i * NV);
  double prm[M1::kParams];
#pragma unroll
  for (int k = 0; k < M1::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][1], xs[W], r[1];
  int64_t off[1];
  gather_stencil<1, W>(P, v, i, geom_tab == nullptr, u, xs, off, j);
  Geom<SCHEME, W> geom = geom_tab ? Geom<SCHEME, W>(geom_tab, P.nx, i) : Geom<SCHEME, W>(xs);
  geom.and_inputs(u, prm);
  if (geom.fast) cell_residual<M1, SCHEME, 1, W, true>(prm, u, geom, r);
  else cell_residual<M1, SCHEME, 1, W, false>(prm, u, geom, r);
  if (r_out) r_out[off[0]] = r[0];
  const double uc = (W == 3 || P.nbl == 1) ? u[1][0] : u[W == 3 ? 1 : 2][0];
  if (y_out) { const double yn = uc + dt * r[0]; y_out[off[0]] = yn; push_halo(P, i, j, m, yn); }                // sol.y += delta_t * f, simulation.py:459
}

template <class M1, int SCHEME, int W, bool FAST>
__device__ __forceinline__ void jacobian_scalar_cell(const sfxm_problem_t &P, const double *__restrict__ x_int,
                                                     const double *__restrict__ y, const double *prm, int64_t i,
                                                     int j, int64_t m, bool interior, const double (&u)[W][1],
                                                     const Geom<SCHEME, W> &geom, double eps, double r0,
                                                     double *__restrict__ p, int nv, uint32_t row_mask) {
  // p: entry (d = 0, r = j, c = 0) of this cell/member; row j of block d is at p + (d*nv*nv + c) * B.
  // Writes the diagonal entry (the off-diagonal entries of a decoupled model are structural zeros: never written,
  // see stored_entry).
  const int64_t B = P.batch;
  const double reps = 1.0 / eps;
  // (Round 2: for one mesh, writing the whole row j - diagonal plus structural zeros, 72 contiguous bytes per slot from
  // the three threads of a cell - so that no 32-byte sector is written partially: ncu r2s shows 459 MB of DRAM traffic for
  // the 152 algorithmic MB of a 1e6-cell assembly, but the kernel is latency-bound at 25 % occupancy (98 registers, IPC
  // 1.4), not DRAM-bound: 0.147 ms with full rows against 0.139 ms; removed.)
  auto emit = [&](int d, double val) {
    double *q = p + (int64_t)d * (nv * nv) * B;
    if (row_mask & (1u << j)) q[(int64_t)j * B] = val;
  };
  if (interior) {
#pragma unroll
    for (int d = 0; d < W; ++d) {
      double up[W][1], r1[1];
#pragma unroll
      for (int s = 0; s < W; ++s) up[s][0] = u[s][0];
      up[d][0] = u[d][0] + eps;                             // simulation.py:337
      cell_residual<M1, SCHEME, 1, W, FAST>(prm, up, geom, r1);                  // simulation.py:345
      emit(d, div_known<FAST>(r1[0] - r0, eps, reps));                           // simulation.py:354
    }
    return;
  }
#pragma unroll 1
  for (int d = 0; d < W; ++d) {
    const int64_t pos = i - P.nbl + d;
    int64_t pc = pos;
    bool live = true, halo = false;
    if (pos < 0) {
      halo = (P.bc_ltype[j] == SFXM_BC_HALO);
      live = halo || (P.bc_ltype[j] == SFXM_BC_PERIODIC);
      pc = halo ? -1 : pos + P.nx;
    } else if (pos >= P.nx) {
      halo = (P.bc_rtype[j] == SFXM_BC_HALO);
      live = halo || (P.bc_rtype[j] == SFXM_BC_PERIODIC);
      pc = halo ? -1 : pos - P.nx;
    }
    double val = 0.0;
    if (live) {
      View vp{y, x_int, m, pc, j, eps};
      double up[W][1], r1[1];
#pragma unroll
      for (int s = 0; s < W; ++s) {
        const int64_t ps = i - P.nbl + s;
        if (ps < 0) up[s][0] = ghost_u_left(P, vp, (int)(-ps - 1), j);
        else if (ps >= P.nx) up[s][0] = ghost_u_right(P, vp, (int)(ps - P.nx), j);
        else up[s][0] = (ps == pc) ? u[s][0] + eps : u[s][0];
        if (halo && ps == pos) up[s][0] = up[s][0] + eps;
      }
      cell_residual<M1, SCHEME, 1, W, FAST>(prm, up, geom, r1);
      val = div_known<FAST>(r1[0] - r0, eps, reps);
    }
    emit(d, val);
  }
}

#ifndef SFXM_SCALAR_JAC_MIN_CTAS
// resident 256-thread CTAs per SM the compiler must allow (register cap).  Measured at nx = 1e6 (gpurun_out r2u): no cap
// (98 registers, 2 CTAs) 0.132 ms; 3 CTAs (80 registers, 12 bytes spilled) 0.109 ms; 4 (64 registers) 0.111 ms; 6 (40
// registers, 624 bytes spilled) 0.166 ms - the kernel is latency-bound, occupancy pays until the spills do.
#define SFXM_SCALAR_JAC_MIN_CTAS 3
#endif
template <class Model, int SCHEME, int NV, int W>
__global__ void __launch_bounds__(256, SFXM_SCALAR_JAC_MIN_CTAS)
jacobian_scalar_kernel(const __grid_constant__ sfxm_problem_t P, const double *__restrict__ x_int,
                       const double *__restrict__ params, int64_t param_stride, const double *__restrict__ y,
                       double eps, double *__restrict__ J, double *__restrict__ r0_out, uint32_t entry_mask,
                       const double *__restrict__ geom_tab) {
  using M1 = typename Model::Scalar;
  int64_t m, t;
  thread_coords(m, t);
  if (m >= P.batch || t >= P.nx * NV) return;
  const int64_t i = t / NV;
  const int j = (int)(t - i * NV);
  double prm[M1::kParams];
#pragma unroll
  for (int k = 0; k < M1::kParams; ++k) prm[k] = __ldg(params + (param_stride ? k * param_stride + m : k));
  View v{y, x_int, m, -1, 0, 0.0};
  double u[W][1], xs[W], r0[1];
  int64_t off[1];
  const bool interior = gather_stencil<1, W>(P, v, i, geom_tab == nullptr, u, xs, off, j);
  Geom<SCHEME, W> geom = geom_tab ? Geom<SCHEME, W>(geom_tab, P.nx, i) : Geom<SCHEME, W>(xs);
  geom.and_inputs(u, prm);
  geom.fast = geom.fast && sane(eps, 120);
  const int64_t B = P.batch;
  double *p = J + ((i * W) * (int64_t)(NV * NV) + (int64_t)j * NV) * B + m;
  const uint32_t row_mask = (entry_mask >> (j * NV)) & ((1u << NV) - 1u);
  if (row_mask == 0u && r0_out == nullptr) return;        // variable outside the requested segments
  if (geom.fast) {
    cell_residual<M1, SCHEME, 1, W, true>(prm, u, geom, r0);                     // simulation.py:320
    jacobian_scalar_cell<M1, SCHEME, W, true>(P, x_int, y, prm, i, j, m, interior, u, geom, eps, r0[0], p, NV, row_mask);
  } else {
    cell_residual<M1, SCHEME, 1, W, false>(prm, u, geom, r0);
    jacobian_scalar_cell<M1, SCHEME, W, false>(P, x_int, y, prm, i, j, m, interior, u, geom, eps, r0[0], p, NV, row_mask);
  }
  if (r0_out) r0_out[off[0]] = r0[0];
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
static int ensemble_block_y() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_BLOCK_Y");
    v = e ? atoi(e) : 1;
    if (v < 1 || v > 8) v = 1;
  }
  return v;
}
static bool warp_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_WARP");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_PAIR");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}
static bool residual_ring_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_RES_RING");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}
static int residual_ring_strip() {      // cells per one-warp CTA of residual_ring_kernel
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_RING_STRIP");
    v = e ? atoi(e) : 16;               // measured on the bench ensemble: 16 -> 0.183 ms, 32 -> 0.190, 64 -> 0.196, 250 -> 0.279
    if (v < 1 || v > 65536) v = 16;
  }
  return v;
}
static int ensemble_strip() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SFXM_STRIP");
    v = e ? atoi(e) : 8;
    if (v < 1 || v > 4096) v = 8;
  }
  return v;
}
static dim3 block_for(int64_t batch) { return batch >= 32 ? dim3(32, ensemble_block_y()) : dim3(1, 256); }
// cells per CTA row: one-warp CTAs walk a strip of cells
static int strip_for(dim3 b) { return (b.x == 32 && b.y == 1) ? ensemble_strip() : 1; }
static dim3 grid_for(int64_t batch, int64_t ncells, dim3 b, int strip = 1) {
  const int64_t ctiles = (ncells + (int64_t)b.y * strip - 1) / ((int64_t)b.y * strip);
  if (b.x == 32) {   // x: member tiles, y/z: cell tiles
    const int64_t gy = ctiles < 65535 ? ctiles : 65535;
    return dim3((unsigned)((batch + 31) / 32), (unsigned)gy, (unsigned)((ctiles + gy - 1) / gy));
  }
  return dim3((unsigned)ctiles, (unsigned)batch);   // x: cell tiles, y: members (< 32)
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
  const double *geom;
  int seg_lo, seg_hi;   // Jacobian: assemble the entries of segments [seg_lo, seg_hi) only
  int64_t warp_lo, warp_hi;   // residual, warp-cooperative kernel only: warps [warp_lo, warp_hi) (warp_hi == 0: all)
};

// host twin of sane() (sfxm_device.cuh)
static bool host_sane(double v, unsigned span) {
  uint64_t bits;
  std::memcpy(&bits, &v, sizeof bits);
  const uint32_t a = (uint32_t)(bits >> 32) & 0x7fffffffu;
  return ((uint32_t)(a - ((1023u - span) << 20)) < ((2u * span) << 20)) || ((a | (uint32_t)bits) == 0u);
}

// the ensemble Jacobian kernel addresses with 32-bit row lengths (PairArgs)
template <int NV, int W>
static bool pair_args_fit(const sfxm_problem_t &P) {
  if (P.nx >= (int64_t)1 << 31 || P.batch * (int64_t)(W * NV * NV) >= (int64_t)1 << 32) return false;
  if (P.batch * (int64_t)SFXM_MAX_PARAMS >= (int64_t)1 << 32) return false;   // parameter rows (stride <= batch)
  for (int j = 0; j < NV; ++j)
    if (P.var_stride[j] * P.batch >= (int64_t)1 << 32) return false;
  return true;
}

template <class Model, int SCHEME, int NV, int W>
static int launch_t(const Launch &L) {
  const sfxm_problem_t &P = *L.P;
  uint32_t mask = 0;
  bool restricted = false;
  for (int r = 0; r < NV; ++r)
    for (int c = 0; c < NV; ++c)
      if (P.var_seg[r] == P.var_seg[c]) {
        if (P.var_seg[r] >= L.seg_lo && P.var_seg[r] < L.seg_hi) mask |= 1u << (r * NV + c);
        else restricted = true;
      }
  if (L.warp_hi > 0 && !(L.op == OP_RESIDUAL && P.batch < 32 && !Model::kDecoupled && warp_enabled() && P.nx < ((int64_t)1 << 32)))
    return set_error("sfxm_residual_warps: this model / batch does not use the warp-cooperative residual kernel");
  if constexpr (Model::kDecoupled) {
   if (P.batch < 32) {
    // (ensembles keep one thread per cell: the per-thread overhead - addresses, parameters, geometry, sanity check -
    // is then shared by the NV variables; measured 0.77 ms vs 1.04 ms on the bench shape)
    dim3 b = block_for(P.batch), g = grid_for(P.batch, P.nx * NV, b);
    if (g.z > 65535) return set_error("mesh too large for one launch");
    if (L.op == OP_RESIDUAL)
      residual_scalar_kernel<Model, SCHEME, NV, W><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom);
    else
      jacobian_scalar_kernel<Model, SCHEME, NV, W><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1, mask, L.geom);
    return check_launch("sfxm kernel launch");
  }
  }
  dim3 b = block_for(P.batch);
  if (L.op == OP_RESIDUAL && P.batch >= 64 && P.batch % 2 == 0 && L.geom != nullptr && pair_enabled()) {
    const int strip = ensemble_strip();
    const int64_t tiles = (P.nx + strip - 1) / strip;
    const int64_t gy = tiles < 65535 ? tiles : 65535;
    dim3 g((unsigned)((P.batch / 2 + 31) / 32), (unsigned)gy, (unsigned)((tiles + gy - 1) / gy));
    if (g.z > 65535) return set_error("mesh too large for one launch");
    if constexpr (W == 3) {
      if (residual_ring_enabled()) {
        const int rstrip = residual_ring_strip();
        const int64_t rtiles = (P.nx + rstrip - 1) / rstrip;
        const int64_t rgy = rtiles < 65535 ? rtiles : 65535;
        dim3 rg((unsigned)((P.batch / 2 + 31) / 32), (unsigned)rgy, (unsigned)((rtiles + rgy - 1) / rgy));
        if (L.out1)
          residual_ring_kernel<Model, SCHEME, NV, true><<<rg, dim3(32, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom, rstrip);
        else
          residual_ring_kernel<Model, SCHEME, NV, false><<<rg, dim3(32, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom, rstrip);
        return check_launch("sfxm kernel launch");
      }
    }
    if (L.out1)
      residual_pair_kernel<Model, SCHEME, NV, W, true><<<g, dim3(32, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom, strip);
    else
      residual_pair_kernel<Model, SCHEME, NV, W, false><<<g, dim3(32, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom, strip);
  } else if (L.op == OP_RESIDUAL && P.batch < 32 && !Model::kDecoupled && warp_enabled() && P.nx < ((int64_t)1 << 32)) {
    const int per_warp = 32 - (W - 1);
    const int64_t warps = (P.nx + per_warp - 1) / per_warp;
    const int64_t w0 = L.warp_hi > 0 ? L.warp_lo : 0, w1 = L.warp_hi > 0 ? (L.warp_hi < warps ? L.warp_hi : warps) : warps;
    if (L.warp_hi > 0 && L.warp_lo < 0) {                  // both end warps, one launch
      dim3 g(warps > 1 ? 2u : 1u, (unsigned)P.batch);
      residual_warp_kernel<Model, SCHEME, NV, W><<<g, dim3(256, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, -1, warps);
      return check_launch("sfxm kernel launch");
    }
    if (w0 < 0 || w0 > w1) return set_error("bad warp range");
    if (w1 > w0) {
      dim3 g((unsigned)((w1 - w0 + 7) / 8), (unsigned)P.batch);
      residual_warp_kernel<Model, SCHEME, NV, W><<<g, dim3(256, 1, 1), 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, w0, w1);
    }
  } else if (L.op == OP_RESIDUAL) {
    const int strip = strip_for(b);
    dim3 g = grid_for(P.batch, P.nx, b, strip);
    if (g.z > 65535) return set_error("mesh too large for one launch");
    residual_kernel<Model, SCHEME, NV, W><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.out0, L.out1, L.scalar, L.geom, strip);
  } else if (P.batch >= 64 && P.batch % 2 == 0 && L.geom != nullptr && pair_enabled() && pair_args_fit<NV, W>(P)) {
    const int64_t ctas = (P.nx + SFXM_PAIR_CELLS - 1) / SFXM_PAIR_CELLS;
    const int64_t gy = ctas < 65535 ? ctas : 65535;
    dim3 g((unsigned)((P.batch / 2 + 31) / 32), (unsigned)gy, (unsigned)((ctas + gy - 1) / gy));
    if (g.z > 65535) return set_error("mesh too large for one launch");
    PairArgs<NV> A;
    for (int j = 0; j < NV; ++j) {
      const int64_t base = (P.var_base[j] + P.var_off[j]) * P.batch;
      A.yv[j] = L.y + base;
      A.r0v[j] = L.out1 ? L.out1 + base : nullptr;
      A.row[j] = (uint32_t)(P.var_stride[j] * P.batch);
    }
    A.batch = (uint32_t)P.batch;
    A.nx = (uint32_t)P.nx;
    A.nbl = (uint32_t)P.nbl;
    A.pstride = (uint32_t)L.pstride;
    A.eps_sane = host_sane(L.scalar, 120) ? 1u : 0u;
    A.reps = 1.0 / L.scalar;
    if (restricted)
      jacobian_pair_kernel<Model, SCHEME, NV, W, true><<<g, dim3(32, 1, 1), 0, L.st>>>(P, A, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1, mask, L.geom);
    else
      jacobian_pair_kernel<Model, SCHEME, NV, W, false><<<g, dim3(32, 1, 1), 0, L.st>>>(P, A, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1, mask, L.geom);
  } else {
    dim3 g = grid_for(P.batch, P.nx, b);
    if (g.z > 65535) return set_error("mesh too large for one launch");
    if (restricted)
      jacobian_kernel<Model, SCHEME, NV, W, true><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1, mask, L.geom);
    else
      jacobian_kernel<Model, SCHEME, NV, W, false><<<g, b, 0, L.st>>>(P, L.x, L.params, L.pstride, L.y, L.scalar, L.out0, L.out1, mask, L.geom);
  }
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

}  // namespace sfxm
