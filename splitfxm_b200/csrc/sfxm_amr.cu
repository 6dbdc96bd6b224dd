// sfxm_amr.cu - adaptive mesh refinement as reduce / mark / scan / compact kernels (sm_100a).
//
// Replaces Refiner.refine + perform_changes (refine.py:90-282, a port of Cantera's refine.cpp), which walk
// Python dicts cell by cell.  The reference's `keep` / `loc` dictionaries are order-independent except for one
// loop:
//   keep[j] ends as  1  if ANY rule assigned 1 to it (a later 1 overwrites -1, refine.py:169-170,194,204-227),
//                   -1  if some rule visited it with r < prune and no rule ever assigned 1 (refine.py:171-173,195-197)
//   loc[j]  ends as  1  if ANY rule asked for a point after j
// so marking is a scatter of idempotent flag stores.  The only sequential step, "never delete two adjacent points"
// (refine.py:231-233), flips every odd offset inside a maximal run of deletions: a max-scan of run starts.
// New positions: kept cell j lands at kept_before(j) + inserted_before(j), the midpoint after j one past the
// last kept cell <= j (the reference appends the midpoints and sorts by x, refine.py:267-279; a midpoint lies
// strictly between its two parents, so the sort is this merge).
//
// All kernels stream x and U once: HBM-bound (~96 B per cell over the mark and apply passes).
//
// Slab-decomposed meshes: every kernel works on a WINDOW of the global mesh - `ne` cells starting at
// global index g0, of which [own_lo, own_hi) are this rank's own and the rest are overlap copied from the neighbours.
// A rule evaluated at cell j reads cells j-2 .. j+2 and sets flags on j-2 .. j+2, so six overlap cells per side make
// the flags of every own cell (and the deletion state of the cell before the first own one) complete.  Global
// quantities cross ranks between the phases: the min/max statistics (one 16-double all-reduce) and the start of the
// deletion run that reaches into the slab (one int64 all-gather).  One mesh on one GPU is the window g0 = 0, ne = n.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"
#include "sfxm_device.cuh"   // div_known / sane: exact quotients by a divisor whose reciprocal is shared (section 3.1)

namespace sfxm {

static constexpr double kUlp = 2.220446049250313e-16;   // math.ulp(1.0), refine.py:7

struct Window {   // a rank's view of the global mesh: array index t <-> global cell g0 + t
  int64_t ne, g0, ng, own_lo, own_hi;
};

struct Stats {   // per variable: min/max of values and of slopes (refine.py:146-153)
  double vmin[SFXM_MAX_NV], vmax[SFXM_MAX_NV], smin[SFXM_MAX_NV], smax[SFXM_MAX_NV];
};

// One mesh on one GPU: what the marking kernel divides by for every cell, prepared once by the statistics kernel.
// The refinement kernels were bound by their IEEE divisions (ncu r2h: issue slots 71-73 % busy, ~17 divisions of ~35
// instructions per cell in the marking kernel); a quotient by a divisor whose correctly rounded reciprocal is known costs
// five FMAs and has the same bits (div_known; each thread checks its operands and otherwise divides).
struct MarkConst {
  double rdmax_slope[SFXM_MAX_NV];   // RN(1 / (slope * (vmax - vmin) + ulp)), refine.py:160
  int32_t ok[SFXM_MAX_NV];           // divisor within 2^+-120
};

// a / b with rb = RN(1 / b): five FMAs when the operands are sane, the IEEE division otherwise (same result)
__device__ __forceinline__ double div_shared(double a, double b, double rb, bool b_ok) {
  return (b_ok && sane(a, 120)) ? div_known<true>(a, b, rb) : div_ieee(a, b);
}
__device__ __forceinline__ bool divisor_ok(double b) { return sane(b, 120) && b != 0.0; }

__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// stage 1: per-block min/max of v_j and s_j = (v_{j+1} - v_j) / (z_{j+1} - z_j); stage 2 (one block) folds them
template <int NV>
__global__ void __launch_bounds__(256)
amr_stats_kernel(const double *__restrict__ z, const double *__restrict__ U, const Window wd, double *__restrict__ part) {
  double lo_v[NV], hi_v[NV], lo_s[NV], hi_s[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { lo_v[i] = lo_s[i] = INFINITY; hi_v[i] = hi_s[i] = -INFINITY; }
  for (int64_t j = wd.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < wd.own_hi; j += (int64_t)gridDim.x * blockDim.x) {
    const bool has_s = wd.g0 + j + 1 < wd.ng;       // the slope of interval j belongs to the rank that owns cell j
    const double dz = has_s ? __ldg(z + j + 1) - __ldg(z + j) : 1.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double v = __ldg(U + j * NV + i);
      lo_v[i] = fmin(lo_v[i], v);
      hi_v[i] = fmax(hi_v[i], v);
      if (has_s) {
        const double s = (__ldg(U + (j + 1) * NV + i) - v) / dz;
        lo_s[i] = fmin(lo_s[i], s);
        hi_s[i] = fmax(hi_s[i], s);
      }
    }
  }
  __shared__ double sh[8][4 * NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double a = warp_min(lo_v[i]), b = warp_max(hi_v[i]), c = warp_min(lo_s[i]), d = warp_max(hi_s[i]);
    if (lane == 0) { sh[w][4 * i] = a; sh[w][4 * i + 1] = b; sh[w][4 * i + 2] = c; sh[w][4 * i + 3] = d; }
  }
  __syncthreads();
  if (threadIdx.x < 4 * NV) {
    const bool is_min = (threadIdx.x & 1) == 0;
    double acc = sh[0][threadIdx.x];
    for (int k = 1; k < 8; ++k) acc = is_min ? fmin(acc, sh[k][threadIdx.x]) : fmax(acc, sh[k][threadIdx.x]);
    part[(int64_t)blockIdx.x * 4 * NV + threadIdx.x] = acc;
  }
}

template <int NV>
__global__ void amr_stats_final_kernel(const double *__restrict__ part, int nblocks, Stats *__restrict__ out) {
  const int t = threadIdx.x;
  if (t >= 4 * NV) return;
  const bool is_min = (t & 1) == 0;
  double acc = part[t];
  for (int k = 1; k < nblocks; ++k) acc = is_min ? fmin(acc, part[(int64_t)k * 4 * NV + t]) : fmax(acc, part[(int64_t)k * 4 * NV + t]);
  const int i = t >> 2, f = t & 3;
  if (f == 0) out->vmin[i] = acc;
  else if (f == 1) out->vmax[i] = acc;
  else if (f == 2) out->smin[i] = acc;
  else out->smax[i] = acc;
}

struct Criteria {
  double ratio, slope, curve, prune, min_range, min_grid;
};

// (Round 2: collecting a thread's flags in a word, exchanging the words between the lanes of the warp and storing every
// cell once per array instead of a dozen scattered, mostly duplicate byte stores was measured SLOWER - 343 us against
// 296 us at 1e7 cells (gpurun_out r2t): the kernel is bound by its instruction stream, not by the stores; removed.)
// flags: keep1[j] / visited[j] / loc[j] (bytes, zero-initialised), comp[i] = component i asked for a point,
// gridpt[j] = the grid-ratio rule named "point j" (refine.py:203,211)
// (register cap) whole pass at 1e7 cells: marking uncapped (58 registers) 0.711 ms; 5 CTAs per SM 0.687 ms; 8: 0.701 ms
#ifndef SFXM_AMR_MARK_MIN_CTAS
#define SFXM_AMR_MARK_MIN_CTAS 5
#endif
template <int NV>
__global__ void __launch_bounds__(256, SFXM_AMR_MARK_MIN_CTAS)
amr_mark_kernel(const double *__restrict__ z, const double *__restrict__ U, const Window wd, const Criteria cr,
                const Stats *__restrict__ st, uint8_t *__restrict__ keep1, uint8_t *__restrict__ visited,
                uint8_t *__restrict__ loc, int32_t *__restrict__ comp, uint8_t *__restrict__ gridpt,
                const MarkConst *__restrict__ mc) {
  // t: index into the window arrays; j: the global cell index every rule of refine.py is written in
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wd.ne) return;
  const int64_t j = wd.g0 + t, n = wd.ng;
  // cells of the overlap whose own stencil leaves the window are someone else's to evaluate
#pragma unroll
  for (int o = -2; o <= 2; ++o)
    if (j + o >= 0 && j + o < n && (t + o < 0 || t + o >= wd.ne)) return;
  if (j == 0 || j == n - 1) keep1[t] = 1;                              // refine.py:120-121
  const bool own = t >= wd.own_lo && t < wd.own_hi;                    // only own cells report the component
  const double zj = __ldg(z + t);
  const double dzj = j + 1 < n ? __ldg(z + t + 1) - zj : 0.0;          // dz[j]
  const double dzj1 = j + 2 < n ? __ldg(z + t + 2) - __ldg(z + t + 1) : 0.0;   // dz[j+1]
  // reciprocals shared by the NV slopes of an interval (mc == nullptr: the window path, plain divisions)
  const bool z0_ok = mc != nullptr && divisor_ok(dzj), z1_ok = mc != nullptr && divisor_ok(dzj1);
  const double rz0 = 1.0 / dzj, rz1 = 1.0 / dzj1;
  const double ulp_dz = div_shared(kUlp, dzj, rz0, z0_ok);                    // kUlp / dz[j], refine.py:188
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double vmin = st->vmin[i], vmax = st->vmax[i], smin = st->smin[i], smax = st->smax[i];
    const double aa = fmax(fabs(vmin), fabs(vmax));
    const double ss = fmax(fabs(smin), fabs(smax));
    const double v0 = __ldg(U + t * NV + i);
    const double v1 = j + 1 < n ? __ldg(U + (t + 1) * NV + i) : 0.0;
    // slope criterion on interval j (refine.py:159-173)
    if (j + 1 < n && (vmax - vmin) > cr.min_range * aa) {
      const double dmax = cr.slope * (vmax - vmin) + kUlp;
      const double r = mc != nullptr ? div_shared(fabs(v1 - v0), dmax, mc->rdmax_slope[i], mc->ok[i] != 0) : fabs(v1 - v0) / dmax;
      if (r > 1.0 && dzj >= 2 * cr.min_grid) { loc[t] = 1; if (own) comp[i] = 1; }
      if (r >= cr.prune) { keep1[t] = 1; keep1[t + 1] = 1; }
      else visited[t] = 1;
    }
    // curvature criterion on the pair of intervals (j, j+1) (refine.py:179-197)
    if (j + 2 < n && (smax - smin) > cr.min_range * ss) {
      const double v2 = __ldg(U + (t + 2) * NV + i);
      const double s0 = div_shared(v1 - v0, dzj, rz0, z0_ok);
      const double s1 = div_shared(v2 - v1, dzj1, rz1, z1_ok);
      const double dmax = cr.curve * (smax - smin);
      const double r = fabs(s1 - s0) / (dmax + ulp_dz);
      if (r > 1.0 && dzj >= 2 * cr.min_grid && dzj1 >= 2 * cr.min_grid) { if (own) comp[i] = 1; loc[t] = 1; loc[t + 1] = 1; }
      if (r >= cr.prune) keep1[t + 1] = 1;
      else visited[t + 1] = 1;
    }
  }
  // grid-ratio rules (refine.py:200-227), j = 1 .. n-2
  if (j >= 1 && j <= n - 2) {
    const double dzm1 = zj - __ldg(z + t - 1);                         // dz[j-1]
    if (dzj > cr.ratio * dzm1) {
      loc[t] = 1; gridpt[t] = 1;
      keep1[t - 1] = 1; keep1[t] = 1; keep1[t + 1] = 1;
      if (j + 2 < n) keep1[t + 2] = 1;
    }
    if (dzj < dzm1 / cr.ratio) {
      loc[t - 1] = 1; gridpt[t - 1] = 1;
      if (j >= 2) keep1[t - 2] = 1;
      keep1[t - 1] = 1; keep1[t] = 1; keep1[t + 1] = 1;
    }
    const double span = __ldg(z + t + 1) - __ldg(z + t - 1);
    if (j > 1 && span > cr.ratio * (__ldg(z + t - 1) - __ldg(z + t - 2))) keep1[t] = 1;
    if (j < n - 2 && span > cr.ratio * dzj1) keep1[t] = 1;
  }
}

// start[t] = global index of cell t if it begins a maximal run of deletions, else -1
__global__ void amr_run_start_kernel(const uint8_t *__restrict__ keep1, const uint8_t *__restrict__ visited, const Window wd,
                                     int64_t *__restrict__ start) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wd.ne) return;
  const bool del = visited[t] && !keep1[t];
  const bool prev = t > 0 && visited[t - 1] && !keep1[t - 1];
  start[t] = (del && !prev) ? wd.g0 + t : -1;
}

struct MaxOp {
  __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; }
};

// final flags: keepf[t] = 0 only for deletions at even offsets of their run (refine.py:231-233); counts for the scans.
// carry: start of the deletion run that reaches this slab from the ranks before it (-1: none / one mesh).
__global__ void amr_finalize_kernel(const uint8_t *__restrict__ keep1, const uint8_t *__restrict__ visited,
                                    const int64_t *__restrict__ run_start, const uint8_t *__restrict__ loc, const Window wd,
                                    int64_t carry, uint8_t *__restrict__ keepf, int64_t *__restrict__ kept_cnt,
                                    int64_t *__restrict__ ins_cnt) {
  const int64_t t = wd.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wd.own_hi) return;
  const int64_t j = wd.g0 + t;
  const bool del0 = visited[t] && !keep1[t];
  const int64_t rs = run_start[t] > carry ? run_start[t] : carry;
  const bool del = del0 && (((j - rs) & 1) == 0);
  keepf[t] = del ? 0 : 1;
  kept_cnt[t] = del ? 0 : 1;
  ins_cnt[t] = (loc[t] && j + 1 < wd.ng) ? 1 : 0;
}

// last own value of the run-start scan: what the ranks after this one need (the latest run start so far)
__global__ void amr_last_run_kernel(const int64_t *__restrict__ run_start, const Window wd, int64_t *__restrict__ out) {
  out[0] = run_start[wd.own_hi - 1];
}

template <int NV>
__global__ void amr_apply_kernel(const double *__restrict__ z, const double *__restrict__ U, const Window wd,
                                 const uint8_t *__restrict__ keepf, const uint8_t *__restrict__ loc,
                                 const int64_t *__restrict__ kept_before, const int64_t *__restrict__ ins_before,
                                 double *__restrict__ z_out, double *__restrict__ U_out) {
  const int64_t t = wd.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wd.own_hi) return;
  const int64_t kb = kept_before[t], ib = ins_before[t];
  const bool ins = loc[t] && wd.g0 + t + 1 < wd.ng;       // the midpoint after an own cell is this rank's
  if (keepf[t]) {
    const int64_t o = kb + ib;
    z_out[o] = __ldg(z + t);
#pragma unroll
    for (int i = 0; i < NV; ++i) U_out[o * NV + i] = __ldg(U + t * NV + i);
  }
  if (ins) {
    const int64_t o = kb + (keepf[t] ? 1 : 0) + ib;
    z_out[o] = 0.5 * (__ldg(z + t + 1) + __ldg(z + t));                               // refine.py:268
#pragma unroll
    for (int i = 0; i < NV; ++i) U_out[o * NV + i] = 0.5 * (__ldg(U + (t + 1) * NV + i) + __ldg(U + t * NV + i));   // :269
  }
}

__global__ void amr_count_kernel(const int64_t *__restrict__ kept_before, const int64_t *__restrict__ ins_before,
                                 const int64_t *__restrict__ kept_cnt, const int64_t *__restrict__ ins_cnt, const Window wd,
                                 int64_t *__restrict__ counts) {
  // counts[0] = new size, counts[1] = inserted, counts[2] = deleted   (own cells)
  const int64_t last = wd.own_hi - 1;
  const int64_t kept = kept_before[last] + kept_cnt[last];
  const int64_t ins = ins_before[last] + ins_cnt[last];
  counts[0] = kept + ins;
  counts[1] = ins;
  counts[2] = (wd.own_hi - wd.own_lo) - kept;
}

static inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

struct AmrWork {
  // layout of the workspace
  int64_t off_part, off_stats, off_keep1, off_visited, off_loc, off_keepf, off_gridpt, off_comp, off_start, off_run,
      off_kcnt, off_icnt, off_kbef, off_ibef, off_cub, cub_bytes, total;
};

static AmrWork amr_layout(int64_t n) {
  AmrWork w{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += align256(bytes); return r; };
  w.off_part = take(1024 * 4 * SFXM_MAX_NV * 8);
  w.off_stats = take(sizeof(Stats));
  w.off_keep1 = take(n);
  w.off_visited = take(n);
  w.off_loc = take(n);
  w.off_keepf = take(n);
  w.off_gridpt = take(n);
  w.off_comp = take(SFXM_MAX_NV * 4);
  w.off_start = take(n * 8);
  w.off_run = take(n * 8);
  w.off_kcnt = take(n * 8);
  w.off_icnt = take(n * 8);
  w.off_kbef = take(n * 8);
  w.off_ibef = take(n * 8);
  size_t b1 = 0, b2 = 0;
  cub::DeviceScan::InclusiveScan((void *)nullptr, b1, (int64_t *)nullptr, (int64_t *)nullptr, MaxOp(), (int)n);
  cub::DeviceScan::ExclusiveSum((void *)nullptr, b2, (int64_t *)nullptr, (int64_t *)nullptr, (int)n);
  w.cub_bytes = (int64_t)(b1 > b2 ? b1 : b2);
  w.off_cub = take(w.cub_bytes + 256);
  w.total = o;
  return w;
}

static bool window_ok(const Window &wd) {
  return wd.ne >= 3 && wd.g0 >= 0 && wd.g0 + wd.ne <= wd.ng && wd.own_lo >= 0 && wd.own_lo < wd.own_hi && wd.own_hi <= wd.ne &&
         wd.ne <= 2147483647LL;
}

// phase 1: local min/max of own values and slopes -> stats_dev[16] = vmin[4] | vmax[4] | smin[4] | smax[4]
template <int NV>
static int amr_stats_t(const double *z, const double *U, const Window &wd, char *work, double *stats_dev, cudaStream_t st) {
  const AmrWork w = amr_layout(wd.ne);
  const int64_t n_own = wd.own_hi - wd.own_lo;
  const unsigned blocks = (unsigned)((n_own + 255) / 256);
  const int sblocks = (int)(blocks < 1024 ? blocks : 1024);
  amr_stats_kernel<NV><<<sblocks, 256, 0, st>>>(z, U, wd, (double *)(work + w.off_part));
  amr_stats_final_kernel<NV><<<1, 32, 0, st>>>((const double *)(work + w.off_part), sblocks, (Stats *)stats_dev);
  return check_launch("sfxm_refine stats");
}

// phase 2: flags over the window, run starts, max-scan over the own cells
template <int NV>
static int amr_mark_t(const double *z, const double *U, const Window &wd, const Criteria &cr, const double *stats_dev,
                      char *work, int64_t *last_run_dev, cudaStream_t st) {
  const AmrWork w = amr_layout(wd.ne);
  const int threads = 256;
  const unsigned blocks = (unsigned)((wd.ne + threads - 1) / threads);
  const int64_t n_own = wd.own_hi - wd.own_lo;
  cudaMemsetAsync(work + w.off_keep1, 0, w.off_comp + align256(SFXM_MAX_NV * 4) - w.off_keep1, st);
  amr_mark_kernel<NV><<<blocks, threads, 0, st>>>(z, U, wd, cr, (const Stats *)stats_dev,
                                                  (uint8_t *)(work + w.off_keep1), (uint8_t *)(work + w.off_visited),
                                                  (uint8_t *)(work + w.off_loc), (int32_t *)(work + w.off_comp),
                                                  (uint8_t *)(work + w.off_gridpt), nullptr);
  amr_run_start_kernel<<<blocks, threads, 0, st>>>((const uint8_t *)(work + w.off_keep1),
                                                   (const uint8_t *)(work + w.off_visited), wd, (int64_t *)(work + w.off_start));
  size_t cb = (size_t)w.cub_bytes;
  cub::DeviceScan::InclusiveScan((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_start) + wd.own_lo,
                                 (int64_t *)(work + w.off_run) + wd.own_lo, MaxOp(), (int)n_own, st);
  if (last_run_dev) amr_last_run_kernel<<<1, 1, 0, st>>>((const int64_t *)(work + w.off_run), wd, last_run_dev);
  return check_launch("sfxm_refine mark");
}

// phase 3: adjacent-deletion rule with the carried-in run start, output slots of the own cells, counts
static int amr_finalize_t(const Window &wd, int64_t carry, char *work, int64_t *counts_dev, cudaStream_t st) {
  const AmrWork w = amr_layout(wd.ne);
  const int threads = 256;
  const int64_t n_own = wd.own_hi - wd.own_lo;
  const unsigned blocks = (unsigned)((n_own + threads - 1) / threads);
  amr_finalize_kernel<<<blocks, threads, 0, st>>>((const uint8_t *)(work + w.off_keep1), (const uint8_t *)(work + w.off_visited),
                                                  (const int64_t *)(work + w.off_run), (const uint8_t *)(work + w.off_loc), wd,
                                                  carry, (uint8_t *)(work + w.off_keepf), (int64_t *)(work + w.off_kcnt),
                                                  (int64_t *)(work + w.off_icnt));
  size_t cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_kcnt) + wd.own_lo,
                                (int64_t *)(work + w.off_kbef) + wd.own_lo, (int)n_own, st);
  cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_icnt) + wd.own_lo,
                                (int64_t *)(work + w.off_ibef) + wd.own_lo, (int)n_own, st);
  amr_count_kernel<<<1, 1, 0, st>>>((const int64_t *)(work + w.off_kbef), (const int64_t *)(work + w.off_ibef),
                                    (const int64_t *)(work + w.off_kcnt), (const int64_t *)(work + w.off_icnt), wd, counts_dev);
  return check_launch("sfxm_refine finalize");
}

template <int NV>
static int amr_apply_t(const double *z, const double *U, const Window &wd, const char *wk, double *x_out, double *U_out,
                       cudaStream_t st) {
  const AmrWork w = amr_layout(wd.ne);
  const int threads = 256;
  const unsigned blocks = (unsigned)((wd.own_hi - wd.own_lo + threads - 1) / threads);
  amr_apply_kernel<NV><<<blocks, threads, 0, st>>>(z, U, wd, (const uint8_t *)(wk + w.off_keepf),
                                                   (const uint8_t *)(wk + w.off_loc), (const int64_t *)(wk + w.off_kbef),
                                                   (const int64_t *)(wk + w.off_ibef), x_out, U_out);
  return check_launch("sfxm_refine_apply");
}

// explicit flags (Refiner.perform_changes(d, loc, keep), refine.py:239-282): keep_in[t] = 0 deletes cell t, loc_in[t] = 1
// inserts the midpoint after cell t
__global__ void amr_set_flags_kernel(const uint8_t *__restrict__ loc_in, const uint8_t *__restrict__ keep_in, int64_t n,
                                     uint8_t *__restrict__ loc, uint8_t *__restrict__ keepf, int64_t *__restrict__ kept_cnt,
                                     int64_t *__restrict__ ins_cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint8_t k = keep_in[t] ? 1 : 0, l = loc_in[t] ? 1 : 0;
  loc[t] = l;
  keepf[t] = k;
  kept_cnt[t] = k;
  ins_cnt[t] = (l && t + 1 < n) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// One mesh on one GPU in THREE launches (round 2; the window phases above stay for slab-decomposed meshes):
//   1. amr_stats_zero_kernel  : per-block min/max partials, the last block to finish folds them (ticket counter);
//                               the same grid zero-fills the flag bytes and the tile descriptors of launch 3
//   2. amr_mark_kernel        : unchanged
//   3. amr_scan_apply_kernel  : ONE pass over the flags with two chained decoupled look-backs per tile - the max-scan
//                               of deletion-run starts (refine.py:231-233), then, with the final keep flags it yields,
//                               the exclusive counts of kept and inserted cells - and the new mesh is written straight
//                               from the tile (capacity 2n - 1; the new size comes back in counts[0]).
// Replaces 10 launches incl. three cub::DeviceScan passes over 8-byte arrays (profile r1: 0.178 ms at nx = 1e6).
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kScanThreads = 256, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

struct TileDesc {                 // status in the two top bits: 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix
  unsigned long long run;         // low 62 bits: 1 + latest deletion-run start (0 = none)
  unsigned long long cnt;         // low 62 bits: kept << 31 | inserted
};

template <int NV>
__global__ void __launch_bounds__(256)
amr_stats_zero_kernel(const double *__restrict__ z, const double *__restrict__ U, const Window wd, double *__restrict__ part,
                      Stats *__restrict__ out, MarkConst *__restrict__ mc, double slope, unsigned int *__restrict__ ticket,
                      uint4 *__restrict__ zero_lo, int64_t zero_n16) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < zero_n16; q += (int64_t)gridDim.x * blockDim.x)
    zero_lo[q] = make_uint4(0u, 0u, 0u, 0u);
  double lo_v[NV], hi_v[NV], lo_s[NV], hi_s[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { lo_v[i] = lo_s[i] = INFINITY; hi_v[i] = hi_s[i] = -INFINITY; }
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < wd.ne; j += (int64_t)gridDim.x * blockDim.x) {
    const bool has_s = j + 1 < wd.ne;
    const double dz = has_s ? __ldg(z + j + 1) - __ldg(z + j) : 1.0;
    const bool dz_ok = divisor_ok(dz);
    const double rz = 1.0 / dz;                            // shared by the NV slopes of the interval
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double v = __ldg(U + j * NV + i);
      lo_v[i] = fmin(lo_v[i], v);
      hi_v[i] = fmax(hi_v[i], v);
      if (has_s) {
        const double sl = div_shared(__ldg(U + (j + 1) * NV + i) - v, dz, rz, dz_ok);
        lo_s[i] = fmin(lo_s[i], sl);
        hi_s[i] = fmax(hi_s[i], sl);
      }
    }
  }
  __shared__ double sh[8][4 * NV];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double a = warp_min(lo_v[i]), b = warp_max(hi_v[i]), c = warp_min(lo_s[i]), d = warp_max(hi_s[i]);
    if (lane == 0) { sh[w][4 * i] = a; sh[w][4 * i + 1] = b; sh[w][4 * i + 2] = c; sh[w][4 * i + 3] = d; }
  }
  __syncthreads();
  if (threadIdx.x < 4 * NV) {
    const bool is_min = (threadIdx.x & 1) == 0;
    double acc = sh[0][threadIdx.x];
    for (int k = 1; k < 8; ++k) acc = is_min ? fmin(acc, sh[k][threadIdx.x]) : fmax(acc, sh[k][threadIdx.x]);
    part[(int64_t)blockIdx.x * 4 * NV + threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // min / max are exact: the fold order does not matter.  256 threads fold the partials of the 4 NV statistics in
  // groups (a one-thread-per-statistic loop over ~600 partials was half of this kernel's time), then 4 NV threads finish.
  __shared__ double fold[256];
  constexpr int NS = 4 * NV, NG = 256 / NS;
  {
    const int v = threadIdx.x % NS, g = threadIdx.x / NS;
    const bool is_min = (v & 1) == 0;
    double acc = is_min ? INFINITY : -INFINITY;
    if (g < NG) {
      const volatile double *vp = part;
      for (unsigned k = g; k < gridDim.x; k += NG) acc = is_min ? fmin(acc, vp[(int64_t)k * NS + v]) : fmax(acc, vp[(int64_t)k * NS + v]);
    }
    fold[threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    const int t = threadIdx.x;
    const bool is_min = (t & 1) == 0;
    double acc = fold[t];
    for (int g = 1; g < NG; ++g) acc = is_min ? fmin(acc, fold[g * NS + t]) : fmax(acc, fold[g * NS + t]);
    const int i = t >> 2, f = t & 3;
    if (f == 0) out->vmin[i] = acc;
    else if (f == 1) out->vmax[i] = acc;
    else if (f == 2) out->smin[i] = acc;
    else out->smax[i] = acc;
    fold[t] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NV) {                                  // the slope criterion's divisor, the same expression as in
    const int i = threadIdx.x;                             // amr_mark_kernel (refine.py:160), and its reciprocal
    const double dmax = slope * (fold[4 * i + 1] - fold[4 * i]) + kUlp;
    mc->rdmax_slope[i] = 1.0 / dmax;
    mc->ok[i] = divisor_ok(dmax) ? 1 : 0;
  }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Exclusive prefix of tile `tile` from the descriptors of the tiles before it (warp 0 of the block; standard decoupled
// look-back, 32 predecessors per step).  MAXOP: combine with max (run starts), else with + (packed counts).
template <bool MAXOP>
__device__ __forceinline__ unsigned long long tile_lookback(const unsigned long long *desc, int stride_u64, int tile) {
  constexpr unsigned long long kMask = (1ull << 62) - 1;
  const int lane = threadIdx.x & 31;
  unsigned long long acc = 0;
  for (int base = tile - 1; base >= 0; base -= 32) {
    const int p = base - lane;
    unsigned long long v = 2ull << 62;                      // tiles before the first one: inclusive prefix = identity
    if (p >= 0) {
      do { v = ld_volatile_u64(desc + (int64_t)p * stride_u64); } while ((v >> 62) == 0ull);
    }
    const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
    const int stop = incl ? (__ffs(incl) - 1) : 31;         // nearest predecessor (lowest lane) with an inclusive prefix
    unsigned long long val = lane <= stop ? (v & kMask) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, val, o);
      val = MAXOP ? (val > other ? val : other) : (val + other);
    }
    acc = MAXOP ? (acc > val ? acc : val) : (acc + val);
    if (incl) break;
  }
  return acc;
}

// Cell <-> (warp, item, lane): warp w of tile T owns the 128 contiguous cells from T * 1024 + w * 128; its item k of
// lane L is cell base + k * 32 + L, so every load and store of a warp instruction touches 32 consecutive cells (the first
// version gave each thread 4 consecutive cells: 32-byte strides between lanes, 4x the sectors - ncu r2m 420 us at 1e7
// cells).  The scans run in cell order: a warp scan per item, the carry of item k - 1 entering item k.
template <int NV>
__global__ void __launch_bounds__(kScanThreads, 5)
amr_scan_apply_kernel(const double *__restrict__ z, const double *__restrict__ U, int64_t n,
                      const uint8_t *__restrict__ keep1, const uint8_t *__restrict__ visited, const uint8_t *__restrict__ loc,
                      uint8_t *__restrict__ keepf, TileDesc *__restrict__ desc, unsigned int *__restrict__ tile_counter,
                      double *__restrict__ z_out, double *__restrict__ U_out, int64_t *__restrict__ counts) {
  constexpr unsigned long long kMask = (1ull << 62) - 1;
  constexpr int kWarps = kScanThreads / 32;
  __shared__ int s_tile;
  __shared__ unsigned long long s_warp[kWarps];
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(tile_counter, 1u);     // tiles start in index order: look-back cannot deadlock
  __syncthreads();
  const int tile = s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t wbase = (int64_t)tile * kScanTile + (int64_t)w * (32 * kScanItems);
  bool del0[kScanItems], lc[kScanItems];
  // flags of the thread's items (coalesced).  The kernel is a chain of dependent memory round trips per tile (flags ->
  // look-back 1 -> look-back 2 -> cells -> stores): what hides them is resident tiles, so registers are capped for five
  // CTAs per SM and the cells are loaded where they are used
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t t = wbase + k * 32 + lane;
    const bool in = t < n;
    del0[k] = in && visited[t] && !keep1[t];
    lc[k] = in && loc[t] && t + 1 < n;
  }
  // ---- scan 1: inclusive max of (1 + run start) in cell order ---------------------------------------------------
  unsigned long long run[kScanItems];          // inclusive over the warp's cells up to this one
  unsigned long long carry = 0;
  {
    bool last_prev = wbase > 0 && wbase - 1 < n && visited[wbase - 1] && !keep1[wbase - 1];    // cell before the warp's first
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int64_t t = wbase + k * 32 + lane;
      const bool up = __shfl_up_sync(0xffffffffu, del0[k] ? 1 : 0, 1) != 0;
      const bool prev = lane == 0 ? last_prev : up;
      unsigned long long v = (del0[k] && !prev) ? (unsigned long long)(t + 1) : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long other = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o && other > v) v = other;
      }
      if (carry > v) v = carry;
      run[k] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
      last_prev = __shfl_sync(0xffffffffu, del0[k] ? 1 : 0, 31) != 0;
    }
  }
  if (lane == 31) s_warp[w] = carry;           // the warp's aggregate
  __syncthreads();
  unsigned long long before = 0, agg = 0;      // max over the warps before this one / over the tile
#pragma unroll
  for (int q = 0; q < kWarps; ++q) {
    const unsigned long long v = s_warp[q];
    if (q < w && v > before) before = v;
    if (v > agg) agg = v;
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long *>(&desc[tile].run) = (tile == 0 ? (2ull << 62) : (1ull << 62)) | agg;
  }
  if (w == 0) {
    unsigned long long excl = 0;
    if (tile > 0) {
      excl = tile_lookback<true>(&desc[0].run, (int)(sizeof(TileDesc) / 8), tile);
      if (lane == 0) {
        const unsigned long long inc = excl > agg ? excl : agg;
        __threadfence();
        *reinterpret_cast<volatile unsigned long long *>(&desc[tile].run) = (2ull << 62) | inc;
      }
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  const unsigned long long base_run = s_prefix > before ? s_prefix : before;
  // final keep flags (refine.py:231-233: inside a run of deletions every odd offset is kept) and the counts
  bool kp[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t t = wbase + k * 32 + lane;
    const unsigned long long rs1 = run[k] > base_run ? run[k] : base_run;               // 1 + run start
    const bool del = del0[k] && (((t - ((int64_t)rs1 - 1)) & 1) == 0);
    kp[k] = (t < n) && !del;
  }
  __syncthreads();                                        // s_warp / s_prefix are reused
  // ---- scan 2: exclusive sums of (kept << 31 | inserted) in cell order -------------------------------------------
  unsigned long long pos[kScanItems];          // exclusive over the warp's cells before this one
  unsigned long long wsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const unsigned long long mine = (kp[k] ? (1ull << 31) : 0ull) + (lc[k] ? 1ull : 0ull);
    unsigned long long v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long other = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += other;
    }
    pos[k] = wsum + v - mine;
    wsum += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 31) s_warp[w] = wsum;
  __syncthreads();
  unsigned long long tbefore = 0, tagg = 0;
#pragma unroll
  for (int q = 0; q < kWarps; ++q) {
    const unsigned long long v = s_warp[q];
    if (q < w) tbefore += v;
    tagg += v;
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long *>(&desc[tile].cnt) = (tile == 0 ? (2ull << 62) : (1ull << 62)) | tagg;
  }
  if (w == 0) {
    unsigned long long excl = 0;
    if (tile > 0) {
      excl = tile_lookback<false>(&desc[0].cnt, (int)(sizeof(TileDesc) / 8), tile);
      if (lane == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned long long *>(&desc[tile].cnt) = (2ull << 62) | ((excl + tagg) & kMask);
      }
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  const unsigned long long tile_pos = s_prefix + tbefore;
  // ---- write the new mesh --------------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t t = wbase + k * 32 + lane;
    if (t >= n) continue;
    keepf[t] = kp[k] ? 1 : 0;
    const unsigned long long p = tile_pos + pos[k];
    const int64_t kb = (int64_t)(p >> 31), ib = (int64_t)(p & ((1ull << 31) - 1));
    if (!kp[k] && !lc[k]) continue;
    const double zc = __ldg(z + t);
    double Uc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) Uc[i] = __ldg(U + t * NV + i);
    if (kp[k]) {
      const int64_t o = kb + ib;
      z_out[o] = zc;
#pragma unroll
      for (int i = 0; i < NV; ++i) U_out[o * NV + i] = Uc[i];
    }
    if (lc[k]) {                                           // the next cell is loaded by the neighbouring lane too: L1 hit
      const int64_t o = kb + (kp[k] ? 1 : 0) + ib;
      z_out[o] = 0.5 * (__ldg(z + t + 1) + zc);                                           // refine.py:268
#pragma unroll
      for (int i = 0; i < NV; ++i) U_out[o * NV + i] = 0.5 * (__ldg(U + (t + 1) * NV + i) + Uc[i]);      // :269
    }
    if (t == n - 1) {                                      // the thread that owns the last cell reports the totals
      const int64_t kept = kb + (kp[k] ? 1 : 0), ins = ib + (lc[k] ? 1 : 0);
      counts[0] = kept + ins;
      counts[1] = ins;
      counts[2] = n - kept;
    }
  }
}

struct FusedLayout {
  int64_t off_part, off_stats, off_stats_ticket, off_mc, off_zero, off_keep1, off_visited, off_loc, off_gridpt, off_comp, off_ticket, off_desc, zero_bytes,
      off_keepf, total;
};

static FusedLayout fused_layout(int64_t n) {
  FusedLayout w{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += align256(bytes); return r; };
  w.off_part = take(1024 * 4 * SFXM_MAX_NV * 8);
  w.off_stats = take(sizeof(Stats));
  w.off_stats_ticket = take(16);                           // last-block ticket of launch 1 (zeroed by a 4-byte memset)
  w.off_mc = take(sizeof(MarkConst));
  w.off_zero = o;                                          // everything from here to off_keepf is zero-filled by launch 1
  w.off_keep1 = take(n);
  w.off_visited = take(n);
  w.off_loc = take(n);
  w.off_gridpt = take(n);
  w.off_comp = take(SFXM_MAX_NV * 4);
  w.off_ticket = take(16);
  w.off_desc = take(((n + kScanTile - 1) / kScanTile) * (int64_t)sizeof(TileDesc));
  w.zero_bytes = o - w.off_zero;
  w.off_keepf = take(n);
  w.total = o;
  return w;
}

template <int NV>
static int amr_fused_t(const double *z, const double *U, int64_t n, const Criteria &cr, char *work, double *x_out, double *U_out,
                       int64_t *counts_dev, cudaStream_t st) {
  const FusedLayout w = fused_layout(n);
  const Window wd{n, 0, n, 0, n};
  const int64_t blocks_all = (n + 255) / 256;
  const int sblocks = (int)(blocks_all < 592 ? blocks_all : 592);                 // 4 x 148 SMs
  unsigned int *stats_ticket = (unsigned int *)(work + w.off_stats_ticket);
  unsigned int *tile_counter = (unsigned int *)(work + w.off_ticket);            // inside the region launch 1 zero-fills
  cudaMemsetAsync(stats_ticket, 0, sizeof(unsigned int), st);
  amr_stats_zero_kernel<NV><<<sblocks, 256, 0, st>>>(z, U, wd, (double *)(work + w.off_part), (Stats *)(work + w.off_stats),
                                                     (MarkConst *)(work + w.off_mc), cr.slope, stats_ticket,
                                                     (uint4 *)(work + w.off_zero), w.zero_bytes / 16);
  amr_mark_kernel<NV><<<(unsigned)blocks_all, 256, 0, st>>>(z, U, wd, cr, (const Stats *)(work + w.off_stats),
                                                            (uint8_t *)(work + w.off_keep1), (uint8_t *)(work + w.off_visited),
                                                            (uint8_t *)(work + w.off_loc), (int32_t *)(work + w.off_comp),
                                                            (uint8_t *)(work + w.off_gridpt), (const MarkConst *)(work + w.off_mc));
  const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
  amr_scan_apply_kernel<NV><<<tiles, kScanThreads, 0, st>>>(z, U, n, (const uint8_t *)(work + w.off_keep1),
                                                            (const uint8_t *)(work + w.off_visited), (const uint8_t *)(work + w.off_loc),
                                                            (uint8_t *)(work + w.off_keepf), (TileDesc *)(work + w.off_desc),
                                                            tile_counter, x_out, U_out, counts_dev);
  return check_launch("sfxm_refine_fused");
}

#define SFXM_NV_SWITCH(nv, CALL)                                   \
  switch (nv) {                                                    \
    case 1: { constexpr int NVV = 1; e = CALL; } break;            \
    case 2: { constexpr int NVV = 2; e = CALL; } break;            \
    case 3: { constexpr int NVV = 3; e = CALL; } break;            \
    case 4: { constexpr int NVV = 4; e = CALL; } break;            \
    default: return set_error("nv must be 1..4");                  \
  }

}  // namespace sfxm

using namespace sfxm;

extern "C" int64_t sfxm_refine_workspace_bytes(int64_t n) { return n > 0 ? amr_layout(n).total : 0; }

extern "C" int sfxm_refine_mark(const double *x, const double *U, int64_t n, int32_t nv, double ratio, double slope,
                                double curve, double prune, double min_range, double min_grid, void *work,
                                int64_t *counts_dev, void *stream) {
  if (!x || !U || !work || !counts_dev) return set_error("sfxm_refine_mark: null argument");
  if (n < 3) return set_error("sfxm_refine_mark: need at least 3 interior cells");
  if (n > 2147483647LL) return set_error("sfxm_refine_mark: mesh too large for one scan");
  const Criteria cr{ratio, slope, curve, prune, min_range, min_grid};
  const Window wd{n, 0, n, 0, n};
  cudaStream_t st = (cudaStream_t)stream;
  char *wk = (char *)work;
  double *stats = (double *)(wk + amr_layout(n).off_stats);
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_stats_t<NVV>(x, U, wd, wk, stats, st)));
  if (e) return e;
  SFXM_NV_SWITCH(nv, (amr_mark_t<NVV>(x, U, wd, cr, stats, wk, nullptr, st)));
  if (e) return e;
  return amr_finalize_t(wd, -1, wk, counts_dev, st);
}

extern "C" int64_t sfxm_refine_fused_workspace_bytes(int64_t n) { return n > 0 ? fused_layout(n).total : 0; }

extern "C" int sfxm_refine_fused(const double *x, const double *U, int64_t n, int32_t nv, double ratio, double slope,
                                 double curve, double prune, double min_range, double min_grid, void *work, double *x_out,
                                 double *U_out, int64_t *counts_dev, void *stream) {
  if (!x || !U || !work || !x_out || !U_out || !counts_dev) return set_error("sfxm_refine_fused: null argument");
  if (n < 3) return set_error("sfxm_refine_fused: need at least 3 interior cells");
  if (n > 1073741823LL) return set_error("sfxm_refine_fused: mesh too large for the packed counts");
  const Criteria cr{ratio, slope, curve, prune, min_range, min_grid};
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_fused_t<NVV>(x, U, n, cr, (char *)work, x_out, U_out, counts_dev, (cudaStream_t)stream)));
  return e;
}

extern "C" int sfxm_refine_fused_flags(const void *work, int64_t n, uint8_t *loc_out, uint8_t *keep_out, uint8_t *gridpt_out,
                                       int32_t *comp_out, void *stream) {
  if (!work || n < 1) return set_error("sfxm_refine_fused_flags: bad argument");
  const FusedLayout w = fused_layout(n);
  const char *wk = (const char *)work;
  cudaStream_t st = (cudaStream_t)stream;
  if (loc_out) cudaMemcpyAsync(loc_out, wk + w.off_loc, n, cudaMemcpyDeviceToDevice, st);
  if (keep_out) cudaMemcpyAsync(keep_out, wk + w.off_keepf, n, cudaMemcpyDeviceToDevice, st);
  if (gridpt_out) cudaMemcpyAsync(gridpt_out, wk + w.off_gridpt, n, cudaMemcpyDeviceToDevice, st);
  if (comp_out) cudaMemcpyAsync(comp_out, wk + w.off_comp, SFXM_MAX_NV * 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("sfxm_refine_fused_flags");
}

extern "C" int sfxm_refine_set_flags(const uint8_t *loc_in, const uint8_t *keep_in, int64_t n, void *work,
                                     int64_t *counts_dev, void *stream) {
  if (!loc_in || !keep_in || !work || !counts_dev) return set_error("sfxm_refine_set_flags: null argument");
  if (n < 2 || n > 2147483647LL) return set_error("sfxm_refine_set_flags: bad mesh size");
  const AmrWork w = amr_layout(n);
  const Window wd{n, 0, n, 0, n};
  cudaStream_t st = (cudaStream_t)stream;
  char *wk = (char *)work;
  amr_set_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(loc_in, keep_in, n, (uint8_t *)(wk + w.off_loc),
                                                                  (uint8_t *)(wk + w.off_keepf), (int64_t *)(wk + w.off_kcnt),
                                                                  (int64_t *)(wk + w.off_icnt));
  size_t cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(wk + w.off_cub), cb, (const int64_t *)(wk + w.off_kcnt), (int64_t *)(wk + w.off_kbef),
                                (int)n, st);
  cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(wk + w.off_cub), cb, (const int64_t *)(wk + w.off_icnt), (int64_t *)(wk + w.off_ibef),
                                (int)n, st);
  amr_count_kernel<<<1, 1, 0, st>>>((const int64_t *)(wk + w.off_kbef), (const int64_t *)(wk + w.off_ibef),
                                    (const int64_t *)(wk + w.off_kcnt), (const int64_t *)(wk + w.off_icnt), wd, counts_dev);
  return check_launch("sfxm_refine_set_flags");
}

extern "C" int sfxm_refine_flags(const void *work, int64_t n, uint8_t *loc_out, uint8_t *keep_out, uint8_t *gridpt_out,
                                 int32_t *comp_out, void *stream) {
  return sfxm_refine_window_flags(work, n, 0, n, loc_out, keep_out, gridpt_out, comp_out, stream);
}

extern "C" int sfxm_refine_apply(const double *x, const double *U, int64_t n, int32_t nv, const void *work, double *x_out,
                                 double *U_out, void *stream) {
  if (!x || !U || !work || !x_out || !U_out) return set_error("sfxm_refine_apply: null argument");
  const Window wd{n, 0, n, 0, n};
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_apply_t<NVV>(x, U, wd, (const char *)work, x_out, U_out, (cudaStream_t)stream)));
  return e;
}

// ---- the same phases on a window of a slab-decomposed mesh ----
extern "C" int sfxm_refine_window_stats(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                                        int64_t own_lo, int64_t own_hi, void *work, double *stats_dev, void *stream) {
  if (!x || !U || !work || !stats_dev) return set_error("sfxm_refine_window_stats: null argument");
  const Window wd{ne, g0, ng, own_lo, own_hi};
  if (!window_ok(wd)) return set_error("sfxm_refine_window_stats: bad window");
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_stats_t<NVV>(x, U, wd, (char *)work, stats_dev, (cudaStream_t)stream)));
  return e;
}

extern "C" int sfxm_refine_window_mark(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                                       int64_t own_lo, int64_t own_hi, double ratio, double slope, double curve,
                                       double prune, double min_range, double min_grid, const double *stats_dev,
                                       void *work, int64_t *last_run_dev, void *stream) {
  if (!x || !U || !work || !stats_dev || !last_run_dev) return set_error("sfxm_refine_window_mark: null argument");
  const Window wd{ne, g0, ng, own_lo, own_hi};
  if (!window_ok(wd)) return set_error("sfxm_refine_window_mark: bad window");
  const Criteria cr{ratio, slope, curve, prune, min_range, min_grid};
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_mark_t<NVV>(x, U, wd, cr, stats_dev, (char *)work, last_run_dev, (cudaStream_t)stream)));
  return e;
}

extern "C" int sfxm_refine_window_finalize(int64_t ne, int64_t g0, int64_t ng, int64_t own_lo, int64_t own_hi,
                                           int64_t carry_run_start, void *work, int64_t *counts_dev, void *stream) {
  if (!work || !counts_dev) return set_error("sfxm_refine_window_finalize: null argument");
  const Window wd{ne, g0, ng, own_lo, own_hi};
  if (!window_ok(wd)) return set_error("sfxm_refine_window_finalize: bad window");
  return amr_finalize_t(wd, carry_run_start, (char *)work, counts_dev, (cudaStream_t)stream);
}

extern "C" int sfxm_refine_window_apply(const double *x, const double *U, int64_t ne, int32_t nv, int64_t g0, int64_t ng,
                                        int64_t own_lo, int64_t own_hi, const void *work, double *x_out, double *U_out,
                                        void *stream) {
  if (!x || !U || !work || !x_out || !U_out) return set_error("sfxm_refine_window_apply: null argument");
  const Window wd{ne, g0, ng, own_lo, own_hi};
  if (!window_ok(wd)) return set_error("sfxm_refine_window_apply: bad window");
  int e = 0;
  SFXM_NV_SWITCH(nv, (amr_apply_t<NVV>(x, U, wd, (const char *)work, x_out, U_out, (cudaStream_t)stream)));
  return e;
}

extern "C" int sfxm_refine_window_flags(const void *work, int64_t ne, int64_t own_lo, int64_t own_hi, uint8_t *loc_out,
                                        uint8_t *keep_out, uint8_t *gridpt_out, int32_t *comp_out, void *stream) {
  if (!work || own_lo < 0 || own_hi > ne || own_lo >= own_hi) return set_error("sfxm_refine_window_flags: bad argument");
  const AmrWork w = amr_layout(ne);
  const char *wk = (const char *)work;
  const int64_t n = own_hi - own_lo;
  cudaStream_t st = (cudaStream_t)stream;
  if (loc_out) cudaMemcpyAsync(loc_out, wk + w.off_loc + own_lo, n, cudaMemcpyDeviceToDevice, st);
  if (keep_out) cudaMemcpyAsync(keep_out, wk + w.off_keepf + own_lo, n, cudaMemcpyDeviceToDevice, st);
  if (gridpt_out) cudaMemcpyAsync(gridpt_out, wk + w.off_gridpt + own_lo, n, cudaMemcpyDeviceToDevice, st);
  if (comp_out) cudaMemcpyAsync(comp_out, wk + w.off_comp, SFXM_MAX_NV * 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("sfxm_refine_flags");
}
