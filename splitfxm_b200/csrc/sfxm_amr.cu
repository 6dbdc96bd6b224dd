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
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {

static constexpr double kUlp = 2.220446049250313e-16;   // math.ulp(1.0), refine.py:7

struct Stats {   // per variable: min/max of values and of slopes (refine.py:146-153)
  double vmin[SFXM_MAX_NV], vmax[SFXM_MAX_NV], smin[SFXM_MAX_NV], smax[SFXM_MAX_NV];
};

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
amr_stats_kernel(const double *__restrict__ z, const double *__restrict__ U, int64_t n, double *__restrict__ part) {
  double lo_v[NV], hi_v[NV], lo_s[NV], hi_s[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { lo_v[i] = lo_s[i] = INFINITY; hi_v[i] = hi_s[i] = -INFINITY; }
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const bool has_s = j + 1 < n;
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

// flags: keep1[j] / visited[j] / loc[j] (bytes, zero-initialised), comp[i] = component i asked for a point,
// gridpt[j] = the grid-ratio rule named "point j" (refine.py:203,211)
template <int NV>
__global__ void __launch_bounds__(256)
amr_mark_kernel(const double *__restrict__ z, const double *__restrict__ U, int64_t n, const Criteria cr,
                const Stats *__restrict__ st, uint8_t *__restrict__ keep1, uint8_t *__restrict__ visited,
                uint8_t *__restrict__ loc, int32_t *__restrict__ comp, uint8_t *__restrict__ gridpt) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (j == 0 || j == n - 1) keep1[j] = 1;                              // refine.py:120-121
  const double zj = __ldg(z + j);
  const double dzj = j + 1 < n ? __ldg(z + j + 1) - zj : 0.0;          // dz[j]
  const double dzj1 = j + 2 < n ? __ldg(z + j + 2) - __ldg(z + j + 1) : 0.0;   // dz[j+1]
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double vmin = st->vmin[i], vmax = st->vmax[i], smin = st->smin[i], smax = st->smax[i];
    const double aa = fmax(fabs(vmin), fabs(vmax));
    const double ss = fmax(fabs(smin), fabs(smax));
    const double v0 = __ldg(U + j * NV + i);
    const double v1 = j + 1 < n ? __ldg(U + (j + 1) * NV + i) : 0.0;
    // slope criterion on interval j (refine.py:159-173)
    if (j + 1 < n && (vmax - vmin) > cr.min_range * aa) {
      const double dmax = cr.slope * (vmax - vmin) + kUlp;
      const double r = fabs(v1 - v0) / dmax;
      if (r > 1.0 && dzj >= 2 * cr.min_grid) { loc[j] = 1; comp[i] = 1; }
      if (r >= cr.prune) { keep1[j] = 1; keep1[j + 1] = 1; }
      else visited[j] = 1;
    }
    // curvature criterion on the pair of intervals (j, j+1) (refine.py:179-197)
    if (j + 2 < n && (smax - smin) > cr.min_range * ss) {
      const double v2 = __ldg(U + (j + 2) * NV + i);
      const double s0 = (v1 - v0) / dzj;
      const double s1 = (v2 - v1) / dzj1;
      const double dmax = cr.curve * (smax - smin);
      const double r = fabs(s1 - s0) / (dmax + (kUlp / dzj));
      if (r > 1.0 && dzj >= 2 * cr.min_grid && dzj1 >= 2 * cr.min_grid) { comp[i] = 1; loc[j] = 1; loc[j + 1] = 1; }
      if (r >= cr.prune) keep1[j + 1] = 1;
      else visited[j + 1] = 1;
    }
  }
  // grid-ratio rules (refine.py:200-227), j = 1 .. n-2
  if (j >= 1 && j <= n - 2) {
    const double dzm1 = zj - __ldg(z + j - 1);                         // dz[j-1]
    if (dzj > cr.ratio * dzm1) {
      loc[j] = 1; gridpt[j] = 1;
      keep1[j - 1] = 1; keep1[j] = 1; keep1[j + 1] = 1;
      if (j + 2 < n) keep1[j + 2] = 1;
    }
    if (dzj < dzm1 / cr.ratio) {
      loc[j - 1] = 1; gridpt[j - 1] = 1;
      if (j >= 2) keep1[j - 2] = 1;
      keep1[j - 1] = 1; keep1[j] = 1; keep1[j + 1] = 1;
    }
    const double span = __ldg(z + j + 1) - __ldg(z + j - 1);
    if (j > 1 && span > cr.ratio * (__ldg(z + j - 1) - __ldg(z + j - 2))) keep1[j] = 1;
    if (j < n - 2 && span > cr.ratio * dzj1) keep1[j] = 1;
  }
}

// start[j] = j if j begins a maximal run of deletions, else -1
__global__ void amr_run_start_kernel(const uint8_t *__restrict__ keep1, const uint8_t *__restrict__ visited, int64_t n,
                                     int64_t *__restrict__ start) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const bool del = visited[j] && !keep1[j];
  const bool prev = j > 0 && visited[j - 1] && !keep1[j - 1];
  start[j] = (del && !prev) ? j : -1;
}

struct MaxOp {
  __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; }
};

// final flags: keepf[j] = 0 only for deletions at even offsets of their run (refine.py:231-233); counts for the scans
__global__ void amr_finalize_kernel(const uint8_t *__restrict__ keep1, const uint8_t *__restrict__ visited,
                                    const int64_t *__restrict__ run_start, const uint8_t *__restrict__ loc, int64_t n,
                                    uint8_t *__restrict__ keepf, int64_t *__restrict__ kept_cnt, int64_t *__restrict__ ins_cnt) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const bool del0 = visited[j] && !keep1[j];
  const bool del = del0 && (((j - run_start[j]) & 1) == 0);
  keepf[j] = del ? 0 : 1;
  kept_cnt[j] = del ? 0 : 1;
  ins_cnt[j] = (loc[j] && j + 1 < n) ? 1 : 0;
}

template <int NV>
__global__ void amr_apply_kernel(const double *__restrict__ z, const double *__restrict__ U, int64_t n,
                                 const uint8_t *__restrict__ keepf, const uint8_t *__restrict__ loc,
                                 const int64_t *__restrict__ kept_before, const int64_t *__restrict__ ins_before,
                                 double *__restrict__ z_out, double *__restrict__ U_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t kb = kept_before[j], ib = ins_before[j];
  if (keepf[j]) {
    const int64_t o = kb + ib;
    z_out[o] = __ldg(z + j);
#pragma unroll
    for (int i = 0; i < NV; ++i) U_out[o * NV + i] = __ldg(U + j * NV + i);
  }
  if (loc[j] && j + 1 < n) {
    const int64_t o = kb + (keepf[j] ? 1 : 0) + ib;
    z_out[o] = 0.5 * (__ldg(z + j + 1) + __ldg(z + j));                               // refine.py:268
#pragma unroll
    for (int i = 0; i < NV; ++i) U_out[o * NV + i] = 0.5 * (__ldg(U + (j + 1) * NV + i) + __ldg(U + j * NV + i));   // :269
  }
}

__global__ void amr_count_kernel(const int64_t *__restrict__ kept_before, const int64_t *__restrict__ ins_before,
                                 const uint8_t *__restrict__ keepf, const uint8_t *__restrict__ loc, int64_t n,
                                 int64_t *__restrict__ counts) {
  // counts[0] = new size, counts[1] = inserted, counts[2] = deleted
  const int64_t last = n - 1;
  const int64_t kept = kept_before[last] + (keepf[last] ? 1 : 0);
  const int64_t ins = ins_before[last];   // loc[n-1] is never set: points are inserted after j <= n-2 only
  counts[0] = kept + ins;
  counts[1] = ins;
  counts[2] = n - kept;
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

template <int NV>
static int refine_mark_t(const double *z, const double *U, int64_t n, const Criteria &cr, char *work, cudaStream_t st) {
  const AmrWork w = amr_layout(n);
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  const int sblocks = (int)(blocks < 1024 ? blocks : 1024);
  cudaMemsetAsync(work + w.off_keep1, 0, w.off_comp + align256(SFXM_MAX_NV * 4) - w.off_keep1, st);
  amr_stats_kernel<NV><<<sblocks, 256, 0, st>>>(z, U, n, (double *)(work + w.off_part));
  amr_stats_final_kernel<NV><<<1, 32, 0, st>>>((const double *)(work + w.off_part), sblocks, (Stats *)(work + w.off_stats));
  amr_mark_kernel<NV><<<blocks, threads, 0, st>>>(z, U, n, cr, (const Stats *)(work + w.off_stats),
                                                  (uint8_t *)(work + w.off_keep1), (uint8_t *)(work + w.off_visited),
                                                  (uint8_t *)(work + w.off_loc), (int32_t *)(work + w.off_comp),
                                                  (uint8_t *)(work + w.off_gridpt));
  amr_run_start_kernel<<<blocks, threads, 0, st>>>((const uint8_t *)(work + w.off_keep1),
                                                   (const uint8_t *)(work + w.off_visited), n, (int64_t *)(work + w.off_start));
  size_t cb = (size_t)w.cub_bytes;
  cub::DeviceScan::InclusiveScan((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_start),
                                 (int64_t *)(work + w.off_run), MaxOp(), (int)n, st);
  amr_finalize_kernel<<<blocks, threads, 0, st>>>((const uint8_t *)(work + w.off_keep1), (const uint8_t *)(work + w.off_visited),
                                                  (const int64_t *)(work + w.off_run), (const uint8_t *)(work + w.off_loc), n,
                                                  (uint8_t *)(work + w.off_keepf), (int64_t *)(work + w.off_kcnt),
                                                  (int64_t *)(work + w.off_icnt));
  cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_kcnt),
                                (int64_t *)(work + w.off_kbef), (int)n, st);
  cb = (size_t)w.cub_bytes;
  cub::DeviceScan::ExclusiveSum((void *)(work + w.off_cub), cb, (const int64_t *)(work + w.off_icnt),
                                (int64_t *)(work + w.off_ibef), (int)n, st);
  return check_launch("sfxm_refine_mark");
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
  Criteria cr{ratio, slope, curve, prune, min_range, min_grid};
  cudaStream_t st = (cudaStream_t)stream;
  int e;
  switch (nv) {
    case 1: e = refine_mark_t<1>(x, U, n, cr, (char *)work, st); break;
    case 2: e = refine_mark_t<2>(x, U, n, cr, (char *)work, st); break;
    case 3: e = refine_mark_t<3>(x, U, n, cr, (char *)work, st); break;
    case 4: e = refine_mark_t<4>(x, U, n, cr, (char *)work, st); break;
    default: return set_error("nv must be 1..4");
  }
  if (e) return e;
  const AmrWork w = amr_layout(n);
  char *wk = (char *)work;
  amr_count_kernel<<<1, 1, 0, st>>>((const int64_t *)(wk + w.off_kbef), (const int64_t *)(wk + w.off_ibef),
                                    (const uint8_t *)(wk + w.off_keepf), (const uint8_t *)(wk + w.off_loc), n, counts_dev);
  return check_launch("sfxm_refine_mark");
}

extern "C" int sfxm_refine_flags(const void *work, int64_t n, uint8_t *loc_out, uint8_t *keep_out, uint8_t *gridpt_out,
                                 int32_t *comp_out, void *stream) {
  if (!work) return set_error("sfxm_refine_flags: null argument");
  const AmrWork w = amr_layout(n);
  const char *wk = (const char *)work;
  cudaStream_t st = (cudaStream_t)stream;
  if (loc_out) cudaMemcpyAsync(loc_out, wk + w.off_loc, n, cudaMemcpyDeviceToDevice, st);
  if (keep_out) cudaMemcpyAsync(keep_out, wk + w.off_keepf, n, cudaMemcpyDeviceToDevice, st);
  if (gridpt_out) cudaMemcpyAsync(gridpt_out, wk + w.off_gridpt, n, cudaMemcpyDeviceToDevice, st);
  if (comp_out) cudaMemcpyAsync(comp_out, wk + w.off_comp, SFXM_MAX_NV * 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("sfxm_refine_flags");
}

extern "C" int sfxm_refine_apply(const double *x, const double *U, int64_t n, int32_t nv, const void *work, double *x_out,
                                 double *U_out, void *stream) {
  if (!x || !U || !work || !x_out || !U_out) return set_error("sfxm_refine_apply: null argument");
  const AmrWork w = amr_layout(n);
  const char *wk = (const char *)work;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
#define SFXM_APPLY(NV)                                                                                                   \
  amr_apply_kernel<NV><<<blocks, threads, 0, st>>>(x, U, n, (const uint8_t *)(wk + w.off_keepf),                         \
                                                   (const uint8_t *)(wk + w.off_loc), (const int64_t *)(wk + w.off_kbef), \
                                                   (const int64_t *)(wk + w.off_ibef), x_out, U_out)
  switch (nv) {
    case 1: SFXM_APPLY(1); break;
    case 2: SFXM_APPLY(2); break;
    case 3: SFXM_APPLY(3); break;
    case 4: SFXM_APPLY(4); break;
    default: return set_error("nv must be 1..4");
  }
#undef SFXM_APPLY
  return check_launch("sfxm_refine_apply");
}
