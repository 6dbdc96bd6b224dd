// sfxm_newton_ctl.cu - per-member control of the Newton / Split-Newton iteration, on the device.
//
// The reference hands the loop to the third-party `splitnewton` (simulation.py:515-523); this repository's specification
// of it is oracle/newton.py (parity unpinned).  Each ensemble member follows the scalar algorithm with its own iteration
// count, criterion and pseudo time step; these two kernels are the per-iteration bookkeeping of
// splitfxm_b200/newton.py (NewtonDriver._begin / _end), which was a dozen eager tensor operations per iteration:
//   begin: run = parent && crit >= 1 && it < maxiter;  shift = -1/dt;  any |= run      (one flag comes back to the host)
//   end:   crit = sqrt(ss) / (abs + rel sqrt(yy));  it += 1;  dt = min(dt |r| / |r_new|, dtmax)   for the running members
//          (ss, yy: one segment's sums, or - at an enclosing Split-Newton level - the sums of the kept per-segment sums)
// Plain IEEE operations in the order of the tensor expressions they replace (-fmad=false), so both give the same bits.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {

__global__ void newton_begin_kernel(int64_t batch, const int32_t *__restrict__ parent, double *__restrict__ crit,
                                    int64_t *__restrict__ it, double maxiter, double *__restrict__ dt, double dt0, int reset,
                                    int32_t *__restrict__ run, double *__restrict__ shift, int32_t *__restrict__ any_flag) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool r = false;
  if (m < batch) {
    if (reset) {                                         // first iteration of a loop: fresh counters, no fill launches
      crit[m] = INFINITY;
      it[m] = 0;
      if (dt) dt[m] = dt0;
    }
    r = (parent == nullptr || parent[m] != 0) && crit[m] >= 1.0 && (double)it[m] < maxiter;
    run[m] = r ? 1 : 0;
    if (shift) shift[m] = -1.0 / dt[m];
  }
  if (__syncthreads_or(r ? 1 : 0) && threadIdx.x == 0) atomicOr(any_flag, 1);
}

__global__ void newton_end_kernel(int64_t batch, const int32_t *__restrict__ run, const double *__restrict__ ss,
                                  const double *__restrict__ yy, int32_t nsum, double abs_tol, double rel_tol,
                                  double *__restrict__ crit, int64_t *__restrict__ it, double *__restrict__ dt,
                                  const double *__restrict__ r_old_sq, const double *__restrict__ r_new_sq, double dtmax,
                                  double *__restrict__ keep_ss, double *__restrict__ keep_yy) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= batch || run[m] == 0) return;
  double a = ss[m], b = yy[m];
  for (int q = 1; q < nsum; ++q) {                     // Split-Newton: the sums of the segments, in segment order
    a += ss[(int64_t)q * batch + m];
    b += yy[(int64_t)q * batch + m];
  }
  if (keep_ss) { keep_ss[m] = a; keep_yy[m] = b; }     // this segment's last step / state norms, for the enclosing level
  crit[m] = sqrt(a) / (abs_tol + rel_tol * sqrt(b));
  it[m] += 1;
  if (dt) {                                            // SER growth of the pseudo time step
    const double rn = sqrt(r_new_sq[m]);
    const double grown = dt[m] * sqrt(r_old_sq[m]) / rn;
    dt[m] = rn > 0.0 ? fmin(grown, dtmax) : dtmax;
  }
}

}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_newton_begin(int64_t batch, const int32_t *parent, double *crit, int64_t *it, double maxiter, double *dt,
                                 double dt0, int32_t reset, int32_t *run, double *shift, int32_t *any_flag, void *stream) {
  if (batch < 1 || !crit || !it || !run || !any_flag || (shift && !dt)) return set_error("sfxm_newton_begin: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(any_flag, 0, sizeof(int32_t), st);
  newton_begin_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(batch, parent, crit, it, maxiter, dt, dt0, reset, run, shift,
                                                                      any_flag);
  return check_launch("sfxm_newton_begin");
}

extern "C" int sfxm_newton_end(int64_t batch, const int32_t *run, const double *ss, const double *yy, int32_t nsum,
                               double abs_tol, double rel_tol, double *crit, int64_t *it, double *dt, const double *r_old_sq,
                               const double *r_new_sq, double dtmax, double *keep_ss, double *keep_yy, void *stream) {
  if (batch < 1 || !run || !ss || !yy || nsum < 1 || !crit || !it || (dt && (!r_old_sq || !r_new_sq)) || (!keep_ss != !keep_yy))
    return set_error("sfxm_newton_end: bad argument");
  newton_end_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(batch, run, ss, yy, nsum, abs_tol, rel_tol,
                                                                                      crit, it, dt, r_old_sq, r_new_sq, dtmax,
                                                                                      keep_ss, keep_yy);
  return check_launch("sfxm_newton_end");
}
