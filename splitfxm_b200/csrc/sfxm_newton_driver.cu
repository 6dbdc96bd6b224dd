// sfxm_newton_driver.cu - the Newton / Split-Newton loop nest of Simulation.steady_state on the host side of the
// library (simulation.py:474-526 hands it to the third-party splitnewton; specification: oracle/newton.py, parity
// unpinned).  Native twin of splitfxm_b200/newton.py (NewtonDriver.newton_segment / split_newton) for the plain
// configuration - full steps, no pseudo-transient term, no Armijo back-tracking, no bounds: the same C-ABI calls in the
// same order (begin, [Jacobian of the segment, block solve, update + norms | fused sweep], end), so iterates and
// iteration counts are bit-identical to the Python driver's.  What it removes is the interpreter between the launches:
// a Newton iteration of one small mesh is ~7 kernels of a few microseconds each, which ~10 ctypes calls spread over
// ~90 us; from here they go out back to back (steady state at nx = 250: 1.0 ms -> see profiles/README.md).
//
// The "does any member still run" flag of every loop level travels to pinned host memory behind an event; as in the
// Python driver the iteration's kernels - all masked by the per-member run flags - are launched before the host looks
// at the flag, so the device never waits for the host.
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <map>
#include <array>
#include <mutex>
#include "../../include/sfxm_b200.h"
#include "sfxm_host.h"

namespace sfxm {
namespace {

constexpr int kMaxLevels = 2 * SFXM_MAX_NV;

struct Run {
  const sfxm_newton_plan_t &p;
  cudaStream_t st;
  cudaEvent_t *ev;
  int64_t launches;
  int K;
  int64_t B;

  int64_t *it(int l) const { return p.it + (int64_t)l * B; }
  double *crit(int l) const { return p.crit + (int64_t)l * B; }
  int32_t *run(int l) const { return p.run + (int64_t)l * B; }

  // begin + flag to the level's pinned word behind its event
  int begin(int l, const int32_t *parent, double maxiter, bool reset) {
    if (int e = sfxm_newton_begin(B, parent, crit(l), it(l), maxiter, nullptr, 0.0, reset ? 1 : 0, run(l), nullptr, p.any_dev, st))
      return e;
    ++launches;
    if (cudaMemcpyAsync(const_cast<int32_t *>(p.any_host) + l, p.any_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(ev[l], st) != cudaSuccess)
      return set_error("sfxm_newton_run: flag copy failed");
    return 0;
  }
  // 1 / 0 when the flag is already known, -1 while it is in flight
  int peek(int l) const { return cudaEventQuery(ev[l]) == cudaSuccess ? (p.any_host[l] != 0) : -1; }
  int wait(int l) const {
    cudaEventSynchronize(ev[l]);
    return p.any_host[l] != 0;
  }

  // Newton on ONE segment (NewtonDriver.newton_segment)
  int segment(int k, double maxiter, const int32_t *parent) {
    const int l = k;
    bool reset = true;
    for (;;) {
      if (int e = begin(l, parent, maxiter, reset)) return e;
      reset = false;
      const int known = peek(l);
      if (known == 0) break;
      // everything below is masked by run(l): launched before the flag is known, a no-op if nobody runs
      const double *ss, *yy;
      if (p.solver == SFXM_NEWTON_FUSED_STEP) {
        if (int e = sfxm_newton_step_fused(p.prob, p.x_int, p.params, p.param_stride, p.geom, p.y, p.eps, nullptr, run(l), k,
                                           p.s, nullptr, p.work_solve, p.status, p.norms2, st))
          return e;
        launches += 1;
      } else {
        if (int e = sfxm_jacobian_segments(p.prob, p.x_int, p.params, p.param_stride, p.geom, p.y, p.eps, k, k + 1, p.J, p.r, st))
          return e;
        int e = 0;
        if (p.solver == SFXM_NEWTON_THOMAS)
          e = sfxm_block_tridiag_solve(p.prob, p.J, p.r, nullptr, run(l), k, k + 1, p.s, p.work_solve, p.status, st);
        else if (p.solver == SFXM_NEWTON_PCR)
          e = sfxm_block_tridiag_solve_pcr(p.prob, p.J, p.r, nullptr, run(l), k, k + 1, p.s, p.work_solve, p.status, st);
        else
          e = sfxm_block_band4_solve_pcr(p.prob, p.J, p.r, nullptr, run(l), k, k + 1, p.s, p.work_solve, p.status, st);
        if (e) return e;
        if (int e2 = sfxm_update_norms(p.y, p.s, nullptr, run(l), 1, p.row_lo[k], p.row_hi[k], B, p.norms2, p.work_norms, st))
          return e2;
        launches += 1 + p.solve_launches[k] + 2;
      }
      ss = p.norms2;
      yy = p.norms2 + B;
      if (int e = sfxm_newton_end(B, run(l), ss, yy, 1, p.abs_tol, p.rel_tol, crit(l), it(l), nullptr, nullptr, nullptr, 0.0,
                                  p.seg_ss + (int64_t)k * B, p.seg_yy + (int64_t)k * B, st))
        return e;
      ++launches;
      if (maxiter == 1.0) break;          // one step (the outer segments of Split-Newton): nobody can run a second time
      if (known < 0 && !wait(l)) break;
    }
    return 0;
  }

  // Split-Newton over segments k0 .. K-1 (NewtonDriver.split_newton)
  int split(int k0, double maxiter, const int32_t *parent) {
    if (k0 == K - 1) return segment(k0, maxiter, parent);
    const int l = K + k0;
    bool reset = true;
    for (;;) {
      if (int e = begin(l, parent, maxiter, reset)) return e;
      reset = false;
      const int known = peek(l);
      if (known == 0) break;
      if (int e = split(k0 + 1, maxiter, run(l))) return e;      // inner segments to convergence
      if (int e = segment(k0, 1.0, run(l))) return e;            // one step on this segment
      // |s| and |y| over segments k0.. from the sums each segment kept of its last step
      if (int e = sfxm_newton_end(B, run(l), p.seg_ss + (int64_t)k0 * B, p.seg_yy + (int64_t)k0 * B, K - k0, p.abs_tol, p.rel_tol,
                                  crit(l), it(l), nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, st))
        return e;
      ++launches;
      if (known < 0 && !wait(l)) break;
    }
    return 0;
  }
};

// events are cheap to keep: one set per device, created on first use
cudaEvent_t *level_events() {
  static std::mutex mu;
  static std::map<int, std::array<cudaEvent_t, kMaxLevels>> pool;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto hit = pool.find(dev);
  if (hit == pool.end()) {
    std::array<cudaEvent_t, kMaxLevels> evs;
    for (auto &e : evs)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    hit = pool.emplace(dev, evs).first;
  }
  return hit->second.data();
}

}  // namespace
}  // namespace sfxm

using namespace sfxm;

extern "C" int sfxm_newton_run(const sfxm_newton_plan_t *plan, double maxiter, int32_t *result_level, int64_t *launches,
                               void *stream) {
  if (!plan || !plan->prob || !plan->x_int || !plan->params || !plan->y || !plan->s || !plan->status || !plan->it ||
      !plan->crit || !plan->run || !plan->seg_ss || !plan->seg_yy || !plan->norms2 || !plan->any_dev || !plan->any_host ||
      !plan->work_solve)
    return set_error("sfxm_newton_run: null argument");
  const sfxm_newton_plan_t &p = *plan;
  if (p.nseg < 1 || p.nseg > SFXM_MAX_NV) return set_error("sfxm_newton_run: bad segment count");
  if (p.solver != SFXM_NEWTON_FUSED_STEP && (!p.J || !p.r || !p.work_norms))
    return set_error("sfxm_newton_run: the unfused path needs J, r and the norm workspace");
  if (!(maxiter >= 0.0)) return set_error("sfxm_newton_run: bad maxiter");
  cudaEvent_t *ev = level_events();
  if (!ev) return set_error("sfxm_newton_run: cannot create events");
  Run r{p, (cudaStream_t)stream, ev, 0, p.nseg, p.prob->batch};
  const int e = r.split(0, std::fmin(maxiter, 1e300), nullptr);
  if (result_level) *result_level = p.nseg == 1 ? 0 : p.nseg;      // the outermost loop's counters
  if (launches) *launches = r.launches;
  return e;
}
