// sfxm_cpasync.cuh - cp.async (LDGSTS) helpers: per-thread asynchronous copies from global into shared memory, used to
// fetch the rows a sequential sweep needs several iterations ahead (sfxm_thomas.cuh, sfxm_fused.cuh, the ensemble
// residual kernel).  Every thread only ever reads what it copied itself, so cp.async.wait_group is the only ordering needed.
#pragma once
#include <cuda_runtime.h>

namespace sfxm {

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16(double2 *smem, const double2 *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

}  // namespace sfxm
