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

// ---- bulk asynchronous copies (TMA, sm_90+; SASS UBLKCP): shared -> global ---------------------------------------------
// One elected thread hands a contiguous row staged in shared memory to the copy engine; the stores leave the SM without
// occupying the load/store path that the warp's own STG instructions use.  `bytes` and both addresses must be multiples
// of 16.  The generic-proxy writes that filled the row must be fenced into the async proxy first (bulk_fence_smem), and
// the CTA may not exit - or reuse the row - before bulk_wait_read().
__device__ __forceinline__ void bulk_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *gmem, const void *smem, unsigned bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace sfxm
