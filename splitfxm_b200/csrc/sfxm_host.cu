// sfxm_host.cu - error plumbing of the C ABI (include/sfxm_b200.h)
#include <cstdio>
#include <cstring>
#include "sfxm_host.h"
#include "../../include/sfxm_b200.h"

namespace sfxm {
static thread_local char g_err[512] = "";

int set_error(const char *msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return 1;
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return 2;
}
}  // namespace sfxm

extern "C" int sfxm_abi_version(void) { return 1; }
extern "C" const char *sfxm_last_error(void) { return sfxm::g_err; }

// ---- step ordering of the fused halo exchange (include/sfxm_b200.h) ----
namespace sfxm {
__global__ void halo_signal_kernel(int64_t *left, int64_t *right, int64_t value) {
  __threadfence_system();   // the step kernel's peer stores were issued before this kernel started (stream order)
  if (left) *reinterpret_cast<volatile int64_t *>(left) = value;
  if (right) *reinterpret_cast<volatile int64_t *>(right) = value;
  __threadfence_system();
}
__global__ void halo_wait_kernel(const int64_t *from_left, const int64_t *from_right, int64_t value, int32_t *timeout_flag) {
  const long long t0 = clock64();
  const long long limit = 4000000000LL;   // ~2 s at 1.9 GHz: never hang the GPU on a lost neighbour
  bool ok = false;
  while (!ok) {
    const bool l = !from_left || *reinterpret_cast<const volatile int64_t *>(from_left) >= value;
    const bool r = !from_right || *reinterpret_cast<const volatile int64_t *>(from_right) >= value;
    ok = l && r;
    if (!ok && clock64() - t0 > limit) {
      if (timeout_flag) *timeout_flag = 1;
      break;
    }
  }
  __threadfence_system();
}
// counter variants: the step number lives in device memory, so the launches carry no per-step argument and a
// (wait, update, signal) sequence can be captured once in a CUDA graph and replayed
__global__ void halo_signal_counter_kernel(int64_t *counter, int64_t *left, int64_t *right) {
  const int64_t v = *counter + 1;
  *counter = v;
  __threadfence_system();
  if (left) *reinterpret_cast<volatile int64_t *>(left) = v;
  if (right) *reinterpret_cast<volatile int64_t *>(right) = v;
  __threadfence_system();
}
__global__ void halo_wait_counter_kernel(const int64_t *counter, const int64_t *from_left, const int64_t *from_right,
                                         int32_t *timeout_flag) {
  const int64_t value = *counter;          // steps I have completed: both neighbours must have completed as many
  const long long t0 = clock64();
  const long long limit = 4000000000LL;
  bool ok = false;
  while (!ok) {
    const bool l = !from_left || *reinterpret_cast<const volatile int64_t *>(from_left) >= value;
    const bool r = !from_right || *reinterpret_cast<const volatile int64_t *>(from_right) >= value;
    ok = l && r;
    if (!ok && clock64() - t0 > limit) {
      if (timeout_flag) *timeout_flag = 1;
      break;
    }
  }
  __threadfence_system();
}
}  // namespace sfxm

extern "C" int sfxm_halo_signal_counter(int64_t *counter, int64_t *left_peer_flag, int64_t *right_peer_flag, void *stream) {
  if (!counter) return sfxm::set_error("sfxm_halo_signal_counter: null counter");
  sfxm::halo_signal_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, left_peer_flag, right_peer_flag);
  return sfxm::check_launch("sfxm_halo_signal_counter");
}

extern "C" int sfxm_halo_wait_counter(const int64_t *counter, const int64_t *from_left_flag, const int64_t *from_right_flag,
                                      int32_t *timeout_flag, void *stream) {
  if (!counter) return sfxm::set_error("sfxm_halo_wait_counter: null counter");
  sfxm::halo_wait_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, from_left_flag, from_right_flag, timeout_flag);
  return sfxm::check_launch("sfxm_halo_wait_counter");
}

extern "C" int sfxm_halo_signal(int64_t *left_peer_flag, int64_t *right_peer_flag, int64_t value, void *stream) {
  sfxm::halo_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(left_peer_flag, right_peer_flag, value);
  return sfxm::check_launch("sfxm_halo_signal");
}

extern "C" int sfxm_halo_wait(const int64_t *from_left_flag, const int64_t *from_right_flag, int64_t value,
                              int32_t *timeout_flag, void *stream) {
  sfxm::halo_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(from_left_flag, from_right_flag, value, timeout_flag);
  return sfxm::check_launch("sfxm_halo_wait");
}
