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
