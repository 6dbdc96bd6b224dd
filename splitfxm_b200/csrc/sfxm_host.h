// sfxm_host.h - host-side helpers shared by the translation units of libsfxm_b200.so
#pragma once
#include <cuda_runtime.h>

namespace sfxm {
// stores the message for sfxm_last_error() and returns a non-zero status
int set_error(const char *msg);
// cudaGetLastError() -> status (0 ok)
int check_launch(const char *what);
}  // namespace sfxm
