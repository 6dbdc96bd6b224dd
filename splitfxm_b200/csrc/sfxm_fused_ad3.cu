// sfxm_fused_ad3.cu - fused Jacobian + block-Thomas kernels instantiated for ModelAD<3> (the three-variable
// advection-diffusion ensemble of BASELINE.json) on the 3-cell schemes.
#include "sfxm_fused.cuh"
namespace sfxm {
int fused_ad3(const FusedLaunch &L) {
  switch (L.P->scheme) {
    case SFXM_FV_LAX_FRIEDRICHS: return fused_launch_scheme<ModelAD<3>, SFXM_FV_LAX_FRIEDRICHS, 3>(L);
    case SFXM_FV_CENTRAL: return fused_launch_scheme<ModelAD<3>, SFXM_FV_CENTRAL, 3>(L);
    case SFXM_FD_CENTRAL: return fused_launch_scheme<ModelAD<3>, SFXM_FD_CENTRAL, 3>(L);
  }
  return -2;   // not instantiated: the caller uses the unfused path
}
}  // namespace sfxm
