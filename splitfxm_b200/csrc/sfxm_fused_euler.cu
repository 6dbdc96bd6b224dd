// sfxm_fused_euler.cu - fused Jacobian + block-Thomas kernels for the Euler models (Lax-Friedrichs, the scheme of the
// reference's Sod tutorials).
#include "sfxm_fused.cuh"
namespace sfxm {
int fused_euler_cons(const FusedLaunch &L) {
  if (L.P->scheme == SFXM_FV_LAX_FRIEDRICHS) return fused_launch_scheme<ModelEulerCons, SFXM_FV_LAX_FRIEDRICHS, 3>(L);
  return -2;
}
int fused_euler_prim(const FusedLaunch &L) {
  if (L.P->scheme == SFXM_FV_LAX_FRIEDRICHS) return fused_launch_scheme<ModelEulerPrim, SFXM_FV_LAX_FRIEDRICHS, 3>(L);
  return -2;
}
}  // namespace sfxm
