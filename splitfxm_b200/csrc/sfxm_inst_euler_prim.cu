// sfxm_inst_euler_prim.cu - residual / Jacobian kernels instantiated for ModelEulerPrim (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_euler_prim(const Launch &L) { return launch_scheme<ModelEulerPrim, 3>(L); }
}  // namespace sfxm
