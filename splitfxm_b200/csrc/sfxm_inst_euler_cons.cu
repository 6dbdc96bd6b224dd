// sfxm_inst_euler_cons.cu - residual / Jacobian kernels instantiated for ModelEulerCons (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_euler_cons(const Launch &L) { return launch_scheme<ModelEulerCons, 3>(L); }
}  // namespace sfxm
