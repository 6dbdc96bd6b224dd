// sfxm_inst_ad3.cu - residual / Jacobian kernels instantiated for ModelAD<3> (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_ad3(const Launch &L) { return launch_scheme<ModelAD<3>, 3>(L); }
}  // namespace sfxm
