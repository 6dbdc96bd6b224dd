// sfxm_inst_ad4.cu - residual / Jacobian kernels instantiated for ModelAD<4> (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_ad4(const Launch &L) { return launch_scheme<ModelAD<4>, 4>(L); }
}  // namespace sfxm
