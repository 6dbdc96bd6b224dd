// sfxm_inst_ad2.cu - residual / Jacobian kernels instantiated for ModelAD<2> (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_ad2(const Launch &L) { return launch_scheme<ModelAD<2>, 2>(L); }
}  // namespace sfxm
