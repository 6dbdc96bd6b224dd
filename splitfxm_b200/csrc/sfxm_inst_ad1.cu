// sfxm_inst_ad1.cu - residual / Jacobian kernels instantiated for ModelAD<1> (all schemes).
#include "sfxm_kernels.cuh"
namespace sfxm {
int launch_ad1(const Launch &L) { return launch_scheme<ModelAD<1>, 1>(L); }
}  // namespace sfxm
