// Sample-prediction network on tcgen05 tensor cores (HR_MLP_BF16X3_TC) -- under construction.
#include "hr_handle.h"

namespace hr {

int pack_mlp_tc(hr_handle*, const hr_params*, const float* const*, const float* const*, cudaStream_t) {
  return hr_fail("HR_MLP_BF16X3_TC: tensor-core sample net not built into this library yet");
}

cudaError_t launch_mlp_tc(const hr_config&, const MlpTcPack&, const float*, float*, long long, int, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace hr
