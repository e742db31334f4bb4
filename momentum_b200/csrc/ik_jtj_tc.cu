#include "ik_jtj_tc.cuh"

namespace mb2 {

bool jtjTensorSupported(int, int) { return false; }
cudaError_t launchJtJTensor(const JtJArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }

} // namespace mb2
