// SPDX-License-Identifier: Apache-2.0
// Host-buffer (zero-copy, shared-memory tile) instantiations of the env-step kernel (TILE=1), see kernel_common.cuh.
#include "step_kernel.cuh"

namespace upkie_b200 {
cudaError_t launch_step_host(const StepArgs& a) { return launch_step_kernels<1>(a); }
}  // namespace upkie_b200
