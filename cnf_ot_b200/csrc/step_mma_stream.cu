// Instantiations of the fused train-step kernel: the warp-level tensor-core engine streaming weights + fragments from global memory.
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_ENG_CASE(M_, E_) \
  if (f.M == M_) return (const void*)&mfc_step_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, E_>;

const void* find_mfc_step_kernel_mma_stream(const FlowLayout& f) {
  STEP_ENG_CASE(1, kEngMmaStream) STEP_ENG_CASE(2, kEngMmaStream) STEP_ENG_CASE(3, kEngMmaStream)
  return nullptr;
}

}  // namespace cnfot
