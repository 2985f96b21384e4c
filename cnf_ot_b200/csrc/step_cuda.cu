// Instantiations of the fused train-step kernel: the CUDA-core conditioner engine (device_common.cuh).
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_CASE(H_, K_, M_)                                                            \
  if (f.H == H_ && f.K == K_ && f.M == M_)                                               \
    return (const void*)&mfc_step_kernel<NetCfg<H_, K_, M_>, Dims<0, 0>, kEngCuda>;

const void* find_mfc_step_kernel_cuda(const FlowLayout& f) {
  if (f.H == 16 && f.K == 5 && f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngCuda>;
  CNFOT_NET_LIST(STEP_CASE)
  return nullptr;
}

}  // namespace cnfot
