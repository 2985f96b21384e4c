// Instantiations of the fused train-step kernel: the CUDA-core conditioner engine (device_common.cuh).
// CNFOT_STEP_SPLIT (step_cuda_split.cu): the instantiations whose kinetic rows are spread over lane groups.
#ifndef CNFOT_STEP_SPLIT
#define CNFOT_STEP_SPLIT 0
#endif
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_CASE(H_, K_, M_)                                                            \
  if (f.H == H_ && f.K == K_ && f.M == M_)                                               \
    return (const void*)&mfc_step_kernel<NetCfg<H_, K_, M_>, Dims<0, 0>, kEngCuda, CNFOT_STEP_SPLIT != 0>;

#if CNFOT_STEP_SPLIT
const void* find_mfc_step_kernel_cuda_split(const FlowLayout& f) {
#else
const void* find_mfc_step_kernel_cuda(const FlowLayout& f) {
#endif
  if (f.H == 16 && f.K == 5 && f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngCuda, CNFOT_STEP_SPLIT != 0>;
  CNFOT_NET_LIST(STEP_CASE)
  return nullptr;
}

}  // namespace cnfot
