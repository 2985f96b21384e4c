// Instantiations of the fused train-step kernel: the tcgen05 engine (tc_engine.cuh).
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_ENG_CASE(M_, E_) \
  if (f.M == M_) return (const void*)&mfc_step_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, E_>;

const void* find_mfc_step_kernel_tc(const FlowLayout& f) {
  if (f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngTc>;
  STEP_ENG_CASE(1, kEngTc) STEP_ENG_CASE(2, kEngTc) STEP_ENG_CASE(3, kEngTc)
  return nullptr;
}

}  // namespace cnfot
