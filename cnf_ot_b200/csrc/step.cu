// Instantiations of the fused train-step kernel.
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_CASE(H_, K_, M_)                                                            \
  if (f.H == H_ && f.K == K_ && f.M == M_)                                               \
    return (const void*)&mfc_step_kernel<NetCfg<H_, K_, M_>, Dims<0, 0>, kEngCuda>;
#define STEP_ENG_CASE(M_, E_)                                                            \
  if (f.M == M_) return (const void*)&mfc_step_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, E_>;

const void* find_mfc_step_kernel(const FlowLayout& f, int engine) {
  if (engine == kEngTc && tc_available(f)) {
    if (f.M == 2 && f.D == 2 && f.L == 2)
      return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngTc>;
    STEP_ENG_CASE(1, kEngTc) STEP_ENG_CASE(2, kEngTc) STEP_ENG_CASE(3, kEngTc)
    return nullptr;
  }
  if (engine == kEngMma && tc_available(f)) {
    if (f.M == 2 && f.D == 2 && f.L == 2)
      return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngMma>;
    STEP_ENG_CASE(1, kEngMma) STEP_ENG_CASE(2, kEngMma) STEP_ENG_CASE(3, kEngMma)
    return nullptr;
  }
  if (engine == kEngMmaStream && tc_available(f)) {
    STEP_ENG_CASE(1, kEngMmaStream) STEP_ENG_CASE(2, kEngMmaStream) STEP_ENG_CASE(3, kEngMmaStream)
    return nullptr;
  }
  if (f.H == 16 && f.K == 5 && f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&mfc_step_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, kEngCuda>;
  CNFOT_NET_LIST(STEP_CASE)
  return nullptr;
}

}  // namespace cnfot
