// Instantiations of the fused train-step kernel: the warp-level tensor-core engine streaming weights + fragments from global memory.
// CNFOT_STEP_SPLIT (step_mma_stream_split.cu): the instantiations whose kinetic rows are spread over lane groups;
// CNFOT_STEP_LAT (step_mma_stream_lat.cu): two CTAs per SM instead of four, for steps of few tile rounds.
#ifndef CNFOT_STEP_SPLIT
#define CNFOT_STEP_SPLIT 0
#endif
#ifndef CNFOT_STEP_LAT
#define CNFOT_STEP_LAT 0
#endif
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define STEP_ENG_CASE(M_, E_) \
  if (f.M == M_) return (const void*)&mfc_step_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, E_, CNFOT_STEP_SPLIT != 0, CNFOT_STEP_LAT != 0>;

#if CNFOT_STEP_SPLIT
const void* find_mfc_step_kernel_mma_stream_split(const FlowLayout& f) {
#elif CNFOT_STEP_LAT
const void* find_mfc_step_kernel_mma_stream_lat(const FlowLayout& f) {
#else
const void* find_mfc_step_kernel_mma_stream(const FlowLayout& f) {
#endif
  STEP_ENG_CASE(1, kEngMmaStream) STEP_ENG_CASE(2, kEngMmaStream) STEP_ENG_CASE(3, kEngMmaStream)
  return nullptr;
}

}  // namespace cnfot
