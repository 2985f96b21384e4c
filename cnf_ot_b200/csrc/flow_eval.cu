// Instantiations of the forward-only and VJP flow kernels.
#include "dispatch.h"
#include "flow_kernels.cuh"

namespace cnfot {

#define EVAL_CASE(H_, K_, M_)                                                            \
  if (f.H == H_ && f.K == K_ && f.M == M_)                                               \
    return (const void*)&flow_eval_kernel<NetCfg<H_, K_, M_>, Dims<0, 0>, false>;
#define VJP_CASE(H_, K_, M_)                                                             \
  if (f.H == H_ && f.K == K_ && f.M == M_)                                               \
    return (const void*)&flow_vjp_kernel<NetCfg<H_, K_, M_>, Dims<0, 0>, false>;
#define EVAL_TC_CASE(M_)                                                                 \
  if (f.M == M_) return (const void*)&flow_eval_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, true>;
#define VJP_TC_CASE(M_)                                                                  \
  if (f.M == M_) return (const void*)&flow_vjp_kernel<NetCfg<16, 5, M_>, Dims<0, 0>, true>;

const void* find_flow_eval_kernel(const FlowLayout& f, bool tc) {
  if (tc && tc_available(f)) {
    EVAL_TC_CASE(1) EVAL_TC_CASE(2) EVAL_TC_CASE(3)
    return nullptr;
  }
  if (f.H == 16 && f.K == 5 && f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&flow_eval_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, false>;
  CNFOT_NET_LIST(EVAL_CASE)
  return nullptr;
}

const void* find_flow_vjp_kernel(const FlowLayout& f, bool tc) {
  if (tc && tc_available(f)) {
    VJP_TC_CASE(1) VJP_TC_CASE(2) VJP_TC_CASE(3)
    return nullptr;
  }
  if (f.H == 16 && f.K == 5 && f.M == 2 && f.D == 2 && f.L == 2)
    return (const void*)&flow_vjp_kernel<NetCfg<16, 5, 2>, Dims<2, 2>, false>;
  CNFOT_NET_LIST(VJP_CASE)
  return nullptr;
}

}  // namespace cnfot
