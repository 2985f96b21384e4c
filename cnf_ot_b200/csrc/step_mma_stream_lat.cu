// The fused train-step kernel, streamed plan, two CTAs per SM (flow_kernels.cuh: LAT).
#define CNFOT_STEP_LAT 1
#include "step_mma_stream.cu"
