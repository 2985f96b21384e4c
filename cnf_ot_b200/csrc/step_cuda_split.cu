// The fused train-step kernel with split kinetic rows (flow_kernels.cuh: SPLIT), engine "cuda".
#define CNFOT_STEP_SPLIT 1
#include "step_cuda.cu"
