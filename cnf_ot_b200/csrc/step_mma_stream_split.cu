// The fused train-step kernel with split kinetic rows (flow_kernels.cuh: SPLIT), engine "mma_stream".
#define CNFOT_STEP_SPLIT 1
#include "step_mma_stream.cu"
