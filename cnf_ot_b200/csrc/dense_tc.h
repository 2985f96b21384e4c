// Host-side interface of the tcgen05 dense-layer kernels (dense_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cnfot {

// W (K x N, row stride ldw; or its transpose) -> hi/lo tiles in MMA order, 2*K*N floats
cudaError_t dense_prep(cudaStream_t s, const float* W, int K, int N, int ldw, bool transpose, float* out);
// Y = epilogue(X * W [+ bias]); epilogue: 0 bias, 1 bias + ReLU, 2 ReLU mask by mask_src, 3 none
cudaError_t dense_forward(cudaStream_t s, const float* X, int64_t rows, int K, int ldx, const float* Bt, int N,
                          const float* bias, const float* mask_src, int ldm, int epilogue, float* Y, int ldy,
                          bool* supported);

// dW (Ka x Nb, row stride ldw) += A^T G over `rows` sample rows, db (Nb, may be NULL) += column sums of G
cudaError_t dense_wgrad(cudaStream_t s, const float* A, int lda, const float* G, int ldg, int64_t rows, int Ka, int Nb,
                        float* dW, int ldw, float* db);

}  // namespace cnfot
