// Host-side interface of the tcgen05 dense-layer kernels (dense_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cnfot {

// element (row n, k) of a [rows][16] K-major SWIZZLE_64B operand tile, in floats
__host__ __device__ __forceinline__ int sw64_pos(int n, int k) {
  return n * 16 + ((((k >> 2) ^ ((n >> 1) & 3)) << 2) | (k & 3));
}

// Row of a conditioner's input matrix W0 ((d+1) x H: row 0 = time, row 1+j = coordinate perm[j], j < d;
// cnf_ot/models/autoregressive.py:94-98,124-128) that multiplies column k of a wide-path state row
// [x_0 .. x_{D-1} | t | 0 ...], or -1 when the conditioner does not read that column.
__host__ __device__ __forceinline__ int w0_row(int k, int D, int d, int rev) {
  if (k == D) return 0;
  if (k > D) return -1;
  const int pos = rev ? D - 1 - k : k;
  return pos < d ? 1 + pos : -1;
}

// destination mapping of dense_wgrad: mode 0 = dW[m][n]; mode 1 = dW[w0_row(n)][m] (input layer of a
// conditioner, operands swapped so the wide feature axis fills the 128-row MMA tile)
struct WgradMap {
  int mode, D, d, rev;
};

// W (K x N, row stride ldw; or its transpose) -> hi/lo tiles in MMA order, 2*K*N floats
cudaError_t dense_prep(cudaStream_t s, const float* W, int K, int N, int ldw, bool transpose, float* out);
// Y = epilogue(X * W [+ bias]); epilogue: 0 bias, 1 bias + ReLU, 2 ReLU mask by mask_src, 3 none, 4 Y += X * W
cudaError_t dense_forward(cudaStream_t s, const float* X, int64_t rows, int K, int ldx, const float* Bt, int N,
                          const float* bias, const float* mask_src, int ldm, int epilogue, float* Y, int ldy,
                          bool* supported);

// dW (Ka x Nb, row stride ldw) += A^T G over `rows` sample rows; db (may be NULL) += column sums of G (Nb values;
// with map->mode == 1: of A, Ka values -- the bias of the layer whose adjoint is A there), summed on the side by the
// CTAs that gather those rows anyway
cudaError_t dense_wgrad(cudaStream_t s, const float* A, int lda, const float* G, int ldg, int64_t rows, int Ka, int Nb,
                        float* dW, int ldw, float* db, const WgradMap* map = nullptr);

}  // namespace cnfot
