// Micro-benchmark (development tool): issue rate and dependent latency of the warp-level
// tensor-core instruction HMMA.1688.F32.TF32 (mma.sync.m16n8k8 tf32) on sm_100a, next to FFMA.
// Also checks the fragment layout + the "k-permutation" chaining trick used by warp_mlp.cuh:
// a C fragment fed straight back as the A fragment of the next layer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// NCHAIN independent accumulator chains per warp, ITER mmas per chain
template <int NCHAIN>
__global__ void mma_rate(float* out, int iters) {
  float d[NCHAIN][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
  for (int c = 0; c < NCHAIN; ++c)
    for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) mma_tf32(d[c], a, b);
  }
  float s = 0.f;
  for (int c = 0; c < NCHAIN; ++c)
    for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCHAIN>
__global__ void ffma_rate(float* out, int iters) {
  float d[NCHAIN];
  float a = 1.0f + threadIdx.x * 1e-6f, b = 1e-3f;
  for (int c = 0; c < NCHAIN; ++c) d[c] = c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) d[c] = fmaf(d[c], a, b);
  }
  float s = 0.f;
  for (int c = 0; c < NCHAIN; ++c) s += d[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: Y = relu(X W1) W2 for 16 rows x 8 -> 8 -> 8, chaining C->A with the permuted k order
__global__ void chain_check(const float* X, const float* W1, const float* W2, float* Y) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  uint32_t a[4], b[2];
  a[0] = __float_as_uint(X[g * 8 + t]);
  a[1] = __float_as_uint(X[(g + 8) * 8 + t]);
  a[2] = __float_as_uint(X[g * 8 + t + 4]);
  a[3] = __float_as_uint(X[(g + 8) * 8 + t + 4]);
  b[0] = __float_as_uint(W1[t * 8 + g]);        // B[k][n] = W1[k][n], row-major (k, n)
  b[1] = __float_as_uint(W1[(t + 4) * 8 + g]);
  float c[4] = {0, 0, 0, 0};
  mma_tf32(c, a, b);
  // c0: (g, 2t) c1: (g, 2t+1) c2: (g+8, 2t) c3: (g+8, 2t+1); next layer: logical k=t := col 2t, k=t+4 := col 2t+1
  uint32_t a2[4] = {__float_as_uint(fmaxf(c[0], 0.f)), __float_as_uint(fmaxf(c[2], 0.f)),
                    __float_as_uint(fmaxf(c[1], 0.f)), __float_as_uint(fmaxf(c[3], 0.f))};
  uint32_t b2[2] = {__float_as_uint(W2[(2 * t) * 8 + g]), __float_as_uint(W2[(2 * t + 1) * 8 + g])};
  float y[4] = {0, 0, 0, 0};
  mma_tf32(y, a2, b2);
  Y[g * 8 + 2 * t] = y[0];
  Y[g * 8 + 2 * t + 1] = y[1];
  Y[(g + 8) * 8 + 2 * t] = y[2];
  Y[(g + 8) * 8 + 2 * t + 1] = y[3];
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, p.multiProcessorCount, clk_khz / 1e3);
  float* out;
  CK(cudaMalloc(&out, 148 * 32 * 1024 * sizeof(float)));
  const int iters = 4096;
  const int sms = p.multiProcessorCount;
  for (int wps : {4, 8, 16, 32}) {   // warps per SM (one CTA per SM of wps warps)
    float ms1 = time_ms([&] { mma_rate<1><<<sms, wps * 32>>>(out, iters); });
    float ms4 = time_ms([&] { mma_rate<4><<<sms, wps * 32>>>(out, iters); });
    float ms8 = time_ms([&] { mma_rate<8><<<sms, wps * 32>>>(out, iters); });
    float f8 = time_ms([&] { ffma_rate<8><<<sms, wps * 32>>>(out, iters); });
    // cycles per mma per SM sub-partition (4 per SM), assuming 1965 MHz
    auto cyc = [&](float ms, int nchain) { return ms * 1e-3 * 1.965e9 / ((double)iters * nchain * wps / 4.0); };
    printf("warps/SM %2d: mma chain1 %.2f cyc/mma/SMSP (lat), chain4 %.2f, chain8 %.2f | ffma chain8 %.2f cyc/ffma/SMSP\n",
           wps, cyc(ms1, 1), cyc(ms4, 4), cyc(ms8, 8), cyc(f8, 8));
  }
  // layout check
  std::vector<float> X(16 * 8), W1(64), W2(64), Y(16 * 8), R(16 * 8);
  srand(1);
  auto rnd = [] { float v = (rand() % 2001 - 1000) / 1000.f; uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; };
  for (auto& v : X) v = rnd();
  for (auto& v : W1) v = rnd();
  for (auto& v : W2) v = rnd();
  for (int r = 0; r < 16; ++r) {
    float h[8];
    for (int j = 0; j < 8; ++j) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += X[r * 8 + k] * W1[k * 8 + j];
      h[j] = s > 0 ? s : 0;
    }
    for (int j = 0; j < 8; ++j) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += h[k] * W2[k * 8 + j];
      R[r * 8 + j] = s;
    }
  }
  float *dX, *dW1, *dW2, *dY;
  CK(cudaMalloc(&dX, 512)); CK(cudaMalloc(&dW1, 256)); CK(cudaMalloc(&dW2, 256)); CK(cudaMalloc(&dY, 512));
  CK(cudaMemcpy(dX, X.data(), 512, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW1, W1.data(), 256, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW2, W2.data(), 256, cudaMemcpyHostToDevice));
  chain_check<<<1, 32>>>(dX, dW1, dW2, dY);
  CK(cudaMemcpy(Y.data(), dY, 512, cudaMemcpyDeviceToHost));
  float err = 0;
  for (int i = 0; i < 128; ++i) err = fmaxf(err, fabsf(Y[i] - R[i]));
  printf("chain layout check: max |err| = %.3e (tf32 truncation of the hidden layer allowed: ~1e-3)\n", err);
  return 0;
}
