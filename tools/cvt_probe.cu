// Does cvt.rna.tf32.f32 leave the low 13 bits of its result zero?  (decides whether the 3xTF32 split needs a mask)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const float* in, uint32_t* out, int n, unsigned long long* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i]));
  out[i] = r;
  if (r & 0x1FFFu) atomicAdd(bad, 1ULL);
}
int main() {
  const int n = 1 << 22;
  float* h = (float*)malloc(n * 4);
  uint32_t s = 12345u;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; uint32_t b = s; if (((b >> 23) & 0xFF) == 0xFF) b &= 0x7F7FFFFFu; h[i] = *(float*)&b; }
  float* d; uint32_t* o; unsigned long long* bad; unsigned long long hb = 0;
  cudaMalloc(&d, n * 4); cudaMalloc(&o, n * 4); cudaMalloc(&bad, 8);
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 8);
  probe<<<n / 256, 256>>>(d, o, n, bad);
  cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("cvt.rna.tf32.f32 over %d random bit patterns: %llu results with non-zero low 13 bits (%s)\n", n, hb, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
