// Device-side plumbing shared by the fused flow / train-step kernels:
// shared-memory carving, the CTA-wide weight-gradient sink, loss reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "flow_math.cuh"

namespace cnfot {

constexpr int kTile = 128;       // rows per CTA tile == threads per CTA (one row per thread)
constexpr int kWarps = kTile / 32;

// Row stride (floats) for a staged [kTile][n] tile, n a multiple of 4: an odd
// number of 16-byte units, so 8 consecutive rows hit 8 distinct bank groups and
// both the per-row float4 stores and the strided float4 reads are conflict-free.
__host__ __device__ inline int staged_stride(int n) {
  int q = (n + 3) / 4;
  if ((q & 1) == 0) q += 1;
  return q * 4;
}

struct SmemPlan {
  int total;   // blob floats
  int lda;     // staging stride of the activation tile
  int ldg;     // staging stride of the adjoint tile
  int off_acc, off_sta, off_stg, floats;
};

inline SmemPlan plan_smem(const FlowLayout& f, bool with_grad) {
  SmemPlan p;
  p.total = f.total;
  int wa = f.H > ((f.D + 3) / 4 * 4) ? f.H : (f.D + 3) / 4 * 4;
  int wg = f.H > f.Pp ? f.H : f.Pp;
  p.lda = staged_stride(wa);
  p.ldg = staged_stride(wg);
  int tot4 = (f.total + 3) / 4 * 4;
  p.off_acc = tot4;
  p.off_sta = with_grad ? 2 * tot4 : tot4;
  p.off_stg = p.off_sta + (with_grad ? kTile * p.lda : 0);
  p.floats = p.off_stg + (with_grad ? kTile * p.ldg : 0);
  return p;
}

// Cooperative copy of the weight blob into shared memory (float4, coalesced).
__device__ inline void load_weights(float* sW, const float* __restrict__ gW, int total) {
  const int n4 = total >> 2;
  const float4* src = reinterpret_cast<const float4*>(gW);
  float4* dst = reinterpret_cast<float4*>(sW);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
  for (int i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) sW[i] = __ldg(gW + i);
}

// CTA-wide weight-gradient reduction.
//
// outer<NAMAX, NG>(w_off, Na, a, g): every thread (= row) contributes the rank-1
// update a (x) g to the (Na x NG) matrix at blob offset w_off and g to the bias
// row that directly follows it.  All threads of the CTA must call it together.
//
// Mechanics: each thread stores its a / g row into padded shared-memory tiles;
// after a barrier the CTA re-partitions the work GEMM-style: a lane owns one
// 4x4 block of the matrix for one eighth of the tile's rows (8 row groups x 4
// blocks per warp), accumulates 16 FMAs per float4 pair, the 8 row groups are
// folded with a reduce-scatter butterfly (14 shuffles) and each lane adds its 2
// results to the CTA's shared-memory accumulator.  Every accumulator element
// has exactly one owner lane per call, so no atomics are needed.
struct DeviceSink {
  float* acc;
  float* stA;
  float* stG;
  int lda, ldg;

  template <int NG>
  __device__ __noinline__ void reduce_tile(int Na, float* dst) {
    constexpr int NCB = NG / 4;
    const int nrb = (Na + 3) >> 2;
    const int nblk = (nrb + 1) * NCB;  // + one row of bias blocks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = lane & 7, slot = lane >> 3;
    for (int base = warp * 4; base < nblk; base += 4 * kWarps) {
      const int b = base + slot;
      const bool valid = b < nblk;
      const int rb = valid ? b / NCB : 0, cb = valid ? b - rb * NCB : 0;
      const bool bias = rb == nrb;
      float c[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) c[e] = 0.f;
      if (valid) {
        const float* pg = stG + rg * ldg + cb * 4;
        const float* pa = stA + rg * lda + (bias ? 0 : rb * 4);
#pragma unroll 4
        for (int k = 0; k < kTile / 8; ++k) {
          float4 g4 = *reinterpret_cast<const float4*>(pg + k * 8 * ldg);
          float4 a4 = *reinterpret_cast<const float4*>(pa + k * 8 * lda);
          if (bias) a4 = make_float4(1.f, 0.f, 0.f, 0.f);
          c[0] += a4.x * g4.x; c[1] += a4.x * g4.y; c[2] += a4.x * g4.z; c[3] += a4.x * g4.w;
          c[4] += a4.y * g4.x; c[5] += a4.y * g4.y; c[6] += a4.y * g4.z; c[7] += a4.y * g4.w;
          c[8] += a4.z * g4.x; c[9] += a4.z * g4.y; c[10] += a4.z * g4.z; c[11] += a4.z * g4.w;
          c[12] += a4.w * g4.x; c[13] += a4.w * g4.y; c[14] += a4.w * g4.z; c[15] += a4.w * g4.w;
        }
      }
      // reduce-scatter over the 8 row groups (lane bits 0..2)
      const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
      float v8[8], v4[4], v2[2];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float keep = b2 ? c[8 + q] : c[q], send = b2 ? c[q] : c[8 + q];
        v8[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float keep = b1 ? v8[4 + q] : v8[q], send = b1 ? v8[q] : v8[4 + q];
        v4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float keep = b0 ? v4[2 + q] : v4[q], send = b0 ? v4[q] : v4[2 + q];
        v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      if (valid) {
        const int i = (b2 ? 2 : 0) + (b1 ? 1 : 0);
        const int j = b0 ? 2 : 0;
        const int row = bias ? Na : rb * 4 + i;
        if (bias ? (i == 0) : (row < Na)) {
          float2* p = reinterpret_cast<float2*>(dst + row * NG + cb * 4 + j);
          float2 cur = *p;
          cur.x += v2[0];
          cur.y += v2[1];
          *p = cur;
        }
      }
    }
  }

  template <int NAMAX, int NG>
  __device__ __forceinline__ void outer(int w_off, int Na, const float* a, const float* g) {
    __syncthreads();  // previous call's readers are done with the staging tiles
    float* ra = stA + threadIdx.x * lda;
    float* rgp = stG + threadIdx.x * ldg;
    if (NAMAX == kMaxDim) {
      const int na4 = (Na + 3) & ~3;
      for (int i = 0; i < na4; ++i) ra[i] = i < Na ? a[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < NAMAX; i += 4)
        *reinterpret_cast<float4*>(ra + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
    }
#pragma unroll
    for (int j = 0; j < NG; j += 4)
      *reinterpret_cast<float4*>(rgp + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
    __syncthreads();
    reduce_tile<NG>(Na, acc + w_off);
  }
};

// Sum a per-thread double over the CTA; result valid in thread 0.
__device__ inline double block_sum(double v, double* scratch /* kWarps doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kWarps; ++w) t += scratch[w];
  return t;
}

}  // namespace cnfot
