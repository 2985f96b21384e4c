// Stand-alone spline kernels behind cnfot_rqs_* (seam 2 of include/cnfot.h):
// one scalar rational-quadratic spline per row with per-row raw parameters,
// replacing distrax.RationalQuadraticSpline(params).forward_and_log_det /
// inverse_and_log_det as called at
// /root/reference/cnf_ot/models/autoregressive.py:100,130.
//
// HBM-bound: per row the kernel must move 4P (params) + 4 (x) in and 8 (y,
// logdet) out = 4P + 12 bytes (76 B at K = 5); the backward moves 8P + 16.
// A CTA owns 128 consecutive rows: the (128 x P) parameter tile is one
// contiguous 512P-byte block, copied with coalesced 128-bit loads into a
// padded shared-memory tile (row stride = odd number of 16-byte units, so the
// per-row 128-bit reads are bank-conflict-free), then every thread evaluates
// its row entirely in registers.  Gradients w.r.t. the parameters leave the
// same way in reverse.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"
#include "rqs_math.cuh"

namespace cnfot {

template <int P>
struct ParamTile {
  static constexpr bool kVec = (P % 4) == 0;
  // vector path: odd number of 16-byte units; scalar path: odd number of words
  static constexpr int kStride = kVec ? ((((P / 4) & 1) == 0) ? P + 4 : P) : ((P & 1) ? P : P + 1);
  static constexpr int kFloats = kTile * kStride;
};

// global (tile of `nrows` x P, contiguous) -> shared padded tile
template <int P>
__device__ __forceinline__ void tile_load(float* s, const float* __restrict__ g, int nrows) {
  using TL = ParamTile<P>;
  if (TL::kVec) {
    constexpr int C = P / 4;  // 16-byte chunks per row
    const float4* src = reinterpret_cast<const float4*>(g);
    if (nrows == kTile && (kTile % C) == 0) {
      // full tile: chunk q = tid + j kTile sits in row (tid / C) + j (kTile / C), column tid % C -- one base address
      // on each side, the C copies are immediate offsets from it and all C loads are in flight together
      const int r = threadIdx.x / C, c = threadIdx.x - r * C;
      const float4* sp = src + threadIdx.x;
      float* dp = s + r * TL::kStride + c * 4;
      float4 v[C];
#pragma unroll
      for (int j = 0; j < C; ++j) v[j] = __ldcs(sp + j * kTile);  // streamed once: evict-first
#pragma unroll
      for (int j = 0; j < C; ++j) *reinterpret_cast<float4*>(dp + j * (kTile / C) * TL::kStride) = v[j];
      return;
    }
    const int n = nrows * C;
#pragma unroll 4
    for (int q = threadIdx.x; q < n; q += kTile) {
      const int r = q / C, c = q - r * C;
      float4 v = __ldcs(src + q);  // streamed once: evict-first
      *reinterpret_cast<float4*>(s + r * TL::kStride + c * 4) = v;
    }
  } else {
    const int n = nrows * P;
#pragma unroll 4
    for (int f = threadIdx.x; f < n; f += kTile) {
      const int r = f / P, c = f - r * P;
      s[r * TL::kStride + c] = __ldcs(g + f);
    }
  }
}

template <int P>
__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* s, int nrows) {
  using TL = ParamTile<P>;
  if (TL::kVec) {
    constexpr int C = P / 4;
    float4* dst = reinterpret_cast<float4*>(g);
    if (nrows == kTile && (kTile % C) == 0) {   // full tile: see tile_load
      const int r = threadIdx.x / C, c = threadIdx.x - r * C;
      float4* gp = dst + threadIdx.x;
      const float* sp = s + r * TL::kStride + c * 4;
#pragma unroll
      for (int j = 0; j < C; ++j) __stcs(gp + j * kTile, *reinterpret_cast<const float4*>(sp + j * (kTile / C) * TL::kStride));
      return;
    }
    const int n = nrows * C;
#pragma unroll 4
    for (int q = threadIdx.x; q < n; q += kTile) {
      const int r = q / C, c = q - r * C;
      __stcs(dst + q, *reinterpret_cast<const float4*>(s + r * TL::kStride + c * 4));
    }
  } else {
    const int n = nrows * P;
#pragma unroll 4
    for (int f = threadIdx.x; f < n; f += kTile) {
      const int r = f / P, c = f - r * P;
      __stcs(g + f, s[r * TL::kStride + c]);
    }
  }
}

template <int P>
__device__ __forceinline__ void row_read(const float* s, float* theta) {
  using TL = ParamTile<P>;
  const float* p = s + threadIdx.x * TL::kStride;
  if (TL::kVec) {
#pragma unroll
    for (int j = 0; j < P; j += 4) {
      float4 v = *reinterpret_cast<const float4*>(p + j);
      theta[j] = v.x; theta[j + 1] = v.y; theta[j + 2] = v.z; theta[j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < P; ++j) theta[j] = p[j];
  }
}

template <int P>
__device__ __forceinline__ void row_write(float* s, const float* theta) {
  using TL = ParamTile<P>;
  float* p = s + threadIdx.x * TL::kStride;
  if (TL::kVec) {
#pragma unroll
    for (int j = 0; j < P; j += 4)
      *reinterpret_cast<float4*>(p + j) = make_float4(theta[j], theta[j + 1], theta[j + 2], theta[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < P; ++j) p[j] = theta[j];
  }
}

template <int K, bool INVERSE>
__global__ void __launch_bounds__(kTile)
rqs_eval_kernel(const float* __restrict__ v, const float* __restrict__ params, int64_t rows,
                SplineConsts<float> sc, float* __restrict__ out, float* __restrict__ logdet,
                int32_t* __restrict__ bin) {
  constexpr int P = 3 * K + 1;
  __shared__ __align__(16) float tile[ParamTile<P>::kFloats];
  const int64_t n_tiles = (rows + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r0 = t * kTile;
    const int nrows = (int)((rows - r0) < kTile ? (rows - r0) : kTile);
    const int64_t r = r0 + threadIdx.x;
    const bool live = threadIdx.x < nrows;
    const float xv = live ? __ldcs(v + r) : 0.f;  // issued before the tile copy: overlaps it
    __syncthreads();                               // previous iteration's readers are done
    tile_load<P>(tile, params + r0 * P, nrows);
    __syncthreads();
    if (live) {
      float theta[P];
      row_read<P>(tile, theta);
      SplineState<float, K> st;
      float o, ld;
      if (INVERSE) rqs_inverse<float, K>(xv, theta, sc, st, o, ld);
      else rqs_forward<float, K>(xv, theta, sc, st, o, ld);
      __stcs(out + r, o);
      __stcs(logdet + r, ld);
      if (bin) __stcs(bin + r, st.idx);
    }
  }
}

template <int K, bool INVERSE>
__global__ void __launch_bounds__(kTile)
rqs_vjp_kernel(const float* __restrict__ v, const float* __restrict__ params,
               const float* __restrict__ g_out, const float* __restrict__ g_ld, int64_t rows,
               SplineConsts<float> sc, float* __restrict__ g_in, float* __restrict__ g_params) {
  constexpr int P = 3 * K + 1;
  __shared__ __align__(16) float tile[ParamTile<P>::kFloats];
  const int64_t n_tiles = (rows + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r0 = t * kTile;
    const int nrows = (int)((rows - r0) < kTile ? (rows - r0) : kTile);
    const int64_t r = r0 + threadIdx.x;
    const bool live = threadIdx.x < nrows;
    const float xv = live ? __ldcs(v + r) : 0.f;
    const float go = live ? __ldcs(g_out + r) : 0.f;
    const float gl = live ? __ldcs(g_ld + r) : 0.f;
    __syncthreads();
    tile_load<P>(tile, params + r0 * P, nrows);
    __syncthreads();
    if (live) {
      float theta[P], gtheta[P];
      row_read<P>(tile, theta);
      SplineState<float, K> st;
      float o, ld, gi;
      if (INVERSE) {
        rqs_inverse<float, K>(xv, theta, sc, st, o, ld);
        gi = rqs_inverse_bwd<float, K>(xv, st, sc, go, gl, gtheta);
      } else {
        rqs_forward<float, K>(xv, theta, sc, st, o, ld);
        gi = rqs_forward_bwd<float, K>(xv, st, sc, go, gl, gtheta);
      }
      __stcs(g_in + r, gi);
      row_write<P>(tile, gtheta);  // each thread overwrites only its own row
    }
    __syncthreads();
    tile_store<P>(g_params + r0 * P, tile, nrows);
  }
}

template <int K>
static cudaError_t launch_eval(bool inverse, cudaStream_t s, const float* v, const float* params,
                               int64_t rows, const SplineConsts<float>& sc, float* out,
                               float* ld, int32_t* bin, int grid) {
  if (inverse) rqs_eval_kernel<K, true><<<grid, kTile, 0, s>>>(v, params, rows, sc, out, ld, bin);
  else rqs_eval_kernel<K, false><<<grid, kTile, 0, s>>>(v, params, rows, sc, out, ld, bin);
  return cudaGetLastError();
}

template <int K>
static cudaError_t launch_vjp(bool inverse, cudaStream_t s, const float* v, const float* params,
                              const float* go, const float* gl, int64_t rows,
                              const SplineConsts<float>& sc, float* gi, float* gp, int grid) {
  if (inverse) rqs_vjp_kernel<K, true><<<grid, kTile, 0, s>>>(v, params, go, gl, rows, sc, gi, gp);
  else rqs_vjp_kernel<K, false><<<grid, kTile, 0, s>>>(v, params, go, gl, rows, sc, gi, gp);
  return cudaGetLastError();
}

#define CNFOT_BINS_LIST(X) X(3) X(4) X(5) X(8) X(10) X(16)

int rqs_grid(int64_t rows, int num_sms) {
  int64_t tiles = (rows + kTile - 1) / kTile;
  int64_t cap = (int64_t)num_sms * 16;  // persistent: up to 16 resident CTAs per SM
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

cudaError_t rqs_eval_dispatch(int K, bool inverse, cudaStream_t s, const float* v,
                              const float* params, int64_t rows, const SplineConsts<float>& sc,
                              float* out, float* ld, int32_t* bin, int num_sms, bool* known) {
  *known = true;
  int grid = rqs_grid(rows, num_sms);
  switch (K) {
#define X(KK) case KK: return launch_eval<KK>(inverse, s, v, params, rows, sc, out, ld, bin, grid);
    CNFOT_BINS_LIST(X)
#undef X
    default: *known = false; return cudaSuccess;
  }
}

cudaError_t rqs_vjp_dispatch(int K, bool inverse, cudaStream_t s, const float* v,
                             const float* params, const float* go, const float* gl, int64_t rows,
                             const SplineConsts<float>& sc, float* gi, float* gp, int num_sms,
                             bool* known) {
  *known = true;
  int grid = rqs_grid(rows, num_sms);
  switch (K) {
#define X(KK) case KK: return launch_vjp<KK>(inverse, s, v, params, go, gl, rows, sc, gi, gp, grid);
    CNFOT_BINS_LIST(X)
#undef X
    default: *known = false; return cudaSuccess;
  }
}

}  // namespace cnfot
