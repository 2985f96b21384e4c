// Wide conditioner layers on tcgen05: Y = epilogue(X W + b) for X (rows x K), W (K x N), fp32 in
// and out with fp32 fidelity (3xTF32), for the H = 512 conditioners of BASELINE config 5
// (/root/reference/cnf_ot/models/flows.py:65-81: hk.nets.MLP([H]*M) -> hk.Linear(P)).
//
// One CTA = one 128-row x NT-column output tile, accumulator in TMEM (NT fp32 columns x 128 lanes).
// The contraction runs in chunks of 16 (one 64-byte swizzled row per operand row):
//   A stage  [128 rows][16 k]  K-major, SWIZZLE_64B -- written by the CTA's threads from X, twice:
//            the values themselves (the tensor core reads their tf32 bits = x_hi) and the exact
//            residuals x_lo = x - x_hi;
//   B stage  [NT n][16 k] hi | lo, same layout -- the weights are prepared ONCE per call into exactly
//            this tile order (dense_prep_kernel), so a stage is one contiguous block that a single
//            thread fetches with a 1-D bulk TMA copy (cp.async.bulk -> mbarrier complete_tx).
//   per chunk: 2 k-steps x { x_lo*w_hi, x_hi*w_lo, x_hi*w_hi } = 6 tcgen05.mma.kind::tf32
//            (M = 128, N = NT, K = 8) issued by one thread, tcgen05.commit -> "stage free" mbarrier.
// Two stages: the loads of chunk i+1 overlap the MMAs of chunk i.  Epilogue: every warp reads its 32
// TMEM lanes with tcgen05.ld (32 columns at a time), adds the bias, applies ReLU or a ReLU mask,
// and writes full 128-byte row segments.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dense_tc.h"

namespace cnfot {

namespace {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_64B shared-memory matrix descriptor: 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;        // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void commit_to(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 3xTF32 split with round-to-nearest parts (zero-mean errors; a truncating split is biased and the bias adds
// up over the contraction): hi = rn_tf32(v), lo = rn_tf32(v - hi); v - hi is exact in fp32.
// (cvt.rna.tf32.f32 leaves the low 13 bits of its result zero -- tools/cvt_probe.cu, 4M bit patterns -- so the result
// is the tf32 value as a float and v - hi is the exact residual)
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
  lo = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
}

}  // namespace

// out[(kc * 2 + part) * N * 16 + sw64_pos(n, kk)] = hi / lo of  Wv(kc * 16 + kk, n),
// Wv(k, n) = transpose ? W[n * ldw + k] : W[k * ldw + n];  K % 16 == 0.
__global__ void dense_prep_kernel(const float* __restrict__ W, int K, int N, int ldw, int transpose,
                                  float* __restrict__ out) {
  const int64_t total = (int64_t)K * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int k, n;
    if (transpose) { k = (int)(e % K); n = (int)(e / K); }   // consecutive threads read consecutive memory
    else { n = (int)(e % N); k = (int)(e / N); }
    const float v = transpose ? W[(int64_t)n * ldw + k] : W[(int64_t)k * ldw + n];
    const float hi = tf32_rn(v);
    const int kc = k >> 4, kk = k & 15;
    float* t = out + (int64_t)kc * 2 * N * 16;
    const int p = sw64_pos(n, kk);
    t[p] = hi;
    t[N * 16 + p] = tf32_rn(v - hi);
  }
}

template <int NT>
struct DenseSmem {
  static constexpr int kAFloats = 128 * 16;
  static constexpr int kBFloats = NT * 16;
  static constexpr int kStageFloats = 2 * kAFloats + 2 * kBFloats;   // A_hi | A_lo | B_hi | B_lo
  static constexpr int kBytes = 2 * kStageFloats * 4 + 1024;
};

__host__ __device__ constexpr int tmem_cols(int nt) { return nt <= 32 ? 32 : nt <= 64 ? 64 : nt <= 128 ? 128 : 256; }

// EPI: 0 = bias, 1 = bias + ReLU, 2 = ReLU mask (Y = mask_src > 0 ? acc : 0), 3 = none, 4 = accumulate (Y += acc)
template <int NT, int EPI>
__global__ void __launch_bounds__(128)
dense_tc_kernel(const float* __restrict__ X, int64_t rows, int K, int ldx, const float* __restrict__ Bt,
                int N_total, const float* __restrict__ bias, const float* __restrict__ mask_src, int ldm,
                float* __restrict__ Y, int ldy, int acc2) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[NT];
  using S = DenseSmem<NT>;
  // acc2: the two small products (x_lo w_hi, x_hi w_lo) go to a second accumulator, NT columns further
  const int kTmemCols = acc2 ? tmem_cols(2 * NT) : tmem_cols(NT);
  const uint32_t lo_off = acc2 ? (uint32_t)NT : 0u;
  const int tid = threadIdx.x, warp = tid >> 5;
  // 1-D grid, column tile fastest: the CTAs that share a row tile are adjacent in launch order, so the second
  // read of the A rows hits L2 (ncu r01: with the column tile on gridDim.y the A matrix came from HBM twice)
  const int ntn = N_total / NT;
  const int64_t row0 = (int64_t)(blockIdx.x / ntn) * 128;
  const int n0 = (int)(blockIdx.x % ntn) * NT;
  const int64_t r = row0 + tid;
  const bool live = r < rows;

  if (tid == 0) {
    mbar_init(s_u32(&bar_full[0]), 1); mbar_init(s_u32(&bar_full[1]), 1);
    mbar_init(s_u32(&bar_free[0]), 1); mbar_init(s_u32(&bar_free[1]), 1);
    mbar_init(s_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the tile's bias values wait in shared memory for the epilogue (read there as broadcasts): loading them from
  // global memory inside the epilogue loop exposed one L2 latency per 16-column step
  if (EPI == 0 || EPI == 1)
    for (int j = tid; j < NT; j += 128) bias_s[j] = __ldg(bias + n0 + j);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_tf32(NT);
  const int n_chunks = K >> 4;
  const int sw = (tid >> 1) & 3;
  const float* xrow = X + r * ldx;

  for (int i = 0; i < n_chunks; ++i) {
    const int s = i & 1;
    float* stage = smem + s * S::kStageFloats;
    // the A chunk of this thread's row (issued before any wait: the loads fly while the stage drains)
    float4 xa[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      xa[c] = live ? __ldg(reinterpret_cast<const float4*>(xrow + i * 16) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= 2) mbar_wait(s_u32(&bar_free[s]), ((i >> 1) - 1) & 1);   // MMAs of chunk i-2 are done with the stage
    if (tid == 0) {
      // weights: one contiguous [hi | lo] block per (chunk, NT columns)
      const float* src = Bt + ((int64_t)i * 2 * N_total + n0) * 16;
      mbar_expect_tx(s_u32(&bar_full[s]), 2 * S::kBFloats * 4);
      bulk_g2s(s_u32(stage + 2 * S::kAFloats), src, S::kBFloats * 4, s_u32(&bar_full[s]));
      bulk_g2s(s_u32(stage + 2 * S::kAFloats + S::kBFloats), src + (int64_t)N_total * 16, S::kBFloats * 4,
               s_u32(&bar_full[s]));
    }
    float* ah = stage + tid * 16;
    float* al = stage + S::kAFloats + tid * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 hi, lo;
      split4(xa[c], hi, lo);
      *reinterpret_cast<float4*>(ah + ((c ^ sw) << 2)) = hi;
      *reinterpret_cast<float4*>(al + ((c ^ sw) << 2)) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      mbar_wait(s_u32(&bar_full[s]), (i >> 1) & 1);
      const uint32_t a_hi = s_u32(stage), a_lo = a_hi + S::kAFloats * 4;
      const uint32_t b_hi = a_hi + 2 * S::kAFloats * 4, b_lo = b_hi + S::kBFloats * 4;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t o = ks * 32;
        umma(tmem + lo_off, desc_sw64(a_lo + o), desc_sw64(b_hi + o), idesc, (i | ks) != 0);
        umma(tmem + lo_off, desc_sw64(a_hi + o), desc_sw64(b_lo + o), idesc, 1);
        umma(tmem, desc_sw64(a_hi + o), desc_sw64(b_hi + o), idesc, acc2 ? (i | ks) != 0 : 1);
      }
      commit_to(s_u32(&bar_free[s]));
      if (i == n_chunks - 1) commit_to(s_u32(&bar_done));
    }
  }
  mbar_wait(s_u32(&bar_done), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");

  // epilogue: thread = row (TMEM lane), 16 columns per tcgen05.ld.  The global-memory operands of a step (the ReLU
  // mask source, EPI 2; the old output, EPI 4) are fetched one step ahead.
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cur[4] = {z4, z4, z4, z4}, nxt[4] = {z4, z4, z4, z4};
  auto prefetch = [&](int c0, float4 (&dst)[4]) {
    if (EPI == 2) {
      const float4* m4 = reinterpret_cast<const float4*>(mask_src + r * ldm + n0 + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = __ldg(m4 + q);
    } else if (EPI == 4) {
      const float4* y4 = reinterpret_cast<const float4*>(Y + r * ldy + n0 + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = y4[q];
    }
  };
  if (live && (EPI == 2 || EPI == 4)) prefetch(0, cur);
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 16) {
    if (live && (EPI == 2 || EPI == 4) && c0 + 16 < NT) prefetch(c0 + 16, nxt);
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (acc2) {
      uint32_t u[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
            "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
          : "r"(taddr + lo_off + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
    }
    if (live) {
      float o[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 bq = z4;
        if (EPI == 0 || EPI == 1) bq = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);
        o[4 * q] = __uint_as_float(v[4 * q]) + bq.x;
        o[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bq.y;
        o[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bq.z;
        o[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bq.w;
      }
      if (EPI == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      if (EPI == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 m = cur[q];
          o[4 * q] = m.x > 0.f ? o[4 * q] : 0.f;
          o[4 * q + 1] = m.y > 0.f ? o[4 * q + 1] : 0.f;
          o[4 * q + 2] = m.z > 0.f ? o[4 * q + 2] : 0.f;
          o[4 * q + 3] = m.w > 0.f ? o[4 * q + 3] : 0.f;
        }
      }
      float4* y4 = reinterpret_cast<float4*>(Y + r * ldy + n0 + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 w = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        if (EPI == 4) {
          const float4 old = cur[q];
          w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
        }
        y4[q] = w;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// ---- weight gradient: dW (Ka x Nb) += A^T G over a range of rows, A (rows x Ka), G (rows x Nb) --------
// tcgen05.kind::tf32 takes K-major operands only (tools/tc_probe.cu: MN-major descriptors are not
// executed), so the contraction index -- the sample rows -- must be contiguous inside a stage row.
// Thread t owns operand row t of the stage (feature m0+t of A, features n0+t and n0+128+t of G) and
// gathers its 16 values of a chunk with 16 loads that are coalesced across the warp (consecutive
// features of one sample row), then writes its 64-byte stage row (values and residuals) with the same
// swizzle as the forward kernel.  Same MMA schedule; the 128 x NT accumulator of the CTA's row range
// is added to dW with red.global.add.v4.f32 (split over row ranges: grid.z).
template <int NT>
__global__ void __launch_bounds__(128)
dense_wgrad_kernel(const float* __restrict__ A, int lda, const float* __restrict__ G, int ldg, int64_t rows,
                   int Ka, int Nb, float* __restrict__ dW, int ldw, const WgradMap map, float* __restrict__ db) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  using S = DenseSmem<NT>;
  constexpr int kTmemCols = tmem_cols(NT);
  constexpr int NB = NT / 128 > 0 ? NT / 128 : 1;   // G features per thread (NT = 128 or 256), or < 128 features
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * NT;
  // this CTA's row range, in whole chunks of 16
  const int64_t chunks_total = (rows + 15) / 16;
  const int64_t per = (chunks_total + gridDim.z - 1) / gridDim.z;
  const int64_t c_lo = blockIdx.z * per, c_hi = c_lo + per < chunks_total ? c_lo + per : chunks_total;
  if (c_lo >= c_hi) return;

  if (tid == 0) {
    mbar_init(s_u32(&bar_free[0]), 1); mbar_init(s_u32(&bar_free[1]), 1);
    mbar_init(s_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_tf32(NT);
  const int sw = (tid >> 1) & 3;
  const bool a_ok = m0 + tid < Ka;
  const float* acol = A + m0 + tid;
  float bsum[NB];
#pragma unroll
  for (int h = 0; h < NB; ++h) bsum[h] = 0.f;
  if (map.mode == 0 ? blockIdx.x != 0 : blockIdx.y != 0) db = nullptr;

  for (int64_t c = c_lo; c < c_hi; ++c) {
    const int i = (int)(c - c_lo), s = i & 1;
    float* stage = smem + s * S::kStageFloats;
    const int64_t r0 = c * 16;
    float av[16], gv[NB][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const bool in = r0 + k < rows;
      av[k] = (in && a_ok) ? __ldg(acol + (r0 + k) * lda) : 0.f;
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        const int n = n0 + h * 128 + tid;
        gv[h][k] = (in && tid + h * 128 < NT && n < Nb) ? __ldg(G + (r0 + k) * ldg + n) : 0.f;
      }
    }
    // bias gradient on the side: column sums of G (mode 0, by the CTAs of the first feature tile) or of A
    // (mode 1: the operands are swapped there, by the CTAs of the first state-column tile)
    if (db) {
      if (map.mode == 0) {
#pragma unroll
        for (int h = 0; h < NB; ++h)
#pragma unroll
          for (int k = 0; k < 16; ++k) bsum[h] += gv[h][k];
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) bsum[0] += av[k];
      }
    }
    if (i >= 2) mbar_wait(s_u32(&bar_free[s]), ((i >> 1) - 1) & 1);
    float* ah = stage + tid * 16;
    float* al = stage + S::kAFloats + tid * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 hi, lo;
      split4(make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]), hi, lo);
      *reinterpret_cast<float4*>(ah + ((q ^ sw) << 2)) = hi;
      *reinterpret_cast<float4*>(al + ((q ^ sw) << 2)) = lo;
    }
#pragma unroll
    for (int h = 0; h < NB; ++h) {
      const int row = h * 128 + tid;
      if (row < NT) {
        float* bh = stage + 2 * S::kAFloats + row * 16;
        float* bl = bh + S::kBFloats;
        const int swb = (row >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 hi, lo;
          split4(make_float4(gv[h][4 * q], gv[h][4 * q + 1], gv[h][4 * q + 2], gv[h][4 * q + 3]), hi, lo);
          *reinterpret_cast<float4*>(bh + ((q ^ swb) << 2)) = hi;
          *reinterpret_cast<float4*>(bl + ((q ^ swb) << 2)) = lo;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a_hi = s_u32(stage), a_lo = a_hi + S::kAFloats * 4;
      const uint32_t b_hi = a_hi + 2 * S::kAFloats * 4, b_lo = b_hi + S::kBFloats * 4;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t o = ks * 32;
        umma(tmem, desc_sw64(a_lo + o), desc_sw64(b_hi + o), idesc, (i | ks) != 0);
        umma(tmem, desc_sw64(a_hi + o), desc_sw64(b_lo + o), idesc, 1);
        umma(tmem, desc_sw64(a_hi + o), desc_sw64(b_hi + o), idesc, 1);
      }
      commit_to(s_u32(&bar_free[s]));
      if (c == c_hi - 1) commit_to(s_u32(&bar_done));
    }
  }
  if (db) {
    if (map.mode == 0) {
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        const int n = n0 + h * 128 + tid;
        if (tid + h * 128 < NT && n < Nb) atomicAdd(db + n, bsum[h]);
      }
    } else if (a_ok) {
      atomicAdd(db + m0 + tid, bsum[0]);
    }
  }
  mbar_wait(s_u32(&bar_done), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  float* wrow = dW + (int64_t)(m0 + tid) * ldw + n0;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (a_ok && map.mode == 1) {
      // transposed destination with the conditioner-input row map: accumulator (m, n) = sum_r A[r][m] G[r][n]
      // is the gradient of W0[w0_row(n)][m] (G = the state rows, A = the adjoint of the first hidden layer)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        const int row = n < Nb ? w0_row(n, map.D, map.d, map.rev) : -1;
        if (row >= 0)
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dW + (int64_t)row * ldw + m0 + tid), "f"(__uint_as_float(v[j]))
                       : "memory");
      }
    } else if (a_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (n0 + c0 + 4 * q < Nb)
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(wrow + c0 + 4 * q), "f"(__uint_as_float(v[4 * q])),
                       "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])),
                       "f"(__uint_as_float(v[4 * q + 3]))
                       : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// cudaFuncSetAttribute once per (kernel instantiation, device): it costs microseconds per call, which adds up over
// the hundreds of launches of a wide-engine step on small batches
static cudaError_t set_smem_once(const void* kern, int bytes, bool (&done)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

template <int NT>
static cudaError_t launch_wgrad(cudaStream_t s, const float* A, int lda, const float* G, int ldg, int64_t rows, int Ka,
                                int Nb, float* dW, int ldw, const WgradMap& map, float* db) {
  auto kern = dense_wgrad_kernel<NT>;
  static bool done[64] = {};
  cudaError_t e = set_smem_once((const void*)kern, DenseSmem<NT>::kBytes, done);
  if (e != cudaSuccess) return e;
  const int mt = (Ka + 127) / 128, nt = (Nb + NT - 1) / NT;
  const int64_t chunks = (rows + 15) / 16;
  int64_t z = (148 * 2 + mt * nt - 1) / (mt * nt);   // about one wave of 2 CTAs per SM
  if (z > chunks) z = chunks;
  if (z > 65535) z = 65535;
  if (z < 1) z = 1;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)z);
  kern<<<grid, 128, DenseSmem<NT>::kBytes, s>>>(A, lda, G, ldg, rows, Ka, Nb, dW, ldw, map, db);
  return cudaGetLastError();
}

cudaError_t dense_wgrad(cudaStream_t s, const float* A, int lda, const float* G, int ldg, int64_t rows, int Ka, int Nb,
                        float* dW, int ldw, float* db, const WgradMap* mapp) {
  if (rows == 0) return cudaSuccess;
  WgradMap map;
  map.mode = 0; map.D = 0; map.d = 0; map.rev = 0;
  if (mapp) map = *mapp;
  if (Nb > 128) return launch_wgrad<256>(s, A, lda, G, ldg, rows, Ka, Nb, dW, ldw, map, db);
  if (Nb > 64) return launch_wgrad<128>(s, A, lda, G, ldg, rows, Ka, Nb, dW, ldw, map, db);
  if (Nb > 16) return launch_wgrad<64>(s, A, lda, G, ldg, rows, Ka, Nb, dW, ldw, map, db);
  return launch_wgrad<16>(s, A, lda, G, ldg, rows, Ka, Nb, dW, ldw, map, db);
}

template <int NT, int EPI>
static cudaError_t launch_dense(cudaStream_t s, const float* X, int64_t rows, int K, int ldx, const float* Bt,
                                int N, const float* bias, const float* mask_src, int ldm, float* Y, int ldy) {
  auto kern = dense_tc_kernel<NT, EPI>;
  static bool done[64] = {};
  cudaError_t e = set_smem_once((const void*)kern, DenseSmem<NT>::kBytes, done);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)(((rows + 127) / 128) * (N / NT)));
  // CNFOT_DENSE_ACC2=1: separate accumulator for the small 3xTF32 terms (needs 2 NT <= 512 TMEM columns)
  int acc2 = 0;
  if (const char* ev = getenv("CNFOT_DENSE_ACC2")) acc2 = ev[0] == '1' && 2 * NT <= 512;
  kern<<<grid, 128, DenseSmem<NT>::kBytes, s>>>(X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy, acc2);
  return cudaGetLastError();
}

cudaError_t dense_prep(cudaStream_t s, const float* W, int K, int N, int ldw, bool transpose, float* out) {
  const int64_t total = (int64_t)K * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dense_prep_kernel<<<blocks, 256, 0, s>>>(W, K, N, ldw, transpose ? 1 : 0, out);
  return cudaGetLastError();
}

cudaError_t dense_forward(cudaStream_t s, const float* X, int64_t rows, int K, int ldx, const float* Bt, int N,
                          const float* bias, const float* mask_src, int ldm, int epilogue, float* Y, int ldy,
                          bool* supported) {
  *supported = true;
  if (rows == 0) return cudaSuccess;
#define DENSE_CASE(NT_)                                                                                   \
  switch (epilogue) {                                                                                    \
    case 0: return launch_dense<NT_, 0>(s, X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy);         \
    case 1: return launch_dense<NT_, 1>(s, X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy);         \
    case 2: return launch_dense<NT_, 2>(s, X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy);         \
    case 3: return launch_dense<NT_, 3>(s, X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy);         \
    case 4: return launch_dense<NT_, 4>(s, X, rows, K, ldx, Bt, N, bias, mask_src, ldm, Y, ldy);         \
  }
  if (N % 256 == 0) { DENSE_CASE(256) }
  else if (N % 128 == 0) { DENSE_CASE(128) }
  else if (N % 64 == 0) { DENSE_CASE(64) }
  else if (N % 16 == 0 && N <= 64) {
    if (N == 16) { DENSE_CASE(16) }
    else if (N == 32) { DENSE_CASE(32) }
    else if (N == 48) { DENSE_CASE(48) }
  }
#undef DENSE_CASE
  *supported = false;
  return cudaSuccess;
}

}  // namespace cnfot
