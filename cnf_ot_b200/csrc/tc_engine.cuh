// tcgen05 engine for the conditioner's dense layers (hidden and output linears,
// forward and data-gradient), used by the fused kernels when hidden == 16 and the
// padded spline-parameter count is 16 (mfc.yaml defaults).
//
// One CTA tile = 128 rows = the 128 TMEM lanes of one accumulator.  Per layer:
//   D[128 x 16] (TMEM, fp32) = X[128 x 16] * B^T[16 x 16]        kind::tf32, M=128 N=16 K=8 x2
// X is the row tile the owning threads just wrote (row-major, 64-byte rows, 16-byte chunks
// XOR-swizzled with (row>>1)&3 -- exactly the canonical K-major SWIZZLE_64B operand layout,
// so the tile that feeds the CUDA-core weight-gradient reduction also feeds the MMA).
// fp32 fidelity: the tensor core truncates fp32 inputs to tf32 (measured: tools/tc_probe.cu),
// so every product is split  x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo  with x_hi = trunc(x)
// taken implicitly by the hardware from the full-precision tile and x_lo = rn_tf32(x - trunc(x))
// written to one scratch tile; the weights' hi/lo tiles are built once per CTA.
// One elected thread issues the 6 MMAs and a tcgen05.commit to an mbarrier; every thread
// then reads its own row of D with tcgen05.ld (32 lanes x 32 bit, 16 columns).
#pragma once

#include "device_common.cuh"

namespace cnfot {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor: 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;        // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=16
__device__ __forceinline__ uint32_t umma_idesc_tf32_m128_n16() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ float tf32_trunc(float v) {
  return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}
// low half of the split: v - trunc(v), rounded to nearest tf32 so the hardware's truncation of
// the operand is exact (keeps the split's error at ~2^-22 |v| and unbiased)
__device__ __forceinline__ float tf32_lo(float v) {
  const float lo = v - tf32_trunc(v);
  return __uint_as_float((__float_as_uint(lo) + 0x00001000u) & 0xFFFFE000u);
}

constexpr int kTcTileFloats = 16 * 16;   // one 16 x 16 weight tile in MMA layout
constexpr int kTcTmemCols = 32;

template <class Net>
struct DeviceCtxTC : DeviceCtx<Net> {
  using Base = DeviceCtx<Net>;
  static constexpr int H = Net::kH, M = Net::kM, Pp = Net::kPp;
  static_assert(H == 16 && Pp == 16, "the tcgen05 engine is written for 16-wide layers");
  uint32_t tmem;    // TMEM base address of the accumulator
  uint32_t mbar;    // shared address of the MMA-completion mbarrier
  uint32_t phase;   // parity of the next completion

  __device__ __forceinline__ void setup(int D, int L, uint64_t* mbar_ptr, uint32_t* tmem_slot) {
    this->load();
    tc_setup(D, L, mbar_ptr, tmem_slot);
  }
  __device__ __forceinline__ void teardown() { tc_teardown(); }

  // swizzled position of element (row n, col k) inside a [16][16] MMA tile
  __device__ __forceinline__ static int tile_pos(int n, int k) {
    return n * 16 + ((((k >> 2) ^ ((n >> 1) & 3)) << 2) | (k & 3));
  }

  // After Base::load(): build the weights' MMA tiles, allocate TMEM, arm the barrier.
  // Per dense layer `mat`: [fwd_hi | fwd_lo | bwd_hi | bwd_lo], fwd(n=j,k=i) = bwd(n=i,k=j) = W[i][j].
  __device__ __forceinline__ void tc_setup(int D, int L, uint64_t* mbar_ptr, uint32_t* tmem_slot) {
    const SmemPlan& p = this->p;
    float* smem = this->smem;
    const float* blob = smem + p.off_w;
    const int n_mlp = L * (D - 1);
    for (int mlp = 0; mlp < n_mlp; ++mlp) {
      const int layer = mlp / (D - 1), d = mlp - layer * (D - 1) + 1;
      const float* Wb = blob + mlp_offset<Net>(D, layer, d) + (d + 1) * H + H;  // first hidden matrix
      for (int slot = 0; slot < M; ++slot) {
        const float* Ws = Wb + slot * (H * H + H);  // hidden slot (H x H) or the output matrix (H x Pp)
        float* t = smem + p.off_wmma + (mlp * M + slot) * 4 * kTcTileFloats;
        for (int e = threadIdx.x; e < 256; e += blockDim.x) {
          const int i = e >> 4, j = e & 15;
          const float v = Ws[i * 16 + j];
          const float lo = tf32_lo(v);
          t[tile_pos(j, i)] = v;
          t[kTcTileFloats + tile_pos(j, i)] = lo;
          t[2 * kTcTileFloats + tile_pos(i, j)] = v;
          t[3 * kTcTileFloats + tile_pos(i, j)] = lo;
        }
      }
    }
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar_ptr)));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if ((threadIdx.x >> 5) == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   :: "r"(smem_u32(tmem_slot)), "r"(kTcTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    tmem = *tmem_slot;
    mbar = smem_u32(mbar_ptr);
    phase = 0;
  }

  __device__ __forceinline__ void tc_teardown() {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcTmemCols));
  }

  // y[0..16) = row of  X * B^T  for the calling thread's row.  xt: the thread's row of the X tile
  // (already stored, full precision), xr: the same 16 values in registers, btile: index of the
  // hi weight tile (its lo tile follows).
  __device__ __forceinline__ void mma_row(const float* xt, int sw, const float* xr, int btile, float* y) {
    const SmemPlan& p = this->p;
    float* lo = this->smem + p.off_lo + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(lo + chunk_at(j, sw)) =
          make_float4(tf32_lo(xr[j]), tf32_lo(xr[j + 1]), tf32_lo(xr[j + 2]), tf32_lo(xr[j + 3]));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // tile writes -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a_hi = smem_u32(xt);  // thread 0's row pointer is the tile base
      const uint32_t a_lo = smem_u32(this->smem + p.off_lo);
      const uint32_t b_hi = smem_u32(this->smem + p.off_wmma + btile * kTcTileFloats);
      const uint32_t b_lo = b_hi + kTcTileFloats * 4;
      const uint32_t idesc = umma_idesc_tf32_m128_n16();
      // K = 16 = two k-steps of 8 (32 bytes inside the 64-byte swizzled row)
      umma_tf32(tmem, umma_desc_sw64(a_hi), umma_desc_sw64(b_hi), idesc, 0);
      umma_tf32(tmem, umma_desc_sw64(a_hi + 32), umma_desc_sw64(b_hi + 32), idesc, 1);
      umma_tf32(tmem, umma_desc_sw64(a_lo), umma_desc_sw64(b_hi), idesc, 1);
      umma_tf32(tmem, umma_desc_sw64(a_lo + 32), umma_desc_sw64(b_hi + 32), idesc, 1);
      umma_tf32(tmem, umma_desc_sw64(a_hi), umma_desc_sw64(b_lo), idesc, 1);
      umma_tf32(tmem, umma_desc_sw64(a_hi + 32), umma_desc_sw64(b_lo + 32), idesc, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   :: "r"(mbar) : "memory");
    }
    while (!mbar_try_wait(mbar, phase)) {}
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(v[j]);
  }

  template <int K, int N>
  __device__ __forceinline__ void dense_fwd(const float* xt, int sw, const float* xr, const float*, int mat, float* y) {
    static_assert(K == 16 && N == 16, "16-wide layers only");
    mma_row(xt, sw, xr, mat * 4, y);
  }
  template <int K, int N>
  __device__ __forceinline__ void dense_bwd(const float* gt, int sw, const float* g, const float*, int mat, float* y) {
    static_assert(K == 16 && N == 16, "16-wide layers only");
    mma_row(gt, sw, g, mat * 4 + 2, y);
  }
};

}  // namespace cnfot
