// Probe for the tcgen05 plumbing used by the conditioner engine (development tool):
//   test 1  D[128x16]  = A[128x16] * B^T     A, B K-major "plane" layout, kind::tf32
//   test 2  D[128x16]  = At[128x128]^T-style: D[m][n] = sum_r A(r,m) G(r,n), both MN-major planes
// Prints max errors against CPU references with truncated / rounded tf32 inputs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
  uint64_t d = (uint64_t)layout << 61;  // 0 none, 2 SW128, 4 SW64, 6 SW32
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;             // D format F32
  d |= 2u << 7;             // A format TF32
  d |= 2u << 10;            // B format TF32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

// mode 0: K-major planes, K = 16 (2 k-steps).  mode 1: MN-major planes, K = 128 rows (16 k-steps)
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);            // up to 32 planes x 2 KB
  float* sB = sA + 32 * 512;                                  // up to 4 planes x 2 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (mode == 5) {
    // row-major 64-byte rows, 16-byte chunks XOR-swizzled with (row >> 1) & 3 (== Swizzle<2,4,3>)
    for (int c = 0; c < 4; ++c)
      for (int q = 0; q < 4; ++q) sA[tid * 16 + ((c ^ ((tid >> 1) & 3)) << 2) + q] = A[tid * 16 + c * 4 + q];
    if (tid < 16)
      for (int c = 0; c < 4; ++c)
        for (int q = 0; q < 4; ++q) sB[tid * 16 + ((c ^ ((tid >> 1) & 3)) << 2) + q] = B[tid * 16 + c * 4 + q];
  } else if (mode == 6) {
    // A[r][m]: 8 tiles (16 features each) of [128 rows][16], swizzled rows; G[r][n]: one such tile
    for (int t = 0; t < 8; ++t)
      for (int c = 0; c < 4; ++c)
        for (int q = 0; q < 4; ++q)
          sA[t * 2048 + tid * 16 + ((c ^ ((tid >> 1) & 3)) << 2) + q] = A[tid * 128 + t * 16 + c * 4 + q];
    for (int c = 0; c < 4; ++c)
      for (int q = 0; q < 4; ++q) sB[tid * 16 + ((c ^ ((tid >> 1) & 3)) << 2) + q] = B[tid * 16 + c * 4 + q];
  } else if (mode == 0) {
    // A[r][k] (128 x 16) -> planes [c][r][4];  B[n][k] (16 x 16) -> planes [c][n][4]
    for (int c = 0; c < 4; ++c)
      for (int q = 0; q < 4; ++q) sA[c * 512 + tid * 4 + q] = A[tid * 16 + c * 4 + q];
    if (tid < 16)
      for (int c = 0; c < 4; ++c)
        for (int q = 0; q < 4; ++q) sB[c * 64 + tid * 4 + q] = B[tid * 16 + c * 4 + q];
  } else {
    const bool a_mn = mode == 1 || mode == 2, b_mn = mode == 1 || mode == 3;
    // A[r][m] (K = 128 rows x 128 feats), G[r][n] (128 x 16)
    if (a_mn) {  // planes [c = m/4][r][4]
      for (int c = 0; c < 32; ++c)
        for (int q = 0; q < 4; ++q) sA[c * 512 + tid * 4 + q] = A[tid * 128 + c * 4 + q];
    } else {     // K-major: planes [c = r/4][m][4]; thread = m
      for (int c = 0; c < 32; ++c)
        for (int q = 0; q < 4; ++q) sA[c * 512 + tid * 4 + q] = A[(c * 4 + q) * 128 + tid];
    }
    if (b_mn) {
      for (int c = 0; c < 4; ++c)
        for (int q = 0; q < 4; ++q) sB[c * 512 + tid * 4 + q] = B[tid * 16 + c * 4 + q];
    } else if (tid < 16) {  // K-major: planes [c = r/4][n][4]
      for (int c = 0; c < 32; ++c)
        for (int q = 0; q < 4; ++q) sB[c * 64 + tid * 4 + q] = B[(c * 4 + q) * 16 + tid];
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;

  if (tid == 0) {
    if (mode == 5) {
      const uint32_t idesc = make_idesc_tf32(128, 16, 0, 0);
      for (int s = 0; s < 2; ++s) {   // K-major SW64: 8-row groups at SBO = 512 B; k-step = +32 B inside the row
        uint64_t ad = make_desc(smem_u32(sA) + s * 32, 16, 512, 4);
        uint64_t bd = make_desc(smem_u32(sB) + s * 32, 16, 512, 4);
        mma_tf32(tmem_d, ad, bd, idesc, s > 0);
      }
    } else if (mode == 6) {
      const uint32_t idesc = make_idesc_tf32(128, 16, 1, 1);
      for (int s = 0; s < 16; ++s) {  // MN-major SW64: 16-feature groups at LBO = tile stride, 8-row K groups at SBO = 512 B
        uint64_t ad = make_desc(smem_u32(sA) + s * 512, 8192, 512, 4);
        uint64_t bd = make_desc(smem_u32(sB) + s * 512, 8192, 512, 4);
        mma_tf32(tmem_d, ad, bd, idesc, s > 0);
      }
    } else if (mode == 0) {
      const uint32_t idesc = make_idesc_tf32(128, 16, 0, 0);
      for (int s = 0; s < 2; ++s) {
        uint64_t ad = make_desc(smem_u32(sA) + s * 2 * 2048, 2048, 128);  // LBO = plane stride, SBO = 8-row group
        uint64_t bd = make_desc(smem_u32(sB) + s * 2 * 256, 256, 128);
        mma_tf32(tmem_d, ad, bd, idesc, s > 0);
      }
    } else {
      const bool a_mn = mode == 1 || mode == 2, b_mn = mode == 1 || mode == 3;
      const uint32_t idesc = make_idesc_tf32(128, 16, a_mn, b_mn);
      for (int s = 0; s < 16; ++s) {
        // MN-major: LBO = 8-row (K) group, SBO = plane;  K-major: LBO = plane (K chunk), SBO = 8-row MN group
        uint64_t ad = a_mn ? make_desc(smem_u32(sA) + s * 128, 128, 2048) : make_desc(smem_u32(sA) + s * 2 * 2048, 2048, 128);
        uint64_t bd = b_mn ? make_desc(smem_u32(sB) + s * 128, 128, 2048) : make_desc(smem_u32(sB) + s * 2 * 256, 256, 128);
        mma_tf32(tmem_d, ad, bd, idesc, s > 0);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(32));
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00001000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int smem = 32 * 2048 + 4 * 2048 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 7; ++mode) {
    const int ka = (mode == 0 || mode == 5) ? 16 : 128;
    const int nA = 128 * ka, nB = (mode == 0 || mode == 5) ? 16 * 16 : 128 * 16;
    std::vector<float> A(nA), B(nB), D(128 * 16);
    srand(7 + mode);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, nA * 4)); CK(cudaMalloc(&dB, nB * 4)); CK(cudaMalloc(&dD, 128 * 16 * 4));
    CK(cudaMemcpy(dA, A.data(), nA * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), nB * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * 16 * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, 128 * 16 * 4, cudaMemcpyDeviceToHost));
    double e_full = 0, e_tr = 0, e_rn = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        double f = 0, t = 0, r = 0;
        const bool small = mode == 0 || mode == 5;
        const int K = small ? 16 : 128;
        for (int k = 0; k < K; ++k) {
          float a = small ? A[m * 16 + k] : A[k * 128 + m];
          float b = small ? B[n * 16 + k] : B[k * 16 + n];
          f += (double)a * b;
          t += (double)trunc_tf32(a) * trunc_tf32(b);
          r += (double)round_tf32(a) * round_tf32(b);
        }
        double d = D[m * 16 + n];
        e_full = fmax(e_full, fabs(d - f)); e_tr = fmax(e_tr, fabs(d - t)); e_rn = fmax(e_rn, fabs(d - r));
      }
    if (mode == 1) {
      FILE* f = fopen("gpurun_out/probe_mode1.bin", "wb");
      if (f) { fwrite(A.data(), 4, nA, f); fwrite(B.data(), 4, nB, f); fwrite(D.data(), 4, 128 * 16, f); fclose(f); }
    }
    printf("mode %d: max|D - fp32 ref| = %.3e   |D - trunc-tf32 ref| = %.3e   |D - rn-tf32 ref| = %.3e   D[0][0..3] = %f %f %f %f\n",
           mode, e_full, e_tr, e_rn, D[0], D[1], D[2], D[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
