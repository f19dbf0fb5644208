// Stand-alone probe of the tcgen05 building blocks conv_umma.cu relies on (sm_100a):
//   * SWIZZLE_NONE K-major smem descriptors over a "position-planar" operand layout
//     ([k-chunk of 8 halves][row][8 halves], SBO = 128 B, LBO = plane stride),
//   * an A start address shifted by whole rows (16 B) that is NOT 128-byte aligned (the kx tap shift),
//   * accumulation over several K=16 steps, N in {16, 32, 64, 160}, TMEM 32x32b loads.
// Prints max |error| against a CPU reference for each case; exit code 0 iff all pass.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma probe_umma.cu && ./probe_umma
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (sm_100)
  return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// A: global [rows_total][K] halves (row-major), B: global [N][K]; C: [128][N] floats.
// C[m][n] = sum_k A[shift + m][k] * B[n][k]
__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                    float* __restrict__ C, int rows_total, int K, int N, int shift) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int kchunks = K / 8;
  const uint32_t a_plane = (uint32_t)rows_total * 16;   // bytes per 8-channel plane of A
  const uint32_t b_plane = (uint32_t)N * 16;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)kchunks * a_plane;
  // fill smem (generic proxy): element (row, k) -> plane k/8, row*16 + (k%8)*2
  for (int i = threadIdx.x; i < rows_total * kchunks; i += blockDim.x) {
    int row = i % rows_total, ch = i / rows_total;
    uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)row * K + ch * 8);
    *reinterpret_cast<uint4*>(a_s + (size_t)ch * a_plane + row * 16) = v;
  }
  for (int i = threadIdx.x; i < N * kchunks; i += blockDim.x) {
    int row = i % N, ch = i / N;
    uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)row * K + ch * 8);
    *reinterpret_cast<uint4*>(b_s + (size_t)ch * b_plane + row * 16) = v;
  }
  const int warp = threadIdx.x >> 5;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t da = make_desc(smem_u32(a_s) + (uint32_t)(2 * ks) * a_plane + (uint32_t)shift * 16, a_plane, 128);
      uint64_t db = make_desc(smem_u32(b_s) + (uint32_t)(2 * ks) * b_plane, b_plane, 128);
      uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = threadIdx.x;   // TMEM lane == warp*32 + lane == threadIdx.x
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) C[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

static int run_case(int K, int N, int shift, int rows_total) {
  std::vector<__half> hA((size_t)rows_total * K), hB((size_t)N * K);
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = __float2half((float)((int)(rand() % 17) - 8) / 8.f);
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = __float2half((float)((int)(rand() % 13) - 6) / 4.f);
  __half *dA, *dB;
  float* dC;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dC, (size_t)128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, (size_t)128 * N * 4);
  size_t smem = (size_t)(K / 8) * rows_total * 16 + (size_t)(K / 8) * N * 16;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dC, rows_total, K, N, shift);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("K=%d N=%d shift=%d: CUDA error %s\n", K, N, shift, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hC((size_t)128 * N);
  cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)__half2float(hA[(size_t)(shift + m) * K + k]) * (double)__half2float(hB[(size_t)n * K + k]);
      double err = fabs(ref - (double)hC[(size_t)m * N + n]);
      if (!(err <= maxerr)) maxerr = err;   // catches NaN too
    }
  printf("K=%3d N=%3d shift=%2d : max|err| = %.3g %s\n", K, N, shift, maxerr, maxerr < 1e-3 ? "OK" : "FAIL");
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dC);
  return maxerr < 1e-3 ? 0 : 1;
}

int main() {
  int fails = 0;
  fails += run_case(16, 16, 0, 136);
  fails += run_case(16, 16, 1, 136);
  fails += run_case(16, 16, 3, 136);
  fails += run_case(32, 16, 5, 136);
  fails += run_case(64, 32, 7, 136);
  fails += run_case(32, 64, 2, 136);
  fails += run_case(32, 160, 1, 136);
  fails += run_case(48, 256, 4, 136);
  printf(fails ? "PROBE FAILED (%d cases)\n" : "PROBE OK\n", fails);
  return fails ? 1 : 0;
}
