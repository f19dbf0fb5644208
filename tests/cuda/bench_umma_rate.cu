// Micro-benchmark: how many cycles does one tcgen05.mma (kind::f16, cta_group::1, A and B from shared memory)
// occupy the tensor pipe for, as a function of M, N, the operand layout and what else uses shared memory?
// One CTA per SM; thread 0 issues `reps` groups of 9 MMAs (distinct start addresses, like the 9 (kz,kx) taps of
// conv_umma_rows.cu), commits once and waits; optional "noise" warps hammer shared memory with LDS/STS to imitate the
// producers.  Numbers only (operands are whatever is in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bench_umma_rate bench_umma_rate.cu && ./bench_umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
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

struct Params {
  int M, N, reps, noise_warps, layout;   // layout: 0 = SWIZZLE_NONE planar, 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
  uint32_t a_lbo, a_sbo, a_tap_stride;   // bytes
  int same_tile;                         // 1: all MMAs accumulate into one TMEM tile, 0: rotate over tiles
  int commits;                           // tcgen05.commit after every group of 9 MMAs (0, 1 or 2 of them)
  int per_group;                         // MMAs per group (1..9)
  int spin_warps;                        // warps polling an mbarrier that never completes (like idle pipeline roles)
  int random_data;                       // fill the operands with pseudo-random fp16 instead of zeros
  int issuers;                           // 1, 2 or 4 warps issue concurrently (own TMEM tiles)
  int waits;                             // try_wait on an already-completed mbarrier before every group (0..4)
};

template <int PG, bool ELECT>
__global__ void __launch_bounds__(1024) rate_kernel(Params p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint64_t side_bar[2];
  __shared__ uint64_t never_bar;
  __shared__ uint64_t done_bar;   // phase 0 already complete
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    stop = 0;
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&side_bar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&side_bar[1])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never_bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 220 * 1024 / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p.random_data) {   // fp16 values in (-2, 2): exponent bits 0x3800..0x3f00 | random mantissa / sign
      uint32_t h = (uint32_t)i * 2654435761u;
      v.x = ((h & 0x83ff83ffu) | 0x38003800u);
      h = h * 1664525u + 1013904223u;
      v.y = ((h & 0x83ff83ffu) | 0x3c003c00u);
      h = h * 1664525u + 1013904223u;
      v.z = ((h & 0x83ff83ffu) | 0x34003400u);
      h = h * 1664525u + 1013904223u;
      v.w = ((h & 0x83ff83ffu) | 0x38003800u);
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool leader;
  if (ELECT) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    leader = pred != 0;
  } else {
    leader = (threadIdx.x & 31) == 0;
  }
  if (warp < p.issuers) {
    if (leader) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem);
    const uint32_t b0 = smem_u32(smem) + 128 * 1024;
    const int slots = p.same_tile ? 1 : 512 / p.N / p.issuers;
    long long t0 = clock64();
    int slot = 0;
    for (int r = 0; r < p.reps; ++r) {
      const uint32_t d = tmem + (uint32_t)(warp * slots + slot) * p.N;
      if (++slot == slots) slot = 0;
      for (int w = 0; w < (p.waits & 7); ++w) {
        if (p.waits & 8) mbar_wait_relaxed(smem_u32(&done_bar), 0); else mbar_wait(smem_u32(&done_bar), 0);
      }
      if (p.waits) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int t = 0; t < PG; ++t) {
        const uint64_t da = make_desc(a0 + (uint32_t)(t / 3) * 3 * p.a_tap_stride * 64 + (uint32_t)(t % 3) * p.a_tap_stride, p.a_lbo, p.a_sbo, p.layout);
        const uint64_t db = make_desc(b0 + (uint32_t)t * 2 * p.N * 16, (uint32_t)p.N * 16, 128, 0);
        const uint32_t acc = t > 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
      for (int cm = 0; cm < p.commits; ++cm)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&side_bar[cm])) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    mbar_wait(smem_u32(&bar[warp]), 0);
    long long t2 = clock64();
    if (warp == 0) stop = 1;
    if (blockIdx.x == 0 && warp == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    }
  } else if (warp >= 4 && warp < 4 + p.noise_warps) {
    // shared-memory noise: LDS.128 + STS.128 in place on a private 16 KB window, like the in-place normalise
    uint4* w = reinterpret_cast<uint4*>(smem + 64 * 1024) + (size_t)(warp - 4) * 128 + (threadIdx.x & 31);
    uint4 acc = make_uint4(0, 0, 0, 0);
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v = w[i * 32];
        acc.x ^= v.x;
        v.y += 1;
        w[i * 32] = v;
      }
    }
    if (acc.x == 0x12345) out[2] = 1;
  } else if (warp >= 4 + p.noise_warps && warp < 4 + p.noise_warps + p.spin_warps) {
    // idle roles: poll a barrier that is still in phase 0 until the issuer is done
    const uint32_t addr = smem_u32(&never_bar);
    while (!stop) {
      uint32_t done;
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(addr), "r"(0)
          : "memory");
      if (done) break;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int PG, bool ELECT>
static void launch(Params p, long long* d) {
  cudaFuncSetAttribute(rate_kernel<PG, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  rate_kernel<PG, ELECT><<<148, 1024, 224 * 1024>>>(p, d);
}

static void run(const char* name, Params p, bool elect = false) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  if (p.issuers == 0) p.issuers = 1;
  if (elect) {
    if (p.per_group == 9) launch<9, true>(p, d); else if (p.per_group == 3) launch<3, true>(p, d); else launch<1, true>(p, d);
  } else {
    if (p.per_group == 9) launch<9, false>(p, d); else if (p.per_group == 3) launch<3, false>(p, d); else launch<1, false>(p, d);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n = (double)p.per_group * p.reps;
  printf("%-44s M=%3d N=%3d %s issuers=%d : issue %.1f cyc/MMA/issuer, issue+drain %.1f\n", name, p.M, p.N, elect ? "elect" : "lane0", p.issuers,
         h[0] / n, h[1] / n);
  cudaFree(d);
}

int main() {
  const int reps = 2000;
  // planar SWIZZLE_NONE, as conv_umma_rows.cu: LBO = plane stride 140 positions * 16 B, SBO = 128 B, tap shift 16 B
  for (int N : {16, 48, 96, 128, 256}) run("planar none lbo=2240", Params{128, N, reps, 0, 0, 2240, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  run("planar none lbo=2240, same tile", Params{128, 48, reps, 0, 0, 2240, 128, 16, 1, 0, 9, 0, 0, 1, 0});
  run("planar none lbo=2048", Params{128, 48, reps, 0, 0, 2048, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  run("planar none lbo=2304 (18*128)", Params{128, 48, reps, 0, 0, 2304, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  run("planar none lbo=2112 (2048+64)", Params{128, 48, reps, 0, 0, 2112, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  run("planar none lbo=2176 (17*128)", Params{128, 48, reps, 0, 0, 2176, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  run("planar none, tap shift 128 B", Params{128, 48, reps, 0, 0, 2240, 128, 128, 0, 0, 9, 0, 0, 1, 0});
  run("planar none, tap shift 0", Params{128, 48, reps, 0, 0, 2240, 128, 0, 0, 0, 9, 0, 0, 1, 0});
  run("none, K-chunks adjacent (lbo=128,sbo=256)", Params{128, 48, reps, 0, 0, 128, 256, 32, 0, 0, 9, 0, 0, 1, 0});
  run("swizzle 32B (sbo=256)", Params{128, 48, reps, 0, 6, 16, 256, 32, 0, 0, 9, 0, 0, 1, 0});
  run("swizzle 128B (sbo=1024)", Params{128, 48, reps, 0, 2, 16, 1024, 32, 0, 0, 9, 0, 0, 1, 0});
  run("M=64 planar none", Params{64, 48, reps, 0, 0, 2240, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  for (int nw : {4, 8, 16}) run("planar none + smem noise", Params{128, 48, reps, nw, 0, 2240, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  for (int nw : {8, 16}) run("planar none N=96 + smem noise", Params{128, 96, reps, nw, 0, 2240, 128, 16, 0, 0, 9, 0, 0, 1, 0});
  for (int cm : {1, 2}) {
    Params q{128, 48, reps, 0, 0, 2240, 128, 16, 0, cm, 9, 0, 0, 1, 0};
    run(cm == 1 ? "9 MMAs + 1 commit per group" : "9 MMAs + 2 commits per group", q);
    q.per_group = 1;
    run(cm == 1 ? "1 MMA + 1 commit per group" : "1 MMA + 2 commits per group", q);
    q.per_group = 3;
    run(cm == 1 ? "3 MMAs + 1 commit per group" : "3 MMAs + 2 commits per group", q);
  }
  for (int sw : {4, 12, 20}) {
    Params q{128, 48, reps, 0, 0, 2240, 128, 16, 0, 2, 9, sw, 0, 1, 0};
    char name[64];
    snprintf(name, sizeof(name), "9 MMAs + 2 commits, %d warps polling", sw);
    run(name, q);
    q.commits = 0;
    snprintf(name, sizeof(name), "9 MMAs, no commit, %d warps polling", sw);
    run(name, q);
  }
  for (int N : {48, 96, 256}) {
    Params q{128, N, reps, 0, 0, 2240, 128, 16, 0, 0, 9, 0, 1, 1, 0};
    run("random operands, no commit", q);
    q.commits = 2;
    run("random operands, 2 commits per group", q);
  }
  for (int N : {16, 48, 96}) {
    for (int iss : {1, 2, 4}) {
      Params q{128, N, reps, 0, 0, 2240, 128, 16, 0, 0, 9, 0, 1, iss, 0};
      run("concurrent issuers", q, false);
      run("concurrent issuers", q, true);
    }
  }
  {
    Params q{128, 48, reps, 0, 0, 2240, 128, 16, 0, 2, 9, 0, 1, 1, 0};
    run("2 commits per 9 MMAs", q, true);
    q.commits = 1;
    run("1 commit per 9 MMAs", q, true);
    q.issuers = 2;
    run("1 commit per 9 MMAs", q, true);
  }
  for (int w : {0, 1, 2, 4, 9, 10}) {
    Params q{128, 48, reps, 0, 0, 2240, 128, 16, 0, 2, 9, 0, 1, 1, w};
    char name[64];
    snprintf(name, sizeof(name), "9 MMAs + 2 commits + %d %s waits", w & 7, (w & 8) ? "relaxed" : "acquire");
    run(name, q, true);
  }
  return 0;
}
