// First layer (Cin = 1) on the tensor cores: the 27 taps of a 3x3x3 kernel (9 of a 1x3x3 one) become the K
// dimension of a tcgen05.mma, K = 32 (16), N = 16 output channels, M = 128 positions of one output row.
//
// Why: on CUDA cores (conv_small_cin_kernel) the layer is issue-bound at ~22 TFLOP/s and takes 10 % of a forward
// although it holds 0.8 % of the FLOPs (profiles/README.md).  Here the arithmetic is two 128 x 16 x 16 MMAs per output
// row; the CUDA cores only build the im2col tile:
//   producers (warps 0-3, thread = output column x): keep a ring of 5 input rows x 3 columns per z plane in registers
//     and slide it along y, so an output row costs 9 two-byte loads per thread (3 new input rows x {x-1, x, x+1})
//     issued one row period before their first use, ~20 packing instructions and 4 (2) 16-byte shared-memory stores:
//     element k = (kz*3 + ky)*3 + kx of position x goes to group k / 8 of the planar UMMA operand layout
//     [8-k group][position][16 B] (SBO 128 B, LBO 2048 B);
//   MMA warp (warp 4): K/16 MMAs per row into a ring of 16 TMEM slots (16 columns each), one thread (elect.sync);
//   epilogue (warps 5-8, TMEM lane quarter = warp % 4): bias, InstanceNorm sums of the rounded values (fp32 in
//     registers, one fp64 atomic per channel per sample per warp), fp16 channels-last store.
// Two CTAs per SM (256 TMEM columns, 80 KB of shared memory each: the ring is sized so that a third CTA cannot become
// resident and block in tcgen05.alloc) overlap each other's barrier latencies.
// The source may carry a pending transform (scale / shift / LeakyReLU per sample, applied at load time: slow path);
// out-of-image taps are zero AFTER it.
// Measured (ncu, 32 patches of 128^3): 1.51 ms against 2.60 ms for conv_small_cin_kernel; with the transform applied
// at load time every load exposed its latency (clock64: 3060 of 3500 cycles per row in load_row) and the kernel was no
// faster than the CUDA-core one.  launch_conv_specialised() selects it unless FNNU_FIRST_LAYER_TC=0.
#include <type_traits>
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

constexpr int kFirstThreads = 288;
constexpr int kFirstStages = 10;
constexpr int kFirstSlots = 16;

struct FirstArgs {
  ConvArgs a;
  int n_units;   // batch * D output planes
};

bool first_umma_supported(const ConvArgs& a) {
  if (a.transposed || a.cin != 1 || a.cout_pad != 16) return false;
  if (a.k[1] != 3 || a.k[2] != 3 || (a.k[0] != 1 && a.k[0] != 3)) return false;
  for (int i = 0; i < 3; ++i) {
    if (a.s[i] != 1 || a.pad[i] != (a.k[i] - 1) / 2 || a.in_d[i] != a.out_d[i]) return false;
  }
  return a.in_d[2] <= 128 && a.w != nullptr;
}

__device__ __forceinline__ void first_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int n = 8, off = 16; n >= 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = hi ? v[i] : v[i + n];
      const float keep = hi ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int NKZ>   // 1 or 3 z taps
__global__ void __launch_bounds__(kFirstThreads, 2) conv_first_umma_kernel(const __grid_constant__ FirstArgs p) {
  constexpr int K = NKZ == 3 ? 32 : 16;
  constexpr int G = K / 8;                       // 16-byte groups per position
  constexpr int STAGE_BYTES = G * 128 * 16;
  constexpr int PZ = (NKZ - 1) / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = p.a;
  uint8_t* ring = smem;
  uint8_t* b_s = smem + kFirstStages * STAGE_BYTES;                       // [K/16][2][16][8 halves]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(b_s + (K / 16) * 512);   // [kFirstStages]
  uint64_t* empty_bar = full_bar + kFirstStages;
  uint64_t* tfull_bar = empty_bar + kFirstStages;                        // [kFirstSlots]
  uint64_t* tempty_bar = tfull_bar + kFirstSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kFirstSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.in_d[0], H = a.in_d[1], W = a.in_d[2];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFirstStages; ++s) {
      mbar_init(&full_bar[s], 4);      // one arrival per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kFirstSlots; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);    // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights: fp32 [tap][cout_pad] -> fp16 K-major UMMA tile, taps beyond NKZ * 9 are zero
  for (int i = threadIdx.x; i < K * 16; i += kFirstThreads) {
    const int k = i >> 4, n = i & 15;
    const float v = (k < NKZ * 9 && n < a.cout) ? __ldg(a.w + (size_t)k * a.cout_pad + n) : 0.f;
    *reinterpret_cast<__half*>(b_s + (k >> 4) * 512 + ((k >> 3) & 1) * 256 + n * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== PRODUCERS: thread = output column ===========================
    const int x = threadIdx.x;
    const size_t row_elems = (size_t)W * a.src_cs;
    int stage = 0;
    uint32_t phase = 0;
    int cur_b = -1;
    float sc = 1.f, sh = 0.f, sl = 1.f;
    // IDENT: the source has no pending transform (the network input).  Then a load has no dependent arithmetic and
    // really is one row period ahead of its first use; with the transform applied at load time every one of the 9
    // loads of a row exposed its latency (clock64: 3060 of 3500 cycles per row in load_row).
    auto produce = [&](auto ident_tag) {
    constexpr bool IDENT = decltype(ident_tag)::value;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int b = u / D, z = u - b * D;
      if (b != cur_b) {
        const ChanMeta m = a.src_meta[0];
        xform_from_stats(a.src_stats + (size_t)b * a.src_stat_stride * 2, m, a.src_inv_count, sc, sh);
        sl = m.eps < 0.f ? 1.f : m.slope;
        cur_b = b;
      }
      const __half* vol = a.src + (size_t)b * D * H * row_elems;
      // rows[kz][slot][kx]: a ring of 5 input rows of plane z+kz-PZ at columns x-1+kx (transformed, 0 outside the
      // image); input row r lives in slot (r + 1) % 5.  Output row y uses rows y-1, y, y+1 and its iteration starts
      // the loads of row y+2, which are first read one row period later.
      __half rows[NKZ][5][3];
      auto load_row = [&](int kz, int yy, __half (&o)[3]) {
        const int zz = z + kz - PZ;
        const bool row_ok = zz >= 0 && zz < D && yy >= 0 && yy < H;
        const __half* row = vol + ((size_t)(row_ok ? zz : 0) * H + (row_ok ? yy : 0)) * row_elems;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          __half hv = __float2half_rn(0.f);
          if (row_ok && xx >= 0 && xx < W) {
            const __half raw = __ldg(row + (size_t)xx * a.src_cs);
            hv = IDENT ? raw : __float2half_rn(lrelu(fmaf(__half2float(raw), sc, sh), sl));
          }
          o[kx] = hv;
        }
      };
#pragma unroll
      for (int kz = 0; kz < NKZ; ++kz) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) rows[kz][0][kx] = __float2half_rn(0.f);   // row -1
        load_row(kz, 0, rows[kz][1]);
        load_row(kz, 1, rows[kz][2]);
      }
      for (int yb = 0; yb < H; yb += 5) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int y = yb + j;
          if (y < H) {
#pragma unroll
            for (int kz = 0; kz < NKZ; ++kz) load_row(kz, y + 2, rows[kz][(j + 3) % 5]);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            // k = (kz * 3 + ky) * 3 + kx; group g holds taps 8 g .. 8 g + 7
            __half taps[K];
#pragma unroll
            for (int k = 0; k < K; ++k) taps[k] = __float2half_rn(0.f);
#pragma unroll
            for (int kz = 0; kz < NKZ; ++kz)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) taps[(kz * 3 + ky) * 3 + kx] = rows[kz][(j + ky) % 5][kx];
            uint8_t* dst = ring + (size_t)stage * STAGE_BYTES + (size_t)x * 16;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              uint4 v;
              __half2 h0 = __halves2half2(taps[8 * g + 0], taps[8 * g + 1]);
              __half2 h1 = __halves2half2(taps[8 * g + 2], taps[8 * g + 3]);
              __half2 h2 = __halves2half2(taps[8 * g + 4], taps[8 * g + 5]);
              __half2 h3 = __halves2half2(taps[8 * g + 6], taps[8 * g + 7]);
              v.x = *reinterpret_cast<uint32_t*>(&h0);
              v.y = *reinterpret_cast<uint32_t*>(&h1);
              v.z = *reinterpret_cast<uint32_t*>(&h2);
              v.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst + (size_t)g * 2048) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_warp(&full_bar[stage]);
            if (++stage == kFirstStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    };
    if (a.src_meta[0].eps < 0.f) produce(std::true_type{}); else produce(std::false_type{});
  } else if (warp == 4) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(ring), 2048, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(b_s), 256, 128);
    int stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      for (int y = 0; y < H; ++y) {
        mbar_wait(&tempty_bar[slot], sphase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)slot * 16u;
          const uint64_t da = a_desc0 + (uint64_t)((uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4));
#pragma unroll
          for (int c = 0; c < K / 16; ++c)
            umma_f16(d, da + (uint64_t)(c * ((2 * 2048) >> 4)), b_desc0 + (uint64_t)(c * (512 >> 4)), idesc, c > 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[slot]);
        }
        __syncwarp();
        if (++stage == kFirstStages) { stage = 0; phase ^= 1; }
        if (++slot == kFirstSlots) { slot = 0; sphase ^= 1; }
      }
    }
  } else {
    // =========================== EPILOGUE ===========================
    const int wq = warp & 3;
    const int x = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const bool vec_store = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0) && a.cout == 16;
    const bool col_ok = x < W;
    float bias[16], s1[16], s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      bias[j] = (a.bias && j < a.cout) ? __ldg(a.bias + j) : 0.f;
      s1[j] = s2[j] = 0.f;
    }
    auto flush_stats = [&](int b) {
      if (!a.dst_stats || b < 0) return;
      first_reduce16(s1, lane);
      first_reduce16(s2, lane);
      const int j = lane >> 1;
      if (!(lane & 1) && j < a.cout) {
        atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + j) * 2 + 0, (double)s1[0]);
        atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + j) * 2 + 1, (double)s2[0]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) s1[i] = s2[i] = 0.f;
    };
    int slot = 0;
    uint32_t sphase = 0;
    int cur_b = -1;
    const size_t out_row = (size_t)W * a.dst_cs;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int b = u / D, z = u - b * D;
      if (b != cur_b) {
        flush_stats(cur_b);
        cur_b = b;
      }
      __half* out_px = a.dst + (((size_t)b * D + z) * H) * out_row + (size_t)x * a.dst_cs;
      for (int y = 0; y < H; ++y, out_px += out_row) {
        mbar_wait(&tfull_bar[slot], sphase);
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(tmem_base + lane_off + (uint32_t)slot * 16u, acc);
        if (col_ok) {
          __half2 hv[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float v0 = __uint_as_float(acc[j]) + bias[j];
            const float v1 = __uint_as_float(acc[j + 1]) + bias[j + 1];
            hv[j >> 1] = __floats2half2_rn(v0, v1);
            // sums of the ROUNDED values, as conv_small_cin_kernel (the kernel this one replaces) keeps them
            const float2 r = __half22float2(hv[j >> 1]);
            s1[j] += r.x;
            s2[j] = fmaf(r.x, r.x, s2[j]);
            s1[j + 1] += r.y;
            s2[j + 1] = fmaf(r.y, r.y, s2[j + 1]);
          }
          if (vec_store) {
            reinterpret_cast<uint4*>(out_px)[0] = *reinterpret_cast<uint4*>(&hv[0]);
            reinterpret_cast<uint4*>(out_px)[1] = *reinterpret_cast<uint4*>(&hv[4]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < a.cout) out_px[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&tempty_bar[slot]);
        if (++slot == kFirstSlots) { slot = 0; sphase ^= 1; }
      }
    }
    flush_stats(cur_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

int launch_conv_first_umma(const ConvArgs& a, cudaStream_t s) {
  if (!first_umma_supported(a)) {
    set_error("conv_first_umma: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  FirstArgs p;
  p.a = a;
  p.n_units = a.batch * a.in_d[0];
  const int K = a.k[0] == 3 ? 32 : 16;
  // the ring is padded to > 227 KB / 3 so that at most two CTAs (2 x 256 TMEM columns) are resident per SM
  size_t smem = (size_t)kFirstStages * (K / 8) * 128 * 16 + (size_t)(K / 16) * 512 + (2 * kFirstStages + 2 * kFirstSlots) * 8 + 64;
  if (smem < 78 * 1024) smem = 78 * 1024;
  int grid = p.n_units < 2 * num_sms() ? p.n_units : 2 * num_sms();
  if (a.k[0] == 3) {
    /* the attribute is per device: set it on every launch (cheap) */ FNNU_CUDA(cudaFuncSetAttribute(conv_first_umma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    conv_first_umma_kernel<3><<<grid, kFirstThreads, smem, s>>>(p);
  } else {
    /* the attribute is per device: set it on every launch (cheap) */ FNNU_CUDA(cudaFuncSetAttribute(conv_first_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    conv_first_umma_kernel<1><<<grid, kFirstThreads, smem, s>>>(p);
  }
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
