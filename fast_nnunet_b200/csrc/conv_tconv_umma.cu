// Transposed convolution with kernel == stride (the decoder's up-sampling step) as ONE GEMM per 128 input voxels:
//   D[voxel][n = tap * CP + co] = sum_ci  A[voxel][ci] * W[ci][tap][co]           tap = (kz, ky, kx), kx fastest
// every output voxel (z sz + kz, y sy + ky, x sx + kx) receives exactly one tap of one input voxel, so there are no
// halos, no accumulation across tiles and no InstanceNorm sums (nnU-Net puts no norm behind the transposed conv): the
// layer is a stream of  cin  fp16 in  ->  ntaps * cout  fp16 out  per voxel and is bound by the HBM write of the
// 8x larger output.  The generic implicit-GEMM kernel (conv_umma.cu) ran it at 1.28 ms per 32 student patches
// (32 -> 16 channels, 64^3 -> 128^3; floor 0.42 ms): it tiles the virtual 128 output channels in groups of 16 and
// re-stages the input for each.  Here:
//   producers (warps 0-3): 16-byte loads of the input voxels' channels, normalise + LeakyReLU of the pending
//     InstanceNorm on the fly, st.shared into the K-major no-swizzle A tile [8-channel group][voxel][16 B]
//   MMA warp (warp 4): cin / 16 MMAs of 128 x NB x 16 per pass (NB = min(ntaps * CP, 256) TMEM columns; layers with
//     more virtual channels take two passes over the same A tile), one commit per pass
//   epilogue (warps 5-12): tcgen05.ld of 16 columns = 16 channels of one tap, + bias, fp16, two 16-byte stores to the
//     tap's output voxel; the two sets of four warps split a pass's columns
// Weights (<= 128 KB as fp16) stay in shared memory for the whole kernel, converted from the engine's fp32
// [tap][cin][cout_pad] copy in the prologue.
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

namespace {

constexpr int kTcThreads = 13 * 32;
constexpr int kTcMaxStages = 8;
constexpr int kTcSmemLimit = 225 * 1024;

struct TcCfg {
  int groups;          // cin / 8
  int chunks;          // cin / 16
  int ntaps;
  int n_total;         // ntaps * cout_pad (virtual output channels)
  int nb;              // TMEM columns of one pass
  int n_pass;
  int slots;
  int tmem_cols;
  int stages;
  int stage_bytes;
  int w_bytes;
  int smem_bytes;
  int ctas_per_sm;
  int tiles_per_sample;
  int n_tiles;
  int tiles_per_cta;
};

struct TcArgs {
  ConvArgs a;
  TcCfg c;
};

bool plan_tconv(const ConvArgs& a, TcCfg& c) {
  memset(&c, 0, sizeof(c));
  if (!a.transposed || a.dst_stats) return false;
  for (int i = 0; i < 3; ++i)
    if (a.k[i] != a.s[i] || (a.s[i] != 1 && a.s[i] != 2) || a.out_d[i] != a.in_d[i] * a.s[i]) return false;
  if (a.cin != 32 && a.cin != 64 && a.cin != 128) return false;      // 8 * groups divides the 128 producer threads
  if (a.src_cs % 8 != 0 || ((uintptr_t)a.src % 16) != 0) return false;
  if (a.dst_cs % 8 != 0 || ((uintptr_t)a.dst % 16) != 0) return false;
  if (a.cout_pad != 16 && a.cout_pad != 32 && a.cout_pad != 64 && a.cout_pad != 128) return false;
  const long long nvox = (long long)a.in_d[0] * a.in_d[1] * a.in_d[2];
  if (nvox % 128 != 0 || nvox < 512) return false;      // tiles never straddle samples
  c.groups = a.cin / 8;
  c.chunks = a.cin / 16;
  c.ntaps = a.s[0] * a.s[1] * a.s[2];
  if (c.ntaps < 2) return false;
  c.n_total = c.ntaps * a.cout_pad;
  if (c.n_total > 512 || c.n_total < 32) return false;
  c.nb = c.n_total > 256 ? 256 : c.n_total;
  c.n_pass = c.n_total / c.nb;
  c.w_bytes = a.cin * c.n_total * 2;
  c.stage_bytes = a.cin * 256;
  const int misc = 1024 + 32 * 8 + 128 * 4;
  if (c.w_bytes + 3 * c.stage_bytes + misc > kTcSmemLimit) return false;
  // two CTAs per SM (256 TMEM columns each) when a CTA's share of shared memory holds the weights and four stages
  c.ctas_per_sm = (c.nb <= 128 && c.w_bytes + 4 * c.stage_bytes + misc <= 110 * 1024) ? 2 : 1;
  c.tmem_cols = c.ctas_per_sm == 2 ? 256 : 512;
  c.slots = c.tmem_cols / c.nb;
  if (c.slots > 4) c.slots = 4;
  const int budget = (c.ctas_per_sm == 2 ? 110 * 1024 : kTcSmemLimit) - misc - c.w_bytes;
  c.stages = budget / c.stage_bytes;
  if (c.stages > kTcMaxStages) c.stages = kTcMaxStages;
  c.smem_bytes = c.w_bytes + c.stages * c.stage_bytes + misc;
  if (c.ctas_per_sm == 2 && c.smem_bytes < 78 * 1024) c.smem_bytes = 78 * 1024;     // never three CTAs: 3 x 256 columns
  if (c.ctas_per_sm == 1 && c.smem_bytes < 116 * 1024) c.smem_bytes = 116 * 1024;
  c.tiles_per_sample = (int)(nvox / 128);
  c.n_tiles = a.batch * c.tiles_per_sample;
  const int ctas = c.ctas_per_sm * num_sms();
  c.tiles_per_cta = (c.n_tiles + ctas - 1) / ctas;
  return true;
}

__device__ __forceinline__ uint4 tc_ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void tc_sts16(uint32_t addr, const uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// y = lrelu((x - m) * s + t) on 8 fp16 channels (the form conv_umma_zrows.cu uses: m is the fp16-rounded mean)
__device__ __forceinline__ uint4 tc_xform8(const uint4 raw, const __half2* m2, const __half2* s2, const __half2* t2, const __half2* l2) {
  const __half2* x = reinterpret_cast<const __half2*>(&raw);
  uint4 o;
  __half2* y = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 v = __hfma2(__hsub2(x[e], m2[e]), s2[e], t2[e]);
    y[e] = __hmax2(v, __hmul2(v, l2[e]));
  }
  return o;
}

template <int GROUPS>   // cin / 8: 4, 8 or 16
__global__ void __launch_bounds__(kTcThreads, GROUPS == 4 ? 2 : 1) conv_tconv_umma_kernel(const __grid_constant__ TcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = p.a;
  const TcCfg& c = p.c;
  uint8_t* w_s = smem;                                               // [cin / 8][n_total][8 halves]
  uint8_t* ring = smem + c.w_bytes;                                  // [stage][cin / 8][128 voxels][16 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + c.stages * c.stage_bytes);   // [kTcMaxStages]
  uint64_t* empty_bar = full_bar + kTcMaxStages;                     // [kTcMaxStages]
  uint64_t* tfull_bar = empty_bar + kTcMaxStages;                    // [4]
  uint64_t* tempty_bar = tfull_bar + 4;                              // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4);
  int* col_off = reinterpret_cast<int*>(tmem_slot + 4);              // [n_total / 16] output offset of a 16-column group
  float* bias_s = reinterpret_cast<float*>(col_off + 32);            // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = a.in_d[1], W = a.in_d[2];
  const int Ho = a.out_d[1], Wo = a.out_d[2];
  const int CP = a.cout_pad;
  const int t_begin = (int)blockIdx.x * c.tiles_per_cta;
  const int t_end = t_begin + c.tiles_per_cta < c.n_tiles ? t_begin + c.tiles_per_cta : c.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(&full_bar[s], 4);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights: fp32 [tap][cin][cout_pad] -> fp16 [ci / 8][n = tap * CP + co][ci % 8]
  for (int i = threadIdx.x; i < a.cin * c.n_total; i += kTcThreads) {
    const int co = i % CP, ci = (i / CP) % a.cin, tap = i / (CP * a.cin);
    const int n = tap * CP + co;
    *reinterpret_cast<__half*>(w_s + ((size_t)(ci >> 3) * c.n_total + n) * 16 + (ci & 7) * 2) = __float2half_rn(__ldg(a.w + i));
  }
  if (threadIdx.x < c.n_total / 16) {
    const int n0 = threadIdx.x * 16;
    const int tap = n0 / CP, co0 = n0 - tap * CP;
    const int kx = tap % a.s[2], ky = (tap / a.s[2]) % a.s[1], kz = tap / (a.s[2] * a.s[1]);
    col_off[threadIdx.x] = ((kz * Ho + ky) * Wo + kx) * a.dst_cs + co0;
  }
  if (threadIdx.x < 128) bias_s[threadIdx.x] = (a.bias && (int)threadIdx.x < a.cout) ? __ldg(a.bias + threadIdx.x) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)c.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== PRODUCERS ===========================
    // item = (voxel, 8-channel group); a quarter warp takes 8 consecutive voxels of one group (conflict-free 128-byte
    // shared-memory rows) and a warp's 32 lanes cover 8 voxels x 4 groups = whole 32-byte sectors of global memory.
    // 128 threads stride the items by 128 = a multiple of 8 * GROUPS, so a thread's group never changes.
    const int tid = threadIdx.x;
    const int r = tid % (8 * GROUPS);
    const int g = r >> 3;
    const int v0 = (tid / (8 * GROUPS)) * 8 + (r & 7);        // + k * (128 / GROUPS)
    constexpr int VSTEP = 128 / GROUPS;
    const uint32_t st0 = smem_u32(ring) + (uint32_t)(g * 128 + v0) * 16u;
    const uint32_t goff0 = (uint32_t)(v0 * a.src_cs + g * 8) * 2u;
    const uint32_t gstep = (uint32_t)(VSTEP * a.src_cs) * 2u;
    const uint32_t tile_bytes = (uint32_t)(128 * a.src_cs) * 2u;
    const long long sample_bytes = (long long)c.tiles_per_sample * tile_bytes;
    __half2 m2[4], s2[4], t2[4], l2[4];
    int cur_b = -1, stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / c.tiles_per_sample, ts = t - b * c.tiles_per_sample;
      if (b != cur_b) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float mh[2], sc[2], sh[2], sl[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = g * 8 + 2 * e + k;
            const ChanMeta m = a.src_meta[ch];
            if (m.eps < 0.f) {
              mh[k] = 0.f; sc[k] = 1.f; sh[k] = 0.f; sl[k] = 1.f;
            } else {
              const double* st = a.src_stats + ((size_t)b * a.src_stat_stride + ch) * 2;
              const double mean = st[0] * a.src_inv_count;
              double var = st[1] * a.src_inv_count - mean * mean;
              if (var < 0.0) var = 0.0;
              const float scale = m.gamma * (float)(1.0 / sqrt(var + (double)m.eps));
              mh[k] = __half2float(__float2half_rn((float)mean));
              sc[k] = scale;
              sh[k] = m.beta - (float)(mean - (double)mh[k]) * scale;
              sl[k] = m.slope;
            }
          }
          m2[e] = __floats2half2_rn(mh[0], mh[1]);
          s2[e] = __floats2half2_rn(sc[0], sc[1]);
          t2[e] = __floats2half2_rn(sh[0], sh[1]);
          l2[e] = __floats2half2_rn(sl[0], sl[1]);
        }
        cur_b = b;
      }
      const char* src = reinterpret_cast<const char*>(a.src) + (long long)b * sample_bytes + (long long)ts * tile_bytes + goff0;
      uint4 v[GROUPS];
#pragma unroll
      for (int k = 0; k < GROUPS; ++k) v[k] = tc_ldg16(src + k * gstep);
      mbar_wait(&empty_bar[stage], phase ^ 1);
#pragma unroll
      for (int k = 0; k < GROUPS; ++k)
        tc_sts16(st0 + (uint32_t)stage * (uint32_t)c.stage_bytes + (uint32_t)(k * VSTEP) * 16u, tc_xform8(v[k], m2, s2, t2, l2));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_warp(&full_bar[stage]);
      if (++stage == c.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 4) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.nb >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(ring), 2048, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(w_s), (uint32_t)c.n_total * 16u, 128);
    int stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      for (int ps = 0; ps < c.n_pass; ++ps) {
        mbar_wait(&tempty_bar[slot], sphase ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(slot * c.nb);
          const uint64_t da = a_desc0 + (uint64_t)((uint32_t)stage * (uint32_t)(c.stage_bytes >> 4));
          const uint64_t db = b_desc0 + (uint64_t)(uint32_t)(ps * c.nb);
          for (int kc = 0; kc < c.chunks; ++kc)
            umma_f16(d, da + (uint64_t)(uint32_t)(kc * 2 * 128), db + (uint64_t)(uint32_t)(kc * 2 * c.n_total), idesc, kc > 0 ? 1u : 0u);
          if (ps == c.n_pass - 1) umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[slot]);
        }
        __syncwarp();
        if (++slot == c.slots) { slot = 0; sphase ^= 1; }
      }
      if (++stage == c.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // =========================== EPILOGUE ===========================
    const int set = (warp - 5) >> 2;
    const int wq = warp & 3;
    const int v_in_tile = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int half_groups = c.nb / 32;                 // 16-column groups of a pass that this set handles
    const bool vec_ok = (a.cout % 16) == 0;
    const bool has_bias = a.bias != nullptr;
    // lanes 2i / 2i+1 are neighbours in x (W is even: 128 divides the voxel count and tiles start at multiples of 128)
    const bool paired = vec_ok && (W % 2) == 0;
    const long long partner_delta = (long long)a.s[2] * a.dst_cs * 2 * ((lane & 1) ? -1 : 1);
    const long long out_sample = (long long)a.out_d[0] * Ho * Wo * a.dst_cs;
    int slot = 0;
    uint32_t sphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / c.tiles_per_sample, ts = t - b * c.tiles_per_sample;
      const int v = ts * 128 + v_in_tile;
      const int x = v % W, y = (v / W) % H, z = v / (W * H);
      __half* out_v = a.dst + (long long)b * out_sample + ((long long)(z * a.s[0]) * Ho + y * a.s[1]) * (long long)Wo * a.dst_cs +
                      (long long)(x * a.s[2]) * a.dst_cs;
      for (int ps = 0; ps < c.n_pass; ++ps) {
        mbar_wait(&tfull_bar[slot], sphase);
        tc_fence_after();
        const int g_first = set * half_groups;
        for (int gi = 0; gi < half_groups; ++gi) {
          const int col = (g_first + gi) * 16;
          uint32_t acc[16];
          tmem_ld16(t_lane + (uint32_t)(slot * c.nb + col), acc);
          if (gi == half_groups - 1) {
            tc_fence_before();
            mbar_arrive_warp(&tempty_bar[slot]);
          }
          const int n0 = ps * c.nb + col;
          const int co0 = n0 % CP;
          if (has_bias) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + co0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 bv = b4[j];
              acc[4 * j + 0] = __float_as_uint(__uint_as_float(acc[4 * j + 0]) + bv.x);
              acc[4 * j + 1] = __float_as_uint(__uint_as_float(acc[4 * j + 1]) + bv.y);
              acc[4 * j + 2] = __float_as_uint(__uint_as_float(acc[4 * j + 2]) + bv.z);
              acc[4 * j + 3] = __float_as_uint(__uint_as_float(acc[4 * j + 3]) + bv.w);
            }
          }
          __half2 hv[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) hv[j >> 1] = __floats2half2_rn(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]));
          __half* q = out_v + col_off[n0 >> 4];
          if (paired) {
            stg32_paired(q, partner_delta, *reinterpret_cast<uint4*>(&hv[0]), *reinterpret_cast<uint4*>(&hv[4]), true, lane);
          } else if (vec_ok) {
            reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
            reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[4]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (co0 + j < a.cout) q[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
          }
        }
        if (++slot == c.slots) { slot = 0; sphase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)c.tmem_cols));
  }
}

}  // namespace

bool tconv_umma_supported(const ConvArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FNNU_TCONV_UMMA");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || a.w == nullptr) return false;
  TcCfg c;
  return plan_tconv(a, c);
}

int launch_tconv_umma(const ConvArgs& a, cudaStream_t s) {
  TcArgs p;
  p.a = a;
  if (!plan_tconv(a, p.c)) {
    set_error("conv_tconv_umma: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  const int grid = (p.c.n_tiles + p.c.tiles_per_cta - 1) / p.c.tiles_per_cta;
#define FNNU_TC_CASE(G)                                                                                                   \
  if (p.c.groups == G) {                                                                                                  \
    FNNU_CUDA(cudaFuncSetAttribute(conv_tconv_umma_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemLimit)); \
    conv_tconv_umma_kernel<G><<<grid, kTcThreads, p.c.smem_bytes, s>>>(p);                                                \
  }
  FNNU_TC_CASE(4) else FNNU_TC_CASE(8) else FNNU_TC_CASE(16)
#undef FNNU_TC_CASE
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
