// First layer (Cin = 1) on the tensor cores, second generation ("z-pair"): the (input plane, ky) taps are the K
// dimension, kx is a descriptor shift, and TWO output z-planes share every MMA.
//
// Round 1's conv_first_umma.cu built a full im2col tile per output row (27 taps per position, 8 KB of shared-memory
// stores, 9 two-byte loads per thread and row) and ran at 1.51 ms per 32 patches of 128^3 against an HBM floor of
// 0.33 ms; the MMA and epilogue warps waited ~90 % of the time for the four producer warps.  Here:
//   A[position x][k = pl * 3 + ky] = in[z0 - PZ + pl][y - 1 + ky][x]       pl < NPL = 2 + NKZ - 1 input planes, K = 16
//   B[kx][n = zt * CP + co][k]     = w[co][kz = pl - zt][ky][kx]           (0 where plane pl does not feed output zt)
//   D[x][n] = sum_kx  A[x + kx - 1][.] * B[kx][n][.]                       3 MMAs of 128 x (2 CP) x 16 per y-step
// so one y-step produces an output row of BOTH planes z0 and z0 + 1, a position's operand is 12 (6) values instead of
// 27, and the producer work per y-step is 4 (2) two-byte loads + two 16-byte stores per thread: 6.75x fewer stored
// bytes and 4.5x fewer loads per output row.  The x-1 / x+1 taps are the SAME shared-memory rows read from a start
// address shifted by 16 bytes (one position), as in conv_umma_zrows.cu.
//   producers (warps 0-3, thread = column x): a register window of rows y-1 .. y+1 per input plane, the row two steps
//     ahead already in flight; out-of-image rows / planes are zeros; the halo positions are zeroed once per kernel
//   MMA warp (warp 4): 3 MMAs + ONE commit per y-step (the barrier frees the stage and publishes the tile)
//   epilogue (warps 5-12, one set of four per output plane): tcgen05.ld, fp32 InstanceNorm sums, fp16 store
// Two CTAs per SM (256 TMEM columns each) overlap each other's barrier latencies.  The source must carry no pending
// transform (it is the gathered image); anything else runs on conv_small_cin_kernel.
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

#include <type_traits>

namespace fnnu {

namespace {

constexpr int kFzThreads = 13 * 32;      // 4 producer warps, 1 MMA warp, 8 epilogue warps
constexpr int kFzStages = 12;
constexpr int kFzProw = 140;             // positions per 8-k group: 1 + 128 + 1, padded to 4 (mod 8)
constexpr int kFzStageBytes = 2 * kFzProw * 16;
constexpr int kFzStepBars = 16;
#ifndef FNNU_FZ_AHEAD
#define FNNU_FZ_AHEAD 4
#endif
constexpr int kFzAhead = FNNU_FZ_AHEAD;   // input rows in flight per plane and thread

struct FzArgs {
  ConvArgs a;
  int n_units;     // batch * ceil(D / 2)
  int units_per_cta;
};

// CP = 16: two CTAs per SM with 256 TMEM columns each.  CP = 32: an epilogue thread carries 64 fp32 InstanceNorm sums, so
// one CTA per SM with the whole register file and 512 columns.
template <int CP, int NKZ>
__global__ void __launch_bounds__(kFzThreads, CP == 16 ? 2 : 1) conv_first_zpair_kernel(const __grid_constant__ FzArgs p) {
  constexpr int NPL = NKZ + 1;             // input planes of a z pair
  constexpr int PZ = (NKZ - 1) / 2;
  constexpr int N = 2 * CP;                // both output planes
  constexpr int SLOTS = 8;
  constexpr uint32_t TMEM_COLS = SLOTS * N;   // 256 or 512
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = p.a;
  uint8_t* ring = smem;                                                   // [stage][q = 2][position][16 B]
  uint8_t* b_s = smem + kFzStages * kFzStageBytes;                        // [kx = 3][q = 2][n][8 halves]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(b_s + 3 * 2 * N * 16);   // [kFzStages]
  uint64_t* step_bar = full_bar + kFzStages;                              // [kFzStepBars]
  uint64_t* tempty_bar = step_bar + kFzStepBars;                          // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 8);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);                 // [32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.in_d[0], H = a.in_d[1], W = a.in_d[2];
  const int n_zp = (D + 1) >> 1;
  // consecutive z pairs per CTA: neighbouring pairs share two input planes (L2 hits a few microseconds apart) and a CTA
  // changes sample - which flushes the InstanceNorm sums - at most a few times
  const int u_begin = (int)blockIdx.x * p.units_per_cta;
  const int u_end = u_begin + p.units_per_cta < p.n_units ? u_begin + p.units_per_cta : p.n_units;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFzStages; ++s) mbar_init(&full_bar[s], 4);      // one arrival per producer warp
    for (int s = 0; s < kFzStepBars; ++s) mbar_init(&step_bar[s], 1);
    for (int s = 0; s < SLOTS; ++s) mbar_init(&tempty_bar[s], 8);        // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights: fp32 [tap = (kz * 3 + ky) * 3 + kx][cout_pad] -> fp16 K-major UMMA tiles, zero where a plane does not feed a tile
  for (int i = threadIdx.x; i < 3 * N * 16; i += kFzThreads) {
    const int k = i & 15, n = (i >> 4) % N, kx = i / (16 * N);
    const int pl = k / 3, ky = k - pl * 3, zt = n / CP, co = n - zt * CP;
    const int kz = pl - zt;
    float v = 0.f;
    if (k < NPL * 3 && kz >= 0 && kz < NKZ && co < a.cout) v = __ldg(a.w + (size_t)((kz * 3 + ky) * 3 + kx) * a.cout_pad + co);
    *reinterpret_cast<__half*>(b_s + ((kx * 2 + (k >> 3)) * N + n) * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  if (threadIdx.x < 32) bias_s[threadIdx.x] = (a.bias && (int)threadIdx.x < a.cout) ? __ldg(a.bias + threadIdx.x) : 0.f;
  // the ring starts as zeros: the producers never write the halo positions (x = -1, x >= W)
  for (int i = threadIdx.x; i < kFzStages * kFzStageBytes / 16; i += kFzThreads)
    reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== PRODUCERS: thread = column x ===========================
    const int x = threadIdx.x;
    const bool col_ok = x < W;
    const uint32_t dst0 = smem_u32(ring) + (uint32_t)(x + 1) * 16u;
    const uint32_t row_stride = (uint32_t)W * (uint32_t)a.src_cs;
    // IDENT: the source carries no pending transform (the gathered image) - the usual case
    auto produce = [&](auto ident_tag) {
      constexpr bool IDENT = decltype(ident_tag)::value;
      int t = 0, stage = 0, cur_b = -1;
      float sc = 1.f, sh = 0.f, sl = 1.f;
      for (int u = u_begin; u < u_end; ++u) {
        const int b = u / n_zp, z0 = 2 * (u - b * n_zp);
        if (!IDENT && b != cur_b) {
          const ChanMeta m = a.src_meta[0];
          xform_from_stats(a.src_stats + (size_t)b * a.src_stat_stride * 2, m, a.src_inv_count, sc, sh);
          sl = m.eps < 0.f ? 1.f : m.slope;
          cur_b = b;
        }
        const __half* vol = a.src + (size_t)b * D * H * W * a.src_cs + (size_t)x * a.src_cs;
        // 32-bit element offsets from the sample base; a plane outside the volume (or a column beyond W) never loads
        uint32_t plane_off[NPL];
        bool plane_ok[NPL];
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          const int zz = z0 - PZ + pl;
          plane_ok[pl] = col_ok && zz >= 0 && zz < D;
          plane_off[pl] = plane_ok[pl] ? (uint32_t)zz * (uint32_t)H * row_stride : 0u;
        }
        auto in_image = [&](int pl, int yy) { return plane_ok[pl] && yy < H; };      // yy >= 0 at every call
        auto load = [&](int pl, int yy) {      // raw value, 0 outside the image
          __half v = __float2half_rn(0.f);
          if (in_image(pl, yy)) v = __ldg(vol + (plane_off[pl] + (uint32_t)yy * row_stride));
          return v;
        };
        auto xf = [&](__half raw, int pl, int yy) {   // the padding stays 0: it pads the TRANSFORMED tensor
          if (IDENT || !in_image(pl, yy)) return raw;
          return __float2half_rn(lrelu(fmaf(__half2float(raw), sc, sh), sl));
        };
        // win[pl][r]: rows y-1, y, y+1 of input plane z0 - PZ + pl at column x; ahead[pl][i]: rows y+1 .. y+kFzAhead,
        // raw, in flight (row r in slot (r - 1) % kFzAhead) - one row of look-ahead left every y-step waiting a full
        // HBM latency (0.82 ms per 32 patches)
        __half win[NPL][3], ahead[NPL][kFzAhead];
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          win[pl][0] = __float2half_rn(0.f);
          win[pl][1] = __float2half_rn(0.f);        // row -1 after the first shift
          win[pl][2] = xf(load(pl, 0), pl, 0);
#pragma unroll
          for (int i = 0; i < kFzAhead; ++i) ahead[pl][i] = load(pl, i + 1);
        }
        for (int yb = 0; yb < H; yb += kFzAhead) {
#pragma unroll
        for (int j = 0; j < kFzAhead; ++j) {
          const int y = yb + j;
          if (y >= H) break;
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            win[pl][0] = win[pl][1];
            win[pl][1] = win[pl][2];
            win[pl][2] = xf(ahead[pl][j], pl, y + 1);
            ahead[pl][j] = load(pl, y + 1 + kFzAhead);
          }
          if (t >= kFzStages) {
            const int tp = t - kFzStages;
            mbar_wait(&step_bar[tp & (kFzStepBars - 1)], (uint32_t)(tp >> 4) & 1u);
          }
          if (col_ok) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              // k = pl * 3 + ky; group q holds k = 8 q .. 8 q + 7
              __half2 h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int k0 = 8 * q + 2 * e, k1 = k0 + 1;
                const __half v0 = k0 < NPL * 3 ? win[k0 / 3 < NPL ? k0 / 3 : 0][k0 % 3] : __float2half_rn(0.f);
                const __half v1 = k1 < NPL * 3 ? win[k1 / 3 < NPL ? k1 / 3 : 0][k1 % 3] : __float2half_rn(0.f);
                h[e] = __halves2half2(v0, v1);
              }
              const uint4 v = *reinterpret_cast<uint4*>(h);
              if (q * 8 < NPL * 3)    // NKZ = 1: k = 0 .. 5, group 1 stays zero
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + (uint32_t)stage * kFzStageBytes + (uint32_t)q * (kFzProw * 16)),
                             "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_warp(&full_bar[stage]);
          if (++stage == kFzStages) stage = 0;
          ++t;
        }
        }
      }
    };
    if (a.src_meta[0].eps < 0.f) produce(std::true_type{}); else produce(std::false_type{});
  } else if (warp == 4) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(ring), kFzProw * 16, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(b_s), N * 16, 128);
    int t = 0, stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    for (int u = u_begin; u < u_end; ++u) {
      for (int y = 0; y < H; ++y, ++t) {
        mbar_wait(&tempty_bar[slot], sphase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t d = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)slot * N, 0);
        const uint32_t da_lo = __shfl_sync(0xffffffffu, (uint32_t)a_desc0 + (uint32_t)stage * (kFzStageBytes >> 4), 0);
        const uint32_t bar = __shfl_sync(0xffffffffu, smem_u32(&step_bar[t & (kFzStepBars - 1)]), 0);
        if (elect_one()) {
          const uint64_t da = (a_desc0 & 0xffffffff00000000ull) | da_lo;
          // kx = 0, 1, 2: the A rows start one position (16 bytes) further each time; B tile kx
          umma_f16_off<0, 0>(d, da, b_desc0, idesc, 0u);
          umma_f16_off<1, 2 * N>(d, da, b_desc0, idesc, 1u);
          umma_f16_off<2, 4 * N>(d, da, b_desc0, idesc, 1u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        if (++stage == kFzStages) { stage = 0; phase ^= 1; }
        if (++slot == SLOTS) { slot = 0; sphase ^= 1; }
      }
    }
  } else {
    // =========================== EPILOGUE (set k = output plane z0 + k) ===========================
    const int k = (warp - 5) >> 2;
    const int wq = warp & 3;
    const int x = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(k * CP);
    const bool col_ok = x < W;
    const bool has_bias = a.bias != nullptr;
    const bool vec_ok = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0);
    const long long partner_delta = (long long)a.dst_cs * 2 * ((lane & 1) ? -1 : 1);     // lanes 2i / 2i+1: columns x, x+1
    const size_t out_row = (size_t)W * a.dst_cs;
    float s1[CP], s2[CP];
#pragma unroll
    for (int j = 0; j < CP; ++j) s1[j] = s2[j] = 0.f;
    int cur_b = -1;
    auto flush_stats = [&](int b) {
      if (!a.dst_stats || b < 0) return;
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        float v1 = s1[j], v2 = s2[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, off);
          v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        if (lane == 0 && j < a.cout) {
          atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + j) * 2 + 0, (double)v1);
          atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + j) * 2 + 1, (double)v2);
        }
        s1[j] = s2[j] = 0.f;
      }
    };
    int t = 0, slot = 0;
    for (int u = u_begin; u < u_end; ++u) {
      const int b = u / n_zp, z = 2 * (u - b * n_zp) + k;
      if (b != cur_b) {
        flush_stats(cur_b);
        cur_b = b;
      }
      const bool plane_ok = z < D;                 // odd depth: the last pair has one plane only
      __half* out_px = a.dst + (((size_t)b * D + (plane_ok ? z : 0)) * H) * out_row + (size_t)x * a.dst_cs;
      for (int y = 0; y < H; ++y, ++t, out_px += out_row) {
        mbar_wait(&step_bar[t & (kFzStepBars - 1)], (uint32_t)(t >> 4) & 1u);
        tc_fence_after();
#pragma unroll
        for (int g0 = 0; g0 < CP; g0 += 16) {
          uint32_t acc[16];
          tmem_ld16(t_lane + (uint32_t)(slot * N + g0), acc);
          if (g0 + 16 >= CP) {
            tc_fence_before();
            mbar_arrive_warp(&tempty_bar[slot]);
          }
          if (has_bias) {      // a branch, not 32 predicated-off loads and adds per row
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + bias_s[g0 + j]);
          }
          const bool ok = col_ok && plane_ok;
          __half2 hv[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) hv[j >> 1] = __floats2half2_rn(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]));
          if (ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              // sums of the ROUNDED (stored) values, as the general and CUDA-core kernels keep them
              const float2 r = __half22float2(hv[j >> 1]);
              s1[g0 + j] += r.x;
              s2[g0 + j] = fmaf(r.x, r.x, s2[g0 + j]);
              s1[g0 + j + 1] += r.y;
              s2[g0 + j + 1] = fmaf(r.y, r.y, s2[g0 + j + 1]);
            }
          }
          __half* q = out_px + g0;
          if (vec_ok && g0 + 16 <= a.cout) {
            // whole 32-byte sectors per instruction (see stg32_paired)
            stg32_paired(q, partner_delta, *reinterpret_cast<uint4*>(&hv[0]), *reinterpret_cast<uint4*>(&hv[4]), ok, lane);
          } else if (ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (g0 + j < a.cout) q[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
          }
        }
        if (++slot == SLOTS) slot = 0;
      }
    }
    flush_stats(cur_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace

bool first_zpair_supported(const ConvArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FNNU_FIRST_ZPAIR");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled) return false;
  if (a.transposed || a.cin != 1 || (a.cout_pad != 16 && a.cout_pad != 32)) return false;
  if (a.k[1] != 3 || a.k[2] != 3 || (a.k[0] != 1 && a.k[0] != 3)) return false;
  for (int i = 0; i < 3; ++i)
    if (a.s[i] != 1 || a.pad[i] != (a.k[i] - 1) / 2 || a.in_d[i] != a.out_d[i]) return false;
  return a.in_d[2] <= 128 && a.w != nullptr;
}

// host check of "no pending transform on the source": the meta lives on the device, so the engine passes the fact
int launch_conv_first_zpair(const ConvArgs& a, cudaStream_t s) {
  if (!first_zpair_supported(a)) {
    set_error("conv_first_zpair: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  FzArgs p;
  p.a = a;
  p.n_units = a.batch * ((a.in_d[0] + 1) / 2);
  const int N = 2 * a.cout_pad;
  // padded so that no more CTAs are resident per SM than the 512 TMEM columns allow (2 x 256 or 1 x 512)
  size_t smem = (size_t)kFzStages * kFzStageBytes + (size_t)3 * 2 * N * 16 + (kFzStages + kFzStepBars + 8) * 8 + 8 + 32 * 4 + 64;
  const size_t smem_min = a.cout_pad == 16 ? 78 * 1024 : 116 * 1024;
  if (smem < smem_min) smem = smem_min;
  const int ctas = (a.cout_pad == 16 ? 2 : 1) * num_sms();
  p.units_per_cta = (p.n_units + ctas - 1) / ctas;
  const int grid = (p.n_units + p.units_per_cta - 1) / p.units_per_cta;
#define FNNU_FZ_CASE(CPV, NKZV)                                                                                               \
  if (a.cout_pad == CPV && a.k[0] == NKZV) {                                                                                  \
    FNNU_CUDA(cudaFuncSetAttribute(conv_first_zpair_kernel<CPV, NKZV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024)); \
    conv_first_zpair_kernel<CPV, NKZV><<<grid, kFzThreads, smem, s>>>(p);                                                     \
  }
  FNNU_FZ_CASE(16, 3) else FNNU_FZ_CASE(16, 1) else FNNU_FZ_CASE(32, 3) else FNNU_FZ_CASE(32, 1)
#undef FNNU_FZ_CASE
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
