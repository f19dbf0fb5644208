// Stride-2 3x3x3 convolution 16 -> (<= 32) channels (the first down-sampling conv of the student / ResEnc encoders) on
// the tensor cores, row-streaming.  The generic implicit-GEMM kernel ran it at 1.95 ms per 32 patches (119 TFLOP/s):
// it gathers a 27-tap im2col tile per 128 outputs, 8 KB of shared-memory stores per tile for 4 KB of input.
//
// Stride 2 along x breaks the "kx = descriptor shift" trick of conv_umma_zrows.cu (output x reads input 2x - 1 + kx),
// but in NDHWC memory the input row IS ALREADY a row of 64 position PAIRS with 32 channels each: pair p holds input
// positions 2p (channels 0-15 of the pair = "sub 0") and 2p + 1 ("sub 1").  In that view the layer is a stride-1 conv
// along x with two taps,
//   pair offset -1: sub 1 only (kx = 0)           pair offset 0: sub 0 (kx = 1) and sub 1 (kx = 2)
// each of them a K = 16 MMA whose A rows start at stored position 0 or 1 of the same shared-memory row - no im2col.
// Along y and z the stride only selects rows: output row (zo, yo) reads input rows 2 yo - 1 + ky of planes
// 2 zo - 1 + kz, so a unit (sample, zo) streams along y: step j stages input rows 2j and 2j + 1 of the three planes
// (6 rows, each loaded once per unit) and its 27 MMAs read this stage plus the odd rows of the previous one.  Padding
// exists only on the low side (D, H, W even): row -1 / plane -1 MMAs are skipped, the pair -1 halo is a zero written
// once.  M = 128 with 64 output positions per row: TMEM lanes 64-127 compute on whatever follows in shared memory and
// are never read.
//   warps 2,3,6,7,9-12 producers (16-byte loads, normalise + LeakyReLU on load, conflict-free st.shared)
//   warp 8            MMA issue, one commit per step
//   warps 0,1,4,5     epilogue (TMEM lanes 0-63 belong to warps with id % 4 in {0, 1}); 16 channels each
// Two CTAs per SM, 256 TMEM columns each (8 accumulator slots of 32).
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

namespace {

constexpr int kS2Threads = 13 * 32;
constexpr int kS2Pitch = 66;                              // stored positions per 8-channel group: 16 * 66 = 8 * 128 + 32 bytes
                                                          // (a quarter warp's 2 pairs x 4 groups hit 8 different 16-byte bank slots)
constexpr int kS2RowBytes = 4 * kS2Pitch * 16;            // 4 groups: (sub, half)
constexpr int kS2StageBytes = 6 * kS2RowBytes;            // 3 planes x (even row, odd row)
constexpr int kS2Stages = 3;
#ifndef FNNU_S2_PREFETCH
#define FNNU_S2_PREFETCH 4
#endif
constexpr int kS2Prefetch = FNNU_S2_PREFETCH;             // y-steps of L2 look-ahead
constexpr int kS2StepBars = 16;
constexpr int kS2Slots = 8;
constexpr int kS2CP = 32;
constexpr int kS2WBytes = 27 * kS2CP * 32;
constexpr int kS2Tail = 1024;                             // the last group's M = 128 read runs 1008 bytes past the ring

struct S2Args {
  ConvArgs a;
  int n_units;          // batch * Do
};

__device__ __forceinline__ uint4 s2_ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void s2_sts16(uint32_t addr, const uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 s2_xform8(const uint4 raw, const __half2* m2, const __half2* s2, const __half2* t2, const __half2* l2) {
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

__global__ void __launch_bounds__(kS2Threads, 2) conv_s2_umma_kernel(const __grid_constant__ S2Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = p.a;
  uint8_t* w_s = smem;                                                     // [tap 27][half 2][n 32][8 halves]
  uint8_t* ring = smem + kS2WBytes;                                        // [stage][plane 3][parity 2][group 4][pitch][16 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kS2Stages * kS2StageBytes + kS2Tail);   // [kS2Stages]
  uint64_t* step_bar = full_bar + 4;                                       // [kS2StepBars]
  uint64_t* tempty_bar = step_bar + kS2StepBars;                           // [kS2Slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kS2Slots);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);                 // [32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.in_d[0], H = a.in_d[1], W = a.in_d[2];
  const int Do = a.out_d[0], Ho = a.out_d[1], Wo = a.out_d[2];
  // unit u -> CTA u % grid: neighbouring CTAs work on neighbouring output planes AT THE SAME TIME, so the input plane two
  // output planes share is fetched from HBM once and hits L2 for the other (consecutive planes per CTA re-read it 64
  // steps later, long after L2 had dropped it: 1.5x the DRAM reads)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kS2Stages; ++s) mbar_init(&full_bar[s], 8);        // one arrival per producer warp
    for (int s = 0; s < kS2StepBars; ++s) mbar_init(&step_bar[s], 1);
    for (int s = 0; s < kS2Slots; ++s) mbar_init(&tempty_bar[s], 4);       // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights: fp32 [tap][ci 16][cout_pad 32] -> fp16 K-major tiles [tap][ci / 8][n][ci % 8]
  for (int i = threadIdx.x; i < 27 * 16 * kS2CP; i += kS2Threads) {
    const int co = i % kS2CP, ci = (i / kS2CP) % 16, tap = i / (kS2CP * 16);
    const float v = co < a.cout ? __ldg(a.w + i) : 0.f;
    *reinterpret_cast<__half*>(w_s + ((size_t)(tap * 2 + (ci >> 3)) * kS2CP + co) * 16 + (ci & 7) * 2) = __float2half_rn(v);
  }
  if (threadIdx.x < 32) bias_s[threadIdx.x] = (a.bias && (int)threadIdx.x < a.cout) ? __ldg(a.bias + threadIdx.x) : 0.f;
  // zeros: the pair -1 halo of every row (never written again) and the tail the last row's MMAs read
  for (int i = threadIdx.x; i < (kS2Stages * kS2StageBytes + kS2Tail) / 16; i += kS2Threads)
    reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool is_epilogue = warp < 8 && (warp & 3) < 2;
  if (warp == 8) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kS2CP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(ring), kS2Pitch * 16, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(w_s), kS2CP * 16, 128);
    int t = 0, stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x) {
      const int zo = u % Do;
      for (int j = 0; j < Ho; ++j, ++t) {
        mbar_wait(&tempty_bar[slot], sphase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(slot * kS2CP);
          const int prev = stage == 0 ? kS2Stages - 1 : stage - 1;
          uint32_t accum = 0;
          for (int kz = zo == 0 ? 1 : 0; kz < 3; ++kz) {
            for (int ky = j == 0 ? 1 : 0; ky < 3; ++ky) {
              // input row 2j - 1 + ky: the previous stage's odd row, this stage's even row, this stage's odd row
              const int st = ky == 0 ? prev : stage;
              const int parity = ky == 1 ? 0 : 1;
              const uint32_t row16 = (uint32_t)(st * kS2StageBytes + (kz * 2 + parity) * kS2RowBytes) >> 4;
              const uint64_t da = a_desc0 + (uint64_t)row16;
              const uint64_t db = b_desc0 + (uint64_t)(uint32_t)(((kz * 3 + ky) * 3) * (kS2CP * 2));
              // kx = 0: sub 1 (groups 2, 3) from stored position 0; kx = 1: sub 0 from position 1; kx = 2: sub 1 from position 1
              umma_f16(d, da + (uint64_t)(2 * kS2Pitch), db, idesc, accum);
              umma_f16(d, da + 1, db + (uint64_t)(kS2CP * 2), idesc, 1u);
              umma_f16(d, da + (uint64_t)(2 * kS2Pitch + 1), db + (uint64_t)(2 * kS2CP * 2), idesc, 1u);
              accum = 1;
            }
          }
          umma_commit(&step_bar[t & (kS2StepBars - 1)]);
        }
        __syncwarp();
        if (++stage == kS2Stages) { stage = 0; phase ^= 1; }
        if (++slot == kS2Slots) { slot = 0; sphase ^= 1; }
      }
    }
  } else if (!is_epilogue) {
    // =========================== PRODUCERS (warps 2, 3, 6, 7, 9 .. 12) ===========================
    const int pw = warp < 4 ? warp - 2 : (warp < 8 ? warp - 4 : warp - 5);
    const int item = pw * 32 + lane;                  // (pair, group): 64 pairs x 4 groups of one input row
    const int pair = item >> 2, grp = item & 3;       // grp = sub * 2 + half
    const bool item_ok = pair * 2 < W;
    const bool pf_lane = item_ok && (grp & 1) == 0;       // one request per 32-byte sector
    const uint32_t goff = (uint32_t)((pair * 2 + (grp >> 1)) * a.src_cs + (grp & 1) * 8) * 2u;
    const uint32_t soff = smem_u32(ring) + (uint32_t)(grp * kS2Pitch + pair + 1) * 16u;
    const uint32_t row_bytes = (uint32_t)(W * a.src_cs) * 2u;
    const long long plane_bytes = (long long)H * row_bytes;
    __half2 m2[4], s2[4], t2[4], l2[4];
    int t = 0, stage = 0, cur_b = -1;
    for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x) {
      const int b = u / Do, zo = u - b * Do;
      if (b != cur_b) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float mh[2], sc[2], sh[2], sl[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = (grp & 1) * 8 + 2 * e + k;
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
      // planes 2 zo - 1 .. 2 zo + 1; plane -1 (zo = 0) is padding: not loaded, its MMAs are skipped
      const char* vol = reinterpret_cast<const char*>(a.src) + ((long long)b * D + (2 * zo - 1)) * plane_bytes + goff;
      const int kz0 = zo == 0 ? 1 : 0;
      // next unit's planes, for the look-ahead across the unit boundary
      const int un = u + (int)gridDim.x;
      const int bn = un / Do, zn = un - bn * Do;
      const char* vol_next = reinterpret_cast<const char*>(a.src) + ((long long)bn * D + (2 * zn - 1)) * plane_bytes + goff;
      for (int j = 0; j < Ho; ++j, ++t) {
        // L2 look-ahead: registers hold one step of loads per thread (24 KB per CTA in flight), far too little for
        // HBM latency; the sectors of step j + kS2Prefetch are requested now, so that their loads hit L2 later
        if (pf_lane) {
          const int jp = j + kS2Prefetch;
          const bool same = jp < Ho;
          if (same || un < p.n_units) {
            const char* pr = (same ? vol : vol_next) + (long long)(2 * (same ? jp : jp - Ho)) * row_bytes;
            const int kzp = (same ? zo : zn) == 0 ? 1 : 0;
#pragma unroll
            for (int kz = 0; kz < 3; ++kz)
#pragma unroll
              for (int par = 0; par < 2; ++par)
                if (kz >= kzp) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + kz * plane_bytes + par * row_bytes));
          }
        }
        uint4 v[6];
        const char* rows = vol + (long long)(2 * j) * row_bytes;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
          for (int par = 0; par < 2; ++par)
            if (kz >= kz0 && item_ok) v[kz * 2 + par] = s2_ldg16(rows + kz * plane_bytes + par * row_bytes);
        // the stage was last read by steps t - S and t - S + 1
        if (t >= kS2Stages) {
          const int tp = t - kS2Stages + 1;
          mbar_wait(&step_bar[tp & (kS2StepBars - 1)], (uint32_t)(tp >> 4) & 1u);
        }
        if (item_ok) {
#pragma unroll
          for (int kz = 0; kz < 3; ++kz)
#pragma unroll
            for (int par = 0; par < 2; ++par)
              if (kz >= kz0)
                s2_sts16(soff + (uint32_t)(stage * kS2StageBytes + (kz * 2 + par) * kS2RowBytes), s2_xform8(v[kz * 2 + par], m2, s2, t2, l2));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_warp(&full_bar[stage]);
        if (++stage == kS2Stages) stage = 0;
      }
    }
  } else {
    // =========================== EPILOGUE (warps 0, 1, 4, 5) ===========================
    const int wq = warp & 3;                 // 0 or 1: TMEM lanes 0-31 / 32-63 = output columns
    const int c0 = (warp >> 2) * 16;         // first output channel of this warp
    const int x = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
    const bool col_ok = x < Wo;
    const bool has_bias = a.bias != nullptr;
    const bool vec_ok = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0) && c0 + 16 <= a.cout;
    const long long partner_delta = (long long)a.dst_cs * 2 * ((lane & 1) ? -1 : 1);
    const size_t out_row = (size_t)Wo * a.dst_cs;
    float s1[16], s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s1[j] = s2[j] = 0.f;
    int cur_b = -1;
    auto flush_stats = [&](int b) {
      if (!a.dst_stats || b < 0) return;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v1 = s1[j], v2 = s2[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, off);
          v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        if (lane == 0 && c0 + j < a.cout) {
          atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + c0 + j) * 2 + 0, (double)v1);
          atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + c0 + j) * 2 + 1, (double)v2);
        }
        s1[j] = s2[j] = 0.f;
      }
    };
    int t = 0, slot = 0;
    for (int u = (int)blockIdx.x; u < p.n_units; u += (int)gridDim.x) {
      const int b = u / Do, zo = u - b * Do;
      if (b != cur_b) {
        flush_stats(cur_b);
        cur_b = b;
      }
      __half* out_px = a.dst + (((size_t)b * Do + zo) * Ho) * out_row + (size_t)x * a.dst_cs + c0;
      for (int j = 0; j < Ho; ++j, ++t, out_px += out_row) {
        mbar_wait(&step_bar[t & (kS2StepBars - 1)], (uint32_t)(t >> 4) & 1u);
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(t_lane + (uint32_t)(slot * kS2CP), acc);
        tc_fence_before();
        mbar_arrive_warp(&tempty_bar[slot]);
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + bias_s[c0 + i]);
        }
        __half2 hv[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) hv[i >> 1] = __floats2half2_rn(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]));
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float2 r = __half22float2(hv[i >> 1]);      // sums of the stored (rounded) values
            s1[i] += r.x;
            s2[i] = fmaf(r.x, r.x, s2[i]);
            s1[i + 1] += r.y;
            s2[i + 1] = fmaf(r.y, r.y, s2[i + 1]);
          }
        }
        if (vec_ok) {
          stg32_paired(out_px, partner_delta, *reinterpret_cast<uint4*>(&hv[0]), *reinterpret_cast<uint4*>(&hv[4]), col_ok, lane);
        } else if (col_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < a.cout) out_px[i] = (i & 1) ? __high2half(hv[i >> 1]) : __low2half(hv[i >> 1]);
        }
        if (++slot == kS2Slots) slot = 0;
      }
    }
    flush_stats(cur_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

}  // namespace

bool s2_umma_supported(const ConvArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FNNU_S2_UMMA");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || a.transposed || a.w == nullptr) return false;
  if (a.cin != 16 || a.cout_pad != kS2CP) return false;
  for (int i = 0; i < 3; ++i)
    if (a.k[i] != 3 || a.s[i] != 2 || a.pad[i] != 1 || (a.in_d[i] & 1) || a.out_d[i] * 2 != a.in_d[i]) return false;
  if (a.in_d[2] > 128 || a.in_d[2] < 32) return false;
  return a.src_cs % 8 == 0 && ((uintptr_t)a.src % 16) == 0;
}

int launch_conv_s2_umma(const ConvArgs& a, cudaStream_t s) {
  if (!s2_umma_supported(a)) {
    set_error("conv_s2_umma: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  S2Args p;
  p.a = a;
  p.n_units = a.batch * a.out_d[0];
  const int ctas = 2 * num_sms();
  const int grid = p.n_units < ctas ? p.n_units : ctas;
  // > 227 KB / 3: never three CTAs (3 x 256 TMEM columns) on an SM
  const size_t smem = (size_t)kS2WBytes + kS2Stages * kS2StageBytes + kS2Tail + (4 + kS2StepBars + kS2Slots) * 8 + 16 + 32 * 4 + 64;
  static_assert(kS2WBytes + kS2Stages * kS2StageBytes + kS2Tail + 512 > 78 * 1024, "two CTAs per SM at most");
  static_assert(kS2WBytes + kS2Stages * kS2StageBytes + kS2Tail + 512 < 112 * 1024, "two CTAs per SM must fit");
  FNNU_CUDA(cudaFuncSetAttribute(conv_s2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
  conv_s2_umma_kernel<<<grid, kS2Threads, smem, s>>>(p);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
