// Conv3d as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// GEMM view (per output plane z, per block of TY output rows):
//   D[position, cout] += A[position + tap shift, cin-chunk] * W[tap][cout, cin-chunk]
// over the taps (kz, ky, kx) and 16-channel chunks of Cin.  M = 128 consecutive positions of a
// zero-padded, flattened (row, x) block of the input plane held in shared memory, N = Cout chunk
// (<= 256), K = 16 channels per tcgen05.mma.  Because the block is flattened with its padding, every
// (ky, kx) tap of every M-tile is the SAME shared-memory block read from a start address shifted by
// (ky * Wp + kx) positions: one copy of the input serves all 9 in-plane taps and all M-tiles.
//
// Shared-memory operand layout (no swizzle, K-major canonical layout of the UMMA descriptor):
//   A stage : [8-channel group q][position p][8 halves]   -> row pitch 16 B, SBO = 128 B, LBO = plane
//   B stage : [chunk][tap][8-channel group][cout n][8 halves]
// A is written by 4 producer warps that read the RAW fp16 output of the previous layer from HBM/L2 and
// apply its InstanceNorm affine + LeakyReLU on the way (the fused "normalise on load"); B (weights,
// pre-packed per stage) arrives by one cp.async.bulk (TMA engine) per stage.  One thread issues the
// MMAs; 4 epilogue warps drain TMEM (tcgen05.ld), add the bias, round to fp16, store channels-last and
// accumulate the InstanceNorm sums of the rounded values (fp32 partials -> fp64 atomics).
//
// Warp roles (288 threads): warps 0-3 producers, warps 4-7 epilogue (TMEM lane quarter = warp % 4),
// warp 8 MMA issuer + TMEM allocator.  Pipelines: smem ring (full/empty mbarriers) between producers
// and MMA; TMEM accumulator buffers (full/empty mbarriers) between MMA and epilogue.
//
// Replaces the cuDNN calls behind `self.network(x)` (predict_from_raw_data.py:543) for stride-1
// Conv3d layers with Cin % 16 == 0; other shapes run on conv_ref.cu.
#include "common.cuh"
#include "ops.cuh"

namespace fnnu {

struct UmmaCfg {
  int ok;
  int Nc, n_chunks;       // N per MMA, number of Cout chunks
  int KC, G;              // 16-channel chunks per stage, stage groups per kz
  int TY, n_yblocks, T;   // output rows per unit, units per plane, M-tiles per unit
  int rows_mode, tiles_per_row;
  int Wp, R, P_fill, P_alloc;
  int pz, py, px, nkz, nky, nkx;
  int stages, a_stage_bytes, b_stage_bytes;
  int tmem_bufs;
  int smem_bytes;
  unsigned wp_magic;
};

struct UmmaArgs {
  ConvArgs a;
  UmmaCfg c;
  int n_units;
};

constexpr int kProducerThreads = 128;
constexpr int kEpilogueThreads = 128;
constexpr int kThreadsUmma = 288;
constexpr int kSmemLimit = 227 * 1024;

__host__ __device__ inline int tile_base_of(const UmmaCfg& c, int i) {
  return c.rows_mode ? (i / c.tiles_per_row) * c.Wp + (i % c.tiles_per_row) * 128 : i * 128;
}

static bool plan_umma(const ConvArgs& a, UmmaCfg& c) {
  memset(&c, 0, sizeof(c));
  if (a.transposed) return false;
  if (a.s[0] != 1 || a.s[1] != 1 || a.s[2] != 1) return false;
  if (a.cin % 16 != 0 || a.cin < 16) return false;
  if (a.src_cs % 8 != 0 || ((uintptr_t)a.src % 16) != 0) return false;
  const int D = a.in_d[0], H = a.in_d[1], W = a.in_d[2];
  (void)D;
  c.nkz = a.k[0]; c.nky = a.k[1]; c.nkx = a.k[2];
  c.pz = a.pad[0]; c.py = a.pad[1]; c.px = a.pad[2];
  const int cout_pad = a.cout_pad;
  if (cout_pad <= 256) {
    c.Nc = cout_pad;
  } else {
    c.Nc = 0;
    for (int n = 256; n >= 16; n -= 16)
      if (cout_pad % n == 0) { c.Nc = n; break; }
    if (!c.Nc) return false;
  }
  c.n_chunks = cout_pad / c.Nc;
  c.Wp = W + 2 * c.px;
  if (c.Wp >= 65536) return false;
  c.rows_mode = (W % 128 == 0);
  c.tiles_per_row = c.rows_mode ? W / 128 : 0;
  const int ntyx = c.nky * c.nkx;
  const int misc = 256 + 3 * a.cin * 4 + 2 * c.Nc * 4 + 1024;
  const int avail = kSmemLimit - misc;
  double best_cost = 1e30;
  int best_ty = 0;
  for (int ty = 1; ty <= H; ++ty) {
    int T = c.rows_mode ? ty * c.tiles_per_row : ((ty - 1) * c.Wp + W + 127) / 128;
    if (T * c.Nc > 512) break;
    int R = ty + 2 * c.py;
    int last_base = c.rows_mode ? ((T - 1) / c.tiles_per_row) * c.Wp + ((T - 1) % c.tiles_per_row) * 128 : (T - 1) * 128;
    int need = last_base + 128 + (c.nky - 1) * c.Wp + (c.nkx - 1);
    if (need < R * c.Wp) need = R * c.Wp;
    int palloc = (need + 7) / 8 * 8 + 4;
    int stage1 = 2 * palloc * 16 + ntyx * 2 * c.Nc * 16;
    if (2 * stage1 > avail) break;
    int nb = (H + ty - 1) / ty;
    double cost = (double)nb * T * (2 * T * c.Nc <= 512 ? 1.0 : 1.06) * (1.0 + 0.15 * (double)(R - ty) / ty);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_ty = ty; }
  }
  if (!best_ty) return false;
  c.TY = best_ty;
  c.n_yblocks = (H + c.TY - 1) / c.TY;
  c.T = c.rows_mode ? c.TY * c.tiles_per_row : ((c.TY - 1) * c.Wp + W + 127) / 128;
  c.R = c.TY + 2 * c.py;
  c.P_fill = c.R * c.Wp;
  {
    int last_base = tile_base_of(c, c.T - 1);
    int need = last_base + 128 + (c.nky - 1) * c.Wp + (c.nkx - 1);
    if (need < c.P_fill) need = c.P_fill;
    c.P_alloc = (need + 7) / 8 * 8 + 4;
  }
  c.tmem_bufs = (2 * c.T * c.Nc <= 512) ? 2 : 1;
  const int chunks = a.cin / 16;
  c.KC = 0;
  for (int kc = 4; kc >= 1; kc >>= 1) {
    if (chunks % kc) continue;
    int stage = kc * (2 * c.P_alloc * 16 + ntyx * 2 * c.Nc * 16);
    int st = avail / stage;
    if (st >= 3 || (kc == 1 && st >= 2)) {
      c.KC = kc;
      c.stages = st > 4 ? 4 : st;
      break;
    }
  }
  if (!c.KC) return false;
  c.G = chunks / c.KC;
  c.a_stage_bytes = c.KC * 2 * c.P_alloc * 16;
  c.b_stage_bytes = c.KC * ntyx * 2 * c.Nc * 16;
  c.smem_bytes = c.stages * (c.a_stage_bytes + c.b_stage_bytes) + misc;
  c.wp_magic = (unsigned)((0x100000000ull + (unsigned)c.Wp - 1) / (unsigned)c.Wp);
  if ((c.P_alloc * 16 >> 4) > 0x3FFF || (c.Nc * 16 >> 4) > 0x3FFF) return false;
  c.ok = 1;
  return true;
}

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100); SWIZZLE_NONE, base offset 0
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct UnitIdx {
  int b, z, yb, nc;
};
__device__ __forceinline__ UnitIdx decode_unit(const UmmaArgs& p, int u) {
  UnitIdx r;
  const int D = p.a.out_d[0];
  r.z = u % D;
  u /= D;
  r.yb = u % p.c.n_yblocks;
  u /= p.c.n_yblocks;
  r.nc = u % p.c.n_chunks;
  r.b = u / p.c.n_chunks;
  return r;
}

__global__ void __launch_bounds__(kThreadsUmma, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const UmmaCfg& c = p.c;
  const ConvArgs& a = p.a;
  const int stage_bytes = c.a_stage_bytes + c.b_stage_bytes;
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)c.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tfull_bar = empty_bar + 4;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* xs = reinterpret_cast<float*>(tmem_slot + 4);
  float* xh = xs + a.cin;
  float* xl = xh + a.cin;
  float* stat_s = xl + a.cin;   // [2][Nc]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.in_d[0], H = a.in_d[1], W = a.in_d[2];
  const int ntyx = c.nky * c.nkx;

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full_bar[s], kProducerThreads);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpilogueThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * c.Nc; i += blockDim.x) stat_s[i] = 0.f;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_cols = (uint32_t)(c.T * c.Nc);

  if (warp < 4) {
    // =========================== PRODUCERS ===========================
    const int tid = threadIdx.x;
    const int Q = 2 * c.KC;               // 8-channel groups per stage (2, 4 or 8)
    const int q = tid % Q;
    const int items = c.P_fill * Q;
    int stage = 0;
    uint32_t phase = 0;
    int cur_b = -1;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitIdx ui = decode_unit(p, u);
      if (ui.b != cur_b) {
        named_bar_sync(1, kProducerThreads);
        for (int ch = tid; ch < a.cin; ch += kProducerThreads) {
          float sc, sh;
          const ChanMeta m = a.src_meta[ch];
          xform_from_stats(a.src_stats + ((size_t)ui.b * a.src_stat_stride + ch) * 2, m, a.src_inv_count, sc, sh);
          xs[ch] = sc;
          xh[ch] = sh;
          xl[ch] = m.eps < 0.f ? 1.f : m.slope;
        }
        named_bar_sync(1, kProducerThreads);
        cur_b = ui.b;
      }
      const int y0 = ui.yb * c.TY;
      for (int kz = 0; kz < c.nkz; ++kz) {
        const int z_in = ui.z + kz - c.pz;
        if (z_in < 0 || z_in >= D) continue;
        const __half* plane = a.src + ((size_t)ui.b * D + z_in) * H * W * a.src_cs;
        for (int g = 0; g < c.G; ++g) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_s = ring + (size_t)stage * stage_bytes;
          uint8_t* b_s = a_s + c.a_stage_bytes;
          if (tid == 0) {
            mbar_expect_tx(&full_bar[stage], (uint32_t)c.b_stage_bytes);
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w_umma) +
                                  ((size_t)(ui.nc * c.nkz + kz) * c.G + g) * c.b_stage_bytes;
            bulk_g2s(b_s, wsrc, (uint32_t)c.b_stage_bytes, &full_bar[stage]);
          }
          const int ch0 = g * c.KC * 16 + q * 8;
          float sc[8], sh[8], sl[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sc[e] = xs[ch0 + e];
            sh[e] = xh[ch0 + e];
            sl[e] = xl[ch0 + e];
          }
          uint8_t* a_q = a_s + (size_t)q * c.P_alloc * 16;
          constexpr int U = 4;
          for (int j0 = tid; j0 < items; j0 += kProducerThreads * U) {
            uint4 raw[U];
            int pos[U];
            bool ok[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
              const int j = j0 + k * kProducerThreads;
              const int pp = j / Q;
              pos[k] = (j < items) ? pp : -1;
              const int r = (int)__umulhi((unsigned)pp, c.wp_magic);
              const int xp = pp - r * c.Wp;
              const int y_in = y0 - c.py + r;
              const int x_in = xp - c.px;
              ok[k] = (j < items) && y_in >= 0 && y_in < H && x_in >= 0 && x_in < W;
              raw[k] = make_uint4(0u, 0u, 0u, 0u);
              if (ok[k]) raw[k] = __ldg(reinterpret_cast<const uint4*>(plane + ((size_t)y_in * W + x_in) * a.src_cs + ch0));
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
              if (pos[k] < 0) continue;
              uint4 o = make_uint4(0u, 0u, 0u, 0u);
              if (ok[k]) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw[k]);
                __half2 r2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __half22float2(h2[e]);
                  float v0 = fmaf(f.x, sc[2 * e], sh[2 * e]);
                  float v1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
                  v0 = fmaxf(v0, v0 * sl[2 * e]);
                  v1 = fmaxf(v1, v1 * sl[2 * e + 1]);
                  r2[e] = __floats2half2_rn(v0, v1);
                }
                o = *reinterpret_cast<uint4*>(r2);
              }
              *reinterpret_cast<uint4*>(a_q + (size_t)pos[k] * 16) = o;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&full_bar[stage]);
          if (++stage == c.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    // =========================== MMA ISSUER ===========================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(c.Nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)c.P_alloc * 16, b_lbo = (uint32_t)c.Nc * 16;
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase[2] = {0, 0};
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const UnitIdx ui = decode_unit(p, u);
        mbar_wait(&tempty_bar[buf], tphase[buf] ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)buf * buf_cols;
        bool first = true;
        for (int kz = 0; kz < c.nkz; ++kz) {
          const int z_in = ui.z + kz - c.pz;
          if (z_in < 0 || z_in >= D) continue;
          for (int g = 0; g < c.G; ++g) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_base = smem_u32(ring + (size_t)stage * stage_bytes);
            const uint32_t b_base = a_base + (uint32_t)c.a_stage_bytes;
            for (int kc = 0; kc < c.KC; ++kc) {
              for (int t = 0; t < ntyx; ++t) {
                const int ky = t / c.nkx, kx = t - ky * c.nkx;
                const uint64_t db = make_desc(b_base + (uint32_t)((kc * ntyx + t) * 2) * b_lbo, b_lbo, 128);
                const uint32_t a_off = a_base + (uint32_t)(kc * 2) * a_lbo + (uint32_t)(ky * c.Wp + kx) * 16;
                const uint32_t accum = (first && kc == 0 && t == 0) ? 0u : 1u;
                for (int i = 0; i < c.T; ++i) {
                  const uint64_t da = make_desc(a_off + (uint32_t)tile_base_of(c, i) * 16, a_lbo, 128);
                  umma_f16(d_base + (uint32_t)(i * c.Nc), da, db, idesc, accum);
                }
              }
            }
            umma_commit(&empty_bar[stage]);
            first = false;
            if (++stage == c.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tfull_bar[buf]);
        tphase[buf] ^= 1;
        if (c.tmem_bufs == 2) buf ^= 1;
      }
    }
    __syncwarp();
  } else {
    // =========================== EPILOGUE ===========================
    const int wq = warp & 3;                     // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 4 * 32;         // 0..127
    const int c0 = c.py * c.Wp + c.px;
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    const bool vec_store = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0);
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitIdx ui = decode_unit(p, u);
      const int y0 = ui.yb * c.TY;
      mbar_wait(&tfull_bar[buf], tphase[buf]);
      tphase[buf] ^= 1;
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(wq * 32) << 16);
      __half* out_plane = a.dst + ((size_t)ui.b * D + ui.z) * H * W * a.dst_cs;
      for (int n0 = 0; n0 < c.Nc; n0 += 16) {
        const int co0 = ui.nc * c.Nc + n0;
        float bias[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) bias[j] = (a.bias && co0 + j < a.cout) ? __ldg(a.bias + co0 + j) : 0.f;
        float s1[16], s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s1[j] = s2[j] = 0.f;
        const bool full16 = co0 + 16 <= a.cout;
        for (int i = 0; i < c.T; ++i) {
          const int pc = tile_base_of(c, i) + c0 + wq * 32 + lane;
          const int r = (int)__umulhi((unsigned)pc, c.wp_magic);
          const int xp = pc - r * c.Wp;
          const int yo = r - c.py, xo = xp - c.px;
          const int y = y0 + yo;
          const bool valid = xo >= 0 && xo < W && yo < c.TY && y < H;
          uint32_t acc[16];
          tmem_ld16(d_base + (uint32_t)(i * c.Nc + n0), acc);
          if (valid) {
            __half hv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              hv[j] = __float2half_rn(__uint_as_float(acc[j]) + bias[j]);
              const float f = __half2float(hv[j]);
              s1[j] += f;
              s2[j] = fmaf(f, f, s2[j]);
            }
            __half* q = out_plane + ((size_t)y * W + xo) * a.dst_cs + co0;
            if (vec_store && full16) {
              reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
              reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[8]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < a.cout) q[j] = hv[j];
            }
          }
        }
        if (a.dst_stats) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
              s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
            }
          }
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              atomicAdd(&stat_s[n0 + j], s1[j]);
              atomicAdd(&stat_s[c.Nc + n0 + j], s2[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      if (c.tmem_bufs == 2) buf ^= 1;
      if (a.dst_stats) {
        named_bar_sync(2, kEpilogueThreads);
        for (int i = et; i < 2 * c.Nc; i += kEpilogueThreads) {
          const int which = i / c.Nc, n = i - which * c.Nc;
          const int co = ui.nc * c.Nc + n;
          if (co < a.cout) atomicAdd(a.dst_stats + ((size_t)ui.b * a.dst_stat_stride + co) * 2 + which, (double)stat_s[i]);
          stat_s[i] = 0.f;
        }
        named_bar_sync(2, kEpilogueThreads);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: [cout][cin][kz][ky][kx] fp32 -> per-stage blobs of fp16
//   [n-chunk][kz][group g][chunk kc][tap (ky,kx)][8-channel half][n < Nc][8 halves]
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_umma_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout,
                                         UmmaCfg c) {
  const int ntyx = c.nky * c.nkx;
  const size_t total = (size_t)c.n_chunks * c.nkz * c.G * c.KC * ntyx * 2 * c.Nc * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % c.Nc); r /= c.Nc;
    const int h = (int)(r % 2); r /= 2;
    const int t = (int)(r % ntyx); r /= ntyx;
    const int kc = (int)(r % c.KC); r /= c.KC;
    const int g = (int)(r % c.G); r /= c.G;
    const int kz = (int)(r % c.nkz); r /= c.nkz;
    const int nc = (int)r;
    const int co = nc * c.Nc + n;
    const int ci = (g * c.KC + kc) * 16 + h * 8 + e;
    const int ky = t / c.nkx, kx = t % c.nkx;
    float v = 0.f;
    if (co < cout) v = w[((((size_t)co * cin + ci) * c.nkz + kz) * c.nky + ky) * c.nkx + kx];
    out[i] = __float2half_rn(v);
  }
}

bool umma_supported(const ConvArgs& a) {
  UmmaCfg c;
  return plan_umma(a, c);
}

size_t umma_packed_weight_bytes(int cin, int cout, int ntaps, int transposed) {
  if (transposed || cin % 16 != 0) return 0;
  const int cout_pad = (cout + 15) / 16 * 16;
  return (size_t)ntaps * cin * cout_pad * sizeof(__half);
}

int launch_pack_weights_umma(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s) {
  UmmaCfg c;
  if (!plan_umma(a, c)) return FNNU_OK;   // shape not covered: the direct kernel runs it
  const size_t total = (size_t)a.ntaps * a.cin * a.cout_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weights_umma_kernel<<<blocks, 256, 0, s>>>(w_dev, (__half*)out, a.cin, a.cout, c);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

int launch_conv_umma(const ConvArgs& a, cudaStream_t s) {
  UmmaArgs p;
  p.a = a;
  if (!plan_umma(a, p.c)) {
    set_error("conv_umma: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  p.n_units = a.batch * p.c.n_chunks * p.c.n_yblocks * a.out_d[0];
  static bool attr_set = false;
  if (!attr_set) {
    FNNU_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  conv_umma_kernel<<<grid, kThreadsUmma, p.c.smem_bytes, s>>>(p);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
