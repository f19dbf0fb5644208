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
// A is written by 16 producer warps that read the RAW fp16 output of the previous layer from HBM/L2 and
// apply its InstanceNorm affine + LeakyReLU on the way (the fused "normalise on load"); B (weights,
// pre-packed per stage) arrives by one cp.async.bulk (TMA engine) per stage.  One thread issues the
// MMAs; 4 epilogue warps drain TMEM (tcgen05.ld), add the bias, round to fp16, store channels-last and
// accumulate the InstanceNorm sums of the rounded values (fp32 partials -> fp64 atomics).
//
// Warp roles (768 threads, setmaxnreg 80 / 120 / 40): warps 0-15 producers, warps 16-19 epilogue (TMEM lane quarter
// = warp % 4), warp 20 MMA issuer (elect.sync) + TMEM allocator, warps 21-23 only donate registers.
// Pipelines: smem ring (full/empty mbarriers) between producers
// and MMA; TMEM accumulator buffers (full/empty mbarriers) between MMA and epilogue.
//
// Strided convolutions keep the same structure through a phase decomposition: with stride s the tap k
// reads input s*(o + d) + a where k - pad = s*d + a, so the producers write one flattened plane per
// phase a (even / odd rows and columns) and every tap is again a shifted read of one of those planes.
// ConvTranspose3d with kernel == stride is the 1-tap case with the taps folded into N
// (N = prod(stride) * Cout) and an epilogue that scatters each 16-channel group to its output voxel.
//
// Replaces the cuDNN calls behind `self.network(x)` (predict_from_raw_data.py:543) for Conv3d
// (kernel 1|3, stride 1|2 per axis) and ConvTranspose3d (kernel == stride) layers with Cin % 16 == 0;
// other shapes (the 1- or 4-channel first layer) run on conv_ref.cu.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

struct UmmaCfg {
  int ok;
  int transposed;
  int Nc, n_chunks;       // N per MMA, number of (virtual) Cout chunks
  int cout_virtual;       // conv: cout_pad; transposed: prod(stride) * cout_pad
  int KC, G;              // 16-channel chunks per stage, stage groups per kz
  int TY, n_yblocks, T;   // output rows per unit, units per plane, M-tiles per unit
  int rows_mode, tiles_per_row;
  int Dz, Ho, Wo;         // extents of the position space the M-tiles cover (conv: output; transposed: input)
  int pitch;              // positions per row of a flattened phase plane
  int lead_y, lead_x;     // leading pad rows / columns of a phase plane
  int nph_y, nph_x;       // phases per axis (= stride)
  int R, P_fill, P_plane, P_alloc;   // rows and positions per phase plane; positions per 8-channel group
  int nkz, pz, sz, sy, sx;
  int ntyx;
  int tap_aoff[9];        // per in-plane tap: phase * P_plane + shifted start (positions)
  // descriptor offsets (16-byte units) of the (kc, tap) operand pair inside a stage, in issue order: .x for A, .y for
  // B.  Read with one uniform constant load per tap; computing them in the issue loop cost ~20 dependent uniform
  // instructions per tap, which paced every layer with 1-3 M-tiles per tap.
  uint2 tap_desc[36];     // KC <= 4, ntyx <= 9
  int stages, a_stage_bytes, b_stage_bytes;
  int tmem_bufs;
  int smem_bytes;
  unsigned pitch_magic;
};

struct UmmaArgs {
  ConvArgs a;
  UmmaCfg c;
  int n_units;
};

// Six warpgroups: 4 of producers (ncu: the producers are instruction-latency bound -- ~40 instructions per 16-byte
// item at ~4.5 cycles each with 2 warps per scheduler -- so 16 warps instead of 8), 1 epilogue, 1 holding the MMA warp
// (its 3 other warps only donate registers).  setmaxnreg: 512 x 80 + 128 x 120 + 128 x 40 = 768 x 80 registers (the producers
// keep U = 8 16-byte loads in flight per thread: the stride-2 layer was bound by memory-level parallelism).
constexpr int kProducerWarps = 16;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueThreads = 128;
constexpr int kMmaWarp = kProducerWarps + 4;
constexpr int kThreadsUmma = (kMmaWarp + 4) * 32;
constexpr int kRegsProducerU = 80, kRegsMmaU = 40, kRegsEpilogueU = 120;   // must sum to <= 768 x 80: setmaxnreg moves registers inside the CTA's own allocation
constexpr int kSmemLimit = 227 * 1024;

__host__ __device__ inline int tile_base_of(const UmmaCfg& c, int i) {
  return c.rows_mode ? (i / c.tiles_per_row) * c.pitch + (i % c.tiles_per_row) * 128 : i * 128;
}

// k - pad = s*d + a  (floor division): tap k reads phase a at index (o + d)
static inline void tap_split(int k, int pad, int s, int& d, int& a) {
  int v = k - pad;
  d = (v >= 0) ? v / s : -((-v + s - 1) / s);
  a = v - d * s;
}

static bool plan_umma(const ConvArgs& a, UmmaCfg& c) {
  memset(&c, 0, sizeof(c));
  if (a.cin % 16 != 0 || a.cin < 16) return false;
  if (a.src_cs % 8 != 0 || ((uintptr_t)a.src % 16) != 0) return false;
  if (a.cout_pad % 16 != 0) return false;
  c.transposed = a.transposed;
  int dy[3], ay[3], dx[3], ax[3];
  int nky, nkx, trail_y = 0, trail_x = 0;
  if (a.transposed) {
    c.Dz = a.in_d[0]; c.Ho = a.in_d[1]; c.Wo = a.in_d[2];
    c.nkz = 1; c.pz = 0; c.sz = 1; c.sy = 1; c.sx = 1;
    nky = nkx = 1;
    dy[0] = dx[0] = ay[0] = ax[0] = 0;
    c.nph_y = c.nph_x = 1;
    c.cout_virtual = a.ntaps * a.cout_pad;
  } else {
    c.Dz = a.out_d[0]; c.Ho = a.out_d[1]; c.Wo = a.out_d[2];
    c.nkz = a.k[0]; c.pz = a.pad[0]; c.sz = a.s[0]; c.sy = a.s[1]; c.sx = a.s[2];
    nky = a.k[1]; nkx = a.k[2];
    for (int k = 0; k < nky; ++k) {
      tap_split(k, a.pad[1], c.sy, dy[k], ay[k]);
      if (-dy[k] > c.lead_y) c.lead_y = -dy[k];
      if (dy[k] > trail_y) trail_y = dy[k];
    }
    for (int k = 0; k < nkx; ++k) {
      tap_split(k, a.pad[2], c.sx, dx[k], ax[k]);
      if (-dx[k] > c.lead_x) c.lead_x = -dx[k];
      if (dx[k] > trail_x) trail_x = dx[k];
    }
    c.nph_y = c.sy; c.nph_x = c.sx;
    c.cout_virtual = a.cout_pad;
  }
  c.ntyx = nky * nkx;
  if (c.cout_virtual <= 256) {
    c.Nc = c.cout_virtual;
  } else {
    c.Nc = 0;
    for (int n = 256; n >= 16; n -= 16)
      if (c.cout_virtual % n == 0) { c.Nc = n; break; }
    if (!c.Nc) return false;
  }
  c.n_chunks = c.cout_virtual / c.Nc;
  c.pitch = c.Wo + c.lead_x + trail_x;
  if (c.pitch >= 32768) return false;
  c.rows_mode = (c.Wo % 128 == 0);
  c.tiles_per_row = c.rows_mode ? c.Wo / 128 : 0;
  const int nph = c.nph_y * c.nph_x;
  const int misc = 512 + 3 * a.cin * 4 + 8 * c.Nc * 4 + (c.cout_virtual / 16) * 4 + a.cout_pad * 4 + 1024;
  const int avail = kSmemLimit - misc;
  const int center = c.lead_y * c.pitch + c.lead_x;
  int max_shift = 0;   // largest shifted start (relative to the tile base) over the taps
  for (int ky = 0; ky < nky; ++ky)
    for (int kx = 0; kx < nkx; ++kx) {
      int sft = center + dy[ky] * c.pitch + dx[kx];
      if (sft > max_shift) max_shift = sft;
    }
  auto plane_positions = [&](int ty, int T) {
    int R = ty + c.lead_y + trail_y;
    int last_base = c.rows_mode ? ((T - 1) / c.tiles_per_row) * c.pitch + ((T - 1) % c.tiles_per_row) * 128 : (T - 1) * 128;
    int need = last_base + 128 + max_shift;
    if (need < R * c.pitch) need = R * c.pitch;
    return (need + 7) / 8 * 8;
  };
  double best_cost = 1e30;
  int best_ty = 0;
  for (int ty = 1; ty <= c.Ho; ++ty) {
    int T = c.rows_mode ? ty * c.tiles_per_row : ((ty - 1) * c.pitch + c.Wo + 127) / 128;
    if (T * c.Nc > 512 || T > 64) break;
    int palloc = nph * plane_positions(ty, T) + 4;
    int stage1 = 2 * palloc * 16 + c.ntyx * 2 * c.Nc * 16;
    if (2 * stage1 > avail) break;
    int nb = (c.Ho + ty - 1) / ty;
    int R = ty + c.lead_y + trail_y;
    double cost = (double)nb * T * (2 * T * c.Nc <= 512 ? 1.0 : 1.06) * (1.0 + 0.15 * (double)(R - ty) / ty);
    if (cost <= best_cost + 1e-9) { best_cost = cost; best_ty = ty; }   // ties: larger blocks, fewer units
  }
  if (!best_ty) return false;
  c.TY = best_ty;
  c.n_yblocks = (c.Ho + c.TY - 1) / c.TY;
  c.T = c.rows_mode ? c.TY * c.tiles_per_row : ((c.TY - 1) * c.pitch + c.Wo + 127) / 128;
  c.R = c.TY + c.lead_y + trail_y;
  c.P_fill = c.R * c.pitch;
  c.P_plane = plane_positions(c.TY, c.T);
  c.P_alloc = nph * c.P_plane + 4;     // +4: plane stride = 64 mod 128 bytes (fewer st.shared bank conflicts)
  for (int ky = 0; ky < nky; ++ky)
    for (int kx = 0; kx < nkx; ++kx)
      c.tap_aoff[ky * nkx + kx] = (ay[ky] * c.nph_x + ax[kx]) * c.P_plane + center + dy[ky] * c.pitch + dx[kx];
  c.tmem_bufs = (2 * c.T * c.Nc <= 512) ? 2 : 1;
  const int chunks = a.cin / 16;
  c.KC = 0;
  for (int kc = 4; kc >= 1; kc >>= 1) {
    if (chunks % kc) continue;
    int stage = kc * (2 * c.P_alloc * 16 + c.ntyx * 2 * c.Nc * 16);
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
  c.b_stage_bytes = c.KC * c.ntyx * 2 * c.Nc * 16;
  c.smem_bytes = c.stages * (c.a_stage_bytes + c.b_stage_bytes) + misc;
  c.pitch_magic = (unsigned)((0x100000000ull + (unsigned)c.pitch - 1) / (unsigned)c.pitch);
  for (int kc = 0; kc < c.KC; ++kc)
    for (int t = 0; t < c.ntyx; ++t) {
      c.tap_desc[kc * c.ntyx + t].x = (unsigned)(kc * 2 * c.P_alloc + c.tap_aoff[t]);
      c.tap_desc[kc * c.ntyx + t].y = (unsigned)((kc * c.ntyx + t) * 2 * c.Nc);
    }
  if ((c.P_alloc * 16 >> 4) > 0x3FFF || (c.Nc * 16 >> 4) > 0x3FFF) return false;
  c.ok = 1;
  return true;
}

// v[0..15] per lane, 32 lanes: afterwards v[0] of lanes 2 i and 2 i + 1 holds the sum over the warp of value i.
__device__ __forceinline__ void warp_transpose_reduce16(float (&v)[16], int lane) {
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

struct UnitIdx {
  int b, z, yb, nc;
};
__device__ __forceinline__ UnitIdx decode_unit(const UmmaArgs& p, int u) {
  UnitIdx r;
  r.z = u % p.c.Dz;
  u /= p.c.Dz;
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
  uint32_t* tile_tab = tmem_slot + 4;                       // [64] tile base positions
  float* xs = reinterpret_cast<float*>(tile_tab + 64);
  float* xh = xs + a.cin;
  float* xl = xh + a.cin;
  float* stat_s = xl + a.cin;                               // [4 epilogue warps][2][Nc] partial sums of one unit
  float* bias_s = stat_s + 8 * c.Nc;                        // [cout_pad]
  uint32_t* grp_tab = reinterpret_cast<uint32_t*>(bias_s + a.cout_pad);   // per 16-column group: co0 | tap offsets

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Din = a.in_d[0], Hin = a.in_d[1], Win = a.in_d[2];

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full_bar[s], kProducerWarps);       // one arrival per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpilogueThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < a.cout_pad; i += blockDim.x) bias_s[i] = (a.bias && i < a.cout) ? __ldg(a.bias + i) : 0.f;
  for (int g = threadIdx.x; g < c.cout_virtual / 16; g += blockDim.x) {
    // (virtual) output channel 16 g: real channel group and, for a transposed conv, the output offset of its tap
    uint32_t e = (uint32_t)(g * 16);
    if (c.transposed) {
      const int tap = (g * 16) / a.cout_pad;
      const int co0 = g * 16 - tap * a.cout_pad;
      const int ddx = tap % a.s[2], ddy = (tap / a.s[2]) % a.s[1], ddz = tap / (a.s[2] * a.s[1]);
      e = (uint32_t)co0 | ((uint32_t)ddz << 16) | ((uint32_t)ddy << 20) | ((uint32_t)ddx << 24);
    }
    grp_tab[g] = e;
  }
  for (int i = threadIdx.x; i < c.T; i += blockDim.x) tile_tab[i] = (uint32_t)tile_base_of(c, i);
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_cols = (uint32_t)(c.T * c.Nc);

  if (warp < kProducerWarps) {
    // =========================== PRODUCERS ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducerU));
    const int tid = threadIdx.x;
    const int Q = 2 * c.KC;               // 8-channel groups per stage (2, 4 or 8)
    const int qshift = (Q == 2) ? 1 : (Q == 4 ? 2 : 3);
    const int q = tid & (Q - 1);
    const int items = c.P_fill << qshift;
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
        const int z_in = ui.z * c.sz + kz - c.pz;
        if (z_in < 0 || z_in >= Din) continue;
        const __half* plane = a.src + ((size_t)ui.b * Din + z_in) * Hin * Win * a.src_cs;
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
          __half2 s2[4], t2[4], l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            s2[e] = __floats2half2_rn(xs[ch0 + 2 * e], xs[ch0 + 2 * e + 1]);
            t2[e] = __floats2half2_rn(xh[ch0 + 2 * e], xh[ch0 + 2 * e + 1]);
            l2[e] = __floats2half2_rn(xl[ch0 + 2 * e], xl[ch0 + 2 * e + 1]);
          }
          // all phase planes of the stage form ONE item space (a strided layer's planes are small: enc1.0 has 4 planes
          // of 780 items for 512 threads, and looping over them one by one exposed a global-load latency per plane)
          {
            uint8_t* a_q = a_s + (size_t)q * c.P_alloc * 16;
            const int n_ph = c.nph_y * c.nph_x;
            const int all_items = n_ph * items;
            constexpr int U = 4;
            for (int j0 = tid; j0 < all_items; j0 += kProducerThreads * U) {
              uint4 raw[U];
              int pos[U];
              bool ok[U];
#pragma unroll
              for (int k = 0; k < U; ++k) {
                // item order (fastest first): channel group q, x phase, position, y phase -- consecutive lanes read
                // consecutive global bytes even with stride 2 in x (per-phase order fetched every 64-byte L2 line twice)
                const int jg = j0 + k * kProducerThreads;
                const int idx = jg >> qshift;
                const int phx = idx & (c.nph_x - 1);                                     // nph_x, nph_y are 1 or 2
                const int t = idx >> (c.nph_x - 1);
                const int phy = (t >= c.P_fill) ? 1 : 0;
                const int pp = t - phy * c.P_fill;
                const int ph = phy * c.nph_x + phx;
                pos[k] = (jg < all_items) ? ph * c.P_plane + pp : -1;
                const int r = (int)__umulhi((unsigned)pp, c.pitch_magic);
                const int xp = pp - r * c.pitch;
                const int y_in = c.sy * (y0 + r - c.lead_y) + phy;
                const int x_in = c.sx * (xp - c.lead_x) + phx;
                ok[k] = (jg < all_items) && y_in >= 0 && y_in < Hin && x_in >= 0 && x_in < Win;
                raw[k] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[k]) raw[k] = __ldg(reinterpret_cast<const uint4*>(plane + ((size_t)y_in * Win + x_in) * a.src_cs + ch0));
              }
#pragma unroll
              for (int k = 0; k < U; ++k) {
                if (pos[k] < 0) continue;
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (ok[k]) {
                  o = xform8_h2(raw[k], s2, t2, l2);
                }
                *reinterpret_cast<uint4*>(a_q + (size_t)pos[k] * 16) = o;
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_warp(&full_bar[stage]);
          if (++stage == c.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMmaU));
    if (warp == kMmaWarp) {
    // The whole warp runs the (uniform) control flow; one lane, chosen with elect.sync, issues (ptxas then knows a
    // single thread executes the UTCHMMAs and does not wrap each one in an R2UR serialisation loop).  Per MMA: one
    // 64-bit add on the A descriptor and one add on the TMEM column, both warp-uniform.
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.Nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lbo = (uint32_t)c.P_alloc * 16, b_lbo = (uint32_t)c.Nc * 16;
    const uint64_t a_desc0 = make_desc(0, a_lbo, 128);
    const uint64_t b_desc0 = make_desc(0, b_lbo, 128);
    const uint64_t desc_hi_a = a_desc0 & 0xffffffff00000000ull, desc_hi_b = b_desc0 & 0xffffffff00000000ull;
    // tile bases: flat mode i*128; one tile per row (W == 128) i*pitch; several tiles per row: nested
    const bool nested = c.rows_mode && c.tiles_per_row > 1;
    const int n_outer = nested ? c.TY : 1;
    const int n_inner = nested ? c.tiles_per_row : c.T;
    const uint32_t inner_step = (c.rows_mode && !nested) ? (uint32_t)c.pitch : 128u;   // positions == 16-byte units
    const uint32_t outer_step = (uint32_t)c.pitch;
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t tphase0 = 0, tphase1 = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitIdx ui = decode_unit(p, u);
      mbar_wait(buf ? &tempty_bar[1] : &tempty_bar[0], (buf ? tphase1 : tphase0) ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)buf * buf_cols;
      bool first = true;
      for (int kz = 0; kz < c.nkz; ++kz) {
        const int z_in = ui.z * c.sz + kz - c.pz;
        if (z_in < 0 || z_in >= Din) continue;
        for (int g = 0; g < c.G; ++g) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + (size_t)stage * stage_bytes);
          const uint32_t b_base = a_base + (uint32_t)c.a_stage_bytes;
          if (elect_one()) {
          // Descriptor arithmetic on the LOW word only: the 14-bit start-address field never carries into the LBO
          // field (shared memory < 256 KB), so each MMA costs one 32-bit uniform add per operand instead of a 64-bit
          // add-with-carry chain (the issue loop, not the tensor pipe, paced layers with few tiles per tap).
          const uint32_t a_lo0 = (uint32_t)a_desc0 + (a_base >> 4), b_lo0 = (uint32_t)b_desc0 + (b_base >> 4);
          const int n_taps = c.KC * c.ntyx;
          for (int kt = 0; kt < n_taps; ++kt) {
            const uint2 off = c.tap_desc[kt];
            const uint64_t db = desc_hi_b | (uint64_t)(b_lo0 + off.y);
            const uint32_t da_lo0 = a_lo0 + off.x;
            const uint32_t accum = (first && kt == 0) ? 0u : 1u;
            // tile i of the unit sits at row (i / tiles_per_row), column block (i % tiles_per_row): pure
            // uniform-register arithmetic per MMA (no memory reads on the issue path)
            uint32_t da_row = da_lo0;
            for (int r = 0; r < n_outer; ++r) {
#pragma unroll 4
              for (int j = 0; j < n_inner; ++j) {
                umma_f16(d_base + (uint32_t)((r * n_inner + j) * c.Nc), desc_hi_a | (uint64_t)(da_row + (uint32_t)j * inner_step), db, idesc, accum);
              }
              da_row += outer_step;
            }
          }
          umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          first = false;
          if (++stage == c.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) umma_commit(buf ? &tfull_bar[1] : &tfull_bar[0]);
      __syncwarp();
      if (buf) tphase1 ^= 1; else tphase0 ^= 1;
      if (c.tmem_bufs == 2) buf ^= 1;
    }
    }
  } else {
    // =========================== EPILOGUE ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogueU));
    const int wq = warp & 3;                     // TMEM lane quarter this warp may access
    const int et = threadIdx.x - kProducerThreads;   // 0..127
    int buf = 0;
    uint32_t tphase0 = 0, tphase1 = 0;
    const bool vec_store = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0);
    const bool do_stats = a.dst_stats != nullptr;
    const int Dout = a.out_d[0], Hout = a.out_d[1], Wout = a.out_d[2];
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const UnitIdx ui = decode_unit(p, u);
      const int y0 = ui.yb * c.TY;
      mbar_wait(buf ? &tfull_bar[1] : &tfull_bar[0], buf ? tphase1 : tphase0);
      if (buf) tphase1 ^= 1; else tphase0 ^= 1;
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(wq * 32) << 16);
      __half* out_b = a.dst + (size_t)ui.b * Dout * Hout * Wout * a.dst_cs;
      for (int n0 = 0; n0 < c.Nc; n0 += 16) {
        const uint32_t ge = grp_tab[(ui.nc * c.Nc + n0) >> 4];
        const int co0 = (int)(ge & 0xffffu);
        int oz = ui.z, oy_off = 0, ox_off = 0, ys = 1, xsn = 1;
        if (c.transposed) {
          oz = ui.z * a.s[0] + (int)((ge >> 16) & 15u);
          oy_off = (int)((ge >> 20) & 15u); ox_off = (int)((ge >> 24) & 15u); ys = a.s[1]; xsn = a.s[2];
        }
        float bias[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + co0 + j);
          bias[j] = bv.x; bias[j + 1] = bv.y; bias[j + 2] = bv.z; bias[j + 3] = bv.w;
        }
        float s1[16], s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s1[j] = s2[j] = 0.f;
        const bool full16 = co0 + 16 <= a.cout;
        __half* out_plane = out_b + (size_t)oz * Hout * Wout * a.dst_cs;
        // two tiles per TMEM round trip: the tcgen05.ld latency was exposed once per tile
        auto emit_tile = [&](int i, const uint32_t (&acc)[16]) {
          const int pos = (int)tile_tab[i] + wq * 32 + lane;
          const int yo = (int)__umulhi((unsigned)pos, c.pitch_magic);
          const int xo = pos - yo * c.pitch;
          const int y = y0 + yo;
          const bool valid = xo < c.Wo && yo < c.TY && y < c.Ho;
          if (valid) {
            __half2 hv[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float v0 = __uint_as_float(acc[j]) + bias[j];
              const float v1 = __uint_as_float(acc[j + 1]) + bias[j + 1];
              if (do_stats) {   // fp32 values: the fp16 rounding error averages out over the patch
                s1[j] += v0;
                s2[j] = fmaf(v0, v0, s2[j]);
                s1[j + 1] += v1;
                s2[j + 1] = fmaf(v1, v1, s2[j + 1]);
              }
              hv[j >> 1] = __floats2half2_rn(v0, v1);
            }
            __half* q = out_plane + ((size_t)(y * ys + oy_off) * Wout + (xo * xsn + ox_off)) * a.dst_cs + co0;
            if (vec_store && full16) {
              reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
              reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[4]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < a.cout) q[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
            }
          }
        };
        for (int i = 0; i < c.T; i += 2) {
          uint32_t acc0[16], acc1[16];
          const bool two = i + 1 < c.T;
          tmem_ld16_nowait(d_base + (uint32_t)(i * c.Nc + n0), acc0);
          if (two) tmem_ld16_nowait(d_base + (uint32_t)((i + 1) * c.Nc + n0), acc1);
          tmem_wait_ld();
          emit_tile(i, acc0);
          if (two) emit_tile(i + 1, acc1);
        }
        if (do_stats) {
          // 16 values x 32 lanes -> lane pair (2 i, 2 i + 1) ends up with the warp total of value i: each exchange step
          // halves the values a lane keeps (16 shuffles per array instead of 80), and the totals go to this warp's own
          // slots with plain stores (the shared-memory atomics of four warps on the same words were the slow part)
          warp_transpose_reduce16(s1, lane);
          warp_transpose_reduce16(s2, lane);
          if (!(lane & 1)) {
            float* slot = stat_s + (size_t)wq * 2 * c.Nc + n0 + (lane >> 1);
            slot[0] = s1[0];
            slot[c.Nc] = s2[0];
          }
        }
      }
      tc_fence_before();
      mbar_arrive_warp(buf ? &tempty_bar[1] : &tempty_bar[0]);
      if (c.tmem_bufs == 2) buf ^= 1;
      if (a.dst_stats) {
        named_bar_sync(2, kEpilogueThreads);
        for (int i = et; i < 2 * c.Nc; i += kEpilogueThreads) {
          const int which = i / c.Nc, n = i - which * c.Nc;
          const int co = ui.nc * c.Nc + n;
          const float tot = (stat_s[i] + stat_s[2 * c.Nc + i]) + (stat_s[4 * c.Nc + i] + stat_s[6 * c.Nc + i]);
          if (co < a.cout) atomicAdd(a.dst_stats + ((size_t)ui.b * a.dst_stat_stride + co) * 2 + which, (double)tot);
        }
        named_bar_sync(2, kEpilogueThreads);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: [cout][cin][kz][ky][kx] fp32 -> per-stage blobs of fp16
//   [n-chunk][kz][group g][chunk kc][tap (ky,kx)][8-channel half][n < Nc][8 halves]
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_umma_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout,
                                         int cout_pad, int ntaps, int nky, int nkx, UmmaCfg c) {
  const size_t total = (size_t)c.n_chunks * c.nkz * c.G * c.KC * c.ntyx * 2 * c.Nc * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % c.Nc); r /= c.Nc;
    const int h = (int)(r % 2); r /= 2;
    const int t = (int)(r % c.ntyx); r /= c.ntyx;
    const int kc = (int)(r % c.KC); r /= c.KC;
    const int g = (int)(r % c.G); r /= c.G;
    const int kz = (int)(r % c.nkz); r /= c.nkz;
    const int nc = (int)r;
    const int v = nc * c.Nc + n;                      // (virtual) output channel
    const int ci = (g * c.KC + kc) * 16 + h * 8 + e;
    float val = 0.f;
    if (c.transposed) {
      const int tap = v / cout_pad, co = v - tap * cout_pad;    // weight [cin][cout][taps]
      if (co < cout) val = w[((size_t)ci * cout + co) * ntaps + tap];
    } else {
      const int ky = t / nkx, kx = t % nkx;                     // weight [cout][cin][kz][ky][kx]
      if (v < cout) val = w[((((size_t)v * cin + ci) * c.nkz + kz) * nky + ky) * nkx + kx];
    }
    out[i] = __float2half_rn(val);
  }
}

bool umma_supported(const ConvArgs& a) {
  UmmaCfg c;
  return plan_umma(a, c);
}

size_t umma_packed_weight_bytes(int cin, int cout, int ntaps, int transposed) {
  (void)transposed;
  if (cin % 16 != 0) return 0;
  const int cout_pad = (cout + 15) / 16 * 16;
  return (size_t)ntaps * cin * cout_pad * sizeof(__half);
}

int launch_pack_weights_umma(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s) {
  UmmaCfg c;
  if (!plan_umma(a, c)) return FNNU_OK;   // shape not covered: the direct kernel runs it
  const size_t total = (size_t)a.ntaps * a.cin * a.cout_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weights_umma_kernel<<<blocks, 256, 0, s>>>(w_dev, (__half*)out, a.cin, a.cout, a.cout_pad, a.ntaps,
                                                  a.transposed ? 1 : a.k[1], a.transposed ? 1 : a.k[2], c);
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
  p.n_units = a.batch * p.c.n_chunks * p.c.n_yblocks * p.c.Dz;
  /* the attribute is per device: set it on every launch (cheap) */ FNNU_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  static const bool show_plan = getenv("FNNU_SHOW_PLAN") != nullptr;   // debugging aid: one line per launch
  if (show_plan) {
    const UmmaCfg& c = p.c;
    fprintf(stderr,
            "conv_umma: cin=%d cout=%d k=%dx%dx%d s=%d%d%d in=%dx%dx%d batch=%d%s | Nc=%d chunks=%d KC=%d G=%d TY=%d yblocks=%d T=%d "
            "rows_mode=%d pitch=%d R=%d P_plane=%d stages=%d a_stage=%d b_stage=%d tmem_bufs=%d units=%d smem=%d\n",
            a.cin, a.cout, a.k[0], a.k[1], a.k[2], a.s[0], a.s[1], a.s[2], a.in_d[0], a.in_d[1], a.in_d[2], a.batch,
            a.transposed ? " transposed" : "", c.Nc, c.n_chunks, c.KC, c.G, c.TY, c.n_yblocks, c.T, c.rows_mode, c.pitch, c.R,
            c.P_plane, c.stages, c.a_stage_bytes, c.b_stage_bytes, c.tmem_bufs, p.n_units, c.smem_bytes);
  }
  conv_umma_kernel<<<grid, kThreadsUmma, p.c.smem_bytes, s>>>(p);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
