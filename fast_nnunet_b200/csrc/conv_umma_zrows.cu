// Row-streaming tcgen05 convolution, second generation: TWO output z-planes per pass ("z-pair rows").
// For Conv3d 3x3x3, stride 1, Cin in {16, 32}, Cout <= 16, 64 <= W <= 128, even D: enc0.1 / dec5.0 / dec5.1 of the
// students, i.e. 48 % of a forward's FLOPs.
//
// What round 1 measured on conv_umma_rows.cu (profiles/README.md) and what this kernel does about it:
//  * a 128 x N x 16 tcgen05.mma with A and B in shared memory is bound by the operand fetch below N = 128
//    (44 / 56 cycles for N = 48 / 96): the 4 KB A tile is re-read for every MMA.  Here a unit covers output planes
//    z0 and z0+1, whose accumulator tiles sit side by side in TMEM, so the two middle input planes (z0, z0+1) feed
//    BOTH tiles with one N = 96 MMA per (kx, 16 input channels): 12 MMAs = 600 cycles per TWO output rows
//    instead of 2 x 9 MMAs = 792, and every input row is staged 2x instead of 3x.
//      weights: B = [kx][chunk][half][n = (2 - kz) * 48 + ky * 16 + co][8]  (N extent 144)
//      plane z0-1: kz = 0 -> tile 0, B columns [96, 144)      plane z0  : kz = 1 | 0 -> tiles 0 | 1, columns [48, 144)
//      plane z0+1: kz = 2 | 1 -> tiles 0 | 1, columns [0, 96)  plane z0+2: kz = 2 -> tile 1, columns [0, 48)
//    (plane z0 is issued first: its N = 96 MMA is the "accumulate = 0" first touch of both tiles).
//  * every mbarrier wait / tcgen05.commit of the issuing thread costs 70-200 cycles that do not overlap with the
//    tensor pipe: one wait pair and ONE commit per y-step (12-24 MMAs) instead of two waits + two commits per 9 MMAs
//    — the same barrier releases the shared-memory stage to the producers and publishes the tiles to the
//    epilogue — and two issuing warps take alternate y-steps so that one warp's waits hide behind the other's MMAs.
//  * producers: LDG.128 -> InstanceNorm + LeakyReLU in registers -> STS.128 (no cp.async + in-place LDS/STS pass);
//    halo positions are zeroed once per kernel.  The transform subtracts the fp16-rounded mean first, so its
//    rounding error scales with the normalised value and not with mean * scale.
//  * eight epilogue warps (one set of four per output plane) drain the tiles: out[y] = T[y-1][ky 0] + T[y][ky 1] +
//    T[y+1][ky 2], the same-lane sum of conv_umma_rows.cu.
//
// Shapes (template <CHUNKS = Cin / 16, CP = padded Cout, ZC = output planes per pass, NKZ = kernel extent along z>):
//   <1|2, 16, 2, 3>  the z-pair form described above (even depth)
//   <1|2, 16, 1, 1|3>, <1|2, 32, 1, 1|3>  one output plane per pass: 1x3x3 kernels (anisotropic first stages), odd depths
//                    and Cout = 32 (two tiles of 96 columns do not leave room for five y-steps in TMEM); with Cout = 32
//                    the two epilogue warp sets split the channels instead of the planes.
// A y-step's shared-memory stage holds NPL = ZC + NKZ - 1 input planes; input plane pl feeds output tiles
// zt in [max(0, pl - NKZ + 1), min(ZC - 1, pl)] with kz = pl - zt, one MMA per (kx, 16 input channels) whose N spans those
// tiles: weights are packed [kx][chunk][half][n = (NKZ - 1 - kz) * 3 CP + ky * CP + co][8].
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

#ifdef FNNU_ZROWS_PROF
// clock64 accounting per warp role of CTA 0 (temporary profiling builds only): [role * 8 + counter]
__device__ long long g_zprof[32];
#define ZPROF_T(var) const long long var = clock64()
#define ZPROF_ADD(slot, expr) do { if (blockIdx.x == 0) zp[slot] += (expr); } while (0)
#else
#define ZPROF_T(var)
#define ZPROF_ADD(slot, expr)
#endif

namespace {

// 12 producer warps with 88 registers each (8 independent 16-byte loads in flight per thread, no spills) measured
// FASTER than 16 warps with 72 registers (enc0.1: 1.51 vs 2.01 ms per 32 patches): the producers are bound by
// latency per warp, and the register budget is what buys loads in flight.
constexpr int kZProducerWarps = 12;
constexpr int kZEpilogueWarp0 = 12;          // warps 12-15: plane z0, warps 16-19: plane z0 + 1
constexpr int kZMmaWarp0 = 20;               // warps 20, 21 issue; 22, 23 only donate registers
constexpr int kZThreads = 24 * 32;
constexpr int kZRegsProducer = 80, kZRegsEpilogue = 96, kZRegsMma = 40;   // 384 x 80 + 256 x 96 + 128 x 40 <= 768 x 80
constexpr int kZProw = 140;                  // positions per (plane, 8-channel group): 1 + 128 + 1, padded to 4 (mod 8)
constexpr int kZMaxSlots = 8;                // y-steps resident in TMEM: 512 / (ZC * 3 * CP) columns, at most 8
constexpr int kZStepBars = 16;               // ring of "y-step done" barriers (> stages, > slots)
constexpr int kZMaxStages = 12;
constexpr int kZSmemLimit = 227 * 1024;

struct ZCfg {
  int D, H, W;
  int chunks, cp, zc, nkz, npl;
  int stage_bytes, stages, w_bytes, smem_bytes;
  int n_yseg, seg_rows;
  int issuers;
  int prefetch_distance;   // y-steps the L2 prefetcher runs ahead of the tensor pipe (0 = off)
};

struct ZArgs {
  ConvArgs a;
  ZCfg c;
  int n_units;
};

bool plan_zrows(const ConvArgs& a, ZCfg& c) {
  memset(&c, 0, sizeof(c));
  if (a.transposed) return false;
  if (a.s[0] != 1 || a.s[1] != 1 || a.s[2] != 1) return false;
  if ((a.k[0] != 3 && a.k[0] != 1) || a.k[1] != 3 || a.k[2] != 3) return false;
  if (a.cin != 16 && a.cin != 32) return false;
  if (a.cout_pad != 16 && a.cout_pad != 32) return false;
  if (a.src_cs % 8 != 0 || ((uintptr_t)a.src % 16) != 0) return false;
  c.D = a.in_d[0]; c.H = a.in_d[1]; c.W = a.in_d[2];
  if (c.W > 128 || c.W < 64 || c.H < 8) return false;
  c.chunks = a.cin / 16;
  c.cp = a.cout_pad;
  c.nkz = a.k[0];
  c.zc = (c.cp == 16 && c.nkz == 3 && !(c.D & 1)) ? 2 : 1;
  c.npl = c.zc + c.nkz - 1;
  c.stage_bytes = c.npl * 2 * c.chunks * kZProw * 16;
  c.w_bytes = 3 * c.chunks * 2 * (c.nkz * 3 * c.cp) * 16;
  const int misc = 2048;
  int st = (kZSmemLimit - misc - c.w_bytes) / c.stage_bytes;
  if (st > kZMaxStages) st = kZMaxStages;
  const int groups = kZProducerWarps / (2 * c.chunks);
  if (st <= groups) return false;
  c.stages = st;
  c.smem_bytes = c.w_bytes + c.stages * c.stage_bytes + misc;
  c.n_yseg = 1;
  while ((long long)a.batch * (c.D / c.zc) * c.n_yseg < 3LL * num_sms() && c.H / (c.n_yseg * 2) >= 8) c.n_yseg *= 2;
  c.seg_rows = (c.H + c.n_yseg - 1) / c.n_yseg;
  static int issuers = 0;
  if (!issuers) {
    const char* e = getenv("FNNU_ZROWS_ISSUERS");
    issuers = (e && atoi(e) == 1) ? 1 : 2;
  }
  c.issuers = issuers;
  static int pf = -1;
  if (pf < 0) {
    const char* e = getenv("FNNU_ZROWS_PREFETCH");
    pf = e ? atoi(e) : 8;
    if (pf < 0) pf = 0;
    if (pf > kZStepBars - 2) pf = kZStepBars - 2;
  }
  c.prefetch_distance = pf;
  return true;
}

// Geometry of one (CP, ZC, NKZ) instantiation, all compile-time.
template <int CP, int ZC, int NKZ>
struct ZGeo {
  static constexpr int TILE_N = 3 * CP;                 // one output plane's accumulator tile: (ky, co)
  static constexpr int SLOT_COLS = ZC * TILE_N;         // a y-step's TMEM columns
  static constexpr int SLOTS = (512 / SLOT_COLS) < kZMaxSlots ? (512 / SLOT_COLS) : kZMaxSlots;
  static constexpr int NPL = ZC + NKZ - 1;              // input planes per stage
  static constexpr int PZ = (NKZ - 1) / 2;              // plane PZ is output plane z0 itself: always inside the volume
  static constexpr int BN = NKZ * TILE_N;               // N extent of the packed weights
  static constexpr int zt_lo(int pl) { return pl - NKZ + 1 > 0 ? pl - NKZ + 1 : 0; }
  static constexpr int zt_hi(int pl) { return pl < ZC - 1 ? pl : ZC - 1; }
  static constexpr int n_of(int pl) { return (zt_hi(pl) - zt_lo(pl) + 1) * TILE_N; }
  static constexpr int d_off(int pl) { return zt_lo(pl) * TILE_N; }
  static constexpr int b_col(int pl) { return (NKZ - 1 - (pl - zt_lo(pl))) * TILE_N; }
};

// MMAs of one input plane: 3 kx x CHUNKS, descriptor offsets (16-byte units) are template constants.
template <int CHUNKS, int PL, int BN, int NCOL0, int I>
__device__ __forceinline__ void z_issue_plane(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t& accum) {
  if constexpr (I < 3 * CHUNKS) {
    constexpr int kx = I / CHUNKS, kc = I % CHUNKS;
    constexpr uint32_t a_off = (uint32_t)((PL * 2 * CHUNKS + kc * 2) * kZProw + kx);
    constexpr uint32_t b_off = (uint32_t)(((kx * CHUNKS + kc) * 2) * BN + NCOL0);
    umma_f16_off<a_off, b_off>(d, da, db, idesc, accum);
    accum = 1;
    z_issue_plane<CHUNKS, PL, BN, NCOL0, I + 1>(d, da, db, idesc, accum);
  }
}

// every plane of a y-step except the first-touch plane PZ: first (FULL = true) the planes whose MMAs span the whole
// slot, then the edge planes — MMAs of equal N back to back (the order 96, 48, 96, 48 measured 20 % slower on the
// Cin = 32 layer than 96, 96, 48, 48)
template <int CHUNKS, int CP, int ZC, int NKZ, int PL, bool FULL>
__device__ __forceinline__ void z_issue_other_planes(uint32_t d, uint64_t da, uint64_t db, uint32_t pl_ok, uint32_t& accum) {
  using G = ZGeo<CP, ZC, NKZ>;
  if constexpr (PL < G::NPL) {
    if constexpr (PL != G::PZ && (G::n_of(PL) == G::SLOT_COLS) == FULL) {
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(G::n_of(PL) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      if ((pl_ok >> PL) & 1)
        z_issue_plane<CHUNKS, PL, G::BN, G::b_col(PL), 0>(d + (uint32_t)G::d_off(PL), da, db, idesc, accum);
    }
    z_issue_other_planes<CHUNKS, CP, ZC, NKZ, PL + 1, FULL>(d, da, db, pl_ok, accum);
  }
}

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void sts16(uint32_t addr, const uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// y = lrelu((x - m) * s + t) on 8 fp16 channels, packed half2 arithmetic
__device__ __forceinline__ uint4 zxform8(const uint4 raw, const __half2* m2, const __half2* s2, const __half2* t2, const __half2* l2) {
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

template <int CHUNKS, int CP, int ZC, int NKZ>
__global__ void __launch_bounds__(kZThreads, 1) conv_umma_zrows_kernel(const __grid_constant__ ZArgs p) {
  using G = ZGeo<CP, ZC, NKZ>;
  constexpr int kZSlots = G::SLOTS;
  constexpr int kZSlotCols = G::SLOT_COLS;
  // epilogue warp sets that have work: one per output plane (ZC = 2) or one per half of 32 output channels (CP = 32)
  constexpr int kSets = (ZC == 2 || CP == 32) ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const ZCfg& c = p.c;
  const ConvArgs& a = p.a;
  uint8_t* w_s = smem;
  uint8_t* ring = smem + c.w_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)c.stages * c.stage_bytes);   // [kZMaxStages]
  uint64_t* step_bar = full_bar + kZMaxStages;                                                  // [kZStepBars]
  uint64_t* tempty_bar = step_bar + kZStepBars;                                                 // [kZMaxSlots]
  uint64_t* w_bar = tempty_bar + kZMaxSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);                                      // [32]
  volatile int* progress = reinterpret_cast<volatile int*>(bias_s + 32);                        // [2] y-steps issued per issuer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kGroupWarps = 2 * CHUNKS;
  constexpr int kGroups = kZProducerWarps / kGroupWarps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) mbar_init(&full_bar[s], kGroupWarps);
    for (int s = 0; s < kZStepBars; ++s) mbar_init(&step_bar[s], 1);
    for (int s = 0; s < kZSlots; ++s) mbar_init(&tempty_bar[s], 4 * kSets);       // one arrival per epilogue warp
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) bias_s[threadIdx.x] = (a.bias && (int)threadIdx.x < a.cout) ? a.bias[threadIdx.x] : 0.f;
  if (threadIdx.x < 2) progress[threadIdx.x] = 0;
  // halo positions (and everything else) start as zeros; producers only ever write in-image positions
  {
    uint4* r4 = reinterpret_cast<uint4*>(ring);
    const int n16 = c.stages * c.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kZThreads) r4[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kZMmaWarp0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(w_bar, (uint32_t)c.w_bytes);
    bulk_g2s(w_s, a.w_umma, (uint32_t)c.w_bytes, w_bar);
  }

  // unit u -> (sample b, z chunk zp, y segment): output planes ZC zp .. ZC zp + ZC - 1, rows [ya, yb)
  const int n_zp = c.D / ZC;
  auto decode = [&](int u, int& b, int& z0, int& ya, int& yb) {
    const int zp = u % n_zp;
    u /= n_zp;
    const int seg = u % c.n_yseg;
    b = u / c.n_yseg;
    z0 = ZC * zp;
    ya = seg * c.seg_rows;
    yb = ya + c.seg_rows;
    if (yb > c.H) yb = c.H;
  };

  // bit pl: input plane z0 - PZ + pl lies inside the volume (the others are skipped by producers and issuers alike)
  auto planes_ok = [&](int z0) {
    uint32_t m = 0;
#pragma unroll
    for (int pl = 0; pl < G::NPL; ++pl) {
      const int zi = z0 - G::PZ + pl;
      if (zi >= 0 && zi < c.D) m |= 1u << pl;
    }
    return m;
  };

  if (warp < kZProducerWarps) {
    // =========================== PRODUCERS ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kZRegsProducer));
    const int grp = warp / kGroupWarps;
    const int wg = warp - grp * kGroupWarps;
    // a warp covers 16 positions x one PAIR of 8-channel groups (32 contiguous bytes per voxel), so "these channels
    // carry no pending transform" (ConvTranspose3d half of a decoder concat) is warp-uniform
    const int q = 2 * (wg & (CHUNKS - 1)) + (lane & 1);
    const int x0 = (wg / CHUNKS) * 16 + (lane >> 1);
    bool q_identity = true;
#pragma unroll
    for (int e = 0; e < 16; ++e) q_identity = q_identity && (a.src_meta[(q & ~1) * 8 + e].eps < 0.f);
    uint32_t in_w = 0;                                       // bit it: x0 + 32 it < W
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (x0 + 32 * it < c.W) in_w |= 1u << it;
    const uint32_t goff0 = (uint32_t)(x0 * a.src_cs + q * 8) * 2u;      // + it * 32 * src_cs * 2
    const uint32_t gstep = (uint32_t)(32 * a.src_cs) * 2u;
    const uint32_t plane_bytes = (uint32_t)(2 * CHUNKS * kZProw) * 16u;
    const uint32_t my_off = (uint32_t)(q * kZProw + x0 + 1) * 16u;      // + it * 512
    // one 64-bit base per unit; everything below it is 32-bit offsets (a plane of 128 x 128 x 32 fp16 is 1 MB)
    const uint32_t row_bytes = (uint32_t)(c.W * a.src_cs * 2);
    const uint32_t zplane_bytes = (uint32_t)c.H * row_bytes;
    const uint32_t ring_u32 = smem_u32(ring);
    const bool full_w = c.W == 128;                          // every lane's four positions lie inside the row
    int t = 0, t_grp = 0, stage = 0;                         // step counter, its group, its stage
    int cur_b = -1;
    __half2 m2[4], s2[4], t2[4], l2[4];
#ifdef FNNU_ZROWS_PROF
    long long zp[4] = {0, 0, 0, 0};
    const long long zt0 = clock64();
#endif
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int b, z0, ya, yb;
      decode(u, b, z0, ya, yb);
      if (b != cur_b) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float mh[2], sc[2], sh[2], sl[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = q * 8 + 2 * e + k;
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
      const int n_rows = (yb - ya) + 2;
      // planes z0 - PZ .. z0 - PZ + NPL - 1; out-of-volume planes are skipped here and by the MMA warps
      const uint32_t pl_ok = planes_ok(z0);
      const char* vol0 = reinterpret_cast<const char*>(a.src) + ((long long)b * c.D + (z0 - G::PZ)) * (long long)zplane_bytes + goff0;
      for (int j = 0; j < n_rows; ++j) {
        const bool mine = t_grp == grp;
        const int my_stage = stage;
        const int my_t = t;
        ++t;
        if (++t_grp == kGroups) t_grp = 0;
        if (++stage == c.stages) stage = 0;
        if (!mine) continue;
        const int y_in = ya - 1 + j;
        const bool row_ok = y_in >= 0 && y_in < c.H;
        const uint32_t st = ring_u32 + (uint32_t)my_stage * (uint32_t)c.stage_bytes + my_off;
        auto wait_stage_free = [&]() {
          // the loads do not need the stage: only the stores wait for its previous y-step to be consumed
          if (my_t >= c.stages) {
            const int tp = my_t - c.stages;
            ZPROF_T(w0);
            mbar_wait(&step_bar[tp & (kZStepBars - 1)], (uint32_t)(tp >> 4) & 1u);
            ZPROF_ADD(0, clock64() - w0);
          }
        };
        if (!row_ok) {
          // a row outside the image: zeros (the conv's zero padding applies to the NORMALISED activations)
          wait_stage_free();
#pragma unroll
          for (int pl = 0; pl < G::NPL; ++pl)
            if ((pl_ok >> pl) & 1) {
#pragma unroll
              for (int it = 0; it < 4; ++it)
                if ((in_w >> it) & 1) sts16(st + pl * plane_bytes + it * 512, make_uint4(0, 0, 0, 0));
            }
        } else {
          const uint32_t row_off = (uint32_t)y_in * row_bytes;
          // unrolled on purpose: with `#pragma unroll 1` enc0.1 went from 1.40 to 1.83 ms per 32 patches
#pragma unroll
          for (int hp = 0; hp < (G::NPL + 1) / 2; ++hp) {
            // two planes per round: 8 independent 16-byte loads in flight per thread; plane validity is CTA-uniform
            const bool va = (pl_ok >> (2 * hp)) & 1, vb = (2 * hp + 1 < G::NPL) && ((pl_ok >> (2 * hp + 1)) & 1);
            const uint32_t oa = row_off + (uint32_t)(2 * hp) * zplane_bytes;
            const uint32_t ob = oa + zplane_bytes;
            uint4 v[8];
            if (full_w) {
              if (va) {
#pragma unroll
                for (int it = 0; it < 4; ++it) v[it] = ldg_nc16(vol0 + (oa + it * gstep));
              }
              if (vb) {
#pragma unroll
                for (int it = 0; it < 4; ++it) v[4 + it] = ldg_nc16(vol0 + (ob + it * gstep));
              }
            } else {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                v[it] = make_uint4(0, 0, 0, 0);
                v[4 + it] = make_uint4(0, 0, 0, 0);
                if (va && ((in_w >> it) & 1)) v[it] = ldg_nc16(vol0 + (oa + it * gstep));
                if (vb && ((in_w >> it) & 1)) v[4 + it] = ldg_nc16(vol0 + (ob + it * gstep));
              }
            }
            if (hp == 0) wait_stage_free();
            if (!q_identity) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = zxform8(v[k], m2, s2, t2, l2);
            }
            if (full_w) {
              if (va) {
#pragma unroll
                for (int it = 0; it < 4; ++it) sts16(st + (2 * hp) * plane_bytes + it * 512, v[it]);
              }
              if (vb) {
#pragma unroll
                for (int it = 0; it < 4; ++it) sts16(st + (2 * hp + 1) * plane_bytes + it * 512, v[4 + it]);
              }
            } else {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                if (va && ((in_w >> it) & 1)) sts16(st + (2 * hp) * plane_bytes + it * 512, v[it]);
                if (vb && ((in_w >> it) & 1)) sts16(st + (2 * hp + 1) * plane_bytes + it * 512, v[4 + it]);
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_warp(&full_bar[my_stage]);
        ZPROF_ADD(1, 1);
      }
    }
#ifdef FNNU_ZROWS_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      g_zprof[0] = zp[0]; g_zprof[1] = zp[1]; g_zprof[2] = clock64() - zt0;
    }
#endif
  } else if (warp >= kZMmaWarp0) {
    // =========================== MMA ISSUERS ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kZRegsMma));
    const int me = warp - kZMmaWarp0;
    if (me < c.issuers) {
      constexpr uint32_t idesc_first = (1u << 4) | ((uint32_t)(G::n_of(G::PZ) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_desc0 = make_desc(smem_u32(ring), kZProw * 16, 128);
      const uint64_t b_desc0 = make_desc(smem_u32(w_s), G::BN * 16, 128);
      const uint32_t stage_u16 = (uint32_t)c.stage_bytes >> 4;
      mbar_wait(w_bar, 0);
      int t = 0, turn = 0, stage = 0, slot = 0;
      uint32_t phase = 0, sphase = 0;
#ifdef FNNU_ZROWS_PROF
      long long zp[4] = {0, 0, 0, 0};
      const long long zt0 = clock64();
#endif
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        int b, z0, ya, yb;
        decode(u, b, z0, ya, yb);
        const int n_rows = (yb - ya) + 2;
        const uint32_t pl_ok = planes_ok(z0);
        for (int j = 0; j < n_rows; ++j) {
          if (turn == me) {
            ZPROF_T(w0);
            mbar_wait(&tempty_bar[slot], sphase ^ 1);
            ZPROF_T(w1);
            mbar_wait(&full_bar[stage], phase);
            ZPROF_T(w2);
            ZPROF_ADD(0, w1 - w0);
            ZPROF_ADD(1, w2 - w1);
            tc_fence_after();
            // lane-0 broadcasts tell ptxas the values are warp-uniform (uniform-register operands for UTCHMMA)
            const uint32_t d = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)slot * kZSlotCols, 0);
            const uint32_t da_lo = __shfl_sync(0xffffffffu, (uint32_t)a_desc0 + (uint32_t)stage * stage_u16, 0);
            const uint32_t bar = __shfl_sync(0xffffffffu, smem_u32(&step_bar[t & (kZStepBars - 1)]), 0);
            if (elect_one()) {
              const uint64_t da_st = (a_desc0 & 0xffffffff00000000ull) | da_lo;
              uint32_t accum = 0;
              // plane PZ (output plane z0 itself, always inside the volume) touches every tile of the step: its first
              // MMA is the "accumulate = 0" one
              static_assert(G::n_of(G::PZ) == G::SLOT_COLS && G::d_off(G::PZ) == 0, "the first-touch plane must cover the whole slot");
              z_issue_plane<CHUNKS, G::PZ, G::BN, G::b_col(G::PZ), 0>(d, da_st, b_desc0, idesc_first, accum);
              z_issue_other_planes<CHUNKS, CP, ZC, NKZ, 0, true>(d, da_st, b_desc0, pl_ok, accum);
              z_issue_other_planes<CHUNKS, CP, ZC, NKZ, 0, false>(d, da_st, b_desc0, pl_ok, accum);
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
              progress[me] = t + c.issuers;  // this issuer's next step (steps are dealt round-robin to the issuers)
            }
            __syncwarp();
            ZPROF_ADD(2, clock64() - w2);
            ZPROF_ADD(3, 1);
          }
          ++t;
          if (++turn == c.issuers) turn = 0;
          if (++stage == c.stages) { stage = 0; phase ^= 1; }
          if (++slot == kZSlots) { slot = 0; sphase ^= 1; }
        }
      }
#ifdef FNNU_ZROWS_PROF
      if (blockIdx.x == 0 && lane == 0) {
        long long* o = g_zprof + 8 + me * 8;
        o[0] = zp[0]; o[1] = zp[1]; o[2] = zp[2]; o[3] = zp[3]; o[4] = clock64() - zt0;
      }
#endif
    } else if (me == 2 && c.prefetch_distance > 0) {
      // =========================== L2 PREFETCHER (a warp that would otherwise only donate registers) ===============
      // The producers are bound by the latency of their loads (the previous layer's output comes from HBM: 4.3 GB per
      // launch against 126 MB of L2).  One bulk prefetch per (plane, row), issued `prefetch_distance` y-steps ahead of
      // the tensor pipe, turns those loads into L2 hits.
      const size_t row_bytes = (size_t)c.W * a.src_cs * 2;
      const size_t zplane_bytes = (size_t)c.H * row_bytes;
      const uint32_t pf_bytes = (uint32_t)((size_t)c.W * a.cin * 2) & ~15u;
      const int dist = c.prefetch_distance;
      int t = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        int b, z0, ya, yb;
        decode(u, b, z0, ya, yb);
        const int n_rows = (yb - ya) + 2;
        const int zi = z0 - G::PZ + lane;                   // lanes 0 .. NPL-1: one input plane each
        const bool pl_ok = lane < G::NPL && zi >= 0 && zi < c.D;
        const char* plane = reinterpret_cast<const char*>(a.src) + ((size_t)b * c.D + zi) * zplane_bytes;
        for (int j = 0; j < n_rows; ++j, ++t) {
          // pace on the issuers' progress counters (plain shared-memory words: no barrier phase to miss, and the
          // counters reach their final value whatever happens to this warp)
          while (true) {
            int done = progress[0];
            if (c.issuers == 2) done = min(done, (int)progress[1]);
            if (done + dist >= t) break;
            __nanosleep(256);
          }
          const int y_in = ya - 1 + j;
          if (pl_ok && y_in >= 0 && y_in < c.H)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(plane + (size_t)y_in * row_bytes), "r"(pf_bytes) : "memory");
        }
      }
    }
  } else {
    // ============ EPILOGUE (warp set k: output plane z0 + k when ZC = 2, channels 16 k .. 16 k + 15 when CP = 32) ======
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kZRegsEpilogue));
    const int k = (warp - kZEpilogueWarp0) >> 2;
    if (k < kSets) {
    const int pk = ZC == 2 ? k : 0;                  // output plane of this set
    const int c0 = CP == 32 ? 16 * k : 0;            // first output channel of this set
    const int wq = warp & 3;
    const int x = wq * 32 + lane;                    // output column == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(pk * G::TILE_N + c0);
    const bool vec_store = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0) && c0 + 16 <= a.cout;
    const bool col_ok = x < c.W;
    const bool has_bias = a.bias != nullptr;
    const size_t out_row_stride = (size_t)c.W * a.dst_cs;
#ifdef FNNU_ZROWS_PAIRED_STORE
    const long long partner_delta = (long long)a.dst_cs * 2 * ((lane & 1) ? -1 : 1);     // lanes 2i / 2i+1: columns x, x+1
#endif
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
    int t = 0, slot_a = 0;                           // step / TMEM slot holding input row y-1 of the next output row
#ifdef FNNU_ZROWS_PROF
    long long zp[4] = {0, 0, 0, 0};
    const long long zt0 = clock64();
#endif
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int b, z0, ya, yb;
      decode(u, b, z0, ya, yb);
      if (b != cur_b) {
        flush_stats(cur_b);
        cur_b = b;
      }
      const int n_out = yb - ya;
      // each issuing warp's steps complete in order; consecutive steps may belong to different warps: wait for the
      // first two steps here and for step yo + 2 inside the loop
      mbar_wait(&step_bar[t & (kZStepBars - 1)], (uint32_t)(t >> 4) & 1u);
      mbar_wait(&step_bar[(t + 1) & (kZStepBars - 1)], (uint32_t)((t + 1) >> 4) & 1u);
      __half* out_px = a.dst + (((size_t)b * c.D + z0 + pk) * c.H + ya) * out_row_stride + (size_t)x * a.dst_cs + c0;
      for (int yo = 0; yo < n_out; ++yo, out_px += out_row_stride) {
        const int tc = t + 2;
        ZPROF_T(w0);
        mbar_wait(&step_bar[tc & (kZStepBars - 1)], (uint32_t)(tc >> 4) & 1u);
        ZPROF_ADD(0, clock64() - w0);
        ZPROF_ADD(1, 1);
        tc_fence_after();
        int slot_b = slot_a + 1;
        if (slot_b == kZSlots) slot_b = 0;
        int slot_c = slot_b + 1;
        if (slot_c == kZSlots) slot_c = 0;
        // one TMEM round per row: the three 16-column groups are in flight together
        {
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16_nowait(t_lane + (uint32_t)(slot_a * kZSlotCols + 0), r0);           // T[y-1], ky = 0
          tmem_ld16_nowait(t_lane + (uint32_t)(slot_b * kZSlotCols + CP), r1);          // T[y],   ky = 1
          tmem_ld16_nowait(t_lane + (uint32_t)(slot_c * kZSlotCols + 2 * CP), r2);      // T[y+1], ky = 2
          tmem_wait_ld();
          // the tile of step t is fully consumed (its ky = 1, 2 groups were used by the two previous rows)
          tc_fence_before();
          mbar_arrive_warp(&tempty_bar[slot_a]);
#ifdef FNNU_ZROWS_PAIRED_STORE
          {
            __half2 hv[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float v0 = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]);
              float v1 = (__uint_as_float(r0[j + 1]) + __uint_as_float(r1[j + 1])) + __uint_as_float(r2[j + 1]);
              if (has_bias) {       // only a convolution WITHOUT a following InstanceNorm keeps its bias (program.py)
                const float2 bj = *reinterpret_cast<const float2*>(bias_s + c0 + j);
                v0 += bj.x;
                v1 += bj.y;
              }
              if (col_ok) {
                s1[j] += v0;
                s2[j] = fmaf(v0, v0, s2[j]);
                s1[j + 1] += v1;
                s2[j + 1] = fmaf(v1, v1, s2[j + 1]);
              }
              hv[j >> 1] = __floats2half2_rn(v0, v1);
            }
            if (vec_store) {
              // whole 32-byte sectors per instruction (umma_ptx.cuh: stg32_paired)
              stg32_paired(out_px, partner_delta, *reinterpret_cast<uint4*>(&hv[0]), *reinterpret_cast<uint4*>(&hv[4]), col_ok, lane);
            } else if (col_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < a.cout) out_px[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
            }
          }
#else
          if (col_ok) {
            __half2 hv[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float v0 = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]);
              float v1 = (__uint_as_float(r0[j + 1]) + __uint_as_float(r1[j + 1])) + __uint_as_float(r2[j + 1]);
              if (has_bias) {       // only a convolution WITHOUT a following InstanceNorm keeps its bias (program.py)
                const float2 bj = *reinterpret_cast<const float2*>(bias_s + c0 + j);
                v0 += bj.x;
                v1 += bj.y;
              }
              s1[j] += v0;
              s2[j] = fmaf(v0, v0, s2[j]);
              s1[j + 1] += v1;
              s2[j + 1] = fmaf(v1, v1, s2[j + 1]);
              hv[j >> 1] = __floats2half2_rn(v0, v1);
            }
            if (vec_store) {
              reinterpret_cast<uint4*>(out_px)[0] = *reinterpret_cast<uint4*>(&hv[0]);
              reinterpret_cast<uint4*>(out_px)[1] = *reinterpret_cast<uint4*>(&hv[4]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < a.cout) out_px[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
            }
          }
#endif
        }
        ++t;
        slot_a = slot_b;
      }
      // the unit's last two steps have no later consumer
      tc_fence_before();
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        mbar_arrive_warp(&tempty_bar[slot_a]);
        if (++slot_a == kZSlots) slot_a = 0;
        ++t;
      }
    }
    flush_stats(cur_b);
#ifdef FNNU_ZROWS_PROF
    if (blockIdx.x == 0 && lane == 0 && wq == 0) {
      long long* o = g_zprof + 24 + k * 4;
      o[0] = zp[0]; o[1] = zp[1]; o[2] = clock64() - zt0;
    }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kZMmaWarp0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// weights [cout][cin][kz][ky][kx] fp32 -> fp16 [kx][chunk][half][n = (nkz - 1 - kz) * 3 cp + ky * cp + co][8]
__global__ void pack_weights_zrows_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout, int cp,
                                          int nkz) {
  const int chunks = cin / 16;
  const int bn = nkz * 3 * cp;
  const size_t total = (size_t)3 * chunks * 2 * bn * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % bn); r /= bn;
    const int h = (int)(r % 2); r /= 2;
    const int kc = (int)(r % chunks); r /= chunks;
    const int kx = (int)r;
    const int kz = nkz - 1 - n / (3 * cp), ky = (n % (3 * cp)) / cp, co = n % cp;
    const int ci = kc * 16 + h * 8 + e;
    float v = 0.f;
    if (co < cout) v = w[((((size_t)co * cin + ci) * nkz + kz) * 3 + ky) * 3 + kx];
    out[i] = __float2half_rn(v);
  }
}

}  // namespace

bool zrows_supported(const ConvArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FNNU_ZROWS");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled) return false;
  ZCfg c;
  if (!plan_zrows(a, c)) return false;
  // FNNU_ZROWS=2: only the z-pair form (the older ky-folded kernel keeps the other shapes) — A/B measurements
  return enabled == 1 || c.zc == 2;
}

int launch_pack_weights_zrows(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s) {
  ZCfg c;
  if (!plan_zrows(a, c)) return FNNU_E_UNSUPPORTED;
  pack_weights_zrows_kernel<<<64, 256, 0, s>>>(w_dev, (__half*)out, a.cin, a.cout, c.cp, c.nkz);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

#ifdef FNNU_ZROWS_PROF
extern "C" int fnnu_debug_zrows_prof(long long* out32) {
  return cudaMemcpyFromSymbol(out32, g_zprof, sizeof(long long) * 32) == cudaSuccess ? 0 : -2;
}
#endif

int launch_conv_zrows(const ConvArgs& a, cudaStream_t s) {
  ZArgs p;
  p.a = a;
  if (!plan_zrows(a, p.c)) {
    set_error("conv_umma_zrows: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  p.n_units = a.batch * (p.c.D / p.c.zc) * p.c.n_yseg;
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  bool launched = false;
#define FNNU_Z_CASE(CH, CPV, ZCV, NKZV)                                                                                       \
  if (!launched && p.c.chunks == CH && p.c.cp == CPV && p.c.zc == ZCV && p.c.nkz == NKZV) {                                  \
    /* the attribute is per device: set it on every launch (cheap) */                                                         \
    FNNU_CUDA(cudaFuncSetAttribute(conv_umma_zrows_kernel<CH, CPV, ZCV, NKZV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                   kZSmemLimit));                                                                             \
    conv_umma_zrows_kernel<CH, CPV, ZCV, NKZV><<<grid, kZThreads, p.c.smem_bytes, s>>>(p);                                    \
    launched = true;                                                                                                          \
  }
  FNNU_Z_CASE(1, 16, 2, 3) FNNU_Z_CASE(2, 16, 2, 3)
  FNNU_Z_CASE(1, 16, 1, 3) FNNU_Z_CASE(2, 16, 1, 3) FNNU_Z_CASE(1, 16, 1, 1) FNNU_Z_CASE(2, 16, 1, 1)
  FNNU_Z_CASE(1, 32, 1, 3) FNNU_Z_CASE(2, 32, 1, 3) FNNU_Z_CASE(1, 32, 1, 1) FNNU_Z_CASE(2, 32, 1, 1)
#undef FNNU_Z_CASE
  if (!launched) {
    set_error("conv_umma_zrows: no instantiation for chunks=%d cp=%d zc=%d nkz=%d", p.c.chunks, p.c.cp, p.c.zc, p.c.nkz);
    return FNNU_E_UNSUPPORTED;
  }
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
