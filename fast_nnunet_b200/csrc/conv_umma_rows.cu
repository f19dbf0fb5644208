// Row-streaming tcgen05 convolution for the thin, full-resolution layers (Cout <= 32, W <= 128 ...):
// the layers that hold most of a nnU-Net's FLOPs (48 % at 128^3 with Cout = 16).
//
// Why a second kernel.  With A and B read from shared memory a 128 x N x 16 tcgen05.mma takes 39 / 44 / 56 / 64 /
// 128 cycles for N = 16 / 48 / 96 / 128 / 256 (tests/cuda/bench_umma_rate.cu): below N = 128 it is bound by the
// operand fetch (4 KB of A + 32 N bytes of B at ~125 B/cycle), its math takes only N/2 cycles.  A 3x3x3 conv with
// Cout = 16 issued as 27 taps x (N = 16) is therefore operand-fetch bound at <= 12 % of the tensor peak.  Here the ky taps
// are FOLDED INTO N:   D'[row r][x, (ky, co)] = sum_{kz, kx, ci} X[z + kz - 1, r, x + kx - 1, ci] * W[kz, ky, kx][ci, co]
// is one accumulator tile per INPUT row r (N = 3 * Cout, 9 MMAs per 16 input channels instead of 27), and
//   out[y][x, co] = D'[y - 1][x, (2, co)] + D'[y][x, (1, co)] + D'[y + 1][x, (0, co)]
// is a sum of three accumulator tiles AT THE SAME LANE, done by the epilogue while it drains TMEM — no
// cross-lane traffic.  Because an input row now feeds exactly one accumulator tile, the kernel streams:
//   producers (16 warps in 8 or 4 groups): one smem stage = one input row x 3 z-planes x Cin, normalise-on-load
//   MMA warp: 9 * Cin/16 MMAs per stage into a ring of TMEM tile slots, commit per row
//   epilogue (4 warps): output row y as soon as tile y+1 is complete; frees tile y-1
// Weights (all 27 taps) stay resident in shared memory for the whole kernel (<= 64 KB).
#include "common.cuh"
#include "ops.cuh"
#include "umma_ptx.cuh"

namespace fnnu {

struct RowsCfg {
  int ok;
  int nkz, nkx, pz, px;
  int Nf;                 // 3 * cout_pad: N of every MMA
  int chunks;             // Cin / 16
  int Q;                  // Cin / 8 (8-channel groups)
  int W, H, D;
  int P_row;              // positions allocated per (kz, group) row in smem
  int stage_bytes, stages;
  int w_bytes;
  int slots;              // TMEM tile slots (Nf columns each)
  int n_yseg, seg_rows;
  int smem_bytes;
  int occ;                // CTAs per SM this launch is planned for (1 or 2): halves smem / TMEM budgets
  int tmem_cols;
};

struct RowsArgs {
  ConvArgs a;
  RowsCfg c;
  int n_units;
};

// 16 producer warps (the producers are latency-bound: ncu shows them busy ~85 % at one instruction per ~8 cycles
// per warp), dealt in groups of 64 * CHUNKS threads (8 groups for Cin 16, 4 groups for Cin 32) so that a thread
// always covers a row in 5 steps of 32 positions.  Warps 16-19: epilogue, warp 20: MMA (one thread chosen with
// elect.sync issues; the other lanes only take part in the mbarrier waits).
constexpr int kRowsProducerWarps = 16;
constexpr int kRowsProducerThreads = kRowsProducerWarps * 32;
constexpr int kRowsMmaWarp = kRowsProducerWarps + 4;
constexpr int kRowsMmaIssuers = 1;      // issuing warps (rows dealt round-robin); 2 measured no faster: the tensor pipe, not the issue path, paces the MMAs
constexpr int kRowsThreads = 24 * 32;   // 6 warpgroups: 4 producer, 1 epilogue, 1 holding the MMA warp (+3 idle warps)
constexpr int kRegsProducer = 72, kRegsMma = 40, kRegsEpilogue = 152;   // setmaxnreg: 512 x 72 + 128 x 40 + 128 x 152 = 768 x 80
__host__ __device__ constexpr int rows_group_threads(int chunks) { return 64 * chunks; }
__host__ __device__ constexpr int rows_groups(int chunks) { return kRowsProducerThreads / (64 * chunks); }
constexpr int kRowsMaxStages = 16;
constexpr int kRowsMaxSlots = 10;
constexpr int kRowsSmemLimit = 227 * 1024;

static bool plan_rows(const ConvArgs& a, RowsCfg& c) {
  memset(&c, 0, sizeof(c));
  if (a.transposed) return false;
  if (a.s[0] != 1 || a.s[1] != 1 || a.s[2] != 1) return false;
  if (a.k[1] != 3 || a.k[2] != 3) return false;          // ky is the folded axis; kx == 3 fixes the row pitch (140)
  if (a.cin != 16 && a.cin != 32) return false;          // producer mapping: 64 threads cover a row in <= 9 steps
  if (a.cout_pad != 16 && a.cout_pad != 32) return false;   // per-thread InstanceNorm partials live in registers
  if (a.src_cs % 8 != 0 || ((uintptr_t)a.src % 16) != 0) return false;
  c.D = a.in_d[0]; c.H = a.in_d[1]; c.W = a.in_d[2];
  if (c.W > 128 || c.W < 64) return false;               // one M-tile per row; below 64 the generic kernel packs better
  c.nkz = a.k[0]; c.nkx = a.k[2]; c.pz = a.pad[0]; c.px = a.pad[2];
  c.Nf = 3 * a.cout_pad;
  c.chunks = a.cin / 16;
  c.Q = a.cin / 8;
  // positions read by an MMA tile: [kx, kx + 128); keep the row stride = 4 (mod 8) positions for the banks
  int need = 128 + (c.nkx - 1);
  if (need < c.W + 2 * c.px) need = c.W + 2 * c.px;
  c.P_row = (need + 7) / 8 * 8 + 4;
  c.stage_bytes = c.nkz * c.Q * c.P_row * 16;
  c.w_bytes = c.nkz * c.nkx * c.chunks * 2 * c.Nf * 16;
  if (c.w_bytes > 80 * 1024) return false;
  const int misc = 1024 + 3 * a.cin * 4 + 1024;
  // One CTA per SM.  Two (256 TMEM columns and ~113 KB of shared memory each) were measured SLOWER (enc0.1: 2.8 ->
  // 4.0 ms per 32 patches): the kernel is bound by shared-memory traffic (A is re-read for each of the 9 taps), not
  // by latency.  The OCC template parameter is kept for that experiment.
  c.occ = 1;
  const int st = (kRowsSmemLimit - misc - c.w_bytes) / c.stage_bytes;
  // stages > producer groups is required: a group publishes row r only while it prefetches row r + groups
  if (st <= rows_groups(c.chunks)) return false;
  c.stages = st > kRowsMaxStages ? kRowsMaxStages : st;
  c.tmem_cols = c.occ == 2 ? 256 : 512;
  c.slots = c.tmem_cols / c.Nf;
  if (c.slots > kRowsMaxSlots) c.slots = kRowsMaxSlots;
  if (c.slots < 4) return false;
  c.smem_bytes = c.w_bytes + c.stages * c.stage_bytes + misc;
  // units: (sample, z, y segment).  Whole columns unless that leaves SMs idle.
  c.n_yseg = 1;
  while ((long long)a.batch * c.D * c.n_yseg < 3LL * num_sms() && c.H / (c.n_yseg * 2) >= 8) c.n_yseg *= 2;
  c.seg_rows = (c.H + c.n_yseg - 1) / c.n_yseg;
  c.ok = 1;
  return true;
}

// The 3 * CHUNKS MMAs of one kz plane of a row; descriptor offsets (16-byte units) are template constants.
template <int CP, int CHUNKS, int KZ, int I>   // I -> (kx = I / CHUNKS, kc = I % CHUNKS)
__device__ __forceinline__ void rows_issue_plane(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t& accum) {
  if constexpr (I < 3 * CHUNKS) {
    constexpr int kx = I / CHUNKS, kc = I % CHUNKS;
    constexpr uint32_t a_off = (uint32_t)((KZ * 2 * CHUNKS + kc * 2) * 140 + kx);       // plan_rows: P_row == 140
    constexpr uint32_t b_off = (uint32_t)((((KZ * 3 + kx) * CHUNKS + kc) * 2) * (3 * CP));
    umma_f16_off<a_off, b_off>(d, da, db, idesc, accum);
    accum = 1;
    rows_issue_plane<CP, CHUNKS, KZ, I + 1>(d, da, db, idesc, accum);
  }
}

template <int CP, int CHUNKS, int OCC>   // cout_pad (16 | 32); CHUNKS = Cin / 16; OCC = CTAs per SM
__global__ void __launch_bounds__(kRowsThreads, OCC) conv_umma_rows_kernel(const __grid_constant__ RowsArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const RowsCfg& c = p.c;
  const ConvArgs& a = p.a;
  uint8_t* w_s = smem;
  uint8_t* ring = smem + c.w_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)c.stages * c.stage_bytes);   // [16]
  uint64_t* empty_bar = full_bar + kRowsMaxStages;                                              // [16]
  uint64_t* tfull_bar = empty_bar + kRowsMaxStages;                                             // [10]
  uint64_t* tempty_bar = tfull_bar + kRowsMaxSlots;                                             // [10]
  uint64_t* w_bar = tempty_bar + kRowsMaxSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* xs = reinterpret_cast<float*>(tmem_slot + 4);
  float* xh = xs + a.cin;
  float* xl = xh + a.cin;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full_bar[s], rows_group_threads(CHUNKS) / 32);   // one arrival per producer warp of the group
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < c.slots; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);                                // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kRowsMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)c.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {   // resident weights: one bulk copy per CTA
    mbar_arrive_expect_tx(w_bar, (uint32_t)c.w_bytes);     // the barrier's single arrival + the byte count
    bulk_g2s(w_s, a.w_umma, (uint32_t)c.w_bytes, w_bar);
  }

  // unit u -> (sample b, plane z, y segment); rows [ya, yb)
  auto decode = [&](int u, int& b, int& z, int& ya, int& yb) {
    z = u % c.D;
    u /= c.D;
    const int seg = u % c.n_yseg;
    b = u / c.n_yseg;
    ya = seg * c.seg_rows;
    yb = ya + c.seg_rows;
    if (yb > c.H) yb = c.H;
  };

  if (warp < kRowsProducerWarps) {
    // =========================== PRODUCERS ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
    const int tid = threadIdx.x;
    constexpr int kGroupThreads = rows_group_threads(CHUNKS);
    constexpr int kGroups = rows_groups(CHUNKS);
    const int grp = tid / kGroupThreads;              // stages are dealt round-robin to the groups
    const int gt = tid - grp * kGroupThreads;
    // A thread owns ONE 8-channel group q (its scale/shift/slope live in registers) and every XP_STEP-th
    // position of a row; consecutive lanes fetch the consecutive 16-byte pieces of whole voxels.  Everything
    // that does not depend on the row (offsets, in-image flags) is computed once, so the per-row loop is
    // one cp.async per item and, one stage later, LDS.128 + 12 half2 ops + STS.128 per item.
    constexpr int kQ_ = 2 * CHUNKS;
    constexpr int XP_STEP = kGroupThreads / kQ_;              // 32 positions per step
    constexpr int N_IT = (140 + XP_STEP - 1) / XP_STEP;       // 5
    // A warp covers 16 positions x one PAIR of channel groups (32 contiguous bytes per voxel: full sectors), so
    // whether a warp's channels carry a pending transform is warp-uniform: the up-sampled half of a decoder concat
    // (ConvTranspose3d output, no norm) skips the in-place normalise pass altogether.
    const int wg = gt >> 5;                                   // warp inside the group (1 or 2 q-pairs x 2 halves)
    const int q = 2 * (wg & (CHUNKS - 1)) + (lane & 1);
    const int xp0 = (wg / CHUNKS) * 16 + (lane >> 1);
    bool q_identity = true;                                   // all 16 channels of the warp's q-pair: eps < 0
#pragma unroll
    for (int e = 0; e < 16; ++e) q_identity = q_identity && (a.src_meta[(q & ~1) * 8 + e].eps < 0.f);
    uint32_t goff[N_IT];                                      // byte offset of the voxel within its row
    uint32_t live = 0, inimg = 0;                             // bit it: position exists / lies inside the image
#pragma unroll
    for (int it = 0; it < N_IT; ++it) {
      const int xp = xp0 + it * XP_STEP;
      const int x_in = xp - c.px;
      const bool in = x_in >= 0 && x_in < c.W;
      if (xp < c.P_row) live |= 1u << it;
      if (in) inimg |= 1u << it;
      goff[it] = (uint32_t)((in ? x_in : 0) * a.src_cs + q * 8) * 2u;
    }
    const uint32_t plane_bytes = (uint32_t)(c.Q * c.P_row) * 16u;
    const uint32_t my_off = (uint32_t)(q * c.P_row + xp0) * 16u;   // this thread's first item inside a plane
    const size_t row_bytes = (size_t)c.W * a.src_cs * 2;
    // ring position of the CTA's next row: group index, stage and phase are advanced incrementally (a 64-bit
    // div/mod by a run-time value is a ~300-cycle subroutine; five of them per row bounded the kernel once)
    int row_grp = 0, row_stage = 0;
    uint32_t row_phase = 0;
    int cur_b = -1;
    __half2 s2[4], t2[4], l2[4];
    // Software pipeline, two stages deep per group: the raw row is fetched with cp.async (zero-filled outside the
    // image, no registers held, latency overlapped with the previous stage's transform), then normalised in place.
    int pend_stage = -1, pend_mask = 0;               // stage whose copies are in flight; bit kz = plane present
    bool pend_row_ok = false;
    auto finish_pending = [&](int keep_in_flight) {
      if (pend_stage < 0) return;
      if (keep_in_flight) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
      if (pend_row_ok && !q_identity) {
        uint8_t* st = ring + (size_t)pend_stage * c.stage_bytes + my_off;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
          if (!((pend_mask >> kz) & 1)) continue;
#pragma unroll
          for (int it = 0; it < N_IT; ++it) {
            if ((inimg >> it) & 1) {
              uint4* ptr = reinterpret_cast<uint4*>(st + kz * plane_bytes + it * (XP_STEP * 16));
              *ptr = xform8_h2(*ptr, s2, t2, l2);
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_warp(&full_bar[pend_stage]);
      pend_stage = -1;
    };
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int b, z, ya, yb;
      decode(u, b, z, ya, yb);
      if (b != cur_b) {
        finish_pending(0);                            // rows of the previous sample use the previous transform
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float sc[2], sh[2], sl[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = q * 8 + 2 * e + k;
            const ChanMeta m = a.src_meta[ch];
            xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + ch) * 2, m, a.src_inv_count, sc[k], sh[k]);
            sl[k] = m.eps < 0.f ? 1.f : m.slope;
          }
          s2[e] = __floats2half2_rn(sc[0], sc[1]);
          t2[e] = __floats2half2_rn(sh[0], sh[1]);
          l2[e] = __floats2half2_rn(sl[0], sl[1]);
        }
        cur_b = b;
      }
      int zmask = 0;
      for (int kz = 0; kz < c.nkz; ++kz)
        if (z + kz - c.pz >= 0 && z + kz - c.pz < c.D) zmask |= 1 << kz;
      const int n_rows = (yb - ya) + 2;
      const char* plane0 = reinterpret_cast<const char*>(a.src) + (((size_t)b * c.D + (z - c.pz)) * c.H) * row_bytes;
      for (int j = 0; j < n_rows; ++j) {
        const int stage = row_stage;
        const uint32_t phase = row_phase;
        const bool mine = row_grp == grp;
        row_grp = (row_grp + 1) & (kGroups - 1);
        if (++row_stage == c.stages) { row_stage = 0; row_phase ^= 1; }
        if (!mine) continue;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int y_in = ya - 1 + j;
        const bool row_ok = y_in >= 0 && y_in < c.H;
        const uint32_t fill = row_ok ? inimg : 0u;    // items that carry data; the rest are zero-filled
        const uint32_t dst0 = smem_u32(ring + (size_t)stage * c.stage_bytes) + my_off;
        const char* row0 = plane0 + (size_t)(row_ok ? y_in : 0) * row_bytes;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
          if (!((zmask >> kz) & 1)) continue;         // the MMA warp skips this plane too
          const char* row = row0 + (size_t)kz * c.H * row_bytes;
#pragma unroll
          for (int it = 0; it < N_IT; ++it) {
            if ((live >> it) & 1)
              cp_async16_zfill(dst0 + kz * plane_bytes + it * (XP_STEP * 16), row + goff[it], ((fill >> it) & 1) ? 16u : 0u);
          }
        }
        cp_async_commit_group();
        finish_pending(1);
        pend_stage = stage;
        pend_mask = zmask;
        pend_row_ok = row_ok;
      }
    }
    finish_pending(0);
  } else if (warp >= kRowsMmaWarp) {
    // =========================== MMA ISSUER (warp 20; warps 21-23 only donate registers) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMma));
    if (warp < kRowsMmaWarp + kRowsMmaIssuers) {
    // Rows are dealt round-robin to kRowsMmaIssuers issuing warps (each row has its own TMEM tile, so the rows are
    // independent): one warp's issue path (R2UR + ELECT + UTCHMMA, ~65 cycles per MMA) is slower than the tensor pipe.
    // One elected lane issues; every descriptor offset of a row's 9 * CHUNKS MMAs is a compile-time constant
    // added to the stage base, so the issue path is ~a dozen instructions per MMA (the MMA itself occupies the
    // tensor pipe for ~64 cycles of operand fetch).
    constexpr uint32_t kProw = 140;                      // plan_rows: nkx == 3 -> P_row == 140
    constexpr uint32_t kNf = 3 * CP;
    constexpr uint32_t kQ = 2 * CHUNKS;
    const uint32_t idesc = (1u << 4) | ((kNf >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(ring), kProw * 16, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(w_s), kNf * 16, 128);
    const uint32_t stage_u16 = (uint32_t)c.stage_bytes >> 4;
    mbar_wait(w_bar, 0);
    int stage = 0, slot = 0, turn = 0;
    const int my_turn = warp - kRowsMmaWarp;
    uint32_t phase = 0, sphase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int b, z, ya, yb;
      decode(u, b, z, ya, yb);
      const int n_rows = (yb - ya) + 2;
      bool kzv[3];
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) kzv[kz] = kz < c.nkz && (z + kz - c.pz) >= 0 && (z + kz - c.pz) < c.D;
      for (int j = 0; j < n_rows; ++j) {
        const bool mine = turn == my_turn;
        if (++turn == kRowsMmaIssuers) turn = 0;
        if (!mine) {
          if (++stage == c.stages) { stage = 0; phase ^= 1; }
          if (++slot == c.slots) { slot = 0; sphase ^= 1; }
          continue;
        }
        mbar_wait(&tempty_bar[slot], sphase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // the two per-row values go through a lane-0 broadcast: ptxas then knows they are warp-uniform and feeds the
        // UTCHMMA uniform-register operands directly instead of through a per-instruction R2UR "waterfall" loop
        const uint32_t d = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)slot * kNf, 0);
        const uint32_t da_lo = __shfl_sync(0xffffffffu, (uint32_t)a_desc0 + (uint32_t)stage * stage_u16, 0);
        if (elect_one()) {
          const uint64_t da_st = (a_desc0 & 0xffffffff00000000ull) | da_lo;
          uint32_t accum = 0;
          if (kzv[0]) rows_issue_plane<CP, CHUNKS, 0, 0>(d, da_st, b_desc0, idesc, accum);
          if (kzv[1]) rows_issue_plane<CP, CHUNKS, 1, 0>(d, da_st, b_desc0, idesc, accum);
          if (kzv[2]) rows_issue_plane<CP, CHUNKS, 2, 0>(d, da_st, b_desc0, idesc, accum);
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[slot]);
        }
        __syncwarp();
        if (++stage == c.stages) { stage = 0; phase ^= 1; }
        if (++slot == c.slots) { slot = 0; sphase ^= 1; }
      }
    }
    }
  } else {
    // =========================== EPILOGUE ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int wq = warp & 3;
    const int x = wq * 32 + lane;                   // output column == TMEM lane
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const bool vec_store = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0);
    int slot_a = 0;                                 // TMEM slot (and its phase) of the unit tile that holds row y-1
    uint32_t phase_a = 0;
    int cur_b = -1;
    constexpr int kBiasRegs = CP == 16 ? 16 : 1;      // CP == 32: registers are needed for the partial sums
    float s1[CP], s2[CP], bias[kBiasRegs];
#pragma unroll
    for (int j = 0; j < CP; ++j) s1[j] = s2[j] = 0.f;
#pragma unroll
    for (int j = 0; j < kBiasRegs; ++j) bias[j] = (a.bias && j < a.cout) ? __ldg(a.bias + j) : 0.f;
    const bool col_ok = x < c.W;
    const size_t out_row_stride = (size_t)c.W * a.dst_cs;
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
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int b, z, ya, yb;
      decode(u, b, z, ya, yb);
      if (b != cur_b) {
        flush_stats(cur_b);
        cur_b = b;
      }
      const int n_out = yb - ya;
      {
        // tiles complete in issue order per issuing warp only: wait for every tile once (rows y-1 and y here, row
        // y+1 inside the loop)
        int sl = slot_a;
        uint32_t ph = phase_a;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          mbar_wait(&tfull_bar[sl], ph);
          if (++sl == c.slots) { sl = 0; ph ^= 1; }
        }
      }
      __half* out_px = a.dst + (((size_t)b * c.D + z) * c.H + ya) * out_row_stride + (size_t)x * a.dst_cs;
      for (int yo = 0; yo < n_out; ++yo, out_px += out_row_stride) {
        // unit tiles yo, yo+1, yo+2 hold input rows y-1, y, y+1; MMAs complete in order: wait for the last
        int slot_b = slot_a + 1;
        uint32_t phase_c = phase_a;
        if (slot_b == c.slots) { slot_b = 0; phase_c ^= 1; }
        int slot_c = slot_b + 1;
        if (slot_c == c.slots) { slot_c = 0; phase_c ^= 1; }
        mbar_wait(&tfull_bar[slot_c], phase_c);
        tc_fence_after();
        const uint32_t t_a = tmem_base + lane_off + (uint32_t)(slot_a * (int)(3 * CP));
        const uint32_t t_b = tmem_base + lane_off + (uint32_t)(slot_b * (int)(3 * CP));
        const uint32_t t_c = tmem_base + lane_off + (uint32_t)(slot_c * (int)(3 * CP));
#pragma unroll
        for (int g0 = 0; g0 < CP; g0 += 16) {
          // out(y) = D'[y-1][ky = 0] + D'[y][ky = 1] + D'[y+1][ky = 2]   (input row r feeds output row r - ky + 1)
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16_nowait(t_a + (uint32_t)(0 * CP + g0), r0);
          tmem_ld16_nowait(t_b + (uint32_t)(1 * CP + g0), r1);
          tmem_ld16_nowait(t_c + (uint32_t)(2 * CP + g0), r2);
          tmem_wait_ld();
          if (col_ok) {
            __half2 hv[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float bj0, bj1;
              if constexpr (CP == 16) {
                bj0 = bias[j];
                bj1 = bias[j + 1];
              } else {
                bj0 = (a.bias && g0 + j < a.cout) ? __ldg(a.bias + g0 + j) : 0.f;
                bj1 = (a.bias && g0 + j + 1 < a.cout) ? __ldg(a.bias + g0 + j + 1) : 0.f;
              }
              const float v0 = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]) + bj0;
              const float v1 = (__uint_as_float(r0[j + 1]) + __uint_as_float(r1[j + 1])) + __uint_as_float(r2[j + 1]) + bj1;
              // InstanceNorm partial sums on the fp32 values (the fp16 rounding error averages out over >= 1e5 voxels)
              s1[g0 + j] += v0;
              s2[g0 + j] = fmaf(v0, v0, s2[g0 + j]);
              s1[g0 + j + 1] += v1;
              s2[g0 + j + 1] = fmaf(v1, v1, s2[g0 + j + 1]);
              hv[j >> 1] = __floats2half2_rn(v0, v1);
            }
            __half* q = out_px + g0;
            if (vec_store && g0 + 16 <= a.cout) {
              reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
              reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[4]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (g0 + j < a.cout) q[j] = (j & 1) ? __high2half(hv[j >> 1]) : __low2half(hv[j >> 1]);
            }
          }
        }
        // unit tile yo is fully consumed (its ky = 1, 2 groups were used by the two previous rows)
        tc_fence_before();
        mbar_arrive_warp(&tempty_bar[slot_a]);
        if (++slot_a == c.slots) { slot_a = 0; phase_a ^= 1; }
      }
      // the last two tiles of the unit have no later consumer
      tc_fence_before();
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        mbar_arrive_warp(&tempty_bar[slot_a]);
        if (++slot_a == c.slots) { slot_a = 0; phase_a ^= 1; }
      }
    }
    flush_stats(cur_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRowsMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)c.tmem_cols));
  }
}

// weights [cout][cin][kz][ky][kx] fp32 -> fp16 [kz][kx][chunk][half][n = ky * cout_pad + co][8]
__global__ void pack_weights_rows_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout,
                                         int cout_pad, RowsCfg c) {
  const size_t total = (size_t)c.nkz * c.nkx * c.chunks * 2 * c.Nf * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % c.Nf); r /= c.Nf;
    const int h = (int)(r % 2); r /= 2;
    const int kc = (int)(r % c.chunks); r /= c.chunks;
    const int kx = (int)(r % c.nkx); r /= c.nkx;
    const int kz = (int)r;
    const int ky = n / cout_pad, co = n - ky * cout_pad;
    const int ci = kc * 16 + h * 8 + e;
    float v = 0.f;
    if (co < cout) v = w[((((size_t)co * cin + ci) * c.nkz + kz) * 3 + ky) * c.nkx + kx];
    out[i] = __float2half_rn(v);
  }
}

bool rows_supported(const ConvArgs& a) {
  RowsCfg c;
  return plan_rows(a, c);
}

int launch_pack_weights_rows(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s) {
  RowsCfg c;
  if (!plan_rows(a, c)) return FNNU_E_UNSUPPORTED;
  const size_t total = (size_t)c.w_bytes / 2;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weights_rows_kernel<<<blocks, 256, 0, s>>>(w_dev, (__half*)out, a.cin, a.cout, a.cout_pad, c);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

int launch_conv_rows(const ConvArgs& a, cudaStream_t s) {
  RowsArgs p;
  p.a = a;
  if (!plan_rows(a, p.c)) {
    set_error("conv_umma_rows: unsupported shape");
    return FNNU_E_UNSUPPORTED;
  }
  p.n_units = a.batch * p.c.D * p.c.n_yseg;
  int grid = p.n_units < num_sms() * p.c.occ ? p.n_units : num_sms() * p.c.occ;
#define FNNU_ROWS_CASE(CPV, CHV, OCCV)                                                                                    \
  if (a.cout_pad == CPV && p.c.chunks == CHV && p.c.occ == OCCV) {                                                                 \
    /* the attribute is per device: set it on every launch (cheap) */ FNNU_CUDA(cudaFuncSetAttribute(conv_umma_rows_kernel<CPV, CHV, OCCV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                     kRowsSmemLimit));                                                                                                                \
    conv_umma_rows_kernel<CPV, CHV, OCCV><<<grid, kRowsThreads, p.c.smem_bytes, s>>>(p);                                   \
  }
  FNNU_ROWS_CASE(16, 1, 1)
  else FNNU_ROWS_CASE(16, 2, 1)
  else FNNU_ROWS_CASE(32, 1, 1)
  else FNNU_ROWS_CASE(32, 2, 1)
  else {
    set_error("conv_umma_rows: no instantiation for cout_pad=%d chunks=%d", a.cout_pad, p.c.chunks);
    return FNNU_E_UNSUPPORTED;
  }
#undef FNNU_ROWS_CASE
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
