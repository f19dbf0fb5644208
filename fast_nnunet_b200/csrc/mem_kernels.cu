// Memory-bound operators of the sliding window: tile gather (+mirror), TTA-mean + Gaussian weight +
// accumulate, weight-sum map, normalise + inf check + argmax, halo add.
// Reference semantics: distillation/nnunetv2/inference/predict_from_raw_data.py:541-631.
// All kernels are HBM-bound; they use 128-bit accesses where the innermost start is 16-byte aligned
// and grids sized in multiples of the SM count.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace fnnu {

static int64_t g_mem_launches = 0;   // kernels launched by the memory-bound entry points (bench evidence)

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// ------------------------------------------------------------------------------------------------
// gather: volume fp32 [C][X][Y][Z] -> tiles fp16 [n_tiles*n_flips][pX][pY][pZ][cs]
// ------------------------------------------------------------------------------------------------
struct FlipList {
  uint8_t m[8];
};

// grid: x = chunks of 256 over (y, z) of one tile plane, y = tile plane x, z = sample (tile * n_flips + flip):
// no 64-bit division on the path (the index arithmetic used to cost more than the three memory operations).
__global__ void __launch_bounds__(256) gather_tiles_kernel(
    const float* __restrict__ vol, int C, int X, int Y, int Z, const int32_t* __restrict__ starts,
    int n_tiles, int pX, int pY, int pZ, FlipList flips, int n_flips, __half* __restrict__ out, int cs) {
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= (uint32_t)(pY * pZ)) return;
  const int y = (int)(idx / (uint32_t)pZ);
  const int z = (int)(idx - (uint32_t)y * (uint32_t)pZ);
  const int x = blockIdx.y;
  const int n = blockIdx.z;
  const int t = n / n_flips;
  const int f = flips.m[n - t * n_flips];
  const int sx = starts[t * 3 + 0], sy = starts[t * 3 + 1], sz = starts[t * 3 + 2];
  const int gx = sx + ((f & 1) ? pX - 1 - x : x);
  const int gy = sy + ((f & 2) ? pY - 1 - y : y);
  const int gz = sz + ((f & 4) ? pZ - 1 - z : z);
  const size_t vstride = (size_t)X * Y * Z;
  const float* src = vol + ((size_t)gx * Y + gy) * Z + gz;
  __half* dst = out + (((size_t)n * pX + x) * pY * pZ + idx) * cs;
  if (cs == 4 && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) {
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = c < C ? __ldg(src + c * vstride) : 0.f;
    __half2 h[2] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3])};
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(h);
    return;
  }
  for (int c = 0; c < C; ++c) dst[c] = __float2half_rn(__ldg(src + c * vstride));
  for (int c = C; c < cs; ++c) dst[c] = __float2half_rn(0.f);
}

// Single-channel volumes (CT): one thread reads 8 consecutive z voxels of the tile ONCE (two 128-bit loads) and writes
// every mirrored copy (one 16-byte store per copy; z-mirrored copies get the 8 values in reverse order).
// Requires pZ % 8 == 0 and cs == 1.  grid: x = chunks of 256 over (y, z / 8), y = tile plane x, z = tile.
__global__ void __launch_bounds__(256) gather_tiles_c1_vec8_kernel(
    const float* __restrict__ vol, int X, int Y, int Z, const int32_t* __restrict__ starts, int n_tiles, int pX,
    int pY, int pZ, FlipList flips, int n_flips, __half* __restrict__ out) {
  const int zq = pZ >> 3;
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= (uint32_t)(pY * zq)) return;
  const int y = (int)(idx / (uint32_t)zq);
  const int z = (int)(idx - (uint32_t)y * (uint32_t)zq) << 3;
  const int x = blockIdx.y;
  const int t = blockIdx.z;
  const int sx = starts[t * 3 + 0], sy = starts[t * 3 + 1], sz = starts[t * 3 + 2];
  const float* src = vol + ((size_t)(sx + x) * Y + (sy + y)) * Z + (sz + z);
  float v[8];
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(src + k);
  }
  __half2 fwd[4], rev[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    fwd[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    rev[k] = __floats2half2_rn(v[7 - 2 * k], v[6 - 2 * k]);
  }
  const size_t pvox = (size_t)pX * pY * pZ;
  __half* out_t = out + (size_t)t * n_flips * pvox;
#pragma unroll
  for (int fi = 0; fi < 8; ++fi) {
    if (fi < n_flips) {
      const int f = flips.m[fi];
      const int dx = (f & 1) ? pX - 1 - x : x;
      const int dy = (f & 2) ? pY - 1 - y : y;
      const bool rz = (f & 4) != 0;
      const int dz = rz ? pZ - 8 - z : z;
      *reinterpret_cast<uint4*>(out_t + (size_t)fi * pvox + ((size_t)dx * pY + dy) * pZ + dz) =
          rz ? *reinterpret_cast<const uint4*>(rev) : *reinterpret_cast<const uint4*>(fwd);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// accumulate: one launch per tile (tiles of a batch overlap, so they are applied in order).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename AccT>
__device__ __forceinline__ void acc_add(AccT* p, float v);
template <>
__device__ __forceinline__ void acc_add<float>(float* p, float v) { *p = __fadd_rn(*p, v); }
template <>
__device__ __forceinline__ void acc_add<__half>(__half* p, float v) {
  *p = __float2half_rn(__fadd_rn(__half2float(*p), v));
}

// Tiles of one launch ("round"): mutually non-overlapping, so they can be applied concurrently; overlapping
// tiles always sit in different rounds, in tile order, which keeps the per-voxel summation order of the reference.
#define FNNU_ROUND_MAX 8
struct TileRound {
  int n;
  int sx[FNNU_ROUND_MAX], sy[FNNU_ROUND_MAX], sz[FNNU_ROUND_MAX];
  int idx[FNNU_ROUND_MAX];    // index of the tile inside the call's prediction batch
};

// Generic path: any heads / stride / alignment.  One thread per tile voxel; blockIdx.y = tile of the round.
template <typename InT, typename AccT>
__global__ void __launch_bounds__(256) accumulate_generic_kernel(
    const InT* __restrict__ preds_all, int ps, int heads, TileRound tr, int pX, int pY, int pZ,
    FlipList flips, int n_flips, const __half* __restrict__ gauss, AccT* __restrict__ acc, int X, int Y,
    int Z) {
  const size_t pvox = (size_t)pX * pY * pZ;
  const int sx = tr.sx[blockIdx.y], sy = tr.sy[blockIdx.y], sz = tr.sz[blockIdx.y];
  const InT* __restrict__ preds = preds_all + (size_t)tr.idx[blockIdx.y] * n_flips * pvox * ps;
  const size_t hstride = (size_t)X * Y * Z;
  const float inv_dummy = (float)n_flips;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pvox;
       i += (size_t)gridDim.x * blockDim.x) {
    int z = (int)(i % pZ);
    int y = (int)((i / pZ) % pY);
    int x = (int)(i / ((size_t)pZ * pY));
    float g = gauss ? __half2float(gauss[i]) : 1.f;
    size_t a = ((size_t)(sx + x) * Y + (sy + y)) * Z + (sz + z);
    for (int h = 0; h < heads; ++h) {
      float s = 0.f;
      for (int f = 0; f < n_flips; ++f) {
        int m = flips.m[f];
        int fx = (m & 1) ? pX - 1 - x : x;
        int fy = (m & 2) ? pY - 1 - y : y;
        int fz = (m & 4) ? pZ - 1 - z : z;
        size_t src = ((size_t)f * pvox + ((size_t)fx * pY + fy) * pZ + fz) * ps + h;
        float v = to_f<InT>(preds[src]);
        s = (f == 0) ? v : __fadd_rn(s, v);
      }
      if (n_flips > 1) s = __fdiv_rn(s, inv_dummy);
      if (gauss) s = __fmul_rn(s, g);
      acc_add<AccT>(acc + h * hstride + a, s);
    }
  }
}

// Fast path: fp16 predictions, 2 heads stored as half2 per voxel, fp32 accumulators, pZ % 4 == 0,
// (sz % 4 == 0 && Z % 4 == 0): one thread handles 4 consecutive z voxels with 128-bit accesses.
__global__ void __launch_bounds__(256) accumulate_h2_vec4_kernel(
    const __half* __restrict__ preds_all, TileRound tr, int pX, int pY, int pZ, FlipList flips,
    int n_flips, const __half* __restrict__ gauss, float* __restrict__ acc, int X, int Y, int Z) {
  const int zq = pZ >> 2;
  const size_t nthreads = (size_t)pX * pY * zq;
  const size_t pvox = (size_t)pX * pY * pZ;
  const int sx = tr.sx[blockIdx.y], sy = tr.sy[blockIdx.y], sz = tr.sz[blockIdx.y];
  const __half* __restrict__ preds = preds_all + (size_t)tr.idx[blockIdx.y] * n_flips * pvox * 2;
  const size_t hstride = (size_t)X * Y * Z;
  const float nf = (float)n_flips;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nthreads;
       i += (size_t)gridDim.x * blockDim.x) {
    int z = (int)(i % zq) << 2;
    int y = (int)((i / zq) % pY);
    int x = (int)(i / ((size_t)zq * pY));
    float s0[4], s1[4];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      if (f < n_flips) {
        int m = flips.m[f];
        int fx = (m & 1) ? pX - 1 - x : x;
        int fy = (m & 2) ? pY - 1 - y : y;
        bool rz = (m & 4) != 0;
        int fz = rz ? pZ - 4 - z : z;
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(
            preds + ((size_t)f * pvox + ((size_t)fx * pY + fy) * pZ + fz) * 2));
        const __half2* hp = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 v = __half22float2(hp[rz ? 3 - k : k]);
          if (f == 0) {
            s0[k] = v.x;
            s1[k] = v.y;
          } else {
            s0[k] = __fadd_rn(s0[k], v.x);
            s1[k] = __fadd_rn(s1[k], v.y);
          }
        }
      }
    }
    float g[4] = {1.f, 1.f, 1.f, 1.f};
    if (gauss) {
      const uint2 graw = __ldg(reinterpret_cast<const uint2*>(gauss + ((size_t)x * pY + y) * pZ + z));
      const __half2* gp = reinterpret_cast<const __half2*>(&graw);
      float2 g01 = __half22float2(gp[0]), g23 = __half22float2(gp[1]);
      g[0] = g01.x; g[1] = g01.y; g[2] = g23.x; g[3] = g23.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (n_flips > 1) {
        s0[k] = __fdiv_rn(s0[k], nf);
        s1[k] = __fdiv_rn(s1[k], nf);
      }
      if (gauss) {
        s0[k] = __fmul_rn(s0[k], g[k]);
        s1[k] = __fmul_rn(s1[k], g[k]);
      }
    }
    size_t a = ((size_t)(sx + x) * Y + (sy + y)) * Z + (sz + z);
    float4* a0 = reinterpret_cast<float4*>(acc + a);
    float4* a1 = reinterpret_cast<float4*>(acc + hstride + a);
    float4 v0 = *a0, v1 = *a1;
    v0.x = __fadd_rn(v0.x, s0[0]); v0.y = __fadd_rn(v0.y, s0[1]);
    v0.z = __fadd_rn(v0.z, s0[2]); v0.w = __fadd_rn(v0.w, s0[3]);
    v1.x = __fadd_rn(v1.x, s1[0]); v1.y = __fadd_rn(v1.y, s1[1]);
    v1.z = __fadd_rn(v1.z, s1[2]); v1.w = __fadd_rn(v1.w, s1[3]);
    *a0 = v0;
    *a1 = v1;
  }
}

// Cluster path (fp16 predictions, fp32 accumulators, z starts / extents multiples of VZ): ONE launch applies a group
// of up to 8 tiles that overlap each other.  A thread owns VZ consecutive z voxels x HC heads of the group's bounding
// box, loads the accumulator once, adds the contribution of every covering tile IN TILE ORDER (the reference's
// summation order, bit for bit: acc = (acc + c_t0) + c_t1 ...) and stores once — overlapping tiles cost one
// read-modify-write per voxel instead of one per tile, and no launch boundary is needed between them.
// Heads are contiguous per voxel in the channels-last prediction: a voxel's HC heads are one 4 / 8 / 16 / 32-byte load.
struct TileCluster {
  int n;
  int sx[FNNU_ROUND_MAX], sy[FNNU_ROUND_MAX], sz[FNNU_ROUND_MAX];
  int idx[FNNU_ROUND_MAX];
  int ox, oy, oz;       // bounding box origin (accumulator coordinates)
  int bx, by, bz;       // bounding box extents
};

template <int HC>
__device__ __forceinline__ void load_heads(const __half* p, float* out) {
  if constexpr (HC == 2) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(p));
    out[0] = v.x; out[1] = v.y;
  } else if constexpr (HC == 4) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float2 v = __half22float2(h[k]);
      out[2 * k] = v.x; out[2 * k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < HC / 8; ++q) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p) + q);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 v = __half22float2(h[k]);
        out[8 * q + 2 * k] = v.x; out[8 * q + 2 * k + 1] = v.y;
      }
    }
  }
}

// CONTIG: the prediction stride equals HC, so a thread's VZ voxels x HC heads are VZ * HC * 2 contiguous bytes (needs
// tile z starts that are multiples of VZ).  !CONTIG: every voxel's HC heads are one aligned 16 / 32-byte piece, loaded
// voxel by voxel, so tile starts may be ANY integer (cfg 5's tiles start at 46, 139, 185 ...): a thread's VZ
// accumulator voxels are aligned in the ACCUMULATOR, and each of them checks its own coverage.
// SINGLE: heads <= HC, one chunk per voxel (no head-chunk index arithmetic).
template <int HC, int VZ, bool CONTIG, bool SINGLE>
__global__ void __launch_bounds__(256, (HC <= 2 ? 4 : (HC <= 4 ? 3 : 2))) accumulate_cluster_kernel(
    const __half* __restrict__ preds_all, int ps, int heads, TileCluster tc, int pX, int pY, int pZ, FlipList flips,
    int n_flips, const __half* __restrict__ gauss, float* __restrict__ acc, int X, int Y, int Z) {
  // grid: x = chunks of 256 over (y, z groups[, head chunks]) of one box plane, y = box plane: 32-bit index arithmetic
  constexpr int NQ = VZ * HC / 8;            // 16-byte loads per flip and thread
  constexpr int G = 8 / NQ;                  // flips whose loads are in flight together (8 x 16 bytes per thread)
  constexpr int PQ = HC >= 8 ? HC / 8 : 1;   // 16-byte pieces per voxel (non-contiguous case)
  static_assert(NQ >= 1 && NQ <= 4, "accumulate_cluster_kernel: unsupported (HC, VZ)");
  static_assert(CONTIG || HC >= 8, "voxel-by-voxel loads need at least 16 bytes of heads per voxel");
  // Head chunks are the FASTEST thread index: with more heads than HC, neighbouring lanes read neighbouring
  // 16 / 32-byte pieces of the same voxel, so a warp's load covers whole 128-byte lines.
  const int zq = tc.bz / VZ;
  const int n_hc = SINGLE ? 1 : (heads + HC - 1) / HC;
  const uint32_t pv32 = (uint32_t)pX * (uint32_t)pY * (uint32_t)pZ;
  const size_t hstride = (size_t)X * Y * Z;
  const float nf = (float)n_flips;
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= (uint32_t)(tc.by * zq * n_hc)) return;
  const uint32_t vox = SINGLE ? idx : idx / (uint32_t)n_hc;
  const int h0 = SINGLE ? 0 : (int)(idx - vox * (uint32_t)n_hc) * HC;
  const int yy = (int)(vox / (uint32_t)zq);
  const int gz = tc.oz + (int)(vox - (uint32_t)yy * (uint32_t)zq) * VZ;     // multiple of VZ in the accumulator
  const int gy = tc.oy + yy;
  const int gx = tc.ox + (int)blockIdx.y;
  float* const a0 = acc + ((size_t)gx * Y + gy) * Z + gz;
  float r[VZ][HC];
  bool any = false;
  for (int t = 0; t < tc.n; ++t) {
    const int lx = gx - tc.sx[t], ly = gy - tc.sy[t], lz = gz - tc.sz[t];
    if ((unsigned)lx >= (unsigned)pX || (unsigned)ly >= (unsigned)pY) continue;
    bool vok[VZ];           // voxel v of this thread lies inside tile t
    bool some = false;
#pragma unroll
    for (int v = 0; v < VZ; ++v) {
      vok[v] = (unsigned)(lz + v) < (unsigned)pZ;
      some = some || vok[v];
    }
    if (CONTIG ? !vok[0] : !some) continue;       // CONTIG: lz is a multiple of VZ, all or nothing
    if (!any) {
      any = true;
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        if (h0 + j < heads) {
          if constexpr (VZ == 4) {
            const float4 v = *reinterpret_cast<const float4*>(a0 + (size_t)(h0 + j) * hstride);
            r[0][j] = v.x; r[1][j] = v.y; r[2][j] = v.z; r[3][j] = v.w;
          } else {
            const float2 v = *reinterpret_cast<const float2*>(a0 + (size_t)(h0 + j) * hstride);
            r[0][j] = v.x; r[1][j] = v.y;
          }
        }
      }
    }
    float g[VZ];
#pragma unroll
    for (int v = 0; v < VZ; ++v) g[v] = 1.f;
    if (gauss) {
      const __half* gp = gauss + ((size_t)lx * pY + ly) * pZ + lz;
      if constexpr (CONTIG) {
#pragma unroll
        for (int v = 0; v < VZ; v += 2) {
          const float2 gv = __half22float2(*reinterpret_cast<const __half2*>(gp + v));
          g[v] = gv.x; g[v + 1] = gv.y;
        }
      } else {
#pragma unroll
        for (int v = 0; v < VZ; ++v)
          if (vok[v]) g[v] = __half2float(gp[v]);
      }
    }
    const __half* __restrict__ preds = preds_all + (size_t)tc.idx[t] * n_flips * pv32 * ps + h0;
    float s[VZ][HC];
#pragma unroll
    for (int f0 = 0; f0 < 8; f0 += G) {
      if (f0 < n_flips) {
        // ---- all loads of G flips first (independent, 8 x 16 bytes in flight per thread) ...
        uint4 raw[G][NQ];
        bool rzs[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int f = f0 + i;
          const int m = flips.m[f];
          const int fx = (m & 1) ? pX - 1 - lx : lx;
          const int fy = (m & 2) ? pY - 1 - ly : ly;
          rzs[i] = (m & 4) != 0;
          // offsets inside one tile's predictions fit 32 bits (8 x 2.1 M voxels x 64 heads = 1.07 G elements)
          const uint32_t row = (uint32_t)f * pv32 + (uint32_t)((fx * pY + fy) * pZ);
          if constexpr (CONTIG) {
            const int fz = rzs[i] ? pZ - VZ - lz : lz;
            const __half* src = preds + (size_t)((row + (uint32_t)fz) * (uint32_t)ps);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              raw[i][q] = make_uint4(0, 0, 0, 0);
              if (f < n_flips) raw[i][q] = __ldg(reinterpret_cast<const uint4*>(src) + q);
            }
          } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const int v = q / PQ;                        // piece q belongs to accumulator voxel v
              raw[i][q] = make_uint4(0, 0, 0, 0);
              if (f < n_flips && vok[v]) {
                const int fz = rzs[i] ? pZ - 1 - (lz + v) : lz + v;
                raw[i][q] = __ldg(reinterpret_cast<const uint4*>(preds + (size_t)((row + (uint32_t)fz) * (uint32_t)ps)) + (q % PQ));
              }
            }
          }
        }
        // ---- ... then the sums, flip after flip (the reference's order)
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int f = f0 + i;
          if (f < n_flips) {
            const uint32_t* w32 = reinterpret_cast<const uint32_t*>(&raw[i][0]);      // one half2 per word
#pragma unroll
            for (int v = 0; v < VZ; ++v) {
#pragma unroll
              for (int j = 0; j < HC; j += 2) {
                uint32_t w = w32[(v * HC + j) >> 1];
                if constexpr (CONTIG) {
                  // the mirrored copy stores the voxels in reverse: select the VALUE (both indices are compile-time
                  // constants, so the loaded words stay in registers)
                  const uint32_t wr = w32[((VZ - 1 - v) * HC + j) >> 1];
                  w = rzs[i] ? wr : w;
                }
                const float2 pv = __half22float2(*reinterpret_cast<const __half2*>(&w));
                s[v][j] = (f == 0) ? pv.x : __fadd_rn(s[v][j], pv.x);
                s[v][j + 1] = (f == 0) ? pv.y : __fadd_rn(s[v][j + 1], pv.y);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VZ; ++v) {
      if (CONTIG || vok[v]) {
#pragma unroll
        for (int j = 0; j < HC; ++j) {
          float c = s[v][j];
          if (n_flips > 1) c = __fdiv_rn(c, nf);
          if (gauss) c = __fmul_rn(c, g[v]);
          r[v][j] = __fadd_rn(r[v][j], c);
        }
      }
    }
  }
  if (any) {
#pragma unroll
    for (int j = 0; j < HC; ++j) {
      if (h0 + j < heads) {
        if constexpr (VZ == 4)
          *reinterpret_cast<float4*>(a0 + (size_t)(h0 + j) * hstride) = make_float4(r[0][j], r[1][j], r[2][j], r[3][j]);
        else
          *reinterpret_cast<float2*>(a0 + (size_t)(h0 + j) * hstride) = make_float2(r[0][j], r[1][j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight sum (n_predictions): gather over the covering tiles, in tile order.
// ------------------------------------------------------------------------------------------------
#define FNNU_MAX_STEPS 160
struct StepLists {
  int32_t s[3][FNNU_MAX_STEPS];
  int n[3];
};

template <typename AccT>
__global__ void __launch_bounds__(256) weight_sum_kernel(StepLists st, int pX, int pY, int pZ,
                                                         const __half* __restrict__ gauss,
                                                         AccT* __restrict__ wsum, int X, int Y, int Z) {
  const size_t total = (size_t)X * Y * Z;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int z = (int)(i % Z);
    int y = (int)((i / Z) % Y);
    int x = (int)(i / ((size_t)Z * Y));
    float w = 0.f;
    for (int a = 0; a < st.n[0]; ++a) {
      int lx = x - st.s[0][a];
      if (lx < 0 || lx >= pX) continue;
      for (int b = 0; b < st.n[1]; ++b) {
        int ly = y - st.s[1][b];
        if (ly < 0 || ly >= pY) continue;
        for (int c = 0; c < st.n[2]; ++c) {
          int lz = z - st.s[2][c];
          if (lz < 0 || lz >= pZ) continue;
          float g = gauss ? __half2float(__ldg(gauss + ((size_t)lx * pY + ly) * pZ + lz)) : 1.f;
          w = __fadd_rn(w, g);
          if (sizeof(AccT) == 2) w = __half2float(__float2half_rn(w));
        }
      }
    }
    if (sizeof(AccT) == 2)
      reinterpret_cast<__half*>(wsum)[i] = __float2half_rn(w);
    else
      reinterpret_cast<float*>(wsum)[i] = w;
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: logits = acc / wsum (rounded to fp16 like the reference's half buffers), inf check,
// argmax with first-maximum tie-break.
// ------------------------------------------------------------------------------------------------
template <typename AccT>
__global__ void __launch_bounds__(256) finalize_kernel(const AccT* __restrict__ acc,
                                                       const AccT* __restrict__ wsum, int heads,
                                                       size_t nvox, size_t hs, __half* __restrict__ logits,
                                                       uint8_t* __restrict__ labels,
                                                       int32_t* __restrict__ inf_flag) {
  bool saw_inf = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox;
       i += (size_t)gridDim.x * blockDim.x) {
    float w = to_f<AccT>(wsum[i]);
    float best = 0.f;
    int arg = 0;
    for (int h = 0; h < heads; ++h) {
      float v = __fdiv_rn(to_f<AccT>(acc[(size_t)h * hs + i]), w);
      __half hv = __float2half_rn(v);
      float r = __half2float(hv);
      if (isinf(r)) saw_inf = true;
      if (logits) logits[(size_t)h * nvox + i] = hv;
      if (h == 0 || r > best) {   // strict '>' keeps the first maximum (numpy argmax)
        best = r;
        arg = h;
      }
    }
    if (labels) labels[i] = (uint8_t)arg;
  }
  if (saw_inf && inf_flag) atomicOr(inf_flag, 1);
}

// 2 heads, fp32 accumulators, nvox % 4 == 0: 128-bit loads, 64-bit logit stores, 32-bit label stores.
__global__ void __launch_bounds__(256) finalize_h2_vec4_kernel(const float* __restrict__ acc,
                                                               const float* __restrict__ wsum,
                                                               size_t nvox, size_t hs, __half* __restrict__ logits,
                                                               uint8_t* __restrict__ labels,
                                                               int32_t* __restrict__ inf_flag) {
  bool saw_inf = false;
  const size_t nq = nvox >> 2;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
       q += (size_t)gridDim.x * blockDim.x) {
    float4 w = __ldg(reinterpret_cast<const float4*>(wsum) + q);
    float4 a0 = __ldg(reinterpret_cast<const float4*>(acc) + q);
    float4 a1 = __ldg(reinterpret_cast<const float4*>(acc + hs) + q);
    float wv[4] = {w.x, w.y, w.z, w.w};
    float v0[4] = {a0.x, a0.y, a0.z, a0.w};
    float v1[4] = {a1.x, a1.y, a1.z, a1.w};
    __half h0[4], h1[4];
    uint32_t lab = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h0[k] = __float2half_rn(__fdiv_rn(v0[k], wv[k]));
      h1[k] = __float2half_rn(__fdiv_rn(v1[k], wv[k]));
      float r0 = __half2float(h0[k]), r1 = __half2float(h1[k]);
      if (isinf(r0) || isinf(r1)) saw_inf = true;
      if (r1 > r0) lab |= (1u << (8 * k));
    }
    if (logits) {
      uint2 o0, o1;
      o0.x = (uint32_t)__half_as_ushort(h0[0]) | ((uint32_t)__half_as_ushort(h0[1]) << 16);
      o0.y = (uint32_t)__half_as_ushort(h0[2]) | ((uint32_t)__half_as_ushort(h0[3]) << 16);
      o1.x = (uint32_t)__half_as_ushort(h1[0]) | ((uint32_t)__half_as_ushort(h1[1]) << 16);
      o1.y = (uint32_t)__half_as_ushort(h1[2]) | ((uint32_t)__half_as_ushort(h1[3]) << 16);
      reinterpret_cast<uint2*>(logits)[q] = o0;
      reinterpret_cast<uint2*>(logits + nvox)[q] = o1;
    }
    if (labels) reinterpret_cast<uint32_t*>(labels)[q] = lab;
  }
  if (saw_inf && inf_flag) atomicOr(inf_flag, 1);
}

// Any number of heads, fp32 accumulators, nvox % 4 == 0: a thread normalises 4 consecutive voxels of every head
// (128-bit loads, 64-bit logit stores, one 32-bit label store).
__global__ void __launch_bounds__(256) finalize_vec4_kernel(const float* __restrict__ acc, const float* __restrict__ wsum,
                                                            int heads, size_t nvox, size_t hs, __half* __restrict__ logits,
                                                            uint8_t* __restrict__ labels, int32_t* __restrict__ inf_flag) {
  bool saw_inf = false;
  const size_t nq = nvox >> 2;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wsum) + q);
    const float wv[4] = {w.x, w.y, w.z, w.w};
    float best[4];
    uint32_t arg[4] = {0, 0, 0, 0};
    for (int h0 = 0; h0 < heads; h0 += 8) {
      float4 a8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)       // eight independent 128-bit loads in flight per thread
        if (h0 + i < heads) a8[i] = __ldg(reinterpret_cast<const float4*>(acc + (size_t)(h0 + i) * hs) + q);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int h = h0 + i;
        if (h < heads) {
          const float av[4] = {a8[i].x, a8[i].y, a8[i].z, a8[i].w};
          __half hv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            hv[k] = __float2half_rn(__fdiv_rn(av[k], wv[k]));
            const float r = __half2float(hv[k]);
            if (isinf(r)) saw_inf = true;
            if (h == 0 || r > best[k]) {      // strict '>' keeps the first maximum (numpy argmax)
              best[k] = r;
              arg[k] = (uint32_t)h;
            }
          }
          if (logits) {
            uint2 o;
            o.x = (uint32_t)__half_as_ushort(hv[0]) | ((uint32_t)__half_as_ushort(hv[1]) << 16);
            o.y = (uint32_t)__half_as_ushort(hv[2]) | ((uint32_t)__half_as_ushort(hv[3]) << 16);
            reinterpret_cast<uint2*>(logits + (size_t)h * nvox)[q] = o;
          }
        }
      }
    }
    if (labels) reinterpret_cast<uint32_t*>(labels)[q] = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  }
  if (saw_inf && inf_flag) atomicOr(inf_flag, 1);
}

__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ acc,
                                                          const float* __restrict__ other, size_t n, int aligned) {
  const size_t nq = aligned ? (n >> 2) : 0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
       q += (size_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(acc)[q];
    float4 b = __ldg(reinterpret_cast<const float4*>(other) + q);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(acc)[q] = a;
  }
  for (size_t i = (nq << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    acc[i] += other[i];
}

__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ x, float f, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = __fmul_rn(x[i], f);
}

static inline int grid_for(size_t work_items, int threads, int waves_cap = 16) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)num_sms() * waves_cap;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int)blocks;
}

}  // namespace fnnu

using namespace fnnu;

extern "C" int fnnu_gather_tiles(const float* volume, int channels, const int vol_dims[3],
                                 const int32_t* starts_dev, int n_tiles, const int patch[3],
                                 const uint8_t* flip_masks, int n_flips, void* out, int c_stride,
                                 void* stream) {
  FNNU_CHECK_ARG(volume && starts_dev && out && vol_dims && patch && flip_masks, "gather: null pointer");
  FNNU_CHECK_ARG(n_tiles > 0 && n_flips >= 1 && n_flips <= 8, "gather: n_tiles=%d n_flips=%d", n_tiles, n_flips);
  FNNU_CHECK_ARG(channels >= 1 && c_stride >= channels, "gather: channels=%d stride=%d", channels, c_stride);
  for (int a = 0; a < 3; ++a)
    FNNU_CHECK_ARG(patch[a] >= 1 && vol_dims[a] >= patch[a], "gather: volume smaller than patch on axis %d", a);
  FlipList fl;
  for (int i = 0; i < 8; ++i) fl.m[i] = i < n_flips ? (flip_masks[i] & 7) : 0;
  FNNU_CHECK_ARG(patch[0] <= 65535 && n_tiles * n_flips <= 65535, "gather: grid extents");
  if (channels == 1 && c_stride == 1 && patch[2] % 8 == 0 && ((uintptr_t)out % 16) == 0) {
    const dim3 g8((unsigned)((patch[1] * (patch[2] / 8) + 255) / 256), (unsigned)patch[0], (unsigned)n_tiles);
    gather_tiles_c1_vec8_kernel<<<g8, 256, 0, (cudaStream_t)stream>>>(
        volume, vol_dims[0], vol_dims[1], vol_dims[2], starts_dev, n_tiles, patch[0], patch[1], patch[2], fl, n_flips,
        (__half*)out);
    FNNU_LAUNCH_CHECK();
    ++g_mem_launches;
    return FNNU_OK;
  }
  const dim3 gg((unsigned)((patch[1] * patch[2] + 255) / 256), (unsigned)patch[0], (unsigned)(n_tiles * n_flips));
  gather_tiles_kernel<<<gg, 256, 0, (cudaStream_t)stream>>>(
      volume, channels, vol_dims[0], vol_dims[1], vol_dims[2], starts_dev, n_tiles, patch[0], patch[1],
      patch[2], fl, n_flips, (__half*)out, c_stride);
  FNNU_LAUNCH_CHECK();
  ++g_mem_launches;
  return FNNU_OK;
}

extern "C" int fnnu_accumulate_tiles(const void* preds, int in_dtype, int p_stride, int heads,
                                     const int32_t* starts_host, int n_tiles, const int patch[3],
                                     const uint8_t* flip_masks, int n_flips, const void* gaussian,
                                     void* acc, int acc_dtype, const int vol_dims[3], void* stream) {
  FNNU_CHECK_ARG(preds && starts_host && acc && vol_dims && patch && flip_masks, "accumulate: null pointer");
  FNNU_CHECK_ARG(n_tiles > 0 && n_flips >= 1 && n_flips <= 8, "accumulate: n_tiles=%d n_flips=%d", n_tiles, n_flips);
  FNNU_CHECK_ARG(heads >= 1 && p_stride >= heads, "accumulate: heads=%d stride=%d", heads, p_stride);
  FNNU_CHECK_ARG(in_dtype == FNNU_IN_F16 || in_dtype == FNNU_IN_F32, "accumulate: in_dtype=%d", in_dtype);
  FNNU_CHECK_ARG(acc_dtype == FNNU_ACC_F32 || acc_dtype == FNNU_ACC_F16, "accumulate: acc_dtype=%d", acc_dtype);
  FlipList fl;
  for (int i = 0; i < 8; ++i) fl.m[i] = i < n_flips ? (flip_masks[i] & 7) : 0;
  const int pX = patch[0], pY = patch[1], pZ = patch[2];
  const int X = vol_dims[0], Y = vol_dims[1], Z = vol_dims[2];
  const size_t pvox = (size_t)pX * pY * pZ;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = in_dtype == FNNU_IN_F16 && acc_dtype == FNNU_ACC_F32 && heads == 2 && p_stride == 2 && (pZ % 4 == 0) &&
             (Z % 4 == 0) && (((uintptr_t)preds) % 16 == 0) && (((uintptr_t)acc) % 16 == 0) &&
             (!gaussian || ((uintptr_t)gaussian) % 8 == 0);
  for (int t = 0; t < n_tiles; ++t) {
    int sx = starts_host[t * 3 + 0], sy = starts_host[t * 3 + 1], sz = starts_host[t * 3 + 2];
    FNNU_CHECK_ARG(sx >= 0 && sy >= 0 && sz >= 0 && sx + pX <= X && sy + pY <= Y && sz + pZ <= Z,
                   "accumulate: tile %d (%d,%d,%d) outside the volume", t, sx, sy, sz);
    if (sz % 4 != 0) vec = false;
  }
  auto overlaps = [&](int t, int u) {
    return abs(starts_host[t * 3] - starts_host[u * 3]) < pX && abs(starts_host[t * 3 + 1] - starts_host[u * 3 + 1]) < pY &&
           abs(starts_host[t * 3 + 2] - starts_host[u * 3 + 2]) < pZ;
  };
  // ---- cluster path: groups of <= 8 consecutive tiles, one launch per group over its bounding box ----
  int hc = 0, vz = 4;
  if (in_dtype == FNNU_IN_F16 && acc_dtype == FNNU_ACC_F32 && ((uintptr_t)preds % 16 == 0) && ((uintptr_t)acc % 16 == 0) &&
      (!gaussian || ((uintptr_t)gaussian) % 4 == 0)) {
    if (p_stride == 2 && heads == 2) hc = 2;
    else if (p_stride == 4) hc = 4;
    else if (p_stride == 8) hc = 8;
    else if (p_stride % 16 == 0) { hc = 16; vz = 2; }
    else if (p_stride % 8 == 0) hc = 8;
    if (hc && Z % vz) hc = 0;                       // the accumulator rows must hold whole z groups
    if (hc && hc == p_stride) {                     // contiguous loads: tile z starts and extents in whole groups
      if (pZ % vz) hc = 0;
      for (int t = 0; t < n_tiles && hc; ++t)
        if (starts_host[t * 3 + 2] % vz) hc = 0;
    }
    static int cluster_enabled = -1;
    if (cluster_enabled < 0) {
      const char* e = getenv("FNNU_ACC_CLUSTER");
      cluster_enabled = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (!cluster_enabled) hc = 0;
  }
  if (hc) {
    int t = 0;
    while (t < n_tiles) {
      // a group: consecutive tiles, each overlapping at least one earlier tile of the group (so that the bounding box
      // stays tight); non-overlapping neighbours start a new group
      TileCluster tc;
      tc.n = 0;
      int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      for (; t < n_tiles && tc.n < FNNU_ROUND_MAX; ++t) {
        bool joins = tc.n == 0;
        for (int k = 0; k < tc.n && !joins; ++k) joins = overlaps(t, tc.idx[k]);
        if (!joins) break;
        const int st[3] = {starts_host[t * 3], starts_host[t * 3 + 1], starts_host[t * 3 + 2]};
        const int pp[3] = {pX, pY, pZ};
        for (int a = 0; a < 3; ++a) {
          if (tc.n == 0 || st[a] < lo[a]) lo[a] = st[a];
          if (tc.n == 0 || st[a] + pp[a] > hi[a]) hi[a] = st[a] + pp[a];
        }
        tc.sx[tc.n] = st[0]; tc.sy[tc.n] = st[1]; tc.sz[tc.n] = st[2];
        tc.idx[tc.n] = t;
        ++tc.n;
      }
      // the box is aligned to whole z groups of the ACCUMULATOR (tile starts may be odd)
      lo[2] = lo[2] / vz * vz;
      hi[2] = (hi[2] + vz - 1) / vz * vz;
      tc.ox = lo[0]; tc.oy = lo[1]; tc.oz = lo[2];
      tc.bx = hi[0] - lo[0]; tc.by = hi[1] - lo[1]; tc.bz = hi[2] - lo[2];
      FNNU_CHECK_ARG((size_t)8 * pvox * p_stride < ((size_t)1 << 32) && tc.bx <= 65535, "accumulate: tile too large for the 32-bit offsets");
      const int n_hc = (heads + hc - 1) / hc;
      const dim3 grid((unsigned)((tc.by * (tc.bz / vz) * n_hc + 255) / 256), (unsigned)tc.bx);
      const __half* pr = (const __half*)preds;
      const __half* gs = (const __half*)gaussian;
      float* ac = (float*)acc;
#define FNNU_ACC_LAUNCH(HCV, VZV, CT, SG) accumulate_cluster_kernel<HCV, VZV, CT, SG><<<grid, 256, 0, s>>>(pr, p_stride, heads, tc, pX, pY, pZ, fl, n_flips, gs, ac, X, Y, Z)
      const bool contig = p_stride == hc;
      if (hc == 2) FNNU_ACC_LAUNCH(2, 4, true, true);
      else if (hc == 4) FNNU_ACC_LAUNCH(4, 4, true, true);
      else if (hc == 8 && contig) FNNU_ACC_LAUNCH(8, 4, true, true);
      else if (hc == 8) FNNU_ACC_LAUNCH(8, 4, false, false);
      else if (contig) FNNU_ACC_LAUNCH(16, 2, true, true);
      else FNNU_ACC_LAUNCH(16, 2, false, false);
#undef FNNU_ACC_LAUNCH
      FNNU_LAUNCH_CHECK();
      ++g_mem_launches;
    }
    return FNNU_OK;
  }
  // ---- round path (any dtype / alignment): round[t] = 1 + max round of the earlier tiles that overlap t ----
  std::vector<int> round(n_tiles, 0);
  int n_rounds = 0;
  for (int t = 0; t < n_tiles; ++t) {
    for (int u = 0; u < t; ++u)
      if (overlaps(t, u) && round[u] + 1 > round[t]) round[t] = round[u] + 1;
    if (round[t] + 1 > n_rounds) n_rounds = round[t] + 1;
  }
  for (int r = 0; r < n_rounds; ++r) {
    int t = 0;
    while (t < n_tiles) {
      TileRound tr;
      tr.n = 0;
      for (; t < n_tiles && tr.n < FNNU_ROUND_MAX; ++t) {
        if (round[t] != r) continue;
        tr.sx[tr.n] = starts_host[t * 3];
        tr.sy[tr.n] = starts_host[t * 3 + 1];
        tr.sz[tr.n] = starts_host[t * 3 + 2];
        tr.idx[tr.n] = t;
        ++tr.n;
      }
      if (tr.n == 0) break;
      if (vec) {
        dim3 grid((unsigned)grid_for(pvox / 4, 256, 16 / (tr.n > 4 ? 4 : tr.n) + 1), (unsigned)tr.n);
        accumulate_h2_vec4_kernel<<<grid, 256, 0, s>>>((const __half*)preds, tr, pX, pY, pZ, fl, n_flips,
                                                       (const __half*)gaussian, (float*)acc, X, Y, Z);
      } else {
        dim3 grid((unsigned)grid_for(pvox, 256), (unsigned)tr.n);
        if (in_dtype == FNNU_IN_F16 && acc_dtype == FNNU_ACC_F32)
          accumulate_generic_kernel<__half, float><<<grid, 256, 0, s>>>((const __half*)preds, p_stride, heads, tr, pX, pY, pZ, fl,
                                                                        n_flips, (const __half*)gaussian, (float*)acc, X, Y, Z);
        else if (in_dtype == FNNU_IN_F16)
          accumulate_generic_kernel<__half, __half><<<grid, 256, 0, s>>>((const __half*)preds, p_stride, heads, tr, pX, pY, pZ, fl,
                                                                         n_flips, (const __half*)gaussian, (__half*)acc, X, Y, Z);
        else if (acc_dtype == FNNU_ACC_F32)
          accumulate_generic_kernel<float, float><<<grid, 256, 0, s>>>((const float*)preds, p_stride, heads, tr, pX, pY, pZ, fl,
                                                                       n_flips, (const __half*)gaussian, (float*)acc, X, Y, Z);
        else
          accumulate_generic_kernel<float, __half><<<grid, 256, 0, s>>>((const float*)preds, p_stride, heads, tr, pX, pY, pZ, fl,
                                                                        n_flips, (const __half*)gaussian, (__half*)acc, X, Y, Z);
      }
      FNNU_LAUNCH_CHECK();
      ++g_mem_launches;
    }
  }
  return FNNU_OK;
}

extern "C" long long fnnu_mem_launches(void) { return (long long)g_mem_launches; }

extern "C" int fnnu_weight_sum(const int32_t* steps_x, int nx, const int32_t* steps_y, int ny,
                               const int32_t* steps_z, int nz, const int patch[3], const void* gaussian,
                               void* wsum, int acc_dtype, const int vol_dims[3], void* stream) {
  FNNU_CHECK_ARG(steps_x && steps_y && steps_z && patch && wsum && vol_dims, "weight_sum: null pointer");
  FNNU_CHECK_ARG(nx >= 1 && ny >= 1 && nz >= 1 && nx <= FNNU_MAX_STEPS && ny <= FNNU_MAX_STEPS && nz <= FNNU_MAX_STEPS,
                 "weight_sum: step counts %d %d %d (max %d per axis)", nx, ny, nz, FNNU_MAX_STEPS);
  FNNU_CHECK_ARG(acc_dtype == FNNU_ACC_F32 || acc_dtype == FNNU_ACC_F16, "weight_sum: acc_dtype=%d", acc_dtype);
  StepLists st;
  st.n[0] = nx; st.n[1] = ny; st.n[2] = nz;
  for (int i = 0; i < nx; ++i) st.s[0][i] = steps_x[i];
  for (int i = 0; i < ny; ++i) st.s[1][i] = steps_y[i];
  for (int i = 0; i < nz; ++i) st.s[2][i] = steps_z[i];
  size_t total = (size_t)vol_dims[0] * vol_dims[1] * vol_dims[2];
  cudaStream_t s = (cudaStream_t)stream;
  if (acc_dtype == FNNU_ACC_F32)
    weight_sum_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(st, patch[0], patch[1], patch[2],
        (const __half*)gaussian, (float*)wsum, vol_dims[0], vol_dims[1], vol_dims[2]);
  else
    weight_sum_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(st, patch[0], patch[1], patch[2],
        (const __half*)gaussian, (__half*)wsum, vol_dims[0], vol_dims[1], vol_dims[2]);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

extern "C" int fnnu_finalize(const void* acc, const void* wsum, int acc_dtype, int heads,
                             const int vol_dims[3], size_t acc_head_stride, void* logits_out,
                             uint8_t* labels_out, int32_t* inf_flag_dev, void* stream) {
  FNNU_CHECK_ARG(acc && wsum && vol_dims, "finalize: null pointer");
  FNNU_CHECK_ARG(heads >= 1 && heads <= 255, "finalize: heads=%d", heads);
  FNNU_CHECK_ARG(acc_dtype == FNNU_ACC_F32 || acc_dtype == FNNU_ACC_F16, "finalize: acc_dtype=%d", acc_dtype);
  size_t nvox = (size_t)vol_dims[0] * vol_dims[1] * vol_dims[2];
  size_t hs = acc_head_stride ? acc_head_stride : nvox;
  FNNU_CHECK_ARG(hs >= nvox, "finalize: head stride %zu < %zu voxels", hs, nvox);
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = acc_dtype == FNNU_ACC_F32 && heads == 2 && nvox % 4 == 0 && hs % 4 == 0 && ((uintptr_t)acc % 16 == 0) &&
             ((uintptr_t)wsum % 16 == 0) && (!logits_out || (uintptr_t)logits_out % 8 == 0) &&
             (!labels_out || (uintptr_t)labels_out % 4 == 0);
  if (vec)
    finalize_h2_vec4_kernel<<<grid_for(nvox / 4, 256), 256, 0, s>>>((const float*)acc, (const float*)wsum, nvox, hs,
                                                                    (__half*)logits_out, labels_out, inf_flag_dev);
  else if (acc_dtype == FNNU_ACC_F32 && nvox % 4 == 0 && hs % 4 == 0 && ((uintptr_t)acc % 16 == 0) && ((uintptr_t)wsum % 16 == 0) &&
           (!logits_out || (uintptr_t)logits_out % 8 == 0) && (!labels_out || (uintptr_t)labels_out % 4 == 0))
    finalize_vec4_kernel<<<grid_for(nvox / 4, 256), 256, 0, s>>>((const float*)acc, (const float*)wsum, heads, nvox, hs,
                                                                 (__half*)logits_out, labels_out, inf_flag_dev);
  else if (acc_dtype == FNNU_ACC_F32)
    finalize_kernel<float><<<grid_for(nvox, 256), 256, 0, s>>>((const float*)acc, (const float*)wsum, heads, nvox, hs,
                                                               (__half*)logits_out, labels_out, inf_flag_dev);
  else
    finalize_kernel<__half><<<grid_for(nvox, 256), 256, 0, s>>>((const __half*)acc, (const __half*)wsum, heads, nvox, hs,
                                                                (__half*)logits_out, labels_out, inf_flag_dev);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

extern "C" int fnnu_scale_inplace_f32(float* x, float factor, size_t n, void* stream) {
  FNNU_CHECK_ARG(x, "scale_inplace: null pointer");
  if (n == 0) return FNNU_OK;
  scale_inplace_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, factor, n);
  FNNU_LAUNCH_CHECK();
  ++g_mem_launches;
  return FNNU_OK;
}

extern "C" int fnnu_add_inplace_f32(float* acc, const float* other, size_t n, void* stream) {
  FNNU_CHECK_ARG(acc && other, "add_inplace: null pointer");
  if (n == 0) return FNNU_OK;
  int aligned = ((uintptr_t)acc % 16 == 0) && ((uintptr_t)other % 16 == 0);
  add_inplace_kernel<<<grid_for(aligned ? n / 4 + 1 : n, 256), 256, 0, (cudaStream_t)stream>>>(acc, other, n, aligned);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}
