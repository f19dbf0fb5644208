// CUDA-core direct convolution / transposed convolution, residual join and average pooling.
//
// This is the general-shape back end of the per-patch network forward (any channel count, kernel
// 1|3 per axis, stride 1|2 per axis).  It is used for the layers the tcgen05 implicit-GEMM kernel
// (conv_umma.cu) does not cover (e.g. the first layer with 1 or 4 input channels) and as the on-device
// cross-check of that kernel.  Semantics follow torch.nn.Conv3d / ConvTranspose3d / InstanceNorm3d /
// LeakyReLU as composed by dynamic_network_architectures' ConvDropoutNormReLU, StackedConvBlocks,
// BasicBlockD and UNetDecoder (call site: predict_from_raw_data.py:543 `self.network(x)`).
//
// Fusion contract shared with conv_umma.cu:
//   * the source is read RAW (pre-norm) and the producing layer's InstanceNorm affine + LeakyReLU is
//     applied while loading (per (sample, channel) scale/shift from the fp64 sums);
//   * the output is written RAW (conv + bias, rounded to fp16) and the per-(sample, channel) sum and
//     sum of squares of the ROUNDED values are accumulated in fp64 for the consumer.
#include <type_traits>
#include <cstdlib>
#include "common.cuh"
#include "ops.cuh"

namespace fnnu {

constexpr int kVX = 4;    // output voxels per thread (consecutive along the innermost axis)
constexpr int kCO = 16;   // output channels per thread
constexpr int kCI = 8;    // input channels per smem weight chunk
constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) conv_direct_kernel(ConvArgs a) {
  extern __shared__ float smem[];
  const int ntaps_blk = a.transposed ? 1 : a.ntaps;
  float* w_s = smem;                                   // [ntaps_blk][kCI][kCO]
  float* xs = w_s + ntaps_blk * kCI * kCO;             // scale [cin]
  float* xh = xs + a.cin;                              // shift [cin]
  float* xl = xh + a.cin;                              // slope [cin]
  float* red = xl + a.cin;                             // [kThreads/32][2*kCO]

  const int b = blockIdx.z;
  const int n_co_chunks = a.cout_pad / kCO;
  const int co_chunk = blockIdx.y % n_co_chunks;
  const int tap_fixed = blockIdx.y / n_co_chunks;      // transposed only
  const int co0 = co_chunk * kCO;

  // Domain the threads iterate over: conv -> output voxels; transposed -> input voxels.
  const int D0 = a.transposed ? a.in_d[0] : a.out_d[0];
  const int D1 = a.transposed ? a.in_d[1] : a.out_d[1];
  const int D2 = a.transposed ? a.in_d[2] : a.out_d[2];
  const int groups_per_row = (D2 + kVX - 1) / kVX;
  const long long n_groups = (long long)D0 * D1 * groups_per_row;
  const long long gidx = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool active = gidx < n_groups;
  int x0 = 0, y = 0, z = 0;
  if (active) {
    x0 = (int)(gidx % groups_per_row) * kVX;
    y = (int)((gidx / groups_per_row) % D1);
    z = (int)(gidx / ((long long)groups_per_row * D1));
  }

  // pending transform of the source channels for this sample
  for (int c = threadIdx.x; c < a.cin; c += kThreads) {
    float sc, sh;
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, a.src_meta[c], a.src_inv_count, sc, sh);
    xs[c] = sc;
    xh[c] = sh;
    xl[c] = a.src_meta[c].eps < 0.f ? 1.f : a.src_meta[c].slope;
  }

  float acc[kVX][kCO];
#pragma unroll
  for (int v = 0; v < kVX; ++v)
#pragma unroll
    for (int o = 0; o < kCO; ++o) acc[v][o] = 0.f;

  const size_t in_row = (size_t)a.src_cs;
  const __half* src_b = a.src + (size_t)b * a.in_d[0] * a.in_d[1] * a.in_d[2] * in_row;
  const bool vec_in = (a.src_cs % 8 == 0) && (((uintptr_t)a.src) % 16 == 0);

  for (int ci0 = 0; ci0 < a.cin; ci0 += kCI) {
    __syncthreads();
    for (int i = threadIdx.x; i < ntaps_blk * kCI * kCO; i += kThreads) {
      int o = i % kCO;
      int c = (i / kCO) % kCI;
      int t = i / (kCO * kCI);
      int tap = a.transposed ? tap_fixed : t;
      int ci = ci0 + c;
      w_s[i] = (ci < a.cin) ? __ldg(a.w + ((size_t)tap * a.cin + ci) * a.cout_pad + co0 + o) : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int t = 0; t < ntaps_blk; ++t) {
      int iz, iy, ixb, xstep;
      if (a.transposed) {
        iz = z; iy = y; ixb = x0; xstep = 1;
      } else {
        int kx = t % a.k[2];
        int ky = (t / a.k[2]) % a.k[1];
        int kz = t / (a.k[2] * a.k[1]);
        iz = z * a.s[0] + kz - a.pad[0];
        iy = y * a.s[1] + ky - a.pad[1];
        ixb = x0 * a.s[2] + kx - a.pad[2];
        xstep = a.s[2];
        if (iz < 0 || iz >= a.in_d[0] || iy < 0 || iy >= a.in_d[1]) continue;
      }
      float in[kVX][kCI];
#pragma unroll
      for (int v = 0; v < kVX; ++v) {
        int ix = ixb + v * xstep;
        bool ok = ix >= 0 && ix < a.in_d[2] && (x0 + v) < D2;
        const __half* p = src_b + (((size_t)iz * a.in_d[1] + iy) * a.in_d[2] + (ok ? ix : 0)) * in_row + ci0;
        if (ok && vec_in && ci0 + kCI <= a.cin) {
          uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int c = 0; c < kCI; c += 2) {
            float2 f = __half22float2(h2[c >> 1]);
            in[v][c] = lrelu(fmaf(f.x, xs[ci0 + c], xh[ci0 + c]), xl[ci0 + c]);
            in[v][c + 1] = lrelu(fmaf(f.y, xs[ci0 + c + 1], xh[ci0 + c + 1]), xl[ci0 + c + 1]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < kCI; ++c) {
            float val = 0.f;
            if (ok && ci0 + c < a.cin) {
              float f = __half2float(p[c]);
              val = lrelu(fmaf(f, xs[ci0 + c], xh[ci0 + c]), xl[ci0 + c]);
            }
            in[v][c] = val;
          }
        }
      }
      const float4* wt = reinterpret_cast<const float4*>(w_s + (size_t)t * kCI * kCO);
#pragma unroll
      for (int c = 0; c < kCI; ++c) {
        float4 w0 = wt[c * 4 + 0], w1 = wt[c * 4 + 1], w2 = wt[c * 4 + 2], w3 = wt[c * 4 + 3];
        float wv[kCO] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                         w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int v = 0; v < kVX; ++v)
#pragma unroll
          for (int o = 0; o < kCO; ++o) acc[v][o] = fmaf(in[v][c], wv[o], acc[v][o]);
      }
    }
  }

  // epilogue: bias, round to fp16, store, InstanceNorm sums of the rounded values
  float s1[kCO], s2[kCO];
#pragma unroll
  for (int o = 0; o < kCO; ++o) s1[o] = s2[o] = 0.f;
  if (active) {
    int oz, oy, oxb, oxstep;
    if (a.transposed) {
      int dx = tap_fixed % a.s[2];
      int dy = (tap_fixed / a.s[2]) % a.s[1];
      int dz = tap_fixed / (a.s[2] * a.s[1]);
      oz = z * a.s[0] + dz; oy = y * a.s[1] + dy; oxb = x0 * a.s[2] + dx; oxstep = a.s[2];
    } else {
      oz = z; oy = y; oxb = x0; oxstep = 1;
    }
    __half* dst_b = a.dst + (size_t)b * a.out_d[0] * a.out_d[1] * a.out_d[2] * a.dst_cs;
    const bool vec_out = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0) && (co0 + kCO <= a.cout);
#pragma unroll
    for (int v = 0; v < kVX; ++v) {
      if (x0 + v >= D2) continue;
      int ox = oxb + v * oxstep;
      __half* q = dst_b + (((size_t)oz * a.out_d[1] + oy) * a.out_d[2] + ox) * a.dst_cs + co0;
      __half hv[kCO];
#pragma unroll
      for (int o = 0; o < kCO; ++o) {
        float val = acc[v][o] + ((a.bias && co0 + o < a.cout) ? __ldg(a.bias + co0 + o) : 0.f);
        hv[o] = __float2half_rn(val);
        float r = __half2float(hv[o]);
        s1[o] += r;
        s2[o] += r * r;
      }
      if (vec_out) {
        reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
        reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[8]);
      } else {
#pragma unroll
        for (int o = 0; o < kCO; ++o)
          if (co0 + o < a.cout) q[o] = hv[o];
      }
    }
  }
  if (a.dst_stats) {
    __syncthreads();
#pragma unroll
    for (int o = 0; o < kCO; ++o) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s1[o] += __shfl_xor_sync(0xffffffffu, s1[o], off);
        s2[o] += __shfl_xor_sync(0xffffffffu, s2[o], off);
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < kCO; ++o) {
        red[warp * 2 * kCO + o] = s1[o];
        red[warp * 2 * kCO + kCO + o] = s2[o];
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kCO) {
      int o = threadIdx.x % kCO;
      int which = threadIdx.x / kCO;
      double tot = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) tot += (double)red[w * 2 * kCO + which * kCO + o];
      if (co0 + o < a.cout) atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + co0 + o) * 2 + which, tot);
    }
  }
}

int launch_conv_direct(const ConvArgs& a, cudaStream_t s) {
  const int D0 = a.transposed ? a.in_d[0] : a.out_d[0];
  const int D1 = a.transposed ? a.in_d[1] : a.out_d[1];
  const int D2 = a.transposed ? a.in_d[2] : a.out_d[2];
  long long groups = (long long)D0 * D1 * ((D2 + kVX - 1) / kVX);
  dim3 grid((unsigned)((groups + kThreads - 1) / kThreads),
            (unsigned)((a.cout_pad / kCO) * (a.transposed ? a.ntaps : 1)), (unsigned)a.batch);
  int ntaps_blk = a.transposed ? 1 : a.ntaps;
  size_t smem = (size_t)(ntaps_blk * kCI * kCO + 3 * a.cin + (kThreads / 32) * 2 * kCO) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv_direct: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return FNNU_E_CUDA;
    }
  }
  conv_direct_kernel<<<grid, kThreads, smem, s>>>(a);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

// ------------------------------------------------------------------------------------------------
// First layer: Cin <= 4 (1 CT channel, 4 MRI channels), stride 1, kernel 1|3 per axis, Cout <= 64.
// K = taps * Cin is too thin for the tensor cores (0.8 % of the network's FLOPs), so this is a
// CUDA-core kernel: weights broadcast from shared memory, each thread owns 4 consecutive x voxels and
// 16 output channels, each loaded input row segment is reused by the 3 kx taps.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(128) conv_small_cin_kernel(ConvArgs a) {
  extern __shared__ float smem[];
  float* w_s = smem;                                    // [ntaps][CIN][16]
  float* red = w_s + a.ntaps * CIN * 16;                // [4][32]
  const int b = blockIdx.z;
  const int co0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < a.ntaps * CIN * 16; i += 128) {
    int o = i & 15, rest = i >> 4;                      // rest = tap * CIN + ci
    w_s[i] = __ldg(a.w + (size_t)rest * a.cout_pad + co0 + o);
  }
  __syncthreads();
  float xs[CIN], xh[CIN], xl[CIN];     // pending transform of the source (identity for the network input)
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const ChanMeta m = a.src_meta[c];
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, m, a.src_inv_count, xs[c], xh[c]);
    xl[c] = m.eps < 0.f ? 1.f : m.slope;
  }
  const int D0 = a.out_d[0], D1 = a.out_d[1], D2 = a.out_d[2];
  const int gpr = (D2 + 3) / 4;
  const long long n_groups = (long long)D0 * D1 * gpr;
  const long long gidx = (long long)blockIdx.x * 128 + threadIdx.x;
  const bool active = gidx < n_groups;
  // accumulators as (o, o+1) pairs: Blackwell's packed fp32 FMA (fma.rn.f32x2) does two FMAs per instruction
  unsigned long long acc2[4][8];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc2[v][o] = 0ull;
  int x0 = 0, y = 0, z = 0;
  if (active) {
    x0 = (int)(gidx % gpr) * 4;
    y = (int)((gidx / gpr) % D1);
    z = (int)(gidx / ((long long)gpr * D1));
    const __half* src_b = a.src + (size_t)b * D0 * D1 * D2 * a.src_cs;
    const int nx = a.k[2] + 3;                           // input columns needed for 4 outputs
    unsigned xok = 0;                                    // bit j: input column j of this thread lies inside the image
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int ix = x0 + j - a.pad[2];
      if (j < nx && ix >= 0 && ix < D2) xok |= 1u << j;
    }
    // the network input has no pending transform: skip the per-load scale / shift / LeakyReLU there (the kernel is
    // issue-bound: 68 % of the issue slots at 25 % occupancy)
    bool ident = true;
#pragma unroll
    for (int c = 0; c < CIN; ++c) ident = ident && (a.src_meta[c].eps < 0.f);
    auto accumulate = [&](auto ident_tag) {
      constexpr bool IDENT = decltype(ident_tag)::value;
      for (int kz = 0; kz < a.k[0]; ++kz) {
        const int iz = z + kz - a.pad[0];
        if (iz < 0 || iz >= D0) continue;
        for (int ky = 0; ky < a.k[1]; ++ky) {
          const int iy = y + ky - a.pad[1];
          if (iy < 0 || iy >= D1) continue;
          const __half* row = src_b + ((size_t)iz * D1 + iy) * D2 * a.src_cs;
          float in[6][CIN];
  #pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int ix = x0 + j - a.pad[2];
            const bool ok = (xok >> j) & 1u;
  #pragma unroll
            for (int c = 0; c < CIN; ++c) {
              const float raw = ok ? __half2float(__ldg(row + (size_t)ix * a.src_cs + c)) : 0.f;
              in[j][c] = (IDENT || !ok) ? raw : lrelu(fmaf(raw, xs[c], xh[c]), xl[c]);
            }
          }
  #pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (kx >= a.k[2]) break;
            const int tap = (kz * a.k[1] + ky) * a.k[2] + kx;
  #pragma unroll
            for (int c = 0; c < CIN; ++c) {
              const ulonglong2* wt = reinterpret_cast<const ulonglong2*>(w_s + ((size_t)tap * CIN + c) * 16);
              const ulonglong2 w01 = wt[0], w23 = wt[1], w45 = wt[2], w67 = wt[3];
              const unsigned long long wp[8] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y};
  #pragma unroll
              for (int v = 0; v < 4; ++v) {
                const unsigned int ib = __float_as_uint(in[v + kx][c]);
                unsigned long long a2;
                asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "r"(ib));
  #pragma unroll
                for (int o = 0; o < 8; ++o)
                  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[v][o]) : "l"(a2), "l"(wp[o]));
              }
            }
          }
        }
      }
    };
    if (ident) accumulate(std::true_type{}); else accumulate(std::false_type{});
  }
  float s1[16], s2[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) s1[o] = s2[o] = 0.f;
  if (active) {
    __half* dst_b = a.dst + (size_t)b * D0 * D1 * D2 * a.dst_cs;
    const bool vec_out = (a.dst_cs % 8 == 0) && (((uintptr_t)a.dst) % 16 == 0) && (co0 + 16 <= a.cout);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (x0 + v >= D2) continue;
      __half* q = dst_b + (((size_t)z * D1 + y) * D2 + x0 + v) * a.dst_cs + co0;
      __half hv[16];
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const float accv = __uint_as_float((o & 1) ? (unsigned int)(acc2[v][o >> 1] >> 32) : (unsigned int)(acc2[v][o >> 1]));
        float val = accv + ((a.bias && co0 + o < a.cout) ? __ldg(a.bias + co0 + o) : 0.f);
        hv[o] = __float2half_rn(val);
        float r = __half2float(hv[o]);
        s1[o] += r;
        s2[o] += r * r;
      }
      if (vec_out) {
        reinterpret_cast<uint4*>(q)[0] = *reinterpret_cast<uint4*>(&hv[0]);
        reinterpret_cast<uint4*>(q)[1] = *reinterpret_cast<uint4*>(&hv[8]);
      } else {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (co0 + o < a.cout) q[o] = hv[o];
      }
    }
  }
  if (a.dst_stats) {
#pragma unroll
    for (int o = 0; o < 16; ++o) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s1[o] += __shfl_xor_sync(0xffffffffu, s1[o], off);
        s2[o] += __shfl_xor_sync(0xffffffffu, s2[o], off);
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        red[warp * 32 + o] = s1[o];
        red[warp * 32 + 16 + o] = s2[o];
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      int o = threadIdx.x & 15, which = threadIdx.x >> 4;
      double tot = 0.0;
      for (int w = 0; w < 4; ++w) tot += (double)red[w * 32 + which * 16 + o];
      if (co0 + o < a.cout) atomicAdd(a.dst_stats + ((size_t)b * a.dst_stat_stride + co0 + o) * 2 + which, tot);
    }
  }
}

static bool small_cin_ok(const ConvArgs& a) {
  if (a.transposed || a.cin > 4) return false;
  for (int i = 0; i < 3; ++i)
    if (a.s[i] != 1 || (a.k[i] != 1 && a.k[i] != 3)) return false;
  return a.cout_pad <= 64;
}

// ------------------------------------------------------------------------------------------------
// Segmentation head: 1x1x1 conv to <= 8 heads.  HBM-bound (reads Cin fp16 per voxel, writes the logits):
// one thread per voxel, 128-bit loads, source transform applied on the fly, no InstanceNorm sums.
// ------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256) conv_pointwise_head_kernel(ConvArgs a) {
  extern __shared__ float smem[];
  float* w_s = smem;                 // [COUT][cin]
  float* xs = w_s + COUT * a.cin;    // scale, shift, slope [cin] per sample (recomputed when b changes)
  float* xh = xs + a.cin;
  float* xl = xh + a.cin;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < COUT * a.cin; i += 256) {
    int co = i / a.cin, ci = i - co * a.cin;
    w_s[i] = co < a.cout ? __ldg(a.w + (size_t)ci * a.cout_pad + co) : 0.f;     // packed [tap=0][cin][cout_pad]
  }
  for (int c = threadIdx.x; c < a.cin; c += 256) {
    float sc, sh;
    const ChanMeta m = a.src_meta[c];
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, m, a.src_inv_count, sc, sh);
    xs[c] = sc;
    xh[c] = sh;
    xl[c] = m.eps < 0.f ? 1.f : m.slope;
  }
  __syncthreads();
  const size_t nv = (size_t)a.in_d[0] * a.in_d[1] * a.in_d[2];
  const __half* src_b = a.src + (size_t)b * nv * a.src_cs;
  __half* dst_b = a.dst + (size_t)b * nv * a.dst_cs;
  float bias[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) bias[o] = (a.bias && o < a.cout) ? __ldg(a.bias + o) : 0.f;
  const bool vec_out = a.dst_cs >= COUT && (a.dst_cs * 2) % (COUT * 2) == 0 && ((uintptr_t)a.dst % (COUT * 2)) == 0;
  for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nv; v += (size_t)gridDim.x * 256) {
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = bias[o];
    const uint4* p = reinterpret_cast<const uint4*>(src_b + v * a.src_cs);
    for (int c8 = 0; c8 < a.cin; c8 += 8) {
      const uint4 raw = __ldg(p + (c8 >> 3));
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(h2[e]);
        const int c = c8 + 2 * e;
        float v0 = fmaf(f.x, xs[c], xh[c]);
        float v1 = fmaf(f.y, xs[c + 1], xh[c + 1]);
        v0 = fmaxf(v0, v0 * xl[c]);
        v1 = fmaxf(v1, v1 * xl[c + 1]);
        // the tensor-core layers feed fp16-rounded activations to the MMA; round here too so that all
        // back ends see the same operand values
        v0 = __half2float(__float2half_rn(v0));
        v1 = __half2float(__float2half_rn(v1));
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v0, w_s[o * a.cin + c], fmaf(v1, w_s[o * a.cin + c + 1], acc[o]));
      }
    }
    __half* q = dst_b + v * a.dst_cs;
    if (vec_out) {
      // the destination holds COUT (padded) heads per voxel: one 4 / 8 / 16-byte store (padding heads are zeros)
      __half2 h[COUT / 2];
#pragma unroll
      for (int o = 0; o < COUT; o += 2) h[o >> 1] = __floats2half2_rn(acc[o], acc[o + 1]);
      if constexpr (COUT == 2) *reinterpret_cast<__half2*>(q) = h[0];
      else if constexpr (COUT == 4) *reinterpret_cast<uint2*>(q) = *reinterpret_cast<uint2*>(h);
      else *reinterpret_cast<uint4*>(q) = *reinterpret_cast<uint4*>(h);
    } else {
#pragma unroll
      for (int o = 0; o < COUT; ++o)
        if (o < a.cout) q[o] = __float2half_rn(acc[o]);
    }
  }
}

// Segmentation head with many classes (9 .. 64 per pass, e.g. the 61 labels of bone_turbo): HBM-bound (32 bytes in,
// 2 x heads bytes out per voxel), 2 x Cin x heads FLOP per voxel on the CUDA cores.  One thread per voxel keeps 64
// fp32 accumulators in registers; the weights sit in shared memory as [cin][64] so that one LDS.128 broadcast feeds four
// FMAs; the padded head vector of a voxel leaves as 16-byte stores.  Heads beyond 64 take further passes.
__global__ void __launch_bounds__(128) conv_pointwise_head_wide_kernel(ConvArgs a) {
  extern __shared__ float smem[];
  float* w_s = smem;                  // [cin][64] of the current pass
  float* xs = w_s + 64 * a.cin;       // scale, shift, slope [cin]
  float* xh = xs + a.cin;
  float* xl = xh + a.cin;
  float* b_s = xl + a.cin;            // [64] bias of the current pass
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < a.cin; c += 128) {
    float sc, sh;
    const ChanMeta m = a.src_meta[c];
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, m, a.src_inv_count, sc, sh);
    xs[c] = sc;
    xh[c] = sh;
    xl[c] = m.eps < 0.f ? 1.f : m.slope;
  }
  const size_t nv = (size_t)a.in_d[0] * a.in_d[1] * a.in_d[2];
  const __half* src_b = a.src + (size_t)b * nv * a.src_cs;
  __half* dst_b = a.dst + (size_t)b * nv * a.dst_cs;
  for (int co0 = 0; co0 < a.cout; co0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * a.cin; i += 128) {
      const int ci = i >> 6, co = co0 + (i & 63);
      w_s[i] = co < a.cout ? __ldg(a.w + (size_t)ci * a.cout_pad + co) : 0.f;       // packed [tap = 0][cin][cout_pad]
    }
    if (threadIdx.x < 64) b_s[threadIdx.x] = (a.bias && co0 + (int)threadIdx.x < a.cout) ? __ldg(a.bias + co0 + threadIdx.x) : 0.f;
    __syncthreads();
    // the store is vectorised when the destination row of a voxel holds the 64 (padded) heads of this pass
    const bool vec_out = a.dst_cs % 8 == 0 && co0 + 64 <= a.dst_cs && ((uintptr_t)a.dst % 16) == 0;
    for (size_t v = (size_t)blockIdx.x * 128 + threadIdx.x; v < nv; v += (size_t)gridDim.x * 128) {
      float acc[64];
#pragma unroll
      for (int o4 = 0; o4 < 16; ++o4) {
        const float4 b4 = reinterpret_cast<const float4*>(b_s)[o4];
        acc[4 * o4] = b4.x; acc[4 * o4 + 1] = b4.y; acc[4 * o4 + 2] = b4.z; acc[4 * o4 + 3] = b4.w;
      }
      const uint4* p = reinterpret_cast<const uint4*>(src_b + v * a.src_cs);
      for (int c8 = 0; c8 < a.cin; c8 += 8) {
        const uint4 raw = __ldg(p + (c8 >> 3));
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 f = __half22float2(h2[e >> 1]);
          const int c = c8 + e;
          float x = fmaf((e & 1) ? f.y : f.x, xs[c], xh[c]);
          x = fmaxf(x, x * xl[c]);
          x = __half2float(__float2half_rn(x));      // every back end feeds fp16-rounded activations to its conv
          const float4* wr = reinterpret_cast<const float4*>(w_s + c * 64);
#pragma unroll
          for (int o4 = 0; o4 < 16; ++o4) {
            const float4 w4 = wr[o4];
            acc[4 * o4 + 0] = fmaf(x, w4.x, acc[4 * o4 + 0]);
            acc[4 * o4 + 1] = fmaf(x, w4.y, acc[4 * o4 + 1]);
            acc[4 * o4 + 2] = fmaf(x, w4.z, acc[4 * o4 + 2]);
            acc[4 * o4 + 3] = fmaf(x, w4.w, acc[4 * o4 + 3]);
          }
        }
      }
      __half* q = dst_b + v * a.dst_cs + co0;
      if (vec_out) {
#pragma unroll
        for (int o8 = 0; o8 < 8; ++o8) {
          __half2 h[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int o = 8 * o8 + 2 * k;
            h[k] = __floats2half2_rn(acc[o], acc[o + 1]);
          }
          reinterpret_cast<uint4*>(q)[o8] = *reinterpret_cast<uint4*>(h);
        }
      } else {
#pragma unroll
        for (int o = 0; o < 64; ++o)
          if (co0 + o < a.cout) q[o] = __float2half_rn(acc[o]);
      }
    }
  }
}

static bool pointwise_head_wide_ok(const ConvArgs& a) {
  if (a.transposed || a.dst_stats) return false;
  if (a.k[0] != 1 || a.k[1] != 1 || a.k[2] != 1 || a.s[0] != 1 || a.s[1] != 1 || a.s[2] != 1) return false;
  return a.cout > 8 && a.cin % 8 == 0 && a.cin <= 64 && a.src_cs % 8 == 0 && ((uintptr_t)a.src % 16) == 0;
}

static bool pointwise_head_ok(const ConvArgs& a) {
  if (a.transposed || a.dst_stats) return false;
  if (a.k[0] != 1 || a.k[1] != 1 || a.k[2] != 1 || a.s[0] != 1 || a.s[1] != 1 || a.s[2] != 1) return false;
  return a.cout <= 8 && a.cin % 8 == 0 && a.cin <= 1024 && a.src_cs % 8 == 0 && ((uintptr_t)a.src % 16) == 0;
}

bool direct_specialised(const ConvArgs& a) { return small_cin_ok(a) || pointwise_head_ok(a) || pointwise_head_wide_ok(a); }
bool prefer_cuda_cores(const ConvArgs& a) { return pointwise_head_ok(a) || pointwise_head_wide_ok(a); }

int launch_conv_specialised(const ConvArgs& a, cudaStream_t s) {
  if (pointwise_head_wide_ok(a)) {
    const size_t nv = (size_t)a.in_d[0] * a.in_d[1] * a.in_d[2];
    int blocks = (int)((nv + 127) / 128);
    const int cap = num_sms() * 16 / (a.batch > 8 ? 8 : a.batch) + 1;
    if (blocks > cap) blocks = cap;
    const dim3 grid((unsigned)blocks, (unsigned)a.batch);
    const size_t smem = (size_t)(64 * a.cin + 3 * a.cin + 64) * sizeof(float);
    conv_pointwise_head_wide_kernel<<<grid, 128, smem, s>>>(a);
    FNNU_LAUNCH_CHECK();
    return FNNU_OK;
  }
  if (pointwise_head_ok(a)) {
    const size_t nv = (size_t)a.in_d[0] * a.in_d[1] * a.in_d[2];
    int blocks = (int)((nv + 255) / 256);
    int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)a.batch);
    const int co = a.cout <= 2 ? 2 : (a.cout <= 4 ? 4 : 8);
    size_t smem = (size_t)(co * a.cin + 3 * a.cin) * sizeof(float);
    if (co == 2) conv_pointwise_head_kernel<2><<<grid, 256, smem, s>>>(a);
    else if (co == 4) conv_pointwise_head_kernel<4><<<grid, 256, smem, s>>>(a);
    else conv_pointwise_head_kernel<8><<<grid, 256, smem, s>>>(a);
    FNNU_LAUNCH_CHECK();
    return FNNU_OK;
  }
  if (small_cin_ok(a)) {
    // Cin == 1 first layers run on the tensor cores (conv_first_zpair.cu, conv_first_umma.cu); FNNU_FIRST_LAYER_TC=0 selects the CUDA-core kernel
    static const bool first_tc = [] { const char* e = getenv("FNNU_FIRST_LAYER_TC"); return !(e && e[0] == '0'); }();
    if (first_tc && first_zpair_supported(a)) return launch_conv_first_zpair(a, s);
    if (first_tc && first_umma_supported(a)) return launch_conv_first_umma(a, s);
    long long groups = (long long)a.out_d[0] * a.out_d[1] * ((a.out_d[2] + 3) / 4);
    dim3 grid((unsigned)((groups + 127) / 128), (unsigned)(a.cout_pad / 16), (unsigned)a.batch);
    size_t smem = (size_t)(a.ntaps * a.cin * 16 + 4 * 32) * sizeof(float);
    switch (a.cin) {
      case 1: conv_small_cin_kernel<1><<<grid, 128, smem, s>>>(a); break;
      case 2: conv_small_cin_kernel<2><<<grid, 128, smem, s>>>(a); break;
      case 3: conv_small_cin_kernel<3><<<grid, 128, smem, s>>>(a); break;
      default: conv_small_cin_kernel<4><<<grid, 128, smem, s>>>(a); break;
    }
    FNNU_LAUNCH_CHECK();
    return FNNU_OK;
  }
  set_error("conv_specialised: unsupported shape");
  return FNNU_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// weight packing: PyTorch layouts -> [tap][cin][cout_pad] fp32
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_direct_kernel(const float* __restrict__ w, float* __restrict__ out, int cin, int cout,
                                           int cout_pad, int ntaps, int transposed) {
  size_t total = (size_t)ntaps * cin * cout_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(i % cout_pad);
    int ci = (int)((i / cout_pad) % cin);
    int tap = (int)(i / ((size_t)cout_pad * cin));
    float v = 0.f;
    if (co < cout) {
      size_t src = transposed ? (((size_t)ci * cout + co) * ntaps + tap) : (((size_t)co * cin + ci) * ntaps + tap);
      v = w[src];
    }
    out[i] = v;
  }
}

int launch_pack_weights_direct(const float* w_dev, float* out, int cin, int cout, int cout_pad, int ntaps,
                               int transposed, cudaStream_t s) {
  size_t total = (size_t)ntaps * cin * cout_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weights_direct_kernel<<<blocks, 256, 0, s>>>(w_dev, out, cin, cout, cout_pad, ntaps, transposed);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

// ------------------------------------------------------------------------------------------------
// residual join: dst = lrelu_slope(xform(a) + xform(b)), stored ready to use
// average pooling: dst = mean over the stride window of xform(src), stored ready to use
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_act_scalar_kernel(EltArgs a) {
  const size_t nv = (size_t)a.d[0] * a.d[1] * a.d[2];
  const size_t total = nv * a.c * a.batch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % a.c);
    size_t v = (i / a.c) % nv;
    int b = (int)(i / ((size_t)a.c * nv));
    float sc, sh, sc2, sh2;
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, a.src_meta[c], a.inv_count, sc, sh);
    xform_from_stats(a.src2_stats + ((size_t)b * a.src2_stat_stride + c) * 2, a.src2_meta[c], a.inv_count, sc2, sh2);
    float sl1 = a.src_meta[c].eps < 0.f ? 1.f : a.src_meta[c].slope;
    float sl2 = a.src2_meta[c].eps < 0.f ? 1.f : a.src2_meta[c].slope;
    float x1 = lrelu(fmaf(__half2float(a.src[((size_t)b * nv + v) * a.src_cs + c]), sc, sh), sl1);
    float x2 = lrelu(fmaf(__half2float(a.src2[((size_t)b * nv + v) * a.src2_cs + c]), sc2, sh2), sl2);
    a.dst[((size_t)b * nv + v) * a.dst_cs + c] = __float2half_rn(lrelu(x1 + x2, a.slope));
  }
}

__global__ void __launch_bounds__(256) avgpool_scalar_kernel(EltArgs a) {
  // a.d = OUTPUT dims; input dims = a.d * a.s
  const size_t nv = (size_t)a.d[0] * a.d[1] * a.d[2];
  const size_t total = nv * a.c * a.batch;
  const int I1 = a.d[1] * a.s[1], I2 = a.d[2] * a.s[2];
  const size_t nvi = nv * a.s[0] * a.s[1] * a.s[2];
  const float inv = 1.f / (float)(a.s[0] * a.s[1] * a.s[2]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % a.c);
    size_t v = (i / a.c) % nv;
    int b = (int)(i / ((size_t)a.c * nv));
    int x = (int)(v % a.d[2]);
    int y = (int)((v / a.d[2]) % a.d[1]);
    int z = (int)(v / ((size_t)a.d[2] * a.d[1]));
    float sc, sh;
    xform_from_stats(a.src_stats + ((size_t)b * a.src_stat_stride + c) * 2, a.src_meta[c], a.inv_count, sc, sh);
    float sl = a.src_meta[c].eps < 0.f ? 1.f : a.src_meta[c].slope;
    float sum = 0.f;
    for (int dz = 0; dz < a.s[0]; ++dz)
      for (int dy = 0; dy < a.s[1]; ++dy)
        for (int dx = 0; dx < a.s[2]; ++dx) {
          size_t iv = ((size_t)(z * a.s[0] + dz) * I1 + (y * a.s[1] + dy)) * I2 + (x * a.s[2] + dx);
          sum += lrelu(fmaf(__half2float(a.src[((size_t)b * nvi + iv) * a.src_cs + c]), sc, sh), sl);
        }
    a.dst[((size_t)b * nv + v) * a.dst_cs + c] = __float2half_rn(sum * inv);
  }
}

// Both kernels: blockIdx.y = sample; the per-channel transforms are evaluated ONCE per block into shared memory
// (they involve fp64 mean / rsqrt), then each thread streams 8 channels of a voxel with 128-bit accesses.
__device__ __forceinline__ void elt_tables(const EltArgs& a, int b, float* t, bool second) {
  // t: [scale | shift | slope][c]
  const double* st = second ? a.src2_stats : a.src_stats;
  const ChanMeta* mt = second ? a.src2_meta : a.src_meta;
  const int stride = second ? a.src2_stat_stride : a.src_stat_stride;
  for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
    float sc, sh;
    const ChanMeta m = mt[c];
    xform_from_stats(st + ((size_t)b * stride + c) * 2, m, a.inv_count, sc, sh);
    t[c] = sc;
    t[a.c + c] = sh;
    t[2 * a.c + c] = m.eps < 0.f ? 1.f : m.slope;
  }
}

__global__ void __launch_bounds__(256) add_act_kernel(EltArgs a) {
  extern __shared__ float tab[];            // [2 sources][3][c]
  const int b = blockIdx.y;
  float* t1 = tab;
  float* t2 = tab + 3 * a.c;
  elt_tables(a, b, t1, false);
  elt_tables(a, b, t2, true);
  __syncthreads();
  const size_t nv = (size_t)a.d[0] * a.d[1] * a.d[2];
  const int groups = a.c >> 3;
  const size_t total = nv * groups;
  const __half* s1 = a.src + (size_t)b * nv * a.src_cs;
  const __half* s2 = a.src2 + (size_t)b * nv * a.src2_cs;
  __half* d = a.dst + (size_t)b * nv * a.dst_cs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t v = i / groups;
    const int c0 = g * 8;
    const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(s1 + v * a.src_cs + c0));
    const uint4 r2 = __ldg(reinterpret_cast<const uint4*>(s2 + v * a.src2_cs + c0));
    const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
    const __half2* h2 = reinterpret_cast<const __half2*>(&r2);
    __half2 o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f1 = __half22float2(h1[e]), f2 = __half22float2(h2[e]);
      const int c = c0 + 2 * e;
      const float x0 = lrelu(fmaf(f1.x, t1[c], t1[a.c + c]), t1[2 * a.c + c]) + lrelu(fmaf(f2.x, t2[c], t2[a.c + c]), t2[2 * a.c + c]);
      const float x1 = lrelu(fmaf(f1.y, t1[c + 1], t1[a.c + c + 1]), t1[2 * a.c + c + 1]) +
                       lrelu(fmaf(f2.y, t2[c + 1], t2[a.c + c + 1]), t2[2 * a.c + c + 1]);
      o[e] = __floats2half2_rn(lrelu(x0, a.slope), lrelu(x1, a.slope));
    }
    *reinterpret_cast<uint4*>(d + v * a.dst_cs + c0) = *reinterpret_cast<uint4*>(o);
  }
}

__global__ void __launch_bounds__(256) avgpool_kernel(EltArgs a) {
  // a.d = OUTPUT dims; input dims = a.d * a.s
  extern __shared__ float tab[];
  const int b = blockIdx.y;
  elt_tables(a, b, tab, false);
  __syncthreads();
  const size_t nv = (size_t)a.d[0] * a.d[1] * a.d[2];
  const int I1 = a.d[1] * a.s[1], I2 = a.d[2] * a.s[2];
  const size_t nvi = nv * a.s[0] * a.s[1] * a.s[2];
  const float inv = 1.f / (float)(a.s[0] * a.s[1] * a.s[2]);
  const int groups = a.c >> 3;
  const size_t total = nv * groups;
  const __half* src = a.src + (size_t)b * nvi * a.src_cs;
  __half* d = a.dst + (size_t)b * nv * a.dst_cs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t v = i / groups;
    const int c0 = g * 8;
    const int x = (int)(v % a.d[2]);
    const int y = (int)((v / a.d[2]) % a.d[1]);
    const int z = (int)(v / ((size_t)a.d[2] * a.d[1]));
    float sum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sum[e] = 0.f;
    for (int dz = 0; dz < a.s[0]; ++dz)
      for (int dy = 0; dy < a.s[1]; ++dy)
        for (int dx = 0; dx < a.s[2]; ++dx) {
          const size_t iv = ((size_t)(z * a.s[0] + dz) * I1 + (y * a.s[1] + dy)) * I2 + (x * a.s[2] + dx);
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(src + iv * a.src_cs + c0));
          const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            const int c = c0 + 2 * e;
            sum[2 * e] += lrelu(fmaf(f.x, tab[c], tab[a.c + c]), tab[2 * a.c + c]);
            sum[2 * e + 1] += lrelu(fmaf(f.y, tab[c + 1], tab[a.c + c + 1]), tab[2 * a.c + c + 1]);
          }
        }
    __half2 o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __floats2half2_rn(sum[2 * e] * inv, sum[2 * e + 1] * inv);
    *reinterpret_cast<uint4*>(d + v * a.dst_cs + c0) = *reinterpret_cast<uint4*>(o);
  }
}

static bool elt_vec_ok(const EltArgs& a) {
  return a.c % 8 == 0 && a.src_cs % 8 == 0 && a.dst_cs % 8 == 0 && ((uintptr_t)a.src % 16) == 0 && ((uintptr_t)a.dst % 16) == 0 &&
         (!a.src2 || (a.src2_cs % 8 == 0 && ((uintptr_t)a.src2 % 16) == 0)) && a.c <= 2048;
}

int launch_add_act(const EltArgs& a, cudaStream_t s) {
  size_t total = (size_t)a.d[0] * a.d[1] * a.d[2] * a.c * a.batch;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (elt_vec_ok(a)) {
    size_t work = (size_t)a.d[0] * a.d[1] * a.d[2] * (a.c / 8);
    int bx = (int)((work + 255) / 256);
    if (bx > num_sms() * 8) bx = num_sms() * 8;
    add_act_kernel<<<dim3((unsigned)bx, (unsigned)a.batch), 256, (size_t)6 * a.c * sizeof(float), s>>>(a);
    FNNU_LAUNCH_CHECK();
    return FNNU_OK;
  }
  add_act_scalar_kernel<<<blocks, 256, 0, s>>>(a);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

int launch_avgpool(const EltArgs& a, cudaStream_t s) {
  size_t total = (size_t)a.d[0] * a.d[1] * a.d[2] * a.c * a.batch;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (elt_vec_ok(a)) {
    size_t work = (size_t)a.d[0] * a.d[1] * a.d[2] * (a.c / 8);
    int bx = (int)((work + 255) / 256);
    if (bx > num_sms() * 8) bx = num_sms() * 8;
    avgpool_kernel<<<dim3((unsigned)bx, (unsigned)a.batch), 256, (size_t)3 * a.c * sizeof(float), s>>>(a);
    FNNU_LAUNCH_CHECK();
    return FNNU_OK;
  }
  avgpool_scalar_kernel<<<blocks, 256, 0, s>>>(a);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

}  // namespace fnnu
