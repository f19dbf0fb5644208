// Pre-processing of one case on the device: everything between the raw image array and the
// network-ready tensor of predict_from_raw_data.py:423-468, i.e. DefaultPreprocessor.run_case_npy
// (preprocessing/preprocessors/default_preprocessor.py:45-118) for a test case:
//   transpose_forward -> crop_to_nonzero (cropping/cropping.py:8-39, incl. binary_fill_holes for the normalisation
//   mask) -> per-channel intensity normalisation (normalization/default_normalization_schemes.py:27-95) -> resampling
//   to the plans' spacing (resampling/default_resampling.py:111-192: cubic spline, nearest along the anisotropic axis).
// skimage.transform.resize(order=3, mode='edge', anti_aliasing=False, clip=True) is scipy.ndimage.zoom(order=3,
// mode='nearest', grid_mode=True) + clipping to the input's min / max.  scipy's zoom: pad 12 edge samples, cubic
// B-spline prefilter per axis (pole sqrt(3) - 2, mirror initialisation, float64), then 4 taps per axis at
// s = (o + 0.5) * in / out - 0.5.  All kernels are memory-bound; the transpose and the crop are index arithmetic inside
// the kernels (no copies), float64 where scipy computes in float64.
#include <limits.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "ops.cuh"

namespace fnnu {

struct View3 {          // raw image [C][d0][d1][d2] (original axis order) seen through transpose_forward
  int d[3];             // original extents
  int tf[3];            // transposed axis k = original axis tf[k]
  int td[3];            // transposed extents
};

__device__ __forceinline__ size_t raw_index(const View3& v, int t0, int t1, int t2) {
  int o[3];
  o[v.tf[0]] = t0;
  o[v.tf[1]] = t1;
  o[v.tf[2]] = t2;
  return ((size_t)o[0] * v.d[1] + o[1]) * v.d[2] + o[2];
}

// ---- bounding box of the voxels where any channel is non-zero (transposed coordinates, hi exclusive) ----
__global__ void __launch_bounds__(256) nonzero_bbox_kernel(const float* __restrict__ img, int C, View3 v, int* __restrict__ bbox) {
  const size_t nvox = (size_t)v.d[0] * v.d[1] * v.d[2];
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {0, 0, 0};
  int inv[3];
  inv[v.tf[0]] = 0; inv[v.tf[1]] = 1; inv[v.tf[2]] = 2;        // original axis -> transposed axis
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (size_t)gridDim.x * blockDim.x) {
    bool nz = false;
    for (int c = 0; c < C; ++c) nz = nz || (img[(size_t)c * nvox + i] != 0.f);
    if (nz) {
      int o[3];
      o[2] = (int)(i % v.d[2]);
      o[1] = (int)((i / v.d[2]) % v.d[1]);
      o[0] = (int)(i / ((size_t)v.d[2] * v.d[1]));
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int t = inv[k];
        lo[t] = min(lo[t], o[k]);
        hi[t] = max(hi[t], o[k] + 1);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
    if ((threadIdx.x & 31) == 0) {
      if (lo[k] != INT_MAX) atomicMin(&bbox[2 * k], lo[k]);
      if (hi[k] != 0) atomicMax(&bbox[2 * k + 1], hi[k]);
    }
  }
}

// ---- fill-holes state of the cropped mask: 0 = non-zero voxel, 1 = zero voxel, 2 = zero voxel connected to outside ----
__global__ void __launch_bounds__(256) mask_state_kernel(const float* __restrict__ img, int C, View3 v, int l0, int l1, int l2,
                                                         int c0, int c1, int c2, uint8_t* __restrict__ state) {
  const size_t nvox = (size_t)v.d[0] * v.d[1] * v.d[2];
  const size_t total = (size_t)c0 * c1 * c2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % c2), y = (int)((i / c2) % c1), x = (int)(i / ((size_t)c2 * c1));
    const size_t r = raw_index(v, x + l0, y + l1, z + l2);
    bool nz = false;
    for (int c = 0; c < C; ++c) nz = nz || (img[(size_t)c * nvox + r] != 0.f);
    const bool face = x == 0 || y == 0 || z == 0 || x == c0 - 1 || y == c1 - 1 || z == c2 - 1;
    state[i] = nz ? 0 : (face ? 2 : 1);
  }
}

// One sweep along `axis` (both directions): "outside" spreads through zero voxels (6-connectivity, the structure
// scipy.ndimage.binary_fill_holes uses by default).  One thread per line.
__global__ void __launch_bounds__(128) fill_sweep_kernel(uint8_t* __restrict__ state, int c0, int c1, int c2, int axis,
                                                         int* __restrict__ changed) {
  const int dims[3] = {c0, c1, c2};
  const size_t strides[3] = {(size_t)c1 * c2, (size_t)c2, 1};
  const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;     // the two other axes, a2 the faster one
  const size_t n_lines = (size_t)dims[a1] * dims[a2];
  const int n = dims[axis];
  const size_t st = strides[axis];
  bool any = false;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_lines; l += (size_t)gridDim.x * blockDim.x) {
    uint8_t* p = state + (l / dims[a2]) * strides[a1] + (l % dims[a2]) * strides[a2];
    bool carry = false;
    for (int i = 0; i < n; ++i) {
      const uint8_t s = p[i * st];
      if (s == 2) carry = true;
      else if (s == 1 && carry) { p[i * st] = 2; any = true; }
      else carry = false;
    }
    carry = false;
    for (int i = n - 1; i >= 0; --i) {
      const uint8_t s = p[i * st];
      if (s == 2) carry = true;
      else if (s == 1 && carry) { p[i * st] = 2; any = true; }
      else carry = false;
    }
  }
  if (any) *changed = 1;
}

// ---- per-channel statistics of the cropped channel (optionally inside the filled mask): sum, sum of squares, count,
// min, max ----
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ chan, View3 v, int l0, int l1, int l2,
                                                            int c0, int c1, int c2, const uint8_t* __restrict__ state,
                                                            double* __restrict__ out /* sum, sumsq, count */,
                                                            int* __restrict__ minmax /* ordered-int min, max */) {
  const size_t total = (size_t)c0 * c1 * c2;
  double s = 0.0, ss = 0.0, n = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (state && state[i] == 2) continue;
    const int z = (int)(i % c2), y = (int)((i / c2) % c1), x = (int)(i / ((size_t)c2 * c1));
    const float f = chan[raw_index(v, x + l0, y + l1, z + l2)];
    s += (double)f;
    ss += (double)f * (double)f;
    n += 1.0;
    mn = fminf(mn, f);
    mx = fmaxf(mx, f);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
    n += __shfl_xor_sync(0xffffffffu, n, off);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, s);
    atomicAdd(out + 1, ss);
    atomicAdd(out + 2, n);
    // order-preserving int encoding of a float
    auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
    if (mn <= mx) {
      atomicMin(minmax + 0, enc(mn));
      atomicMax(minmax + 1, enc(mx));
    }
  }
}

// ---- crop + transpose + normalise one channel: float32 arithmetic, operation by operation as numpy does it ----
// mode 0: copy; 1: (clip(x, lo, hi) - a) / b   (CTNormalization); 2: (x - a) / b   (z-score, rescale-to-01, RGB);
// 3: like 2 but only inside the mask (use_mask_for_norm)
__global__ void __launch_bounds__(256) crop_normalize_kernel(const float* __restrict__ chan, View3 v, int l0, int l1, int l2,
                                                             int c0, int c1, int c2, int mode, float a, float b, float lo,
                                                             float hi, const uint8_t* __restrict__ state,
                                                             float* __restrict__ out) {
  const size_t total = (size_t)c0 * c1 * c2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % c2), y = (int)((i / c2) % c1), x = (int)(i / ((size_t)c2 * c1));
    float f = chan[raw_index(v, x + l0, y + l1, z + l2)];
    if (mode == 1) {
      f = fminf(fmaxf(f, lo), hi);
      f = __fdiv_rn(__fsub_rn(f, a), b);
    } else if (mode == 2 || (mode == 3 && state[i] != 2)) {
      f = __fdiv_rn(__fsub_rn(f, a), b);
    }
    out[i] = f;
  }
}

// ---- cubic-spline resampling (scipy.ndimage.zoom order 3, mode nearest, grid_mode) ----
// pad the float32 channel with `pad[k]` edge samples per side into a float64 array
__global__ void __launch_bounds__(256) spline_pad_kernel(const float* __restrict__ src, int n0, int n1, int n2, int p0, int p1,
                                                         int p2, double* __restrict__ dst) {
  const int m0 = n0 + 2 * p0, m1 = n1 + 2 * p1, m2 = n2 + 2 * p2;
  const size_t total = (size_t)m0 * m1 * m2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int z = (int)(i % m2) - p2, y = (int)((i / m2) % m1) - p1, x = (int)(i / ((size_t)m2 * m1)) - p0;
    x = min(max(x, 0), n0 - 1);
    y = min(max(y, 0), n1 - 1);
    z = min(max(z, 0), n2 - 1);
    dst[i] = (double)src[((size_t)x * n1 + y) * n2 + z];
  }
}

// in-place cubic B-spline prefilter along one axis, one thread per line (ni_splines.c: gain, causal pass with the
// exact mirror initialisation, anti-causal pass)
__global__ void __launch_bounds__(128) spline_prefilter_kernel(double* __restrict__ data, int m0, int m1, int m2, int axis) {
  const int dims[3] = {m0, m1, m2};
  const size_t strides[3] = {(size_t)m1 * m2, (size_t)m2, 1};
  const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
  const size_t n_lines = (size_t)dims[a1] * dims[a2];
  const int n = dims[axis];
  const size_t st = strides[axis];
  const double z = -0.26794919243112270647;       // sqrt(3) - 2
  const double gain = (1.0 - z) * (1.0 - 1.0 / z);
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_lines; l += (size_t)gridDim.x * blockDim.x) {
    double* c = data + (l / dims[a2]) * strides[a1] + (l % dims[a2]) * strides[a2];
    if (n < 2) continue;
    // causal initialisation (mirror): c0 = (sum_{i} z^i (c[i] + z^(n-1) c[n-1-i])) / (1 - z^(2n-2)), everything times gain
    const double zn1 = pow(z, (double)(n - 1));
    double c0 = c[0] * gain + zn1 * (c[(size_t)(n - 1) * st] * gain);
    double zi = z;
    for (int i = 1; i < n - 1; ++i) {
      c0 += zi * (c[(size_t)i * st] * gain + zn1 * (c[(size_t)(n - 1 - i) * st] * gain));
      zi *= z;
      if (fabs(zi) < 1e-40) break;                 // beyond double precision: the remaining terms vanish
    }
    double prev = c0 / (1.0 - zn1 * zn1);
    c[0] = prev;
    for (int i = 1; i < n; ++i) {
      prev = c[(size_t)i * st] * gain + z * prev;
      c[(size_t)i * st] = prev;
    }
    double nxt = (z * c[(size_t)(n - 2) * st] + c[(size_t)(n - 1) * st]) * z / (z * z - 1.0);
    c[(size_t)(n - 1) * st] = nxt;
    for (int i = n - 2; i >= 0; --i) {
      nxt = z * (nxt - c[(size_t)i * st]);
      c[(size_t)i * st] = nxt;
    }
  }
}

struct ResampleArgs {
  const double* coef;       // padded coefficients (cubic axes pre-filtered) or the padded data itself (linear / nearest)
  int in_d[3], pad[3], out_d[3];
  int mode[3];              // 0 = cubic, 1 = nearest, 2 = linear
  double scale[3];          // in / out
  const float* clip;        // [n][2] min, max of the input (per slice of clip_axis, or one pair)
  int clip_axis;            // -1: one pair for the channel
  float* out;
};

__device__ __forceinline__ void axis_taps(int mode, double s, int n, int pad, int* idx, double* w, int& ntap, int& src_slice) {
  // s is the coordinate in the UNPADDED input grid
  if (mode == 1) {
    int j = (int)floor(s + 0.5);
    j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    idx[0] = j + pad;
    w[0] = 1.0;
    ntap = 1;
    src_slice = j;
  } else if (mode == 2) {
    double c = s < 0.0 ? 0.0 : (s > (double)(n - 1) ? (double)(n - 1) : s);
    int j = (int)floor(c);
    if (j > n - 1) j = n - 1;
    const double t = c - (double)j;
    idx[0] = j + pad;
    idx[1] = (j + 1 > n - 1 ? n - 1 : j + 1) + pad;
    w[0] = 1.0 - t;
    w[1] = t;
    ntap = 2;
    src_slice = j;
  } else {
    const double c = s + (double)pad;
    const double f = floor(c);
    const double y = c - f, zc = 1.0 - y;
    w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
    w[2] = (zc * zc * (zc - 2.0) * 3.0 + 4.0) / 6.0;
    w[0] = zc * zc * zc / 6.0;
    w[3] = 1.0 - w[0] - w[1] - w[2];
    const int j = (int)f - 1;
    const int m = n + 2 * pad;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int q = j + k;
      idx[k] = q < 0 ? 0 : (q > m - 1 ? m - 1 : q);
    }
    ntap = 4;
    src_slice = 0;
  }
}

__global__ void __launch_bounds__(256) spline_resample_kernel(const ResampleArgs a) {
  const size_t total = (size_t)a.out_d[0] * a.out_d[1] * a.out_d[2];
  const size_t s1 = (size_t)(a.in_d[2] + 2 * a.pad[2]);
  const size_t s0 = (size_t)(a.in_d[1] + 2 * a.pad[1]) * s1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int o[3];
    o[2] = (int)(i % a.out_d[2]);
    o[1] = (int)((i / a.out_d[2]) % a.out_d[1]);
    o[0] = (int)(i / ((size_t)a.out_d[2] * a.out_d[1]));
    int idx[3][4], nt[3], sl[3];
    double w[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double s = ((double)o[k] + 0.5) * a.scale[k] - 0.5;
      axis_taps(a.mode[k], s, a.in_d[k], a.pad[k], idx[k], w[k], nt[k], sl[k]);
    }
    // scipy accumulates  sum over taps of coefficient * (w0 * w1 * w2)  with the first axis outermost
    double v = 0.0;
    for (int i0 = 0; i0 < nt[0]; ++i0)
      for (int i1 = 0; i1 < nt[1]; ++i1) {
        const double* row = a.coef + (size_t)idx[0][i0] * s0 + (size_t)idx[1][i1] * s1;
        const double w01 = w[0][i0] * w[1][i1];
        for (int i2 = 0; i2 < nt[2]; ++i2) v += row[idx[2][i2]] * (w01 * w[2][i2]);
      }
    if (a.clip) {
      const float* cl = a.clip + (a.clip_axis >= 0 ? 2 * sl[a.clip_axis] : 0);
      const double lo = (double)cl[0], hi = (double)cl[1];
      v = v < lo ? lo : (v > hi ? hi : v);
    }
    a.out[i] = (float)v;
  }
}

// min / max of a float32 channel, per slice along `axis` (or one pair when axis < 0): out[n][2]
__global__ void __launch_bounds__(256) slice_minmax_kernel(const float* __restrict__ src, int n0, int n1, int n2, int axis,
                                                           int* __restrict__ enc /* [n][2] ordered-int */) {
  const size_t total = (size_t)n0 * n1 * n2;
  auto encf = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
  // the loop bound is warp-uniform so that every lane takes part in the shuffles
  const size_t rounded = (total + 31) / 32 * 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += (size_t)gridDim.x * blockDim.x) {
    const bool live = i < total;
    const size_t ii = live ? i : total - 1;
    const int c[3] = {(int)(ii / ((size_t)n2 * n1)), (int)((ii / n2) % n1), (int)(ii % n2)};
    const int s = axis >= 0 ? c[axis] : 0;
    int lo = encf(src[ii]), hi = lo;
    const int s0 = __shfl_sync(0xffffffffu, s, 0);
    if (__all_sync(0xffffffffu, s == s0)) {        // the usual case: one slice per warp -> one atomic pair per warp
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
      }
      if ((threadIdx.x & 31) == 0) {
        atomicMin(enc + 2 * s, lo);
        atomicMax(enc + 2 * s + 1, hi);
      }
    } else {
      atomicMin(enc + 2 * s, lo);
      atomicMax(enc + 2 * s + 1, hi);
    }
  }
}

__global__ void decode_minmax_kernel(const int* __restrict__ enc, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int e = enc[i];
    out[i] = __int_as_float(e >= 0 ? e : e ^ 0x7fffffff);
  }
}

static inline unsigned blocks_for(size_t items, int threads) {
  size_t b = (items + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b == 0) b = 1;
  return (unsigned)b;
}

static int make_view(const int dims[3], const int tf[3], View3& v) {
  int seen = 0;
  for (int k = 0; k < 3; ++k) {
    if (dims[k] < 1 || tf[k] < 0 || tf[k] > 2) return -1;
    seen |= 1 << tf[k];
    v.d[k] = dims[k];
    v.tf[k] = tf[k];
  }
  if (seen != 7) return -1;
  for (int k = 0; k < 3; ++k) v.td[k] = dims[tf[k]];
  return 0;
}

}  // namespace fnnu

using namespace fnnu;

extern "C" int fnnu_pre_nonzero_bbox(const float* image, int channels, const int dims[3], const int transpose_forward[3],
                                     int32_t* bbox6_dev, void* stream) {
  FNNU_CHECK_ARG(image && dims && transpose_forward && bbox6_dev && channels >= 1, "pre_nonzero_bbox: bad argument");
  View3 v;
  FNNU_CHECK_ARG(make_view(dims, transpose_forward, v) == 0, "pre_nonzero_bbox: dims / transpose_forward");
  const int init[6] = {INT_MAX, 0, INT_MAX, 0, INT_MAX, 0};
  FNNU_CUDA(cudaMemcpyAsync(bbox6_dev, init, sizeof(init), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  const size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
  nonzero_bbox_kernel<<<blocks_for(nvox, 256), 256, 0, (cudaStream_t)stream>>>(image, channels, v, bbox6_dev);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

extern "C" int fnnu_pre_filled_mask(const float* image, int channels, const int dims[3], const int transpose_forward[3],
                                    const int bbox_lo[3], const int crop_dims[3], uint8_t* state, int32_t* changed_dev,
                                    int max_iterations, void* stream) {
  FNNU_CHECK_ARG(image && dims && transpose_forward && bbox_lo && crop_dims && state && changed_dev, "pre_filled_mask: null pointer");
  View3 v;
  FNNU_CHECK_ARG(make_view(dims, transpose_forward, v) == 0, "pre_filled_mask: dims / transpose_forward");
  for (int k = 0; k < 3; ++k)
    FNNU_CHECK_ARG(bbox_lo[k] >= 0 && crop_dims[k] >= 1 && bbox_lo[k] + crop_dims[k] <= v.td[k], "pre_filled_mask: crop outside the image on axis %d", k);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t total = (size_t)crop_dims[0] * crop_dims[1] * crop_dims[2];
  mask_state_kernel<<<blocks_for(total, 256), 256, 0, s>>>(image, channels, v, bbox_lo[0], bbox_lo[1], bbox_lo[2], crop_dims[0],
                                                           crop_dims[1], crop_dims[2], state);
  FNNU_LAUNCH_CHECK();
  for (int it = 0; it < max_iterations; ++it) {
    FNNU_CUDA(cudaMemsetAsync(changed_dev, 0, sizeof(int32_t), s));
    for (int axis = 2; axis >= 0; --axis) {
      const size_t lines = total / crop_dims[axis];
      fill_sweep_kernel<<<blocks_for(lines, 128), 128, 0, s>>>(state, crop_dims[0], crop_dims[1], crop_dims[2], axis, changed_dev);
      FNNU_LAUNCH_CHECK();
    }
    int32_t changed = 0;
    FNNU_CUDA(cudaMemcpyAsync(&changed, changed_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FNNU_CUDA(cudaStreamSynchronize(s));
    if (!changed) return FNNU_OK;
  }
  set_error("pre_filled_mask: flood fill did not converge in %d iterations", max_iterations);
  return FNNU_E_UNSUPPORTED;
}

extern "C" int fnnu_pre_channel_stats(const float* channel, const int dims[3], const int transpose_forward[3],
                                      const int bbox_lo[3], const int crop_dims[3], const uint8_t* state_or_null,
                                      double* stats5_host, void* stream) {
  FNNU_CHECK_ARG(channel && dims && transpose_forward && bbox_lo && crop_dims && stats5_host, "pre_channel_stats: null pointer");
  View3 v;
  FNNU_CHECK_ARG(make_view(dims, transpose_forward, v) == 0, "pre_channel_stats: dims / transpose_forward");
  cudaStream_t s = (cudaStream_t)stream;
  double* d_out = nullptr;
  FNNU_CUDA(cudaMallocAsync((void**)&d_out, 3 * sizeof(double) + 2 * sizeof(int), s));
  int* d_mm = reinterpret_cast<int*>(d_out + 3);
  const double zero[3] = {0, 0, 0};
  const int mm[2] = {INT_MAX, INT_MIN};
  FNNU_CUDA(cudaMemcpyAsync(d_out, zero, sizeof(zero), cudaMemcpyHostToDevice, s));
  FNNU_CUDA(cudaMemcpyAsync(d_mm, mm, sizeof(mm), cudaMemcpyHostToDevice, s));
  const size_t total = (size_t)crop_dims[0] * crop_dims[1] * crop_dims[2];
  channel_stats_kernel<<<blocks_for(total, 256), 256, 0, s>>>(channel, v, bbox_lo[0], bbox_lo[1], bbox_lo[2], crop_dims[0],
                                                              crop_dims[1], crop_dims[2], state_or_null, d_out, d_mm);
  FNNU_LAUNCH_CHECK();
  double h[3];
  int hm[2];
  FNNU_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, s));
  FNNU_CUDA(cudaMemcpyAsync(hm, d_mm, sizeof(hm), cudaMemcpyDeviceToHost, s));
  FNNU_CUDA(cudaStreamSynchronize(s));
  FNNU_CUDA(cudaFreeAsync(d_out, s));
  auto dec = [](int e) { int i = e >= 0 ? e : e ^ 0x7fffffff; float f; memcpy(&f, &i, 4); return (double)f; };
  stats5_host[0] = h[0];
  stats5_host[1] = h[1];
  stats5_host[2] = h[2];
  stats5_host[3] = dec(hm[0]);
  stats5_host[4] = dec(hm[1]);
  return FNNU_OK;
}

extern "C" int fnnu_pre_crop_normalize(const float* channel, const int dims[3], const int transpose_forward[3],
                                       const int bbox_lo[3], const int crop_dims[3], int mode, float a, float b, float lo,
                                       float hi, const uint8_t* state_or_null, float* out, void* stream) {
  FNNU_CHECK_ARG(channel && dims && transpose_forward && bbox_lo && crop_dims && out, "pre_crop_normalize: null pointer");
  FNNU_CHECK_ARG(mode >= 0 && mode <= 3 && (mode != 3 || state_or_null), "pre_crop_normalize: mode=%d", mode);
  View3 v;
  FNNU_CHECK_ARG(make_view(dims, transpose_forward, v) == 0, "pre_crop_normalize: dims / transpose_forward");
  for (int k = 0; k < 3; ++k)
    FNNU_CHECK_ARG(bbox_lo[k] >= 0 && crop_dims[k] >= 1 && bbox_lo[k] + crop_dims[k] <= v.td[k], "pre_crop_normalize: crop outside the image on axis %d", k);
  const size_t total = (size_t)crop_dims[0] * crop_dims[1] * crop_dims[2];
  crop_normalize_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      channel, v, bbox_lo[0], bbox_lo[1], bbox_lo[2], crop_dims[0], crop_dims[1], crop_dims[2], mode, a, b, lo, hi, state_or_null, out);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}

extern "C" size_t fnnu_pre_resample_workspace_bytes(const int in_dims[3], const int axis_mode[3]) {
  size_t n = 1, slices = 1;
  for (int k = 0; k < 3; ++k) {
    n *= (size_t)(in_dims[k] + (axis_mode[k] == 0 ? 24 : 0));
    if (axis_mode[k] == 1 && (size_t)in_dims[k] > slices) slices = in_dims[k];
  }
  return n * sizeof(double) + (slices * 2 + 64) * (sizeof(int) + sizeof(float));
}

extern "C" int fnnu_pre_resample_channel(const float* src, const int in_dims[3], const int out_dims[3], const int axis_mode[3],
                                         int clip_to_input_range, void* workspace, size_t workspace_bytes, float* out,
                                         void* stream) {
  FNNU_CHECK_ARG(src && in_dims && out_dims && axis_mode && workspace && out, "pre_resample_channel: null pointer");
  FNNU_CHECK_ARG(workspace_bytes >= fnnu_pre_resample_workspace_bytes(in_dims, axis_mode), "pre_resample_channel: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  ResampleArgs a;
  int nearest_axis = -1, n_nearest = 0;
  size_t padded = 1;
  for (int k = 0; k < 3; ++k) {
    FNNU_CHECK_ARG(in_dims[k] >= 1 && out_dims[k] >= 1 && axis_mode[k] >= 0 && axis_mode[k] <= 2, "pre_resample_channel: axis %d", k);
    a.in_d[k] = in_dims[k];
    a.out_d[k] = out_dims[k];
    a.mode[k] = axis_mode[k];
    a.pad[k] = axis_mode[k] == 0 ? 12 : 0;
    a.scale[k] = (double)in_dims[k] / (double)out_dims[k];
    padded *= (size_t)(in_dims[k] + 2 * a.pad[k]);
    if (axis_mode[k] == 1) { nearest_axis = k; ++n_nearest; }
  }
  FNNU_CHECK_ARG(n_nearest <= 1, "pre_resample_channel: at most one nearest (separate-z) axis");
  double* coef = (double*)workspace;
  int* enc = (int*)(coef + padded);
  const int n_slices = nearest_axis >= 0 ? in_dims[nearest_axis] : 1;
  float* clipv = (float*)(enc + 2 * n_slices + 32);
  spline_pad_kernel<<<blocks_for(padded, 256), 256, 0, s>>>(src, in_dims[0], in_dims[1], in_dims[2], a.pad[0], a.pad[1], a.pad[2], coef);
  FNNU_LAUNCH_CHECK();
  const int m[3] = {in_dims[0] + 2 * a.pad[0], in_dims[1] + 2 * a.pad[1], in_dims[2] + 2 * a.pad[2]};
  for (int axis = 0; axis < 3; ++axis) {
    if (axis_mode[axis] != 0) continue;
    spline_prefilter_kernel<<<blocks_for(padded / m[axis], 128), 128, 0, s>>>(coef, m[0], m[1], m[2], axis);
    FNNU_LAUNCH_CHECK();
  }
  a.clip = nullptr;
  a.clip_axis = nearest_axis;
  if (clip_to_input_range) {
    std::vector<int> init(2 * n_slices);
    for (int i = 0; i < n_slices; ++i) { init[2 * i] = INT_MAX; init[2 * i + 1] = INT_MIN; }
    FNNU_CUDA(cudaMemcpyAsync(enc, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    const size_t total = (size_t)in_dims[0] * in_dims[1] * in_dims[2];
    slice_minmax_kernel<<<blocks_for(total, 256), 256, 0, s>>>(src, in_dims[0], in_dims[1], in_dims[2], nearest_axis, enc);
    FNNU_LAUNCH_CHECK();
    decode_minmax_kernel<<<(2 * n_slices + 127) / 128, 128, 0, s>>>(enc, clipv, 2 * n_slices);
    FNNU_LAUNCH_CHECK();
    a.clip = clipv;
  }
  a.coef = coef;
  a.out = out;
  const size_t total_out = (size_t)out_dims[0] * out_dims[1] * out_dims[2];
  spline_resample_kernel<<<blocks_for(total_out, 256), 256, 0, s>>>(a);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}
