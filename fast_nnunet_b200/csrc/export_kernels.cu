// Logits -> label map export on the device, one kernel, one byte written per voxel:
//   resample the fp16 logits to the pre-resampling shape (linear; nearest along the anisotropic axis when the
//   reference resamples "separate z"), round to fp16 as the reference's output array does, argmax (first maximum
//   wins), insert into the un-cropped canvas, undo transpose_forward.
// Replaces inference/export_prediction.py:14-71 (convert_predicted_logits_to_segmentation_with_correct_shape) with
// preprocessing/resampling/default_resampling.py:89-192 (resample_data_or_seg, order 1 / order_z 0) underneath, whose
// skimage.transform.resize(order=1, mode='edge', anti_aliasing=False) is scipy.ndimage.zoom(order=1, mode='nearest',
// grid_mode=True): source coordinate = (o + 0.5) * in / out - 0.5, edge-clamped, interpolated in float64.
#include "common.cuh"
#include "ops.cuh"

namespace fnnu {

struct ExportArgs {
  const __half* logits;     // [heads][in_d0][in_d1][in_d2]
  int heads;
  int in_d[3];              // grid of the logits (network spacing)
  int mid_d[3];             // shape_after_cropping_and_before_resampling
  int nearest[3];           // 1: order-0 sampling on this axis
  int lo[3];                // bbox_used_for_cropping lower corner
  int canvas_d[3];          // shape_before_cropping (transposed axes)
  int tb[3];                // transpose_backward: out axis k = canvas axis tb[k]
  int out_d[3];             // canvas_d permuted by tb
  double scale[3];          // in_d / mid_d
  uint8_t* out;
};

__global__ void __launch_bounds__(256) export_labels_kernel(const ExportArgs a) {
  const size_t total = (size_t)a.out_d[0] * a.out_d[1] * a.out_d[2];
  const size_t in_plane = (size_t)a.in_d[1] * a.in_d[2];
  const size_t in_vox = in_plane * a.in_d[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int o[3];
    o[2] = (int)(i % a.out_d[2]);
    o[1] = (int)((i / a.out_d[2]) % a.out_d[1]);
    o[0] = (int)(i / ((size_t)a.out_d[2] * a.out_d[1]));
    int c[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ax = a.tb[k];
      const int v = o[k] - a.lo[ax];
      c[ax] = v;
      inside = inside && v >= 0 && v < a.mid_d[ax];
    }
    uint8_t label = 0;
    if (inside) {
      int i0[3], i1[3];
      double t[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int n = a.in_d[k];
        double s = ((double)c[k] + 0.5) * a.scale[k] - 0.5;
        if (a.nearest[k]) {
          int j = (int)floor(s + 0.5);
          j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
          i0[k] = i1[k] = j;
          t[k] = 0.0;
        } else {
          if (s < 0.0) s = 0.0;
          if (s > (double)(n - 1)) s = (double)(n - 1);
          int j = (int)floor(s);
          if (j > n - 1) j = n - 1;
          i0[k] = j;
          i1[k] = j + 1 > n - 1 ? n - 1 : j + 1;
          t[k] = s - (double)j;
        }
      }
      const size_t b00 = (size_t)i0[0] * in_plane + (size_t)i0[1] * a.in_d[2];
      const size_t b01 = (size_t)i0[0] * in_plane + (size_t)i1[1] * a.in_d[2];
      const size_t b10 = (size_t)i1[0] * in_plane + (size_t)i0[1] * a.in_d[2];
      const size_t b11 = (size_t)i1[0] * in_plane + (size_t)i1[1] * a.in_d[2];
      float best = 0.f;
      int arg = 0;
      for (int h = 0; h < a.heads; ++h) {
        const __half* p = a.logits + (size_t)h * in_vox;
        const double v000 = __half2float(p[b00 + i0[2]]), v001 = __half2float(p[b00 + i1[2]]);
        const double v010 = __half2float(p[b01 + i0[2]]), v011 = __half2float(p[b01 + i1[2]]);
        const double v100 = __half2float(p[b10 + i0[2]]), v101 = __half2float(p[b10 + i1[2]]);
        const double v110 = __half2float(p[b11 + i0[2]]), v111 = __half2float(p[b11 + i1[2]]);
        const double w0 = 1.0 - t[0], w1 = 1.0 - t[1], w2 = 1.0 - t[2];
        // the sum over the 8 corners with product weights, as the spline evaluation of scipy's zoom does
        double v = (w0 * w1 * w2) * v000 + (w0 * w1 * t[2]) * v001 + (w0 * t[1] * w2) * v010 + (w0 * t[1] * t[2]) * v011 +
                   (t[0] * w1 * w2) * v100 + (t[0] * w1 * t[2]) * v101 + (t[0] * t[1] * w2) * v110 + (t[0] * t[1] * t[2]) * v111;
        const float r = __half2float(__double2half(v));     // the reference stores the resampled logits in fp16
        if (h == 0 || r > best) {
          best = r;
          arg = h;
        }
      }
      label = (uint8_t)arg;
    }
    a.out[i] = label;
  }
}

}  // namespace fnnu

using namespace fnnu;

extern "C" int fnnu_export_labels(const void* logits, int heads, const int in_dims[3], const int mid_dims[3],
                                  const int nearest_axis[3], const int bbox_lo[3], const int canvas_dims[3],
                                  const int transpose_backward[3], uint8_t* labels_out, void* stream) {
  FNNU_CHECK_ARG(logits && in_dims && mid_dims && nearest_axis && bbox_lo && canvas_dims && transpose_backward && labels_out,
                 "export_labels: null pointer");
  FNNU_CHECK_ARG(heads >= 1 && heads <= 255, "export_labels: heads=%d", heads);
  ExportArgs a;
  a.logits = (const __half*)logits;
  a.heads = heads;
  a.out = labels_out;
  int seen = 0;
  for (int k = 0; k < 3; ++k) {
    FNNU_CHECK_ARG(in_dims[k] >= 1 && mid_dims[k] >= 1 && canvas_dims[k] >= 1, "export_labels: non-positive extent on axis %d", k);
    FNNU_CHECK_ARG(bbox_lo[k] >= 0 && bbox_lo[k] + mid_dims[k] <= canvas_dims[k],
                   "export_labels: bounding box [%d, %d) outside the canvas extent %d on axis %d", bbox_lo[k],
                   bbox_lo[k] + mid_dims[k], canvas_dims[k], k);
    FNNU_CHECK_ARG(transpose_backward[k] >= 0 && transpose_backward[k] < 3, "export_labels: transpose_backward[%d]=%d", k,
                   transpose_backward[k]);
    seen |= 1 << transpose_backward[k];
    a.in_d[k] = in_dims[k];
    a.mid_d[k] = mid_dims[k];
    a.nearest[k] = nearest_axis[k] ? 1 : 0;
    a.lo[k] = bbox_lo[k];
    a.canvas_d[k] = canvas_dims[k];
    a.tb[k] = transpose_backward[k];
    a.scale[k] = (double)in_dims[k] / (double)mid_dims[k];
  }
  FNNU_CHECK_ARG(seen == 7, "export_labels: transpose_backward is not a permutation");
  for (int k = 0; k < 3; ++k) a.out_d[k] = a.canvas_d[a.tb[k]];
  const size_t total = (size_t)a.out_d[0] * a.out_d[1] * a.out_d[2];
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  export_labels_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  FNNU_LAUNCH_CHECK();
  return FNNU_OK;
}
