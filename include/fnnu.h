/*
 * fnnu.h — C ABI of libfnnu.so, the B200 (sm_100a) sliding-window inference engine.
 *
 * The reference (77even/Fast-nnUNet) has no FFI/plugin interface for this path: its boundary is the
 * Python class nnUNetPredictor (distillation/nnunetv2/inference/predict_from_raw_data.py:39-680) and,
 * one level below, the call `self.network(x)` (:543, :555).  Each entry point below names the piece
 * of that file it replaces.  All pointers are plain device or host pointers (no torch types); every
 * function returns 0 on success or a negative FNNU_E_* code, and fnnu_last_error() returns a
 * thread-local message.  All work is enqueued on the caller's cudaStream_t (passed as void*) and is
 * asynchronous with respect to the host; the library allocates no device memory of its own — the
 * caller provides the parameter arena and the activation workspace (sizes are queried first).
 *
 * Tensor layouts
 *   volume      : float32 [C][X][Y][Z]           (reference layout, predict_from_raw_data.py:649)
 *   tile batch  : fp16    [n][pX][pY][pZ][Cs]    channels-last, n = tiles x flips
 *   accumulator : float32 or fp16 [H][X][Y][Z]   (predicted_logits, :587-589)
 *   weight sum  : float32 or fp16 [X][Y][Z]      (n_predictions, :590)
 *   gaussian    : fp16    [pX][pY][pZ]           (compute_gaussian, sliding_window_prediction.py:10-27)
 */
#ifndef FNNU_H_
#define FNNU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNNU_ABI_VERSION 1

enum {
  FNNU_OK = 0,
  FNNU_E_INVALID = -1,   /* bad argument (AssertionError in the reference, e.g. :649, :548) */
  FNNU_E_CUDA = -2,      /* CUDA runtime error */
  FNNU_E_UNSUPPORTED = -3,
  FNNU_E_INF = -4        /* inf in the normalised logits (RuntimeError at :622-625) */
};

enum { FNNU_ACC_F32 = 0, FNNU_ACC_F16 = 1 };   /* accumulator arithmetic; F16 reproduces :587-590,613 */
enum { FNNU_IN_F16 = 0, FNNU_IN_F32 = 1 };     /* dtype of per-tile predictions handed to accumulate */

int fnnu_abi_version(void);
const char* fnnu_last_error(void);
/* 1 if the binary holds sm_100a code and the current device can run it. */
int fnnu_device_ok(void);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound operators of the sliding window (predict_from_raw_data.py:560-631)
 * ---------------------------------------------------------------------------------------------- */

/* Replaces the producer thread `torch.clone(d[s][None]).to(device)` (:568-571) and the
 * `torch.flip(x, axes)` of the mirror loop (:555): cuts n_tiles tiles out of the device-resident
 * volume and writes every requested mirrored copy, fp32 -> fp16, channels-last.
 * starts: device int32 [n_tiles][3]; flip_masks: host uint8 [n_flips], bit0/1/2 = flip axis X/Y/Z;
 * out: fp16 [n_tiles*n_flips][pX][pY][pZ][c_stride] (sample index = tile*n_flips + flip). */
int fnnu_gather_tiles(const float* volume, int channels, const int vol_dims[3],
                      const int32_t* starts_dev, int n_tiles, const int patch[3],
                      const uint8_t* flip_masks, int n_flips,
                      void* out, int c_stride, void* stream);

/* Replaces `prediction += flip(net(flip(x)))`, `/= n` (:555-556), `prediction *= gaussian` (:612) and
 * `predicted_logits[sl] += prediction` (:613) for n_tiles tiles, processed in tile order.
 * preds: [n_tiles*n_flips][pX][pY][pZ][p_stride] (fp16 or fp32, see in_dtype), heads <= p_stride;
 * gaussian: fp16 [pX][pY][pZ] or NULL (use_gaussian=False); acc: [heads][X][Y][Z] of acc_dtype. */
int fnnu_accumulate_tiles(const void* preds, int in_dtype, int p_stride, int heads,
                          const int32_t* starts_host, int n_tiles, const int patch[3],
                          const uint8_t* flip_masks, int n_flips,
                          const void* gaussian, void* acc, int acc_dtype, const int vol_dims[3],
                          void* stream);

/* Replaces `n_predictions[sl[1:]] += gaussian` (:614).  n_predictions does not depend on the image,
 * so it is produced in one pass: every voxel sums the map values of the tiles covering it, in tile
 * order, in acc_dtype arithmetic.  steps_*: host int32 per-axis tile starts
 * (compute_steps_for_sliding_window).  gaussian may be NULL (weight 1 per tile). */
int fnnu_weight_sum(const int32_t* steps_x, int nx, const int32_t* steps_y, int ny,
                    const int32_t* steps_z, int nz, const int patch[3], const void* gaussian,
                    void* wsum, int acc_dtype, const int vol_dims[3], void* stream);

/* Replaces `torch.div(predicted_logits, n_predictions, out=...)`, the inf check (:620-625) and
 * LabelManager.convert_logits_to_segmentation (label_handling.py:184-195: argmax over heads, first
 * maximum wins).  logits_out (fp16 [H][X][Y][Z]) and labels_out (uint8 [X][Y][Z]) may each be NULL.
 * inf_flag_dev: device int32, set to 1 if any normalised logit is +-inf (checked by the caller).
 * acc_head_stride: elements between consecutive heads of acc (0 = dense, X*Y*Z), so that a slab view
 * of a larger accumulator can be normalised in place; wsum and the outputs are dense over vol_dims. */
int fnnu_finalize(const void* acc, const void* wsum, int acc_dtype, int heads, const int vol_dims[3],
                  size_t acc_head_stride, void* logits_out, uint8_t* labels_out, int32_t* inf_flag_dev,
                  void* stream);

/* Number of kernels the memory-bound entry points above have launched in this process (bench.py's
 * gpu_launches evidence; fnnu_engine_launch_counts gives the network's). */
long long fnnu_mem_launches(void);

/* Fold ensembling (predict_from_raw_data.py:483-500: `prediction += ...; prediction /= n_folds`): every fold's
 * tile predictions are accumulated into the SAME accumulator, so the fold mean is a division of the weight
 * sum: x *= factor over n elements (fp32). */
int fnnu_scale_inplace_f32(float* x, float factor, size_t n, void* stream);

/* Multi-GPU halo step: acc += other over a contiguous range of n elements (fp32). */
int fnnu_add_inplace_f32(float* acc, const float* other, size_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Export: logits -> label map in the ORIGINAL image geometry (inference/export_prediction.py:14-71)
 * ---------------------------------------------------------------------------------------------- */

/* Replaces convert_predicted_logits_to_segmentation_with_correct_shape for non-region label maps:
 * resampling_fn_probabilities (default_resampling.py:89-192: order 1; order 0 along `nearest_axis` axes, the
 * reference's "separate z" branch) of the fp16 logits [heads][in_dims] to `mid_dims`
 * (shape_after_cropping_and_before_resampling), rounding to fp16 like the reference's output array, argmax (first
 * maximum), insert_crop_into_image at bbox_lo into a zero canvas of canvas_dims (shape_before_cropping) and
 * `.transpose(transpose_backward)`.  labels_out: uint8, dense, extents canvas_dims permuted by
 * transpose_backward.  All arrays are in the transposed (network) axis order. */
int fnnu_export_labels(const void* logits, int heads, const int in_dims[3], const int mid_dims[3],
                       const int nearest_axis[3], const int bbox_lo[3], const int canvas_dims[3],
                       const int transpose_backward[3], uint8_t* labels_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pre-processing of one case (preprocessing/preprocessors/default_preprocessor.py:45-118, test-case branch)
 * image: float32 [C][dims] in the image's own axis order; "transposed" = after transpose_forward.
 * ---------------------------------------------------------------------------------------------- */

/* crop_to_nonzero's bounding box (cropping/cropping.py:8-39; binary_fill_holes cannot change it):
 * bbox6_dev = {lo0, hi0, lo1, hi1, lo2, hi2} in transposed coordinates, hi exclusive; lo = INT_MAX when the image
 * is all zeros (the caller then takes the whole image). */
int fnnu_pre_nonzero_bbox(const float* image, int channels, const int dims[3], const int transpose_forward[3],
                          int32_t* bbox6_dev, void* stream);

/* create_nonzero_mask with binary_fill_holes (6-connectivity) on the cropped box: state[crop] = 0 non-zero voxel,
 * 1 enclosed zero voxel (a filled hole), 2 zero voxel connected to the outside.  The reference's seg is
 * `state == 2 ? -1 : 0`.  Synchronises the stream once per flood-fill iteration. */
int fnnu_pre_filled_mask(const float* image, int channels, const int dims[3], const int transpose_forward[3],
                         const int bbox_lo[3], const int crop_dims[3], uint8_t* state, int32_t* changed_dev,
                         int max_iterations, void* stream);

/* sum, sum of squares, count, min, max of one cropped channel (inside the mask when state != NULL), float64 sums;
 * stats5_host is written after a stream synchronisation. */
int fnnu_pre_channel_stats(const float* channel, const int dims[3], const int transpose_forward[3],
                           const int bbox_lo[3], const int crop_dims[3], const uint8_t* state_or_null,
                           double* stats5_host, void* stream);

/* transpose + crop + intensity normalisation of one channel (default_normalization_schemes.py:27-95), float32
 * arithmetic operation by operation as numpy performs it.  mode 0: copy (NoNormalization); 1: (clip(x, lo, hi) - a) / b
 * (CTNormalization); 2: (x - a) / b (ZScore, RescaleTo01, RGBTo01); 3: mode 2 inside the mask only
 * (ZScore with use_mask_for_norm).  out: float32 [crop_dims]. */
int fnnu_pre_crop_normalize(const float* channel, const int dims[3], const int transpose_forward[3],
                            const int bbox_lo[3], const int crop_dims[3], int mode, float a, float b, float lo,
                            float hi, const uint8_t* state_or_null, float* out, void* stream);

/* resampling_fn_data (default_resampling.py:111-192) of one channel: axis_mode[k] = 0 cubic spline (order 3), 1 nearest
 * (the separate-z axis, order_z 0), 2 linear (order 1); clip_to_input_range = skimage.transform.resize's clip=True
 * (per slice of the nearest axis, as the reference resizes slice by slice there).  float64 coefficients live in the
 * caller's workspace (fnnu_pre_resample_workspace_bytes). */
size_t fnnu_pre_resample_workspace_bytes(const int in_dims[3], const int axis_mode[3]);
int fnnu_pre_resample_channel(const float* src, const int in_dims[3], const int out_dims[3], const int axis_mode[3],
                              int clip_to_input_range, void* workspace, size_t workspace_bytes, float* out,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-patch network forward (replaces `self.network(x)`, :543/:555; PlainConvUNet /
 * ResidualEncoderUNet of dynamic_network_architectures as built by get_network_from_plans.py:9-43)
 * ---------------------------------------------------------------------------------------------- */

typedef struct fnnu_engine fnnu_engine;

enum {
  FNNU_OP_CONV = 0,      /* Conv3d(k, stride, pad=(k-1)/2) [+bias] writing raw output + InstanceNorm sums */
  FNNU_OP_TCONV = 1,     /* ConvTranspose3d(kernel == stride) [+bias] */
  FNNU_OP_ADD_ACT = 2,   /* dst = lrelu(xform(src) + xform(src2))   (residual join of BasicBlockD) */
  FNNU_OP_AVGPOOL = 3    /* AvgPool3d(stride, stride) of xform(src) */
};

/* One activation buffer: channels-last fp16 [batch][dims0][dims1][dims2][channels]. */
typedef struct {
  int32_t dims[3];
  int32_t channels;
} fnnu_buffer_desc;

/* One operator.  Sources are read through their pending per-channel transform (InstanceNorm affine
 * + LeakyReLU of the producing layer, applied at load); CONV with has_norm=1 leaves its own output
 * pending in the same way.  Host pointers are fp32 in PyTorch layouts and are copied/packed into the
 * parameter arena by fnnu_engine_create. */
typedef struct {
  int32_t op;
  int32_t src, src_coff;          /* source buffer index, first channel */
  int32_t src2, src2_coff;        /* second source (ADD_ACT) or -1 */
  int32_t dst, dst_coff;
  int32_t cin, cout;
  int32_t kernel[3], stride[3];
  int32_t has_bias, has_norm;
  float   norm_eps;
  float   act_slope;              /* slope applied after the norm when the output is consumed; 1 = none */
  const float* weight;            /* CONV: [cout][cin][k0][k1][k2]; TCONV: [cin][cout][s0][s1][s2] */
  const float* bias;              /* [cout] or NULL */
  const float* gamma;             /* [cout] or NULL */
  const float* beta;              /* [cout] or NULL */
} fnnu_op_desc;

/* Sizes the caller must allocate (device bytes) for a program and a maximum batch. */
int fnnu_engine_sizes(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops,
                      int max_batch, size_t* param_bytes, size_t* workspace_bytes);

/* Builds the engine: validates the program, packs the weights into param_arena (device), lays the
 * activation buffers out in workspace (device).  Both arenas must outlive the engine. */
int fnnu_engine_create(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops,
                       int max_batch, void* param_arena, size_t param_bytes,
                       void* workspace, size_t workspace_bytes, void* stream, fnnu_engine** out);
void fnnu_engine_destroy(fnnu_engine* e);

/* Device pointer of activation buffer `index` (fp16 [max_batch][dims][channels]). */
void* fnnu_engine_buffer(fnnu_engine* e, int index);

/* Runs every operator of the program for `batch` samples.  The caller has filled buffer
 * `input_index` beforehand (fnnu_gather_tiles) and reads buffer `output_index` afterwards. */
int fnnu_engine_forward(fnnu_engine* e, int batch, void* stream);

/* Selects the convolution back end: 0 = tcgen05 implicit GEMM where the shape is supported, CUDA-core
 * direct kernel elsewhere (default); 1 = CUDA-core direct kernel everywhere (debug cross-check). */
int fnnu_engine_set_backend(fnnu_engine* e, int backend);
/* Number of kernels the last fnnu_engine_forward launched, and how many of them were tcgen05. */
int fnnu_engine_launch_counts(fnnu_engine* e, int* total, int* umma);

/* Measurement hook (bench.py): fnnu_engine_profile_op(e, i) makes every following fnnu_engine_forward bracket
 * operator i with two CUDA events on the launching stream (i < 0 switches it off); fnnu_engine_profile_ms
 * synchronises on the second event and returns the device time of that operator in the LAST forward. */
int fnnu_engine_profile_op(fnnu_engine* e, int op_index);
int fnnu_engine_profile_ms(fnnu_engine* e, float* ms);

/* Device pointer of the InstanceNorm sums of buffer `index`: double [max_batch][channels][2]
 * (sum, sum of squares of the stored fp16 values), valid after fnnu_engine_forward. */
double* fnnu_engine_stats(fnnu_engine* e, int index);

#ifdef __cplusplus
}
#endif
#endif  /* FNNU_H_ */
