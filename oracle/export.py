"""Oracle restatement of the logits -> segmentation export.  Test infrastructure only (see oracle/__init__.py).

Follows /root/reference/distillation/nnunetv2/
  inference/export_prediction.py:14-71                      convert_predicted_logits_to_segmentation_with_correct_shape
  preprocessing/resampling/default_resampling.py:37-72      determine_do_sep_z_and_axis
  preprocessing/resampling/default_resampling.py:111-192    resample_data_or_seg (is_seg=False branch)
  utilities/label_handling/label_handling.py:143-195        argmax on the logits, first maximum wins
Third-party arithmetic: `skimage.transform.resize(img, shape, order, mode='edge', anti_aliasing=False)` (scikit-image,
un-vendored and un-pinned by the reference; absent from this image) is, from scikit-image 0.19 on
(skimage/transform/_warps.py), `scipy.ndimage.zoom(img, shape / img.shape, order=order, mode='nearest',
grid_mode=True)`; scipy IS installed here, so the arithmetic below executes the real library.  `insert_crop_into_image`
(acvl_utils, un-vendored) is `canvas[bbox slices] = crop`.  PARITY: pinned to scipy.ndimage for the resampling
arithmetic, unpinned for the two wrappers named above.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
from scipy.ndimage import map_coordinates, zoom

ANISO_THRESHOLD = 3


def _resize(img, new_shape, order):
    """skimage.transform.resize(img, new_shape, order, mode='edge', anti_aliasing=False) as scikit-image >= 0.19 runs it."""
    img = np.asarray(img, dtype=np.float64)
    if tuple(img.shape) == tuple(int(s) for s in new_shape):
        return img.copy()
    factors = [float(n) / float(o) for n, o in zip(new_shape, img.shape)]
    out = zoom(img, factors, order=order, mode='nearest', grid_mode=True)
    assert tuple(out.shape) == tuple(int(s) for s in new_shape)
    return out


def get_do_separate_z(spacing, anisotropy_threshold=ANISO_THRESHOLD):
    return (np.max(spacing) / np.min(spacing)) > anisotropy_threshold


def get_lowres_axis(new_spacing):
    return np.where(max(new_spacing) / np.array(new_spacing) == 1)[0]


def determine_do_sep_z_and_axis(force_separate_z, current_spacing, new_spacing, threshold=ANISO_THRESHOLD):
    """default_resampling.py:37-72."""
    if force_separate_z is not None:
        do_separate_z = force_separate_z
        axis = get_lowres_axis(current_spacing) if force_separate_z else None
    else:
        if get_do_separate_z(current_spacing, threshold):
            do_separate_z, axis = True, get_lowres_axis(current_spacing)
        elif get_do_separate_z(new_spacing, threshold):
            do_separate_z, axis = True, get_lowres_axis(new_spacing)
        else:
            do_separate_z, axis = False, None
    if axis is not None:
        if len(axis) == 3 or len(axis) == 2:
            do_separate_z, axis = False, None
        else:
            axis = axis[0]
    return do_separate_z, axis


def resample_data_or_seg(data, new_shape, axis=None, order=3, do_separate_z=False, order_z=0):
    """default_resampling.py:111-192, is_seg=False.  The output array has the INPUT's dtype (fp16 logits in, fp16 out)."""
    assert data.ndim == 4
    shape = np.array(data[0].shape)
    new_shape = np.array([int(s) for s in new_shape])
    dtype_out = data.dtype
    out = np.zeros((data.shape[0], *new_shape), dtype=dtype_out)
    if not np.any(shape != new_shape):
        return data
    data = data.astype(float, copy=False)
    if do_separate_z:
        assert axis is not None
        new_shape_2d = new_shape[[i for i in range(3) if i != axis]]
        for c in range(data.shape[0]):
            tmp = deepcopy(new_shape)
            tmp[axis] = shape[axis]
            here = np.zeros(tmp)
            for s in range(shape[axis]):
                sl = [slice(None)] * 3
                sl[axis] = s
                here[tuple(sl)] = _resize(data[c][tuple(sl)], new_shape_2d, order)
            if shape[axis] != new_shape[axis]:
                rows, cols, dim = new_shape
                orig_rows, orig_cols, orig_dim = here.shape
                mr, mc, md = np.mgrid[:rows, :cols, :dim]
                mr = (float(orig_rows) / rows) * (mr + 0.5) - 0.5
                mc = (float(orig_cols) / cols) * (mc + 0.5) - 0.5
                md = (float(orig_dim) / dim) * (md + 0.5) - 0.5
                out[c] = map_coordinates(here, np.array([mr, mc, md]), order=order_z, mode='nearest')[None]
            else:
                out[c] = here
    else:
        for c in range(data.shape[0]):
            out[c] = _resize(data[c], new_shape, order)
    return out


def resample_data_or_seg_to_shape(data, new_shape, current_spacing, new_spacing, order=3, order_z=0,
                                  force_separate_z=None):
    """default_resampling.py:89-108."""
    do_separate_z, axis = determine_do_sep_z_and_axis(force_separate_z, current_spacing, new_spacing)
    return resample_data_or_seg(data, new_shape, axis, order, do_separate_z, order_z=order_z)


def convert_predicted_logits_to_segmentation_with_correct_shape(predicted_logits, plans_spacing, transpose_forward,
                                                                transpose_backward, properties_dict, num_foreground=1):
    """export_prediction.py:14-71 with return_probabilities=False.  predicted_logits: numpy (heads, x, y, z), fp16
    as the predictor returns it."""
    spacing_transposed = [properties_dict['spacing'][i] for i in transpose_forward]
    mid = properties_dict['shape_after_cropping_and_before_resampling']
    current_spacing = list(plans_spacing) if len(plans_spacing) == len(mid) else [spacing_transposed[0], *plans_spacing]
    logits = resample_data_or_seg_to_shape(np.asarray(predicted_logits), mid, current_spacing, spacing_transposed,
                                           order=1, order_z=0, force_separate_z=None)
    seg = np.argmax(logits, 0)                                   # label_handling.py:143-182, non-region branch
    canvas = np.zeros(properties_dict['shape_before_cropping'], dtype=np.uint8 if num_foreground < 255 else np.uint16)
    sl = tuple(slice(b[0], b[1]) for b in properties_dict['bbox_used_for_cropping'])
    canvas[sl] = seg                                             # insert_crop_into_image
    return canvas.transpose(transpose_backward)
