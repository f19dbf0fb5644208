"""Logits -> label map in the original image geometry, on the device.

Drop-in for inference/export_prediction.py:14-71 `convert_predicted_logits_to_segmentation_with_correct_shape`
(same name, argument meaning and return value) for non-region label maps: the (heads, x, y, z) fp16 logits stay on
the GPU, ONE libfnnu kernel resamples them to `shape_after_cropping_and_before_resampling` (order 1; order 0 along
the anisotropic axis exactly when the reference's `determine_do_sep_z_and_axis`,
preprocessing/resampling/default_resampling.py:37-72, says so), arg-maxes, inserts the crop into the canvas and undoes
`transpose_forward`; one uint8 per voxel crosses PCIe instead of 2 x heads bytes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

ANISO_THRESHOLD = 3          # nnunetv2/configuration.py


def determine_do_sep_z_and_axis(force_separate_z: Optional[bool], current_spacing, new_spacing,
                                separate_z_anisotropy_threshold: float = ANISO_THRESHOLD) -> Tuple[bool, Optional[int]]:
    """default_resampling.py:37-72, same decisions."""
    def do_sep(spacing):
        return (np.max(spacing) / np.min(spacing)) > separate_z_anisotropy_threshold

    def lowres_axis(spacing):
        return np.where(max(spacing) / np.array(spacing) == 1)[0]

    if force_separate_z is not None:
        do_separate_z = force_separate_z
        axis = lowres_axis(current_spacing) if force_separate_z else None
    elif do_sep(current_spacing):
        do_separate_z, axis = True, lowres_axis(current_spacing)
    elif do_sep(new_spacing):
        do_separate_z, axis = True, lowres_axis(new_spacing)
    else:
        do_separate_z, axis = False, None
    if axis is not None:
        if len(axis) in (2, 3):
            do_separate_z, axis = False, None
        else:
            axis = int(axis[0])
    return bool(do_separate_z), axis


def axis_modes(in_shape: Sequence[int], out_shape: Sequence[int], current_spacing, new_spacing, order_z: int = 0,
               force_separate_z: Optional[bool] = None) -> Tuple[int, int, int]:
    """1 where resample_data_or_seg (default_resampling.py:111-192) samples an axis with order 0: the separate-z axis,
    when its extent changes and order_z == 0.  (When it does not change the slices are copied: identical to both.)"""
    do_sep, axis = determine_do_sep_z_and_axis(force_separate_z, current_spacing, new_spacing)
    modes = [0, 0, 0]
    if do_sep and axis is not None:
        if order_z != 0:
            raise NotImplementedError('order_z != 0 for the separate-z axis is not on the B200 export path')
        modes[axis] = 1
    return tuple(modes)


def source_coordinates(n_in: int, n_out: int, nearest: bool):
    """Host mirror of the kernel's per-axis sampling (float64): indices i0, i1 and weight t of output voxel o.
    scipy.ndimage.zoom(grid_mode=True, mode='nearest'): s = (o + 0.5) * n_in / n_out - 0.5, clamped to the grid."""
    s = (np.arange(n_out, dtype=np.float64) + 0.5) * (float(n_in) / float(n_out)) - 0.5
    if nearest:
        j = np.clip(np.floor(s + 0.5).astype(np.int64), 0, n_in - 1)
        return j, j, np.zeros(n_out)
    s = np.clip(s, 0.0, float(n_in - 1))
    i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    return i0, i1, s - i0


def export_labels(logits: torch.Tensor, mid_shape: Sequence[int], nearest_axes: Sequence[int], bbox,
                  canvas_shape: Sequence[int], transpose_backward: Sequence[int]) -> torch.Tensor:
    """libfnnu call: (heads, x, y, z) fp16 device logits -> uint8 device label map in the original axis order."""
    lib = _lib.load()
    assert logits.is_cuda and logits.dtype == torch.float16 and logits.ndim == 4 and logits.is_contiguous()
    tb = [int(t) for t in transpose_backward]
    out_shape = tuple(int(canvas_shape[t]) for t in tb)
    out = torch.empty(out_shape, dtype=torch.uint8, device=logits.device)
    lo = [int(b[0]) for b in bbox]
    with torch.cuda.device(logits.device):
        _lib.check(lib.fnnu_export_labels(C.c_void_p(logits.data_ptr()), int(logits.shape[0]), _lib.i3(logits.shape[1:]),
                                          _lib.i3(mid_shape), _lib.i3(nearest_axes), _lib.i3(lo), _lib.i3(canvas_shape),
                                          _lib.i3(tb), C.c_void_p(out.data_ptr()), _lib.stream_ptr()))
    return out


def convert_predicted_logits_to_segmentation_with_correct_shape(predicted_logits: torch.Tensor, plans_manager,
                                                                configuration_manager, label_manager,
                                                                properties_dict: dict,
                                                                return_probabilities: bool = False,
                                                                num_threads_torch: int = 8, to_host=None):
    """export_prediction.py:14-71 for logits that live on the GPU.  Returns the numpy label map (uint8) in the
    image's own axis order and shape.  Region-based label maps and probability export go through the host path
    (fast_nnunet_b200.prepost)."""
    if return_probabilities or label_manager.has_regions:
        from . import prepost
        return prepost.logits_to_segmentation_with_correct_shape(predicted_logits.cpu(), properties_dict, plans_manager,
                                                                 label_manager, return_probabilities)
    if len(label_manager.foreground_labels) >= 255:
        raise NotImplementedError('more than 254 foreground labels need uint16 label maps (not on the B200 export path)')
    tf = list(plans_manager.transpose_forward)
    spacing_transposed = [properties_dict['spacing'][i] for i in tf]
    mid = tuple(int(v) for v in properties_dict['shape_after_cropping_and_before_resampling'])
    current_spacing = list(configuration_manager.spacing) if len(configuration_manager.spacing) == len(mid) else \
        [spacing_transposed[0], *configuration_manager.spacing]
    kw = dict(getattr(configuration_manager, 'resampling_fn_probabilities_kwargs', None) or
              {'is_seg': False, 'order': 1, 'order_z': 0, 'force_separate_z': None})
    if kw.get('order', 1) != 1 or kw.get('is_seg', False):
        raise NotImplementedError(f'resampling_fn_probabilities_kwargs {kw} (the B200 export path implements order 1)')
    logits = predicted_logits
    if not logits.is_cuda:
        raise RuntimeError('the B200 export path takes logits that live on the GPU; there is no CPU path')
    logits = logits.to(torch.float16).contiguous()
    modes = (0, 0, 0)
    if tuple(logits.shape[1:]) != mid:
        modes = axis_modes(logits.shape[1:], mid, current_spacing, spacing_transposed, kw.get('order_z', 0),
                           kw.get('force_separate_z', None))
    lab = export_labels(logits, mid, modes, properties_dict['bbox_used_for_cropping'],
                        properties_dict['shape_before_cropping'], plans_manager.transpose_backward)
    if to_host is not None:
        return to_host(lab).numpy()
    return lab.cpu().numpy()
