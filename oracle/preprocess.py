"""Oracle restatement of the pre-processing of one case for inference.  Test infrastructure only
(see oracle/__init__.py).

Follows /root/reference/distillation/nnunetv2/
  preprocessing/preprocessors/default_preprocessor.py:45-118   run_case_npy (seg=None branch)
  preprocessing/cropping/cropping.py:8-39                      create_nonzero_mask / crop_to_nonzero
  preprocessing/normalization/default_normalization_schemes.py:27-95
  preprocessing/resampling/default_resampling.py:75-192        compute_new_shape, resample_data_or_seg (order 3, order_z 0)
Third-party arithmetic: skimage.transform.resize (see oracle/export.py: scipy.ndimage.zoom(order, mode='nearest',
grid_mode=True) followed by `clip=True`, i.e. clipping to the input's own min / max — skimage/transform/_warps.py
_clip_warp_output); acvl_utils.get_bbox_from_mask (first / last index with any True per axis, half-open).  scipy is
installed and executes here.  PARITY: pinned to scipy.ndimage; unpinned for the skimage / acvl_utils wrappers.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import binary_fill_holes, map_coordinates, zoom

from .export import determine_do_sep_z_and_axis


def create_nonzero_mask(data):
    m = data[0] != 0
    for c in range(1, data.shape[0]):
        m |= data[c] != 0
    return binary_fill_holes(m)


def get_bbox_from_mask(mask):
    out = []
    for ax in range(mask.ndim):
        other = tuple(i for i in range(mask.ndim) if i != ax)
        idx = np.where(mask.any(axis=other))[0]
        out.append([0, mask.shape[ax]] if len(idx) == 0 else [int(idx[0]), int(idx[-1]) + 1])
    return out


def crop_to_nonzero(data, nonzero_label=-1):
    mask = create_nonzero_mask(data)
    bbox = get_bbox_from_mask(mask)
    sl = tuple(slice(b[0], b[1]) for b in bbox)
    seg = np.where(mask[sl][None], np.int8(0), np.int8(nonzero_label))
    return data[(slice(None),) + sl], seg, bbox


def normalize(data, seg, schemes, use_mask, props_per_channel):
    """default_preprocessor.py:120-135 + the scheme classes, float32 in place like numpy does it."""
    for c in range(data.shape[0]):
        img = data[c]
        scheme = schemes[c]
        if scheme == 'ZScoreNormalization':
            if use_mask[c]:
                mask = seg[0] >= 0
                mean, std = img[mask].mean(), img[mask].std()
                img[mask] = (img[mask] - mean) / (max(std, 1e-8))
            else:
                mean, std = img.mean(), img.std()
                img -= mean
                img /= (max(std, 1e-8))
        elif scheme == 'CTNormalization':
            ip = props_per_channel[str(c)]
            np.clip(img, ip['percentile_00_5'], ip['percentile_99_5'], out=img)
            img -= ip['mean']
            img /= max(ip['std'], 1e-8)
        elif scheme == 'NoNormalization':
            pass
        elif scheme == 'RescaleTo01Normalization':
            img -= img.min()
            img /= np.clip(img.max(), a_min=1e-8, a_max=None)
        elif scheme == 'RGBTo01Normalization':
            img /= 255.
        else:
            raise NotImplementedError(scheme)
    return data


def compute_new_shape(old_shape, old_spacing, new_spacing):
    return np.array([int(round(i / j * k)) for i, j, k in zip(old_spacing, new_spacing, old_shape)])


def _resize(img, new_shape, order):
    """skimage.transform.resize(img, new_shape, order, mode='edge', anti_aliasing=False) [clip=True]."""
    img = np.asarray(img, dtype=np.float64)
    if tuple(img.shape) == tuple(int(s) for s in new_shape):
        return img.copy()
    out = zoom(img, [float(n) / float(o) for n, o in zip(new_shape, img.shape)], order=order, mode='nearest', grid_mode=True)
    return np.clip(out, img.min(), img.max())


def resample_data(data, new_shape, current_spacing, new_spacing, order=3, order_z=0, force_separate_z=None):
    """resample_data_or_seg_to_shape (default_resampling.py:89-108) -> resample_data_or_seg, is_seg=False."""
    do_separate_z, axis = determine_do_sep_z_and_axis(force_separate_z, current_spacing, new_spacing)
    shape = np.array(data[0].shape)
    new_shape = np.array([int(s) for s in new_shape])
    if not np.any(shape != new_shape):
        return data
    out = np.zeros((data.shape[0], *new_shape), dtype=data.dtype)
    data = data.astype(float, copy=False)
    if do_separate_z:
        new_shape_2d = new_shape[[i for i in range(3) if i != axis]]
        for c in range(data.shape[0]):
            tmp = new_shape.copy()
            tmp[axis] = shape[axis]
            here = np.zeros(tmp)
            for s in range(shape[axis]):
                sl = [slice(None)] * 3
                sl[axis] = s
                here[tuple(sl)] = _resize(data[c][tuple(sl)], new_shape_2d, order)
            if shape[axis] != new_shape[axis]:
                rows, cols, dim = new_shape
                orig = here.shape
                mr, mc, md = np.mgrid[:rows, :cols, :dim]
                mr = (float(orig[0]) / rows) * (mr + 0.5) - 0.5
                mc = (float(orig[1]) / cols) * (mc + 0.5) - 0.5
                md = (float(orig[2]) / dim) * (md + 0.5) - 0.5
                out[c] = map_coordinates(here, np.array([mr, mc, md]), order=order_z, mode='nearest')[None]
            else:
                out[c] = here
    else:
        for c in range(data.shape[0]):
            out[c] = _resize(data[c], new_shape, order)
    return out


def run_case_npy(data, properties, transpose_forward, target_spacing, schemes, use_mask, props_per_channel,
                 resampling_kwargs=None):
    """default_preprocessor.py:45-118 for a test case (no segmentation).  Returns (data float32, properties)."""
    kw = resampling_kwargs or {'order': 3, 'order_z': 0, 'force_separate_z': None}
    data = np.asarray(data).astype(np.float32)
    data = data.transpose([0, *[i + 1 for i in transpose_forward]])
    original_spacing = [properties['spacing'][i] for i in transpose_forward]
    props = dict(properties)
    props['shape_before_cropping'] = data.shape[1:]
    data, seg, bbox = crop_to_nonzero(data)
    data = np.ascontiguousarray(data)
    props['bbox_used_for_cropping'] = bbox
    props['shape_after_cropping_and_before_resampling'] = data.shape[1:]
    target_spacing = list(target_spacing)
    if len(target_spacing) < len(data.shape[1:]):
        target_spacing = [original_spacing[0]] + target_spacing
    new_shape = compute_new_shape(data.shape[1:], original_spacing, target_spacing)
    data = normalize(data, seg, schemes, use_mask, props_per_channel)
    data = resample_data(data, new_shape, original_spacing, target_spacing, kw.get('order', 3), kw.get('order_z', 0),
                         kw.get('force_separate_z', None))
    return data, props
