"""Oracle restatement of tile-start computation, Gaussian importance map and
padding.  Test infrastructure only (see oracle/__init__.py).

Follows /root/reference/distillation/nnunetv2/inference/sliding_window_prediction.py
  compute_gaussian                    :10-27
  compute_steps_for_sliding_window    :30-54
and predict_from_raw_data.py:506-538 (slicer order) and the published
behaviour of acvl_utils.cropping_and_padding.padding.pad_nd_image
(call site predict_from_raw_data.py:657-659).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def steps_for_sliding_window(image_size, tile_size, tile_step_size):
    """sliding_window_prediction.py:30-54 — per-axis tile starts."""
    assert all(i >= j for i, j in zip(image_size, tile_size))
    assert 0 < tile_step_size <= 1
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step_value = image_size[dim] - tile_size[dim]
        if num_steps[dim] > 1:
            actual = max_step_value / (num_steps[dim] - 1)
        else:
            actual = 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def gaussian_map(tile_size, sigma_scale=1.0 / 8, value_scaling_factor=1.0, dtype=torch.float16):
    """sliding_window_prediction.py:10-27 — importance map, cast to `dtype`,
    zeros replaced by the smallest non-zero entry."""
    tmp = np.zeros(tile_size)
    center = [i // 2 for i in tile_size]
    sigmas = [i * sigma_scale for i in tile_size]
    tmp[tuple(center)] = 1
    g = gaussian_filter(tmp, sigmas, 0, mode='constant', cval=0)
    g = torch.from_numpy(g)
    g /= (torch.max(g) / value_scaling_factor)
    g = g.to(dtype=dtype)
    mask = g == 0
    g[mask] = torch.min(g[~mask])
    return g


def slicers_for(image_size, patch_size, tile_step_size):
    """predict_from_raw_data.py:526-537 — 3-D branch, order sx (outer), sy, sz."""
    steps = steps_for_sliding_window(image_size, patch_size, tile_step_size)
    out = []
    for sx in steps[0]:
        for sy in steps[1]:
            for sz in steps[2]:
                out.append((slice(None),) + tuple(slice(s, s + t) for s, t in zip((sx, sy, sz), patch_size)))
    return out


def pad_to_patch(image: torch.Tensor, patch_size):
    """acvl_utils pad_nd_image(image, patch, 'constant', {'value': 0}, True, None):
    symmetric zero pad of the trailing len(patch) axes up to patch size
    (below = diff // 2, above = diff // 2 + diff % 2); returns the padded tensor
    and the slicer that crops it back."""
    nd = len(patch_size)
    old = np.array(image.shape[-nd:])
    new = np.maximum(old, np.array(patch_size))
    diff = new - old
    below = diff // 2
    above = diff // 2 + diff % 2
    if diff.sum() > 0:
        pad = []
        for b, a in zip(below[::-1], above[::-1]):
            pad += [int(b), int(a)]
        image = torch.nn.functional.pad(image, pad, mode='constant', value=0)
    lead = image.ndim - nd
    slicer = tuple([slice(None)] * lead + [slice(int(b), int(b + o)) for b, o in zip(below, old)])
    return image, slicer
