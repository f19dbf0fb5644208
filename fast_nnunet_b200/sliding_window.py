"""Host-side tile geometry for the sliding window: tile starts, tile order,
Gaussian importance map, padding.  Integer/float64 host arithmetic; the results
are uploaded once per (volume shape, patch) and consumed by the CUDA kernels.

Bit-exact counterparts of the reference's
  inference/sliding_window_prediction.py:30-54  (compute_steps_for_sliding_window)
  inference/sliding_window_prediction.py:10-27  (compute_gaussian)
  inference/predict_from_raw_data.py:526-537    (tile order: sx outer, sy, sz inner)
checked against golden vectors produced by the reference's own code
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np


def compute_steps_for_sliding_window(image_size: Sequence[int], tile_size: Sequence[int],
                                     tile_step_size: float) -> List[List[int]]:
    """Per-axis tile starts (same name and semantics as the reference function)."""
    assert all(i >= j for i, j in zip(image_size, tile_size)), \
        'image size must be as large or larger than patch_size'
    assert 0 < tile_step_size <= 1, 'step_size must be larger than 0 and smaller or equal to 1'
    steps = []
    for img, tile in zip(image_size, tile_size):
        target = tile * tile_step_size                      # float64, as in the reference
        n = int(math.ceil((img - tile) / target)) + 1
        span = img - tile
        if n > 1:
            actual = span / (n - 1)
            # np.round == round-half-to-even on float64; Python's round() on floats matches.
            steps.append([int(round(actual * i)) for i in range(n)])
        else:
            steps.append([0])
    return steps


def tile_starts(image_size: Sequence[int], tile_size: Sequence[int], tile_step_size: float) -> np.ndarray:
    """All tile origins as int32 [n_tiles, 3], in the reference's loop order
    (first axis outermost)."""
    steps = compute_steps_for_sliding_window(image_size, tile_size, tile_step_size)
    grid = np.stack(np.meshgrid(*[np.asarray(s, dtype=np.int32) for s in steps], indexing='ij'), -1)
    return np.ascontiguousarray(grid.reshape(-1, len(tile_size)))


def _gaussian_kernel1d(sigma: float, radius: int) -> np.ndarray:
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


@lru_cache(maxsize=8)
def _gaussian_f64(tile_size: Tuple[int, ...], sigma_scale: float, value_scaling_factor: float) -> np.ndarray:
    """float64 map before the dtype cast.  A delta filtered by a separable, truncated Gaussian with
    constant-0 borders is the outer product of the 1-D kernels sampled around the centre, multiplied
    axis by axis in the order the separable filter runs (axis 0 first); the centre lies within the
    truncation radius int(4*sigma + 0.5) for every tile size, so no tap is cut off."""
    prof = []
    for n in tile_size:
        sigma = n * sigma_scale
        radius = int(4.0 * float(sigma) + 0.5)
        k = _gaussian_kernel1d(sigma, radius)
        c = n // 2
        off = np.arange(n) - c
        p = np.zeros(n, dtype=np.float64)
        ok = np.abs(off) <= radius
        p[ok] = k[off[ok] + radius]
        prof.append(p)
    g = prof[0]
    for p in prof[1:]:
        g = g[..., None] * p
    g = g / (g.max() / value_scaling_factor)
    return g


def compute_gaussian(tile_size: Sequence[int], sigma_scale: float = 1. / 8, value_scaling_factor: float = 1,
                     dtype=np.float16) -> np.ndarray:
    """Importance map as a numpy array of `dtype`; zeros are replaced by the smallest non-zero entry,
    exactly as the reference does after its cast."""
    g = _gaussian_f64(tuple(int(t) for t in tile_size), float(sigma_scale), float(value_scaling_factor))
    # torch's float64 -> float16 cast (what the reference runs) rounds through float32; do the same
    # so that the fp16 bits agree (104 of 1.47 M entries differ for (160, 96, 96) otherwise).
    g = g.astype(np.float32).astype(dtype)
    mask = g == 0
    if mask.any():
        g[mask] = g[~mask].min()
    return g


def pad_amounts(shape: Sequence[int], patch_size: Sequence[int]):
    """Symmetric zero padding up to the patch size (acvl_utils.pad_nd_image semantics used at
    predict_from_raw_data.py:657-659): returns (below, above) per spatial axis."""
    below, above = [], []
    for s, p in zip(shape, patch_size):
        d = max(p - s, 0)
        below.append(d // 2)
        above.append(d // 2 + d % 2)
    return below, above


def weight_sum_map(image_size: Sequence[int], tile_size: Sequence[int], starts: np.ndarray,
                   gaussian16: np.ndarray, tile_order_fp16: bool = True) -> np.ndarray:
    """n_predictions of predict_from_raw_data.py:614 — input independent, so it is computed once on
    the host: fp16 running sum in tile order when `tile_order_fp16` (the reference's arithmetic),
    else exact float32."""
    if tile_order_fp16:
        acc = np.zeros(image_size, dtype=np.float16)
        for s in starts:
            sl = tuple(slice(int(a), int(a) + int(t)) for a, t in zip(s, tile_size))
            acc[sl] = (acc[sl].astype(np.float32) + gaussian16.astype(np.float32)).astype(np.float16)
        return acc
    acc = np.zeros(image_size, dtype=np.float32)
    g = gaussian16.astype(np.float32)
    for s in starts:
        sl = tuple(slice(int(a), int(a) + int(t)) for a, t in zip(s, tile_size))
        acc[sl] += g
    return acc
