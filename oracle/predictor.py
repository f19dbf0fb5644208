"""Oracle restatement of the sliding-window loop, mirror TTA, Gaussian-weighted
accumulation, normalisation and final argmax.  Test infrastructure only (see
oracle/__init__.py).

Follows /root/reference/distillation/nnunetv2/inference/predict_from_raw_data.py
  _internal_maybe_mirror_and_predict              :541-557
  _internal_predict_sliding_window_return_logits  :560-631
  predict_sliding_window_return_logits            :634-680
and utilities/label_handling/label_handling.py:143-195 (argmax on the logits,
first maximum wins).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from .sliding_window import gaussian_map, pad_to_patch, slicers_for  # noqa: F401 (pad_to_patch re-exported)


def mirror_axes_combinations(mirror_axes):
    """predict_from_raw_data.py:550-553 — (2,),(3,),(4,),(2,3),(2,4),(3,4),(2,3,4) for axes (0,1,2)."""
    axes = [m + 2 for m in mirror_axes]
    return [c for i in range(len(axes)) for c in itertools.combinations(axes, i + 1)]


@torch.inference_mode()
def mirror_and_predict(network, x, mirror_axes):
    """predict_from_raw_data.py:541-557."""
    prediction = network(x)
    if mirror_axes is not None:
        assert max(mirror_axes) <= x.ndim - 3
        combos = mirror_axes_combinations(mirror_axes)
        for axes in combos:
            prediction += torch.flip(network(torch.flip(x, axes)), axes)
        prediction /= (len(combos) + 1)
    return prediction


@torch.inference_mode()
def accumulate_tiles(tile_predictions, slicers, volume_shape, num_heads, patch_size, use_gaussian=True,
                     acc_dtype=torch.half):
    """predict_from_raw_data.py:587-625 for already-computed per-tile predictions
    (list of (heads, *patch) tensors): fp16 accumulators, `pred *= g`, two `+=`, one `div`."""
    dev = tile_predictions[0].device       # CPU in the reference; the GPU box runs the same arithmetic faster
    logits = torch.zeros((num_heads, *volume_shape), dtype=acc_dtype, device=dev)
    n_pred = torch.zeros(volume_shape, dtype=acc_dtype, device=dev)
    g = gaussian_map(tuple(patch_size), 1. / 8, 10).to(dev) if use_gaussian else 1
    for pred, sl in zip(tile_predictions, slicers):
        pred = pred.clone()
        if use_gaussian:
            pred *= g
        logits[sl] += pred
        n_pred[sl[1:]] += g
    torch.div(logits, n_pred, out=logits)
    if torch.any(torch.isinf(logits)):
        raise RuntimeError('Encountered inf in predicted array.')
    return logits, n_pred


@torch.inference_mode()
def predict_sliding_window_return_logits(network, input_image, patch_size, tile_step_size=0.5, use_gaussian=True,
                                         mirror_axes=(0, 1, 2), tile_subset=None, return_tile_predictions=False,
                                         autocast_device=None, acc_dtype=torch.half, return_n_predictions=False):
    """predict_from_raw_data.py:634-680 (CPU semantics: fp32 network, fp16 accumulators; pass
    autocast_device='cuda' with a CUDA network/input for the reference's GPU semantics).
    `tile_subset` restricts the loop to some tile indices (bounded CPU-baseline samples).
    `acc_dtype=torch.float32` is NOT the reference's arithmetic: it is the exact-accumulation variant used
    to separate network error from the reference's fp16-accumulator noise in the parity report."""
    assert isinstance(input_image, torch.Tensor) and input_image.ndim == 4
    network.eval()
    ctx = torch.autocast(autocast_device, enabled=True) if autocast_device else _Null()
    with ctx:
        data, revert = pad_to_patch(input_image, patch_size)
        slicers = slicers_for(data.shape[1:], patch_size, tile_step_size)
        if tile_subset is not None:
            slicers = [slicers[i] for i in tile_subset]
        heads = None
        g = gaussian_map(tuple(patch_size), 1. / 8, 10).to(data.device) if use_gaussian else 1
        logits = n_pred = None
        tile_preds = []
        for sl in slicers:
            workon = torch.clone(data[sl][None], memory_format=torch.contiguous_format)
            pred = mirror_and_predict(network, workon, mirror_axes)[0]
            if logits is None:
                heads = pred.shape[0]
                logits = torch.zeros((heads, *data.shape[1:]), dtype=acc_dtype, device=data.device)
                n_pred = torch.zeros(data.shape[1:], dtype=acc_dtype, device=data.device)
            if return_tile_predictions:
                tile_preds.append(pred.clone())
            if use_gaussian:
                pred *= g
            logits[sl] += pred
            n_pred[sl[1:]] += g
        if tile_subset is None:
            torch.div(logits, n_pred, out=logits)
            if torch.any(torch.isinf(logits)):
                raise RuntimeError('Encountered inf in predicted array.')
        logits = logits[(slice(None), *revert[1:])]
        n_pred = n_pred[tuple(revert[1:])]
    if return_tile_predictions:
        return logits, tile_preds, slicers
    if return_n_predictions:
        return logits, n_pred
    return logits


def logits_to_segmentation(logits):
    """label_handling.py:143-182 non-region branch: argmax over heads, first max wins."""
    if isinstance(logits, torch.Tensor):
        logits = logits.float().cpu().numpy()
    return np.argmax(logits, 0)


def dice_per_class(a, b, num_classes):
    """evaluation/evaluate_predictions.py:110 — 2TP / (2TP + FP + FN)."""
    out = []
    for c in range(num_classes):
        ma, mb = (a == c), (b == c)
        tp = np.logical_and(ma, mb).sum()
        fp = np.logical_and(~ma, mb).sum()
        fn = np.logical_and(ma, ~mb).sum()
        den = 2 * tp + fp + fn
        out.append(float('nan') if den == 0 else float(2 * tp / den))
    return out


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
