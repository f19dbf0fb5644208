"""Host-side steps either side of the hot path, for `predict_single_npy_array` only.

These are rows f1/f2 of SURVEY.md §8 ("next"); what is here is the subset needed for an array that
is already at the plans' spacing: transpose, crop-to-nonzero, intensity normalisation, and the way
back (argmax, un-crop, inverse transpose).  Resampling is NOT implemented and raises.

Follows preprocessing/preprocessors/default_preprocessor.py:45-118, preprocessing/cropping/cropping.py:8-39,
preprocessing/normalization/default_normalization_schemes.py:27-95, inference/export_prediction.py:14-71.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import binary_fill_holes


def _bbox_from_mask(mask: np.ndarray):
    if not mask.any():
        return [[0, s] for s in mask.shape]
    out = []
    for ax in range(mask.ndim):
        other = tuple(i for i in range(mask.ndim) if i != ax)
        idx = np.where(mask.any(axis=other))[0]
        out.append([int(idx[0]), int(idx[-1]) + 1])
    return out


def _normalize(data, seg, schemes, use_mask, props_per_channel):
    for c in range(data.shape[0]):
        scheme = schemes[c]
        img = data[c]
        if scheme == 'ZScoreNormalization':
            if use_mask[c]:
                mask = seg[0] >= 0
                mean, std = img[mask].mean(), img[mask].std()
                img[mask] = (img[mask] - mean) / max(std, 1e-8)
            else:
                mean, std = img.mean(), img.std()
                img -= mean
                img /= max(std, 1e-8)
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
            raise NotImplementedError(f'normalization scheme {scheme}')
    return data


def preprocess_npy(image, properties, seg_prev, plans_manager, configuration_manager, dataset_json, label_manager):
    data = np.asarray(image).astype(np.float32)
    assert data.ndim == 4, 'input_image must be (c, x, y, z)'
    tf = list(plans_manager.transpose_forward)
    data = data.transpose([0, *[i + 1 for i in tf]])
    original_spacing = [properties['spacing'][i] for i in tf]
    props = dict(properties)
    props['shape_before_cropping'] = data.shape[1:]
    nonzero = data[0] != 0
    for c in range(1, data.shape[0]):
        nonzero |= data[c] != 0
    nonzero = binary_fill_holes(nonzero)
    bbox = _bbox_from_mask(nonzero)
    sl = tuple(slice(b[0], b[1]) for b in bbox)
    seg = np.where(nonzero[sl], np.int8(0), np.int8(-1))[None]
    data = np.ascontiguousarray(data[(slice(None),) + sl])
    props['bbox_used_for_cropping'] = bbox
    props['shape_after_cropping_and_before_resampling'] = data.shape[1:]
    target = list(configuration_manager.spacing)
    if not np.allclose(np.asarray(original_spacing, dtype=np.float64), np.asarray(target, dtype=np.float64), rtol=1e-3):
        raise NotImplementedError(f'resampling from spacing {original_spacing} to {target} is outside the B200 '
                                  f'inference path (SURVEY.md §8 f2); pass data at the plans\' spacing')
    data = _normalize(data, seg, configuration_manager.normalization_schemes, configuration_manager.use_mask_for_norm,
                      plans_manager.foreground_intensity_properties_per_channel)
    if seg_prev is not None:
        sp = np.asarray(seg_prev).transpose([0, *[i + 1 for i in tf]])[(slice(None),) + sl]
        onehot = np.stack([(sp[0] == l) for l in label_manager.foreground_labels]).astype(np.float32)
        data = np.concatenate([data, onehot], 0)
    return data, props


def logits_to_segmentation_with_correct_shape(logits: torch.Tensor, props, plans_manager, label_manager,
                                              return_probabilities=False):
    shape = tuple(props['shape_after_cropping_and_before_resampling'])
    assert tuple(logits.shape[1:]) == shape, 'resampling of logits is not implemented (SURVEY.md §8 f1)'
    seg = label_manager.convert_logits_to_segmentation(logits)
    seg = seg.numpy() if isinstance(seg, torch.Tensor) else seg
    out = np.zeros(props['shape_before_cropping'],
                   dtype=np.uint8 if len(label_manager.foreground_labels) < 255 else np.uint16)
    sl = tuple(slice(b[0], b[1]) for b in props['bbox_used_for_cropping'])
    out[sl] = seg
    out = out.transpose(plans_manager.transpose_backward)
    if not return_probabilities:
        return out
    lf = logits.float()
    probs = torch.sigmoid(lf) if label_manager.has_regions else torch.softmax(lf, 0)
    full = np.zeros((probs.shape[0], *props['shape_before_cropping']), dtype=np.float32)
    if not label_manager.has_regions:
        full[0] = 1
    full[(slice(None),) + sl] = probs.numpy()
    full = full.transpose([0] + [i + 1 for i in plans_manager.transpose_backward])
    return out, full
