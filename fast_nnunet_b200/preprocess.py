"""Pre-processing of one case on the device.

Drop-in for DefaultPreprocessor.run_case_npy (preprocessing/preprocessors/default_preprocessor.py:45-118) in its
test-case form (no segmentation), as nnUNetPredictor.predict_single_npy_array reaches it through
PreprocessAdapterFromNpy (inference/data_iterators.py:17-58): the raw image goes to the GPU once; transpose_forward,
crop_to_nonzero, the intensity normalisation of every channel and the resampling to the plans' spacing are libfnnu
kernels (csrc/preprocess_kernels.cu); the returned tensor stays on the device, ready for the sliding window.
`properties` receives the same keys the reference writes (shape_before_cropping, bbox_used_for_cropping,
shape_after_cropping_and_before_resampling).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .export import determine_do_sep_z_and_axis


def compute_new_shape(old_shape, old_spacing, new_spacing):
    """default_resampling.py:75-86."""
    assert len(old_spacing) == len(old_shape) == len(new_spacing)
    return [int(round(i / j * k)) for i, j, k in zip(old_spacing, new_spacing, old_shape)]


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def nonzero_bbox(image: torch.Tensor, transpose_forward: Sequence[int]):
    """cropping.py:19-39: bounding box (transposed axes, half-open) of the voxels where any channel is non-zero."""
    lib = _lib.load()
    box = torch.empty(6, dtype=torch.int32, device=image.device)
    _lib.check(lib.fnnu_pre_nonzero_bbox(_ptr(image), int(image.shape[0]), _lib.i3(image.shape[1:]),
                                         _lib.i3(transpose_forward), _ptr(box), _lib.stream_ptr()))
    b = box.cpu().numpy().tolist()
    tdims = [int(image.shape[1 + t]) for t in transpose_forward]
    out = []
    for k in range(3):
        lo, hi = b[2 * k], b[2 * k + 1]
        out.append([0, tdims[k]] if hi <= lo else [int(lo), int(hi)])      # an all-zero image keeps its full extent
    return out


def filled_mask_state(image: torch.Tensor, transpose_forward, bbox) -> torch.Tensor:
    """create_nonzero_mask (cropping.py:8-17) on the cropped box: uint8 state, 2 = outside (the reference's seg == -1)."""
    lib = _lib.load()
    crop = [b[1] - b[0] for b in bbox]
    state = torch.empty(crop, dtype=torch.uint8, device=image.device)
    changed = torch.zeros(1, dtype=torch.int32, device=image.device)
    _lib.check(lib.fnnu_pre_filled_mask(_ptr(image), int(image.shape[0]), _lib.i3(image.shape[1:]),
                                        _lib.i3(transpose_forward), _lib.i3([b[0] for b in bbox]), _lib.i3(crop),
                                        _ptr(state), _ptr(changed), 4 * max(crop) + 16, _lib.stream_ptr()))
    return state


def channel_stats(channel: torch.Tensor, transpose_forward, bbox, state: Optional[torch.Tensor]):
    """(mean, std, min, max) of one cropped channel (inside the mask when given) — float64 sums on the device."""
    lib = _lib.load()
    out = (C.c_double * 5)()
    _lib.check(lib.fnnu_pre_channel_stats(_ptr(channel), _lib.i3(channel.shape), _lib.i3(transpose_forward),
                                          _lib.i3([b[0] for b in bbox]), _lib.i3([b[1] - b[0] for b in bbox]),
                                          _ptr(state), out, _lib.stream_ptr()))
    s, ss, n, mn, mx = (float(v) for v in out)
    if n == 0:
        return 0.0, 0.0, 0.0, 0.0
    mean = s / n
    var = max(ss / n - mean * mean, 0.0)
    return mean, float(np.sqrt(var)), mn, mx


def crop_normalize_channel(channel: torch.Tensor, transpose_forward, bbox, scheme: str, use_mask: bool,
                           intensity_props: Optional[dict], state: Optional[torch.Tensor], out: torch.Tensor):
    """default_normalization_schemes.py:27-95 for one channel; float32 arithmetic as numpy performs it."""
    lib = _lib.load()
    mode, a, b, lo, hi = 0, 0.0, 1.0, 0.0, 0.0
    st = None
    if scheme == 'CTNormalization':
        assert intensity_props is not None, 'CTNormalization requires intensity properties'
        mode = 1
        a = float(np.float32(intensity_props['mean']))
        b = float(np.float32(max(intensity_props['std'], 1e-8)))
        lo = float(np.float32(intensity_props['percentile_00_5']))
        hi = float(np.float32(intensity_props['percentile_99_5']))
    elif scheme == 'ZScoreNormalization':
        masked = bool(use_mask)
        mean, std, _, _ = channel_stats(channel, transpose_forward, bbox, state if masked else None)
        mode = 3 if masked else 2
        st = state if masked else None
        a = float(np.float32(mean))
        b = float(max(np.float32(std), 1e-8))
    elif scheme == 'RescaleTo01Normalization':
        _, _, mn, mx = channel_stats(channel, transpose_forward, bbox, None)
        mode = 2
        a = float(np.float32(mn))
        b = float(np.clip(np.float32(mx) - np.float32(mn), a_min=1e-8, a_max=None))
    elif scheme == 'RGBTo01Normalization':
        mode, a, b = 2, 0.0, 255.0
    elif scheme == 'NoNormalization':
        mode = 0
    else:
        raise NotImplementedError(f'normalization scheme {scheme}')
    _lib.check(lib.fnnu_pre_crop_normalize(_ptr(channel), _lib.i3(channel.shape), _lib.i3(transpose_forward),
                                           _lib.i3([q[0] for q in bbox]), _lib.i3([q[1] - q[0] for q in bbox]), mode,
                                           C.c_float(a), C.c_float(b), C.c_float(lo), C.c_float(hi), _ptr(st), _ptr(out),
                                           _lib.stream_ptr()))


def resample_channels(data: torch.Tensor, new_shape, current_spacing, new_spacing, order: int = 3, order_z: int = 0,
                      force_separate_z: Optional[bool] = None) -> torch.Tensor:
    """resample_data_or_seg_to_shape (default_resampling.py:89-192, is_seg=False) channel by channel."""
    lib = _lib.load()
    in_shape = tuple(int(v) for v in data.shape[1:])
    new_shape = tuple(int(v) for v in new_shape)
    if in_shape == new_shape:
        return data
    if order not in (1, 3):
        raise NotImplementedError(f'resampling order {order} (the B200 path implements orders 1 and 3)')
    do_sep, axis = determine_do_sep_z_and_axis(force_separate_z, current_spacing, new_spacing)
    base = 0 if order == 3 else 2
    modes = [base, base, base]
    if do_sep and axis is not None:
        if order_z != 0:
            raise NotImplementedError('order_z != 0 is not on the B200 pre-processing path')
        modes[axis] = 1
    ws_bytes = int(lib.fnnu_pre_resample_workspace_bytes(_lib.i3(in_shape), _lib.i3(modes)))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=data.device)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    out = torch.empty((data.shape[0], *new_shape), dtype=torch.float32, device=data.device)
    for c in range(data.shape[0]):
        _lib.check(lib.fnnu_pre_resample_channel(_ptr(data[c]), _lib.i3(in_shape), _lib.i3(new_shape), _lib.i3(modes), 1,
                                                 C.c_void_p(ws_ptr), ws_bytes, _ptr(out[c]), _lib.stream_ptr()))
    return out


def run_case_npy(image, properties: dict, seg_prev, plans_manager, configuration_manager, dataset_json, label_manager,
                 device: torch.device):
    """default_preprocessor.py:45-118 (seg=None) + the cascade's one-hot previous-stage channels
    (data_iterators.py:33-39, same-geometry case).  Returns (float32 device tensor (c, x, y, z), properties)."""
    if isinstance(image, torch.Tensor):
        raw = image.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    else:
        raw = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device, non_blocking=True)
    assert raw.ndim == 4, 'input_image must be (c, x, y, z)'
    tf = [int(t) for t in plans_manager.transpose_forward]
    original_spacing = [properties['spacing'][i] for i in tf]
    props = dict(properties)
    tdims = [int(raw.shape[1 + t]) for t in tf]
    props['shape_before_cropping'] = tuple(tdims)
    with torch.cuda.device(device):
        bbox = nonzero_bbox(raw, tf)
        props['bbox_used_for_cropping'] = bbox
        crop = tuple(b[1] - b[0] for b in bbox)
        props['shape_after_cropping_and_before_resampling'] = crop
        target_spacing = list(configuration_manager.spacing)
        if len(target_spacing) < 3:
            target_spacing = [original_spacing[0]] + target_spacing
        new_shape = compute_new_shape(crop, original_spacing, target_spacing)
        schemes = configuration_manager.normalization_schemes
        use_mask = configuration_manager.use_mask_for_norm
        need_mask = any(s == 'ZScoreNormalization' and bool(m) for s, m in zip(schemes, use_mask))
        state = filled_mask_state(raw, tf, bbox) if need_mask else None
        data = torch.empty((raw.shape[0], *crop), dtype=torch.float32, device=device)
        ipp = plans_manager.foreground_intensity_properties_per_channel
        for c in range(raw.shape[0]):
            crop_normalize_channel(raw[c], tf, bbox, schemes[c], use_mask[c], (ipp or {}).get(str(c)), state, data[c])
        kw = configuration_manager.resampling_fn_data_kwargs
        data = resample_channels(data, new_shape, original_spacing, target_spacing, kw.get('order', 3),
                                 kw.get('order_z', 0), kw.get('force_separate_z', None))
        if seg_prev is not None:
            if tuple(new_shape) != crop:
                raise NotImplementedError('cascade input with resampling is not on the B200 pre-processing path')
            sp = torch.as_tensor(np.asarray(seg_prev)).to(device).permute([0, *[i + 1 for i in tf]])
            sp = sp[(slice(None), *[slice(b[0], b[1]) for b in bbox])]
            onehot = torch.stack([(sp[0] == l) for l in label_manager.foreground_labels]).to(torch.float32)
            data = torch.cat([data, onehot], 0)
    return data, props
