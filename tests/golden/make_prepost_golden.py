"""Golden vectors for the callers either side of the hot path, produced by
EXECUTING the reference's own source files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_prepost_golden.py

The files are imported by path; the third-party modules they import but that are absent from this image are replaced
by stubs.  What each golden section therefore pins:

  resampling_decisions   default_resampling.py get_do_separate_z / get_lowres_axis / compute_new_shape /
                         determine_do_sep_z_and_axis — reference code only (no stub is reached)
  resampling_arrays      default_resampling.py resample_data_or_seg_to_shape — the reference's control flow (separate-z
                         loops, z pass through scipy.ndimage.map_coordinates, dtype handling) with
                         skimage.transform.resize STUBBED by scipy.ndimage.zoom(mode='nearest', grid_mode=True) + clip,
                         the implementation scikit-image >= 0.19 documents for itself; skimage stays unpinned
  normalization          default_normalization_schemes.py — reference code only (numpy)
  nonzero_mask           cropping.py create_nonzero_mask — reference code + scipy.ndimage.binary_fill_holes
  crop_to_nonzero        cropping.py crop_to_nonzero — with acvl_utils' get_bbox_from_mask / bounding_box_to_slice STUBBED
                         (restated from the published package); pins everything but those two helpers
  preprocessor           default_preprocessor.py DefaultPreprocessor.run_case_npy (no segmentation) — the reference's flow
                         (transpose, crop_to_nonzero, normalise BEFORE resampling, order-3 data resampling, recorded
                         properties); leans on the skimage and bounding-box stubs above
  export                 export_prediction.py convert_predicted_logits_to_segmentation_with_correct_shape — the reference's
                         function end to end (resample with the plans' default resampling_fn_probabilities, argmax,
                         un-crop, inverse transpose); leans on the skimage stub above and on a restated
                         acvl_utils insert_crop_into_image
  label_manager          label_handling.py LabelManager — properties, convert_logits_to_segmentation (argmax and
                         region thresholds), reference code + torch

Writes tests/golden/prepost_golden.npz (+ .json for the scalar tables).
"""
import importlib.util
import json
import os
import re
import sys
import types

import numpy as np
import torch
from scipy.ndimage import zoom

REF = '/root/reference/distillation/nnunetv2'
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def skimage_resize_stub(image, output_shape, order=1, mode='edge', anti_aliasing=False, clip=True, **kwargs):
    """skimage.transform.resize as scikit-image >= 0.19 implements it for the arguments the reference passes
    (mode='edge', anti_aliasing=False): scipy.ndimage.zoom with grid_mode=True, mode='nearest', then clipping to the
    input range (clip=True is skimage's default)."""
    assert mode == 'edge' and not anti_aliasing and not kwargs
    image = np.asarray(image, dtype=np.float64)
    if tuple(image.shape) == tuple(int(s) for s in output_shape):
        out = image.copy()
    else:
        factors = [float(n) / float(o) for n, o in zip(output_shape, image.shape)]
        out = zoom(image, factors, order=order, mode='nearest', grid_mode=True)
    if clip and order > 1:
        out = np.clip(out, image.min(), image.max())
    return out


def install_stubs():
    cfg_src = open(os.path.join(REF, 'configuration.py')).read()
    aniso = int(re.search(r'^ANISO_THRESHOLD\s*=\s*(\d+)', cfg_src, re.M).group(1))
    nproc = int(re.search(r'default_num_processes\s*=\s*(\d+)', cfg_src).group(1))
    _module('nnunetv2')
    _module('nnunetv2.configuration', ANISO_THRESHOLD=aniso, default_num_processes=nproc)
    _module('nnunetv2.utilities')
    _module('nnunetv2.utilities.find_class_by_name', recursive_find_python_class=lambda *a, **k: None)
    helpers = _load(os.path.join(REF, 'utilities/helpers.py'), 'nnunetv2.utilities.helpers')
    sys.modules['nnunetv2.utilities.helpers'] = helpers
    _module('batchgenerators')
    _module('batchgenerators.augmentations')
    _module('batchgenerators.augmentations.utils', resize_segmentation=None)
    _module('batchgenerators.utilities')
    _module('batchgenerators.utilities.file_and_folder_operations', join=os.path.join)
    _module('skimage')
    _module('skimage.transform', resize=skimage_resize_stub)

    # acvl_utils.cropping_and_padding.bounding_boxes, restated from the published package (stubs, see the docstring)
    def get_bbox_from_mask(mask):
        bbox = []
        for ax in range(mask.ndim):
            other = tuple(i for i in range(mask.ndim) if i != ax)
            idx = np.where(np.any(mask, axis=other))[0]
            bbox.append([int(idx[0]), int(idx[-1]) + 1] if len(idx) else [0, 0])
        return bbox

    def bounding_box_to_slice(bbox):
        return tuple(slice(*b) for b in bbox)

    _module('acvl_utils')
    _module('acvl_utils.cropping_and_padding')
    def insert_crop_into_image(image, crop, bbox):
        """restated from the published acvl_utils for a crop that lies inside the image (the export call)"""
        sl = tuple([slice(None)] * (image.ndim - len(bbox)) + [slice(int(a), int(b)) for a, b in bbox])
        image[sl] = crop if not isinstance(crop, torch.Tensor) else crop.numpy()
        return image

    _module('acvl_utils.cropping_and_padding.bounding_boxes', get_bbox_from_mask=get_bbox_from_mask,
            bounding_box_to_slice=bounding_box_to_slice, insert_crop_into_image=insert_crop_into_image)
    _module('batchgenerators.utilities.file_and_folder_operations', join=os.path.join, load_json=None, save_pickle=None)
    _module('nnunetv2.training')
    _module('nnunetv2.training.dataloading')
    _module('nnunetv2.training.dataloading.nnunet_dataset', nnUNetDatasetBlosc2=object)
    _module('nnunetv2.utilities.plans_handling')
    _module('nnunetv2.utilities.plans_handling.plans_handler', PlansManager=object, ConfigurationManager=object)
    _module('nnunetv2.utilities.label_handling')
    _module('nnunetv2.inference')
    from typing import List, Tuple, Union
    _module('SimpleITK')
    _module('batchgenerators.utilities.file_and_folder_operations', join=os.path.join, load_json=None, save_pickle=None,
            isdir=os.path.isdir, isfile=os.path.isfile, List=List, Tuple=Tuple, Union=Union, os=os,
            __all__=['join', 'load_json', 'save_pickle', 'isdir', 'isfile', 'List', 'Tuple', 'Union', 'os'])
    _module('nnunetv2.paths', nnUNet_preprocessed=None, nnUNet_raw=None)
    _module('nnunetv2.utilities.dataset_name_id_conversion', maybe_convert_to_dataset_name=None)
    _module('nnunetv2.utilities.utils', get_filenames_of_train_images_and_targets=None)
    _module('nnunetv2.preprocessing')
    _module('nnunetv2.preprocessing.cropping')
    _module('nnunetv2.preprocessing.resampling')
    return aniso


def main():
    aniso = install_stubs()
    res = _load(os.path.join(REF, 'preprocessing/resampling/default_resampling.py'), 'ref_resampling')
    norm = _load(os.path.join(REF, 'preprocessing/normalization/default_normalization_schemes.py'), 'ref_norm')
    crop = _load(os.path.join(REF, 'preprocessing/cropping/cropping.py'), 'ref_cropping')
    lab = _load(os.path.join(REF, 'utilities/label_handling/label_handling.py'), 'nnunetv2.utilities.label_handling.label_handling')
    sys.modules['nnunetv2.utilities.label_handling.label_handling'] = lab
    exp = _load(os.path.join(REF, 'inference/export_prediction.py'), 'ref_export')
    sys.modules['nnunetv2.preprocessing.cropping.cropping'] = crop
    sys.modules['nnunetv2.preprocessing.resampling.default_resampling'] = res
    sys.modules['nnunetv2'].__path__ = ['/nonexistent']
    sys.modules['nnunetv2.utilities.find_class_by_name'].recursive_find_python_class = \
        lambda folder, class_name, current_module: getattr(norm, class_name, None)
    prep = _load(os.path.join(REF, 'preprocessing/preprocessors/default_preprocessor.py'), 'ref_preprocessor')

    arrays, table = {}, {'aniso_threshold': aniso}

    # ---- resampling decisions ------------------------------------------------------------------------------------
    spacings = [(1.0, 1.0, 1.0), (3.0, 0.7, 0.7), (3.01, 1.0, 1.0), (2.0, 0.977, 0.977), (0.24, 1.25, 1.25),
                (1.5, 1.5, 5.0), (1.0, 4.0, 1.0), (5.0, 5.0, 1.0), (0.5, 0.5, 0.5), (2.5, 0.8, 0.8), (1.0, 1.0, 3.0),
                (0.8, 0.8, 2.4), (0.8, 0.8, 2.41)]
    decisions = []
    for cur in spacings:
        for new in [(1.0, 1.0, 1.0), (3.0, 0.7, 0.7), (0.8, 0.8, 2.5), (1.0, 4.0, 1.0)]:
            for force in (None, True, False):
                do, ax = res.determine_do_sep_z_and_axis(force, cur, new)
                decisions.append({'current': list(cur), 'new': list(new), 'force': force, 'do_separate_z': bool(do),
                                  'axis': None if ax is None else int(ax)})
    table['resampling_decisions'] = decisions
    table['lowres_axis'] = [{'spacing': list(s), 'axis': [int(a) for a in res.get_lowres_axis(s)]} for s in spacings]
    shapes = []
    for shp, old, new in [((155, 240, 240), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0)), ((64, 512, 512), (3.0, 0.7, 0.7), (2.0, 0.977, 0.977)),
                          ((37, 53, 61), (1.3, 0.9, 0.75), (1.0, 1.0, 1.0)), ((41, 41, 41), (1.0, 1.0, 1.0), (2.0, 2.0, 2.0)),
                          ((21, 33, 45), (1.0, 1.0, 1.0), (0.7, 0.7, 0.7)), ((7, 9, 11), (2.5, 2.5, 2.5), (1.0, 1.0, 1.0))]:
        shapes.append({'shape': list(shp), 'old': list(old), 'new': list(new),
                       'new_shape': [int(v) for v in res.compute_new_shape(shp, old, new)]})
    table['compute_new_shape'] = shapes

    # ---- resampled arrays (reference control flow, scipy-backed resize stub) ------------------------------------
    rng = np.random.default_rng(7)
    cases = [
        # (name, data shape, new shape, current spacing, new spacing, order, order_z, force_separate_z, dtype)
        ('iso_up_o3', (2, 9, 11, 13), (14, 17, 16), (1.5, 1.5, 1.5), (1.0, 1.0, 1.0), 3, 0, None, np.float32),
        ('iso_down_o1', (3, 12, 10, 14), (7, 6, 9), (1.0, 1.0, 1.0), (1.7, 1.7, 1.6), 1, 0, None, np.float16),
        ('sepz_axis0_o3', (1, 5, 12, 14), (9, 17, 19), (4.0, 1.0, 1.0), (2.0, 0.7, 0.7), 3, 0, None, np.float32),
        ('sepz_axis2_o1', (2, 10, 12, 4), (15, 16, 7), (0.8, 0.8, 3.0), (0.6, 0.6, 1.5), 1, 0, None, np.float16),
        ('sepz_axis1_forced', (1, 8, 5, 9), (11, 5, 13), (1.0, 2.5, 1.0), (0.8, 2.5, 0.8), 3, 0, True, np.float32),
        ('sepz_samez_o3', (1, 6, 10, 10), (6, 15, 14), (5.0, 1.0, 1.0), (5.0, 0.7, 0.7), 3, 0, None, np.float32),
        ('identity', (2, 6, 7, 8), (6, 7, 8), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), 3, 0, None, np.float32),
    ]
    meta = []
    for name, shp, new_shape, cur, new, order, order_z, force, dt in cases:
        x = (rng.standard_normal(shp) * 3 + rng.integers(-2, 3, size=(shp[0], 1, 1, 1))).astype(dt)
        y = res.resample_data_or_seg_to_shape(x, new_shape, cur, new, is_seg=False, order=order, order_z=order_z,
                                              force_separate_z=force)
        arrays[f'res_{name}_in'] = x
        arrays[f'res_{name}_out'] = np.asarray(y)
        meta.append({'name': name, 'new_shape': list(new_shape), 'current': list(cur), 'new': list(new), 'order': order,
                     'order_z': order_z, 'force': force, 'out_dtype': str(np.asarray(y).dtype)})
    table['resampling_arrays'] = meta

    # ---- normalisation schemes --------------------------------------------------------------------------------------
    props = {'mean': 87.3, 'std': 41.9, 'percentile_00_5': -48.0, 'percentile_99_5': 212.0}
    img = (rng.standard_normal((9, 10, 11)) * 90 + 60).astype(np.float32)
    seg = np.where(rng.random((9, 10, 11)) > 0.3, 1, -1).astype(np.int8)
    img_u8 = rng.integers(0, 256, size=(6, 7, 8)).astype(np.uint8)
    arrays['norm_img'], arrays['norm_seg'], arrays['norm_img_u8'] = img, seg, img_u8
    arrays['norm_ct'] = norm.CTNormalization(use_mask_for_norm=False, intensityproperties=props).run(img.copy(), seg)
    arrays['norm_zscore'] = norm.ZScoreNormalization(use_mask_for_norm=False, intensityproperties={}).run(img.copy(), seg)
    arrays['norm_zscore_mask'] = norm.ZScoreNormalization(use_mask_for_norm=True, intensityproperties={}).run(img.copy(), seg)
    arrays['norm_none'] = norm.NoNormalization(use_mask_for_norm=False, intensityproperties={}).run(img.copy(), seg)
    arrays['norm_rescale01'] = norm.RescaleTo01Normalization(use_mask_for_norm=False, intensityproperties={}).run(img.copy(), seg)
    arrays['norm_rgb01'] = norm.RGBTo01Normalization(use_mask_for_norm=False, intensityproperties={}).run(img_u8.copy(), None)
    table['norm_props'] = props

    # ---- non-zero mask / crop ---------------------------------------------------------------------------------------
    vol = np.zeros((2, 14, 16, 15), dtype=np.float32)
    vol[0, 3:11, 4:13, 2:12] = rng.standard_normal((8, 9, 10))
    vol[0, 5:8, 6:10, 5:8] = 0                # an enclosed hole: binary_fill_holes closes it
    vol[1, 2:5, 9:14, 8:13] = 1.0
    vol[0, 12, 1, 1] = 0.5                    # an isolated voxel widens the box
    arrays['crop_in'] = vol
    arrays['crop_mask'] = crop.create_nonzero_mask(vol)
    d, s, bbox = crop.crop_to_nonzero(vol.copy(), None, nonzero_label=-1)
    arrays['crop_data'], arrays['crop_seg'] = d, s
    table['crop_bbox'] = [[int(a), int(b)] for a, b in bbox]

    # ---- LabelManager -----------------------------------------------------------------------------------------------
    label_cases = {
        'plain': ({'background': 0, 'liver': 1, 'tumor': 2, 'vessel': 3}, None),
        'ignore': ({'background': 0, 'organ': 1, 'lesion': 2, 'ignore': 3}, None),
        'regions': ({'background': 0, 'whole_tumor': [1, 2, 3], 'tumor_core': [2, 3], 'enhancing_tumor': [3]}, [1, 2, 3]),
    }
    lm_table = {}
    for name, (ld, order) in label_cases.items():
        lm = lab.LabelManager(ld, order)
        heads = lm.num_segmentation_heads
        logits = torch.from_numpy(rng.standard_normal((heads, 6, 7, 8)).astype(np.float32) * 2)
        if not lm.has_regions:
            logits[1, 0, 0, :4] = logits[0, 0, 0, :4]          # exact ties: the first maximum wins
            logits[:, 1, 1, 1] = 0.25
        seg_out = lm.convert_logits_to_segmentation(logits)
        arrays[f'lm_{name}_logits'] = logits.numpy()
        arrays[f'lm_{name}_seg'] = np.asarray(seg_out)
        lm_table[name] = {
            'label_dict': ld, 'regions_class_order': order, 'has_regions': bool(lm.has_regions),
            'has_ignore_label': bool(lm.has_ignore_label), 'ignore_label': lm.ignore_label,
            'all_labels': [int(i) for i in lm.all_labels],
            'all_regions': None if lm.all_regions is None else [list(r) if isinstance(r, (tuple, list)) else int(r) for r in lm.all_regions],
            'foreground_labels': [int(i) for i in lm.foreground_labels],
            'foreground_regions': None if not lm.has_regions else [list(r) if isinstance(r, (tuple, list)) else int(r) for r in lm.foreground_regions],
            'num_segmentation_heads': int(heads), 'seg_dtype': str(np.asarray(seg_out).dtype)}
    table['label_manager'] = lm_table

    # ---- export: logits -> label map in the original geometry ---------------------------------------------------------
    from functools import partial

    class _PM:
        def __init__(self, tf):
            self.transpose_forward = list(tf)
            self.transpose_backward = [int(i) for i in np.argsort(tf)]

    class _CM:
        def __init__(self, spacing):
            self.spacing = list(spacing)
            # the plans' default: "resampling_fn_probabilities": "resample_data_or_seg_to_shape",
            # kwargs {"is_seg": false, "order": 1, "order_z": 0, "force_separate_z": null}
            self.resampling_fn_probabilities = partial(res.resample_data_or_seg_to_shape, is_seg=False, order=1, order_z=0,
                                                       force_separate_z=None)

    export_cases = [
        # name, logits shape, plans spacing, transpose_forward, original spacing (file order), shape before cropping
        # (transposed), bbox, label dict
        ('iso', (3, 10, 12, 9), (1.5, 1.5, 1.5), (0, 1, 2), (1.0, 1.0, 1.0), (19, 20, 16), [[2, 17], [1, 19], [0, 14]],
         {'background': 0, 'a': 1, 'b': 2}),
        ('sepz_transposed', (2, 6, 14, 12), (3.0, 0.8, 0.8), (2, 0, 1), (0.6, 0.6, 4.5), (11, 24, 20), [[1, 10], [2, 23], [0, 18]],
         {'background': 0, 'a': 1}),
        ('same_shape', (4, 7, 8, 9), (1.0, 1.0, 1.0), (1, 2, 0), (1.0, 1.0, 1.0), (9, 8, 12), [[1, 8], [0, 8], [2, 11]],
         {'background': 0, 'a': 1, 'b': 2, 'c': 3}),
    ]
    exp_meta = []
    for name, lshape, pspacing, tf, ospacing, before, bbox, ld in export_cases:
        logits = (rng.standard_normal(lshape) * 2).astype(np.float16)
        mid = [b - a for a, b in bbox]
        props = {'spacing': list(ospacing), 'shape_after_cropping_and_before_resampling': mid,
                 'shape_before_cropping': list(before), 'bbox_used_for_cropping': bbox}
        lm = lab.LabelManager(ld, None)
        pm = _PM(tf)
        seg_out = exp.convert_predicted_logits_to_segmentation_with_correct_shape(logits.copy(), pm, _CM(pspacing), lm, props,
                                                                                  return_probabilities=False)
        arrays[f'exp_{name}_logits'] = logits
        arrays[f'exp_{name}_seg'] = np.asarray(seg_out)
        exp_meta.append({'name': name, 'plans_spacing': list(pspacing), 'transpose_forward': list(tf),
                         'transpose_backward': pm.transpose_backward, 'properties': props,
                         'num_foreground': len(lm.foreground_labels), 'seg_dtype': str(np.asarray(seg_out).dtype),
                         'seg_shape': list(np.asarray(seg_out).shape)})
    table['export'] = exp_meta

    # ---- DefaultPreprocessor.run_case_npy ------------------------------------------------------------------------------
    class _PPM:
        def __init__(self, tf, props_per_channel):
            self.transpose_forward = list(tf)
            self.foreground_intensity_properties_per_channel = props_per_channel

    class _PCM:
        def __init__(self, spacing, schemes, use_mask):
            self.spacing = list(spacing)
            self.normalization_schemes = list(schemes)
            self.use_mask_for_norm = list(use_mask)
            # plans defaults: resampling_fn_data = resample_data_or_seg_to_shape(is_seg=False, order=3, order_z=0,
            # force_separate_z=None); the segmentation branch is not part of the inference path
            self.resampling_fn_data = partial(res.resample_data_or_seg_to_shape, is_seg=False, order=3, order_z=0,
                                              force_separate_z=None)
            self.resampling_fn_seg = lambda seg, new_shape, *a, **k: np.zeros((seg.shape[0], *new_shape), dtype=seg.dtype)

    pre_cases = [
        # name, raw shape (c, file order), spacing (file order), transpose_forward, target spacing, schemes, use_mask
        ('ct_iso', (1, 14, 16, 15), (1.0, 1.0, 1.0), (0, 1, 2), (1.4, 1.4, 1.4), ['CTNormalization'], [False]),
        ('mri_aniso_transposed', (2, 12, 13, 6), (0.8, 0.8, 3.5), (2, 0, 1), (3.5, 1.1, 1.1), ['ZScoreNormalization', 'ZScoreNormalization'], [True, True]),
        ('no_resampling', (1, 9, 10, 11), (1.0, 1.0, 1.0), (0, 1, 2), (1.0, 1.0, 1.0), ['ZScoreNormalization'], [False]),
    ]
    pre_meta = []
    ct_props = {'0': {'mean': 0.4, 'std': 1.3, 'percentile_00_5': -1.5, 'percentile_99_5': 2.5},
                '1': {'mean': 0.0, 'std': 1.0, 'percentile_00_5': -2.0, 'percentile_99_5': 2.0}}
    for name, shp, spacing, tf, target, schemes, use_mask in pre_cases:
        raw = rng.standard_normal(shp).astype(np.float32) + 0.5
        raw[:, :2] = 0                          # a zero border on the first file axis: cropped away
        raw[:, -1] = 0
        raw[:, :, :, :1] = 0
        raw[:, 4:7, 5:8, 2:4] = 0               # enclosed zeros: stay inside the mask (fill-holes)
        props_in = {'spacing': list(spacing)}
        data, seg, props = prep.DefaultPreprocessor(verbose=False).run_case_npy(
            raw.copy(), None, dict(props_in), _PPM(tf, ct_props), _PCM(target, schemes, use_mask), {})
        arrays[f'pre_{name}_raw'] = raw
        arrays[f'pre_{name}_data'] = np.asarray(data)
        pre_meta.append({'name': name, 'spacing': list(spacing), 'transpose_forward': list(tf), 'target_spacing': list(target),
                         'schemes': schemes, 'use_mask': use_mask, 'props_per_channel': ct_props,
                         'shape_before_cropping': [int(v) for v in props['shape_before_cropping']],
                         'bbox_used_for_cropping': [[int(a), int(b)] for a, b in props['bbox_used_for_cropping']],
                         'shape_after_cropping_and_before_resampling': [int(v) for v in props['shape_after_cropping_and_before_resampling']],
                         'data_dtype': str(np.asarray(data).dtype), 'data_shape': list(np.asarray(data).shape)})
    table['preprocessor'] = pre_meta

    np.savez_compressed(os.path.join(HERE, 'prepost_golden.npz'), **arrays)
    with open(os.path.join(HERE, 'prepost_golden.json'), 'w') as f:
        json.dump(table, f, indent=1)
    size = os.path.getsize(os.path.join(HERE, 'prepost_golden.npz'))
    print(f'wrote prepost_golden.npz ({size / 1024:.0f} KB, {len(arrays)} arrays) and prepost_golden.json')


if __name__ == '__main__':
    main()
