"""CPU: the oracle restatements and the product's host logic for pre-processing, export and label handling against
golden vectors produced by EXECUTING the reference's own files (tests/golden/make_prepost_golden.py says what each
section pins and which third-party helpers were stubbed)."""
import json
import os

import numpy as np
import pytest
import torch

from fast_nnunet_b200 import export as P_export
from fast_nnunet_b200 import plans as P_plans
from fast_nnunet_b200 import preprocess as P_pre
from oracle import export as O_export
from oracle import preprocess as O_pre

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'prepost_golden.npz'))
T = json.load(open(os.path.join(HERE, 'golden', 'prepost_golden.json')))


def test_separate_z_decisions_oracle_and_product():
    assert O_export.ANISO_THRESHOLD == T['aniso_threshold']
    for d in T['resampling_decisions']:
        want = (d['do_separate_z'], d['axis'])
        got_o = O_export.determine_do_sep_z_and_axis(d['force'], d['current'], d['new'])
        got_p = P_export.determine_do_sep_z_and_axis(d['force'], d['current'], d['new'])
        assert (bool(got_o[0]), None if got_o[1] is None else int(got_o[1])) == want, d
        assert (bool(got_p[0]), None if got_p[1] is None else int(got_p[1])) == want, d
    for d in T['lowres_axis']:
        assert [int(a) for a in O_export.get_lowres_axis(d['spacing'])] == d['axis'], d


def test_compute_new_shape_oracle_and_product():
    for d in T['compute_new_shape']:
        assert [int(v) for v in O_pre.compute_new_shape(d['shape'], d['old'], d['new'])] == d['new_shape'], d
        assert [int(v) for v in P_pre.compute_new_shape(d['shape'], d['old'], d['new'])] == d['new_shape'], d


@pytest.mark.parametrize('case', T['resampling_arrays'], ids=lambda c: c['name'])
def test_oracle_resampling_equals_reference_control_flow(case):
    """The reference's resample_data_or_seg_to_shape (separate-z loops, map_coordinates z pass, dtype of the result) with
    a scipy-backed skimage.resize: oracle/preprocess.py (order 3 path of f2) and oracle/export.py (order 1 path of f1)
    must give the same arrays."""
    x, want = G[f"res_{case['name']}_in"], G[f"res_{case['name']}_out"]
    assert str(want.dtype) == case['out_dtype']
    got = O_pre.resample_data(x.copy(), case['new_shape'], case['current'], case['new'], order=case['order'],
                              order_z=case['order_z'], force_separate_z=case['force'])
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)
    if case['order'] == 1:          # no overshoot, so export's un-clipped resize is the same arithmetic
        got = O_export.resample_data_or_seg_to_shape(x.copy(), case['new_shape'], case['current'], case['new'],
                                                     order=1, order_z=case['order_z'], force_separate_z=case['force'])
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)


def test_oracle_normalisation_schemes():
    img, seg, u8 = G['norm_img'], G['norm_seg'], G['norm_img_u8']
    props = {'0': T['norm_props']}

    def run(scheme, use_mask, data):
        d = data.astype(np.float32)[None].copy()
        return O_pre.normalize(d, seg[None], [scheme], [use_mask], props)[0]

    np.testing.assert_array_equal(run('CTNormalization', False, img), G['norm_ct'])
    np.testing.assert_array_equal(run('ZScoreNormalization', False, img), G['norm_zscore'])
    np.testing.assert_array_equal(run('ZScoreNormalization', True, img), G['norm_zscore_mask'])
    np.testing.assert_array_equal(run('NoNormalization', False, img), G['norm_none'])
    np.testing.assert_array_equal(run('RescaleTo01Normalization', False, img), G['norm_rescale01'])
    np.testing.assert_array_equal(run('RGBTo01Normalization', False, u8), G['norm_rgb01'])


def test_oracle_nonzero_mask_and_crop():
    vol = G['crop_in']
    np.testing.assert_array_equal(O_pre.create_nonzero_mask(vol), G['crop_mask'])
    data, seg, bbox = O_pre.crop_to_nonzero(vol.copy(), nonzero_label=-1)
    assert [[int(a), int(b)] for a, b in bbox] == T['crop_bbox']
    np.testing.assert_array_equal(data, G['crop_data'])
    np.testing.assert_array_equal(seg, G['crop_seg'])


@pytest.mark.parametrize('name', sorted(T['label_manager']))
def test_label_manager_equals_reference(name):
    d = T['label_manager'][name]
    lm = P_plans.LabelManager(d['label_dict'], d['regions_class_order'])
    assert bool(lm.has_regions) == d['has_regions']
    assert bool(lm.has_ignore_label) == d['has_ignore_label']
    assert lm.ignore_label == d['ignore_label']
    assert [int(i) for i in lm.all_labels] == d['all_labels']
    assert [int(i) for i in lm.foreground_labels] == d['foreground_labels']
    assert int(lm.num_segmentation_heads) == d['num_segmentation_heads']

    def norm(regions):
        return None if regions is None else [list(r) if isinstance(r, (tuple, list)) else int(r) for r in regions]

    assert norm(lm.all_regions) == d['all_regions']
    if d['has_regions']:
        assert norm(lm.foreground_regions) == d['foreground_regions']
    logits = torch.from_numpy(G[f'lm_{name}_logits'])
    seg = lm.convert_logits_to_segmentation(logits)
    seg = seg.numpy() if isinstance(seg, torch.Tensor) else np.asarray(seg)
    np.testing.assert_array_equal(seg.astype(np.int64), G[f'lm_{name}_seg'].astype(np.int64))


@pytest.mark.parametrize('case', T['export'], ids=lambda c: c['name'])
def test_oracle_export_equals_executed_reference(case):
    """export_prediction.convert_predicted_logits_to_segmentation_with_correct_shape, run from the reference's own file
    (plans' default order-1 probability resampling, LabelManager argmax, un-crop, inverse transpose)."""
    logits, want = G[f"exp_{case['name']}_logits"], G[f"exp_{case['name']}_seg"]
    assert str(want.dtype) == case['seg_dtype'] and list(want.shape) == case['seg_shape']
    got = O_export.convert_predicted_logits_to_segmentation_with_correct_shape(
        logits.copy(), case['plans_spacing'], case['transpose_forward'], case['transpose_backward'], case['properties'],
        num_foreground=case['num_foreground'])
    assert got.dtype == want.dtype and got.shape == want.shape
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('case', T['preprocessor'], ids=lambda c: c['name'])
def test_oracle_preprocessor_equals_executed_reference(case):
    """DefaultPreprocessor.run_case_npy, run from the reference's own file: transpose, crop to the filled non-zero mask,
    normalise before resampling, order-3 resampling, and the properties the export needs later."""
    raw, want = G[f"pre_{case['name']}_raw"], G[f"pre_{case['name']}_data"]
    assert str(want.dtype) == case['data_dtype'] and list(want.shape) == case['data_shape']
    data, props = O_pre.run_case_npy(raw.copy(), {'spacing': case['spacing']}, case['transpose_forward'],
                                     case['target_spacing'], case['schemes'], case['use_mask'], case['props_per_channel'])
    assert [int(v) for v in props['shape_before_cropping']] == case['shape_before_cropping']
    assert [[int(a), int(b)] for a, b in props['bbox_used_for_cropping']] == case['bbox_used_for_cropping']
    assert [int(v) for v in props['shape_after_cropping_and_before_resampling']] == case['shape_after_cropping_and_before_resampling']
    assert data.dtype == want.dtype and data.shape == want.shape
    np.testing.assert_array_equal(data, want)
