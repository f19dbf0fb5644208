"""f1: logits -> label map export.
CPU: the kernel's per-axis sampling rule (fast_nnunet_b200.export.source_coordinates, the host mirror of
csrc/export_kernels.cu) against the oracle, which executes scipy.ndimage.zoom / map_coordinates as the reference's
resample_data_or_seg does; the separate-z decision against the oracle's restatement.
GPU: fnnu_export_labels against the oracle, bit-exact label maps (integer result)."""
import numpy as np
import pytest
import torch

from fast_nnunet_b200 import export as X
from oracle import export as OX


def _mirror_resample(data, new_shape, modes):
    """fp64 evaluation of the kernel's rule on the host: out = sum over the 8 corners of w * v, rounded to data.dtype."""
    out = np.asarray(data, dtype=np.float64)
    for ax in range(3):
        i0, i1, t = X.source_coordinates(out.shape[ax + 1], int(new_shape[ax]), bool(modes[ax]))
        a = np.take(out, i0, axis=ax + 1)
        b = np.take(out, i1, axis=ax + 1)
        shp = [1, 1, 1, 1]
        shp[ax + 1] = -1
        t = t.reshape(shp)
        out = a * (1.0 - t) + b * t
    return out.astype(data.dtype)


CASES = [
    # (logits shape, target shape, current spacing, new spacing)
    ((3, 20, 24, 28), (27, 31, 40), (1.5, 1.5, 1.5), (1.1, 1.16, 1.05)),        # isotropic up-sampling
    ((2, 30, 33, 21), (17, 20, 16), (1.0, 1.0, 1.0), (1.76, 1.65, 1.31)),       # down-sampling
    ((4, 12, 40, 44), (19, 50, 57), (5.0, 0.8, 0.8), (3.16, 0.64, 0.62)),       # separate z (axis 0), extent changes
    ((2, 16, 30, 26), (16, 41, 37), (4.0, 1.0, 1.0), (4.0, 0.73, 0.70)),        # separate z, z extent unchanged
    ((2, 25, 18, 30), (25, 18, 30), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0)),          # no resampling
]


@pytest.mark.parametrize('shape,target,cur,new', CASES)
def test_sampling_rule_matches_scipy(shape, target, cur, new):
    g = np.random.default_rng(0)
    logits = (g.normal(size=shape) * 4).astype(np.float16)
    want = OX.resample_data_or_seg_to_shape(logits, target, cur, new, order=1, order_z=0, force_separate_z=None)
    modes = X.axis_modes(shape[1:], target, cur, new) if tuple(shape[1:]) != tuple(target) else (0, 0, 0)
    got = _mirror_resample(logits, target, modes)
    assert got.dtype == np.float16 and got.shape == want.shape
    # identical up to the last bit of the float64 sum: compare the fp16 results, allowing a one-ulp flip on exact ties
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert (diff > 0).mean() <= 2e-4, f'{(diff > 0).mean():.2e} of voxels differ'
    assert diff.max() <= 2.0 ** -6
    assert np.mean(np.argmax(got, 0) == np.argmax(want, 0)) >= 0.9999


@pytest.mark.parametrize('cur,new', [((5.0, 0.8, 0.8), (3.0, 0.7, 0.7)), ((1.0, 1.0, 1.0), (3.5, 1.0, 1.0)),
                                     ((0.24, 1.25, 1.25), (0.24, 1.0, 1.0)), ((1.0, 1.0, 1.0), (1.2, 1.1, 1.0)),
                                     ((1.0, 4.0, 1.0), (1.0, 2.0, 1.0))])
def test_separate_z_decision_matches_oracle(cur, new):
    a = X.determine_do_sep_z_and_axis(None, cur, new)
    b = OX.determine_do_sep_z_and_axis(None, cur, new)
    assert a[0] == bool(b[0]) and (a[1] is None) == (b[1] is None)
    if a[1] is not None:
        assert a[1] == int(b[1])


GPU_CASES = CASES + [((61, 18, 22, 20), (25, 30, 33), (2.0, 0.98, 0.98), (1.44, 0.72, 0.59))]


@pytest.mark.gpu
@pytest.mark.parametrize('shape,target,cur,new', GPU_CASES)
@pytest.mark.parametrize('tf', [(0, 1, 2), (2, 0, 1), (1, 0, 2)])
def test_export_kernel_matches_oracle(shape, target, cur, new, tf):
    dev = torch.device('cuda', 0)
    g = np.random.default_rng(1)
    logits = (g.normal(size=shape) * 4).astype(np.float16)
    tb = tuple(int(i) for i in np.argsort(tf))
    canvas = tuple(int(t) + 7 + 3 * k for k, t in enumerate(target))
    bbox = [[2 + k, 2 + k + int(t)] for k, t in enumerate(target)]
    # properties are expressed in the ORIGINAL axis order: spacing[i] for i in transpose_forward == new (transposed)
    spacing = [0.0, 0.0, 0.0]
    for k, i in enumerate(tf):
        spacing[i] = new[k]
    props = {'spacing': spacing, 'shape_after_cropping_and_before_resampling': target,
             'shape_before_cropping': canvas, 'bbox_used_for_cropping': bbox}
    want = OX.convert_predicted_logits_to_segmentation_with_correct_shape(logits, cur, tf, tb, props)
    modes = X.axis_modes(shape[1:], target, cur, new) if tuple(shape[1:]) != tuple(target) else (0, 0, 0)
    got = X.export_labels(torch.from_numpy(logits).to(dev), target, modes, bbox, canvas, tb).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.uint8
    mismatch = float((got != want).mean())
    assert mismatch <= 1e-4, f'{mismatch:.2e} of voxels differ'      # exact fp16 rounding ties of the float64 sum only
    assert np.array_equal(got == 0, want == 0) or mismatch <= 1e-4
    # the integer steps are exact: everything outside the bounding box is background
    outside = np.ones(canvas, dtype=bool)
    outside[tuple(slice(b[0], b[1]) for b in bbox)] = False
    assert not got.transpose(tf)[outside].any()
