"""f2: pre-processing of one case on the device against the oracle, which executes numpy / scipy
exactly as the reference's run_case_npy does (oracle/preprocess.py).
Bars: bounding box, shapes and the fill-holes mask are integer results -> exact; CT normalisation is float32
arithmetic operation by operation -> bit-exact; z-score uses float64 sums where numpy uses float32 pairwise sums ->
relative 2e-6; cubic resampling evaluates scipy's own recurrence in float64 -> 2e-6 absolute on O(1) data (the
result is rounded to float32 in both)."""
import numpy as np
import pytest
import torch

from oracle import preprocess as OP

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


class _Cfg:
    def __init__(self, spacing, schemes, use_mask, kw=None):
        self.spacing = spacing
        self.normalization_schemes = schemes
        self.use_mask_for_norm = use_mask
        self.resampling_fn_data_kwargs = kw or {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None}


class _Plans:
    def __init__(self, tf, ipp):
        self.transpose_forward = tf
        self.transpose_backward = [int(i) for i in np.argsort(tf)]
        self.foreground_intensity_properties_per_channel = ipp


def _image(shape, channels, seed, hole=True):
    g = np.random.default_rng(seed)
    img = g.normal(100, 300, size=(channels, *shape)).astype(np.float32)
    mask = np.zeros(shape, dtype=bool)
    mask[3:-4, 2:-3, 4:-2] = True
    img[:, ~mask] = 0
    if hole:                               # an enclosed zero pocket (filled by binary_fill_holes) and an open notch
        img[:, 8:12, 8:13, 9:14] = 0
        img[:, 3:9, 10:12, 4:8] = 0
    return img


IPP = {'0': {'mean': 80.0, 'std': 250.0, 'percentile_00_5': -600.0, 'percentile_99_5': 900.0}}

CASES = [
    # shape, channels, transpose_forward, image spacing (original axis order), plans spacing, schemes, use_mask
    ((30, 34, 38), 1, [0, 1, 2], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0], ['CTNormalization'], [False]),
    ((30, 34, 38), 1, [0, 1, 2], [1.3, 0.9, 0.8], [1.0, 1.0, 1.0], ['CTNormalization'], [False]),
    ((24, 40, 36), 2, [2, 0, 1], [0.8, 1.1, 1.25], [1.0, 1.0, 1.0], ['ZScoreNormalization', 'ZScoreNormalization'], [False, True]),
    ((20, 44, 40), 1, [0, 1, 2], [5.0, 0.8, 0.8], [3.0, 1.0, 1.0], ['CTNormalization'], [False]),      # separate z
    ((20, 44, 40), 1, [0, 1, 2], [4.0, 0.7, 0.7], [4.0, 1.0, 1.0], ['ZScoreNormalization'], [True]),   # separate z, z kept
    ((28, 28, 30), 3, [1, 2, 0], [1.0, 1.0, 1.0], [1.4, 1.4, 1.4], ['RescaleTo01Normalization', 'NoNormalization',
                                                                      'ZScoreNormalization'], [False, False, False]),
]


@pytest.mark.parametrize('shape,C,tf,spacing,target,schemes,use_mask', CASES)
def test_run_case_npy_matches_oracle(shape, C, tf, spacing, target, schemes, use_mask):
    from fast_nnunet_b200 import preprocess as P
    img = _image(shape, C, seed=sum(shape))
    props = {'spacing': spacing}
    want, wprops = OP.run_case_npy(img.copy(), dict(props), tf, target, schemes, use_mask, IPP)
    got, gprops = P.run_case_npy(img, dict(props), None, _Plans(tf, IPP), _Cfg(target, schemes, use_mask), None, None, DEV)
    assert [list(b) for b in gprops['bbox_used_for_cropping']] == [list(b) for b in wprops['bbox_used_for_cropping']]
    for k in ('shape_before_cropping', 'shape_after_cropping_and_before_resampling'):
        assert tuple(gprops[k]) == tuple(wprops[k]), k
    got = got.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    scale = max(1.0, float(np.abs(want).max()))
    print(f'{schemes} {spacing}->{target}: shape {want.shape}, max|d|={d.max():.3e} (range {scale:.1f})')
    assert d.max() <= 4e-6 * scale


def test_ct_normalisation_is_bit_exact():
    from fast_nnunet_b200 import preprocess as P
    img = _image((26, 30, 34), 1, seed=5)
    want, _ = OP.run_case_npy(img.copy(), {'spacing': [1, 1, 1]}, [0, 1, 2], [1, 1, 1], ['CTNormalization'], [False], IPP)
    got, _ = P.run_case_npy(img, {'spacing': [1, 1, 1]}, None, _Plans([0, 1, 2], IPP),
                            _Cfg([1, 1, 1], ['CTNormalization'], [False]), None, None, DEV)
    assert np.array_equal(got.cpu().numpy().view(np.int32), want.view(np.int32))


def test_fill_holes_mask_matches_scipy():
    from fast_nnunet_b200 import preprocess as P
    g = np.random.default_rng(3)
    img = np.zeros((1, 40, 36, 44), dtype=np.float32)
    # a shell with enclosed pockets, tunnels to the outside and random speckle
    img[0, 5:35, 4:30, 6:40] = 1
    img[0, 10:20, 10:20, 10:30] = 0            # enclosed
    img[0, 25:30, 0:20, 15:18] = 0             # tunnel to the face
    img[0][g.random(img.shape[1:]) < 0.02] = 0
    mask = OP.create_nonzero_mask(img)
    bbox = OP.get_bbox_from_mask(mask)
    sl = tuple(slice(b[0], b[1]) for b in bbox)
    raw = torch.from_numpy(img).to(DEV)
    assert P.nonzero_bbox(raw, [0, 1, 2]) == bbox
    state = P.filled_mask_state(raw, [0, 1, 2], bbox).cpu().numpy()
    assert np.array_equal(state != 2, mask[sl])


def test_all_zero_image_keeps_its_extent():
    from fast_nnunet_b200 import preprocess as P
    raw = torch.zeros((1, 8, 9, 10), device=DEV)
    assert P.nonzero_bbox(raw, [0, 1, 2]) == [[0, 8], [0, 9], [0, 10]]
