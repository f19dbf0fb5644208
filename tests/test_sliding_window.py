"""Tile starts / Gaussian map / padding: oracle and product host code against golden vectors made by
the reference's own sliding_window_prediction.py (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from fast_nnunet_b200 import sliding_window as sw
from oracle import sliding_window as osw

G = os.path.join(os.path.dirname(__file__), 'golden')
META = json.load(open(os.path.join(G, 'sliding_window_golden.json')))
NPZ = np.load(os.path.join(G, 'sliding_window_golden.npz'))


@pytest.mark.parametrize('case', META['steps'], ids=lambda c: 'x'.join(map(str, c['image'])))
def test_tile_starts_bit_exact(case):
    want = case['steps']
    assert osw.steps_for_sliding_window(case['image'], case['tile'], case['step']) == want
    assert sw.compute_steps_for_sliding_window(case['image'], case['tile'], case['step']) == want
    starts = sw.tile_starts(case['image'], case['tile'], case['step'])
    ref = [(a, b, c) for a in want[0] for b in want[1] for c in want[2]]
    assert starts.dtype == np.int32 and starts.tolist() == [list(r) for r in ref]
    # oracle slicer order == product tile order
    sl = osw.slicers_for(case['image'], case['tile'], case['step'])
    assert [[s.start for s in t[1:]] for t in sl] == starts.tolist()


def test_known_tile_counts():
    # SURVEY.md §8(a1)
    assert len(sw.tile_starts((400, 512, 512), (128,) * 3, 0.5)) == 294
    assert len(sw.tile_starts((160,) * 3, (128,) * 3, 0.5)) == 8
    assert len(sw.tile_starts((155, 240, 240), (128,) * 3, 0.5)) == 18
    assert len(sw.tile_starts((1200, 512, 512), (160, 96, 96), 0.5)) == 1400


@pytest.mark.parametrize('meta', META['gaussian'], ids=lambda m: 'x'.join(map(str, m['tile'])))
def test_gaussian_bit_exact(meta):
    tile = tuple(meta['tile'])
    big = np.prod(tile) > 64 ** 3
    ours16 = sw.compute_gaussian(tile, 1. / 8, 10, np.float16)
    ours32 = sw.compute_gaussian(tile, 1. / 8, 10, np.float32)
    assert hashlib.sha256(ours16.tobytes()).hexdigest() == meta['sha256_fp16']
    assert hashlib.sha256(ours32.tobytes()).hexdigest() == meta['sha256_fp32']
    assert float(ours16.astype(np.float32).max()) == meta['max']
    assert float(ours16.astype(np.float32).min()) == meta['min']
    assert int((ours16 == ours16.min()).sum()) == meta['n_at_floor']
    key = 'gauss16_' + 'x'.join(map(str, tile))
    if key in NPZ:
        assert np.array_equal(NPZ[key].view(np.uint16), ours16.view(np.uint16))
    if not big:   # the scipy oracle takes ~1 s per 128^3 map; keep the CPU suite short
        o16 = osw.gaussian_map(tile, 1. / 8, 10, torch.float16).numpy()
        assert hashlib.sha256(o16.tobytes()).hexdigest() == meta['sha256_fp16']


def test_gaussian_oracle_full_size():
    meta = META['gaussian'][0]
    o16 = osw.gaussian_map(tuple(meta['tile']), 1. / 8, 10, torch.float16).numpy()
    assert hashlib.sha256(o16.tobytes()).hexdigest() == meta['sha256_fp16']


@pytest.mark.parametrize('shape,patch', [((1, 20, 40, 33), (32, 32, 32)), ((2, 32, 32, 32), (32, 32, 32)),
                                         ((1, 5, 70, 31), (16, 64, 32))])
def test_padding(shape, patch):
    x = torch.arange(int(np.prod(shape)), dtype=torch.float32).reshape(shape)
    padded, slicer = osw.pad_to_patch(x, patch)
    below, above = sw.pad_amounts(shape[1:], patch)
    assert list(padded.shape[1:]) == [s + b + a for s, b, a in zip(shape[1:], below, above)]
    assert torch.equal(padded[slicer], x)
    assert all(p >= q for p, q in zip(padded.shape[1:], patch))
    assert float(padded.double().sum()) == float(x.double().sum())
    for d, (b, a) in enumerate(zip(below, above)):
        assert a - b in (0, 1)


def test_weight_sum_matches_oracle_accumulation():
    img, tile = (40, 48, 36), (32, 32, 32)
    starts = sw.tile_starts(img, tile, 0.5)
    g16 = sw.compute_gaussian(tile, 1. / 8, 10, np.float16)
    w = sw.weight_sum_map(img, tile, starts, g16, tile_order_fp16=True)
    n = torch.zeros(img, dtype=torch.half)
    g = torch.from_numpy(g16)
    for sl in osw.slicers_for(img, tile, 0.5):
        n[sl[1:]] += g
    assert np.array_equal(w.view(np.uint16), n.numpy().view(np.uint16))
