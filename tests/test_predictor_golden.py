"""CPU: the oracle's sliding-window loop (oracle/predictor.py) against logits produced by RUNNING the reference's own
nnUNetPredictor on the CPU (tests/golden/make_predictor_golden.py: predict_from_raw_data.py imported by path, the
network an opaque torch module with seeded weights).  Pins SURVEY.md section 8 rows a2, a5, a6, a7, a10 of the oracle to
executed reference code; the `padded` case also depends on the restated pad_nd_image stub named there.

Same machine, same torch: the arrays are bit-identical.  The fixture may be checked on a host whose CPU convolution
kernels round differently, so the assertion allows fp16-ulp differences on at most 0.1 % of the values."""
import json
import os

import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import model_folder as M
from fast_nnunet_b200 import sliding_window as sw
from oracle import networks as N
from oracle import predictor as OP

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'predictor_golden.npz'))
T = json.load(open(os.path.join(HERE, 'golden', 'predictor_golden.json')))


@pytest.mark.parametrize('case', T['cases'], ids=lambda c: c['name'])
def test_oracle_loop_equals_executed_reference(case):
    spec = getattr(nets, T['net'])
    torch.set_num_threads(8)
    sds = [M.synthesize_state_dict(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], seed=4321 + f, randomize_affine=True)
           for f in range(case['folds'])]
    net = N.build_from_arch(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], allow_init=False)
    x = nets.ct_like_volume(tuple(case['volume'][1:]), case['volume'][0], seed=11)
    axes = tuple(case['axes']) if case['mirroring'] else None
    pred = None
    for sd in sds:                     # predict_from_raw_data.py:483-500: sum of the folds' fp16 logits, one division
        net.load_state_dict(sd, strict=True)
        out = OP.predict_sliding_window_return_logits(net, x, tuple(case['patch']), case['step'], case['gaussian'], axes)
        pred = out.clone() if pred is None else pred + out
    if len(sds) > 1:
        pred = pred / len(sds)
    want = G[case['name']]
    got = pred.numpy()
    assert got.dtype == want.dtype == np.float16 and got.shape == want.shape
    same = float((got == want).mean())
    print(f"{case['name']}: identical values {same:.6f}, max|d| {np.abs(got.astype(np.float32) - want.astype(np.float32)).max():.5f}")
    assert same >= 0.999
    np.testing.assert_allclose(got.astype(np.float32), want.astype(np.float32), atol=4e-3, rtol=4e-3)


@pytest.mark.parametrize('case', T['cases'], ids=lambda c: c['name'])
def test_product_tile_order_equals_executed_reference(case):
    """fast_nnunet_b200.sliding_window.tile_starts against the slicers the reference's predictor built."""
    padded = tuple(max(v, p) for v, p in zip(case['volume'][1:], case['patch']))
    starts = sw.tile_starts(padded, tuple(case['patch']), case['step'])
    assert len(starts) == case['n_tiles']
    for st, sl in zip(starts, case['first_slicers']):
        assert [int(s) for s in st] == [a for a, _ in sl]
        assert [int(s) + p for s, p in zip(st, case['patch'])] == [b for _, b in sl]


def test_model_folder_fixture_is_what_the_reference_reads(tmp_path):
    """The reference's initialize_from_trained_model_folder (real plans_handler.py / label_handling.py) read a folder
    written by model_folder.write_model_folder — the fixture the GPU predictor tests load — and predicted a two-fold
    ensemble from it; the oracle loop on the same two parameter sets gives the same logits, and the product's plans /
    label handling reads the same facts from the folder."""
    import os as _os
    from fast_nnunet_b200 import plans as P
    m = T['model_folder']
    spec = getattr(nets, T['net'])
    torch.set_num_threads(8)
    folder = str(tmp_path / 'nnUNetTrainer__nnUNetPlans__3d_fullres')
    sds = []
    for f, seed in zip(m['folds'], m['seeds']):
        sd = M.synthesize_state_dict(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], seed=seed, randomize_affine=True)
        M.write_model_folder(folder, spec['cls'], spec['kw'], spec['patch'], sd, spec['in_ch'], spec['heads'], fold=f,
                             mirror_axes=(0, 1, 2))
        sds.append(sd)
    # what the reference read from the folder == what the product's own readers see
    pm = P.PlansManager(P.load_json(_os.path.join(folder, 'plans.json')))
    cm = pm.get_configuration('3d_fullres')
    lm = pm.get_label_manager(P.load_json(_os.path.join(folder, 'dataset.json')))
    ck = torch.load(_os.path.join(folder, 'fold_0', 'checkpoint_final.pth'), map_location='cpu', weights_only=False)
    assert ck['trainer_name'] == m['trainer_name']
    assert list(ck['inference_allowed_mirroring_axes']) == m['allowed_mirroring_axes']
    assert list(cm.patch_size) == m['patch_size'] and cm.network_arch_class_name == m['network_arch_class_name']
    assert int(lm.num_segmentation_heads) == m['num_segmentation_heads']
    assert [int(i) for i in lm.all_labels] == m['all_labels']
    # the ensemble
    net = N.build_from_arch(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], allow_init=False)
    x = nets.ct_like_volume(tuple(m['volume'][1:]), m['volume'][0], seed=12)
    pred = None
    for sd in sds:
        net.load_state_dict(sd, strict=True)
        out = OP.predict_sliding_window_return_logits(net, x, tuple(m['patch_size']), 0.5, True, (0, 1, 2))
        pred = out.clone() if pred is None else pred + out
    pred = pred / len(sds)
    want = G['model_folder']
    got = pred.numpy()
    same = float((got == want).mean())
    assert got.shape == want.shape and same >= 0.999
    np.testing.assert_allclose(got.astype(np.float32), want.astype(np.float32), atol=4e-3, rtol=4e-3)
