"""GPU end-to-end parity: model folder -> nnUNetPredictor (libfnnu) vs the oracle's restatement of the
reference loop (fp32 network, fp16 accumulators), same weights, same volume."""
import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import model_folder as M
from fast_nnunet_b200 import nnUNetPredictor
from oracle import predictor as OP

pytestmark = pytest.mark.gpu


def _folder(tmp_path, spec, sd, **kw):
    return M.write_model_folder(str(tmp_path / 'nnUNetTrainer__nnUNetPlans__3d_fullres'), spec['cls'], spec['kw'],
                                spec['patch'], sd, spec['in_ch'], spec['heads'], **kw)


from parity import DEV, compare as _compare, oracle_volume as _oracle  # noqa: E402


TRAIN_STEPS = {'SMALL_PLAIN16': 200, 'SMALL_PLAIN': 200, 'SMALL_RESENC': 300, 'ANISO_PLAIN': 300}


def _fixture(name):
    spec = getattr(nets, name)
    sd, net = nets.train_oracle(spec, steps=TRAIN_STEPS[name])
    return spec, sd, net


@pytest.mark.parametrize('name,vol', [('SMALL_PLAIN16', (48, 40, 56)), ('SMALL_RESENC', (40, 40, 40)),
                                      ('ANISO_PLAIN', (30, 40, 50))])
def test_sliding_window_matches_oracle(tmp_path, name, vol):
    spec, sd, net = _fixture(name)
    folder = _folder(tmp_path, spec, sd)
    x, _ = nets.phantom_volume(vol, spec['in_ch'], spec['heads'], seed=3)
    p = nnUNetPredictor(tile_step_size=0.5, use_gaussian=True, use_mirroring=True, device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    got = p.predict_sliding_window_return_logits(x)
    assert got.dtype == torch.float16 and tuple(got.shape) == (spec['heads'], *vol) and got.device.type == 'cuda'
    _compare(got, _oracle(net, x.half().float(), spec['patch']), spec['heads'])
    # fused argmax == argmax of the logits of the SAME run (bit-exact); a second run may differ in a few near-ties
    # because the fp64 InstanceNorm sums are accumulated with atomics in arbitrary order
    data, revert = p._pad(x)
    lg, lb = p._internal_predict_sliding_window_return_logits(data, p._internal_get_sliding_window_slicers(data.shape[1:]),
                                                              True, return_labels=True)
    assert np.array_equal(lb.cpu().numpy(), OP.logits_to_segmentation(lg).astype(np.uint8))
    labels = p.predict_sliding_window_return_segmentation(x)
    assert float((labels.cpu().numpy() == OP.logits_to_segmentation(got)).mean()) >= 0.9999
    assert p.last_launches > 0


def test_small_volume_is_padded(tmp_path):
    spec, sd, net = _fixture('SMALL_PLAIN16')
    folder = _folder(tmp_path, spec, sd)
    x, _ = nets.phantom_volume((20, 33, 30), 1, 2, seed=4)
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=None)
    got = p.predict_sliding_window_return_logits(x)
    assert tuple(got.shape) == (2, 20, 33, 30)
    _compare(got, _oracle(net, x.half().float(), spec['patch']), 2)


def test_no_mirroring_no_gaussian_and_fp16_accumulators(tmp_path):
    spec, sd, net = _fixture('SMALL_PLAIN16')
    folder = _folder(tmp_path, spec, sd, mirror_axes=None)
    x, _ = nets.phantom_volume((40, 40, 48), 1, 2, seed=5)
    p = nnUNetPredictor(use_gaussian=False, use_mirroring=True, device=DEV, allow_tqdm=False,
                        accumulator_dtype=torch.float16, tiles_per_batch=3)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    assert p.allowed_mirroring_axes is None
    got = p.predict_sliding_window_return_logits(x)
    _compare(got, _oracle(net, x.half().float(), spec['patch'], False, None), 2)


def test_fold_ensemble_and_cpu_return(tmp_path):
    spec = nets.SMALL_PLAIN
    sd0, net0 = nets.train_oracle(spec, steps=200, seed=1)
    sd1, net1 = nets.train_oracle(spec, steps=200, seed=2)
    folder = _folder(tmp_path, spec, sd0, fold=0)
    M.write_model_folder(folder, spec['cls'], spec['kw'], spec['patch'], sd1, 1, 2, fold=1)
    x, _ = nets.phantom_volume((32, 40, 32), 1, 2, seed=6)
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=None)
    assert len(p.list_of_parameters) == 2
    got = p.predict_logits_from_preprocessed_data(x)
    assert got.device.type == 'cpu'
    o0 = _oracle(net0, x.half().float(), spec['patch'])
    o1 = _oracle(net1, x.half().float(), spec['patch'])
    both = ((o0[0].float() + o1[0].float()) / 2, torch.minimum(o0[1], o1[1]), (o0[2] + o1[2]) / 2)
    _compare(got, both, 2)


def test_manual_initialization_with_live_module(tmp_path):
    from fast_nnunet_b200.plans import PlansManager, load_json
    import os
    spec, sd, net = _fixture('SMALL_PLAIN16')
    folder = _folder(tmp_path, spec, sd)
    pm = PlansManager(load_json(os.path.join(folder, 'plans.json')))
    dj = load_json(os.path.join(folder, 'dataset.json'))
    p = nnUNetPredictor(device=DEV, allow_tqdm=False, use_mirroring=False)
    p.manual_initialization(net, pm, pm.get_configuration('3d_fullres'), None, dj, 'nnUNetTrainer', (0, 1, 2))
    x, _ = nets.phantom_volume((32, 32, 48), 1, 2, seed=8)
    got = p.predict_sliding_window_return_logits(x)
    _compare(got, _oracle(net, x.half().float(), spec['patch'], True, None), 2)


@pytest.mark.parametrize('spacing', [[1.0, 1.0, 1.0], [1.6, 0.8, 0.9]])
def test_predict_single_npy_array(tmp_path, spacing):
    """Raw array in, label map in the image's own geometry out (predict_from_raw_data.py:423-468): device
    pre-processing (f2) -> sliding window -> device export (f1), against the oracle's chain
    run_case_npy -> predict_sliding_window_return_logits -> convert_predicted_logits_to_segmentation_with_correct_shape."""
    from oracle import export as OX
    from oracle import preprocess as OPP
    spec, sd, net = _fixture('SMALL_PLAIN16')
    folder = _folder(tmp_path, spec, sd, normalization='CTNormalization')
    from fast_nnunet_b200.plans import PlansManager, load_json
    import os
    pm = PlansManager(load_json(os.path.join(folder, 'plans.json')))
    cm = pm.get_configuration('3d_fullres')
    ip = pm.foreground_intensity_properties_per_channel['0']
    # a phantom in HU-like units such that CT normalisation maps it to the range the fixture was trained on
    x, _ = nets.phantom_volume((44, 40, 48), 1, 2, seed=9)
    img = (x.numpy() * ip['std'] + ip['mean']).astype(np.float32)
    img[:, :5] = 0
    img[:, :, -4:] = 0
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    seg = p.predict_single_npy_array(img, {'spacing': spacing})
    assert seg.shape == img.shape[1:] and seg.dtype == np.uint8
    data, props = OPP.run_case_npy(img.copy(), {'spacing': spacing}, list(pm.transpose_forward), cm.spacing,
                                   cm.normalization_schemes, cm.use_mask_for_norm,
                                   pm.foreground_intensity_properties_per_channel)
    logits = OP.predict_sliding_window_return_logits(net.to(DEV), torch.from_numpy(data).to(DEV), spec['patch'], 0.5, True,
                                                     (0, 1, 2)).cpu().numpy()
    net.cpu()
    want = OX.convert_predicted_logits_to_segmentation_with_correct_shape(logits, cm.spacing, list(pm.transpose_forward),
                                                                          list(pm.transpose_backward), props)
    agree = float((seg == want).mean())
    dice = [d for d in OP.dice_per_class(seg, want, 2) if d == d]
    print(f'predict_single_npy_array spacing {spacing}: labels agree={agree:.6f} dice={dice}')
    assert agree >= 0.999 and min(dice) >= 0.999
