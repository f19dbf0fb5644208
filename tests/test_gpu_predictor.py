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
DEV = torch.device('cuda', 0)


def _folder(tmp_path, spec, sd, **kw):
    return M.write_model_folder(str(tmp_path / 'nnUNetTrainer__nnUNetPlans__3d_fullres'), spec['cls'], spec['kw'],
                                spec['patch'], sd, spec['in_ch'], spec['heads'], **kw)


def _compare(got_logits, want_logits, heads, tol_max, tol_mean):
    got = got_logits.float().cpu()
    want = want_logits.float().cpu()
    d = (got - want).abs()
    seg_g = OP.logits_to_segmentation(got)
    seg_w = OP.logits_to_segmentation(want)
    agree = float((seg_g == seg_w).mean())
    top2 = torch.topk(want, 2, dim=0).values
    margin = (top2[0] - top2[1]).numpy()
    confident = margin > 2 * tol_max
    agree_conf = float((seg_g == seg_w)[confident].mean()) if confident.any() else 1.0
    dice = OP.dice_per_class(seg_g, seg_w, heads)
    print(f'max|d|={d.max():.4f} mean|d|={d.mean():.5f} agree={agree:.5f} agree(margin>{2 * tol_max})={agree_conf:.6f} '
          f'({confident.mean():.3f} of voxels) dice={["%.4f" % x for x in dice]}')
    assert d.max().item() <= tol_max and d.mean().item() <= tol_mean
    assert agree_conf >= 0.999
    return agree, dice


@pytest.mark.parametrize('name,vol', [('SMALL_PLAIN16', (48, 40, 56)), ('SMALL_RESENC', (40, 40, 40)),
                                      ('ANISO_PLAIN', (30, 40, 50))])
def test_sliding_window_matches_oracle(tmp_path, name, vol):
    spec = getattr(nets, name)
    sd, net = nets.make(spec)
    folder = _folder(tmp_path, spec, sd)
    x = nets.ct_like_volume(vol, spec['in_ch'])
    p = nnUNetPredictor(tile_step_size=0.5, use_gaussian=True, use_mirroring=True, device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    got = p.predict_sliding_window_return_logits(x)
    assert got.dtype == torch.float16 and tuple(got.shape) == (spec['heads'], *vol) and got.device.type == 'cuda'
    want = OP.predict_sliding_window_return_logits(net, x.half().float(), spec['patch'], 0.5, True, (0, 1, 2))
    _compare(got, want, spec['heads'], 0.1, 0.012)
    labels = p.predict_sliding_window_return_segmentation(x)
    assert np.array_equal(labels.cpu().numpy(), OP.logits_to_segmentation(got).astype(np.uint8))
    assert p.last_launches > 0


def test_small_volume_is_padded(tmp_path):
    spec = nets.SMALL_PLAIN16
    sd, net = nets.make(spec)
    folder = _folder(tmp_path, spec, sd)
    x = nets.ct_like_volume((20, 33, 30), 1)
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=None)
    got = p.predict_sliding_window_return_logits(x)
    assert tuple(got.shape) == (2, 20, 33, 30)
    want = OP.predict_sliding_window_return_logits(net, x.half().float(), spec['patch'], 0.5, True, (0, 1, 2))
    _compare(got, want, 2, 0.1, 0.012)


def test_no_mirroring_no_gaussian_and_fp16_accumulators(tmp_path):
    spec = nets.SMALL_PLAIN16
    sd, net = nets.make(spec)
    folder = _folder(tmp_path, spec, sd, mirror_axes=None)
    x = nets.ct_like_volume((40, 40, 48), 1)
    p = nnUNetPredictor(use_gaussian=False, use_mirroring=True, device=DEV, allow_tqdm=False,
                        accumulator_dtype=torch.float16, tiles_per_batch=3)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    assert p.allowed_mirroring_axes is None
    got = p.predict_sliding_window_return_logits(x)
    want = OP.predict_sliding_window_return_logits(net, x.half().float(), spec['patch'], 0.5, False, None)
    _compare(got, want, 2, 0.1, 0.012)


def test_fold_ensemble_and_cpu_return(tmp_path):
    spec = nets.SMALL_PLAIN
    sd0, net0 = nets.make(spec, seed=1)
    sd1, net1 = nets.make(spec, seed=2)
    folder = _folder(tmp_path, spec, sd0, fold=0)
    M.write_model_folder(folder, spec['cls'], spec['kw'], spec['patch'], sd1, 1, 2, fold=1)
    x = nets.ct_like_volume((32, 40, 32), 1)
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=None)
    assert len(p.list_of_parameters) == 2
    got = p.predict_logits_from_preprocessed_data(x)
    assert got.device.type == 'cpu'
    w0 = OP.predict_sliding_window_return_logits(net0, x.half().float(), spec['patch'], 0.5, True, (0, 1, 2))
    w1 = OP.predict_sliding_window_return_logits(net1, x.half().float(), spec['patch'], 0.5, True, (0, 1, 2))
    want = (w0 + w1) / 2
    _compare(got, want, 2, 0.1, 0.012)


def test_manual_initialization_with_live_module(tmp_path):
    from fast_nnunet_b200.plans import PlansManager, load_json
    import os
    spec = nets.SMALL_PLAIN16
    sd, net = nets.make(spec)
    folder = _folder(tmp_path, spec, sd)
    pm = PlansManager(load_json(os.path.join(folder, 'plans.json')))
    dj = load_json(os.path.join(folder, 'dataset.json'))
    p = nnUNetPredictor(device=DEV, allow_tqdm=False, use_mirroring=False)
    p.manual_initialization(net, pm, pm.get_configuration('3d_fullres'), None, dj, 'nnUNetTrainer', (0, 1, 2))
    x = nets.ct_like_volume((32, 32, 48), 1)
    got = p.predict_sliding_window_return_logits(x)
    want = OP.predict_sliding_window_return_logits(net, x.half().float(), spec['patch'], 0.5, True, None)
    _compare(got, want, 2, 0.1, 0.012)


def test_predict_single_npy_array(tmp_path):
    spec = nets.SMALL_PLAIN16
    sd, net = nets.make(spec)
    folder = _folder(tmp_path, spec, sd, normalization='CTNormalization')
    g = np.random.default_rng(0)
    img = g.normal(-350, 450, size=(1, 40, 36, 44)).astype(np.float32)
    img[:, :4] = 0
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    seg = p.predict_single_npy_array(img, {'spacing': [1.0, 1.0, 1.0]})
    assert seg.shape == (40, 36, 44) and seg.dtype == np.uint8
    with pytest.raises(NotImplementedError):
        p.predict_single_npy_array(img, {'spacing': [2.0, 1.0, 1.0]})
