"""GPU: every BASELINE.json configuration at its REAL patch size with TRAINED oracle weights (tests/nets.py) —
  cfg 1  distilled student r=2 on the full 1 x 160^3 volume (8 tiles x 8 mirror passes), end to end
  cfg 2  the same network on a 2 x 2 x 3-tile crop (192 x 192 x 256) of the 512 x 512 x 400 volume, end to end
  cfg 3  full-size teacher PlainConvUNet, forward of one patch
  cfg 4  ResEnc-M distilled student, 4-channel 155 x 240 x 240 volume in full (18 tiles), end to end
  cfg 5  bone_turbo-shaped anisotropic network with 61 heads: forward, and a 2 x 2 x 2-tile crop end to end
each against the fp32 oracle (network on the GPU with TF32 off) with north_star's label bar asserted on all voxels."""
import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import model_folder as M
from fast_nnunet_b200 import nnUNetPredictor
from fast_nnunet_b200.predictor import CompiledNetwork
from parity import DEV, compare, oracle_volume

pytestmark = pytest.mark.gpu

STEPS = {'STUDENT': 120, 'TEACHER': 80, 'RESENC_M_STUDENT': 150, 'BONE_TURBO': 150}


def _trained(name):
    return nets.train_oracle(getattr(nets, name), steps=STEPS[name], batch=1)


@pytest.mark.parametrize('name,batch', [('TEACHER', 1), ('RESENC_M_STUDENT', 2), ('BONE_TURBO', 2)])
def test_full_size_forward_matches_oracle(name, batch):
    spec = getattr(nets, name)
    sd, net = _trained(name)
    patch = spec['patch']
    vol, _ = nets.phantom_volume((patch[0], patch[1], patch[2] * batch), spec['in_ch'], spec['heads'], seed=21)
    x = torch.stack([vol[..., i * patch[2]:(i + 1) * patch[2]] for i in range(batch)])
    cn = CompiledNetwork(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], patch)
    cn.load_state_dict(sd)
    got = cn(x.to(DEV)).float()
    eng = cn.engine(DEV, batch)
    total, umma = eng.launch_counts()
    with torch.no_grad():
        want = net.to(DEV)(x.half().float().to(DEV))
    net.cpu()
    d = (got - want).abs()
    rng = want.abs().max().item()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    print(f'{name}: {umma}/{total} launches on tcgen05, GFLOP/forward={eng.program.total_flops() / 1e9:.1f}, '
          f'max|d|={d.max():.4f} mean|d|={d.mean():.5f} range={rng:.2f} argmax agreement={agree:.6f}')
    assert torch.isfinite(got).all()
    assert d.max().item() <= 0.08 * max(1.0, rng / 8) and d.mean().item() <= 0.008 * max(1.0, rng / 8)
    assert agree >= 0.999
    assert umma >= total // 2


def _plain_folder(tmp_path, spec, sd):
    return M.write_model_folder(str(tmp_path / 'nnUNetTrainer__nnUNetPlans__3d_fullres'), spec['cls'], spec['kw'],
                                spec['patch'], sd, spec['in_ch'], spec['heads'])


def _volume_parity(p, net, spec, vol_shape, seed):
    x, _ = nets.phantom_volume(vol_shape, spec['in_ch'], spec['heads'], seed=seed)
    got = p.predict_sliding_window_return_logits(x)
    assert tuple(got.shape) == (spec['heads'], *vol_shape)
    compare(got, oracle_volume(net, x.half().float(), spec['patch'], on_gpu=True), spec['heads'])
    labels = p.predict_sliding_window_return_segmentation(x)
    assert tuple(labels.shape) == tuple(vol_shape) and labels.dtype == torch.uint8
    assert float((labels == got.argmax(0)).float().mean()) >= 0.9999


def test_cfg1_student_full_volume(tmp_path):
    """BASELINE configs[0] in full: 1 x 160 x 160 x 160, 8 tiles x 8 mirror passes."""
    spec = nets.STUDENT
    sd, net = _trained('STUDENT')
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(_plain_folder(tmp_path, spec, sd), use_folds=(0,))
    assert len(p._internal_get_sliding_window_slicers((160, 160, 160))) == 8
    _volume_parity(p, net, spec, (160, 160, 160), seed=31)


def test_cfg2_student_crop(tmp_path):
    """A 2 x 2 x 3-tile crop of configs[1] (tile pitch 64 as in the 512 x 512 x 400 volume)."""
    spec = nets.STUDENT
    sd, net = _trained('STUDENT')
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(_plain_folder(tmp_path, spec, sd), use_folds=(0,))
    assert len(p._internal_get_sliding_window_slicers((192, 192, 256))) == 12
    _volume_parity(p, net, spec, (192, 192, 256), seed=32)


def test_cfg4_resenc_full_volume(tmp_path):
    """configs[3] in full (4 x 155 x 240 x 240, 18 tiles x 8 mirror passes) from a distillation-trainer folder."""
    spec = nets.RESENC_M_STUDENT
    sd, net = _trained('RESENC_M_STUDENT')
    folder = M.write_model_folder(str(tmp_path / 'nnUNetDistillationTrainer__nnUNetResEncUNetMPlans__3d_fullres'),
                                  spec['cls'], M.resenc_arch_kwargs([32, 64, 128, 256, 320, 320], [[3, 3, 3]] * 6,
                                                                    [[1, 1, 1]] + [[2, 2, 2]] * 5, [1, 3, 4, 6, 6, 6]),
                                  spec['patch'], sd, 4, 4, trainer_name='nnUNetDistillationTrainer',
                                  plans_name='nnUNetResEncUNetMPlans',
                                  init_args_extra={'feature_reduction_factor': 2, 'block_reduction_strategy': 'keep'})
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(folder, use_folds=(0,))
    assert p.network.network_class_name.endswith('ResidualEncoderUNet')
    assert p.network.arch_kwargs['features_per_stage'] == [16, 32, 64, 128, 160, 160]
    assert len(p._internal_get_sliding_window_slicers((155, 240, 240))) == 18
    _volume_parity(p, net, spec, (155, 240, 240), seed=33)


def test_cfg5_bone_turbo_crop(tmp_path):
    """A 2 x 2 x 2-tile crop of configs[4]: 61 heads, patch (160, 96, 96), anisotropic strides."""
    spec = nets.BONE_TURBO
    sd, net = _trained('BONE_TURBO')
    p = nnUNetPredictor(device=DEV, allow_tqdm=False)
    p.initialize_from_trained_model_folder(_plain_folder(tmp_path, spec, sd), use_folds=(0,))
    assert len(p._internal_get_sliding_window_slicers((240, 144, 144))) == 8
    _volume_parity(p, net, spec, (240, 144, 144), seed=34)
