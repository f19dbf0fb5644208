"""Model-folder layout the predictor consumes, plus writers for synthetic folders (tests / bench).

Layout (predict_from_raw_data.py:67-129, nnUNetTrainer.save_checkpoint nnUNetTrainer.py:1149-1172):
    <dir>/dataset.json, <dir>/plans.json, <dir>/fold_{k}/<checkpoint>.pth with keys
    network_weights, trainer_name, init_args{configuration, ...}, inference_allowed_mirroring_axes.
Distilled students (which the reference's predictor cannot load) use the same layout
with trainer_name nnUNetDistillationTrainer[DA5] and init_args carrying feature_reduction_factor /
block_reduction_strategy (nnUNetDistillationTrainer.py:441-451); the architecture is rebuilt with the
rule of :678 (features) and :688-708 (blocks) and cross-checked against the checkpoint's shapes.
"""
from __future__ import annotations

import json
import math
import os
from copy import deepcopy
from typing import Dict, Optional, Sequence

import torch

PLAIN = 'dynamic_network_architectures.architectures.unet.PlainConvUNet'
RESENC = 'dynamic_network_architectures.architectures.unet.ResidualEncoderUNet'
STUDENT_TRAINERS = ('nnUNetDistillationTrainer', 'nnUNetDistillationTrainerDA5')


def plain_arch_kwargs(features, kernel_sizes, strides, n_conv_per_stage=2, n_conv_per_stage_decoder=2) -> dict:
    n = len(features)
    return {
        'n_stages': n, 'features_per_stage': list(features), 'conv_op': 'torch.nn.modules.conv.Conv3d',
        'kernel_sizes': [list(k) for k in kernel_sizes], 'strides': [list(s) for s in strides],
        'n_conv_per_stage': [n_conv_per_stage] * n if isinstance(n_conv_per_stage, int) else list(n_conv_per_stage),
        'n_conv_per_stage_decoder': [n_conv_per_stage_decoder] * (n - 1)
        if isinstance(n_conv_per_stage_decoder, int) else list(n_conv_per_stage_decoder),
        'conv_bias': True, 'norm_op': 'torch.nn.modules.instancenorm.InstanceNorm3d',
        'norm_op_kwargs': {'eps': 1e-05, 'affine': True}, 'dropout_op': None, 'dropout_op_kwargs': None,
        'nonlin': 'torch.nn.LeakyReLU', 'nonlin_kwargs': {'inplace': True},
    }


def resenc_arch_kwargs(features, kernel_sizes, strides, n_blocks_per_stage, n_conv_per_stage_decoder=1) -> dict:
    n = len(features)
    kw = plain_arch_kwargs(features, kernel_sizes, strides, 1, n_conv_per_stage_decoder)
    del kw['n_conv_per_stage']
    kw['n_blocks_per_stage'] = list(n_blocks_per_stage)
    assert len(kw['n_blocks_per_stage']) == n
    return kw


def student_features(features, reduction_factor):
    """nnUNetDistillationTrainer.py:678."""
    return [max(int(f) // int(reduction_factor), 8) for f in features]


def student_blocks(n_blocks, features, lite_features, strategy):
    """nnUNetDistillationTrainer.py:688-708."""
    if strategy == 'reduce':
        return [max(n // 2, 1) for n in n_blocks]
    if strategy == 'increase':
        return [min(n + 1, 8) for n in n_blocks]
    if strategy == 'adaptive':
        return [min(n + max(0, int((o / r) / 4)), 8) for n, o, r in zip(n_blocks, features, lite_features)]
    return list(n_blocks)


def effective_arch(network_class_name: str, arch_kwargs: dict, trainer_name: str, init_args: dict,
                   plans_name: str = '', state_dict: Optional[dict] = None) -> (str, dict):
    """Architecture actually stored in the checkpoint: the plans' arch_kwargs for a teacher; the reduced
    one for a distilled student.  Plain vs residual encoder is read off the checkpoint's own keys when a
    state_dict is given (a LiteNNUNetStudent distilled from a ResEnc teacher keeps the teacher's plans)."""
    kw = deepcopy(arch_kwargs)
    cls = network_class_name
    if trainer_name in STUDENT_TRAINERS:
        r = int(init_args.get('feature_reduction_factor', 2))
        feats = [int(f) for f in kw['features_per_stage']]
        lite = student_features(feats, r)
        kw['features_per_stage'] = lite
        if state_dict is not None:
            from .program import clean_state_dict
            is_resenc = 'encoder.stem.convs.0.conv.weight' in clean_state_dict(state_dict)
        else:
            # the reference's rule (nnUNetDistillationTrainer.py:615)
            is_resenc = 'ResEnc' in (init_args.get('student_plans_identifier') or plans_name or '')
        if is_resenc:
            nb = kw.get('n_blocks_per_stage') or [1, 3, 4, 6, 6, 6][:kw['n_stages']]
            kw['n_blocks_per_stage'] = student_blocks(list(nb), feats, lite,
                                                      init_args.get('block_reduction_strategy', 'keep'))
            kw.pop('n_conv_per_stage', None)
            cls = RESENC
        else:
            cls = PLAIN
    return cls, kw


def _kaiming(shape, fan_in, gen, a=1e-2):
    std = math.sqrt(2.0 / (1 + a * a)) / math.sqrt(fan_in)
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def synthesize_state_dict(network_class_name: str, arch_kwargs: dict, in_channels: int, num_heads: int,
                          seed: int = 1234, randomize_affine: bool = False, deep_supervision_keys: bool = True,
                          prefix: str = '') -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's initialisation (utilities/network_initialization.py:4-12:
    Kaiming-normal a=0.01, zero bias; InstanceNorm gamma=1, beta=0) under the exact key layout real
    checkpoints carry, duplicates (`all_modules.*`, `decoder.encoder.*`) included.  `randomize_affine`
    draws bias/gamma/beta at random instead so that tests exercise them."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    n = int(arch_kwargs['n_stages'])
    feats = [int(f) for f in arch_kwargs['features_per_stage']]
    ks = arch_kwargs['kernel_sizes']
    kernels = [tuple(ks for _ in range(3))] * n if isinstance(ks, int) else [tuple(int(i) for i in k) for k in ks]
    strides = [tuple(int(i) for i in s) for s in arch_kwargs['strides']]
    bias = bool(arch_kwargs.get('conv_bias', False))
    resenc = network_class_name.endswith('ResidualEncoderUNet')

    def vec(c, base):
        if randomize_affine:
            return base + 0.25 * torch.randn(c, generator=gen)
        return torch.full((c,), float(base))

    def conv_norm(pre, cin, cout, k, use_bias, nonlin=True):
        w = _kaiming((cout, cin, *k), cin * k[0] * k[1] * k[2], gen)
        sd[pre + '.conv.weight'] = w
        sd[pre + '.all_modules.0.weight'] = w
        if use_bias:
            bvec = vec(cout, 0.0)
            sd[pre + '.conv.bias'] = bvec
            sd[pre + '.all_modules.0.bias'] = bvec
        g, be = vec(cout, 1.0), vec(cout, 0.0)
        sd[pre + '.norm.weight'] = g
        sd[pre + '.norm.bias'] = be
        sd[pre + '.all_modules.1.weight'] = g
        sd[pre + '.all_modules.1.bias'] = be

    if resenc:
        nb = arch_kwargs['n_blocks_per_stage']
        nb = [nb] * n if isinstance(nb, int) else list(nb)
        conv_norm('encoder.stem.convs.0', in_channels, feats[0], kernels[0], bias)
        cin = feats[0]
        for s in range(n):
            for blk in range(nb[s]):
                pre = f'encoder.stages.{s}.blocks.{blk}'
                st = strides[s] if blk == 0 else (1, 1, 1)
                conv_norm(pre + '.conv1', cin, feats[s], kernels[s], bias)
                conv_norm(pre + '.conv2', feats[s], feats[s], kernels[s], bias, nonlin=False)
                has_stride = any(i != 1 for i in st)
                if cin != feats[s]:
                    conv_norm(pre + f'.skip.{1 if has_stride else 0}', cin, feats[s], (1, 1, 1), False, nonlin=False)
                cin = feats[s]
    else:
        nc = arch_kwargs['n_conv_per_stage']
        nc = [nc] * n if isinstance(nc, int) else list(nc)
        cin = in_channels
        for s in range(n):
            for j in range(nc[s]):
                conv_norm(f'encoder.stages.{s}.0.convs.{j}', cin, feats[s], kernels[s], bias)
                cin = feats[s]
    nd = arch_kwargs['n_conv_per_stage_decoder']
    nd = [nd] * (n - 1) if isinstance(nd, int) else list(nd)
    for l in range(n - 1):
        s = n - 2 - l
        below, c, st = feats[s + 1], feats[s], strides[s + 1]
        # torch's fan_in for ConvTranspose3d weight (cin, cout, *k) is size(1) * prod(k)
        sd[f'decoder.transpconvs.{l}.weight'] = _kaiming((below, c, *st), c * st[0] * st[1] * st[2], gen)
        if bias:
            sd[f'decoder.transpconvs.{l}.bias'] = vec(c, 0.0)
        cin = 2 * c
        for j in range(nd[l]):
            conv_norm(f'decoder.stages.{l}.convs.{j}', cin, c, kernels[s], bias)
            cin = c
        if deep_supervision_keys or l == n - 2:
            sd[f'decoder.seg_layers.{l}.weight'] = _kaiming((num_heads, c, 1, 1, 1), c, gen)
            sd[f'decoder.seg_layers.{l}.bias'] = vec(num_heads, 0.0)
    for k in [k for k in sd if k.startswith('encoder.')]:
        sd['decoder.' + k] = sd[k]
    if prefix:
        sd = {prefix + k: v for k, v in sd.items()}
    return sd


def write_model_folder(folder: str, network_class_name: str, arch_kwargs: dict, patch_size: Sequence[int],
                       state_dict: Dict[str, torch.Tensor], in_channels: int, num_heads: int,
                       trainer_name: str = 'nnUNetTrainer', configuration: str = '3d_fullres',
                       plans_name: str = 'nnUNetPlans', init_args_extra: Optional[dict] = None,
                       mirror_axes=(0, 1, 2), fold=0, checkpoint_name: str = 'checkpoint_final.pth',
                       spacing=(1.0, 1.0, 1.0), normalization: str = 'ZScoreNormalization',
                       regions: bool = False) -> str:
    """Writes dataset.json, plans.json and fold_<fold>/<checkpoint_name>.  `arch_kwargs` are the PLANS'
    kwargs (teacher-sized for a student); `state_dict` holds the checkpoint's actual weights."""
    os.makedirs(os.path.join(folder, f'fold_{fold}'), exist_ok=True)
    labels = {'background': 0}
    for i in range(1, num_heads):
        labels[f'class_{i}'] = i
    dataset_json = {'channel_names': {str(i): f'ch{i}' for i in range(in_channels)}, 'labels': labels,
                    'numTraining': 0, 'file_ending': '.nii.gz'}
    plans = {
        'dataset_name': 'Dataset999_Synthetic', 'plans_name': plans_name,
        'original_median_spacing_after_transp': list(spacing),
        'original_median_shape_after_transp': list(patch_size),
        'image_reader_writer': 'SimpleITKIO', 'transpose_forward': [0, 1, 2], 'transpose_backward': [0, 1, 2],
        'experiment_planner_used': 'ExperimentPlanner', 'label_manager': 'LabelManager',
        'foreground_intensity_properties_per_channel': {
            str(i): {'max': 1207.0, 'mean': -350.0, 'median': -350.0, 'min': -1100.0,
                     'percentile_00_5': -1024.0, 'percentile_99_5': 1000.0, 'std': 450.0}
            for i in range(in_channels)},
        'configurations': {
            configuration: {
                'data_identifier': f'{plans_name}_{configuration}', 'preprocessor_name': 'DefaultPreprocessor',
                'batch_size': 2, 'patch_size': list(patch_size), 'median_image_size_in_voxels': list(patch_size),
                'spacing': list(spacing), 'normalization_schemes': [normalization] * in_channels,
                'use_mask_for_norm': [False] * in_channels,
                'resampling_fn_data': 'resample_data_or_seg_to_shape',
                'resampling_fn_seg': 'resample_data_or_seg_to_shape',
                'resampling_fn_data_kwargs': {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None},
                'resampling_fn_seg_kwargs': {'is_seg': True, 'order': 1, 'order_z': 0, 'force_separate_z': None},
                'resampling_fn_probabilities': 'resample_data_or_seg_to_shape',
                'resampling_fn_probabilities_kwargs': {'is_seg': False, 'order': 1, 'order_z': 0,
                                                       'force_separate_z': None},
                'architecture': {'network_class_name': network_class_name, 'arch_kwargs': arch_kwargs,
                                 '_kw_requires_import': ['conv_op', 'norm_op', 'dropout_op', 'nonlin']},
                'batch_dice': False,
            }
        },
    }
    with open(os.path.join(folder, 'dataset.json'), 'w') as f:
        json.dump(dataset_json, f, indent=1)
    with open(os.path.join(folder, 'plans.json'), 'w') as f:
        json.dump(plans, f, indent=1)
    init_args = {'plans': plans, 'configuration': configuration, 'fold': fold, 'dataset_json': dataset_json}
    init_args.update(init_args_extra or {})
    ckpt = {'network_weights': state_dict, 'optimizer_state': None, 'grad_scaler_state': None, 'logging': None,
            '_best_ema': None, 'current_epoch': 1000, 'init_args': init_args, 'trainer_name': trainer_name,
            'inference_allowed_mirroring_axes': tuple(mirror_axes) if mirror_axes is not None else None}
    torch.save(ckpt, os.path.join(folder, f'fold_{fold}', checkpoint_name))
    return folder
