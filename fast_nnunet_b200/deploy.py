"""Deployment formats of the reference's inference engines: the engine `.ini`
(/root/reference/engine/config/fast_nnunet_bone_turbo.ini:1-23) and the Python inferencers' JSON
(/root/reference/inference/config/3d_fullres/sample_config.json:1-19) next to an exported ONNX model
(distillation/fast_nnunet_distillation_export_onnx.py:432-472).  Released models such as bone_turbo ship as exactly
this pair; `predictor_from_deployment` turns it into a ready nnUNetPredictor on the B200 engine — the role the
withheld TensorRT runtime plays upstream.

    [model]         file_name, input_name, output_name, num_class
    [input]         depth, height, width, patch_size, target_spacing
    [preprocessing] mean, std_dev, lower_bound, upper_bound          (CT normalisation)
    [inference]     use_mirroring, step_size, use_gaussian
    JSON: patch_size, target_spacing, intensity_properties{mean, std, percentile_00_5, percentile_99_5}, model_path
"""
from __future__ import annotations

import configparser
import json
import os
from typing import Optional

import torch


def _floats(s: str):
    return [float(v) for v in s.replace(';', ',').split(',') if v.strip()]


def _bool(s: str) -> bool:
    return str(s).strip().lower() in ('1', 'true', 'yes', 'on')


def read_engine_ini(path: str) -> dict:
    """The TensorRT engine's config -> one flat description."""
    cp = configparser.ConfigParser()
    with open(path) as f:
        cp.read_file(f)
    m, i, p = cp['model'], cp['input'], cp['preprocessing']
    inf = cp['inference'] if cp.has_section('inference') else {}
    patch = [int(v) for v in _floats(i['patch_size'])] if 'patch_size' in i else \
        [int(i['depth']), int(i['height']), int(i['width'])]
    return {
        'model_file': m.get('file_name'), 'input_name': m.get('input_name', 'input'),
        'output_name': m.get('output_name', 'output'), 'num_class': int(m['num_class']),
        'patch_size': patch, 'target_spacing': _floats(i['target_spacing']),
        'intensity_properties': {'mean': float(p['mean']), 'std': float(p['std_dev']),
                                 'percentile_00_5': float(p['lower_bound']), 'percentile_99_5': float(p['upper_bound'])},
        'use_mirroring': _bool(inf.get('use_mirroring', 'false')), 'step_size': float(inf.get('step_size', 0.5)),
        'use_gaussian': _bool(inf.get('use_gaussian', 'true')),
    }


def read_inferencer_json(path: str) -> dict:
    """inference/config/<configuration>/*.json of the ONNX / TensorRT Python inferencers."""
    d = json.load(open(path))
    ip = d['intensity_properties']
    return {
        'model_file': d.get('model_path'), 'input_name': 'input', 'output_name': 'output',
        'num_class': d.get('num_class'), 'patch_size': [int(v) for v in d['patch_size']],
        'target_spacing': [float(v) for v in d['target_spacing']],
        'intensity_properties': {k: float(ip[k]) for k in ('mean', 'std', 'percentile_00_5', 'percentile_99_5')},
        'use_mirroring': bool(d.get('use_mirroring', False)), 'step_size': float(d.get('step_size', 0.5)),
        'use_gaussian': bool(d.get('use_gaussian', True)),
    }


def read_deployment_config(path: str) -> dict:
    return read_inferencer_json(path) if path.lower().endswith('.json') else read_engine_ini(path)


def plans_from_deployment(cfg: dict, arch_kwargs: dict, in_channels: int) -> (dict, dict):
    """plans.json / dataset.json as a trained-model folder would carry them, from the deployment config."""
    from .model_folder import PLAIN
    n_cls = int(cfg['num_class'])
    plans = {
        'dataset_name': 'Deployment', 'plans_name': 'nnUNetPlans',
        'original_median_spacing_after_transp': list(cfg['target_spacing']),
        'image_reader_writer': 'SimpleITKIO', 'transpose_forward': [0, 1, 2], 'transpose_backward': [0, 1, 2],
        'experiment_planner_used': 'deployment', 'label_manager': 'LabelManager',
        'foreground_intensity_properties_per_channel': {str(c): dict(cfg['intensity_properties']) for c in range(in_channels)},
        'configurations': {'3d_fullres': {
            'data_identifier': 'nnUNetPlans_3d_fullres', 'preprocessor_name': 'DefaultPreprocessor', 'batch_size': 2,
            'patch_size': list(cfg['patch_size']), 'spacing': list(cfg['target_spacing']),
            'normalization_schemes': ['CTNormalization'] * in_channels, 'use_mask_for_norm': [False] * in_channels,
            'resampling_fn_data': 'resample_data_or_seg_to_shape', 'resampling_fn_seg': 'resample_data_or_seg_to_shape',
            'resampling_fn_data_kwargs': {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_seg_kwargs': {'is_seg': True, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_probabilities': 'resample_data_or_seg_to_shape',
            'resampling_fn_probabilities_kwargs': {'is_seg': False, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'architecture': {'network_class_name': PLAIN, 'arch_kwargs': arch_kwargs,
                             '_kw_requires_import': ['conv_op', 'norm_op', 'dropout_op', 'nonlin']},
            'batch_dice': False}},
    }
    dataset = {'channel_names': {str(c): 'CT' for c in range(in_channels)},
               'labels': {'background': 0, **{f'class_{k}': k for k in range(1, n_cls)}},
               'numTraining': 0, 'file_ending': '.nii.gz'}
    return plans, dataset


def predictor_from_deployment(config_path: str, onnx_path: Optional[str] = None, device=torch.device('cuda'), **kwargs):
    """Engine `.ini` (or inferencer JSON) + exported ONNX -> nnUNetPredictor (mirroring / step size / Gaussian as the
    config says; the eight-fold mirror TTA can be switched on with use_mirroring=True)."""
    from .onnx_import import plain_conv_unet_from_onnx
    from .plans import PlansManager
    from .predictor import CompiledNetwork, nnUNetPredictor
    cfg = read_deployment_config(config_path)
    if onnx_path is None:
        onnx_path = cfg['model_file']
        if onnx_path and not os.path.isabs(onnx_path):
            onnx_path = os.path.join(os.path.dirname(os.path.abspath(config_path)), onnx_path)
        if onnx_path and not onnx_path.lower().endswith('.onnx'):       # the .ini names the TensorRT plan: same stem
            onnx_path = os.path.splitext(onnx_path)[0] + '.onnx'
    kw, sd_np, info = plain_conv_unet_from_onnx(onnx_path)
    if cfg.get('num_class') is None:
        cfg['num_class'] = info['num_heads']
    if int(cfg['num_class']) != info['num_heads']:
        raise ValueError(f"config says num_class = {cfg['num_class']} but the ONNX head has {info['num_heads']} outputs")
    if (cfg['input_name'], cfg['output_name']) != (info['input_name'], info['output_name']):
        raise ValueError(f"config names tensors {cfg['input_name']!r} / {cfg['output_name']!r}, the ONNX graph "
                         f"{info['input_name']!r} / {info['output_name']!r}")
    plans, dataset = plans_from_deployment(cfg, kw, info['input_channels'])
    pm = PlansManager(plans)
    cm = pm.get_configuration('3d_fullres')
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    net = CompiledNetwork(cm.network_arch_class_name, cm.network_arch_init_kwargs, info['input_channels'],
                          info['num_heads'], cm.patch_size)
    net.load_state_dict(sd)
    pred = nnUNetPredictor(tile_step_size=kwargs.pop('tile_step_size', cfg['step_size']),
                           use_gaussian=kwargs.pop('use_gaussian', cfg['use_gaussian']),
                           use_mirroring=kwargs.pop('use_mirroring', cfg['use_mirroring']), device=device,
                           allow_tqdm=False, **kwargs)
    pred.manual_initialization(net, pm, cm, [sd], dataset, 'nnUNetTrainer', (0, 1, 2))
    return pred
