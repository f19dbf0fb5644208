"""Golden table for plans handling, produced by EXECUTING the reference's utilities/plans_handling/plans_handler.py.

Run in the build container only (needs /root/reference):
    python tests/golden/make_plans_golden.py

Stubs (modules absent from the image): dynamic_network_architectures' convert_dim_to_conv_op / get_matching_instancenorm
(restated: dimension -> torch.nn.ConvNd / InstanceNormNd), the resampling-function and reader/writer registries
(look-ups by name, not reached except for the resampling function), batchgenerators' load_json / join.
Pins fast_nnunet_b200/plans.py: configuration inheritance, every accessor the predictor reads, and the upgrade of
old-format plans (UNet_class_name ...) to the 'architecture' entry.

Writes tests/golden/plans_golden.json (inputs and the reference's answers).
"""
import importlib.util
import json
import os
import sys
import types
from copy import deepcopy

import torch

REF = '/root/reference/distillation/nnunetv2'
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


NEW_PLANS = {
    'dataset_name': 'Dataset501_Golden', 'plans_name': 'nnUNetPlans',
    'original_median_spacing_after_transp': [2.5, 0.8, 0.8], 'original_median_shape_after_transp': [120, 512, 512],
    'image_reader_writer': 'SimpleITKIO', 'transpose_forward': [2, 0, 1], 'transpose_backward': [1, 2, 0],
    'experiment_planner_used': 'ExperimentPlanner', 'label_manager': 'LabelManager',
    'foreground_intensity_properties_per_channel': {'0': {'max': 3071.0, 'mean': 99.4, 'median': 101.0, 'min': -1024.0,
                                                          'percentile_00_5': -41.0, 'percentile_99_5': 207.0, 'std': 39.4}},
    'configurations': {
        '3d_fullres': {
            'data_identifier': 'nnUNetPlans_3d_fullres', 'preprocessor_name': 'DefaultPreprocessor', 'batch_size': 2,
            'patch_size': [64, 192, 160], 'median_image_size_in_voxels': [120.0, 512.0, 512.0], 'spacing': [2.5, 0.8, 0.8],
            'normalization_schemes': ['CTNormalization'], 'use_mask_for_norm': [False],
            'resampling_fn_data': 'resample_data_or_seg_to_shape', 'resampling_fn_seg': 'resample_data_or_seg_to_shape',
            'resampling_fn_data_kwargs': {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_seg_kwargs': {'is_seg': True, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_probabilities': 'resample_data_or_seg_to_shape',
            'resampling_fn_probabilities_kwargs': {'is_seg': False, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'architecture': {
                'network_class_name': 'dynamic_network_architectures.architectures.unet.PlainConvUNet',
                'arch_kwargs': {'n_stages': 6, 'features_per_stage': [32, 64, 128, 256, 320, 320], 'conv_op': 'torch.nn.modules.conv.Conv3d',
                                'kernel_sizes': [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                                'strides': [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]],
                                'n_conv_per_stage': [2, 2, 2, 2, 2, 2], 'n_conv_per_stage_decoder': [2, 2, 2, 2, 2], 'conv_bias': True,
                                'norm_op': 'torch.nn.modules.instancenorm.InstanceNorm3d', 'norm_op_kwargs': {'eps': 1e-05, 'affine': True},
                                'dropout_op': None, 'dropout_op_kwargs': None, 'nonlin': 'torch.nn.LeakyReLU', 'nonlin_kwargs': {'inplace': True}},
                '_kw_requires_import': ['conv_op', 'norm_op', 'dropout_op', 'nonlin']},
            'batch_dice': False},
        '3d_lowres': {'inherits_from': '3d_fullres', 'data_identifier': 'nnUNetPlans_3d_lowres', 'spacing': [3.1, 1.6, 1.6],
                      'patch_size': [80, 160, 160], 'batch_dice': True, 'next_stage': '3d_cascade_fullres'},
        '3d_cascade_fullres': {'inherits_from': '3d_fullres', 'previous_stage': '3d_lowres'},
    },
}

OLD_PLANS = {
    'dataset_name': 'Dataset502_OldFormat', 'plans_name': 'nnUNetPlans', 'transpose_forward': [0, 1, 2], 'transpose_backward': [0, 1, 2],
    'image_reader_writer': 'SimpleITKIO', 'experiment_planner_used': 'ExperimentPlanner', 'label_manager': 'LabelManager',
    'foreground_intensity_properties_per_channel': {'0': {'mean': 0.0, 'std': 1.0, 'percentile_00_5': -1.0, 'percentile_99_5': 1.0}},
    'configurations': {
        '3d_fullres': {
            'data_identifier': 'nnUNetPlans_3d_fullres', 'preprocessor_name': 'DefaultPreprocessor', 'batch_size': 2,
            'patch_size': [128, 128, 128], 'median_image_size_in_voxels': [155.0, 240.0, 240.0], 'spacing': [1.0, 1.0, 1.0],
            'normalization_schemes': ['ZScoreNormalization'] * 4, 'use_mask_for_norm': [True] * 4,
            'UNet_class_name': 'PlainConvUNet', 'UNet_base_num_features': 32, 'n_conv_per_stage_encoder': [2, 2, 2, 2, 2, 2],
            'n_conv_per_stage_decoder': [2, 2, 2, 2, 2], 'num_pool_per_axis': [5, 5, 5],
            'pool_op_kernel_sizes': [[1, 1, 1], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
            'conv_kernel_sizes': [[3, 3, 3]] * 6, 'unet_max_num_features': 320,
            'resampling_fn_data': 'resample_data_or_seg_to_shape', 'resampling_fn_seg': 'resample_data_or_seg_to_shape',
            'resampling_fn_data_kwargs': {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_seg_kwargs': {'is_seg': True, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'resampling_fn_probabilities': 'resample_data_or_seg_to_shape',
            'resampling_fn_probabilities_kwargs': {'is_seg': False, 'order': 1, 'order_z': 0, 'force_separate_z': None},
            'batch_dice': False},
        '3d_resenc': {'inherits_from': '3d_fullres', 'UNet_class_name': 'ResidualEncoderUNet', 'n_conv_per_stage_encoder': [1, 3, 4, 6, 6, 6]},
    },
}


def main():
    nothing = lambda *a, **k: None      # noqa: E731

    def resampling_fn(name):
        return {'resample_data_or_seg_to_shape': resample_data_or_seg_to_shape}[name]

    def resample_data_or_seg_to_shape(*a, **k):
        raise AssertionError('not reached')

    _module('nnunetv2').__path__ = ['/nonexistent']
    _module('nnunetv2.preprocessing')
    _module('nnunetv2.preprocessing.resampling')
    _module('nnunetv2.preprocessing.resampling.utils', recursive_find_resampling_fn_by_name=resampling_fn)
    _module('batchgenerators')
    _module('batchgenerators.utilities')
    _module('batchgenerators.utilities.file_and_folder_operations', load_json=nothing, join=os.path.join)
    _module('nnunetv2.imageio')
    _module('nnunetv2.imageio.reader_writer_registry', recursive_find_reader_writer_by_name=nothing)
    _module('nnunetv2.utilities')
    _module('nnunetv2.utilities.find_class_by_name', recursive_find_python_class=nothing)
    _module('nnunetv2.utilities.label_handling')
    _module('nnunetv2.utilities.label_handling.label_handling', get_labelmanager_class_from_plans=nothing)
    _module('dynamic_network_architectures')
    _module('dynamic_network_architectures.building_blocks')
    _module('dynamic_network_architectures.building_blocks.helper',
            convert_dim_to_conv_op=lambda dim: {1: torch.nn.Conv1d, 2: torch.nn.Conv2d, 3: torch.nn.Conv3d}[dim],
            get_matching_instancenorm=lambda conv_op=None, dimension=None: {1: torch.nn.InstanceNorm1d, 2: torch.nn.InstanceNorm2d,
                                                                            3: torch.nn.InstanceNorm3d}[dimension])
    ph = _load(os.path.join(REF, 'utilities/plans_handling/plans_handler.py'), 'ref_plans_handler')

    import warnings
    out = {'plans': {'new': NEW_PLANS, 'old': OLD_PLANS}, 'answers': {}}
    for tag, plans in (('new', NEW_PLANS), ('old', OLD_PLANS)):
        pm = ph.PlansManager(deepcopy(plans))
        ans = {'dataset_name': pm.dataset_name, 'plans_name': pm.plans_name, 'transpose_forward': pm.transpose_forward,
               'transpose_backward': pm.transpose_backward, 'available_configurations': pm.available_configurations,
               'foreground_intensity_properties_per_channel': pm.foreground_intensity_properties_per_channel, 'configurations': {}}
        for name in pm.available_configurations:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                cm = pm.get_configuration(name)
            ans['configurations'][name] = {
                'patch_size': cm.patch_size, 'spacing': cm.spacing, 'batch_size': cm.batch_size, 'data_identifier': cm.data_identifier,
                'normalization_schemes': cm.normalization_schemes, 'use_mask_for_norm': cm.use_mask_for_norm,
                'network_arch_class_name': cm.network_arch_class_name, 'network_arch_init_kwargs': cm.network_arch_init_kwargs,
                'network_arch_init_kwargs_req_import': list(cm.network_arch_init_kwargs_req_import),
                'pool_op_kernel_sizes': cm.pool_op_kernel_sizes, 'previous_stage_name': cm.previous_stage_name,
                'next_stage_names': cm.next_stage_names,
                'resampling_fn_data_kwargs': dict(cm.resampling_fn_data.keywords),
                'resampling_fn_seg_kwargs': dict(cm.resampling_fn_seg.keywords),
                'resampling_fn_probabilities_kwargs': dict(cm.resampling_fn_probabilities.keywords),
            }
        out['answers'][tag] = ans
    # error behaviour
    pm = ph.PlansManager(deepcopy(NEW_PLANS))
    try:
        pm.get_configuration('2d')
        out['missing_configuration_error'] = None
    except Exception as e:          # noqa: BLE001
        out['missing_configuration_error'] = type(e).__name__
    with open(os.path.join(HERE, 'plans_golden.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote plans_golden.json', {k: list(v['configurations']) for k, v in out['answers'].items()}, out['missing_configuration_error'])


if __name__ == '__main__':
    main()
