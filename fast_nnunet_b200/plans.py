"""plans.json / dataset.json accessors the inference path reads, with the reference's names.

Mirrors (only the members the hot path and its callers touch):
  utilities/plans_handling/plans_handler.py:31-211  ConfigurationManager
  utilities/plans_handling/plans_handler.py:214-325 PlansManager
  utilities/label_handling/label_handling.py:21-245 LabelManager
  utilities/label_handling/label_handling.py:294-311 determine_num_input_channels
"""
from __future__ import annotations

import json
from copy import deepcopy
from typing import List, Tuple, Union

import numpy as np


def load_json(path: str):
    with open(path, 'r') as f:
        return json.load(f)


class ConfigurationManager(object):
    def __init__(self, configuration_dict: dict):
        self.configuration = configuration_dict
        if 'architecture' not in self.configuration:
            # old plans format -> new 'architecture' entry (plans_handler.py:36-97)
            name = self.configuration['UNet_class_name']
            if name == 'PlainConvUNet':
                cls = 'dynamic_network_architectures.architectures.unet.PlainConvUNet'
                key = 'n_conv_per_stage'
            elif name == 'ResidualEncoderUNet':
                cls = 'dynamic_network_architectures.architectures.residual_unet.ResidualEncoderUNet'
                key = 'n_blocks_per_stage'
            else:
                raise RuntimeError(f'Unknown architecture {name}. This conversion only supports '
                                   f'PlainConvUNet and ResidualEncoderUNet')
            n_stages = len(self.configuration['n_conv_per_stage_encoder'])
            dim = len(self.configuration['patch_size'])
            self.configuration['architecture'] = {
                'network_class_name': cls,
                'arch_kwargs': {
                    'n_stages': n_stages,
                    'features_per_stage': [min(self.configuration['UNet_base_num_features'] * 2 ** i,
                                               self.configuration['unet_max_num_features'])
                                           for i in range(n_stages)],
                    'conv_op': f'torch.nn.modules.conv.Conv{dim}d',
                    'kernel_sizes': deepcopy(self.configuration['conv_kernel_sizes']),
                    'strides': deepcopy(self.configuration['pool_op_kernel_sizes']),
                    key: deepcopy(self.configuration['n_conv_per_stage_encoder']),
                    'n_conv_per_stage_decoder': deepcopy(self.configuration['n_conv_per_stage_decoder']),
                    'conv_bias': True,
                    'norm_op': f'torch.nn.modules.instancenorm.InstanceNorm{dim}d',
                    'norm_op_kwargs': {'eps': 1e-05, 'affine': True},
                    'dropout_op': None, 'dropout_op_kwargs': None,
                    'nonlin': 'torch.nn.LeakyReLU', 'nonlin_kwargs': {'inplace': True},
                },
                '_kw_requires_import': ['conv_op', 'norm_op', 'dropout_op', 'nonlin'],
            }
            for k in ('UNet_class_name', 'UNet_base_num_features', 'n_conv_per_stage_encoder',
                      'n_conv_per_stage_decoder', 'num_pool_per_axis', 'pool_op_kernel_sizes',
                      'conv_kernel_sizes', 'unet_max_num_features'):
                self.configuration.pop(k, None)

    def __repr__(self):
        return self.configuration.__repr__()

    @property
    def patch_size(self) -> List[int]:
        return self.configuration['patch_size']

    @property
    def spacing(self) -> List[float]:
        return self.configuration['spacing']

    @property
    def batch_size(self) -> int:
        return self.configuration['batch_size']

    @property
    def data_identifier(self) -> str:
        return self.configuration['data_identifier']

    @property
    def normalization_schemes(self) -> List[str]:
        return self.configuration['normalization_schemes']

    @property
    def use_mask_for_norm(self) -> List[bool]:
        return self.configuration['use_mask_for_norm']

    @property
    def resampling_fn_data_kwargs(self) -> dict:
        return dict(self.configuration.get('resampling_fn_data_kwargs') or
                    {'is_seg': False, 'order': 3, 'order_z': 0, 'force_separate_z': None})

    @property
    def resampling_fn_seg_kwargs(self) -> dict:
        return dict(self.configuration.get('resampling_fn_seg_kwargs') or
                    {'is_seg': True, 'order': 1, 'order_z': 0, 'force_separate_z': None})

    @property
    def resampling_fn_probabilities_kwargs(self) -> dict:
        return dict(self.configuration.get('resampling_fn_probabilities_kwargs') or
                    {'is_seg': False, 'order': 1, 'order_z': 0, 'force_separate_z': None})

    @property
    def network_arch_class_name(self) -> str:
        return self.configuration['architecture']['network_class_name']

    @property
    def network_arch_init_kwargs(self) -> dict:
        return self.configuration['architecture']['arch_kwargs']

    @property
    def network_arch_init_kwargs_req_import(self) -> Union[Tuple[str, ...], List[str]]:
        return self.configuration['architecture']['_kw_requires_import']

    @property
    def pool_op_kernel_sizes(self):
        return self.configuration['architecture']['arch_kwargs']['strides']

    @property
    def previous_stage_name(self) -> Union[str, None]:
        return self.configuration.get('previous_stage')

    @property
    def next_stage_names(self) -> Union[List[str], None]:
        ret = self.configuration.get('next_stage')
        if isinstance(ret, str):
            ret = [ret]
        return ret


class LabelManager(object):
    def __init__(self, label_dict: dict, regions_class_order: Union[List[int], None],
                 force_use_labels: bool = False, inference_nonlin=None):
        if 'background' not in label_dict:
            raise RuntimeError('Background label not declared (remember that this should be label 0!)')
        bg = label_dict['background']
        if isinstance(bg, (tuple, list)):
            raise RuntimeError(f'Background label must be 0. Not a list. Not a tuple. Your background label: {bg}')
        assert int(bg) == 0, f'Background label must be 0. Your background label: {bg}'
        self.label_dict = label_dict
        self.regions_class_order = regions_class_order
        self._force_use_labels = force_use_labels
        self._has_regions = False if force_use_labels else any(
            isinstance(i, (tuple, list)) and len(i) > 1 for i in label_dict.values())
        self._ignore_label = label_dict.get('ignore')
        if self._ignore_label is not None:
            assert isinstance(self._ignore_label, int), 'Ignore label has to be an integer.'
        labels = []
        for k, r in label_dict.items():
            if k == 'ignore':
                continue
            if isinstance(r, (tuple, list)):
                labels += [int(ri) for ri in r]
            else:
                labels.append(int(r))
        self._all_labels = sorted(int(i) for i in np.unique(labels))
        self._regions = None
        if self._has_regions:
            assert regions_class_order is not None, \
                'if region-based training is requested then you need to define regions_class_order!'
            regions = []
            for k, r in label_dict.items():
                if k == 'ignore':
                    continue
                if (np.isscalar(r) and r == 0) or \
                        (isinstance(r, (tuple, list)) and len(np.unique(r)) == 1 and np.unique(r)[0] == 0):
                    continue
                regions.append(tuple(r) if isinstance(r, list) else r)
            assert len(regions_class_order) == len(regions)
            self._regions = regions
        if self.has_ignore_label:
            assert self.ignore_label == max(self.all_labels) + 1
        self.inference_nonlin = inference_nonlin

    @property
    def has_regions(self) -> bool:
        return self._has_regions

    @property
    def has_ignore_label(self) -> bool:
        return self._ignore_label is not None

    @property
    def ignore_label(self):
        return self._ignore_label

    @property
    def all_labels(self) -> List[int]:
        return self._all_labels

    @property
    def all_regions(self):
        return self._regions

    @property
    def foreground_labels(self) -> List[int]:
        return [i for i in self.all_labels if i != 0]

    @property
    def foreground_regions(self):
        return [i for i in self.all_regions if i != 0 and i != (0,)] if self._regions is not None else None

    @property
    def num_segmentation_heads(self) -> int:
        if self.has_regions:
            return len(self.foreground_regions)
        return len(self.all_labels)

    def convert_logits_to_segmentation(self, predicted_logits):
        """label_handling.py:184-195 on host arrays: argmax over heads (first maximum) for label
        training; sigmoid > 0.5 painted in regions_class_order for region training."""
        import torch
        is_torch = isinstance(predicted_logits, torch.Tensor)
        arr = predicted_logits.detach().float().cpu().numpy() if is_torch else np.asarray(predicted_logits)
        assert arr.shape[0] == self.num_segmentation_heads
        if self.has_regions:
            prob = 1.0 / (1.0 + np.exp(-arr.astype(np.float32)))
            seg = np.zeros(arr.shape[1:], dtype=np.uint16)
            for i, c in enumerate(self.regions_class_order):
                seg[prob[i] > 0.5] = c
        else:
            seg = arr.argmax(0)
        return torch.from_numpy(seg.astype(np.int64)) if is_torch else seg


class PlansManager(object):
    def __init__(self, plans_file_or_dict: Union[str, dict]):
        self.plans = plans_file_or_dict if isinstance(plans_file_or_dict, dict) else load_json(plans_file_or_dict)
        self._cache = {}

    def __repr__(self):
        return self.plans.__repr__()

    def _internal_resolve_configuration_inheritance(self, configuration_name: str, visited=None) -> dict:
        if configuration_name not in self.plans['configurations']:
            raise ValueError(f'The configuration {configuration_name} does not exist in the plans I have. Valid '
                             f'configuration names are {list(self.plans["configurations"].keys())}.')
        configuration = deepcopy(self.plans['configurations'][configuration_name])
        if 'inherits_from' in configuration:
            parent = configuration['inherits_from']
            if visited is None:
                visited = (configuration_name,)
            else:
                if parent in visited:
                    raise RuntimeError(f'Circular dependency detected while resolving {visited} -> {parent}')
                visited = (*visited, configuration_name)
            base = self._internal_resolve_configuration_inheritance(parent, visited)
            base.update(configuration)
            configuration = base
        return configuration

    def get_configuration(self, configuration_name: str) -> ConfigurationManager:
        if configuration_name not in self.plans['configurations']:
            raise RuntimeError(f'Requested configuration {configuration_name} not found in plans. '
                               f'Available configurations: {list(self.plans["configurations"].keys())}')
        if configuration_name not in self._cache:
            self._cache[configuration_name] = ConfigurationManager(
                self._internal_resolve_configuration_inheritance(configuration_name))
        return self._cache[configuration_name]

    @property
    def dataset_name(self) -> str:
        return self.plans['dataset_name']

    @property
    def plans_name(self) -> str:
        return self.plans['plans_name']

    @property
    def transpose_forward(self) -> List[int]:
        return self.plans['transpose_forward']

    @property
    def transpose_backward(self) -> List[int]:
        return self.plans['transpose_backward']

    @property
    def available_configurations(self) -> List[str]:
        return list(self.plans['configurations'].keys())

    def get_label_manager(self, dataset_json: dict, **kwargs) -> LabelManager:
        return LabelManager(label_dict=dataset_json['labels'],
                            regions_class_order=dataset_json.get('regions_class_order'), **kwargs)

    @property
    def foreground_intensity_properties_per_channel(self) -> dict:
        if 'foreground_intensity_properties_per_channel' not in self.plans:
            if 'foreground_intensity_properties_by_modality' in self.plans:
                return self.plans['foreground_intensity_properties_by_modality']
        return self.plans['foreground_intensity_properties_per_channel']


def determine_num_input_channels(plans_manager: PlansManager, configuration_or_config_manager,
                                 dataset_json: dict) -> int:
    cm = plans_manager.get_configuration(configuration_or_config_manager) \
        if isinstance(configuration_or_config_manager, str) else configuration_or_config_manager
    lm = plans_manager.get_label_manager(dataset_json)
    n_mod = len(dataset_json['modality']) if 'modality' in dataset_json else len(dataset_json['channel_names'])
    if cm.previous_stage_name is not None:
        return n_mod + len(lm.foreground_labels)
    return n_mod
