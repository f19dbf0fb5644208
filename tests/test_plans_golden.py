"""CPU: fast_nnunet_b200.plans against answers produced by EXECUTING the reference's plans_handler.py
(tests/golden/make_plans_golden.py): configuration inheritance, the accessors the predictor reads, old-format plans."""
import json
import os
import warnings
from copy import deepcopy

import pytest

from fast_nnunet_b200 import plans as P

HERE = os.path.dirname(os.path.abspath(__file__))
T = json.load(open(os.path.join(HERE, 'golden', 'plans_golden.json')))


@pytest.mark.parametrize('tag', ['new', 'old'])
def test_plans_manager_equals_reference(tag):
    want = T['answers'][tag]
    pm = P.PlansManager(deepcopy(T['plans'][tag]))
    assert pm.dataset_name == want['dataset_name'] and pm.plans_name == want['plans_name']
    assert list(pm.transpose_forward) == want['transpose_forward']
    assert list(pm.transpose_backward) == want['transpose_backward']
    assert list(pm.available_configurations) == want['available_configurations']
    assert pm.foreground_intensity_properties_per_channel == want['foreground_intensity_properties_per_channel']
    for name, w in want['configurations'].items():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            cm = pm.get_configuration(name)
        got = {
            'patch_size': cm.patch_size, 'spacing': cm.spacing, 'batch_size': cm.batch_size, 'data_identifier': cm.data_identifier,
            'normalization_schemes': cm.normalization_schemes, 'use_mask_for_norm': cm.use_mask_for_norm,
            'network_arch_class_name': cm.network_arch_class_name, 'network_arch_init_kwargs': cm.network_arch_init_kwargs,
            'network_arch_init_kwargs_req_import': list(cm.network_arch_init_kwargs_req_import),
            'pool_op_kernel_sizes': cm.pool_op_kernel_sizes, 'previous_stage_name': cm.previous_stage_name,
            'next_stage_names': cm.next_stage_names,
            'resampling_fn_data_kwargs': cm.resampling_fn_data_kwargs, 'resampling_fn_seg_kwargs': cm.resampling_fn_seg_kwargs,
            'resampling_fn_probabilities_kwargs': cm.resampling_fn_probabilities_kwargs,
        }
        for k in w:
            assert json.loads(json.dumps(got[k])) == w[k], (tag, name, k, got[k], w[k])


def test_missing_configuration_raises_like_the_reference():
    pm = P.PlansManager(deepcopy(T['plans']['new']))
    with pytest.raises(Exception) as e:
        pm.get_configuration('2d')
    assert type(e.value).__name__ == T['missing_configuration_error']
