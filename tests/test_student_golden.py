"""CPU: model_folder.effective_arch (which architecture a distilled-student checkpoint holds) against the keyword
arguments the reference's own nnUNetDistillationTrainer.build_network_architecture passes to LiteNNUNetStudent /
LiteResEncStudent — the method's source executed unchanged by tests/golden/make_student_golden.py (108 combinations of
plans, student_plans_identifier, feature_reduction_factor and block_reduction_strategy)."""
import json
import os

from fast_nnunet_b200 import model_folder as M

HERE = os.path.dirname(os.path.abspath(__file__))
T = json.load(open(os.path.join(HERE, 'golden', 'student_golden.json')))


def test_effective_arch_equals_executed_reference():
    checked = 0
    for c in T['cases']:
        built = c['built']
        if 'error' in built:          # plain student from ResEnc plans: the reference itself raises
            continue
        cls_in = M.RESENC if 'n_blocks_per_stage' in c['arch_kwargs'] else M.PLAIN
        init_args = {'feature_reduction_factor': c['feature_reduction_factor'],
                     'block_reduction_strategy': c['block_reduction_strategy'],
                     'student_plans_identifier': c['student_plans_identifier']}
        cls, kw = M.effective_arch(cls_in, c['arch_kwargs'], 'nnUNetDistillationTrainer', init_args)
        want_resenc = built['class'] == 'LiteResEncStudent'
        assert (cls == M.RESENC) == want_resenc, c
        assert kw['n_stages'] == built['n_stages']
        assert [int(f) for f in kw['features_per_stage']] == built['features_per_stage'], c
        assert [list(k) for k in kw['kernel_sizes']] == built['kernel_sizes']
        assert [list(s) for s in kw['strides']] == built['strides']
        assert list(kw['n_conv_per_stage_decoder']) == built['n_conv_per_stage_decoder']
        assert bool(kw['conv_bias']) == built['conv_bias']
        if want_resenc:
            assert list(kw['n_blocks_per_stage']) == built['n_blocks_per_stage'], c
        else:
            assert list(kw['n_conv_per_stage']) == built['n_conv_per_stage'], c
        checked += 1
    assert checked >= 100
