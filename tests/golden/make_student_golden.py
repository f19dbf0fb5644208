"""Golden table for the distilled-student architecture rule, produced by EXECUTING the reference's own method
nnUNetDistillationTrainer.build_network_architecture (training/nnUNetTrainer/variants/nnUNetDistillationTrainer.py:605-759).

Run in the build container only (needs /root/reference):
    python tests/golden/make_student_golden.py

The trainer module cannot be imported here (it pulls in the whole training stack), so the METHOD's source is cut out of
the file with `ast`, compiled unchanged and called with a stand-in `self`; LiteNNUNetStudent / LiteResEncStudent are
recorders that keep the keyword arguments the method passes.  Pins fast_nnunet_b200.model_folder.effective_arch
(feature reduction, floor of 8 features, the four ResEnc block strategies, plain vs ResEnc by student_plans_identifier).

Writes tests/golden/student_golden.json.
"""
import ast
import json
import os
import types
from typing import List, Tuple, Union

SRC = '/root/reference/distillation/nnunetv2/training/nnUNetTrainer/variants/nnUNetDistillationTrainer.py'
HERE = os.path.dirname(os.path.abspath(__file__))

ISO_K = [[3, 3, 3]] * 6
ISO_S = [[1, 1, 1]] + [[2, 2, 2]] * 5
PLAIN_KW = {'n_stages': 6, 'features_per_stage': [32, 64, 128, 256, 320, 320], 'conv_op': 'torch.nn.modules.conv.Conv3d',
            'kernel_sizes': ISO_K, 'strides': ISO_S, 'n_conv_per_stage': [2] * 6, 'n_conv_per_stage_decoder': [2] * 5,
            'conv_bias': True, 'norm_op': 'torch.nn.modules.instancenorm.InstanceNorm3d', 'norm_op_kwargs': {'eps': 1e-05, 'affine': True},
            'dropout_op': None, 'dropout_op_kwargs': None, 'nonlin': 'torch.nn.LeakyReLU', 'nonlin_kwargs': {'inplace': True}}
RESENC_KW = dict(PLAIN_KW)
RESENC_KW.pop('n_conv_per_stage')
RESENC_KW['n_blocks_per_stage'] = [1, 3, 4, 6, 6, 6]
RESENC_KW['n_conv_per_stage_decoder'] = [1] * 5
ANISO_KW = dict(PLAIN_KW)
ANISO_KW.update({'n_stages': 5, 'features_per_stage': [32, 64, 128, 256, 320],
                 'kernel_sizes': [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                 'strides': [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]],
                 'n_conv_per_stage': [2] * 5, 'n_conv_per_stage_decoder': [2] * 4})


class _Recorder:
    def __init__(self, name):
        self.name = name

    def __call__(self, **kw):
        return {'class': self.name, **{k: v for k, v in kw.items() if k not in ('conv_op', 'norm_op', 'nonlin')}}


def main():
    tree = ast.parse(open(SRC).read())
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == 'build_network_architecture':
            fn = node
    assert fn is not None
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {'Union': Union, 'List': List, 'Tuple': Tuple, 'LiteNNUNetStudent': _Recorder('LiteNNUNetStudent'),
          'LiteResEncStudent': _Recorder('LiteResEncStudent'), 'Conv3d': 'Conv3d', 'InstanceNorm3d': 'InstanceNorm3d',
          'LeakyReLU': 'LeakyReLU', 'determine_num_input_channels': None}
    exec(compile(mod, SRC, 'exec'), ns)
    build = ns['build_network_architecture']

    cases = []
    for tag, kw in (('plain', PLAIN_KW), ('resenc', RESENC_KW), ('aniso', ANISO_KW)):
        for ident in ('nnUNetPlans', 'nnUNetResEncUNetMPlans'):
            for r in (1, 2, 3, 4, 6, 16):
                for strategy in ('reduce', 'keep', 'increase', 'adaptive', 'something_else'):
                    if ident == 'nnUNetPlans' and strategy != 'keep':
                        continue                   # the strategy only matters for the ResEnc student
                    me = types.SimpleNamespace(
                        student_plans_identifier=ident, feature_reduction_factor=r, block_reduction_strategy=strategy,
                        print_to_log_file=lambda *a, **k: None, _do_i_compile=lambda: False,
                        configuration_manager=types.SimpleNamespace(configuration={'architecture': {'arch_kwargs': json.loads(json.dumps(kw))}}))
                    try:
                        got = build(me, num_input_channels=1, num_output_channels=3, enable_deep_supervision=False)
                        got = json.loads(json.dumps(got))        # tuples -> lists
                    except KeyError as e:                        # a plain student asked from ResEnc plans: the reference fails
                        got = {'error': 'KeyError', 'key': str(e)}
                    cases.append({'tag': tag, 'arch_kwargs': kw, 'student_plans_identifier': ident, 'feature_reduction_factor': r,
                                  'block_reduction_strategy': strategy, 'built': got})
    with open(os.path.join(HERE, 'student_golden.json'), 'w') as f:
        json.dump({'cases': cases}, f)
    print('wrote student_golden.json,', len(cases), 'cases;', sum('error' in c['built'] for c in cases), 'where the reference raises; e.g.', cases[7]['built'])


if __name__ == '__main__':
    main()
