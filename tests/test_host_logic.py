"""CPU: program lowering, plans/label accessors, model-folder round trip, predictor initialisation."""
import os

import numpy as np
import pytest
import torch

from fast_nnunet_b200 import _lib, model_folder as M, program as P
from fast_nnunet_b200.plans import ConfigurationManager, LabelManager, PlansManager
from fast_nnunet_b200.predictor import nnUNetPredictor, _infer_arch_from_weights

KW_STUDENT = M.plain_arch_kwargs([16, 32, 64, 128, 160, 160], [[3, 3, 3]] * 6, [[1, 1, 1]] + [[2, 2, 2]] * 5)
KW_TEACHER = M.plain_arch_kwargs([32, 64, 128, 256, 320, 320], [[3, 3, 3]] * 6, [[1, 1, 1]] + [[2, 2, 2]] * 5)


def test_flops_match_survey():
    sd = M.synthesize_state_dict(M.PLAIN, KW_STUDENT, 1, 2)
    p = P.build_program(M.PLAIN, sd, KW_STUDENT, 1, 2, (128, 128, 128))
    assert abs(p.total_flops() / 1e9 - 239.6) < 0.1
    assert len([o for o in p.ops if o.op == _lib.OP_CONV]) == 23 and len([o for o in p.ops if o.op == _lib.OP_TCONV]) == 5
    kw = M.resenc_arch_kwargs([16, 32, 64, 128, 160, 160], [[3, 3, 3]] * 6, [[1, 1, 1]] + [[2, 2, 2]] * 5, [1, 3, 4, 6, 6, 6])
    sd = M.synthesize_state_dict(M.RESENC, kw, 4, 4)
    p = P.build_program(M.RESENC, sd, kw, 4, 4, (128, 128, 128))
    assert abs(p.total_flops() / 1e9 - 365.2) < 0.1


def test_concat_is_written_in_place():
    sd = M.synthesize_state_dict(M.PLAIN, KW_STUDENT, 1, 2)
    p = P.build_program(M.PLAIN, sd, KW_STUDENT, 1, 2, (128, 128, 128))
    tconvs = [o for o in p.ops if o.op == _lib.OP_TCONV]
    for t in tconvs:
        dims, ch = p.buffers[t.dst]
        assert t.dst_coff == 0 and ch == 2 * t.cout
        # the encoder conv that produced the skip wrote channels [c, 2c) of the same buffer
        writers = [o for o in p.ops if o.dst == t.dst and o.op == _lib.OP_CONV]
        assert len(writers) == 1 and writers[0].dst_coff == t.cout and writers[0].cout == t.cout
        readers = [o for o in p.ops if o.src == t.dst and o.cin == 2 * t.cout]
        assert len(readers) == 1


def test_student_rule_and_inference_from_weights():
    assert M.student_features([32, 64, 128, 256, 320, 320], 2) == [16, 32, 64, 128, 160, 160]
    assert M.student_features([32, 64, 128, 256, 320, 320], 8) == [8, 8, 16, 32, 40, 40]
    assert M.student_blocks([1, 3, 4, 6, 6, 6], None, None, 'reduce') == [1, 1, 2, 3, 3, 3]
    assert M.student_blocks([1, 3, 4, 6, 6, 6], None, None, 'increase') == [2, 4, 5, 7, 7, 7]
    cls, kw = M.effective_arch(M.PLAIN, KW_TEACHER, 'nnUNetDistillationTrainer', {'feature_reduction_factor': 2})
    assert cls == M.PLAIN and kw['features_per_stage'] == [16, 32, 64, 128, 160, 160]
    sd = M.synthesize_state_dict(M.PLAIN, KW_STUDENT, 1, 2)
    kw2 = _infer_arch_from_weights(M.PLAIN, KW_TEACHER, sd)     # plans say teacher, weights say student
    assert kw2['features_per_stage'] == [16, 32, 64, 128, 160, 160]
    assert kw2['n_conv_per_stage'] == [2] * 6 and kw2['n_conv_per_stage_decoder'] == [2] * 5 and kw2['conv_bias']


def test_prefix_stripping():
    sd = M.synthesize_state_dict(M.RESENC, M.resenc_arch_kwargs([8, 16], [[3, 3, 3]] * 2, [[1, 1, 1], [2, 2, 2]], [1, 2]),
                                 1, 2, prefix='module.network.')
    clean = P.clean_state_dict(sd)
    assert 'encoder.stem.convs.0.conv.weight' in clean and not any(k.startswith('module.') for k in clean)


def test_missing_weight_raises():
    sd = M.synthesize_state_dict(M.PLAIN, KW_STUDENT, 1, 2)
    del sd['decoder.transpconvs.2.weight']
    with pytest.raises(KeyError):
        P.build_program(M.PLAIN, sd, KW_STUDENT, 1, 2, (128, 128, 128))
    with pytest.raises(RuntimeError):
        P.build_program('some.other.Primus', sd, KW_STUDENT, 1, 2, (128, 128, 128))


def test_label_manager():
    lm = LabelManager({'background': 0, 'a': 1, 'b': 2}, None)
    assert lm.num_segmentation_heads == 3 and lm.foreground_labels == [1, 2] and not lm.has_regions
    logits = torch.tensor([[[[0.5]], [[1.0]]], [[[0.5]], [[1.0]]], [[[0.1]], [[2.0]]]])
    assert lm.convert_logits_to_segmentation(logits).flatten().tolist() == [0, 2]     # tie -> first maximum
    with pytest.raises(RuntimeError):
        LabelManager({'a': 1}, None)
    lr = LabelManager({'background': 0, 'whole': [1, 2, 3], 'core': [2, 3], 'enh': 3}, [1, 2, 3])
    assert lr.has_regions and lr.num_segmentation_heads == 3


def test_old_plans_format_is_upgraded():
    cm = ConfigurationManager({'UNet_class_name': 'PlainConvUNet', 'UNet_base_num_features': 32,
                               'unet_max_num_features': 320, 'n_conv_per_stage_encoder': [2] * 6,
                               'n_conv_per_stage_decoder': [2] * 5, 'num_pool_per_axis': [5, 5, 5],
                               'pool_op_kernel_sizes': [[1, 1, 1]] + [[2, 2, 2]] * 5, 'conv_kernel_sizes': [[3, 3, 3]] * 6,
                               'patch_size': [128, 128, 128]})
    assert cm.network_arch_init_kwargs['features_per_stage'] == [32, 64, 128, 256, 320, 320]
    assert cm.network_arch_class_name.endswith('PlainConvUNet')


def test_configuration_inheritance():
    pm = PlansManager({'configurations': {'a': {'patch_size': [8, 8, 8], 'spacing': [1, 1, 1], 'architecture': {}},
                                          'b': {'inherits_from': 'a', 'patch_size': [16, 16, 16]}}})
    assert pm.get_configuration('b').patch_size == [16, 16, 16] and pm.get_configuration('b').spacing == [1, 1, 1]
    with pytest.raises(RuntimeError):
        pm.get_configuration('zzz')


def test_model_folder_round_trip_and_student_loading(tmp_path):
    sd = M.synthesize_state_dict(M.PLAIN, KW_STUDENT, 1, 2)
    folder = M.write_model_folder(str(tmp_path / 'nnUNetDistillationTrainer__nnUNetPlans__3d_fullres'), M.PLAIN,
                                  KW_TEACHER, (128, 128, 128), sd, 1, 2, trainer_name='nnUNetDistillationTrainer',
                                  init_args_extra={'feature_reduction_factor': 2, 'block_reduction_strategy': 'keep'})
    p = nnUNetPredictor(device=torch.device('cpu'))
    assert p.perform_everything_on_device is False
    p.initialize_from_trained_model_folder(folder, use_folds=None)
    assert p.trainer_name == 'nnUNetDistillationTrainer' and p.allowed_mirroring_axes == (0, 1, 2)
    assert p.label_manager.num_segmentation_heads == 2 and len(p.list_of_parameters) == 1
    assert p.network.arch_kwargs['features_per_stage'] == [16, 32, 64, 128, 160, 160]
    assert p._flip_masks() == bytes([0, 1, 2, 4, 3, 5, 6, 7])
    p.use_mirroring = False
    assert p._flip_masks() == bytes([0])
    sl = p._internal_get_sliding_window_slicers((160, 160, 160))
    assert len(sl) == 8 and sl[1] == (slice(None), slice(0, 128), slice(0, 128), slice(32, 160))
    # no CPU path: predicting without a CUDA device must fail loudly
    with pytest.raises((RuntimeError, AssertionError, FileNotFoundError)):
        p.predict_sliding_window_return_logits(torch.zeros(1, 128, 128, 128))
    with pytest.raises(AssertionError):
        p.predict_sliding_window_return_logits(np.zeros((1, 128, 128, 128), dtype=np.float32))
