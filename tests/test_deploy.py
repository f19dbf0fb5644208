"""f4: deployment formats — the ONNX files the reference's exporters write and the engine `.ini` /
inferencer JSON configs.  tests/golden/tiny_plain_unet.onnx was written by torch's own ONNX exporter
(tests/golden/make_onnx_golden.py); the importer must give back the exact weights and the architecture."""
import json
import os

import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import deploy, onnx_import

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tiny_plain_unet.onnx')

INI = """[model]
file_name = tiny_plain_unet.trt
input_name = input
output_name = output
num_class = 3

[input]
depth = 16
height = 16
width = 16
patch_size = 16, 16, 16
target_spacing = 2.0, 0.9765625, 0.9765625

[preprocessing]
mean = 418.6798400878906
std_dev = 412.1883239746094
lower_bound = -60.0
upper_bound = 3068.0

[inference]
use_mirroring = false
step_size = 0.5
use_gaussian = true
"""


def test_onnx_graph_gives_back_weights_and_architecture():
    spec = nets.TINY_ONNX
    sd, _ = nets.make(spec, seed=77, randomize_affine=True)
    g = onnx_import.read_onnx(GOLDEN)
    assert g.opset == 17 and g.inputs == ['input'] and g.outputs == ['output']
    kw, got, info = onnx_import.plain_conv_unet_from_onnx(GOLDEN)
    assert info['input_channels'] == 1 and info['num_heads'] == 3
    for k in ('n_stages', 'features_per_stage', 'kernel_sizes', 'strides', 'n_conv_per_stage', 'n_conv_per_stage_decoder',
              'conv_bias'):
        assert kw[k] == spec['kw'][k], (k, kw[k], spec['kw'][k])
    assert abs(info['negative_slope'] - 0.01) < 1e-8
    assert len(got) == 4 * 5 + 4 * 3 + 2 * 2 + 2            # conv + norm tensors, transposed convs, head
    for k, v in got.items():
        assert np.array_equal(v, sd[k].numpy()), k            # bit-exact


def test_imported_weights_lower_to_the_same_program():
    from fast_nnunet_b200.program import build_program
    spec = nets.TINY_ONNX
    sd, _ = nets.make(spec, seed=77, randomize_affine=True)
    kw, got, info = onnx_import.plain_conv_unet_from_onnx(GOLDEN)
    a = build_program(spec['cls'], sd, spec['kw'], 1, 3, spec['patch'])
    b = build_program(spec['cls'], {k: torch.from_numpy(v) for k, v in got.items()}, kw, 1, 3, spec['patch'])
    assert a.buffers == b.buffers and len(a.ops) == len(b.ops)
    for x, y in zip(a.ops, b.ops):
        assert (x.op, x.src, x.dst, x.cin, x.cout, tuple(x.kernel), tuple(x.stride)) == \
               (y.op, y.src, y.dst, y.cin, y.cout, tuple(y.kernel), tuple(y.stride))
        assert np.array_equal(x.weight, y.weight) and abs(x.act_slope - y.act_slope) < 1e-8


def test_non_unet_graph_is_refused(tmp_path):
    # a truncated file and a graph with an operator outside the family must raise, not guess
    bad = tmp_path / 'bad.onnx'
    bad.write_bytes(open(GOLDEN, 'rb').read()[:4096])
    with pytest.raises(Exception):
        onnx_import.plain_conv_unet_from_onnx(str(bad))


def test_engine_ini_and_inferencer_json(tmp_path):
    ini = tmp_path / 'model.ini'
    ini.write_text(INI)
    c = deploy.read_engine_ini(str(ini))
    assert c['num_class'] == 3 and c['patch_size'] == [16, 16, 16] and c['target_spacing'] == [2.0, 0.9765625, 0.9765625]
    assert c['intensity_properties'] == {'mean': 418.6798400878906, 'std': 412.1883239746094, 'percentile_00_5': -60.0,
                                         'percentile_99_5': 3068.0}
    assert c['use_mirroring'] is False and c['step_size'] == 0.5 and c['use_gaussian'] is True
    js = tmp_path / 'cfg.json'
    js.write_text(json.dumps({'patch_size': [56, 160, 192], 'target_spacing': [3.0, 0.78, 0.78],
                              'intensity_properties': {'mean': 85.8, 'std': 108.0, 'percentile_00_5': -913.0,
                                                       'percentile_99_5': 284.0}, 'model_path': 'm.onnx'}))
    d = deploy.read_deployment_config(str(js))
    assert d['patch_size'] == [56, 160, 192] and d['intensity_properties']['percentile_00_5'] == -913.0
    plans, dataset = deploy.plans_from_deployment(c, nets.TINY_ONNX['kw'], 1)
    from fast_nnunet_b200.plans import PlansManager
    pm = PlansManager(plans)
    cm = pm.get_configuration('3d_fullres')
    assert list(cm.patch_size) == [16, 16, 16] and cm.normalization_schemes == ['CTNormalization']
    assert pm.get_label_manager(dataset).num_segmentation_heads == 3


@pytest.mark.gpu
def test_predictor_from_deployment_matches_oracle(tmp_path):
    """`.ini` + ONNX -> predictor -> raw CT array in, label map out, against the oracle chain on the same weights."""
    import shutil
    from oracle import export as OX
    from oracle import predictor as OPR
    from oracle import preprocess as OPP
    dev = torch.device('cuda', 0)
    spec = nets.TINY_ONNX
    _, net = nets.make(spec, seed=77, randomize_affine=True)
    (tmp_path / 'model.ini').write_text(INI)
    shutil.copy(GOLDEN, tmp_path / 'tiny_plain_unet.onnx')
    pred = deploy.predictor_from_deployment(str(tmp_path / 'model.ini'), device=dev)
    assert pred.use_mirroring is False and pred.label_manager.num_segmentation_heads == 3
    g = np.random.default_rng(0)
    img = g.normal(400, 500, size=(1, 20, 36, 40)).astype(np.float32)       # no zero border: nothing is cropped
    spacing = [2.0, 0.9765625, 0.9765625]
    seg = pred.predict_single_npy_array(img, {'spacing': spacing})
    cfg = deploy.read_engine_ini(str(tmp_path / 'model.ini'))
    data, props = OPP.run_case_npy(img.copy(), {'spacing': spacing}, [0, 1, 2], cfg['target_spacing'], ['CTNormalization'],
                                   [False], {'0': cfg['intensity_properties']})
    logits = OPR.predict_sliding_window_return_logits(net.to(dev), torch.from_numpy(data).to(dev), spec['patch'], 0.5,
                                                      True, None).cpu().numpy()
    want = OX.convert_predicted_logits_to_segmentation_with_correct_shape(logits, cfg['target_spacing'], [0, 1, 2],
                                                                          [0, 1, 2], props)
    assert seg.shape == img.shape[1:] == tuple(logits.shape[1:])
    # random-init weights: near-tied logits, so the bar is on the logits' consequence only where the margin is real
    lg = torch.from_numpy(logits.astype(np.float32))
    top2 = torch.topk(lg, 2, dim=0).values
    confident = ((top2[0] - top2[1]) > 0.05).numpy()
    assert float((seg == want)[confident].mean()) >= 0.9999
    assert float((seg == want).mean()) >= 0.98
