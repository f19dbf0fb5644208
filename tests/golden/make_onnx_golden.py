"""Writes tests/golden/tiny_plain_unet.onnx with torch's own TorchScript ONNX exporter, called the way the
reference's exporter calls it (distillation/fast_nnunet_distillation_export_onnx.py:455-467: opset 17, names
'input' / 'output', dynamic batch, constant folding).  The `onnx` Python package is absent from this image; the
exporter only needs it for a post-processing hook that is a no-op for this graph, so that hook is bypassed.
Run here (CPU):  python tests/golden/make_onnx_golden.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import nets  # noqa: E402
from torch.onnx._internal.torchscript_exporter import onnx_proto_utils as U  # noqa: E402

U._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes

sd, net = nets.make(nets.TINY_ONNX, seed=77, randomize_affine=True)
x = torch.randn((1, nets.TINY_ONNX['in_ch'], *nets.TINY_ONNX['patch']))
out = os.path.join(ROOT, 'tests', 'golden', 'tiny_plain_unet.onnx')
torch.onnx.export(net, x, out, export_params=True, opset_version=17, do_constant_folding=True, input_names=['input'],
                  output_names=['output'], dynamic_axes={'input': {0: 'batch_size'}, 'output': {0: 'batch_size'}},
                  training=torch.onnx.TrainingMode.EVAL, dynamo=False)
print(out, os.path.getsize(out), 'bytes')
