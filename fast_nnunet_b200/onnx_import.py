"""Weights and architecture of a PlainConvUNet from the ONNX files the reference's exporters write
(distillation/fast_nnunet_distillation_export_onnx.py:432-472 — torch.onnx.export, opset 17, graph
input 'input', output 'output', dynamic batch; fast_nnunet_distillation_export_onnx_dynamic.py likewise with opset 11).

No `onnx` package is needed (it is absent from this image): the file is protobuf wire format, and the handful of
messages a weight import needs (ModelProto.graph, GraphProto.node / initializer / input / output, NodeProto,
AttributeProto, TensorProto) are decoded here directly; field numbers are those of onnx/onnx.proto.

The graph is read in execution order, which is how the TorchScript exporter emits it:
    encoder stage s : n x (Conv -> InstanceNormalization -> LeakyRelu); a stride on the stage's first Conv
    decoder level l : ConvTranspose -> Concat(up, skip) -> n x (Conv -> InstanceNormalization -> LeakyRelu)
    head            : Conv 1x1x1 (deep supervision is off in the exported graphs)
From it the importer rebuilds plans-style `arch_kwargs` and a `state_dict` with the key layout real nnU-Net
checkpoints carry, so that everything downstream (program lowering, engines, predictor) is the checkpoint path.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------- protobuf wire format


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yields (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    mv = memoryview(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(mv[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(mv[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = bytes(mv[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        yield fno, wt, v


def _packed_varints(v: bytes) -> List[int]:
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    """TensorProto: dims = 1, data_type = 2, float_data = 4, int32_data = 5, int64_data = 7, name = 8, raw_data = 9."""
    dims, dtype, name, raw = [], 1, '', None
    floats, i32, i64 = [], [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v) if wt == 2 else [v]
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = v.decode()
        elif fno == 9:
            raw = v
        elif fno == 4:
            floats += list(struct.unpack(f'<{len(v) // 4}f', v)) if wt == 2 else [struct.unpack('<f', v)[0]]
        elif fno == 5:
            i32 += _packed_varints(v) if wt == 2 else [v]
        elif fno == 7:
            i64 += _packed_varints(v) if wt == 2 else [v]
    if dtype not in _DTYPES:
        raise ValueError(f'tensor {name!r}: ONNX data type {dtype} is not supported')
    np_t = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_t).newbyteorder('<')).astype(np_t)
    elif floats:
        arr = np.asarray(floats, dtype=np_t)
    elif i64:
        arr = np.asarray([_signed(x) for x in i64], dtype=np_t)
    else:
        arr = np.asarray([_signed(x) for x in i32], dtype=np_t)
    return name, arr.reshape([int(d) for d in dims]) if dims else arr.reshape(())


def _attribute(buf: bytes):
    """AttributeProto: name = 1, f = 2, i = 3, s = 4, t = 5, floats = 7, ints = 8."""
    name, val = '', None
    ints, floats = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            val = struct.unpack('<f', v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = v.decode(errors='replace')
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            floats += list(struct.unpack(f'<{len(v) // 4}f', v)) if wt == 2 else [struct.unpack('<f', v)[0]]
        elif fno == 8:
            ints += [_signed(x) for x in _packed_varints(v)] if wt == 2 else [_signed(v)]
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


class Node:
    def __init__(self, op_type, inputs, outputs, attrs, name):
        self.op_type, self.inputs, self.outputs, self.attrs, self.name = op_type, inputs, outputs, attrs, name

    def __repr__(self):
        return f'Node({self.op_type}, in={self.inputs}, out={self.outputs}, {self.attrs})'


def _node(buf: bytes) -> Node:
    """NodeProto: input = 1, output = 2, name = 3, op_type = 4, attribute = 5."""
    ins, outs, name, op, attrs = [], [], '', '', {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            ins.append(v.decode())
        elif fno == 2:
            outs.append(v.decode())
        elif fno == 3:
            name = v.decode()
        elif fno == 4:
            op = v.decode()
        elif fno == 5:
            k, a = _attribute(v)
            attrs[k] = a
    return Node(op, ins, outs, attrs, name)


def _value_info_name(buf: bytes) -> str:
    for fno, wt, v in _fields(buf):
        if fno == 1:
            return v.decode()
    return ''


class Graph:
    def __init__(self):
        self.nodes: List[Node] = []
        self.initializers: Dict[str, np.ndarray] = {}
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.opset: Optional[int] = None


def read_onnx(path: str) -> Graph:
    """ModelProto: graph = 7, opset_import = 8 (OperatorSetIdProto.version = 2);
    GraphProto: node = 1, initializer = 5, input = 11, output = 12."""
    data = open(path, 'rb').read()
    g = Graph()
    for fno, wt, v in _fields(data):
        if fno == 8:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    g.opset = v2 if g.opset is None else max(g.opset, v2)
        elif fno == 7:
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    g.nodes.append(_node(v2))
                elif f2 == 5:
                    name, arr = _tensor(v2)
                    g.initializers[name] = arr
                elif f2 == 11:
                    g.inputs.append(_value_info_name(v2))
                elif f2 == 12:
                    g.outputs.append(_value_info_name(v2))
    g.inputs = [i for i in g.inputs if i not in g.initializers]
    return g


# ---------------------------------------------------------------- graph -> architecture + state_dict

class UnsupportedGraph(RuntimeError):
    pass


def _ints(node: Node, key: str, default):
    v = node.attrs.get(key)
    return list(default) if v is None else [int(i) for i in v]


def plain_conv_unet_from_onnx(path: str):
    """Returns (arch_kwargs, state_dict of numpy arrays, info) for a PlainConvUNet-shaped graph.
    info: {'input_channels', 'num_heads', 'input_name', 'output_name', 'opset', 'negative_slope'}."""
    g = read_onnx(path)
    if len(g.inputs) != 1 or len(g.outputs) != 1:
        raise UnsupportedGraph(f'expected one graph input and one output, found {g.inputs} / {g.outputs}')
    consts = dict(g.initializers)
    for n in g.nodes:                       # Constant nodes (unfolded scalars) count as initializers
        if n.op_type == 'Constant' and 'value' in n.attrs:
            consts[n.outputs[0]] = n.attrs['value']
    # conv blocks in execution order
    blocks = []          # dicts: kind 'conv' | 'tconv', node, w, b, gamma, beta, eps, slope
    by_out = {}
    for n in g.nodes:
        for o in n.outputs:
            by_out[o] = n
    slope_seen = set()
    i, nodes = 0, [n for n in g.nodes if n.op_type not in ('Constant', 'Identity', 'Cast')]
    while i < len(nodes):
        n = nodes[i]
        if n.op_type in ('Conv', 'ConvTranspose'):
            if n.inputs[1] not in consts:
                raise UnsupportedGraph(f'{n.op_type} weight {n.inputs[1]!r} is not an initializer')
            blk = {'kind': 'conv' if n.op_type == 'Conv' else 'tconv', 'node': n, 'w': consts[n.inputs[1]],
                   'b': consts.get(n.inputs[2]) if len(n.inputs) > 2 and n.inputs[2] else None,
                   'gamma': None, 'beta': None, 'eps': 1e-5, 'slope': None}
            if _ints(n, 'dilations', [1, 1, 1]) != [1, 1, 1] or int(n.attrs.get('group', 1)) != 1:
                raise UnsupportedGraph(f'{n.name}: dilation / groups are not supported')
            j = i + 1
            if j < len(nodes) and nodes[j].op_type == 'InstanceNormalization' and nodes[j].inputs[0] == n.outputs[0]:
                inn = nodes[j]
                blk['gamma'], blk['beta'] = consts[inn.inputs[1]], consts[inn.inputs[2]]
                blk['eps'] = float(inn.attrs.get('epsilon', 1e-5))
                j += 1
                if j < len(nodes) and nodes[j].op_type in ('LeakyRelu', 'Relu'):
                    blk['slope'] = float(nodes[j].attrs.get('alpha', 0.01)) if nodes[j].op_type == 'LeakyRelu' else 0.0
                    slope_seen.add(blk['slope'])
                    j += 1
            blocks.append(blk)
            i = j
        elif n.op_type in ('Concat',):
            i += 1
        else:
            raise UnsupportedGraph(f'operator {n.op_type} ({n.name}) is not part of a PlainConvUNet graph; ResidualEncoderUNet '
                                   f'graphs must be imported from their checkpoints')
    if len(slope_seen) > 1:
        raise UnsupportedGraph(f'more than one activation slope in the graph: {sorted(slope_seen)}')
    if not blocks or blocks[-1]['gamma'] is not None or blocks[-1]['kind'] != 'conv':
        raise UnsupportedGraph('the graph does not end in a 1x1x1 segmentation head')
    head = blocks.pop()
    first_t = next((k for k, b in enumerate(blocks) if b['kind'] == 'tconv'), len(blocks))
    enc, dec = blocks[:first_t], blocks[first_t:]
    if any(b['kind'] != 'conv' or b['gamma'] is None for b in enc):
        raise UnsupportedGraph('encoder convolutions must each be followed by InstanceNormalization')
    # encoder stages: a new stage starts at every strided conv (and at the first conv)
    stages: List[List[dict]] = []
    for b in enc:
        strides = _ints(b['node'], 'strides', [1, 1, 1])
        if not stages or any(s != 1 for s in strides):
            stages.append([])
        stages[-1].append(b)
    n_stages = len(stages)
    levels: List[Tuple[dict, List[dict]]] = []
    for b in dec:
        if b['kind'] == 'tconv':
            levels.append((b, []))
        else:
            levels[-1][1].append(b)
    if len(levels) != n_stages - 1:
        raise UnsupportedGraph(f'{n_stages} encoder stages but {len(levels)} decoder levels')
    conv_bias = enc[0]['b'] is not None
    kw = {
        'n_stages': n_stages,
        'features_per_stage': [int(st[-1]['w'].shape[0]) for st in stages],
        'conv_op': 'torch.nn.modules.conv.Conv3d',
        'kernel_sizes': [[int(k) for k in st[0]['w'].shape[2:]] for st in stages],
        'strides': [_ints(st[0]['node'], 'strides', [1, 1, 1]) for st in stages],
        'n_conv_per_stage': [len(st) for st in stages],
        'n_conv_per_stage_decoder': [len(cv) for _, cv in levels],
        'conv_bias': conv_bias,
        'norm_op': 'torch.nn.modules.instancenorm.InstanceNorm3d',
        'norm_op_kwargs': {'eps': float(enc[0]['eps']), 'affine': True},
        'dropout_op': None, 'dropout_op_kwargs': None,
        'nonlin': 'torch.nn.LeakyReLU' if (not slope_seen or next(iter(slope_seen)) != 0.0) else 'torch.nn.ReLU',
        'nonlin_kwargs': {'inplace': True, **({'negative_slope': next(iter(slope_seen))} if slope_seen and
                                              next(iter(slope_seen)) not in (0.0,) else {})},
    }
    if any(len(k) != 3 for k in kw['kernel_sizes']):
        raise UnsupportedGraph('only 3-D convolutions are on the path')
    sd: Dict[str, np.ndarray] = {}

    def put_conv(prefix, b):
        sd[prefix + '.conv.weight'] = np.ascontiguousarray(b['w'], dtype=np.float32)
        if b['b'] is not None:
            sd[prefix + '.conv.bias'] = np.ascontiguousarray(b['b'], dtype=np.float32)
        sd[prefix + '.norm.weight'] = np.ascontiguousarray(b['gamma'], dtype=np.float32)
        sd[prefix + '.norm.bias'] = np.ascontiguousarray(b['beta'], dtype=np.float32)

    for s, st in enumerate(stages):
        for j, b in enumerate(st):
            pads = _ints(b['node'], 'pads', [0] * 6)
            k = [int(x) for x in b['w'].shape[2:]]
            if pads[:3] != [(x - 1) // 2 for x in k] or pads[3:] != pads[:3]:
                raise UnsupportedGraph(f'{b["node"].name}: padding {pads} is not (k - 1) / 2')
            put_conv(f'encoder.stages.{s}.0.convs.{j}', b)
    for l, (t, convs) in enumerate(levels):
        st = _ints(t['node'], 'strides', [1, 1, 1])
        if [int(x) for x in t['w'].shape[2:]] != st:
            raise UnsupportedGraph(f'{t["node"].name}: ConvTranspose kernel != stride')
        sd[f'decoder.transpconvs.{l}.weight'] = np.ascontiguousarray(t['w'], dtype=np.float32)
        if t['b'] is not None:
            sd[f'decoder.transpconvs.{l}.bias'] = np.ascontiguousarray(t['b'], dtype=np.float32)
        for j, b in enumerate(convs):
            if b['gamma'] is None:
                raise UnsupportedGraph('decoder convolutions must each be followed by InstanceNormalization')
            put_conv(f'decoder.stages.{l}.convs.{j}', b)
    sd[f'decoder.seg_layers.{n_stages - 2}.weight'] = np.ascontiguousarray(head['w'], dtype=np.float32)
    sd[f'decoder.seg_layers.{n_stages - 2}.bias'] = np.ascontiguousarray(
        head['b'] if head['b'] is not None else np.zeros(head['w'].shape[0]), dtype=np.float32)
    info = {'input_channels': int(stages[0][0]['w'].shape[1]), 'num_heads': int(head['w'].shape[0]),
            'input_name': g.inputs[0], 'output_name': g.outputs[0], 'opset': g.opset,
            'negative_slope': next(iter(slope_seen)) if slope_seen else 0.01}
    return kw, sd, info
