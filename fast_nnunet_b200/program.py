"""Lowers a PlainConvUNet / ResidualEncoderUNet description (plans `arch_kwargs` + a checkpoint
`state_dict`) into the flat buffer/operator program libfnnu executes (include/fnnu.h).

Module structure and `state_dict` key layout follow dynamic_network_architectures as instantiated by
the reference at utilities/get_network_from_plans.py:9-43 and, for distilled students,
training/nnUNetTrainer/variants/nnUNetDistillationTrainer.py:74-274,678-749.

Fusion decisions encoded here:
  * every Conv->InstanceNorm->LeakyReLU writes its RAW output; the norm/activation is a per-channel
    pending transform applied by whichever operator loads the buffer next;
  * `torch.cat((up, skip), 1)` is never executed: the transposed conv writes channels [0, C) and the
    encoder stage's last conv writes channels [C, 2C) of one pre-allocated buffer;
  * only the highest-resolution seg layer is evaluated (deep supervision off at inference);
  * the bias of a convolution that feeds an InstanceNorm is dropped (IN(x + b) = IN(x)).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib

LRELU_SLOPE = 0.01   # torch.nn.LeakyReLU default; plans carry nonlin_kwargs {'inplace': True} only


class UnsupportedArchitecture(RuntimeError):
    pass


def check_arch(arch_kwargs: dict) -> float:
    """The engine fuses InstanceNorm(affine) + LeakyReLU/ReLU; the reference builds whatever the plans say
    (utilities/get_network_from_plans.py:17-38).  Anything else is refused here, loudly, instead of producing
    silently different logits.  Dropout is the identity in eval mode, the only mode on this path.
    Returns the activation's negative slope."""
    def name(v):
        return v if isinstance(v, str) else getattr(v, '__name__', str(v))
    norm = name(arch_kwargs.get('norm_op') or 'InstanceNorm3d')
    if not norm.endswith('InstanceNorm3d'):
        raise UnsupportedArchitecture(f'unsupported architecture: norm_op={norm!r} (the B200 engine fuses InstanceNorm3d only)')
    if not (arch_kwargs.get('norm_op_kwargs') or {}).get('affine', True):
        raise UnsupportedArchitecture('unsupported architecture: norm_op_kwargs.affine=False')
    nonlin = name(arch_kwargs.get('nonlin') or 'LeakyReLU')
    kw = arch_kwargs.get('nonlin_kwargs') or {}
    if nonlin.endswith('LeakyReLU'):
        return float(kw.get('negative_slope', LRELU_SLOPE))
    if nonlin.split('.')[-1] == 'ReLU':
        return 0.0
    raise UnsupportedArchitecture(f'unsupported architecture: nonlin={nonlin!r} (LeakyReLU and ReLU are supported)')


@dataclass
class Op:
    op: int
    src: int
    dst: int
    cin: int
    cout: int
    src_coff: int = 0
    dst_coff: int = 0
    src2: int = -1
    src2_coff: int = 0
    kernel: Sequence[int] = (1, 1, 1)
    stride: Sequence[int] = (1, 1, 1)
    weight: Optional[np.ndarray] = None
    bias: Optional[np.ndarray] = None
    gamma: Optional[np.ndarray] = None
    beta: Optional[np.ndarray] = None
    eps: float = 1e-5
    act_slope: float = 1.0
    name: str = ''

    @property
    def has_norm(self):
        return self.gamma is not None

    def flops(self, in_dims, out_dims) -> float:
        if self.op == _lib.OP_CONV:
            return 2.0 * self.cin * self.cout * float(np.prod(self.kernel)) * float(np.prod(out_dims))
        if self.op == _lib.OP_TCONV:
            return 2.0 * self.cin * self.cout * float(np.prod(self.stride)) * float(np.prod(in_dims))
        return 0.0


@dataclass
class Program:
    buffers: List[tuple] = field(default_factory=list)     # (dims(3), channels)
    ops: List[Op] = field(default_factory=list)
    input_buffer: int = 0
    output_buffer: int = 0
    in_channels: int = 1
    num_heads: int = 2
    patch_size: Sequence[int] = (128, 128, 128)

    def add_buffer(self, dims, channels) -> int:
        self.buffers.append((tuple(int(d) for d in dims), int(channels)))
        return len(self.buffers) - 1

    def total_flops(self) -> float:
        return sum(o.flops(self.buffers[o.src][0], self.buffers[o.dst][0]) for o in self.ops)

    def activation_elements(self) -> int:
        return sum(int(np.prod(d)) * c for d, c in self.buffers)


def _np(t) -> np.ndarray:
    if hasattr(t, 'detach'):
        t = t.detach().float().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def clean_state_dict(sd: Dict) -> Dict:
    """Strips wrapper prefixes (DDP `module.`, torch.compile `_orig_mod.`, LiteResEncStudent `network.`;
    nnUNetDistillationTrainer.py:1038-1048, :248)."""
    out = {}
    for k, v in sd.items():
        changed = True
        while changed:
            changed = False
            for p in ('module.', '_orig_mod.', 'network.'):
                if k.startswith(p):
                    k = k[len(p):]
                    changed = True
        out[k] = v
    return out


def padded_heads(num_heads: int) -> int:
    for c in (2, 4, 8):
        if num_heads <= c:
            return c
    return (num_heads + 15) // 16 * 16


def _t3(v):
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(i) for i in v)
    assert len(v) == 3, 'only 3-D networks are on the path'
    return v


def _conv_out(dims, kernel, stride):
    return tuple((d + 2 * ((k - 1) // 2) - k) // s + 1 for d, k, s in zip(dims, kernel, stride))


class _Builder:
    def __init__(self, sd, patch_size, in_channels, num_heads, conv_bias, eps, slope):
        self.slope = slope
        self.sd = sd
        self.p = Program(in_channels=in_channels, num_heads=num_heads, patch_size=tuple(patch_size))
        self.conv_bias = conv_bias
        self.eps = eps

    def get(self, key):
        if key not in self.sd:
            raise KeyError(f'checkpoint is missing parameter {key!r}')
        return _np(self.sd[key])

    def conv_norm(self, prefix, src, src_coff, cin, cout, kernel, stride, dst, dst_coff, slope, bias=None, name=''):
        """ConvDropoutNormReLU under `prefix` ('....convs.0' etc.): keys conv.weight, conv.bias, norm.*"""
        # A bias in front of InstanceNorm is a mathematical no-op (IN(x + b) = IN(x): the mean absorbs it, the variance
        # does not see it), so it is not lowered: one add per output less in every epilogue, and the fp16 raw
        # activations keep their precision when |b| is large against the channel's spread.
        use_bias = False
        w = self.get(prefix + '.conv.weight')
        assert w.shape == (cout, cin, *kernel), f'{prefix}: weight {w.shape} != {(cout, cin, *kernel)}'
        self.p.ops.append(Op(op=_lib.OP_CONV, src=src, src_coff=src_coff, dst=dst, dst_coff=dst_coff, cin=cin,
                             cout=cout, kernel=kernel, stride=stride, weight=w,
                             bias=self.get(prefix + '.conv.bias') if use_bias else None,
                             gamma=self.get(prefix + '.norm.weight'), beta=self.get(prefix + '.norm.bias'),
                             eps=self.eps, act_slope=slope, name=name or prefix))


def _decoder(b: _Builder, feats, kernels, strides, stage_dims, skip_bufs, bottleneck_buf, n_conv_dec, num_heads):
    """UNetDecoder: level l (0 = lowest resolution) upsamples from stage n-1-l to stage n-2-l."""
    p = b.p
    n_stages = len(feats)
    low = bottleneck_buf
    low_c = feats[-1]
    for l in range(n_stages - 1):
        s = n_stages - 2 - l
        c = feats[s]
        st = strides[s + 1]
        cat = skip_bufs[s]                                   # [0,c) up ; [c,2c) skip (already written)
        w = b.get(f'decoder.transpconvs.{l}.weight')
        assert w.shape == (low_c, c, *st), f'transpconv {l}: weight {w.shape}'
        p.ops.append(Op(op=_lib.OP_TCONV, src=low, dst=cat, dst_coff=0, cin=low_c, cout=c, kernel=st, stride=st,
                        weight=w, bias=b.get(f'decoder.transpconvs.{l}.bias') if b.conv_bias else None,
                        name=f'decoder.transpconvs.{l}'))
        src, cin = cat, 2 * c
        for j in range(n_conv_dec[l]):
            dst = p.add_buffer(stage_dims[s], c)
            b.conv_norm(f'decoder.stages.{l}.convs.{j}', src, 0, cin, c, kernels[s], (1, 1, 1), dst, 0, b.slope)
            src, cin = dst, c
        low, low_c = src, c
    # highest-resolution seg layer only (deep supervision off: UNetDecoder.forward uses seg_layers[-1])
    last = n_stages - 2
    w = b.get(f'decoder.seg_layers.{last}.weight')
    assert w.shape[:2] == (num_heads, feats[0]), f'seg layer weight {w.shape}'
    # heads are padded per voxel to 2 / 4 / 8 / a multiple of 16 fp16 values so that the accumulate kernel reads a
    # voxel's heads with one aligned 4 / 8 / 16 / 32-byte load (padding channels are never read back)
    out = p.add_buffer(stage_dims[0], padded_heads(num_heads))
    p.ops.append(Op(op=_lib.OP_CONV, src=low, dst=out, cin=feats[0], cout=num_heads, kernel=(1, 1, 1),
                    stride=(1, 1, 1), weight=w, bias=b.get(f'decoder.seg_layers.{last}.bias'),
                    name=f'decoder.seg_layers.{last}'))
    p.output_buffer = out


def build_plain_conv_unet(state_dict, arch_kwargs, in_channels, num_heads, patch_size) -> Program:
    sd = clean_state_dict(state_dict)
    n_stages = int(arch_kwargs['n_stages'])
    feats = [int(f) for f in arch_kwargs['features_per_stage']]
    kernels = [_t3(k) for k in arch_kwargs['kernel_sizes']]
    strides = [_t3(s) for s in arch_kwargs['strides']]
    n_conv = arch_kwargs['n_conv_per_stage']
    n_conv = [n_conv] * n_stages if isinstance(n_conv, int) else list(n_conv)
    n_dec = arch_kwargs['n_conv_per_stage_decoder']
    n_dec = [n_dec] * (n_stages - 1) if isinstance(n_dec, int) else list(n_dec)
    eps = float((arch_kwargs.get('norm_op_kwargs') or {}).get('eps', 1e-5))
    b = _Builder(sd, patch_size, in_channels, num_heads, bool(arch_kwargs.get('conv_bias', False)), eps,
                 check_arch(arch_kwargs))
    p = b.p
    p.input_buffer = p.add_buffer(patch_size, in_channels)
    dims = tuple(patch_size)
    stage_dims, skip_bufs = [], []
    src, src_coff, cin = p.input_buffer, 0, in_channels
    for s in range(n_stages):
        out_dims = _conv_out(dims, kernels[s], strides[s])
        stage_dims.append(out_dims)
        c = feats[s]
        is_last_stage = s == n_stages - 1
        # the stage output doubles as the skip: it lives in channels [c, 2c) of the concat buffer
        cat = None if is_last_stage else p.add_buffer(out_dims, 2 * c)
        for j in range(n_conv[s]):
            last_conv = j == n_conv[s] - 1
            if last_conv and not is_last_stage:
                dst, dst_coff = cat, c
            else:
                dst, dst_coff = p.add_buffer(out_dims, c), 0
            b.conv_norm(f'encoder.stages.{s}.0.convs.{j}', src, src_coff, cin, c, kernels[s],
                        strides[s] if j == 0 else (1, 1, 1), dst, dst_coff, b.slope)
            src, src_coff, cin = dst, dst_coff, c
        skip_bufs.append(cat)
        dims = out_dims
    assert src_coff == 0
    _decoder(b, feats, kernels, strides, stage_dims, skip_bufs, src, n_dec, num_heads)
    return p


def build_residual_encoder_unet(state_dict, arch_kwargs, in_channels, num_heads, patch_size) -> Program:
    sd = clean_state_dict(state_dict)
    n_stages = int(arch_kwargs['n_stages'])
    feats = [int(f) for f in arch_kwargs['features_per_stage']]
    ks = arch_kwargs['kernel_sizes']
    kernels = [_t3(ks)] * n_stages if isinstance(ks, int) else [_t3(k) for k in ks]
    strides = [_t3(s) for s in arch_kwargs['strides']]
    n_blocks = arch_kwargs['n_blocks_per_stage']
    n_blocks = [n_blocks] * n_stages if isinstance(n_blocks, int) else list(n_blocks)
    n_dec = arch_kwargs['n_conv_per_stage_decoder']
    n_dec = [n_dec] * (n_stages - 1) if isinstance(n_dec, int) else list(n_dec)
    eps = float((arch_kwargs.get('norm_op_kwargs') or {}).get('eps', 1e-5))
    b = _Builder(sd, patch_size, in_channels, num_heads, bool(arch_kwargs.get('conv_bias', False)), eps,
                 check_arch(arch_kwargs))
    p = b.p
    p.input_buffer = p.add_buffer(patch_size, in_channels)
    dims = tuple(patch_size)
    stem = p.add_buffer(dims, feats[0])
    b.conv_norm('encoder.stem.convs.0', p.input_buffer, 0, in_channels, feats[0], kernels[0], (1, 1, 1), stem, 0,
                b.slope)
    x, x_coff, cin = stem, 0, feats[0]
    stage_dims, skip_bufs = [], []
    for s in range(n_stages):
        c = feats[s]
        out_dims = _conv_out(dims, kernels[s], strides[s])
        stage_dims.append(out_dims)
        is_last_stage = s == n_stages - 1
        cat = None if is_last_stage else p.add_buffer(out_dims, 2 * c)
        for blk in range(n_blocks[s]):
            pre = f'encoder.stages.{s}.blocks.{blk}'
            stride = strides[s] if blk == 0 else (1, 1, 1)
            in_dims = dims if blk == 0 else out_dims
            has_stride = any(i != 1 for i in stride)
            proj = cin != c
            t1 = p.add_buffer(out_dims, c)
            b.conv_norm(pre + '.conv1', x, x_coff, cin, c, kernels[s], stride, t1, 0, b.slope)
            t2 = p.add_buffer(out_dims, c)
            b.conv_norm(pre + '.conv2', t1, 0, c, c, kernels[s], (1, 1, 1), t2, 0, 1.0)
            r, r_coff = x, x_coff
            idx = 0
            if has_stride:
                pooled = p.add_buffer(out_dims, cin)
                p.ops.append(Op(op=_lib.OP_AVGPOOL, src=r, src_coff=r_coff, dst=pooled, cin=cin, cout=cin,
                                stride=stride, kernel=stride, name=pre + '.skip.0'))
                r, r_coff = pooled, 0
                idx = 1
            if proj:
                pr = p.add_buffer(out_dims, c)
                b.conv_norm(pre + f'.skip.{idx}', r, r_coff, cin, c, (1, 1, 1), (1, 1, 1), pr, 0, 1.0, bias=False)
                r, r_coff = pr, 0
            last_block = blk == n_blocks[s] - 1
            if last_block and not is_last_stage:
                dst, dst_coff = cat, c
            else:
                dst, dst_coff = p.add_buffer(out_dims, c), 0
            p.ops.append(Op(op=_lib.OP_ADD_ACT, src=t2, src2=r, src2_coff=r_coff, dst=dst, dst_coff=dst_coff, cin=c,
                            cout=c, act_slope=b.slope, name=pre + '.add'))
            x, x_coff, cin = dst, dst_coff, c
            del in_dims
        skip_bufs.append(cat)
        dims = out_dims
    assert x_coff == 0
    _decoder(b, feats, kernels, strides, stage_dims, skip_bufs, x, n_dec, num_heads)
    return p


def build_program(network_class_name: str, state_dict, arch_kwargs, in_channels, num_heads, patch_size) -> Program:
    name = network_class_name.split('.')[-1]
    if name in ('PlainConvUNet', 'LiteNNUNetStudent'):
        return build_plain_conv_unet(state_dict, arch_kwargs, in_channels, num_heads, patch_size)
    if name in ('ResidualEncoderUNet', 'LiteResEncStudent'):
        return build_residual_encoder_unet(state_dict, arch_kwargs, in_channels, num_heads, patch_size)
    raise RuntimeError(f'network class {network_class_name!r} is not supported by the B200 inference engine '
                       f'(PlainConvUNet and ResidualEncoderUNet are)')
