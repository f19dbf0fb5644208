"""Oracle restatement of the network families on the hot path (torch.nn, runs on
CPU and CUDA).  Test infrastructure only (see oracle/__init__.py).

The arithmetic lives in the third-party package ``dynamic_network_architectures``
(MIC-DKFZ, PyPI ``dynamic-network-architectures``; NOT vendored under
/root/reference and not pinned by it: distillation/setup.py:7-10).  This file
restates its published module structure so that ``state_dict`` keys are
interchangeable with real nnU-Net checkpoints:

  encoder.stages.{s}.0.convs.{j}.{conv,norm}.{weight,bias}   (+ all_modules.{0,1} aliases)
  decoder.encoder.*                                          (alias of encoder.*)
  decoder.transpconvs.{l}.{weight,bias}, decoder.stages.{l}.convs.{j}.*, decoder.seg_layers.{l}.*
  ResEnc: encoder.stem.convs.0.*, encoder.stages.{s}.blocks.{b}.{conv1,conv2}.*, ...blocks.{b}.skip.{i}.*

Call sites restated: utilities/get_network_from_plans.py:9-43 (constructor
kwargs), nnUNetDistillationTrainer.py:74-177 (LiteNNUNetStudent),
:180-274 (LiteResEncStudent), :678 (feature reduction rule),
utilities/network_initialization.py:4-12 (He init).  PARITY UNPINNED for the
forward arithmetic (no reference golden vectors exist for it).
"""
from __future__ import annotations

import torch
from torch import nn


def _tup(v, n=3):
    if isinstance(v, int):
        return (v,) * n
    return tuple(int(i) for i in v)


class ConvNormAct(nn.Module):
    """conv -> InstanceNorm(affine) -> LeakyReLU; `act=False` drops the non-linearity."""

    def __init__(self, cin, cout, kernel, stride, bias, eps=1e-5, act=True, slope=0.01):
        super().__init__()
        kernel, stride = _tup(kernel), _tup(stride)
        self.conv = nn.Conv3d(cin, cout, kernel, stride, padding=[(k - 1) // 2 for k in kernel], bias=bias)
        self.norm = nn.InstanceNorm3d(cout, eps=eps, affine=True)
        mods = [self.conv, self.norm]
        if act:
            self.nonlin = nn.LeakyReLU(slope, inplace=True)
            mods.append(self.nonlin)
        self.all_modules = nn.Sequential(*mods)

    def forward(self, x):
        return self.all_modules(x)


class ConvStack(nn.Module):
    def __init__(self, n, cin, cout, kernel, first_stride, bias, eps):
        super().__init__()
        chans = cout if isinstance(cout, (list, tuple)) else [cout] * n
        blocks = [ConvNormAct(cin, chans[0], kernel, first_stride, bias, eps)]
        for i in range(1, n):
            blocks.append(ConvNormAct(chans[i - 1], chans[i], kernel, 1, bias, eps))
        self.convs = nn.Sequential(*blocks)

    def forward(self, x):
        return self.convs(x)


class PlainEncoder(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, kernel_sizes, strides, n_conv_per_stage,
                 conv_bias, eps):
        super().__init__()
        stages = []
        cin = input_channels
        for s in range(n_stages):
            stages.append(nn.Sequential(ConvStack(n_conv_per_stage[s], cin, features_per_stage[s], kernel_sizes[s],
                                                  strides[s], conv_bias, eps)))
            cin = features_per_stage[s]
        self.stages = nn.Sequential(*stages)
        self.output_channels = list(features_per_stage)
        self.strides = [_tup(s) for s in strides]
        self.kernel_sizes = [_tup(k) for k in kernel_sizes]
        self.conv_bias = conv_bias
        self.eps = eps

    def forward(self, x):
        skips = []
        for st in self.stages:
            x = st(x)
            skips.append(x)
        return skips


class ResidualBlock(nn.Module):
    """BasicBlockD: lrelu(IN(conv2(lrelu(IN(conv1(x))))) + skip(x))."""

    def __init__(self, cin, cout, kernel, stride, bias, eps):
        super().__init__()
        stride = _tup(stride)
        self.conv1 = ConvNormAct(cin, cout, kernel, stride, bias, eps)
        self.conv2 = ConvNormAct(cout, cout, kernel, 1, bias, eps, act=False)
        self.nonlin2 = nn.LeakyReLU(0.01, inplace=True)
        has_stride = any(s != 1 for s in stride)
        proj = cin != cout
        if has_stride or proj:
            ops = []
            if has_stride:
                ops.append(nn.AvgPool3d(stride, stride))
            if proj:
                ops.append(ConvNormAct(cin, cout, 1, 1, False, eps, act=False))
            self.skip = nn.Sequential(*ops)
        else:
            self.skip = lambda x: x

    def forward(self, x):
        r = self.skip(x)
        out = self.conv2(self.conv1(x))
        out = out + r
        return self.nonlin2(out)


class ResidualStack(nn.Module):
    def __init__(self, n, cin, cout, kernel, first_stride, bias, eps):
        super().__init__()
        blocks = [ResidualBlock(cin, cout, kernel, first_stride, bias, eps)]
        for _ in range(1, n):
            blocks.append(ResidualBlock(cout, cout, kernel, 1, bias, eps))
        self.blocks = nn.Sequential(*blocks)

    def forward(self, x):
        return self.blocks(x)


class ResidualEncoder(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, kernel_sizes, strides, n_blocks_per_stage,
                 conv_bias, eps):
        super().__init__()
        self.stem = ConvStack(1, input_channels, features_per_stage[0], kernel_sizes[0], 1, conv_bias, eps)
        cin = features_per_stage[0]
        stages = []
        for s in range(n_stages):
            stages.append(ResidualStack(n_blocks_per_stage[s], cin, features_per_stage[s], kernel_sizes[s],
                                        strides[s], conv_bias, eps))
            cin = features_per_stage[s]
        self.stages = nn.Sequential(*stages)
        self.output_channels = list(features_per_stage)
        self.strides = [_tup(s) for s in strides]
        self.kernel_sizes = [_tup(k) for k in kernel_sizes]
        self.conv_bias = conv_bias
        self.eps = eps

    def forward(self, x):
        x = self.stem(x)
        skips = []
        for st in self.stages:
            x = st(x)
            skips.append(x)
        return skips


class Decoder(nn.Module):
    """UNetDecoder: per level ConvTranspose(k=s) -> cat((up, skip), 1) -> conv stack; one 1x1x1
    seg layer per level, only the highest-resolution one is evaluated without deep supervision."""

    def __init__(self, encoder, num_classes, n_conv_per_stage, deep_supervision):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        n_enc = len(encoder.output_channels)
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * (n_enc - 1)
        assert len(n_conv_per_stage) == n_enc - 1
        stages, ups, segs = [], [], []
        for s in range(1, n_enc):
            below = encoder.output_channels[-s]
            skip = encoder.output_channels[-(s + 1)]
            st = encoder.strides[-s]
            ups.append(nn.ConvTranspose3d(below, skip, st, st, bias=encoder.conv_bias))
            stages.append(ConvStack(n_conv_per_stage[s - 1], 2 * skip, skip, encoder.kernel_sizes[-(s + 1)], 1,
                                    encoder.conv_bias, encoder.eps))
            segs.append(nn.Conv3d(skip, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(ups)
        self.seg_layers = nn.ModuleList(segs)

    def forward(self, skips):
        low = skips[-1]
        outs = []
        for s in range(len(self.stages)):
            x = self.transpconvs[s](low)
            x = torch.cat((x, skips[-(s + 2)]), 1)
            x = self.stages[s](x)
            if self.deep_supervision:
                outs.append(self.seg_layers[s](x))
            elif s == len(self.stages) - 1:
                outs.append(self.seg_layers[-1](x))
            low = x
        outs = outs[::-1]
        return outs if self.deep_supervision else outs[0]


def _eps_of(norm_op_kwargs):
    return 1e-5 if not norm_op_kwargs else float(norm_op_kwargs.get('eps', 1e-5))


def _he_init(module, slope=1e-2):
    """utilities/network_initialization.py:4-12."""
    if isinstance(module, (nn.Conv3d, nn.ConvTranspose3d)):
        nn.init.kaiming_normal_(module.weight, a=slope)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)


class PlainConvUNet(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, conv_op=None, kernel_sizes=3, strides=1,
                 n_conv_per_stage=2, num_classes=2, n_conv_per_stage_decoder=2, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None,
                 deep_supervision=False, nonlin_first=False):
        super().__init__()
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * n_stages
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * n_stages
        if isinstance(strides, int):
            strides = [strides] * n_stages
        self.encoder = PlainEncoder(input_channels, n_stages, features_per_stage, kernel_sizes, strides,
                                    n_conv_per_stage, conv_bias, _eps_of(norm_op_kwargs))
        self.decoder = Decoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision)

    def forward(self, x):
        return self.decoder(self.encoder(x))

    @staticmethod
    def initialize(module):
        _he_init(module)


class ResidualEncoderUNet(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, conv_op=None, kernel_sizes=3, strides=1,
                 n_blocks_per_stage=2, num_classes=2, n_conv_per_stage_decoder=1, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None,
                 deep_supervision=False, block=None, bottleneck_channels=None, stem_channels=None):
        super().__init__()
        if isinstance(n_blocks_per_stage, int):
            n_blocks_per_stage = [n_blocks_per_stage] * n_stages
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * n_stages
        if isinstance(strides, int):
            strides = [strides] * n_stages
        self.encoder = ResidualEncoder(input_channels, n_stages, features_per_stage, kernel_sizes, strides,
                                       n_blocks_per_stage, conv_bias, _eps_of(norm_op_kwargs))
        self.decoder = Decoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision)

    def forward(self, x):
        return self.decoder(self.encoder(x))

    @staticmethod
    def initialize(module):
        _he_init(module)


class LiteNNUNetStudent(nn.Module):
    """nnUNetDistillationTrainer.py:74-177 — same blocks, attribute names encoder/decoder."""

    def __init__(self, input_channels, num_classes, n_stages, features_per_stage, kernel_sizes, strides,
                 n_conv_per_stage, n_conv_per_stage_decoder, conv_bias=True, norm_op_kwargs=None,
                 deep_supervision=True, **_):
        super().__init__()
        self.encoder = PlainEncoder(input_channels, n_stages, features_per_stage, kernel_sizes, strides,
                                    n_conv_per_stage, conv_bias, _eps_of(norm_op_kwargs))
        self.decoder = Decoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision)

    def forward(self, x):
        return self.decoder(self.encoder(x))


class LiteResEncStudent(nn.Module):
    """nnUNetDistillationTrainer.py:180-274 — wraps the network under `.network`."""

    def __init__(self, input_channels, num_classes, n_stages, features_per_stage, kernel_sizes, strides,
                 n_blocks_per_stage, n_conv_per_stage_decoder, conv_bias=True, norm_op_kwargs=None,
                 deep_supervision=True, **_):
        super().__init__()
        self.network = ResidualEncoderUNet(input_channels, n_stages, features_per_stage, None, kernel_sizes, strides,
                                           n_blocks_per_stage, num_classes, n_conv_per_stage_decoder, conv_bias,
                                           None, norm_op_kwargs, deep_supervision=deep_supervision)

    def forward(self, x):
        return self.network(x)

    @property
    def decoder(self):
        return self.network.decoder


def student_features(features_per_stage, reduction_factor):
    """nnUNetDistillationTrainer.py:678."""
    return [max(f // reduction_factor, 8) for f in features_per_stage]


def student_blocks(n_blocks_per_stage, features, lite_features, strategy):
    """nnUNetDistillationTrainer.py:688-708."""
    if strategy == 'reduce':
        return [max(n // 2, 1) for n in n_blocks_per_stage]
    if strategy == 'increase':
        return [min(n + 1, 8) for n in n_blocks_per_stage]
    if strategy == 'adaptive':
        ratios = [o / r for o, r in zip(features, lite_features)]
        return [min(n + max(0, int(ratio / 4)), 8) for n, ratio in zip(n_blocks_per_stage, ratios)]
    return list(n_blocks_per_stage)


def build_from_arch(network_class_name, arch_kwargs, input_channels, num_classes, deep_supervision=False,
                    allow_init=True):
    """get_network_from_plans.py:9-43 restated for the two families on the path."""
    kw = dict(arch_kwargs)
    for k in ('conv_op', 'norm_op', 'dropout_op', 'nonlin', 'dropout_op_kwargs', 'nonlin_kwargs'):
        kw.pop(k, None)
    name = network_class_name.split('.')[-1]
    if name == 'PlainConvUNet':
        net = PlainConvUNet(input_channels=input_channels, num_classes=num_classes,
                            deep_supervision=deep_supervision, **kw)
    elif name == 'ResidualEncoderUNet':
        net = ResidualEncoderUNet(input_channels=input_channels, num_classes=num_classes,
                                  deep_supervision=deep_supervision, **kw)
    else:
        raise ImportError(f'Network class {network_class_name} is not on the path')
    if allow_init:
        net.apply(net.initialize)
    return net
