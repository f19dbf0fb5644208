"""Shared builders for the GPU network/predictor tests (oracle side lives here, not in the package)."""
import numpy as np
import torch

from fast_nnunet_b200 import model_folder as M
from oracle import networks as N

SMALL_PLAIN = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(32, 32, 32),
                   kw=M.plain_arch_kwargs([8, 16, 32], [[3, 3, 3]] * 3, [[1, 1, 1], [2, 2, 2], [2, 2, 2]]))
SMALL_PLAIN16 = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(32, 32, 32),
                     kw=M.plain_arch_kwargs([16, 32, 64], [[3, 3, 3]] * 3, [[1, 1, 1], [2, 2, 2], [2, 2, 2]]))
ANISO_PLAIN = dict(cls=M.PLAIN, in_ch=2, heads=5, patch=(20, 24, 32),
                   kw=M.plain_arch_kwargs([16, 32, 48, 64], [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                                          [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 1, 1]], [2, 1, 2, 2], [1, 2, 1]))
SMALL_RESENC = dict(cls=M.RESENC, in_ch=4, heads=4, patch=(32, 32, 32),
                    kw=M.resenc_arch_kwargs([16, 32, 32], [[3, 3, 3]] * 3, [[1, 1, 1], [2, 2, 2], [2, 2, 2]], [1, 2, 2]))
# shapes that exercise the row-streaming (ky-folded) tcgen05 kernel: W in [64, 128], Cout 16 / 32
ROWS_W128 = dict(cls=M.PLAIN, in_ch=1, heads=3, patch=(8, 16, 128),
                 kw=M.plain_arch_kwargs([16, 32], [[3, 3, 3]] * 2, [[1, 1, 1], [2, 2, 2]]))
ROWS_W96 = dict(cls=M.PLAIN, in_ch=2, heads=2, patch=(12, 20, 96),
                kw=M.plain_arch_kwargs([16, 32], [[1, 3, 3], [3, 3, 3]], [[1, 1, 1], [1, 2, 2]]))
ROWS_W64_C32 = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(8, 24, 64),
                    kw=M.plain_arch_kwargs([32, 64], [[3, 3, 3]] * 2, [[1, 1, 1], [2, 2, 2]]))
STUDENT = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(128, 128, 128),
               kw=M.plain_arch_kwargs([16, 32, 64, 128, 160, 160], [[3, 3, 3]] * 6, [[1, 1, 1]] + [[2, 2, 2]] * 5))


# the other BASELINE.json configurations, at their real patch sizes
TEACHER = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(128, 128, 128),
               kw=M.plain_arch_kwargs([32, 64, 128, 256, 320, 320], [[3, 3, 3]] * 6, [[1, 1, 1]] + [[2, 2, 2]] * 5))
RESENC_M_STUDENT = dict(cls=M.RESENC, in_ch=4, heads=4, patch=(128, 128, 128),
                        kw=M.resenc_arch_kwargs([16, 32, 64, 128, 160, 160], [[3, 3, 3]] * 6,
                                                [[1, 1, 1]] + [[2, 2, 2]] * 5, [1, 3, 4, 6, 6, 6]))
# topology returned by the reference's get_pool_and_conv_props for spacing (2.0, 0.977, 0.977), patch (160, 96, 96)
# (tests/golden/sliding_window_golden.json 'topology'[0]); 61 labels as in engine/config/fast_nnunet_bone_turbo.ini
BONE_TURBO = dict(cls=M.PLAIN, in_ch=1, heads=61, patch=(160, 96, 96),
                  kw=M.plain_arch_kwargs([16, 32, 64, 128, 160, 160],
                                         [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                                         [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 1, 1]]))


def make(spec, seed=1234, randomize_affine=True):
    sd = M.synthesize_state_dict(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], seed=seed,
                                 randomize_affine=randomize_affine)
    net = N.build_from_arch(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], allow_init=False)
    net.load_state_dict(sd, strict=True)
    net.eval()
    return sd, net


def ct_like_volume(shape, channels=1, seed=0):
    """clip(N(-350, 450^2), -1100, 1207) then CT normalisation with the bone_turbo ini stats,
    plus a smooth low-frequency component so that neighbouring voxels correlate like an image."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn((channels, *shape), generator=g) * 450.0 - 350.0
    low = torch.nn.functional.interpolate(torch.randn((1, channels, *[max(s // 16, 2) for s in shape]), generator=g),
                                          size=shape, mode='trilinear', align_corners=False)[0] * 400.0
    v = (v * 0.5 + low).clamp_(-1100, 1207)
    return ((v - (-350.0)) / 450.0).contiguous()
