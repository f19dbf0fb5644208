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
# z-pair row-streaming kernel (conv_umma_zrows.cu): 3x3x3, Cout 16, even depth; W below 128, odd row counts, y segments
ZROWS_W96 = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(6, 40, 96),
                 kw=M.plain_arch_kwargs([16, 32], [[3, 3, 3]] * 2, [[1, 1, 1], [2, 2, 2]]))
ZROWS_W72_H9 = dict(cls=M.PLAIN, in_ch=1, heads=3, patch=(4, 18, 72),
                    kw=M.plain_arch_kwargs([16, 32], [[3, 3, 3]] * 2, [[1, 1, 1], [2, 2, 2]]))
# the network behind tests/golden/tiny_plain_unet.onnx (anisotropic first stage, 3 heads, 2 / 1 convs per stage)
TINY_ONNX = dict(cls=M.PLAIN, in_ch=1, heads=3, patch=(16, 16, 16),
                 kw=M.plain_arch_kwargs([8, 16, 16], [[1, 3, 3], [3, 3, 3], [3, 3, 3]], [[1, 1, 1], [1, 2, 2], [2, 2, 2]],
                                        [2, 1, 2], [1, 2]))
# odd depth: one output plane per pass (ZC = 1) with the full 3x3x3 kernel; anisotropic second stage keeps D = 5
ZROWS_ODD_D = dict(cls=M.PLAIN, in_ch=1, heads=2, patch=(5, 16, 128),
                   kw=M.plain_arch_kwargs([16, 32], [[3, 3, 3]] * 2, [[1, 1, 1], [1, 2, 2]]))
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


def phantom_codes(channels, n_used):
    """Intensity code of each used class, (n_used, channels), already-normalised units: indicator channels when there
    are enough of them (MR-like: every structure bright in its own sequence), points on a circle for 2 channels,
    evenly spaced levels for 1 channel (CT-like: air / fat / soft tissue / bone ...)."""
    if channels >= n_used:
        c = torch.full((n_used, channels), -0.8)
        c[torch.arange(n_used), torch.arange(n_used)] = 1.6
        return c
    if channels == 1:
        return torch.linspace(-1.6, 1.6, n_used)[:, None]
    ang = torch.arange(n_used, dtype=torch.float32) * (2 * np.pi / n_used)
    c = torch.zeros((n_used, channels))
    c[:, 0], c[:, 1] = 1.6 * torch.cos(ang), 1.6 * torch.sin(ang)
    return c


def phantom_volume(shape, channels=1, n_classes=2, seed=0, cell=12, noise=0.08, max_used=None):
    """Piecewise-constant "organ" phantom: a random class map on a coarse grid (cells of ~`cell` voxels, sharp
    faces like organ boundaries), one intensity code per class (phantom_codes) plus white noise.  At most `max_used`
    of the `n_classes` labels occur (labels spread over the label range), as in a field of view that holds a few of
    the structures a many-class model knows.  Returns (volume float32 (C, *shape), class map int64 (*shape)).
    Real CT/MR volumes are piecewise smooth; a trained network's logits are saturated away from the faces, which is
    the regime north_star's label bar (agreement >= 99.9 %, Dice >= 0.999) is stated for."""
    g = torch.Generator().manual_seed(seed)
    if max_used is None:
        max_used = 3 if channels == 1 else 6      # what a few hundred training steps separate with real margins
    n_used = min(n_classes, max_used)
    used = torch.linspace(0, n_classes - 1, n_used).round().long()      # seed-independent: every phantom uses the same labels
    coarse = [max(2, -(-s // cell)) for s in shape]
    lab = torch.randint(0, n_used, coarse, generator=g)
    flat = lab.view(-1)
    flat[torch.randperm(flat.numel(), generator=g)[:n_used]] = torch.arange(n_used)   # every used class is present
    lab = torch.nn.functional.interpolate(lab[None, None].float(), size=shape, mode='nearest')[0, 0].long()
    codes = phantom_codes(channels, n_used)                              # (n_used, C)
    vol = codes.T[:, lab] + noise * torch.randn((channels, *shape), generator=g)
    return vol.contiguous().float(), used[lab]


@torch.no_grad()
def fit_seg_head(net, sd, patches, labels, gain=8.0, ridge=1e-3, device=None):
    """Replaces the He-init 1x1x1 segmentation head (whose class margins are ~0 everywhere) by the ridge-regression
    read-out of the ORACLE's penultimate features onto `gain` x one-hot(labels): the logits then carry real margins,
    as a trained model's do.  patches: (n, C, *patch) fp32; labels: (n, *patch) int64.  Updates both the oracle
    module and the state_dict (every alias key of the last seg layer)."""
    dec = net.decoder
    seg = dec.seg_layers[-1]
    feats = []
    h = seg.register_forward_pre_hook(lambda m, inp: feats.append(inp[0].detach()))
    dev = device or next(net.parameters()).device
    net.to(dev)
    for i in range(patches.shape[0]):
        net(patches[i:i + 1].to(dev))
    h.remove()
    F = torch.cat([f.permute(0, 2, 3, 4, 1).reshape(-1, f.shape[1]) for f in feats]).double()
    K = seg.out_channels
    y = labels.reshape(-1).to(F.device)
    T = torch.nn.functional.one_hot(y, K).double() * gain - gain / K
    A = torch.cat([F, torch.ones((F.shape[0], 1), dtype=F.dtype, device=F.device)], 1)
    G = A.T @ A
    G += ridge * F.shape[0] * torch.eye(G.shape[0], dtype=G.dtype, device=G.device)
    Wb = torch.linalg.solve(G, A.T @ T)                # (C + 1, K)
    W = Wb[:-1].T.float().cpu().reshape(K, -1, 1, 1, 1).contiguous()
    b = Wb[-1].float().cpu().contiguous()
    seg.weight.copy_(W.to(seg.weight.device))
    seg.bias.copy_(b.to(seg.bias.device))
    last = len(dec.seg_layers) - 1
    for k in list(sd.keys()):
        if k.endswith(f'decoder.seg_layers.{last}.weight'):
            sd[k] = W.clone()
        elif k.endswith(f'decoder.seg_layers.{last}.bias'):
            sd[k] = b.clone()
    net.cpu()
    return sd, net


def tiles_of(volume, labels, patch, max_tiles=4):
    """A few patch-sized crops of a phantom (corners first) for fit_seg_head."""
    import itertools
    xs = []
    ls = []
    starts = [sorted({0, max(0, s - p)}) for s, p in zip(volume.shape[1:], patch)]
    for st in itertools.islice(itertools.product(*starts), max_tiles):
        sl = tuple(slice(a, a + p) for a, p in zip(st, patch))
        v = volume[(slice(None), *sl)]
        l = labels[sl]
        pad = []
        for have, want in zip(v.shape[1:][::-1], patch[::-1]):
            pad += [0, want - have]
        if any(pad):
            v = torch.nn.functional.pad(v, pad)
            l = torch.nn.functional.pad(l, pad)
        xs.append(v)
        ls.append(l)
    return torch.stack(xs), torch.stack(ls)


_TRAINED = {}


def train_oracle(spec, steps=150, seed=7, lr=1e-2, batch=2, device=None, cell=12, noise=0.08, verbose=False,
                 max_used=None):
    """Trains the ORACLE network of `spec` for `steps` Adam steps on phantom patches (cross-entropy against the
    phantom's class map), so that the parity fixtures carry weights with real class margins, non-trivial
    InstanceNorm affine parameters and biases — the regime north_star's label bar is stated for.  With the
    reference's own initialisation (He-normal, gamma 1, beta 0) the two logits of every voxel are nearly tied
    and label agreement measures nothing but the sign of rounding noise.  Returns (state_dict on the CPU with
    every alias key, oracle module in eval mode on the CPU).  Cached per (spec, steps, seed) within a process."""
    key = (repr(sorted((k, str(v)) for k, v in spec.items())), steps, seed, cell, noise, max_used)
    if key in _TRAINED:
        sd, state = _TRAINED[key]
        net = N.build_from_arch(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], allow_init=False)
        net.load_state_dict(state, strict=True)
        net.eval()
        return {k: v.clone() for k, v in sd.items()}, net
    dev = device or (torch.device('cuda', 0) if torch.cuda.is_available() else torch.device('cpu'))
    sd0, net = make(spec, seed=1234, randomize_affine=False)
    net.to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    patch = tuple(spec['patch'])
    big = tuple(int(p * 1.5) for p in patch)
    g = torch.Generator().manual_seed(seed)
    for it in range(steps):
        if it % 8 == 0:
            vol, lab = phantom_volume(big, spec['in_ch'], spec['heads'], seed=seed * 1000 + it, cell=cell, noise=noise,
                                      max_used=max_used)
        xs, ls = [], []
        for _ in range(batch):
            st = [int(torch.randint(0, b - p + 1, (1,), generator=g)) for b, p in zip(big, patch)]
            sl = tuple(slice(a, a + p) for a, p in zip(st, patch))
            xs.append(vol[(slice(None), *sl)])
            ls.append(lab[sl])
        x = torch.stack(xs).to(dev)
        y = torch.stack(ls).to(dev)
        for gparam in opt.param_groups:
            gparam['lr'] = lr * (0.1 ** (it / max(steps, 1)))
        loss = torch.nn.functional.cross_entropy(net(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if verbose and (it % 25 == 0 or it == steps - 1):
            print(f'  train_oracle step {it}: loss {loss.item():.4f}')
    net.eval().cpu()
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd = {k: v.clone() for k, v in state.items()}
    # keys a real checkpoint carries beyond the module's own (deep-supervision heads etc.) stay as synthesised
    for k, v in sd0.items():
        sd.setdefault(k, v)
    _TRAINED[key] = (sd, state)
    return {k: v.clone() for k, v in sd.items()}, net
