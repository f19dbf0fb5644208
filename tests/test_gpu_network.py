"""GPU parity of the per-patch network forward (libfnnu engine) against the oracle network.

Tolerances (stated, per north_star): the engine stores activations in fp16 with fp32 accumulation and
fp64 InstanceNorm sums; the oracle runs fp32.  Logits are O(1); we require max|d| <= 0.08 and
mean|d| <= 0.01 against the fp32 oracle for the shallow test nets, and every intermediate raw conv
output within 2% of its dynamic range."""
import numpy as np
import pytest
import torch

import nets
from fast_nnunet_b200 import _lib
from fast_nnunet_b200.predictor import CompiledNetwork

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _run(spec, batch, backend, seed=0):
    sd, net = nets.make(spec)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, spec['in_ch'], *spec['patch']), generator=g)
    cn = CompiledNetwork(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
    cn.load_state_dict(sd)
    eng = cn.engine(DEV, batch)
    eng.set_backend(backend)
    got = cn(x.to(DEV)).float().cpu()
    with torch.no_grad():
        want = net(x.half().float())       # the engine sees the fp16-rounded input
    return got, want, eng


@pytest.mark.parametrize('name', ['SMALL_PLAIN', 'SMALL_PLAIN16', 'ANISO_PLAIN', 'SMALL_RESENC', 'ROWS_W128', 'ROWS_W96', 'ROWS_W64_C32',
                                  'ZROWS_W96', 'ZROWS_W72_H9', 'ZROWS_ODD_D'])
@pytest.mark.parametrize('backend', [1, 0])
def test_forward_matches_oracle(name, backend):
    spec = getattr(nets, name)
    got, want, eng = _run(spec, 3, backend)
    assert got.shape == want.shape
    d = (got - want).abs()
    scale = want.abs().max().item()
    print(f'{name} backend={backend}: max|d|={d.max():.4f} mean|d|={d.mean():.5f} logit range={scale:.2f} '
          f'launches={eng.launch_counts()}')
    assert torch.isfinite(got).all()
    assert d.max().item() <= 0.08 * max(1.0, scale / 4)
    assert d.mean().item() <= 0.01 * max(1.0, scale / 4)


def test_batch_entries_are_independent():
    """InstanceNorm statistics are per (sample, channel): a sample's output must not depend on its batch-mates."""
    spec = nets.SMALL_PLAIN16
    sd, _ = nets.make(spec)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((4, 1, *spec['patch']), generator=g).to(DEV)
    cn = CompiledNetwork(spec['cls'], spec['kw'], 1, 2, spec['patch'])
    cn.load_state_dict(sd)
    full = cn(x)
    single = cn(x[2:3])
    d = (full[2:3].float() - single.float()).abs().max().item()
    assert d <= 2e-2, d      # only the order of the fp64 atomic sums may differ


def test_intermediate_buffers_and_stats():
    """First encoder conv: raw output and its InstanceNorm sums against torch."""
    spec = nets.SMALL_PLAIN16
    sd, net = nets.make(spec)
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 1, *spec['patch']), generator=g)
    cn = CompiledNetwork(spec['cls'], spec['kw'], 1, 2, spec['patch'])
    cn.load_state_dict(sd)
    cn(x.to(DEV))
    eng = cn.engine(DEV, 2)
    op0 = eng.program.ops[0]
    raw = eng.buffer_tensor(op0.dst, 2)[..., op0.dst_coff:op0.dst_coff + op0.cout].float().cpu()
    conv = net.encoder.stages[0][0].convs[0].conv
    with torch.no_grad():
        # the engine does not lower a bias that feeds an InstanceNorm (program.py: IN(x + b) = IN(x))
        want = torch.nn.functional.conv3d(x.half().float(), conv.weight, None, conv.stride, conv.padding).permute(0, 2, 3, 4, 1)
    assert (raw - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-3
    st = eng.stats_tensor(op0.dst, 2)[:, op0.dst_coff:op0.dst_coff + op0.cout].cpu()
    s1 = raw.double().sum(dim=(1, 2, 3))
    s2 = (raw.double() ** 2).sum(dim=(1, 2, 3))
    assert torch.allclose(st[..., 0], s1, rtol=1e-5, atol=1e-2)
    assert torch.allclose(st[..., 1], s2, rtol=1e-5, atol=1e-2)


def test_student_128_forward_matches_oracle():
    """Full-size distilled student (r=2), trained oracle weights (tests/nets.py: real margins, non-trivial
    gamma / beta / bias), 128^3 phantom patches: north_star's label bar on every voxel."""
    spec = nets.STUDENT
    sd, net = nets.train_oracle(spec, steps=120, batch=1)
    vol, _ = nets.phantom_volume((128, 128, 256), 1, 2, seed=11)
    x = torch.stack([vol[:, :, :, :128], vol[:, :, :, 128:]])
    cn = CompiledNetwork(spec['cls'], spec['kw'], 1, 2, spec['patch'])
    cn.load_state_dict(sd)
    got = cn(x.to(DEV)).float().cpu()
    with torch.no_grad():
        want = net.to(DEV)(x.half().float().to(DEV)).cpu()      # fp32 oracle network (TF32 off), on the GPU for speed
    d = (got - want).abs()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    rng = want.abs().max().item()
    print(f'student128 (trained): max|d|={d.max():.4f} mean|d|={d.mean():.5f} range={rng:.2f} argmax agreement={agree:.6f}')
    assert d.max().item() <= 0.06 * max(1.0, rng / 8) and d.mean().item() <= 0.006 * max(1.0, rng / 8)
    assert agree >= 0.999


def test_student_128_he_init_forward_matches_oracle():
    """The same network with the reference's own initialisation (He-normal, gamma 1, beta 0): logits only, the two
    class logits of every voxel are nearly tied, so label agreement is not a criterion here."""
    spec = nets.STUDENT
    sd, net = nets.make(spec, randomize_affine=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 1, 128, 128, 128), generator=g)
    cn = CompiledNetwork(spec['cls'], spec['kw'], 1, 2, spec['patch'])
    cn.load_state_dict(sd)
    got = cn(x.to(DEV)).float().cpu()
    with torch.no_grad():
        want = net.to(DEV)(x.half().float().to(DEV)).cpu()
    d = (got - want).abs()
    rng = want.abs().max().item()
    print(f'student128 (He init, random affine): max|d|={d.max():.4f} mean|d|={d.mean():.5f} range={rng:.2f}')
    assert d.max().item() <= 0.15 * max(1.0, rng / 8) and d.mean().item() <= 0.01 * max(1.0, rng / 8)


@pytest.mark.parametrize('name', ['SMALL_PLAIN16', 'ANISO_PLAIN', 'SMALL_RESENC', 'ROWS_W128', 'ROWS_W96', 'ROWS_W64_C32',
                                  'ZROWS_W96', 'ZROWS_W72_H9', 'ZROWS_ODD_D', 'TINY_ONNX'])
def test_tcgen05_layers_match_direct_kernel(name):
    """Every activation buffer written with the tcgen05 implicit-GEMM back end against the CUDA-core direct
    kernel (same fp16 inputs, fp32 accumulation in both): differences are fp32 summation order + one fp16 ulp."""
    spec = getattr(nets, name)
    sd, _ = nets.make(spec)
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, spec['in_ch'], *spec['patch']), generator=g).to(DEV)
    cn = CompiledNetwork(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
    cn.load_state_dict(sd)
    eng = cn.engine(DEV, 3)
    eng.set_backend(1)
    cn(x)
    ref = [eng.buffer_tensor(i, 3).float().clone() for i in range(len(eng.program.buffers))]
    ref_stats = [eng.stats_tensor(i, 3).clone() for i in range(len(eng.program.buffers))]
    eng.set_backend(0)
    cn(x)
    total, umma = eng.launch_counts()
    assert umma > 0, 'no layer ran on the tcgen05 back end'
    worst = 0.0
    for op in eng.program.ops:
        got = eng.buffer_tensor(op.dst, 3).float()[..., op.dst_coff:op.dst_coff + op.cout]
        want = ref[op.dst][..., op.dst_coff:op.dst_coff + op.cout]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs().max().item() / scale
        worst = max(worst, err)
        assert err <= 4e-3, f'{op.name}: relative error {err:.5f} (cin={op.cin} cout={op.cout} k={op.kernel} s={op.stride})'
        if op.has_norm:
            st = eng.stats_tensor(op.dst, 3)[:, op.dst_coff:op.dst_coff + op.cout]
            rs = ref_stats[op.dst][:, op.dst_coff:op.dst_coff + op.cout]
            # a few fp16 roundings flip with the fp32 summation order, so compare the implied mean / std
            n = float(np.prod(eng.program.buffers[op.dst][0]))
            mean_a, mean_b = st[..., 0] / n, rs[..., 0] / n
            std_a = (st[..., 1] / n - mean_a ** 2).clamp_min(0).sqrt()
            std_b = (rs[..., 1] / n - mean_b ** 2).clamp_min(0).sqrt()
            assert ((mean_a - mean_b).abs() <= 2e-3 * std_b + 1e-6).all(), f'{op.name}: InstanceNorm mean differs'
            assert ((std_a / std_b - 1).abs() <= 2e-3).all(), f'{op.name}: InstanceNorm std differs'
    print(f'{name}: {umma}/{total} launches on tcgen05, worst relative layer error {worst:.2e}')
