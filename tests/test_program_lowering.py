"""CPU: program.py's lowering of PlainConvUNet / ResidualEncoderUNet checkpoints, executed by the plain-PyTorch program
interpreter (tests/program_interp.py), against the oracle network in fp32."""
import pytest
import torch

import nets
from fast_nnunet_b200 import model_folder as M
from fast_nnunet_b200.program import build_program
from oracle import networks as N
from program_interp import run_program


@pytest.mark.parametrize('name', ['SMALL_PLAIN', 'ANISO_PLAIN', 'SMALL_RESENC', 'TINY_ONNX', 'ROWS_W96', 'ZROWS_ODD_D'])
def test_lowered_program_equals_oracle_network(name):
    spec = getattr(nets, name)
    sd, net = nets.make(spec)                       # random weights, biases, gamma, beta
    prog = build_program(spec['cls'], sd, spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
    x = torch.randn((2, spec['in_ch'], *spec['patch']), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = net(x)
        got = run_program(prog, x)
    assert got.shape == want.shape
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    print(f'{name}: {len(prog.ops)} ops, {len(prog.buffers)} buffers, max|d| {err:.2e} (logit range {scale:.2f})')
    assert err <= 2e-4 * max(1.0, scale)


def test_student_checkpoint_from_teacher_plans_lowers_and_matches():
    """A distilled student: the plans describe the teacher, the checkpoint holds the reduced network."""
    spec = nets.SMALL_PLAIN16
    cls, kw = M.effective_arch(spec['cls'], spec['kw'], 'nnUNetDistillationTrainer', {'feature_reduction_factor': 2})
    assert kw['features_per_stage'] == [8, 16, 32]
    sd = M.synthesize_state_dict(cls, kw, 1, 2, seed=5, randomize_affine=True)
    net = N.build_from_arch(cls, kw, 1, 2, allow_init=False)
    net.load_state_dict(sd, strict=True)
    net.eval()
    prog = build_program(cls, sd, kw, 1, 2, spec['patch'])
    x = torch.randn((1, 1, *spec['patch']), generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        err = (run_program(prog, x) - net(x)).abs().max().item()
    assert err <= 2e-4


@pytest.mark.parametrize('name,patch', [('STUDENT', (64, 64, 64)), ('TEACHER', (64, 64, 64)), ('RESENC_M_STUDENT', (64, 64, 64)),
                                        ('BONE_TURBO', (32, 32, 32))])
def test_baseline_architectures_lower_correctly(name, patch):
    """The BASELINE.json architectures (all stages, block counts, anisotropic kernels / strides, 61 heads) at a patch the
    CPU finishes in seconds: the lowering does not depend on the patch size beyond the buffer extents."""
    spec = dict(getattr(nets, name))
    spec['patch'] = patch
    sd, net = nets.make(spec)
    prog = build_program(spec['cls'], sd, spec['kw'], spec['in_ch'], spec['heads'], patch)
    x = torch.randn((1, spec['in_ch'], *patch), generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = net(x)
        got = run_program(prog, x)
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    print(f'{name}: {len(prog.ops)} ops, max|d| {err:.2e} (logit range {scale:.2f})')
    assert got.shape == want.shape and err <= 5e-4 * max(1.0, scale)
