#!/usr/bin/env python
"""Per-operator device time of one launch sequence (CUDA events inside libfnnu, no profiler):
    python tools/time_ops.py [student|teacher|resenc|bone] [batch] [repeats]
Environment switches of libfnnu (FNNU_ZROWS, FNNU_ZROWS_ISSUERS, FNNU_FIRST_LAYER_TC ...) apply."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import nets  # noqa: E402
from fast_nnunet_b200.predictor import CompiledNetwork  # noqa: E402

name = {'student': 'STUDENT', 'teacher': 'TEACHER', 'resenc': 'RESENC_M_STUDENT', 'bone': 'BONE_TURBO'}[sys.argv[1] if len(sys.argv) > 1 else 'student']
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
spec = getattr(nets, name)
sd, _ = nets.make(spec, randomize_affine=False)
dev = torch.device('cuda', 0)
cn = CompiledNetwork(spec['cls'], spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
cn.load_state_dict(sd)
n_ops = int(os.environ.get('TIME_OPS_TRUNCATE', '0'))
if n_ops:
    # run only the first n operators (the instrumented builds report the LAST row-streaming launch)
    from fast_nnunet_b200.engine import NetworkEngine
    from fast_nnunet_b200.program import build_program
    prog = build_program(spec['cls'], sd, spec['kw'], spec['in_ch'], spec['heads'], spec['patch'])
    prog.ops = prog.ops[:n_ops]
    eng = NetworkEngine(prog, batch, dev)
else:
    eng = cn.engine(dev, batch)
prog = eng.program
inp = eng.buffer_tensor(prog.input_buffer, batch)
inp.copy_(torch.randn(inp.shape, device=dev).half())
for _ in range(2):
    eng.forward(batch)
torch.cuda.synchronize()
total = 0.0
tf = 0.0
rows = []
for i, op in enumerate(prog.ops):
    eng.profile_op(i)
    ts = []
    for _ in range(reps):
        eng.forward(batch)
        ts.append(eng.profile_ms())
    ms = float(np.median(ts))
    fl = op.flops(prog.buffers[op.src][0], prog.buffers[op.dst][0]) * batch
    rows.append((i, op.name, op.cin, op.cout, prog.buffers[op.dst][0], ms, fl / ms / 1e9 if ms > 0 else 0))
    total += ms
    tf += fl
eng.profile_op(-1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    eng.forward(batch)
ev1.record()
torch.cuda.synchronize()
whole = ev0.elapsed_time(ev1) / reps
print(f'{name} batch={batch} env={ {k: v for k, v in os.environ.items() if k.startswith("FNNU_")} }')
for r in rows:
    print(f'{r[0]:3d} {r[1]:42s} {r[2]:4d}->{r[3]:<4d} {str(r[4]):18s} {r[5]:8.3f} ms {r[6]:8.1f} TFLOP/s')
print(f'sum of ops {total:.3f} ms; whole forward {whole:.3f} ms; {tf / whole / 1e9:.1f} TFLOP/s '
      f'({tf / whole / 1e9 / 1398.2 * 100:.1f} % of 1398.2 sustained)')
try:
    import ctypes
    from fast_nnunet_b200 import _lib as _L
    lib = ctypes.CDLL(_L.LIB_PATH)
    if hasattr(lib, 'fnnu_debug_zrows_prof'):
        buf = (ctypes.c_longlong * 32)()
        lib.fnnu_debug_zrows_prof(buf)
        v = list(buf)
        print('zrows role cycles of CTA 0, LAST zrows launch:')
        print(f'  producer warp 0: wait stage-free {v[0]}, stages filled {v[1]}, total {v[2]}')
        for m in range(2):
            o = 8 + m * 8
            print(f'  MMA warp {m}: wait tempty {v[o]}, wait full {v[o + 1]}, issue+commit {v[o + 2]}, steps {v[o + 3]}, total {v[o + 4]}')
        for k in range(2):
            o = 24 + k * 4
            print(f'  epilogue set {k}: wait step {v[o]}, rows {v[o + 1]}, total {v[o + 2]}')
except Exception as e:  # profiling builds only
    print('no zrows profile:', e)
