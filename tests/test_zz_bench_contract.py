"""bench.py prints ONE JSON line with the fields the driver reads.  The reference arm (CPU, oracle port) is checked
here; our arm needs a B200 (`-m gpu`).  Named zz so that it runs after the parity tests."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline', 'gpu_launches'}


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                         timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = _run(['--impl', 'reference', '--workload', 'cfg1', '--steps', '1', '--warmup', '0'],
             {'FNNU_BENCH_REF_TILES': '1'})
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d['impl'] == 'reference' and d['unit'] == 'Mvoxel/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0 and d['vs_baseline'] is None
    assert d['config']['workload']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'tiles' in cb['sample']


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    from fast_nnunet_b200 import sliding_window as sw
    expect = {'cfg1': 8, 'cfg2': 294, 'cfg3': 294, 'cfg4': 18, 'cfg5': 1400}
    for name, n in expect.items():
        vol, patch = bench.WORKLOADS[name][0], bench.WORKLOADS[name][4]
        assert len(sw.tile_starts(vol[1:], patch, 0.5)) == n, name
    dom = {'cin': 32, 'cout': 16, 'out_dims': [128, 128, 128], 'kernel': [3, 3, 3]}
    assert bench._conv_kernel_name(dom) == 'conv_umma_zrows_kernel'       # z-pair row streaming: Cout <= 16, 3x3x3
    dom = {'cin': 32, 'cout': 32, 'out_dims': [64, 64, 64], 'kernel': [3, 3, 3]}
    assert bench._conv_kernel_name(dom) == 'conv_umma_rows_kernel'
    dom = {'cin': 16, 'cout': 16, 'out_dims': [160, 96, 96], 'kernel': [1, 3, 3]}
    assert bench._conv_kernel_name(dom) == 'conv_umma_rows_kernel'
    dom = {'cin': 64, 'cout': 32, 'out_dims': [64, 64, 64], 'kernel': [3, 3, 3]}
    assert bench._conv_kernel_name(dom) == 'conv_umma_kernel'


@pytest.mark.gpu
def test_our_arm_contract():
    d = _run(['--workload', 'cfg1', '--steps', '3', '--warmup', '3', '--no-cpu-baseline', '--no-torch-gpu-baseline'])
    assert (BASE_KEYS - {'cpu_baseline'}) <= set(d), sorted(BASE_KEYS - set(d))
    assert d['n_gpus'] == 1 and d['unit'] == 'Mvoxel/s' and d['value'] > 0 and d['gpu_launches'] > 0
    for key in ('roofline', 'roofline_network', 'roofline_aggregation'):
        r = d[key]
        assert r['peak'] > 0 and 0 < r['frac'] < 1 and r['bound'] in ('hbm', 'tensor'), (key, r)
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    # the end-to-end figure includes the copies; cfg1 is a 40 ms volume, so allow run-to-run noise between the two loops
    assert e['value'] <= d['value'] * 1.5
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
