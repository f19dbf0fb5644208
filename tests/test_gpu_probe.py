"""GPU: the tcgen05 descriptor/layout conventions conv_umma.cu is built on, checked in isolation."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_umma_probe(tmp_path):
    exe = str(tmp_path / 'probe_umma')
    subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O2', '-o', exe,
                    os.path.join(HERE, 'cuda', 'probe_umma.cu')], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and 'PROBE OK' in r.stdout
