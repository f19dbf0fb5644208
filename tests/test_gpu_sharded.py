"""GPU, >= 2 devices: one volume sharded over the GPUs of the box (x-slab tiles + NCCL halo exchange) equals the
single-GPU result.  Skipped on a 1-GPU box; the partition/exchange logic itself is covered on CPU over gloo
(tests/test_sharding.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_equals_single(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    port = 29700 + world
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
                        '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(HERE, 'sharded_check.py')],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and 'SHARDED OK' in r.stdout
