"""Builds fast_nnunet_b200/libfnnu.so (sm_100a only) with nvcc, in-tree.

    python -m fast_nnunet_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.environ.get('FNNU_BUILD_OUT') or os.path.join(HERE, 'libfnnu.so')
OBJ = os.path.join(HERE, 'csrc', '_build' + ('_alt' if os.environ.get('FNNU_BUILD_OUT') else ''))
SOURCES = ['engine.cu', 'mem_kernels.cu', 'export_kernels.cu', 'preprocess_kernels.cu', 'conv_ref.cu', 'conv_umma.cu', 'conv_umma_rows.cu', 'conv_umma_zrows.cu',
           'conv_first_umma.cu', 'conv_first_zpair.cu', 'conv_tconv_umma.cu', 'conv_s2_umma.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-Xptxas', '-v'] + os.environ.get('FNNU_BUILD_DEFINES', '').split()


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for name in sorted(os.listdir(root)):
            p = os.path.join(root, name)
            if os.path.isfile(p) and name.endswith(('.cu', '.cuh', '.h')):
                h.update(name.encode())
                h.update(open(p, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


LAST_ACTION = ''      # what the last build() call did: 'compiled N sources in T s' or 'up to date (digest ...)'


def build(force: bool = False, verbose: bool = False) -> str:
    global LAST_ACTION
    import time
    stamp = os.path.join(OBJ, 'stamp')
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        LAST_ACTION = f'up to date (source digest {dig[:12]}), nothing compiled'
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    t0 = time.time()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        cmd = [NVCC, *FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        with open(obj + '.ptxas.log', 'w') as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, '-shared', '-o', OUT, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(dig)
    LAST_ACTION = f'compiled {len(SOURCES)} sources for sm_100a in {time.time() - t0:.0f} s (source digest {dig[:12]})'
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
