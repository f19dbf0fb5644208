#!/usr/bin/env python
"""Device time (CUDA events) and achieved algorithmic GB/s of the memory-bound kernels on one tile batch:
    python tools/time_mem.py        (FNNU_ACC_CLUSTER=0 selects the per-round accumulate kernels)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from fast_nnunet_b200 import _lib, engine as E  # noqa: E402
from fast_nnunet_b200 import sliding_window as sw  # noqa: E402

dev = torch.device('cuda', 0)
FLIPS8 = bytes([0, 1, 2, 4, 3, 5, 6, 7])
PEAK = 6452.8


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run(name, patch, heads, ps, n_tiles, vol, zstarts=None):
    P = int(np.prod(patch))
    step = patch[2] // 2
    starts = np.array([[0, 0, (zstarts[i] if zstarts else i * step)] for i in range(n_tiles)], dtype=np.int32)
    preds = (torch.randn((n_tiles * 8, *patch, ps), device=dev) * 2).half()
    g16 = torch.from_numpy(sw.compute_gaussian(patch, 1. / 8, 10, np.float16)).to(dev)
    acc = torch.zeros((heads, *vol), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def acc_fn():
        E.accumulate_tiles(preds.data_ptr(), _lib.IN_F16, ps, heads, starts, patch, FLIPS8, g16, acc)

    n0 = E.mem_launches()
    ms = timed(acc_fn)
    launches = (E.mem_launches() - n0) // 6
    alg = n_tiles * (8 * heads * P * 2 + P * 2 + 2 * heads * P * 4)
    print(f'{name}: accumulate {n_tiles} tiles x 8 flips, heads {heads} (stride {ps}): {ms * 1e3:8.1f} us in {launches} launch(es), '
          f'{alg / ms / 1e6:7.0f} GB/s algorithmic = {alg / ms / 1e6 / PEAK * 100:5.1f} % of {PEAK}')
    wsum = torch.empty(vol, dtype=torch.float32, device=dev)
    steps = sw.compute_steps_for_sliding_window(vol, patch, 0.5)
    E.weight_sum(steps, patch, g16, wsum)
    labels = torch.empty(vol, dtype=torch.uint8, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    V = int(np.prod(vol))

    def fin_fn():
        flush.fill_(1)
        E.finalize(acc, wsum, None, labels, flag)

    def flush_fn():
        flush.fill_(1)

    ms = timed(fin_fn) - timed(flush_fn)
    alg = heads * V * 4 + V * 4 + V
    print(f'{name}: finalize (labels only) {V / 1e6:.1f} Mvoxel: {ms * 1e3:8.1f} us, {alg / ms / 1e6:7.0f} GB/s = {alg / ms / 1e6 / PEAK * 100:5.1f} %')
    del preds, acc


print({k: v for k, v in os.environ.items() if k.startswith('FNNU_')})
run('cfg2 (2 heads, 128^3)', (128, 128, 128), 2, 2, 4, (128, 128, 320))
run('cfg4 (4 heads, 128^3)', (128, 128, 128), 4, 4, 4, (128, 128, 320))
run('cfg5 (61 heads, 160x96x96)', (160, 96, 96), 61, 64, 4, (160, 96, 240))
run('cfg5, real tile starts 0 46 92 139', (160, 96, 96), 61, 64, 4, (160, 96, 236), zstarts=[0, 46, 92, 139])
# gather
vol = torch.randn((1, 400, 512, 512), device=dev)
starts = sw.tile_starts((400, 512, 512), (128, 128, 128), 0.5)[:4]
sd = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int32)).to(dev)
out = torch.empty((32, 128, 128, 128, 1), dtype=torch.float16, device=dev)
ms = timed(lambda: E.gather_tiles(vol, sd, 4, (128, 128, 128), FLIPS8, out.data_ptr(), 1))
P = 128 ** 3
alg = 4 * (P * 4 + 8 * P * 2)
print(f'gather 4 tiles x 8 flips (C=1): {ms * 1e3:8.1f} us, {alg / ms / 1e6:7.0f} GB/s = {alg / ms / 1e6 / PEAK * 100:5.1f} %')
