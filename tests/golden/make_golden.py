"""Generate golden vectors by importing the REFERENCE's own pure functions.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Imports /root/reference/distillation/nnunetv2/inference/sliding_window_prediction.py
by file path with a stub for its one missing import (acvl_utils pad_nd_image, which the
two functions we call never touch), and
/root/reference/distillation/nnunetv2/experiment_planning/experiment_planners/network_topology.py.
Writes tests/golden/sliding_window_golden.npz (small: tile starts for many shapes, fp16
Gaussian maps for small tiles, and SHA-256 digests + probe values for the full-size maps).
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/distillation/nnunetv2'
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    stub = types.ModuleType('acvl_utils')
    stub_c = types.ModuleType('acvl_utils.cropping_and_padding')
    stub_p = types.ModuleType('acvl_utils.cropping_and_padding.padding')
    stub_p.pad_nd_image = lambda *a, **k: None
    sys.modules.update({'acvl_utils': stub, 'acvl_utils.cropping_and_padding': stub_c,
                        'acvl_utils.cropping_and_padding.padding': stub_p})
    swp = _load(os.path.join(REF, 'inference/sliding_window_prediction.py'), 'ref_swp')
    topo = _load(os.path.join(REF, 'experiment_planning/experiment_planners/network_topology.py'), 'ref_topo')

    out = {}
    # ---- tile starts: BASELINE configs + edge cases (equal size, ragged, step 1.0, tiny steps)
    cases = [
        ((160, 160, 160), (128, 128, 128), 0.5),
        ((400, 512, 512), (128, 128, 128), 0.5),
        ((155, 240, 240), (128, 128, 128), 0.5),
        ((1200, 512, 512), (160, 96, 96), 0.5),
        ((128, 128, 128), (128, 128, 128), 0.5),
        ((129, 128, 255), (128, 128, 128), 0.5),
        ((110, 110, 110), (64, 64, 64), 0.5),
        ((300, 301, 302), (96, 112, 128), 0.5),
        ((200, 200, 200), (64, 64, 64), 1.0),
        ((97, 131, 77), (32, 48, 40), 0.25),
        ((48, 64, 56), (32, 32, 32), 0.5),
        ((40, 40, 40), (32, 32, 32), 0.5),
        ((37, 53, 61), (32, 32, 32), 0.75),
        ((122, 101, 96), (28, 96, 96), 0.5),
    ]
    step_cases = []
    for img, tile, st in cases:
        steps = swp.compute_steps_for_sliding_window(img, tile, st)
        step_cases.append({'image': list(img), 'tile': list(tile), 'step': st, 'steps': steps})
    # ---- gaussian maps
    gauss_meta = []
    for tile in [(128, 128, 128), (160, 96, 96), (32, 32, 32), (16, 24, 20), (28, 96, 96), (8, 8, 8), (64, 64, 64)]:
        swp.compute_gaussian.cache_clear()
        g16 = swp.compute_gaussian(tuple(tile), sigma_scale=1. / 8, value_scaling_factor=10,
                                   dtype=torch.float16, device=torch.device('cpu'))
        swp.compute_gaussian.cache_clear()
        g32 = swp.compute_gaussian(tuple(tile), sigma_scale=1. / 8, value_scaling_factor=10,
                                   dtype=torch.float32, device=torch.device('cpu'))
        a16 = g16.numpy()
        a32 = g32.numpy()
        meta = {
            'tile': list(tile),
            'sha256_fp16': hashlib.sha256(a16.tobytes()).hexdigest(),
            'sha256_fp32': hashlib.sha256(a32.tobytes()).hexdigest(),
            'max': float(a16.astype(np.float32).max()),
            'min': float(a16.astype(np.float32).min()),
            'sum_fp32_of_fp16': float(a16.astype(np.float64).sum()),
            'n_at_floor': int((a16 == a16.min()).sum()),
        }
        gauss_meta.append(meta)
        if np.prod(tile) <= 32 ** 3:
            out['gauss16_' + 'x'.join(map(str, tile))] = a16
            out['gauss32_' + 'x'.join(map(str, tile))] = a32
        else:
            # keep the three central axis profiles + one diagonal (tiny) for eyeballing failures
            c = [t // 2 for t in tile]
            out['gauss16_prof0_' + 'x'.join(map(str, tile))] = a16[:, c[1], c[2]].copy()
            out['gauss16_prof1_' + 'x'.join(map(str, tile))] = a16[c[0], :, c[2]].copy()
            out['gauss16_prof2_' + 'x'.join(map(str, tile))] = a16[c[0], c[1], :].copy()
    # ---- topology for the bone_turbo-shaped config (cfg5) and the isotropic 128^3 one
    topo_cases = []
    for spacing, patch in [((2.0, 0.9765625, 0.9765625), (160, 96, 96)), ((1.0, 1.0, 1.0), (128, 128, 128))]:
        res = topo.get_pool_and_conv_props(spacing, patch, 4, 999999)
        topo_cases.append({'spacing': list(spacing), 'patch': list(patch),
                           'num_pool_per_axis': [int(i) for i in res[0]],
                           'pool_op_kernel_sizes': [[int(j) for j in i] for i in res[1]],
                           'conv_kernel_sizes': [[int(j) for j in i] for i in res[2]],
                           'patch_size': [int(i) for i in res[3]],
                           'must_be_divisible_by': [int(i) for i in res[4]]})
    np.savez_compressed(os.path.join(HERE, 'sliding_window_golden.npz'), **out)
    with open(os.path.join(HERE, 'sliding_window_golden.json'), 'w') as f:
        json.dump({'steps': step_cases, 'gaussian': gauss_meta, 'topology': topo_cases,
                   'generated_from': 'reference sliding_window_prediction.py:10-54, network_topology.py:30-108',
                   'torch': torch.__version__, 'numpy': np.__version__}, f, indent=1)
    print('wrote', len(step_cases), 'step cases,', len(gauss_meta), 'gaussian maps,', len(topo_cases), 'topologies')


if __name__ == '__main__':
    main()
