#!/usr/bin/env python
"""Headline benchmark: sliding-window segmentation of a synthetic 512x512x400 CT with the distilled
student (r=2 PlainConvUNet, 128^3 patches, step 0.5, mirror TTA) — BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one whole volume through the hot path.  Prints ONE JSON line (rank 0).  See DESIGN.md
section "Measurement" for how every field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ncu --set full on the dominant launch, conv_umma_rows_kernel<16,2> = dec5.0 with 32 patches
# (profiles/r01_ncu_final_rows_dec5.txt): dram read 5.070 GB + write 2.121 GB for 4.295 + 2.147 GB algorithmic
# -> measured DRAM traffic / algorithmic bytes (the 18 % extra reads are z-neighbour planes that missed L2)
NCU_TRAFFIC_OVER_ALGORITHMIC = (5.0698 + 2.1215) / (4.2950 + 2.1475)
# bounded CPU samples: ~3.3 s per 128^3 student tile x 8 passes on 16 cores
CPU_BASELINE_TILES = 4      # cpu_baseline of our arm: ~13 s of CPU work
REF_TILES_PER_STEP = int(os.environ.get('FNNU_BENCH_REF_TILES', '3'))   # --impl reference: ~10 s per step
STUDENT_FEATS = [16, 32, 64, 128, 160, 160]
TEACHER_FEATS = [32, 64, 128, 256, 320, 320]
ISO_K = [[3, 3, 3]] * 6
ISO_S = [[1, 1, 1]] + [[2, 2, 2]] * 5

BONE_K = [[1, 3, 3]] + [[3, 3, 3]] * 5
BONE_S = [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 1, 1]]

WORKLOADS = {
    # name: (volume (C,X,Y,Z), features, kernels, strides, patch, heads, description[, ResEnc blocks per stage])
    'cfg1': ((1, 160, 160, 160), STUDENT_FEATS, ISO_K, ISO_S, (128, 128, 128), 2,
             'student r=2 PlainConvUNet, 1x160^3 CT, 128^3 patches, step 0.5, mirror TTA'),
    'cfg2': ((1, 400, 512, 512), STUDENT_FEATS, ISO_K, ISO_S, (128, 128, 128), 2,
             'student r=2 PlainConvUNet, 512x512x400 CT, 128^3 patches, step 0.5, mirror TTA'),
    'cfg3': ((1, 400, 512, 512), TEACHER_FEATS, ISO_K, ISO_S, (128, 128, 128), 2,
             'teacher PlainConvUNet, 512x512x400 CT, 128^3 patches, step 0.5, mirror TTA'),
    'cfg4': ((4, 155, 240, 240), STUDENT_FEATS, ISO_K, ISO_S, (128, 128, 128), 4,
             'ResEnc-M student r=2 (blocks 1,3,4,6,6,6), 4-channel MRI 240x240x155, 128^3 patches, mirror TTA',
             [1, 3, 4, 6, 6, 6]),
    'cfg5': ((1, 1200, 512, 512), STUDENT_FEATS, BONE_K, BONE_S, (160, 96, 96), 61,
             'bone_turbo-shaped r=2 PlainConvUNet, 61 labels, 512x512x1200 CT, 160x96x96 patches, mirror TTA'),
}


def synth_volume(shape, seed=0):
    """CT-like intensities, then the reference's CT normalisation."""
    g = torch.Generator().manual_seed(seed)
    c = shape[0]
    low_shape = [max(s // 16, 2) for s in shape[1:]]
    low = torch.nn.functional.interpolate(torch.randn((1, c, *low_shape), generator=g), size=tuple(shape[1:]),
                                          mode='trilinear', align_corners=False)[0]
    v = torch.randn(shape, generator=g).mul_(225.0).add_(low.mul_(400.0)).sub_(350.0).clamp_(-1100, 1207)
    return v.add_(350.0).div_(450.0).contiguous()


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}',
                 '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            parts = [p.strip() for p in r.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(np.max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def build_folder(tmp, wl):
    from fast_nnunet_b200 import model_folder as M
    vol, feats, ks, ss, patch, heads = WORKLOADS[wl][:6]
    blocks = WORKLOADS[wl][7] if len(WORKLOADS[wl]) > 7 else None
    cls = M.RESENC if blocks else M.PLAIN
    kw = M.resenc_arch_kwargs(feats, ks, ss, blocks) if blocks else M.plain_arch_kwargs(feats, ks, ss)
    sd = M.synthesize_state_dict(cls, kw, vol[0], heads, seed=1234)
    folder = os.path.join(tmp, 'nnUNetTrainer__nnUNetPlans__3d_fullres')
    M.write_model_folder(folder, cls, kw, patch, sd, vol[0], heads)
    return folder, kw, sd


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU torch restatement of the reference path) on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, n_tiles_sample, threads, seed_vol=0):
    """Times `n_tiles_sample` tiles x 8 mirror passes of the reference's CPU arithmetic (fp32 network, fp16
    accumulators) through the oracle and extrapolates to the whole volume by tile count."""
    from fast_nnunet_b200 import model_folder as M
    from fast_nnunet_b200 import sliding_window as sw
    from oracle import networks as N
    from oracle import predictor as OP
    vol, feats, ks, ss, patch, heads = WORKLOADS[wl][:6]
    blocks = WORKLOADS[wl][7] if len(WORKLOADS[wl]) > 7 else None
    cls = M.RESENC if blocks else M.PLAIN
    kw = M.resenc_arch_kwargs(feats, ks, ss, blocks) if blocks else M.plain_arch_kwargs(feats, ks, ss)
    sd = M.synthesize_state_dict(cls, kw, vol[0], heads, seed=1234)
    net = N.build_from_arch(cls, kw, vol[0], heads, allow_init=False)
    net.load_state_dict(sd)
    net.eval()
    n_total = len(sw.tile_starts(vol[1:], patch, 0.5))
    torch.set_num_threads(threads)
    # a sub-volume that holds exactly the sampled tiles keeps host memory small; the arithmetic per tile is identical
    x = synth_volume((vol[0], patch[0], patch[1], patch[2] + (n_tiles_sample - 1) * (patch[2] // 2)), seed_vol)
    t0 = time.perf_counter()
    OP.predict_sliding_window_return_logits(net, x, patch, 0.5, True, (0, 1, 2))
    dt = time.perf_counter() - t0
    sec_per_volume = dt / n_tiles_sample * n_total
    return sec_per_volume, dt, n_total


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vol = WORKLOADS[wl][0]
    nvox = float(np.prod(vol[1:]))
    cores = os.cpu_count() or 1
    times = []
    for i in range(args.warmup + args.steps):
        spv, dt, n_total = cpu_reference_sample(wl, REF_TILES_PER_STEP, cores)
        if i >= args.warmup:
            times.append(spv)
    spv = float(np.mean(times))
    value = nvox / spv / 1e6
    sample = f'{REF_TILES_PER_STEP} of {n_total} tiles x 8 mirror passes per step (oracle port of the reference CPU path, fp32 network, ' \
             f'fp16 accumulators), extrapolated by tile count'
    line = {
        'impl': 'reference', 'metric': 'sliding-window inference throughput', 'value': value, 'unit': 'Mvoxel/s',
        'sec_per_volume': spv, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': spv * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'fp32', 'data': 'synthetic', 'config': {'workload': WORKLOADS[wl][6]},
        'cpu_baseline': {'value': value, 'unit': 'Mvoxel/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'Mvoxel/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def _conv_kernel_name(dom):
    """Which tcgen05 kernel the engine picks for a stride-1 3x3x3 conv (mirrors plan_rows in conv_umma_rows.cu): the
    row-streaming kernel for Cin in {16, 32}, Cout <= 32 and 64 <= W <= 128, else the general implicit GEMM."""
    cout_pad = (dom['cout'] + 15) // 16 * 16
    w = dom['out_dims'][2]
    rows = dom['cin'] in (16, 32) and cout_pad in (16, 32) and 64 <= w <= 128
    return 'conv_umma_rows_kernel' if rows else 'conv_umma_kernel'


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch.distributed as dist
    from fast_nnunet_b200 import nnUNetPredictor
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    vol, feats, ks, ss, patch, heads, desc = WORKLOADS[wl][:7]
    nvox = float(np.prod(vol[1:]))

    with tempfile.TemporaryDirectory() as tmp:
        folder, kw, sd = build_folder(tmp, wl)
        pred = nnUNetPredictor(tile_step_size=0.5, use_gaussian=True, use_mirroring=True, device=dev,
                               allow_tqdm=False, tiles_per_batch=args.tiles_per_batch)
        pred.initialize_from_trained_model_folder(folder, use_folds=(0,))
    host = synth_volume(vol, 0).pin_memory()
    data = host.to(dev)
    peaks = measured_peaks()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        if world > 1:
            return pred.predict_sliding_window_sharded(data, return_labels=True, gather_to=0)
        return pred.predict_sliding_window_return_segmentation(data)

    # ---- warm-up (also builds the engine)
    for _ in range(max(args.warmup, 1)):
        out = step_resident()
    del out
    sync_all()
    # ---- timed: K volumes, device-resident input, label map left on the device
    pred.collect_timing = True
    pred.timing = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
        launches += pred.last_launches
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    phase = pred.timing_summary()
    pred.collect_timing = False
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- e2e: the reference-facing call with HOST buffers (pinned input, logits returned on the CPU)
    e2e_ms = None
    h2d = d2h = 0
    if world == 1:
        for _ in range(1):
            pred.predict_logits_from_preprocessed_data(host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            res = pred.predict_logits_from_preprocessed_data(host)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
        h2d = host.numel() * host.element_size()
        d2h = res.numel() * res.element_size()
        del res
    else:
        sync_all()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            d = host.to(dev, non_blocking=True)
            lab = pred.predict_sliding_window_sharded(d, return_labels=True, gather_to=0)
            if rank == 0:
                lab_host = lab.cpu()
        sync_all()
        e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
        h2d = host.numel() * host.element_size()
        d2h = int(nvox) if rank == 0 else 0
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    dom = pred.profile_dominant_op(data) if rank == 0 else None
    if rank == 0:
        prog = pred.network.program
        flops_fwd = prog.total_flops()
        n_tiles = len(pred._internal_get_sliding_window_slicers(vol[1:]))
        flops_vol = flops_fwd * n_tiles * 8
        conv_ms = phase.get('forward_ms')        # per volume, rank 0, CUDA events around every engine forward
        acc_ms = phase.get('accumulate_ms')
        pvox = float(np.prod(patch))
        acc_bytes = n_tiles * (8 * heads * pvox * 2 + pvox * 2 + 2 * heads * pvox * 4) / world
        # dominant kernel: the operator with the most FLOPs (dec5.0, Conv3d 32->16 @128^3 for the student), timed by
        # CUDA events inside libfnnu on the launching stream; `traffic` is dram read+write of one ncu --set full
        # capture of the same kernel class (profiles/README.md), scaled to this launch's patch count
        roof = {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': peaks['bf16_tflops_sustained'],
                'achieved': dom['flop_per_launch'] / (dom['ms'] * 1e-3) / 1e12,
                'traffic': (dom['algorithmic_bytes_per_launch'] * NCU_TRAFFIC_OVER_ALGORITHMIC
                            if _conv_kernel_name(dom) == 'conv_umma_rows_kernel' else None),
                'algorithmic_bytes': dom['algorithmic_bytes_per_launch'],
                'peak_source': peaks['source'] + ' (sustained cuBLAS bf16; fp16 runs at the same tcgen05 rate)',
                'kernel': _conv_kernel_name(dom) + ' (' + dom['op'] + f", Conv3d {dom['cin']}->{dom['cout']} @{dom['out_dims']}"
                          f", {dom['patches_per_launch']} patches per launch)",
                'ms_per_launch': dom['ms'], 'flop_per_launch': dom['flop_per_launch']}
        roof['frac'] = roof['achieved'] / roof['peak']
        roof_all = {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': peaks['bf16_tflops_sustained'],
                    'achieved': (flops_vol / world / (conv_ms * 1e-3) / 1e12) if conv_ms else None,
                    'kernel': 'every launch of the network forward (convs, transposed convs, first layer, seg head)',
                    'flop_per_forward': flops_fwd}
        roof_all['frac'] = roof_all['achieved'] / roof_all['peak'] if roof_all['achieved'] else None
        roof_mem = {'bound': 'hbm', 'unit': 'GB/s', 'peak': peaks['hbm_gbs'],
                    'achieved': (acc_bytes / (acc_ms * 1e-3) / 1e9) if acc_ms else None, 'traffic': None,
                    'kernel': 'accumulate_h2_vec4_kernel', 'bytes_per_tile': acc_bytes * world / n_tiles}
        roof_mem['frac'] = roof_mem['achieved'] / roof_mem['peak'] if roof_mem['achieved'] else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            spv, dt, n_total = cpu_reference_sample(wl, CPU_BASELINE_TILES, cores)
            cpu = {'value': nvox / spv / 1e6, 'unit': 'Mvoxel/s', 'cores': cores, 'kind': 'port',
                   'sec_per_volume': spv,
                   'sample': f'{CPU_BASELINE_TILES} of {n_total} tiles x 8 mirror passes ({dt:.1f} s of CPU work), '
                             f'extrapolated by tile count'}
        line = {
            'metric': 'sliding-window inference throughput', 'value': nvox / (ms * 1e-3) / 1e6, 'unit': 'Mvoxel/s',
            'sec_per_volume': ms * 1e-3, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'fp16 storage / fp32 accumulate (tcgen05 kind::f16)', 'data': 'synthetic',
            'config': {'workload': desc, 'tiles': n_tiles, 'mirror_passes': 8, 'tiles_per_batch': pred.last_tiles_per_batch,
                       'l2': f'inputs larger than L2 (volume {host.numel() * 4 / 1e6:.0f} MB, activations '
                             f'{prog.activation_elements() * 2 * pred.last_tiles_per_batch * 8 / 1e9:.1f} GB per launch sequence)',
                       'parallelism': f'x-slab tile sharding over {world} GPU(s), one halo exchange'},
            'e2e': {'value': nvox / (e2e_ms * 1e-3) / 1e6, 'unit': 'Mvoxel/s', 'sec_per_volume': e2e_ms * 1e-3,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'call': 'predict_logits_from_preprocessed_data(host tensor) -> host fp16 logits' if world == 1
                    else 'host volume -> predict_sliding_window_sharded -> host uint8 label map'},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'roofline_network': roof_all,
            'roofline_aggregation': roof_mem,
            'phase_ms_per_volume': phase, 'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--tiles-per-batch', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == '__main__':
    main()
