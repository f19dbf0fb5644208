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

# `roofline.traffic` = dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture of the dominant
# kernel (same kernel, same patches per launch), written by tools/ncu_traffic.py into profiles/; null when the
# committed capture does not match the kernel this run measured.
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
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
    """Which tcgen05 kernel the engine picks for a stride-1 conv (mirrors plan_zrows / plan_rows in csrc/): the z-pair
    row-streaming kernel for 3x3x3, Cin in {16, 32}, Cout <= 16, even depth, 64 <= W <= 128; the ky-folded
    row-streaming kernel for Cout <= 32 otherwise; else the general implicit GEMM."""
    cout_pad = (dom['cout'] + 15) // 16 * 16
    d, _, w = dom['out_dims']
    k = tuple(dom.get('kernel', (3, 3, 3)))
    thin = dom['cin'] in (16, 32) and 64 <= w <= 128 and k[1:] == (3, 3)
    if thin and k == (3, 3, 3) and cout_pad == 16 and d % 2 == 0 and os.environ.get('FNNU_ZROWS', '1') != '0':
        return 'conv_umma_zrows_kernel'
    if thin and cout_pad in (16, 32):
        return 'conv_umma_rows_kernel'
    return 'conv_umma_kernel'


def _ncu_traffic(kernel, patches, op_name):
    """Bytes of DRAM traffic per launch from the committed ncu capture of the SAME kernel and patch count."""
    try:
        d = json.load(open(TRAFFIC_FILE))
    except Exception:
        return None
    for e in d.get('captures', []):
        if e.get('kernel') == kernel and e.get('op') == op_name and e.get('patches_per_launch') == patches:
            return float(e['dram_bytes_read'] + e['dram_bytes_write'])
    return None


def torch_gpu_baseline(wl, dev, n_tiles_sample=6):
    """The reference's own CUDA path as BASELINE.md section 3 describes it: the network under
    torch.autocast('cuda') fp16 with cudnn.benchmark (predict_from_raw_data.py:60, :648), batch 1, eight sequential
    mirror passes per tile (:541-557), Gaussian-weighted fp16 accumulation (:587-620) — the oracle's restatement of that
    loop, run on THIS GPU over a sub-volume holding `n_tiles_sample` tiles, extrapolated by tile count."""
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
    net.eval().to(dev)
    n_total = len(sw.tile_starts(vol[1:], patch, 0.5))
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    x = synth_volume((vol[0], patch[0], patch[1], patch[2] + (n_tiles_sample - 1) * (patch[2] // 2)), 0).to(dev)
    try:
        for _ in range(2):      # warm-up incl. cuDNN autotuning
            OP.predict_sliding_window_return_logits(net, x, patch, 0.5, True, (0, 1, 2), autocast_device='cuda')
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        OP.predict_sliding_window_return_logits(net, x, patch, 0.5, True, (0, 1, 2), autocast_device='cuda')
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        torch.backends.cudnn.benchmark = old
        del net, x
        torch.cuda.empty_cache()
    spv = dt / n_tiles_sample * n_total
    return {'sec_per_volume': spv, 'value': float(np.prod(vol[1:])) / spv / 1e6, 'unit': 'Mvoxel/s',
            'what': "oracle restatement of the reference's CUDA path: torch.autocast('cuda') fp16, cudnn.benchmark=True, "
                    'batch 1, 8 sequential mirror passes per tile, fp16 accumulators, on this GPU',
            'sample': f'{n_tiles_sample} of {n_total} tiles ({dt:.2f} s), extrapolated by tile count'}


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

    # ---- e2e: HOST volume in (pinned) -> HOST label map out, the SAME call at every N (each rank uploads only the
    # planes its tiles read; rank 0 downloads the uint8 label map through a pinned buffer); at N = 1 additionally the
    # reference-facing logits call (host fp16 logits out) as `e2e_logits`
    def e2e_step():
        lab = pred.predict_sliding_window_sharded(host, return_labels=True, gather_to=0)
        if rank == 0:
            return pred.to_host(lab)
        return None

    e2e_step()
    sync_all()
    n_e2e = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    h2d = host.numel() * host.element_size()
    d2h = int(nvox) if rank == 0 else 0
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_logits = None
    if world == 1 and heads <= 8:
        pred.predict_logits_from_preprocessed_data(host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = pred.predict_logits_from_preprocessed_data(host)
        torch.cuda.synchronize()
        ms_l = (time.perf_counter() - t0) / n_e2e * 1e3
        e2e_logits = {'value': nvox / (ms_l * 1e-3) / 1e6, 'unit': 'Mvoxel/s', 'sec_per_volume': ms_l * 1e-3,
                      'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': res.numel() * res.element_size(),
                      'call': 'predict_logits_from_preprocessed_data(host tensor) -> host fp16 logits (pageable, as the '
                              'reference returns them)'}
        del res

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
                'traffic': _ncu_traffic(_conv_kernel_name(dom), dom['patches_per_launch'], dom['op']),
                'algorithmic_bytes': dom['algorithmic_bytes_per_launch'],
                'peak_source': peaks['source'] + ' (sustained cuBLAS bf16; fp16 runs at the same tcgen05 rate)',
                'kernel': _conv_kernel_name(dom) + ' (' + dom['op'] + f", Conv3d {dom['cin']}->{dom['cout']} @{dom['out_dims']}"
                          f", {dom['patches_per_launch']} patches per launch)",
                'ms_per_launch': dom['ms'], 'flop_per_launch': dom['flop_per_launch']}
        roof['frac'] = roof['achieved'] / roof['peak']
        # the other kernels with a committed ncu --set full capture (profiles/r02_ncu_traffic.json): live device time of
        # the same operator, against whichever of the two roofs bounds it
        roof_more = []
        try:
            for cap in json.load(open(TRAFFIC_FILE)).get('captures', []):
                if cap.get('patches_per_launch') != dom['patches_per_launch'] or cap.get('op') == dom['op']:
                    continue
                try:
                    k = pred.profile_dominant_op(data, n_batches=3, op_name=cap['op'])
                except KeyError:
                    continue
                if (k['cin'], k['cout']) != (cap.get('cin'), cap.get('cout')):
                    continue          # same operator name in another network (teacher, ResEnc): not the captured kernel
                t = k['ms'] * 1e-3
                f_tensor = k['flop_per_launch'] / t / 1e12 / peaks['bf16_tflops_sustained']
                f_hbm = k['algorithmic_bytes_per_launch'] / t / 1e9 / peaks['hbm_gbs']
                hbm = f_hbm >= f_tensor
                roof_more.append({'kernel': cap['kernel'], 'op': k['op'], 'bound': 'hbm' if hbm else 'tensor',
                                  'unit': 'GB/s' if hbm else 'TFLOP/s',
                                  'achieved': (k['algorithmic_bytes_per_launch'] / t / 1e9) if hbm else (k['flop_per_launch'] / t / 1e12),
                                  'peak': peaks['hbm_gbs'] if hbm else peaks['bf16_tflops_sustained'],
                                  'frac': f_hbm if hbm else f_tensor,
                                  'traffic': float(cap['dram_bytes_read'] + cap['dram_bytes_write']),
                                  'algorithmic_bytes': k['algorithmic_bytes_per_launch'], 'ms_per_launch': k['ms'],
                                  'patches_per_launch': k['patches_per_launch'], 'ncu_report': cap.get('report')})
        except Exception as ex:      # evidence only: never lose the bench line over it
            roof_more = {'error': repr(ex)}
        roof_all = {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': peaks['bf16_tflops_sustained'],
                    'achieved': (flops_vol / world / (conv_ms * 1e-3) / 1e12) if conv_ms else None,
                    'kernel': 'every launch of the network forward (convs, transposed convs, first layer, seg head)',
                    'flop_per_forward': flops_fwd}
        roof_all['frac'] = roof_all['achieved'] / roof_all['peak'] if roof_all['achieved'] else None
        roof_mem = {'bound': 'hbm', 'unit': 'GB/s', 'peak': peaks['hbm_gbs'],
                    'achieved': (acc_bytes / (acc_ms * 1e-3) / 1e9) if acc_ms else None, 'traffic': None,
                    'kernel': 'accumulate_cluster_kernel (TTA mean + Gaussian weight + accumulate, one launch per group of '
                              'overlapping tiles)', 'bytes_per_tile': acc_bytes * world / n_tiles}
        roof_mem['frac'] = roof_mem['achieved'] / roof_mem['peak'] if roof_mem['achieved'] else None
        gpu_ref = None
        if world == 1 and not args.no_torch_gpu_baseline:
            try:
                gpu_ref = torch_gpu_baseline(wl, dev)
                gpu_ref['ours_over_torch_gpu'] = gpu_ref['sec_per_volume'] / (ms * 1e-3)
            except Exception as e:        # a baseline, never a reason to lose the measurement
                gpu_ref = {'unavailable': repr(e)[:200]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            spv, dt, n_total = cpu_reference_sample(wl, CPU_BASELINE_TILES, cores)
            cpu = {'value': nvox / spv / 1e6, 'unit': 'Mvoxel/s', 'cores': cores, 'kind': 'port',
                   'sec_per_volume': spv,
                   'sample': f'{CPU_BASELINE_TILES} of {n_total} tiles x 8 mirror passes ({dt:.1f} s of CPU work), '
                             f'extrapolated by tile count'}
            # the reference caps its own CPU run at default_num_processes = 8 threads (predict_from_raw_data.py:479-480):
            # the same sample, smaller, with that cap (SURVEY.md section 8d asks for both)
            try:
                capped = min(8, cores)
                spv8, dt8, _ = cpu_reference_sample(wl, 2, capped)
                cpu['reference_default_threads'] = {'cores': capped, 'value': nvox / spv8 / 1e6, 'unit': 'Mvoxel/s',
                                                    'sec_per_volume': spv8,
                                                    'sample': f'2 of {n_total} tiles x 8 mirror passes ({dt8:.1f} s of CPU work)'}
                torch.set_num_threads(cores)
            except Exception as ex:
                cpu['reference_default_threads'] = {'error': repr(ex)}
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
                    'call': 'pinned host volume -> predict_sliding_window_sharded(return_labels) -> pinned host uint8 label '
                            'map (same call at every N)'},
            'e2e_logits': e2e_logits, 'torch_gpu_baseline': gpu_ref, 'roofline_kernels': roof_more,
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
    ap.add_argument('--no-torch-gpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == '__main__':
    main()
