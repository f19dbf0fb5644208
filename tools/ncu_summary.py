#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` report for profiles/: launch shape, duration, DRAM / L2 / L1
traffic, pipe utilisation, and the CUDA source lines holding most warp-stall samples (tools/ncu_stalls.py).
    python tools/ncu_summary.py report.ncu-rep "what was captured" > profiles/r02_ncu_x.txt"""
import csv
import io
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2 sectors read by SMs'), ('lts__t_sectors_srcunit_tex_op_write.sum', 'L2 sectors written by SMs'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput % of peak'),
    ('l1tex__m_l1tex2xbar_write_bytes.sum', 'L1 -> L2 write bytes'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput % of peak'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (elapsed)'),
    ('sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'UTCHMMA fp16->fp32 % of peak'),
    ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of max'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
]
rep, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
print(f'# {os.path.basename(rep)}: {what}')
print('# ncu --set full --clock-control none --import-source on, one launch; times under the profiler are not bench values')
for r in rows[2:]:
    print('kernel:', r[h.index('Kernel Name')])
    for k, label in KEYS:
        if k in h:
            print(f'  {label:40s} {r[h.index(k)]:>18s} {u[h.index(k)]}')
print()
print('warp-stall samples per CUDA source line (top lines; reasons in % of all samples):')
st = subprocess.run([sys.executable, os.path.join(HERE, 'ncu_stalls.py'), rep, 'x', '18'], capture_output=True, text=True).stdout
print('\n'.join(line[:200] for line in st.splitlines()[1:]))
