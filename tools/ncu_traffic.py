#!/usr/bin/env python
"""Reads `ncu --set full` reports and writes the per-launch DRAM traffic of each captured kernel into
profiles/r02_ncu_traffic.json (what bench.py's roofline.traffic quotes) plus a text summary.
    python tools/ncu_traffic.py gpurun_out/x.ncu-rep:op_name:patches:cin:cout [more...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def read(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = float(r[i].replace(',', ''))
                if units[i] in UNIT:
                    v *= UNIT[units[i]]
                d[k] = v
                d[k + ' [unit]'] = units[i]
        res.append(d)
    return res


def main():
    caps, lines = [], []
    for arg in sys.argv[1:]:
        rep, op, patches, cin, cout = (arg.split(':') + ['', '0', '0', '0'])[:5]
        for d in read(rep):
            name = d['kernel'].split('(')[0].split('::')[-1].split('<')[0].strip()
            caps.append({'kernel': name, 'kernel_full': d['kernel'][:120], 'op': op, 'patches_per_launch': int(patches), 'cin': int(cin), 'cout': int(cout),
                         'dram_bytes_read': d.get('dram__bytes_read.sum'), 'dram_bytes_write': d.get('dram__bytes_write.sum'),
                         'report': os.path.basename(rep)})
            lines.append(f"{os.path.basename(rep)}: {d['kernel'][:100]}")
            for k in KEYS:
                if k in d:
                    lines.append(f'    {k:75s} {d[k]:.6g} {d[k + " [unit]"] if d[k + " [unit]"] not in UNIT else "byte"}')
    path = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
    old = {'captures': []}
    if os.path.exists(path):
        old = json.load(open(path))
    keep = [c for c in old['captures'] if (c['kernel'], c['op'], c['patches_per_launch']) not in
            {(n['kernel'], n['op'], n['patches_per_launch']) for n in caps}]
    json.dump({'captures': keep + caps}, open(path, 'w'), indent=1)
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
