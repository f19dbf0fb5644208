#!/usr/bin/env python
"""Warp-stall samples of one captured kernel, summed per CUDA source line (ncu --set full --import-source on):
    python tools/ncu_stalls.py report.ncu-rep [source-file-substring] [top]
Reads `ncu -i report --page source --csv --print-source cuda,sass` and prints the lines that hold most samples,
split by stall reason, so that a profile summary under profiles/ can name the code that waits."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per_line = defaultdict(lambda: defaultdict(float))
hdr, col, key_all, stall_cols, cur, path = None, {}, None, [], None, ''
for r in rows:
    if r and r[0] == 'File Path':
        path = r[1].rsplit('/', 1)[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        col = {c: i for i, c in enumerate(hdr)}
        key_all = col['Warp Stall Sampling (All Samples)']
        stall_cols = [(c, i) for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != '':
        cur = f'{path}:{r[0]}  {r[1].strip()}'     # a CUDA source line introduces the SASS that follows it
        continue
    try:
        n = float(r[key_all] or 0)
    except ValueError:
        continue
    k = cur or '?'
    per_line[k]['all'] += n
    for c, i in stall_cols:
        try:
            per_line[k][c] += float(r[i] or 0)
        except ValueError:
            pass
if hdr is None:
    sys.exit('no source page in ' + rep)
total = sum(v['all'] for v in per_line.values()) or 1.0
print(f'{rep}: {int(total)} samples')
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1]['all'])[:top]:
    reasons = sorted(((c, x) for c, x in v.items() if c != 'all' and x > 0), key=lambda t: -t[1])[:3]
    rs = ', '.join(f'{c[6:]} {x / total * 100:.1f}%' for c, x in reasons)
    print(f'{v["all"] / total * 100:5.1f}%  {k[:110]}   [{rs}]')
