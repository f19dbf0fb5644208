"""Builds the per-layer table of profiles/README.md from an ncu launch list of `bench.py --workload cfg1`
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): one launch sequence = gather, the 28 network
launches of the distilled student (feature_reduction_factor 2, 128^3 patches, 32 patches per launch), the 4 accumulate
rounds.  Usage: python profiles/make_launch_table.py profiles/r01_launches_final.csv"""
import csv
import sys

FEATS = [16, 32, 64, 128, 160, 160]
PATCHES = 32


def student_layers():
    """(name, GFLOP for 32 patches) in launch order; mirrors fast_nnunet_b200/program.py's lowering."""
    out = []
    dims = [128 >> i for i in range(6)]
    cin = 1
    for s, f in enumerate(FEATS):
        d = dims[s]
        stride = '' if s == 0 else ' stride 2'
        out.append((f'enc{s}.0 {cin}->{f}{stride or " @%d^3" % d}', 2.0 * PATCHES * d ** 3 * 27 * cin * f / 1e9))
        out.append((f'enc{s}.1 {f}->{f} @{d}^3', 2.0 * PATCHES * d ** 3 * 27 * f * f / 1e9))
        cin = f
    for k in range(1, 6):
        s = 5 - k                      # decoder stage k writes resolution level s
        f_below, f = FEATS[s + 1], FEATS[s]
        d = dims[s]
        out.append((f'up{k} {f_below}->{f} (transposed)', 2.0 * PATCHES * d ** 3 * f_below * f / 1e9))
        out.append((f'dec{k}.0 {2 * f}->{f} @{d}^3', 2.0 * PATCHES * d ** 3 * 27 * 2 * f * f / 1e9))
        out.append((f'dec{k}.1 {f}->{f} @{d}^3', 2.0 * PATCHES * d ** 3 * 27 * f * f / 1e9))
    out.append(('seg head 16->2 (1x1x1)', 2.0 * PATCHES * 128 ** 3 * 16 * 2 / 1e9))
    return out


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    H = rows[hdr]
    ki, vi = H.index('Kernel Name'), H.index('Metric Value')
    L = [(r[ki].split('(')[0].replace('void ', '').replace('fnnu::', ''), float(r[vi]) / 1000.0) for r in rows[hdr + 1:]]
    start = [i for i, l in enumerate(L) if 'gather' in l[0]][0]
    seq = L[start:start + 33]
    names = [('gather (4 tiles x 8 mirror copies)', None)] + student_layers() + [(f'accumulate round {i}', None) for i in range(1, 5)]
    assert len(names) == len(seq), (len(names), len(seq))
    total = sum(t for _, t in seq)
    gf_total = sum(g for _, g in names if g)
    print('| # | layer | kernel | time (us) | GFLOP (x32 patches) | TFLOP/s | share |')
    print('|---|---|---|---|---|---|---|')
    for i, ((name, gf), (kern, us)) in enumerate(zip(names, seq)):
        gfs = f'{gf:.0f}' if gf else '-'
        tf = f'{gf / us * 1e3:.0f}' if gf else '-'     # GFLOP / us = PFLOP/s -> x1000 = TFLOP/s
        print(f'| {i} | {name} | `{kern}` | {us:.0f} | {gfs} | {tf} | {100 * us / total:.1f} % |')
    print(f'| | **one launch sequence = 4 tiles x 8 mirror copies = 32 patches** | | **{total:.0f}** | {gf_total:.0f} | '
          f'**{gf_total / total * 1e3:.0f}** | 100 % |')


if __name__ == '__main__':
    main(sys.argv[1])
