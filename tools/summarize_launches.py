#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.

    python tools/summarize_launches.py gpurun_out/launches_c2.csv > profiles/r01_launches_c2.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iname, ival, igrid, iblk = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size',
                                                       'Block Size'))
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[iname]).replace('void ', '').replace('<unnamed>::', '')
        if name.startswith('at::') or 'elementwise_kernel' in name:
            name = 'torch: ' + name[:60]
        a = agg.setdefault(name, [0, 0.0, r[igrid], r[iblk]])
        a[0] += 1
        a[1] += float(r[ival])
    total = sum(a[1] for a in agg.values())
    print(f'# {path}: {len(rows) - 1} launches, {total/1e3:.1f} us total (cold-cache, serialised by ncu)')
    print(f'{"kernel":58s} {"n":>4s} {"avg us":>10s} {"share":>7s}  grid / block')
    for name, (n, t, grid, blk) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{name[:58]:58s} {n:4d} {t/n/1e3:10.2f} {t/total:7.1%}  {grid} / {blk}')


if __name__ == '__main__':
    main(sys.argv[1])
