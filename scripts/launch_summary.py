#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= mi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', '')
        v = float(r[mi].replace(',', ''))
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[ui], 1.0)
        agg[name][0] += 1
        agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised)')
    print(f'{"kernel":60s} {"n":>6s} {"total_us":>10s} {"share":>7s} {"avg_us":>9s}')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{k[:60]:60s} {v[0]:6d} {v[1]:10.1f} {v[1] / tot:7.3f} {v[1] / v[0]:9.2f}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
