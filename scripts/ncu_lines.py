#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line.
usage: ncu_lines.py report.ncu-rep [kernel-index] [top]"""
import csv
import collections
import subprocess
import sys


def main(rep, which=0, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'File Path']
    starts.append(len(rows))
    s, e = starts[which], starts[which + 1]
    print('#', rows[s + 1][1][:120])
    hdr = rows[s + 2]
    li, si = 0, 1
    samp = hdr.index('# Samples')
    inst = hdr.index('Instructions Executed')
    agg = collections.OrderedDict()
    cur = None
    for r in rows[s + 3:e]:
        if len(r) <= samp:
            continue
        if r[li]:
            cur = (int(r[li]), r[si])
            agg.setdefault(cur, [0, 0])
        if cur is None:
            continue
        try:
            agg[cur][0] += int(r[samp] or 0)
            agg[cur][1] += int(r[inst] or 0)
        except ValueError:
            pass
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f'# samples {tot}, warp instructions {toti}')
    for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'{ln:5d} {v[0] / tot:6.3f} {v[1] / toti:6.3f}  {src.strip()[:110]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
