#!/usr/bin/env python
"""SASS opcode mix (executed warp instructions) of one launch in an .ncu-rep.
usage: ncu_opmix.py report.ncu-rep [launch-index] [top]"""
import collections, csv, subprocess, sys
def main(rep, which=0, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    s, e = starts[which], starts[which + 1]
    print('#', rows[s][1][:100])
    hdr = rows[s + 1]
    ie, so = hdr.index('Instructions Executed'), hdr.index('Source')
    tot, byop = 0, collections.Counter()
    for r in rows[s + 2:e]:
        try:
            n = int(r[ie])
        except (ValueError, IndexError):
            continue
        toks = r[so].strip().split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        byop[op.split('.')[0]] += n
        tot += n
    print('# total warp instructions', tot)
    for k, v in byop.most_common(top):
        print(f'{k:12s} {v / tot:6.3f} {v:12d}')
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
