#!/usr/bin/env python
"""Per-kernel DRAM traffic and device time from an `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv` log of the bench; writes a text table and the
JSON bench.py reads for `roofline.traffic` (bytes per launch, averaged over the launches of the
kernel class in the captured steps)."""
import collections
import csv
import json
import re
import sys

CLASS = [('mbconv_umma', 'mbconv_fused'), ('node_umma', 'node_fused'), ('pw_persist', 'pw'), ('pw_umma', 'pw'), ('pw_dp4a', 'pw_head_out'), ('dw_umma', 'dw5'), ('dw_kernel', 'dw3'), ('stem', 'stem'),
         ('add_kernel', 'fuse_add'), ('preprocess', 'K1_preprocess'), ('postprocess', 'K6_postprocess'),
         ('tracker_update', 'K7_tracker'), ('velocity_update', 'K8_velocity'), ('pack_detections', 'pack')]
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3,
         'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}


def main(path, out_json=None):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, ni, ui, vi, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
    per = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', '')
        d = per.setdefault(r[ii], {'name': name})
        d[r[ni]] = float(r[vi].replace(',', '')) * SCALE.get(r[ui], 1.0)
    agg = collections.OrderedDict()
    for d in per.values():
        cls = next((c for pat, c in CLASS if pat in d['name']), None)
        if cls is None:
            continue
        a = agg.setdefault(cls, {'launches': 0, 'us': 0.0, 'read': 0.0, 'write': 0.0})
        a['launches'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['read'] += d.get('dram__bytes_read.sum', 0.0)
        a['write'] += d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a['us'] for a in agg.values())
    print(f'# {path}: device time and DRAM traffic per kernel class (cold cache, serialised launches)')
    print(f'{"class":16s} {"launches":>8s} {"total_us":>10s} {"share":>6s} {"dram_MB":>10s} {"MB/launch":>10s} {"GB/s":>8s}')
    out = {}
    for c, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        b = a['read'] + a['write']
        print(f'{c:16s} {a["launches"]:8d} {a["us"]:10.1f} {a["us"] / tot:6.3f} {b / 1e6:10.1f} '
              f'{b / 1e6 / a["launches"]:10.2f} {b / 1e3 / a["us"]:8.1f}')
        out[c] = b / a['launches']
    if out_json:
        with open(out_json, 'w') as f:
            json.dump({k: round(v) for k, v in out.items()}, f, indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
