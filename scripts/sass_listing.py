"""SASS evidence for profiles/: per kernel of libvbt_b200.so, how many tcgen05 / TMEM / TMA / bulk-copy
instructions it holds (cuobjdump -sass; no GPU needed).  usage: python scripts/sass_listing.py > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'vbt_b200', 'libvbt_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
KEYS = ['UTCIMMA', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCCP', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'ELECT', 'IDP', 'LDGSTS', 'UCGABAR', 'MAPA', 'BRA.U.ANY']
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]['_all'] += 1
    for k in KEYS:
        if op.startswith(k):
            counts[cur][k] += 1
            total[k] += 1
print('# cuobjdump -sass vbt_b200/libvbt_b200.so (sm_100a): instruction counts per kernel')
print('# UTCIMMA / UTCHMMA = tcgen05.mma kind::i8 / kind::f16; LDTM = tcgen05.ld (TMEM -> registers); UTCBAR = tcgen05.commit;')
print('# UBLKCP = cp.async.bulk (TMA engine, 1-D); LDGSTS = cp.async; IDP = dp4a; SYNCS = mbarrier ops; ELECT = elect.sync;')
print('# UCGABAR / MAPA = cluster barrier / distributed-shared-memory address mapping; BRA.U.ANY = uniformisation loops')
print(f'{"kernel":70s} {"instrs":>7s} ' + ' '.join(f'{k:>8s}' for k in KEYS))
for fn, c in counts.items():
    name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'\(.*', '', name)[:70]
    print(f'{name:70s} {c["_all"]:7d} ' + ' '.join(f'{c[k]:8d}' for k in KEYS))
print(f'{"TOTAL":70s} {sum(c["_all"] for c in counts.values()):7d} ' + ' '.join(f'{total[k]:8d}' for k in KEYS))
