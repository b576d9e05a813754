"""Tile sweep of the fused MBConv kernel: every Lite0 block shape at batch 64, VBT_MB_TW x VBT_MB_TH
forced in turn (one subprocess per pair: the switches are read once per process); prints us per
block and the best pair.  usage: python scripts/mbconv_sweep.py [child TW TH]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
BLOCKS = [  # h, w, cin, cexp, cout, k, s, res, expand
    (160, 160, 32, 32, 16, 3, 1, False, False), (160, 160, 16, 96, 24, 3, 2, False, True),
    (80, 80, 24, 144, 24, 3, 1, True, True), (80, 80, 24, 144, 40, 5, 2, False, True),
    (40, 40, 40, 240, 40, 5, 1, True, True), (40, 40, 40, 240, 80, 3, 2, False, True),
    (20, 20, 80, 480, 80, 3, 1, True, True), (20, 20, 80, 480, 112, 5, 1, False, True),
    (20, 20, 112, 672, 112, 5, 1, True, True), (20, 20, 112, 672, 192, 5, 2, False, True),
    (10, 10, 192, 1152, 192, 5, 1, True, True), (10, 10, 192, 1152, 320, 3, 1, False, True)]

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np
    import torch
    import micrograph as MG
    from vbt_b200.interpreter import Detector
    out = []
    for (h, w, cin, cexp, cout, k, s, res, ex) in BLOCKS:
        g = MG.mbconv_graph(h, w, cin, cexp, cout, k, s, residual=res, seed=1, expand=ex)
        _, xp = MG.random_input(g, 64, 1)
        det = Detector(g, max_batch=64)
        dev = torch.as_tensor(np.ascontiguousarray(xp), device='cuda').view(torch.uint8)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                det.network(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                det.network(dev)
            b.record()
            st.synchronize()
        out.append(a.elapsed_time(b) * 100.0 if det.plan()[0] > 1 else -1.0)
    print('RES ' + ' '.join(f'{v:.1f}' for v in out))
    sys.exit(0)

pairs = [(0, 0)] + [(tw, th) for tw in (8, 12, 16, 20, 24, 32) for th in (4, 6, 8, 10, 16)]
best = [(1e9, None)] * len(BLOCKS)
for tw, th in pairs:
    env = dict(os.environ)
    if tw:
        env['VBT_MB_TW'], env['VBT_MB_TH'] = str(tw), str(th)
    r = subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith('RES ')]
    if not line:
        print(tw, th, 'failed', r.stderr[-300:])
        continue
    vals = [float(v) for v in line[0].split()[1:]]
    print(f'TW {tw:2d} TH {th:2d}: ' + ' '.join(f'{v:7.1f}' for v in vals), flush=True)
    for i, v in enumerate(vals):
        if 0 < v < best[i][0]:
            best[i] = (v, (tw, th))
for blk, (v, p) in zip(BLOCKS, best):
    print(blk, f'best {v:.1f} us at TW,TH = {p}')
