"""Times a few fused MBConv block shapes at batch 64 (graph replay, 20 iterations each).
usage: python scripts/mb_one.py [indices into BLOCKS ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import micrograph as MG
from vbt_b200.interpreter import Detector
BLOCKS = [(160, 160, 32, 32, 16, 3, 1, False, False), (160, 160, 16, 96, 24, 3, 2, False, True),
    (80, 80, 24, 144, 24, 3, 1, True, True), (80, 80, 24, 144, 40, 5, 2, False, True),
    (40, 40, 40, 240, 40, 5, 1, True, True), (40, 40, 40, 240, 80, 3, 2, False, True),
    (20, 20, 80, 480, 80, 3, 1, True, True), (20, 20, 80, 480, 112, 5, 1, False, True),
    (20, 20, 112, 672, 112, 5, 1, True, True), (20, 20, 112, 672, 192, 5, 2, False, True),
    (10, 10, 192, 1152, 192, 5, 1, True, True), (10, 10, 192, 1152, 320, 3, 1, False, True)]
idx = [int(v) for v in sys.argv[1:]] or list(range(len(BLOCKS)))
out = []
for i in idx:
    (h, w, cin, cexp, cout, k, s, res, ex) = BLOCKS[i]
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
        for _ in range(20):
            det.network(dev)
        b.record()
        st.synchronize()
    out.append(a.elapsed_time(b) * 50.0)
print(' '.join(f'{v:7.1f}' for v in out), f'sum {sum(out):.1f}')
