"""Cycle counters of CTA (0,0) of the fused MBConv kernel for every Lite0 block shape at batch 64
(run with VBT_MB_DBG=1 VBT_GRAPH=0): prints the tile geometry the chooser picked and cycles per phase."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import numpy as np, torch
import micrograph as MG
from vbt_b200.interpreter import Detector
BLOCKS = [  # h, w, cin, cexp, cout, k, s, res, expand (the Lite0 block shapes, as scripts/mbconv_sweep.py)
    (160, 160, 32, 32, 16, 3, 1, False, False), (160, 160, 16, 96, 24, 3, 2, False, True),
    (80, 80, 24, 144, 24, 3, 1, True, True), (80, 80, 24, 144, 40, 5, 2, False, True),
    (40, 40, 40, 240, 40, 5, 1, True, True), (40, 40, 40, 240, 80, 3, 2, False, True),
    (20, 20, 80, 480, 80, 3, 1, True, True), (20, 20, 80, 480, 112, 5, 1, False, True),
    (20, 20, 112, 672, 112, 5, 1, True, True), (20, 20, 112, 672, 192, 5, 2, False, True),
    (10, 10, 192, 1152, 192, 5, 1, True, True), (10, 10, 192, 1152, 320, 3, 1, False, True)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (h, w, cin, cexp, cout, k, s, res, ex) in BLOCKS:
    g = MG.mbconv_graph(h, w, cin, cexp, cout, k, s, residual=res, seed=1, expand=ex)
    _, xp = MG.random_input(g, B, 1)
    det = Detector(g, max_batch=B)
    dev = torch.as_tensor(np.ascontiguousarray(xp), device='cuda').view(torch.uint8)
    det.network(dev); torch.cuda.synchronize()
