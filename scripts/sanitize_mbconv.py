"""A few fused-MBConv / stem-block / node micro cases for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_mbconv.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import micrograph as MG
from vbt_b200.interpreter import Detector
cases = [(20, 20, 24, 144, 24, 3, 1, True, 2), (17, 13, 24, 144, 40, 5, 2, False, 3), (10, 10, 192, 1152, 192, 5, 1, True, 3),
         (33, 31, 16, 96, 24, 3, 2, False, 1), (10, 10, 192, 1152, 320, 3, 1, False, 2), (1, 1, 16, 96, 16, 3, 1, True, 2)]
for (h, w, cin, cexp, cout, k, s, res, B) in cases:
    g = MG.mbconv_graph(h, w, cin, cexp, cout, k, s, residual=res, seed=1)
    _, xp = MG.random_input(g, B, 1)
    det = Detector(g, max_batch=B)
    det.network(torch.as_tensor(np.ascontiguousarray(xp), device='cuda').view(torch.uint8))
    torch.cuda.synchronize()
for (h, w, cout, k, s, B) in [(64, 64, 16, 3, 1, 2), (97, 71, 16, 3, 1, 2), (33, 33, 16, 5, 1, 2)]:
    g = MG.stem_block_graph(h, w, cout, k, s, seed=2)
    det = Detector(g, max_batch=B)
    det.network(torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device='cuda'))
    torch.cuda.synchronize()
# a split (cluster) launch: 10x10 at batch 64
g = MG.mbconv_graph(10, 10, 192, 1152, 192, 5, 1, residual=True, seed=3)
_, xp = MG.random_input(g, 64, 1)
det = Detector(g, max_batch=64)
det.network(torch.as_tensor(np.ascontiguousarray(xp), device='cuda').view(torch.uint8))
torch.cuda.synchronize()
print('sanitize cases done')
