"""Runs the detector network a few times on a fixed random batch, launch by launch (for ncu: VBT_GRAPH=0).
usage: python scripts/net_once.py [variant] [batch] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vbt_b200 import effdet
from vbt_b200.interpreter import Detector

variant = sys.argv[1] if len(sys.argv) > 1 else 'lite0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = effdet.build_synthetic(variant)
det = Detector(g, max_batch=B)
x = torch.randint(0, 256, (B, g.S, g.S, 3), dtype=torch.uint8, device='cuda')
for _ in range(iters):
    det.network(x)
torch.cuda.synchronize()
print('done')
