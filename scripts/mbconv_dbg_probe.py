"""Cycle counters of CTA (0,0) of the fused MBConv kernel for a few block shapes at batch 64
(run with VBT_MB_DBG=1 VBT_GRAPH=0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import micrograph as MG
from vbt_b200.interpreter import Detector
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'stem'):
    g = MG.stem_block_graph(320, 320, 16, 3, 1, seed=1)
    det = Detector(g, max_batch=64)
    x = torch.randint(0, 256, (64, 320, 320, 3), dtype=torch.uint8, device='cuda')
    det.network(x); torch.cuda.synchronize()
if which in ('all', 'blocks'):
    for (h,w,cin,cexp,cout,k,s,res) in [(10,10,192,1152,192,5,1,True),(160,160,16,96,24,3,2,False),(40,40,40,240,40,5,1,True),(20,20,112,672,112,5,1,True),(80,80,24,144,24,3,1,True)]:
        g = MG.mbconv_graph(h,w,cin,cexp,cout,k,s,residual=res,seed=1)
        x, xp = MG.random_input(g, 64, 1)
        det = Detector(g, max_batch=64)
        dev = torch.as_tensor(np.ascontiguousarray(xp), device='cuda')
        det.network(dev.view(torch.uint8)); torch.cuda.synchronize()
