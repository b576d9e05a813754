import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import micrograph as MG
from vbt_b200.interpreter import Detector
for (h,w,cin,cexp,cout,k,s,res) in [(10,10,192,1152,192,5,1,True),(160,160,16,96,24,3,2,False),(40,40,40,240,40,5,1,True),(20,20,112,672,112,5,1,True)]:
    g = MG.mbconv_graph(h,w,cin,cexp,cout,k,s,residual=res,seed=1)
    x, xp = MG.random_input(g, 64, 1)
    det = Detector(g, max_batch=64)
    dev = torch.as_tensor(np.ascontiguousarray(xp), device='cuda')
    det.network(dev.view(torch.uint8)); torch.cuda.synchronize()
