"""Per-launch device times of the detector network alone (no pipeline): vbt_detect on a fixed random
batch with per-op events (vbt_model_profile), plus the time of the graph-replayed network.
usage: python scripts/net_probe.py [variant] [batch] [iters] [filter]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vbt_b200 import effdet
from vbt_b200.interpreter import Detector

variant = sys.argv[1] if len(sys.argv) > 1 else 'lite0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
flt = sys.argv[4] if len(sys.argv) > 4 else ''
g = effdet.build_synthetic(variant)
det = Detector(g, max_batch=B)
x = torch.randint(0, 256, (B, g.S, g.S, 3), dtype=torch.uint8, device='cuda')
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3):
        det.network(x)
    st.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        det.network(x)
    b.record()
    st.synchronize()
    print(f'network (graph replay): {a.elapsed_time(b) / iters * 1e3:.1f} us per {B} frames')
    det.profile(True)
    for _ in range(iters):
        det.network(x)
    st.synchronize()
    ms, calls = det.op_times()
    det.profile(False)
plan, kinds = det.plan(), det.plan_kinds()
tot = {}
for i, op in enumerate(g.ops):
    if plan[i] == 0:
        continue
    us = 1e3 * ms[i] / max(calls, 1)
    kind = 'mbconv' if kinds[i] == 1 else ('node' if plan[i] > 1 else {1: 'stem', 2: 'pw', 3: 'dw', 4: 'add', 5: 'pool'}[op.type])
    tot[kind] = tot.get(kind, 0.0) + us
    if flt and (flt in kind or flt in op.name):
        t = g.tensors[op.inputs[0]]
        print(f'{i:4d} {kind:7s} {op.name:16s} {t.h}x{t.w}x{t.c:<5d} {us:8.2f} us')
print('per-class us:', {k: round(v, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}, 'sum', round(sum(tot.values()), 1))
