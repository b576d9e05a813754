#!/usr/bin/env python
"""How much do two detection lanes overlap?  Times vbt_detect (CUDA-graph replay of the layer
program) for Lite0 at frame batch 64: one stream back to back, then two streams / two detectors
concurrently.  usage (on a B200): python scripts/lane_overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vbt_b200 import effdet
from vbt_b200.interpreter import Detector

g = effdet.build_synthetic('lite0')
B, N = 64, 60
dets = [Detector(g, max_batch=B) for _ in range(3)]
x = torch.randint(0, 256, (B, g.S, g.S, 3), dtype=torch.uint8, device='cuda')
streams = [torch.cuda.Stream() for _ in range(3)]


def run(n_lanes):
    for _ in range(4):
        for i in range(n_lanes):
            with torch.cuda.stream(streams[i]):
                dets[i].network(x, stream=streams[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:n_lanes]:
        s.wait_event(e0)
    for k in range(N):
        i = k % n_lanes
        with torch.cuda.stream(streams[i]):
            dets[i].network(x, stream=streams[i])
    for s in streams[:n_lanes]:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N


for n in (1, 2, 3):
    ms = run(n)
    print(f'lanes={n}: {ms:.3f} ms per 64-frame batch (network only) = {B / ms * 1e3:.0f} frames/s')
