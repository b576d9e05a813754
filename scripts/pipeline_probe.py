#!/usr/bin/env python
"""Where does a pipeline step go?  Lite0, frame batch 64, frames resident in HBM:
(a) detection side only (K1 -> network -> K6 -> pack, two lanes), (b) the same with the tracker /
velocity stream attached, (c) the tracker stream alone on the recorded detection tables.
usage (on a B200): python scripts/pipeline_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vbt_b200 import effdet
from vbt_b200.interpreter import Detector
from vbt_b200.pipeline import VideoPipeline
from vbt_b200.synth import plate_trajectory, render_clip

B, N = 64, 28
g = effdet.build_synthetic('lite0')
det = Detector(g, max_batch=B)
clip = render_clip(B * N, 1080, 1920, seed=0, device='cuda', trajectory=plate_trajectory(B * N, 30.0, seed=0))
numbers = torch.arange(1, B * N + 1, dtype=torch.int32, device='cuda')


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


pipe = VideoPipeline(det, 30.0, 0.5, row_cap=1 << 17)


def detect_only():
    for b in range(N):
        pipe.process(clip[b * B:(b + 1) * B], numbers[b * B:(b + 1) * B], track=False)
    pipe._sync_streams()


def full():
    pipe.reset()
    for b in range(N):
        pipe.process(clip[b * B:(b + 1) * B], numbers[b * B:(b + 1) * B])
    pipe._sync_streams()


detect_only(); table = pipe.detection_table()
a = timed(detect_only) / N
table = pipe.detection_table()
full()
b_ = timed(full) / N


def track_only():
    pipe.reset()
    pipe.track_table(*table)
    pipe._sync_streams()


c = timed(track_only) / N
print(f'detection side only : {a:.3f} ms per 64-frame step ({B / a * 1e3:.0f} frames/s)')
print(f'full pipeline       : {b_:.3f} ms per step ({B / b_ * 1e3:.0f} frames/s)')
print(f'tracker stream alone: {c:.3f} ms per step ({B / c * 1e3:.0f} frames/s)')
