"""Synthetic inputs (there is no network for datasets or checkpoints; the reference's
.tflite weights and videos are not in its checkout -- SURVEY.md F3).

* ``synthetic_model_inputs`` -- small uint8 frames at the model's input size, used to
  calibrate synthetic models and in parity tests (numpy, host).
* ``plate_trajectory`` / ``render_clip`` -- a 1080p clip whose plates follow a lift-like
  vertical oscillation; rendered ON THE DEVICE with torch (noise background + filled
  ellipses), the workload SURVEY.md 8(d) prescribes for the throughput bench.
"""
from __future__ import annotations

import numpy as np


def synthetic_model_inputs(n, size, seed=0):
    """uint8 [n,size,size,3]: uniform noise with 1-3 filled ellipses per frame."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n):
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0.2, 0.8, 2) * size
            ry, rx = rng.uniform(0.05, 0.2, 2) * size
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            frames[i][m] = rng.integers(0, 256, 3, dtype=np.uint8)
    return frames


def plate_trajectory(n_frames, fps=30.0, reps=6, seed=0):
    """Normalised (x, y, w, h) of one plate per frame: `reps` squat-like repetitions."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_frames) / fps
    period = (n_frames / fps) / (reps + 1)
    phase = np.clip((t - period / 2) / period, 0, reps)
    y = 0.35 + 0.2 * (1 - np.cos(2 * np.pi * phase)) / 2 + rng.normal(0, 0.001, n_frames)
    x = 0.5 + rng.normal(0, 0.001, n_frames)
    w = np.full(n_frames, 0.22)
    h = np.full(n_frames, 0.39)       # 0.22 * 1920 px wide == 0.39 * 1080 px tall
    return np.stack([x, y, w, h], axis=1)


def render_clip(n_frames, height=1080, width=1920, seed=0, device='cuda', chunk=32,
                trajectory=None):
    """uint8 [n_frames,height,width,3] BGR clip on `device` (torch), seed-deterministic:
    a static uniform-noise scene (fixed camera) with one dark plate moving over it."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    traj = plate_trajectory(n_frames, seed=seed) if trajectory is None else trajectory
    out = torch.empty((n_frames, height, width, 3), dtype=torch.uint8, device=device)
    yy = torch.arange(height, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(width, device=device, dtype=torch.float32)[None, :]
    bg = torch.randint(0, 256, (height, width, 3), dtype=torch.uint8, device=device, generator=g)
    for s in range(0, n_frames, chunk):
        e = min(s + chunk, n_frames)
        out[s:e] = bg
        for i in range(s, e):
            x, y, w, h = (float(v) for v in traj[i])
            m = ((yy - y * height) / (h * height / 2)) ** 2 + \
                ((xx - x * width) / (w * width / 2)) ** 2 <= 1.0
            out[i][m] = torch.tensor([40, 40, 40], dtype=torch.uint8, device=device)
    return out
