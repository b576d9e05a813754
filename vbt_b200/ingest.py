"""Row-sparse ingest of host frames.

The reference resizes on the host (odt.preprocess_image, odt.py:10-19) and hands the model a
320x320 image; here the resize runs on the GPU, so the 1080p frame has to cross PCIe -- the
bus, not any kernel, is what bounds end-to-end throughput.  The bilinear
kernel without antialiasing reads only two source rows per output row, so only those rows are
sent: for 1080 -> 320 that is 16 of every 27 rows (59 % of the bytes).  The touched rows
repeat with period H / gcd(H, S), which turns "gather 640 scattered rows per frame" into
16 strided 2-D DMA copies per BATCH (`vbt_copy_rows_h2d`); K1 then reads the compacted row
table through a row map (`vbt_preprocess_rows_u8`).
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib


def touched_rows(in_size, out_size):
    """Sorted source rows the half-pixel bilinear resize reads (float32, as the kernel and
    tf.image.resize compute them)."""
    scale = np.float32(in_size) / np.float32(out_size)
    o = np.arange(out_size, dtype=np.float32)
    src = (o + np.float32(0.5)) * scale - np.float32(0.5)
    lo = np.maximum(np.floor(src).astype(np.int64), 0)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    return np.unique(np.concatenate([lo, hi]))


class RowSparseIngest:
    """Plans and performs the host->device transfer of the rows K1 needs."""

    def __init__(self, max_batch, H, W, S, n_buffers=3):
        self.torch = t = _lib.require_cuda()
        self.B, self.H, self.W, self.S = max_batch, H, W, S
        rows = touched_rows(H, S)
        period = H // math.gcd(H, S)
        in_period = rows[rows < period]
        tiled = (in_period[None, :] + period * np.arange(H // period)[:, None]).reshape(-1)
        if not np.array_equal(tiled, rows):          # edge clamping broke the pattern: one period
            period, in_period = H, rows
        self.period = int(period)
        self.rows_in_period = np.ascontiguousarray(in_period, dtype=np.int32)
        self.n_rows = len(in_period)
        self.rows_per_frame = (H // self.period) * self.n_rows
        row_map = np.full(H, -1, np.int32)
        for k in range(H // self.period):
            row_map[in_period + k * self.period] = k * self.n_rows + np.arange(self.n_rows)
        row_map[row_map < 0] = 0
        self.row_map = t.as_tensor(row_map, device='cuda')
        self.tables = [t.empty((max_batch, self.rows_per_frame, W * 3), dtype=t.uint8, device='cuda')
                       for _ in range(n_buffers)]
        self.free = [None] * n_buffers               # event: K1 has read table i
        self.copy_stream = t.cuda.Stream()
        self.cursor = 0
        self.bytes_per_frame = self.rows_per_frame * W * 3

    def upload(self, host_frames):
        """host_frames: pinned uint8 [n,H,W,3].  Starts the DMA on the ingest stream and returns
        (table, n, slot, ready event); the caller records `self.free[slot]` once K1 has consumed
        the table."""
        t = self.torch
        n = int(host_frames.shape[0])
        if not host_frames.is_pinned():
            raise ValueError('host frames must be in pinned memory (tensor.pin_memory())')
        assert n <= self.B and tuple(host_frames.shape[1:]) == (self.H, self.W, 3) and host_frames.is_contiguous()
        slot = self.cursor % len(self.tables)
        self.cursor += 1
        cs = self.copy_stream
        if self.free[slot] is not None:
            cs.wait_event(self.free[slot])
        _lib.check(_lib.lib().vbt_copy_rows_h2d(host_frames.data_ptr(), n, self.H, self.W, self.period,
                                                _lib.ptr(self.rows_in_period), self.n_rows,
                                                self.tables[slot].data_ptr(), cs.cuda_stream))
        ready = t.cuda.Event()
        ready.record(cs)
        return self.tables[slot], n, slot, ready
