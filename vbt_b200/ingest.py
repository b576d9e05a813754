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


class DecodeRing:
    """Video file -> ring of pinned host batches, filled by a background decode thread.

    replaces: the `cv2.VideoCapture` / `cap.read()` front of the reference loop (track.py:135-171).  The
    reference decodes a frame, converts it, runs the network, then decodes the next one; here a thread
    decodes straight INTO page-locked batch buffers (`VideoCapture.retrieve(dst)`: no staging copy)
    while the GPU works on earlier batches, frames the stride skips are only `grab()`-bed (demuxed, not
    colour-converted), and the consumer hands each full batch to `VideoPipeline.process`, whose
    row-sparse DMA (RowSparseIngest) moves just the rows K1 reads.  SURVEY.md 8f rank 2, the
    host-decode form: the GPU boxes of this pool expose no NVDEC user library (libnvcuvid is absent,
    profiles/r2_gpu_box_video_libs.txt), so on-GPU decode cannot be built or tested here.

    Frame numbering is the reference's: `frame_count` starts at 1 and counts every read attempt
    (track.py:160-161); a frame is kept when `frame_count % stride == 0` (track.py:166).

        ring = DecodeRing(path, batch=64, stride=16)
        for frames, numbers, slot in ring:           # frames: uint8 [n,H,W,3] BGR (pinned), n <= batch
            pipe.process(frames, numbers_on_device, swap_rb=True)
            ring.release(slot, pipe.input_consumed)   # event after which the buffer may be refilled
    """

    def __init__(self, src, batch=64, stride=1, n_slots=3, pin=None):
        import queue
        import threading
        import cv2
        import torch
        self.cap = cv2.VideoCapture(src)
        if not self.cap.isOpened():
            raise FileNotFoundError(src)
        self.fps = self.cap.get(cv2.CAP_PROP_FPS)
        self.W = int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.H = int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        self.batch, self.stride = int(batch), int(stride)
        pin = torch.cuda.is_available() if pin is None else pin
        self.slots = []
        for _ in range(n_slots):
            t = torch.empty((self.batch, self.H, self.W, 3), dtype=torch.uint8)
            self.slots.append(t.pin_memory() if pin else t)
        self._views = [s.numpy() for s in self.slots]        # the decoder writes through these
        self._free = queue.Queue()
        self._full = queue.Queue(maxsize=n_slots)
        for i in range(n_slots):
            self._free.put((i, None))
        self.frames_read = 0
        self._error = None
        self._thread = threading.Thread(target=self._decode, daemon=True)
        self._thread.start()

    def _decode(self):
        try:
            frame_count, n, numbers, slot = 0, 0, [], None
            while True:
                ok = self.cap.grab()
                frame_count += 1                               # counts from 1, before the `ret` check
                if not ok:
                    break
                if frame_count % self.stride:
                    continue                                   # skipped frames are never colour-converted
                if slot is None:
                    slot, ev = self._free.get()
                    if ev is not None:
                        ev.synchronize()                       # the DMA that last read this buffer is done
                dst = self._views[slot][n]
                ok, out = self.cap.retrieve(dst)
                if not ok:
                    break
                if out is not dst:                             # decoder returned its own buffer (size / layout change)
                    np.copyto(dst, out)
                numbers.append(frame_count)
                n += 1
                if n == self.batch:
                    self._full.put((slot, n, numbers))
                    n, numbers, slot = 0, [], None
            if n:
                self._full.put((slot, n, numbers))
            self.frames_read = frame_count - 1
        except Exception as e:                                 # surfaced in the consumer thread
            self._error = e
        finally:
            self.cap.release()
            self._full.put(None)

    def __iter__(self):
        while True:
            item = self._full.get()
            if item is None:
                if self._error is not None:
                    raise self._error
                return
            slot, n, numbers = item
            yield self.slots[slot][:n], numbers, slot

    def release(self, slot, event=None):
        """The batch of `slot` has been handed on; `event` (torch.cuda.Event or None) marks when its
        bytes have left host memory."""
        self._free.put((slot, event))
