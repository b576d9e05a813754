"""`OCSort`-shaped facade over the K7 tracker kernel (vbt_b200/csrc/tracker.cu).

Mirrors the surface track.py uses (track.py:17,157,186-199):

    tracker = OCSort(max_age=30, asso_func="diou", iou_threshold=0.1)
    out = tracker.update(dets_f64_Nx6, [])          # -> [M,7] x1,y1,x2,y2,id,cls,conf
    for trk in tracker.trackers: trk.id, trk.kf.x   # 0-based id, 7x1 state

One instance = one video, state on the device.  ``BatchedTracker`` is the production
entry: V videos x F frames per launch, rows appended on the device.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import _lib

MAX_TRACKS = 256
MAX_DETS = 32


def _params(det_thresh, max_age, min_hits, iou_threshold, delta_t, inertia,
            vdc_uses_class_column):
    return _lib.TrackerParams(det_thresh, iou_threshold, inertia, max_age, min_hits, delta_t,
                              1 if vdc_uses_class_column else 0)


class BatchedTracker:
    """V independent OC-SORT instances stepped together (one warp per video)."""

    def __init__(self, n_videos=1, row_cap=1 << 16, det_thresh=0.2, max_age=30, min_hits=3,
                 iou_threshold=0.1, delta_t=3, inertia=0.2, vdc_uses_class_column=True,
                 max_tracks=MAX_TRACKS, keep_details=False):
        self.torch = t = _lib.require_cuda()
        self.V, self.row_cap = n_videos, row_cap
        p = _params(det_thresh, max_age, min_hits, iou_threshold, delta_t, inertia,
                    vdc_uses_class_column)
        h = C.c_void_p()
        _lib.check(_lib.lib().vbt_tracker_create(n_videos, max_tracks, C.byref(p), C.byref(h)))
        self.handle = h
        self.rows = t.zeros((n_videos, row_cap, _lib.ROW_COLS), dtype=t.float64, device='cuda')
        self.row_count = t.zeros(n_videos, dtype=t.int32, device='cuda')
        self.details = None
        if keep_details:     # (xmin,ymin,xmax,ymax,score) of every row: what track.py:190 unpacks for its overlay
            self.details = t.zeros((n_videos, row_cap, 5), dtype=t.float64, device='cuda')
            _lib.check(_lib.lib().vbt_tracker_row_details(self.handle, self.details.data_ptr()))

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                _lib.lib().vbt_tracker_destroy(h)
            except Exception:      # interpreter shutdown: module globals are gone
                pass

    def reset(self):
        _lib.check(_lib.lib().vbt_tracker_reset(self.handle, _lib.stream_ptr()))
        self.row_count.zero_()

    def update(self, dets, det_count, frame_no, fps, n_frames, last_out=None,
               last_out_count=None, stream=None):
        """dets f64 [V,F,D,6], det_count i32 [V,F], frame_no i32 [V,F], fps f64 [V],
        n_frames i32 [V] -- all CUDA tensors.  Appends to self.rows / self.row_count."""
        V, F, D, _ = dets.shape
        assert V == self.V
        _lib.check(_lib.lib().vbt_tracker_update(
            self.handle, dets.data_ptr(), det_count.data_ptr(), frame_no.data_ptr(),
            fps.data_ptr(), n_frames.data_ptr(), F, D, self.rows.data_ptr(),
            self.row_count.data_ptr(), self.row_cap, _lib.ptr(last_out),
            _lib.ptr(last_out_count), _lib.stream_ptr(stream)))

    def check_status(self):
        st = np.zeros(self.V, dtype=np.int32)
        _lib.check(_lib.lib().vbt_tracker_status(self.handle, _lib.ptr(st), _lib.stream_ptr()))

    def peek(self, v=0):
        buf = np.zeros((MAX_TRACKS, 9), dtype=np.float64)
        n = _lib.check(_lib.lib().vbt_tracker_peek(self.handle, v, _lib.ptr(buf),
                                                   _lib.stream_ptr()))
        return buf[:n]

    def rows_host(self, v=0):
        n = int(self.row_count[v].item())
        return self.rows[v, :n].cpu().numpy()


class OCSort:
    """Per-frame facade with the reference package's call shape."""

    def __init__(self, det_thresh=0.2, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3,
                 asso_func='iou', inertia=0.2, use_byte=False, vdc_uses_class_column=True):
        if asso_func != 'diou':
            raise NotImplementedError(
                'only asso_func="diou" (the reference configuration, track.py:157) is built')
        if use_byte:
            raise NotImplementedError('use_byte=True is not part of the reference configuration')
        self.det_thresh, self.max_age, self.min_hits = det_thresh, max_age, min_hits
        self.iou_threshold, self.delta_t, self.inertia = iou_threshold, delta_t, inertia
        self._bt = BatchedTracker(1, row_cap=MAX_DETS, det_thresh=det_thresh, max_age=max_age,
                                  min_hits=min_hits, iou_threshold=iou_threshold, delta_t=delta_t,
                                  inertia=inertia, vdc_uses_class_column=vdc_uses_class_column)
        t = self._bt.torch
        self.frame_count = 0
        self._fps = t.ones(1, dtype=t.float64, device='cuda')
        self._one = t.ones(1, dtype=t.int32, device='cuda')
        self._out = t.zeros((1, MAX_DETS, 9), dtype=t.float64, device='cuda')
        self._out_n = t.zeros(1, dtype=t.int32, device='cuda')

    def update(self, dets, _=None):
        t = self._bt.torch
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 6)
        n = dets.shape[0]
        if n > MAX_DETS:
            raise _lib.VbtError(_lib.ECAPACITY, f'{n} detections in one frame (max {MAX_DETS})')
        self.frame_count += 1
        if n == 0:
            # the reference never calls update() on an empty frame (track.py:180-181);
            # upstream would still age the tracks -- not reproduced, fail loudly instead
            raise ValueError('OCSort.update called with no detections; the reference skips '
                             'empty frames before calling the tracker (track.py:180-181)')
        host = np.zeros((1, 1, MAX_DETS, 6), dtype=np.float64)
        host[0, 0, :n] = dets
        d = t.as_tensor(host, device='cuda')
        cnt = t.tensor([[n]], dtype=t.int32, device='cuda')
        fno = t.tensor([[self.frame_count]], dtype=t.int32, device='cuda')
        self._bt.row_count.zero_()
        self._bt.update(d, cnt, fno, self._fps, self._one, self._out, self._out_n)
        self._bt.check_status()
        m = int(self._out_n.item())
        if m == 0:
            return np.empty((0, 5))
        return self._out[0, :m, :7].cpu().numpy()

    @property
    def trackers(self):
        """Live tracks in list order; each has .id (0-based), .time_since_update and
        .kf.x shaped (7,1) like filterpy's state vector."""
        out = []
        for r in self._bt.peek(0):
            out.append(SimpleNamespace(id=int(r[7]), time_since_update=int(r[8]),
                                       kf=SimpleNamespace(x=r[:7].reshape(7, 1).copy())))
        return out
