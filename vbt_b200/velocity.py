"""Host-side mirror of the reference's velocity classes over the K8 CUDA kernel.

Drop-in surfaces (SURVEY.md 8b "Velocity classes"):

* ``Phase``           <- Phase.py:6-40 (attributes, ``y_diff``, ``duration``, ``str()``,
                         constants CONCENTRIC/ECCENTRIC/HOLD = 0/1/2)
* ``RunningAverage``  <- RunningAverage.py:9-27 (``update`` + window_size/window/total/count)
* ``VelocityTracker`` <- VelocityTracker.py:15-230 (``process_measurements``,
                         ``end_processing``, ``phases``, ``current_phase``, ``max_y_diff``)
* ``analyze_df``      <- plot.py:33-47; ``smooth_and_analyze`` adds plot.py:87-95
* ``analyze_batch``   -- many (video, id) series in one kernel launch (north_star item 4)

All arithmetic runs in ``vbt_b200/csrc/velocity.cu``; this module only buffers samples,
moves them to the device and wraps the results.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
from collections import deque

import numpy as np

from . import _lib

PATH_CAP = 1 << 15      # samples one phase path may hold
PHASE_CAP = 512


class Phase:
    """One concentric / eccentric phase of a set (value object)."""

    CONCENTRIC = 0
    ECCENTRIC = 1
    HOLD = 2
    _NAMES = {0: 'concentric', 1: 'eccentric'}
    __slots__ = ('time_start', 'time_end', 'y_start', 'y_end', 'type', 'rom')

    def __init__(self, time_start, time_end, y_start, y_end, rom, phase_type):
        self.time_start, self.time_end = time_start, time_end
        self.y_start, self.y_end = y_start, y_end
        self.rom = rom              # range of motion [m]
        self.type = phase_type

    y_diff = property(lambda self: abs(self.y_start - self.y_end))
    duration = property(lambda self: self.time_end - self.time_start)

    def __str__(self):
        kind = self._NAMES.get(self.type, 'hold')
        return (f'{kind}, t_start: {self.time_start}, t_end: {self.time_end}, '
                f'y_start: {self.y_start}, y_end: {self.y_end}')

    def __repr__(self):
        return f'Phase({self})'

    @classmethod
    def _from_row(cls, r):
        return cls(float(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4]), int(r[5]))


def _phases_from(arr, count):
    return [Phase._from_row(arr[i]) for i in range(count)]


class _Lanes:
    """Owns a vbt_velocity handle plus the small device index tensors its kernel needs."""

    def __init__(self, n_lanes, path_cap=PATH_CAP, phase_cap=PHASE_CAP):
        self.torch = _lib.require_cuda()
        self.L, self.path_cap, self.phase_cap = n_lanes, path_cap, phase_cap
        h = C.c_void_p()
        _lib.check(_lib.lib().vbt_velocity_create(n_lanes, path_cap, phase_cap, C.byref(h)))
        self.handle = h

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                _lib.lib().vbt_velocity_destroy(h)
            except Exception:      # interpreter shutdown: module globals are gone
                pass

    def reset(self):
        _lib.check(_lib.lib().vbt_velocity_reset(self.handle, _lib.stream_ptr()))

    def update(self, rows, row_count, row_cap, lane_table, lane_id, lane_begin, n_lanes,
               plate_diameter, diff_threshold, min_distance, smooth, finish):
        _lib.check(_lib.lib().vbt_velocity_update(
            self.handle, _lib.ptr(rows), _lib.ptr(row_count), row_cap, _lib.ptr(lane_table),
            _lib.ptr(lane_id), _lib.ptr(lane_begin), n_lanes, plate_diameter, diff_threshold,
            min_distance, int(smooth), int(finish), _lib.stream_ptr()))

    def read(self):
        phases = np.empty((self.L, self.phase_cap, _lib.PHASE_COLS), dtype=np.float64)
        count = np.empty(self.L, dtype=np.int32)
        state = np.empty((self.L, 8), dtype=np.float64)
        _lib.check(_lib.lib().vbt_velocity_read(self.handle, _lib.ptr(phases), _lib.ptr(count),
                                                _lib.ptr(state), _lib.stream_ptr()))
        return phases, count, state


class RunningAverage:
    """Windowed running mean with a running float sum; state lives on the device."""

    def __init__(self, window_size):
        self.window_size = window_size
        self._torch = _lib.require_cuda()
        self._state = self._torch.zeros(window_size + 3, dtype=self._torch.float64, device='cuda')

    def update_many(self, values):
        t = self._torch
        v = t.as_tensor(np.asarray(values, dtype=np.float64).reshape(-1), device='cuda')
        out = t.empty_like(v)
        _lib.check(_lib.lib().vbt_running_average(self._state.data_ptr(), self.window_size,
                                                  v.data_ptr(), v.numel(), out.data_ptr(),
                                                  _lib.stream_ptr()))
        return out.cpu().numpy()

    def update(self, value):
        return float(self.update_many([value])[0])

    def _host(self):
        return self._state.cpu().numpy()

    @property
    def total(self):
        return float(self._host()[self.window_size])

    @property
    def count(self):
        return int(self._host()[self.window_size + 1])

    @property
    def window(self):
        s = self._host()
        w, n, head = self.window_size, int(s[self.window_size + 1]), int(s[self.window_size + 2])
        return deque(float(s[(head + i) % w]) for i in range(n))


class VelocityTracker:
    """Streaming HOLD / CONCENTRIC / ECCENTRIC state machine over (t, x, y, w, h) samples.

    Samples are buffered on the host and pushed through the K8 kernel (one lane, no
    smoothing -- the reference class does none) whenever results are read.
    """

    def __init__(self, plate_diameter, diff_threshold=0.6, min_distance=0.1):
        self.plate_diameter = plate_diameter
        self.min_distance = min_distance
        self.diff_threshold = diff_threshold
        self._lanes = _Lanes(1)
        self._pending = []
        self._phases = []
        self._state = np.array([Phase.HOLD, np.nan, 0, 0, 0, np.nan, 0, 0], dtype=np.float64)

    def process_measurements(self, time, x, y, dx, dy, norm_plate_height, norm_plate_width):
        self._pending.append((0.0, time, x, y, dx, dy, norm_plate_height, norm_plate_width))

    def end_processing(self):
        self._flush(finish=True)

    def _flush(self, finish=False):
        if not self._pending and not finish:
            return
        t = self._lanes.torch
        n = len(self._pending)
        rows = np.asarray(self._pending, dtype=np.float64).reshape(n, _lib.ROW_COLS)
        self._pending = []
        cap = max(n, 1)
        d_rows = t.zeros((1, cap, _lib.ROW_COLS), dtype=t.float64, device='cuda')
        if n:
            d_rows[0, :n] = t.as_tensor(rows, device='cuda')
        d_cnt = t.tensor([n], dtype=t.int32, device='cuda')
        d_tab = t.zeros(1, dtype=t.int32, device='cuda')
        d_id = t.full((1,), -1, dtype=t.int32, device='cuda')
        d_beg = t.zeros(1, dtype=t.int32, device='cuda')
        self._lanes.update(d_rows, d_cnt, cap, d_tab, d_id, d_beg, 1, self.plate_diameter,
                           self.diff_threshold, self.min_distance, smooth=False, finish=finish)
        phases, count, state = self._lanes.read()
        self._phases = _phases_from(phases[0], int(count[0]))
        self._state = state[0]

    @property
    def phases(self):
        self._flush()
        return self._phases

    @property
    def current_phase(self):
        self._flush()
        return int(self._state[0])

    @property
    def max_y_diff(self):
        self._flush()
        return None if np.isnan(self._state[1]) else float(self._state[1])

    @property
    def y_prev(self):
        self._flush()
        return None if np.isnan(self._state[5]) else float(self._state[5])


def analyze_batch(series, plate_diameter=0.45, diff_threshold=0.6, min_distance=0.1,
                  smooth=False, return_state=False):
    """Phases of many independent series in ONE kernel launch.

    series: list of float64 [n_i,7] arrays (time,x,y,dx,dy,norm_plate_height,
    norm_plate_width), each one id of one video, time ordered.  smooth=True applies
    plot.py:90-95 on the device first.  Returns a list of ``list[Phase]``.
    """
    torch = _lib.require_cuda()
    L = len(series)
    if L == 0:
        return ([], np.zeros((0, 8))) if return_state else []
    lens = [int(np.asarray(s).reshape(-1, 7).shape[0]) for s in series]
    cap = max(max(lens), 1)
    host = np.zeros((L, cap, _lib.ROW_COLS), dtype=np.float64)
    for i, s in enumerate(series):
        host[i, :lens[i], 1:] = np.asarray(s, dtype=np.float64).reshape(-1, 7)
    lanes = _Lanes(L, path_cap=cap + 1)
    d_rows = torch.as_tensor(host, device='cuda')
    d_cnt = torch.as_tensor(np.asarray(lens, dtype=np.int32), device='cuda')
    d_tab = torch.arange(L, dtype=torch.int32, device='cuda')
    d_id = torch.full((L,), -1, dtype=torch.int32, device='cuda')
    d_beg = torch.zeros(L, dtype=torch.int32, device='cuda')
    lanes.update(d_rows, d_cnt, cap, d_tab, d_id, d_beg, L, plate_diameter, diff_threshold,
                 min_distance, smooth=smooth, finish=True)
    phases, count, state = lanes.read()
    out = [_phases_from(phases[i], int(count[i])) for i in range(L)]
    return (out, state) if return_state else out


def analyze_df(df, plate_diameter):
    """plot.analyze_df (plot.py:33-47): rows are unpacked POSITIONALLY as
    time,x,y,dx,dy,norm_plate_height,norm_plate_width (column order is the contract)."""
    return analyze_batch([df.to_numpy(dtype=np.float64)], plate_diameter)[0]


def smooth_and_analyze(df, tracking_id, plate_diameter=0.45):
    """plot.py:87-95 + 163 for one pickled DataFrame: select the id, smooth on the
    device (rolling 5 / expanding means) and segment phases."""
    d = df[df['id'] == tracking_id].drop(columns=['id'])
    return analyze_batch([d.to_numpy(dtype=np.float64)], plate_diameter, smooth=True)[0]
