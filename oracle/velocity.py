"""CPU oracle for the velocity stage (rows a10-a13 of SURVEY.md section 8) -- TEST INFRASTRUCTURE.

A flat, array-based restatement of the reference's streaming classes, written in the
same shape as the CUDA kernel (`vbt_b200/csrc/velocity.cu`) so the two can be compared
field by field:

* ``kahan_window_mean``   <- plot.py:90-95 (`rolling(5, min_periods=1).mean()` and
                             `expanding().mean()`), i.e. pandas' roll_mean (Kahan add /
                             remove with separate compensations, consecutive-same-value
                             and sign fix-ups).  Pinned bit-exactly against pandas 3.0.2.
* ``running_average``     <- RunningAverage.py:9-27.
* ``velocity_phases``     <- VelocityTracker.py:92-230 + Phase.py:6-40.

Parity status: PINNED.  tests/test_oracle_velocity.py checks this file against (a) the
live reference classes imported from /root/reference when present, (b) the committed
golden vectors generated from them (tests/golden/make_golden.py) and (c) the ROM/ACV
labels extracted from figs_ocsort/*.pdf.
"""
from __future__ import annotations

import math

import numpy as np

CONCENTRIC, ECCENTRIC, HOLD = 0, 1, 2          # Phase.py:12-14
START_THRESHOLD, END_THRESHOLD = 3, 1          # VelocityTracker.py:11-12
RA_WINDOW = 30                                 # VelocityTracker.py:44


def kahan_window_mean(values, window):
    """Windowed mean the way pandas computes it (window=0 means expanding).

    plot.py:90-92 uses window 5 on x,y,dx,dy; plot.py:94-95 uses the expanding form on
    the two plate columns.  min_periods is 1 in both.
    """
    v = np.asarray(values, dtype=np.float64)
    n = v.shape[0]
    out = np.empty(n, dtype=np.float64)
    comp_add = comp_rem = total = 0.0
    nobs = neg = same = 0
    prev = v[0] if n else 0.0
    for i in range(n):
        if window and i >= window:               # value leaving the window
            x = v[i - window]
            nobs -= 1
            y = -x - comp_rem
            t = total + y
            comp_rem = t - total - y
            total = t
            if math.copysign(1.0, x) < 0:
                neg -= 1
        x = v[i]                                 # value entering the window
        nobs += 1
        y = x - comp_add
        t = total + y
        comp_add = t - total - y
        total = t
        if math.copysign(1.0, x) < 0:
            neg += 1
        same = same + 1 if x == prev else 1
        prev = x
        r = total / nobs
        if same >= nobs:
            r = prev
        elif neg == 0 and r < 0:
            r = 0.0
        elif neg == nobs and r > 0:
            r = 0.0
        out[i] = r
    return out


def smooth_rows(rows):
    """plot.py:90-95 on a [n,7] array (time,x,y,dx,dy,h,w) of ONE id, time-ordered."""
    rows = np.array(rows, dtype=np.float64, copy=True)
    for c in (1, 2, 3, 4):
        rows[:, c] = kahan_window_mean(rows[:, c], 5)
    for c in (5, 6):
        rows[:, c] = kahan_window_mean(rows[:, c], 0)
    return rows


def running_average(values, window_size):
    """RunningAverage.py:16-27 applied to a whole sequence; returns the emitted means."""
    ring = []
    total = 0.0
    count = 0
    out = []
    for val in values:
        ring.append(val)
        total += val
        count += 1
        if count >= window_size:
            out.append(total / window_size)
            total -= ring.pop(0)
            count -= 1
        else:
            out.append(total / count)
    return np.array(out, dtype=np.float64)


class _Lane:
    """All state one (video, id) series carries; same fields as the kernel's lane."""

    def __init__(self):
        self.phase = HOLD
        self.have_max = False
        self.max_y_diff = 0.0
        self.have_prev = False
        self.y_prev = 0.0
        self.neg = 0
        self.pos = 0
        self.path = []            # rows (x, y, width, height, time)
        self.ra_ring = []
        self.ra_total = 0.0
        self.ra_count = 0
        self.phases = []          # rows (t_start, t_end, y_start, y_end, rom, type)


def _ra_step(s, val):
    s.ra_ring.append(val)
    s.ra_total += val
    s.ra_count += 1
    if s.ra_count >= RA_WINDOW:
        avg = s.ra_total / RA_WINDOW
        s.ra_total -= s.ra_ring.pop(0)
        s.ra_count -= 1
        return avg
    return s.ra_total / s.ra_count


def _drop_small(s):
    """VelocityTracker.py:50-67."""
    lim = s.max_y_diff / 2
    s.phases = [p for p in s.phases if not (abs(p[2] - p[3]) < lim)]


def _close_phase(s, plate_diameter, diff_threshold, min_distance):
    """VelocityTracker.py:171-222."""
    ys = [p[1] for p in s.path]
    hi = int(np.argmax(ys))
    lo = int(np.argmin(ys))
    a, b = (hi, lo) if s.phase == CONCENTRIC else (lo, hi)
    y_diff = abs(ys[a] - ys[b])
    if (not s.have_max) or y_diff > s.max_y_diff:
        s.have_max = True
        s.max_y_diff = y_diff
        _drop_small(s)
    if y_diff > s.max_y_diff * diff_threshold:
        dist = 0.0
        for i in range(a + 1, b + 1):
            p, q = s.path[i], s.path[i - 1]
            ddx = abs(p[0] - q[0]) / ((p[2] + q[2]) / 2) * plate_diameter
            ddy = abs(p[1] - q[1]) / ((p[3] + q[3]) / 2) * plate_diameter
            dist += ddx + ddy
        if not (dist < min_distance):
            s.phases.append((s.path[a][4], s.path[b][4], ys[a], ys[b], dist, s.phase))
            _drop_small(s)
    s.phase = HOLD
    s.neg = 0
    s.pos = 0


def velocity_step(s, row, plate_diameter, diff_threshold=0.6, min_distance=0.1):
    """One call of VelocityTracker.process_measurements (VelocityTracker.py:92-158)."""
    time, x, y, _dx, dy, h, w = row
    width = _ra_step(s, w)        # :98  the SAME running average is fed width then
    height = _ra_step(s, h)       # :99  height (VelocityTracker.py:44-45 quirk)
    if s.have_prev:
        dy = y - s.y_prev
    if s.phase != HOLD:
        s.path.append((x, y, width, height, time))
    if s.phase == CONCENTRIC:
        if dy > 0:
            s.pos += 1
            s.neg = 0
            if s.pos >= END_THRESHOLD:
                _close_phase(s, plate_diameter, diff_threshold, min_distance)
        else:
            s.pos = 0
    if s.phase == ECCENTRIC:
        if dy < 0:
            s.neg += 1
            s.pos = 0
            if s.neg >= END_THRESHOLD:
                _close_phase(s, plate_diameter, diff_threshold, min_distance)
        else:
            s.neg = 0
            s.pos += 1
    if dy < 0 and s.phase == HOLD:
        s.neg += 1
        s.pos = 0
        if s.neg == 1:
            s.path = []
        else:
            s.path.append((x, y, width, height, time))
        if s.neg >= START_THRESHOLD:
            s.phase, s.neg, s.pos = CONCENTRIC, 0, 0
    if dy > 0 and s.phase == HOLD:
        s.pos += 1
        s.neg = 0
        if s.pos == 1:
            s.path = []
        else:
            s.path.append((x, y, width, height, time))
        if s.pos >= START_THRESHOLD:
            s.phase, s.neg, s.pos = ECCENTRIC, 0, 0
    s.have_prev = True
    s.y_prev = y


def velocity_phases(rows, plate_diameter=0.45, diff_threshold=0.6, min_distance=0.1,
                    finish=True, lane=None):
    """plot.analyze_df (plot.py:33-47) on an [n,7] array (time,x,y,dx,dy,h,w).

    Returns (phases[k,6] float64 = t_start,t_end,y_start,y_end,rom,type ; lane)."""
    s = lane if lane is not None else _Lane()
    for r in np.asarray(rows, dtype=np.float64).reshape(-1, 7):
        velocity_step(s, tuple(float(t) for t in r), plate_diameter, diff_threshold,
                      min_distance)
    if finish and s.phase != HOLD:           # VelocityTracker.py:224-230
        _close_phase(s, plate_diameter, diff_threshold, min_distance)
    ph = np.array(s.phases, dtype=np.float64).reshape(-1, 6)
    return ph, s


def analyze_series(rows, plate_diameter=0.45):
    """plot.py:87-95 + 163: smoothing then phases for one id's raw rows."""
    return velocity_phases(smooth_rows(rows), plate_diameter)[0]
