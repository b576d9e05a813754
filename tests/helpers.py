"""Shared test helpers: golden vectors and fixture-derived detections."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REFERENCE = os.environ.get('VBT_REFERENCE', '/root/reference')

_cache = {}


def golden_tables():
    """{name: (rows[n,8] append order, index[n])} for the 34 dfs_ocsort fixtures."""
    if 'tables' not in _cache:
        z = np.load(os.path.join(GOLDEN, 'dfs_ocsort.npz'))
        names = sorted({k.split('|')[0] for k in z.files})
        _cache['tables'] = {n.split('/')[1]: (z[n + '|rows'], z[n + '|index']) for n in names}
    return _cache['tables']


def golden_phases():
    """{(set/name, id): (phases[k,6], raw[n,7] or None)}."""
    if 'phases' not in _cache:
        z = np.load(os.path.join(GOLDEN, 'velocity_phases.npz'))
        out = {}
        for k in z.files:
            name, tid, kind = k.split('|')
            if kind == 'phases':
                raw_key = f'{name}|{tid}|raw'
                out[(name, int(tid))] = (z[k], z[raw_key] if raw_key in z.files else None)
        _cache['phases'] = out
    return _cache['phases']


def golden_labels():
    with open(os.path.join(GOLDEN, 'figs_ocsort_labels.json')) as f:
        return json.load(f)


def series_of(name, tid):
    """Raw (time,x,y,dx,dy,h,w) rows of one id of a dfs_ocsort fixture, time ordered
    (= what plot.py:88 selects)."""
    rows, _ = golden_tables()[name]
    sel = rows[rows[:, 0] == tid]
    order = np.argsort(sel[:, 1], kind='stable')
    return sel[order][:, 1:]


def all_series():
    """[(key, raw[n,7], phases[k,6])] over dfs_ocsort (every id) and qualysis_dfs."""
    out = []
    for (name, tid), (ph, raw) in sorted(golden_phases().items()):
        sub, base = name.split('/')
        if raw is None:
            raw = series_of(base, tid)
        out.append(((name, tid), raw, ph))
    return out


def fixture_fps(rows):
    t = np.unique(rows[:, 1])
    return float(round(1.0 / np.min(np.diff(t))))


def fixture_detections(name, score=0.9):
    """Rebuild per-frame detections from the emitted rows of a fixture:
    box = centre +- size/2, detections of a frame ordered by id.  Returns
    (fps, frame_numbers[list], dets[list of [n,6]])."""
    rows, _ = golden_tables()[name]
    fps = fixture_fps(rows)
    fr = np.rint(rows[:, 1] * fps).astype(np.int64)
    frames = {}
    for k, r in zip(fr, rows):
        frames.setdefault(int(k), []).append(r)
    keys = sorted(frames)
    dets = []
    for k in keys:
        rs = sorted(frames[k], key=lambda r: r[0])
        dets.append(np.array([[r[2] - r[7] / 2, r[3] - r[6] / 2, r[2] + r[7] / 2,
                               r[3] + r[6] / 2, score, 0.0] for r in rs], dtype=np.float64))
    return fps, keys, dets


def labels_from_phases(phases):
    """plot.py:171-186 text labels of the concentric phases, as a sorted multiset."""
    out = []
    for p in phases:
        if int(p[5]) == 0:
            rom, dur = p[4], p[1] - p[0]
            out.append(f'{rom:0.2f}')
            out.append(f'{rom / dur:0.2f}')
    return sorted(out)
