"""Pins oracle/velocity.py (CPU, no GPU needed) against: the committed golden phases
produced by the LIVE reference classes, the figs_ocsort PDF labels, pandas' own
rolling/expanding means, and -- when /root/reference is present -- the reference classes
themselves."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
from hypothesis import given, settings, strategies as st

import helpers
from oracle import velocity as ov


def test_window_mean_matches_pandas():
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 300))
        v = rng.normal(size=n) * rng.choice([1e-3, 1.0, 1e3])
        if trial % 3 == 0:
            v = np.abs(v)
        if trial % 5 == 0:
            v[rng.integers(0, n, size=n // 2)] = v[0]
        s = pd.Series(v)
        assert np.array_equal(ov.kahan_window_mean(v, 5),
                              s.rolling(window=5, center=False, min_periods=1).mean().to_numpy())
        assert np.array_equal(ov.kahan_window_mean(v, 0),
                              s.expanding(min_periods=1).mean().to_numpy())


def test_golden_phases_all_series():
    series = helpers.all_series()
    assert len(series) >= 34
    n_ph = 0
    for key, raw, want in series:
        got = ov.analyze_series(raw, 0.45)
        assert got.shape == want.shape, key
        assert np.array_equal(got, want), key      # bit-exact, fp64
        n_ph += len(want)
    assert n_ph > 500


def test_figs_ocsort_labels_34_of_34():
    labels = helpers.golden_labels()
    assert len(labels) == 34
    for name, want in labels.items():
        tid = int(name.split('_id')[1].split('_')[0])
        got = helpers.labels_from_phases(ov.analyze_series(helpers.series_of(name, tid), 0.45))
        assert got == want, name


def test_running_average_window_edges():
    vals = np.arange(1.0, 80.0) * 0.37
    got = ov.running_average(vals, 30)
    # before the window fills: plain prefix means; afterwards: mean of the last 30
    assert np.allclose(got[:29], np.cumsum(vals)[:29] / np.arange(1, 30))
    assert np.allclose(got[40], vals[11:41].mean())
    assert ov.running_average([], 30).shape == (0,)


def test_empty_and_short_series():
    assert ov.velocity_phases(np.zeros((0, 7)))[0].shape == (0, 6)
    one = np.array([[0.1, 0.5, 0.5, 0.0, 0.0, 0.1, 0.1]])
    assert ov.analyze_series(one).shape == (0, 6)


@pytest.mark.skipif(not os.path.isdir(helpers.REFERENCE), reason='reference checkout absent')
def test_live_reference_classes():
    sys.path.insert(0, helpers.REFERENCE)
    try:
        from VelocityTracker import VelocityTracker   # the reference's own class
        from RunningAverage import RunningAverage
    finally:
        sys.path.remove(helpers.REFERENCE)
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(5, 400))
        t = np.arange(1, n + 1) / 30.0
        y = 0.5 + 0.25 * np.sin(np.linspace(0, rng.uniform(2, 30), n)) + rng.normal(0, 0.004, n)
        x = 0.5 + rng.normal(0, 0.003, n)
        h = 0.12 + rng.normal(0, 0.002, n)
        w = 0.2 + rng.normal(0, 0.002, n)
        rows = np.stack([t, x, y, np.zeros(n), np.gradient(y), h, w], axis=1)
        vt = VelocityTracker(0.45)
        for r in rows:
            vt.process_measurements(*r)
        vt.end_processing()
        want = np.array([[p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type]
                         for p in vt.phases]).reshape(-1, 6)
        got = ov.velocity_phases(rows, 0.45)[0]
        assert np.array_equal(got, want)
    ra = RunningAverage(30)
    vals = rng.normal(size=100)
    assert np.array_equal(np.array([ra.update(v) for v in vals]), ov.running_average(vals, 30))


@settings(max_examples=25, deadline=None)
@given(st.lists(st.floats(-1.0, 1.0, allow_nan=False, width=64), min_size=1, max_size=120))
def test_streaming_equals_whole(ys):
    """Feeding a series in two chunks == feeding it at once (the kernel relies on it)."""
    n = len(ys)
    rows = np.zeros((n, 7))
    rows[:, 0] = np.arange(1, n + 1) / 30.0
    rows[:, 1] = 0.5
    rows[:, 2] = ys
    rows[:, 5] = 0.1
    rows[:, 6] = 0.2
    whole = ov.velocity_phases(rows)[0]
    _, lane = ov.velocity_phases(rows[: n // 2], finish=False)
    split = ov.velocity_phases(rows[n // 2:], lane=lane)[0]
    assert np.array_equal(whole, split)
