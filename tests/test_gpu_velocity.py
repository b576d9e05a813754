"""K8 parity on a B200: CUDA kernel (through the C ABI) vs oracle and reference goldens.
Bar: bit-exact fp64."""
import numpy as np
import pytest

import helpers
from oracle import velocity as ov

pytestmark = pytest.mark.gpu


def phases_array(phases):
    return np.array([[p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type]
                     for p in phases], dtype=np.float64).reshape(-1, 6)


def test_all_fixture_series_one_launch():
    from vbt_b200.velocity import analyze_batch
    series = helpers.all_series()
    got = analyze_batch([raw for _, raw, _ in series], 0.45, smooth=True)
    assert len(got) == len(series)
    total = 0
    for (key, _, want), ph in zip(series, got):
        arr = phases_array(ph)
        assert arr.shape == want.shape, key
        assert np.array_equal(arr, want), key
        total += len(want)
    assert total > 500


def test_figs_ocsort_labels_34_of_34():
    from vbt_b200.velocity import analyze_batch
    labels = helpers.golden_labels()
    names = sorted(labels)
    raws = [helpers.series_of(n, int(n.split('_id')[1].split('_')[0])) for n in names]
    got = analyze_batch(raws, 0.45, smooth=True)
    for n, ph in zip(names, got):
        assert helpers.labels_from_phases(phases_array(ph)) == labels[n], n


def test_velocity_tracker_class_streaming():
    """The drop-in class fed one sample at a time, results read mid-stream."""
    from vbt_b200.velocity import VelocityTracker, Phase
    series = [s for s in helpers.all_series() if len(s[2]) > 2][:3]
    for key, raw, want in series:
        sm = ov.smooth_rows(raw)
        vt = VelocityTracker(0.45)
        assert vt.current_phase == Phase.HOLD and vt.max_y_diff is None and vt.phases == []
        for i, r in enumerate(sm):
            vt.process_measurements(*r)
            if i == len(sm) // 2:
                mid, _ = ov.velocity_phases(sm[:i + 1], 0.45, finish=False)
                assert np.array_equal(phases_array(vt.phases), mid)
        vt.end_processing()
        assert np.array_equal(phases_array(vt.phases), want), key
        for ph in vt.phases[:1]:
            assert str(ph).startswith(('concentric', 'eccentric'))
            assert ph.duration == ph.time_end - ph.time_start and ph.y_diff == abs(ph.y_start - ph.y_end)


def test_random_series_vs_oracle():
    from vbt_b200.velocity import analyze_batch
    rng = np.random.default_rng(11)
    series = []
    for _ in range(64):
        n = int(rng.integers(1, 600))
        t = np.arange(1, n + 1) / 30.0
        y = 0.5 + 0.25 * np.sin(np.linspace(0, rng.uniform(1, 40), n)) + rng.normal(0, 0.01, n)
        rows = np.stack([t, 0.5 + rng.normal(0, 0.01, n), y, rng.normal(0, 0.01, n),
                         rng.normal(0, 0.01, n), 0.1 + rng.normal(0, 0.005, n),
                         0.2 + rng.normal(0, 0.005, n)], axis=1)
        series.append(rows)
    series.append(np.zeros((0, 7)))                      # empty series
    series.append(series[0][:1])                         # single sample
    for smooth in (False, True):
        got = analyze_batch(series, 0.45, smooth=smooth)
        for rows, ph in zip(series, got):
            want = ov.velocity_phases(ov.smooth_rows(rows) if smooth and len(rows) else rows, 0.45)[0]
            assert np.array_equal(phases_array(ph), want)


def test_running_average_class():
    from vbt_b200.velocity import RunningAverage
    rng = np.random.default_rng(3)
    vals = rng.normal(size=95)
    ra = RunningAverage(30)
    got = [ra.update(v) for v in vals[:40]] + list(ra.update_many(vals[40:]))
    assert np.array_equal(np.array(got), ov.running_average(vals, 30))
    assert ra.window_size == 30 and ra.count == 29 and len(ra.window) == 29
    assert ra.window[-1] == vals[-1]


def test_path_capacity_fails_loudly():
    from vbt_b200 import _lib
    from vbt_b200.velocity import _Lanes
    import torch
    n = 400
    rows = np.zeros((1, n, 8))
    rows[0, :, 1] = np.arange(1, n + 1) / 30
    rows[0, :, 3] = np.linspace(0.9, 0.1, n)            # one endless concentric phase
    rows[0, :, 6] = 0.1
    rows[0, :, 7] = 0.2
    lanes = _Lanes(1, path_cap=64, phase_cap=4)
    z = lambda *a: torch.zeros(*a, dtype=torch.int32, device='cuda')
    lanes.update(torch.as_tensor(rows, device='cuda'), torch.tensor([n], dtype=torch.int32, device='cuda'),
                 n, z(1), torch.full((1,), -1, dtype=torch.int32, device='cuda'), z(1), 1, 0.45, 0.6,
                 0.1, smooth=False, finish=True)
    with pytest.raises(_lib.VbtError):
        lanes.read()
