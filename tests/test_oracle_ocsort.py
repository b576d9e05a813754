"""Pins oracle/ocsort.py (CPU) against the dfs_ocsort fixtures the way SURVEY.md section 4
prescribes: replaying detections rebuilt from the emitted rows must give (1) only ids /
frames that exist in the fixture for tracks visible from their first hit, (2) rows that
are a subset of the fixture rows, (3) kf velocities dx,dy bit-exact on that subset."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

import helpers
from oracle import ocsort as oo

# Known answers of the replay over ALL 34 fixtures (measured once with this oracle, then pinned):
# name prefix -> (replay rows, rows whose id equals the fixture's id, rows with bit-exact kf dx,dy).
# Every replay row has a fixture row with the same frame and centre (geometry match = 100 %).
KNOWN = {
    '001': (5453, 5453, 2462), '002': (1525, 1306, 1206), '003': (1001, 1001, 1001), '004': (1984, 1984, 1984),
    '005': (2136, 2136, 2136), '006': (1807, 1807, 1807), '007': (936, 936, 936), '008': (3032, 2598, 1841),
    '009': (1663, 1663, 1663), '010': (2102, 2102, 2102), '011': (2109, 2109, 2109), '012': (2124, 2026, 1975),
    '013': (1964, 1964, 1964), '014': (1986, 1986, 1986), '015': (2282, 2282, 2282), '016': (1767, 1767, 1736),
    '017': (699, 699, 699), '018': (976, 976, 976), '019': (923, 923, 923), '020': (1129, 1129, 1129),
    '021': (1039, 1039, 1039), '022': (2036, 2035, 2035), '023': (915, 915, 915), '024': (901, 901, 901),
    '025': (1340, 1332, 1330), '026': (1282, 1282, 1282), '027': (945, 945, 945), '028': (1900, 1895, 1888),
    '029': (2411, 2411, 1994), '030': (3243, 3243, 3243), '031': (3133, 3133, 3133), '032': (2106, 2106, 1036),
    '033': (1270, 1270, 1270), '034': (998, 961, 961),
}
TOTAL = (61117, 60315, 54889)          # of 61,461 fixture rows


def replay(name):
    fps, keys, dets = helpers.fixture_detections(name)
    rows = oo.track_rows(dets, fps, frame_numbers=keys)
    fx_rows, _ = helpers.golden_tables()[name]
    by_frame = {}
    for r in fx_rows:
        by_frame.setdefault(int(round(r[1] * fps)), []).append(r)
    return rows, fx_rows, by_frame, fps


_stats = {}


@pytest.mark.parametrize('name', sorted(helpers.golden_tables()))
def test_fixture_replay(name):
    """All 34 fixtures.  Replay tracks are aligned to fixture tracks by GEOMETRY (same frame, same
    centre), not by raw id: a birth the fixture never shows (fewer than min_hits hits, so it was never
    emitted) still consumes an id, and detections rebuilt from emitted rows cannot contain it -- ids of
    later tracks are then shifted by a constant (002: 2 -> 4, 008: 2 -> 3)."""
    rows, fx_rows, by_frame, fps = replay(name)
    assert len(rows) >= 0.94 * len(fx_rows)
    first_seen, frames_of = {}, {}
    for r in fx_rows:
        fr = int(round(r[1] * fps))
        first_seen[int(r[0])] = min(first_seen.get(int(r[0]), 1 << 30), fr)
        frames_of.setdefault(int(r[0]), []).append(fr)
    id_map, exact, same_id = {}, 0, 0
    diverged = set()
    for r in rows:
        fr = int(round(r[1] * fps))
        cands = [f for f in by_frame.get(fr, []) if abs(f[2] - r[2]) < 1e-12 and abs(f[3] - r[3]) < 1e-12]
        # (2) replay rows are a subset of the fixture rows: emitted geometry is the detection itself
        assert len(cands) == 1, (name, fr, int(r[0]))
        f = cands[0]
        # birth rows (frame_count <= min_hits) come from the filter state, whose w/(h+1e-6) round
        # trip moves h,w by ~5e-7
        assert abs(r[6] - f[6]) < 2e-6 and abs(r[7] - f[7]) < 2e-6
        fid = int(f[0])
        # (1) ids: one replay track = one fixture track, never two
        assert id_map.setdefault(int(r[0]), fid) == fid, (name, fr, int(r[0]), fid)
        same_id += fid == int(r[0])
        if (r[4], r[5]) == (f[4], f[5]):
            exact += 1
        elif fid not in diverged:
            diverged.add(fid)
            # (3) a track may only leave bit-exactness after a gap in ITS OWN fixture rows (the 2
            # hits hidden by min_hits are not in the fixture) or when it was born late (same
            # reason) -- never while it has been continuously visible
            prev = sorted(k for k in frames_of[fid] if k <= fr)
            had_gap = any(b - a > 1 for a, b in zip(prev, prev[1:]))
            assert had_gap or first_seen[fid] > 3, (name, fr, fid)
    # (several replay tracks may land on one fixture track: after a stepped gap the fixture's track
    # lived on through two hidden hits, the replay has to start a new one; and a birth the fixture
    # hides shifts the fixture's later ids up.)  Birth ORDER is kept either way
    order = [id_map[k] for k in sorted(id_map)]
    assert order == sorted(order), id_map
    _stats[name[:3]] = (len(rows), same_id, exact)
    assert _stats[name[:3]] == KNOWN[name[:3]]


def test_total_exact_rows():
    """Known-answer totals over the 34 fixtures (runs after the per-fixture replays)."""
    if len(_stats) < len(KNOWN):
        pytest.skip('needs the 34 per-fixture replays of this session (run the whole file)')
    assert tuple(sum(v[i] for v in _stats.values()) for i in range(3)) == TOTAL
    assert tuple(sum(v[i] for v in KNOWN.values()) for i in range(3)) == TOTAL


def test_assignment_is_optimal():
    rng = np.random.default_rng(2)
    for _ in range(200):
        n, m = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        c = rng.normal(size=(n, m))
        pairs = oo.assign_min_cost(c)
        assert len(pairs) == min(n, m)
        assert len({a for a, _ in pairs}) == len(pairs) and len({b for _, b in pairs}) == len(pairs)
        r, cc = linear_sum_assignment(c)
        assert np.isclose(sum(c[a, b] for a, b in pairs), c[r, cc].sum())
        assert pairs == sorted(zip(r.tolist(), cc.tolist()))      # unique optimum a.s.


def test_empty_frames_do_not_step():
    d = np.array([[0.1, 0.1, 0.3, 0.3, 0.9, 0.0]])
    a = oo.track_rows([d, d, d, d], 30.0)
    b = oo.track_rows([d, np.zeros((0, 6)), d, np.zeros((0, 6)), d, d], 30.0,
                      frame_numbers=[1, 2, 3, 4, 5, 6])
    assert np.array_equal(a[:, 2:], b[:, 2:]) and np.array_equal(a[:, 0], b[:, 0])
    assert list(b[:, 1]) == [1 / 30.0, 3 / 30.0, 5 / 30.0, 6 / 30.0]


def test_min_hits_gating_and_reid():
    """A late birth shows on its 3rd hit; ids are 1-based and restart per instance."""
    a = np.array([[0.1, 0.1, 0.3, 0.3, 0.9, 0.0]])
    b = np.array([[0.6, 0.6, 0.8, 0.8, 0.9, 0.0]])
    both = np.vstack([a, b])
    frames = [a] * 5 + [both] * 4
    rows = oo.track_rows(frames, 30.0)
    ids_by_frame = {}
    for r in rows:
        ids_by_frame.setdefault(int(round(r[1] * 30)), []).append(int(r[0]))
    assert ids_by_frame[5] == [1]
    assert ids_by_frame[6] == [1] and ids_by_frame[7] == [1] and ids_by_frame[8] == [1]
    assert ids_by_frame[9] == [2, 1]          # birth + 2 hidden hits, newest track first
