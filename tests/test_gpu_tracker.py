"""K7 parity on a B200: tracker kernel (through the C ABI) vs oracle/ocsort.py.
Bar: row tables bit-exact (fp64), ids and row order identical."""
import numpy as np
import pytest

import helpers
from oracle import ocsort as oo

pytestmark = pytest.mark.gpu


def run_kernel(videos, max_det=None, **kw):
    """videos: list of (fps, frame_numbers, dets list).  One launch, one warp per video."""
    import torch
    from vbt_b200.ocsort import BatchedTracker
    V = len(videos)
    F = max(len(v[2]) for v in videos)
    D = max_det or max(max((len(d) for d in v[2]), default=1) for v in videos)
    D = max(D, 1)
    dets = np.zeros((V, F, D, 6))
    cnt = np.zeros((V, F), np.int32)
    fno = np.zeros((V, F), np.int32)
    for i, (fps, keys, ds) in enumerate(videos):
        for f, (k, d) in enumerate(zip(keys, ds)):
            d = np.asarray(d).reshape(-1, 6)
            dets[i, f, :len(d)] = d
            cnt[i, f] = len(d)
            fno[i, f] = k
    bt = BatchedTracker(V, row_cap=max(1, F * D), **kw)
    dev = lambda a: torch.as_tensor(a, device='cuda')
    bt.update(dev(dets), dev(cnt), dev(fno), dev(np.array([v[0] for v in videos], np.float64)),
              dev(np.array([len(v[2]) for v in videos], np.int32)))
    bt.check_status()
    return [bt.rows_host(i) for i in range(V)], bt


def test_all_34_fixture_videos_one_launch():
    names = sorted(helpers.golden_tables())
    videos = [helpers.fixture_detections(n) for n in names]
    got, _ = run_kernel(videos)
    for n, v, rows in zip(names, videos, got):
        want = oo.track_rows(v[2], v[0], frame_numbers=v[1])
        assert rows.shape == want.shape, n
        assert np.array_equal(rows, want), n


def crowded_scene(seed, n_obj=8, n_frames=120, clutter=0.3, drop=0.15):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.2, 0.8, (n_obj, 2))
    vel = rng.normal(0, 0.01, (n_obj, 2))
    size = rng.uniform(0.05, 0.2, (n_obj, 2))
    frames = []
    for _ in range(n_frames):
        pos += vel + rng.normal(0, 0.002, pos.shape)
        vel += rng.normal(0, 0.002, vel.shape)
        d = []
        for o in range(n_obj):
            if rng.random() < drop:
                continue
            c, s = pos[o] + rng.normal(0, 0.003, 2), size[o] * rng.uniform(0.95, 1.05, 2)
            d.append([c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2,
                      rng.uniform(0.3, 1.0), 0.0])
        while rng.random() < clutter:
            c, s = rng.uniform(0, 1, 2), rng.uniform(0.02, 0.3, 2)
            d.append([c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2,
                      rng.uniform(0.1, 1.0), 0.0])
        rng.shuffle(d)
        frames.append(np.array(d, dtype=np.float64).reshape(-1, 6))
    return frames


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_crowded_scenes_exercise_assignment_ocr_reupdate(seed):
    frames = crowded_scene(seed)
    keys = list(range(1, len(frames) + 1))
    want = oo.track_rows(frames, 30.0, frame_numbers=keys)
    got, _ = run_kernel([(30.0, keys, frames)])
    assert len(want) > 100
    assert got[0].shape == want.shape
    assert np.array_equal(got[0], want)


def test_velocity_direction_cost_with_confidence_column():
    """vdc multiplied by the score (5-column upstream behaviour): arccos may differ in
    the last bit between libm and CUDA, so geometry is compared to 1e-9."""
    frames = crowded_scene(7, n_obj=5, n_frames=80)
    keys = list(range(1, len(frames) + 1))
    want = oo.track_rows(frames, 30.0, frame_numbers=keys, vdc_uses_class_column=False)
    got, _ = run_kernel([(30.0, keys, frames)], vdc_uses_class_column=False)
    assert got[0].shape == want.shape
    assert np.array_equal(got[0][:, 0], want[:, 0])
    assert np.allclose(got[0], want, rtol=0, atol=1e-9)


def test_streaming_in_batches_equals_single_call():
    import torch
    from vbt_b200.ocsort import BatchedTracker
    frames = crowded_scene(5, n_obj=4, n_frames=90)
    keys = list(range(1, len(frames) + 1))
    whole, _ = run_kernel([(30.0, keys, frames)], max_det=16)
    bt = BatchedTracker(1, row_cap=90 * 16)
    dev = lambda a: torch.as_tensor(a, device='cuda')
    for s in range(0, 90, 32):
        chunk = frames[s:s + 32]
        dets = np.zeros((1, 32, 16, 6))
        cnt = np.zeros((1, 32), np.int32)
        fno = np.zeros((1, 32), np.int32)
        for f, d in enumerate(chunk):
            dets[0, f, :len(d)] = d
            cnt[0, f] = len(d)
            fno[0, f] = keys[s + f]
        bt.update(dev(dets), dev(cnt), dev(fno), dev(np.array([30.0])),
                  dev(np.array([len(chunk)], np.int32)))
    bt.check_status()
    assert np.array_equal(bt.rows_host(0), whole[0])


def test_ocsort_facade_per_frame():
    from vbt_b200.ocsort import OCSort
    frames = crowded_scene(9, n_obj=3, n_frames=40, clutter=0.0, drop=0.1)
    trk = OCSort(max_age=30, asso_func='diou', iou_threshold=0.1)
    ref = oo.OCSortOracle(max_age=30, iou_threshold=0.1)
    for d in frames:
        if len(d) == 0:
            continue
        out = trk.update(d, [])
        want = ref.update(d)
        if len(want) == 0:
            assert out.shape == (0, 5)
            continue
        assert np.array_equal(out, want[:, :7])
        by_id = {t.id: t for t in trk.trackers}
        for row in want:
            x = by_id[int(row[4]) - 1].kf.x
            assert x.shape == (7, 1) and x[4, 0] == row[7] and x[5, 0] == row[8]   # track.py:199
    with pytest.raises(NotImplementedError):
        OCSort(asso_func='giou')


def test_track_capacity_fails_loudly():
    from vbt_b200 import _lib
    rng = np.random.default_rng(0)
    frames = []
    for _ in range(40):     # 20 far-apart new boxes per frame -> > 256 live tracks
        c = rng.uniform(0, 50, (20, 2))
        frames.append(np.concatenate([c, c + 0.01, np.full((20, 1), 0.9), np.zeros((20, 1))], axis=1))
    with pytest.raises(_lib.VbtError):
        run_kernel([(30.0, list(range(1, 41)), frames)])
