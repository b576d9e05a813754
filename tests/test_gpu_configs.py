"""The BASELINE.json configurations as parity cases on a B200:

  configs[0/1]  track.py pipeline on one synthetic clip (Lite0), through the CLI's `track()` and the
                pickled DataFrame that plot.py consumes;
  configs[2]    Lite1 384x384 detection + tracking at frame batch 256;
  configs[3]    Lite2 448x448 at frame batch 256 (int8);
  configs[4]    many videos sharded by whole videos + one gather of the track tables.

At full batch the CPU oracle would take minutes, so exactness is carried by a size-independent
property -- a frame's result does not depend on the batch it travels in -- checked over the
whole batch against small batches, which in turn are checked against the oracle."""
import os

import numpy as np
import pytest

from oracle import effdet as OE, ocsort as oo, postprocess as OP, resize as OR, velocity as ov

pytestmark = pytest.mark.gpu

_graphs = {}


def graph(variant):
    if variant not in _graphs:
        from vbt_b200 import effdet
        _graphs[variant] = effdet.build_synthetic(variant)
    return _graphs[variant]


def oracle_rows(g, frames, fps, thr, numbers=None):
    imgs = OR.preprocess_batch(frames, g.S, swap_rb=True)
    cls, box, _ = OE.run(g, imgs)
    a = g.anchors()
    dets = []
    for b in range(len(frames)):
        ob, oc, osc, cnt, _ = OP.detection_postprocess(cls[b], box[b], a, g.box_scale, g.box_zp)
        dets.append(OP.tracker_inputs(OP.detect_results(ob, osc, cnt, thr)).reshape(-1, 6))
    return oo.track_rows(dets, fps, max_age=30, iou_threshold=0.1, frame_numbers=numbers)


def synthetic_frames(n, h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    return np.stack([np.roll(base, 4 * i, axis=1) for i in range(n)])


def test_config1_track_cli_on_a_synthetic_clip(tmp_path):
    """cv2-decoded clip -> track() -> DataFrame pickle with the reference's schema and file name
    -> the plot.py analysis; rows equal the CPU oracle chain on the same decoded frames."""
    import cv2
    import pandas as pd
    from vbt_b200.interpreter import Interpreter
    from vbt_b200.pipeline import COLUMNS, export_dataframe
    from vbt_b200.track import track
    from vbt_b200.velocity import smooth_and_analyze
    g = graph('lite0')
    path = str(tmp_path / '007_squat_3reps.avi')
    frames = synthetic_frames(12, 240, 320, seed=5)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'MJPG'), 30.0, (320, 240))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(path)
    fps = cap.get(cv2.CAP_PROP_FPS)
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    decoded = np.stack(decoded)
    interp = Interpreter(model_path=g, num_threads=4)
    interp.allocate_tensors()
    thr = 0.3
    data = track(path, interp, thr, frame_stride=1, batch=8)
    assert list(data) == COLUMNS                                   # key order (track.py:144-145)
    want = oracle_rows(g, decoded, fps, thr)
    got = np.column_stack([np.asarray(data[c], dtype=np.float64) for c in COLUMNS])
    assert len(want) > 0 and got.shape == want.shape
    assert np.array_equal(got, want)
    # HEAD's stride: every 16th frame only (track.py:166) -> frame 16 does not exist in a 12-frame clip
    assert len(track(path, interp, thr)['id']) == 0
    df, out = export_dataframe(data, path, 'models/efficientdet_lite0_whole.tflite', str(tmp_path))
    assert os.path.basename(out).startswith('007_squat_3reps_id') and out.endswith('_efficientdet_lite0_whole.pkl.gz')
    back = pd.read_pickle(out)
    assert list(back.columns) == COLUMNS and back['id'].dtype == np.int64
    assert back[['id', 'time']].values.tolist() == sorted(back[['id', 'time']].values.tolist())
    tid = int(os.path.basename(out).split('_id')[1].split('_')[0])
    phases = smooth_and_analyze(back, tid, 0.45)
    sel = back[back['id'] == tid].drop(columns=['id']).to_numpy(dtype=np.float64)
    want_ph = ov.analyze_series(sel, 0.45)
    got_ph = np.array([[p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type] for p in phases]).reshape(-1, 6)
    assert np.array_equal(got_ph, want_ph)


@pytest.mark.parametrize('variant,cfg', [('lite1', 'configs[2]'), ('lite2', 'configs[3]')])
def test_batch_256_detection_and_tracking(variant, cfg):
    import torch
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = graph(variant)
    n, thr = 256, 0.3
    frames = synthetic_frames(n, 135, 240, seed=9)
    dev = torch.as_tensor(frames, device='cuda')
    big = Detector(g, max_batch=256)
    small = Detector(g, max_batch=8)
    # (a) one launch set over 256 frames == 32 launch sets over 8 frames, raw outputs bit for bit
    big.network(big.preprocess(dev, True))
    cls_big = big.raw_cls[:n, :g.n_anchors].cpu().numpy()
    box_big = big.raw_box[:n, :g.n_anchors].cpu().numpy()
    for s in range(0, n, 8):
        small.network(small.preprocess(dev[s:s + 8], True))
        assert np.array_equal(small.raw_cls[:8, :g.n_anchors].cpu().numpy(), cls_big[s:s + 8]), (cfg, s)
        assert np.array_equal(small.raw_box[:8, :g.n_anchors].cpu().numpy(), box_big[s:s + 8]), (cfg, s)
    # (b) the small batch is the oracle's: first and last frames
    pick = [0, 1, n - 1]
    imgs = OR.preprocess_batch(frames[pick], g.S, swap_rb=True)
    want_cls, want_box, _ = OE.run(g, imgs)
    assert np.array_equal(cls_big[pick], want_cls) and np.array_equal(box_big[pick], want_box)
    # (c) tracking + velocity over the 256-frame batch == the same clip in batches of 8
    numbers = torch.arange(1, n + 1, dtype=torch.int32, device='cuda')
    p_big = VideoPipeline(big, 30.0, thr, row_cap=1 << 14)
    p_big.process(dev, numbers, swap_rb=True)
    r_big = p_big.finish()
    p_small = VideoPipeline(small, 30.0, thr, row_cap=1 << 14)
    for s in range(0, n, 8):
        p_small.process(dev[s:s + 8], numbers[s:s + 8], swap_rb=True)
    r_small = p_small.finish()
    assert len(r_big['rows']) > 0
    assert np.array_equal(r_big['rows'], r_small['rows'])
    assert sorted(r_big['phases']) == sorted(r_small['phases'])
    for tid in r_big['phases']:
        a = [(p.time_start, p.time_end, p.rom, p.type) for p in r_big['phases'][tid]]
        b = [(p.time_start, p.time_end, p.rom, p.type) for p in r_small['phases'][tid]]
        assert a == b


def test_config4_videos_sharded_and_gathered():
    """Several clips of different lengths through shard.track_videos (one rank here; the
    multi-rank plan and gather are covered on CPU by tests/test_shard_gloo.py): per-video tables
    equal the oracle chain, tracker state never leaks from one video into the next."""
    import torch
    from vbt_b200 import shard
    from vbt_b200.interpreter import Detector
    g = graph('lite0')
    det = Detector(g, max_batch=4)
    lens = [5, 9, 3, 7]
    vids = [{'fps': 30.0 if i % 2 == 0 else 60.0,
             'frames': torch.as_tensor(synthetic_frames(n, 135, 240, seed=20 + i), device='cuda')}
            for i, n in enumerate(lens)]
    thr = 0.3
    tables, phases = shard.track_videos(vids, det, detection_threshold=thr, row_cap=4096)
    assert sorted(tables) == list(range(len(lens)))
    for i, v in enumerate(vids):
        want = oracle_rows(g, v['frames'].cpu().numpy(), v['fps'], thr)
        assert tables[i].shape == want.shape, i
        assert np.array_equal(tables[i], want), i
    # frame stride (HEAD's `frame_count % 16`, here 2): frames 2, 4, 6, ... keep their numbers
    t2, _ = shard.track_videos(vids[1:2], det, detection_threshold=thr, frame_stride=2, row_cap=4096)
    f = vids[1]['frames'].cpu().numpy()
    want = oracle_rows(g, f[1::2], 60.0, thr, numbers=list(range(2, 10, 2)))
    assert np.array_equal(t2[0], want)


def test_config3_lite2_bf16_heads_equal_int8_heads():
    """configs[3], "int8 vs bf16 heads": the class / box nets' pointwise convs on the bf16 tensor
    path (tcgen05.mma.kind::f16, fp32 accumulate) must reproduce the int8 path's raw outputs bit
    for bit at frame batch 256 -- and both are the oracle's on the frames it is run on."""
    import torch
    from vbt_b200 import effdet
    from vbt_b200.interpreter import Detector
    g8 = graph('lite2')
    g16 = effdet.build_synthetic('lite2', head_dtype='bf16')
    n = 256
    frames = synthetic_frames(n, 135, 240, seed=10)
    dev = torch.as_tensor(frames, device='cuda')
    d8, d16 = Detector(g8, max_batch=n), Detector(g16, max_batch=n)
    plan = d16.plan()
    heads = [i for i, op in enumerate(g16.ops) if op.branch > 0]
    assert all(plan[i] in (0, 2) for i in heads), 'every head stage must run in the fused kernel'
    d8.network(d8.preprocess(dev, True))
    d16.network(d16.preprocess(dev, True))
    cls8, box8 = d8.raw_cls[:n, :g8.n_anchors].cpu().numpy(), d8.raw_box[:n, :g8.n_anchors].cpu().numpy()
    cls16, box16 = d16.raw_cls[:n, :g8.n_anchors].cpu().numpy(), d16.raw_box[:n, :g8.n_anchors].cpu().numpy()
    assert np.array_equal(cls8, cls16) and np.array_equal(box8, box16)
    pick = [0, n - 1]
    imgs = OR.preprocess_batch(frames[pick], g8.S, swap_rb=True)
    want_cls, want_box, _ = OE.run(g16, imgs)
    assert np.array_equal(cls16[pick], want_cls) and np.array_equal(box16[pick], want_box)


def test_one_video_in_frame_chunks_equals_one_pass():
    """shard.track_video_chunks: detection on contiguous frame chunks (here three chunks run one
    after the other on this GPU, then concatenated as the gather would), tracker + velocity over
    the whole table -- rows and phases equal the ordinary one-pass pipeline and the oracle chain.
    The collective itself is covered on CPU by tests/test_shard_gloo.py."""
    import torch
    from vbt_b200 import shard
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = graph('lite0')
    det = Detector(g, max_batch=8)
    n, thr, fps = 37, 0.3, 30.0
    frames_np = synthetic_frames(n, 135, 240, seed=31)
    frames = torch.as_tensor(frames_np, device='cuda')
    video = {'fps': fps, 'frames': frames}
    one = shard.track_video_chunks(video, det, detection_threshold=thr, row_cap=4096)     # world of 1
    want = oracle_rows(g, frames_np, fps, thr)
    assert len(want) > 0 and np.array_equal(one['rows'], want)
    for stride, world in ((1, 3), (2, 2)):
        keep = torch.arange(stride, n + 1, stride, dtype=torch.int32)
        loaded = []

        def load(a, b):
            loaded.append((a, b))
            return frames[a:b]
        vid = {'fps': fps, 'n_frames': n, 'load': load}
        parts = []
        for lo, hi in shard.chunk_bounds(len(keep), world):
            pipe = VideoPipeline(det, fps, thr, row_cap=4096)
            parts.append(shard.detect_chunk(pipe, vid, keep[lo:hi], stride))
        table = [torch.cat([p[i] for p in parts]) for i in range(3)]
        assert table[2].cpu().tolist() == keep.tolist()
        assert all(b - a <= 8 * stride for a, b in loaded)           # a rank only loads its own batches
        pipe = VideoPipeline(det, fps, thr, row_cap=4096)
        pipe.track_table(*table)
        res = pipe.finish()
        ref = VideoPipeline(det, fps, thr, row_cap=4096)
        for s in range(0, len(keep), 8):
            idx = keep[s:s + 8].long() - 1
            ref.process(frames[idx.cuda()].contiguous(), keep[s:s + 8].cuda(), swap_rb=True)
        r = ref.finish()
        assert np.array_equal(res['rows'], r['rows'])
        if stride == 1:
            assert np.array_equal(res['rows'], want)
        assert sorted(res['phases']) == sorted(r['phases'])
        for tid in res['phases']:
            assert [(p.time_start, p.time_end, p.rom, p.type) for p in res['phases'][tid]] == \
                   [(p.time_start, p.time_end, p.rom, p.type) for p in r['phases'][tid]]


def test_track_cli_writes_the_annotated_video(tmp_path):
    """--video_dir (track.py:152-154,241-242): one mp4v frame per processed frame that had a detection
    >= threshold; the per-row overlay inputs (tracker output box + score, track.py:190) equal the
    oracle tracker's return values; the written frames carry the white overlay."""
    import cv2
    from vbt_b200.interpreter import Interpreter
    from vbt_b200.track import track
    g = graph('lite0')
    path = str(tmp_path / 'clip.avi')
    frames = synthetic_frames(10, 240, 320, seed=6)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'MJPG'), 30.0, (320, 240))
    for f in frames:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(path)
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    decoded = np.stack(decoded)
    interp = Interpreter(model_path=g, num_threads=4)
    thr = 0.3
    out = str(tmp_path / 'clip.mp4')
    data, res = track(path, interp, thr, video_path=out, frame_stride=1, batch=4, return_pipeline_result=True)
    # overlay inputs vs the oracle: tracker.update()'s own (xmin,ymin,xmax,ymax,...,score) per emitted row
    imgs = OR.preprocess_batch(decoded, g.S, swap_rb=True)
    cls, box, _ = OE.run(g, imgs)
    trk = oo.OCSortOracle(max_age=30, iou_threshold=0.1)
    want, with_results = [], []
    for b in range(len(decoded)):
        ob, _, osc, cnt, _ = OP.detection_postprocess(cls[b], box[b], g.anchors(), g.box_scale, g.box_zp, min_score_q=-128)
        d = OP.tracker_inputs(OP.detect_results(ob, osc, cnt, thr)).reshape(-1, 6)
        if len(d) == 0:
            continue
        with_results.append(b + 1)
        for row in trk.update(d):
            want.append([row[0], row[1], row[2], row[3], row[6]])
    assert res['frames_with_results'] == with_results and len(with_results) > 0
    assert np.array_equal(res['details'], np.asarray(want, dtype=np.float64).reshape(-1, 5))
    assert len(res['details']) == len(data['id']) > 0
    # the file: frame count, size, and overlay pixels (mp4v is lossy: compare loosely)
    cap = cv2.VideoCapture(out)
    n_written = 0
    first = None
    while True:
        ok, f = cap.read()
        if not ok:
            break
        first = f if first is None else first
        n_written += 1
    cap.release()
    assert n_written == len(with_results) and first.shape == (240, 320, 3)
    t0 = min(data['time'])
    r0 = [i for i, t in enumerate(data['time']) if t == t0][0]
    xmin, ymin, xmax, ymax, _ = res['details'][r0]
    x0, x1, y0 = max(int(xmin * 320), 0), min(int(xmax * 320), 319), int(ymin * 240)
    if 2 <= y0 < 238 and x1 - x0 > 8:
        edge = first[y0, x0 + 3:x1 - 3].astype(int)
        assert edge.mean() > 200, 'the top edge of the first box should be (nearly) white'


def test_next_video_overlaps_and_equals_finish_reset():
    """VideoPipeline.next_video(): three clips back to back without draining between them give,
    per clip, exactly what finish() + reset() gives (fresh tracker and velocity state per video,
    track.py:88-101,157), whatever order the handles are collected in."""
    import torch
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = graph('lite0')
    det = Detector(g, max_batch=8)
    thr = 0.3
    clips = [(30.0, synthetic_frames(n, 135, 240, seed=40 + i)) for i, n in enumerate((19, 8, 27, 5))]
    clips[1] = (60.0, clips[1][1])

    def feed(pipe, frames):
        dev = torch.as_tensor(frames, device='cuda')
        for s in range(0, len(frames), 8):
            e = min(len(frames), s + 8)
            pipe.process(dev[s:e], torch.arange(s + 1, e + 1, dtype=torch.int32, device='cuda'), swap_rb=True)

    ref_pipe = VideoPipeline(det, clips[0][0], thr, row_cap=4096)
    want = []
    for fps, frames in clips:
        ref_pipe.reset(fps)
        feed(ref_pipe, frames)
        want.append(ref_pipe.finish())
    pipe = VideoPipeline(det, clips[0][0], thr, row_cap=4096)
    handles = []
    for i, (fps, frames) in enumerate(clips):
        if i:
            handles.append(pipe.next_video(fps))
        feed(pipe, frames)
    got = [None] * len(clips)
    got[-1] = pipe.finish()
    for i in reversed(range(len(handles))):          # out of order on purpose
        got[i] = handles[i].result()
    assert sum(len(w['rows']) for w in want) > 0
    for i, (w, r) in enumerate(zip(want, got)):
        assert np.array_equal(w['rows'], r['rows']), i
        assert sorted(w['phases']) == sorted(r['phases']), i
        for tid in w['phases']:
            assert [(p.time_start, p.time_end, p.rom, p.type) for p in w['phases'][tid]] == \
                   [(p.time_start, p.time_end, p.rom, p.type) for p in r['phases'][tid]]
    assert np.array_equal(oracle_rows(g, clips[2][1], 30.0, thr), got[2]['rows'])


# ---- many videos per GPU, lane capacity, the bench configuration -----------------------------------

def _phase_tuples(phases):
    return {tid: [(p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type) for p in ps] for tid, ps in phases.items()}


@pytest.mark.parametrize('V', [2, 4])
def test_videos_sharing_batches_equal_videos_alone(V):
    """V videos whose frames travel in the same detection batches (K7: one warp per video, K8: one lane
    per (video, id)) give exactly what each video gives alone -- and what the CPU oracle gives.  The
    unit of independence is the video (track.py:88-101,157)."""
    import torch
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = graph('lite0')
    F, f, n, thr, fps = 8, 8 // V, 24, 0.3, 30.0
    det = Detector(g, max_batch=F)
    clips = [synthetic_frames(n, 135, 240, seed=40 + v) for v in range(V)]
    alone = []
    for v in range(V):
        pipe = VideoPipeline(det, fps, thr, row_cap=4096)
        for s in range(0, n, F):
            pipe.process(torch.as_tensor(clips[v][s:s + F], device='cuda'),
                         torch.arange(s + 1, min(s + F, n) + 1, dtype=torch.int32, device='cuda'), swap_rb=True)
        alone.append(pipe.finish())
        assert np.array_equal(alone[v]['rows'], oracle_rows(g, clips[v], fps, thr))
    assert any(len(a['rows']) for a in alone)
    pipe = VideoPipeline(det, fps, thr, row_cap=4096, n_videos=V)
    for s in range(0, n, f):                               # video-major batches: f frames of each video
        batch = np.concatenate([clips[v][s:s + f] for v in range(V)])
        numbers = np.concatenate([np.arange(s + 1, s + f + 1) for _ in range(V)]).astype(np.int32)
        pipe.process(torch.as_tensor(batch, device='cuda'), torch.as_tensor(numbers, device='cuda'), swap_rb=True)
    handle = pipe.next_video()
    together = handle.result()
    assert len(together) == V
    for v in range(V):
        assert np.array_equal(together[v]['rows'], alone[v]['rows']), v
        assert _phase_tuples(together[v]['phases']) == _phase_tuples(alone[v]['phases'])
        assert together[v]['path'] == alone[v]['path']


def test_more_tracks_than_velocity_lanes_is_a_capacity_error():
    """A track id beyond `id_lanes` would get no phases: VBT_ECAPACITY instead of a silent loss."""
    import torch
    from vbt_b200 import _lib
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = graph('lite0')
    det = Detector(g, max_batch=8)
    pipe = VideoPipeline(det, 30.0, 0.3, row_cap=4096, id_lanes=2)
    D = det.max_det
    dets = torch.zeros((6, D, 6), dtype=torch.float64, device='cuda')
    for i, (x, y) in enumerate([(0.1, 0.1), (0.5, 0.5), (0.8, 0.2)]):          # three well separated plates
        dets[:, i] = torch.tensor([x, y, x + 0.1, y + 0.1, 0.9, 0.0], dtype=torch.float64)
    counts = torch.full((6,), 3, dtype=torch.int32, device='cuda')
    numbers = torch.arange(1, 7, dtype=torch.int32, device='cuda')
    pipe.track_table(dets, counts, numbers)
    with pytest.raises(_lib.VbtError) as e:
        pipe.finish()
    assert e.value.code == _lib.ECAPACITY
    ok = VideoPipeline(det, 30.0, 0.3, row_cap=4096, id_lanes=3)
    ok.track_table(dets, counts, numbers)
    assert sorted(ok.finish()['phases']) == [1, 2, 3]


def test_bench_configuration_matches_the_oracle():
    """The BENCH line's exact path: Lite0, 1080p, frame batch 64, two detection lanes, CUDA-graph replay,
    device-resident frames AND row-sparse ingest from pinned host memory -- rows and phases of three
    batches (two replays of each lane's graph) against the CPU oracle chain (~20 s of CPU)."""
    import torch
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    from vbt_b200.synth import plate_trajectory, render_clip
    g = graph('lite0')
    B, n, fps, thr = 64, 160, 30.0, 0.5
    det = Detector(g, max_batch=B)
    clip = render_clip(n, 1080, 1920, seed=0, device='cuda', trajectory=plate_trajectory(1800, fps, seed=0)[:n])
    numbers = torch.arange(1, n + 1, dtype=torch.int32, device='cuda')
    results = []
    for mode in ('device', 'host'):
        pipe = VideoPipeline(det, fps, thr, row_cap=1 << 14, n_lanes=2)
        assert len(pipe.detectors) == 2
        if mode == 'host':
            assert pipe.use_row_sparse_ingest(1080, 1920) is not None
            host = [torch.empty((B, 1080, 1920, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for rep in range(2):                              # second pass replays every lane's captured graph
            for i, s in enumerate(range(0, n, B)):
                e = min(s + B, n)
                if mode == 'device':
                    src = clip[s:e]
                else:
                    if pipe.input_consumed is not None:
                        pipe.input_consumed.synchronize()
                    host[i % 2][:e - s].copy_(clip[s:e])
                    src = host[i % 2][:e - s]
                pipe.process(src, numbers[s:e], swap_rb=True)
            res = pipe.next_video().result() if rep == 0 else pipe.finish()
            results.append(res)
    frames_np = clip.cpu().numpy()
    want = oracle_rows(g, frames_np, fps, thr)
    assert len(want) >= n // 2                              # the plate (and some spurious boxes) are tracked
    rows_by_id = {}
    for tid in np.unique(want[:, 0]).astype(int):
        rows_by_id[int(tid)] = ov.analyze_series(want[want[:, 0] == tid][:, 1:], 0.45)
    for res in results:
        assert np.array_equal(res['rows'], want)
        assert sorted(res['phases']) == sorted(rows_by_id)
        for tid, ps in res['phases'].items():
            got = np.array([[p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type] for p in ps]).reshape(-1, 6)
            assert np.array_equal(got, rows_by_id[tid])
