"""Per-video pipeline on one GPU: frame batches -> detections -> track rows -> phases.

This is the body of the reference's hot loop (track.py:159-247) re-cut for a GPU: instead
of one frame per iteration through four libraries, a batch of frames already in HBM goes
through K1 (preprocess), the network, K6 (post-process), the threshold/packing kernel,
K7 (tracker, sequential over the batch's frames inside one warp) and K8 (velocity lanes,
streaming) without leaving the device.  Only `finish()` copies the row table and the
phases to the host, where the DataFrame of track.py:103-126 is assembled with pandas.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from .interpreter import Detector, score_to_q
from .ocsort import BatchedTracker
from .velocity import _Lanes, _phases_from

COLUMNS = ['id', 'time', 'x', 'y', 'dx', 'dy', 'norm_plate_height', 'norm_plate_width']


class _VideoState:
    """Tracker + velocity state of ONE video -- or of `n_videos` videos that advance in lock step
    (their frames share the detection batches; K7 runs one warp per video, K8 one lane per (video,
    id)).  The pipeline owns two sets and alternates, so the next video's batches enter the device
    while the previous video's tail (last tracker steps, end_processing, read-back) is still in
    flight -- `VideoPipeline.next_video`."""

    def __init__(self, t, row_cap, id_lanes, keep_details, tracker_kw, fps=30.0, n_videos=1):
        # the tracker recurrence is one warp per video and nearly as long as a detection step: each
        # video gets its own (high-priority) stream, so the tail of one video's recurrence runs
        # beside the head of the next one's instead of in front of it
        self.side = t.cuda.Stream(priority=-1)
        # K8 (velocity lanes) of batch i only needs K7's rows of batch i: on its own stream it runs beside K7 of
        # batch i + 1 instead of in front of it (the tracker writes rows before their count, velocity.cu reads
        # [lane_begin, row_count) -- whichever count it sees, the rows below it are complete)
        self.side2 = t.cuda.Stream(priority=-1)
        self.d_fps = t.full((n_videos,), float(fps), dtype=t.float64, device='cuda')
        self.tracker = BatchedTracker(n_videos, row_cap=row_cap, keep_details=keep_details, **(tracker_kw or {}))
        self.lanes = _Lanes(n_videos * id_lanes, path_cap=min(row_cap, 1 << 15))
        # lane v * id_lanes + l follows id l + 1 of video v (row table v)
        self.lane_table = t.arange(n_videos, dtype=t.int32, device='cuda').repeat_interleave(id_lanes).contiguous()
        self.lane_begin = t.zeros(n_videos * id_lanes, dtype=t.int32, device='cuda')
        self.frame_log = []             # (frame numbers, detection counts) per batch, for the overlay export
        self.pending = None             # _PendingVideo not collected yet


class _PendingVideo:
    """A finished video whose results are still on the device; `result()` waits for its own tail
    only (not for the videos that followed) and builds what `finish()` returns."""

    def __init__(self, pipe, state, done):
        self.pipe, self.state, self.done, self._res = pipe, state, done, None

    def result(self):
        if self._res is None:
            t = self.pipe.torch
            t.cuda.current_stream().wait_event(self.done)
            self._res = self.pipe._collect(self.state)
            self.state.pending = None
        return self._res


class VideoPipeline:
    """One video on one GPU.

    Streams: `n_lanes` detection lanes (each its own stream, CUDA graph and detector
    buffers) take the batches round robin, so two batches of the same video are in flight
    and one lane's small late-network kernels overlap the other's large early ones;
    tracking and velocity (K7, K8: a sequential recurrence over frames on one warp) follow
    in batch order on the video's `side` stream, off the detector's critical path
    (SURVEY.md 7.3-6).  Nothing returns to the host before `finish()` / a `next_video()` handle."""

    # ring of tracker-input slots between the lanes and the side streams: deep enough that detection
    # never waits for the tracker inside a 60 s clip (1.2 kB per frame and slot)
    N_SLOTS = max(2, int(os.environ.get('VBT_SLOTS', '32')))

    def __init__(self, detector: Detector, fps, detection_threshold=0.5, plate_diameter=0.45,
                 row_cap=1 << 17, id_lanes=64, tracker_kw=None, diff_threshold=0.6,
                 min_distance=0.1, n_lanes=None, keep_details=False, n_videos=1):
        """n_videos = V > 1: V videos advance together.  Every batch handed to `process` holds the
        same number of frames of each, video-major ([v0 f0..fk, v1 f0..fk, ...]); `finish()` /
        `next_video().result()` return a list of V result dicts.  The videos are independent
        (track.py:88-101,157: fresh tracker per video) -- sharing batches changes no result, it
        turns K7's one-warp recurrence into V warps of a V-times shorter one."""
        self.torch = t = _lib.require_cuda()
        if n_lanes is None:
            n_lanes = max(1, int(os.environ.get('VBT_LANES', '2')))
        self.det = detector
        self.fps = float(fps)
        self.threshold = float(detection_threshold)
        self.plate_diameter, self.diff_threshold, self.min_distance = plate_diameter, diff_threshold, min_distance
        self.F = detector.max_batch
        self.keep_details = keep_details
        self.id_lanes = id_lanes
        self.V = int(n_videos)
        if self.V < 1 or detector.max_batch % self.V:
            raise ValueError(f'n_videos={n_videos} must divide the detector batch {detector.max_batch}')
        self._state_args = (t, row_cap, id_lanes, keep_details, tracker_kw, self.fps, self.V)
        self.states = [_VideoState(*self._state_args), _VideoState(*self._state_args)]
        self.cur = 0
        D = detector.max_det
        S = self.N_SLOTS
        self.dets = t.zeros((S, 1, self.F, D, 6), dtype=t.float64, device='cuda')
        self.det_count = t.zeros((S, 1, self.F), dtype=t.int32, device='cuda')
        self.frame_no = t.zeros((S, 1, self.F), dtype=t.int32, device='cuda')
        self.n_frames = t.zeros((S, self.V), dtype=t.int32, device='cuda')
        self.lane_id = t.arange(1, id_lanes + 1, dtype=t.int32, device='cuda').repeat(self.V).contiguous()
        self.n_velocity_lanes = self.V * id_lanes
        self.detectors = [detector] + [Detector(detector.source, max_batch=detector.max_batch,
                                                iou_threshold=detector.iou_threshold,
                                                max_det=detector.max_det) for _ in range(n_lanes - 1)]
        self.det_streams = [t.cuda.Stream() for _ in range(n_lanes)]
        self.active_lanes = n_lanes                  # bench.py profiles with a single lane
        self.det_ready = [t.cuda.Event() for _ in range(S)]
        self.slot_free = [t.cuda.Event() for _ in range(S)]
        self.input_consumed = None     # event: the last batch's frames have been read (K1 / DMA done)
        self.ingest = None             # ingest.RowSparseIngest for host frames (use_row_sparse_ingest)
        self.batches = 0
        self.last_slot = 0
        self.frames_done = 0
        self.stage_events = None      # bench.py: list of per-step event tuples when profiling

    # the current video's state (what process() feeds)
    tracker = property(lambda self: self.states[self.cur].tracker)
    lanes = property(lambda self: self.states[self.cur].lanes)
    lane_table = property(lambda self: self.states[self.cur].lane_table)
    lane_begin = property(lambda self: self.states[self.cur].lane_begin)
    _frame_log = property(lambda self: self.states[self.cur].frame_log)
    side = property(lambda self: self.states[self.cur].side)
    d_fps = property(lambda self: self.states[self.cur].d_fps)

    def use_row_sparse_ingest(self, H, W):
        """Host frames of this size are sent row-sparse (only the rows the resize reads)."""
        from .ingest import RowSparseIngest
        if (W * 3) % 16 or (H * W * 3) % 16:       # K1's bulk-copy path needs 16-byte rows
            self.ingest = None
        else:
            self.ingest = RowSparseIngest(self.F, H, W, self.det.S)
        return self.ingest

    def _mark(self, marks, stream=None):
        if marks is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append(e)

    def _sync_streams(self):
        cur = self.torch.cuda.current_stream()
        for s in self.det_streams:
            cur.wait_stream(s)
        for st in self.states:
            cur.wait_stream(st.side)
            cur.wait_stream(st.side2)

    def reset(self, fps=None):
        t = self.torch
        self._sync_streams()
        if fps is not None:
            self.fps = float(fps)
            self.d_fps.fill_(self.fps)
        self.tracker.reset()
        self.lanes.reset()
        self.lane_begin.zero_()
        self._frame_log.clear()
        self.frames_done = 0
        cur = t.cuda.current_stream()
        for s in self.det_streams:
            s.wait_stream(cur)
        for st in self.states:
            st.side.wait_stream(cur)
            st.side2.wait_stream(cur)

    def process(self, frames, frame_numbers, swap_rb=True, track=True):
        """frames: uint8 CUDA [n,H,W,3] (n <= detector.max_batch); frame_numbers: int32 CUDA
        tensor [n] with the 1-based frame_count of each (track.py:161).  Both must be ready on
        the current stream; `self.input_consumed` is recorded once K1 has read `frames`.
        track=False: detection only (K1 -> network -> K6 -> pack); the packed tracker inputs of
        the batch are kept on the device for `detection_table()` (frame-chunk sharding of one
        long video, shard.track_video_chunks: the tracker runs later, on the gathered table)."""
        t = self.torch
        n = frames.shape[0]
        lane = self.batches % self.active_lanes
        k = self.batches % self.N_SLOTS
        self.batches += 1
        self.last_slot = k
        det, ds = self.detectors[lane], self.det_streams[lane]
        marks = [] if self.stage_events is not None else None
        arrived = t.cuda.Event()
        arrived.record()
        ds.wait_event(arrived)
        table = None
        if frames.is_cuda:
            frames.record_stream(ds)
        elif self.ingest is not None and tuple(frames.shape[1:3]) == (self.ingest.H, self.ingest.W):
            # host frames: DMA only the rows K1 touches (ingest.py), on the ingest stream
            table, n, tslot, ready = self.ingest.upload(frames)
            self.input_consumed = ready            # the host buffer is free once the DMA is done
            ds.wait_event(ready)
        # else: pinned host frames are read in place by K1 (zero copy)
        frame_numbers.record_stream(ds)
        with t.cuda.stream(ds):
            self._mark(marks)
            if table is not None:
                images = det.preprocess_rows(self.ingest, table, n, swap_rb)
                self.ingest.free[tslot] = t.cuda.Event()
                self.ingest.free[tslot].record(ds)
            else:
                images = det.preprocess(frames, swap_rb)
                self.input_consumed = t.cuda.Event()
                self.input_consumed.record(ds)
            self._mark(marks)
            det.network(images)
            self._mark(marks)
            boxes, _, scores, count, _ = det.postprocess(n, score_to_q(self.threshold))
            self._mark(marks)
            ds.wait_event(self.slot_free[k])           # K7 has read this slot's previous batch
            _lib.check(_lib.lib().vbt_pack_detections(
                boxes.data_ptr(), scores.data_ptr(), count.data_ptr(), n, det.max_det,
                self.threshold, self.dets[k].data_ptr(), self.det_count[k].data_ptr(),
                _lib.stream_ptr(ds)))
            self.frame_no[k, 0, :n].copy_(frame_numbers, non_blocking=True)
            if n % self.V:
                raise ValueError(f'a batch of {n} frames cannot hold equally many of {self.V} videos')
            self.n_frames[k].fill_(n // self.V)
            if self.keep_details:
                self._frame_log.append((self.frame_no[k, 0, :n].clone(), self.det_count[k, 0, :n].clone()))
            self._mark(marks)
            if not track:
                if not hasattr(self, '_table'):
                    self._table = []
                self._table.append((self.dets[k, 0, :n].clone(), self.det_count[k, 0, :n].clone(),
                                    self.frame_no[k, 0, :n].clone()))
                self.slot_free[k].record(ds)
                self.frames_done += n
                return
            self.det_ready[k].record(ds)
        side = self.side
        side.wait_event(self.det_ready[k])             # batches reach the side stream in order
        with t.cuda.stream(side):
            self._mark(marks, side)
            # video-major batch: the first n frame slots viewed as [V, n / V, ...] are each video's own
            f = n // self.V
            self.tracker.update(self.dets[k, 0, :n].view(self.V, f, -1, 6), self.det_count[k, 0, :n].view(self.V, f),
                                self.frame_no[k, 0, :n].view(self.V, f), self.d_fps, self.n_frames[k], stream=side)
            self._mark(marks, side)
            self.slot_free[k].record(side)
            tracked = t.cuda.Event()
            tracked.record(side)
        side2 = self.states[self.cur].side2
        side2.wait_event(tracked)
        with t.cuda.stream(side2):
            self.lanes.update(self.tracker.rows, self.tracker.row_count, self.tracker.row_cap,
                              self.lane_table, self.lane_id, self.lane_begin, self.n_velocity_lanes,
                              self.plate_diameter, self.diff_threshold, self.min_distance,
                              smooth=True, finish=False)
            self._mark(marks, side2)
        if marks is not None:
            self.stage_events.append(marks)
        self.frames_done += n

    def detection_table(self):
        """(dets f64 [n,25,6], counts i32 [n], frame numbers i32 [n]) of every batch processed with
        track=False since the last call, in order, on the device."""
        t = self.torch
        self._sync_streams()
        parts, self._table = getattr(self, '_table', []), []
        D = self.det.max_det
        if not parts:
            return (t.zeros((0, D, 6), dtype=t.float64, device='cuda'), t.zeros(0, dtype=t.int32, device='cuda'),
                    t.zeros(0, dtype=t.int32, device='cuda'))
        return tuple(t.cat([p[i] for p in parts]) for i in range(3))

    def track_table(self, dets, counts, frame_numbers):
        """K7 + K8 over a detection table (the output of `detection_table`, possibly gathered from
        several ranks), in frame order, `max_batch` frames per tracker call."""
        t = self.torch
        assert self.V == 1, 'track_table replays ONE video\'s detection table'
        self._sync_streams()
        side = self.side
        side.wait_stream(t.cuda.current_stream())
        n_all = int(dets.shape[0])
        with t.cuda.stream(side):
            for s in range(0, n_all, self.F):
                n = min(self.F, n_all - s)
                k = self.batches % self.N_SLOTS
                self.batches += 1
                self.dets[k, 0, :n].copy_(dets[s:s + n])
                self.det_count[k, 0, :n].copy_(counts[s:s + n])
                self.frame_no[k, 0, :n].copy_(frame_numbers[s:s + n])
                self.n_frames[k].fill_(n)
                self.tracker.update(self.dets[k], self.det_count[k], self.frame_no[k], self.d_fps,
                                    self.n_frames[k], stream=side)
                self.lanes.update(self.tracker.rows, self.tracker.row_count, self.tracker.row_cap,
                                  self.lane_table, self.lane_id, self.lane_begin, self.n_velocity_lanes,
                                  self.plate_diameter, self.diff_threshold, self.min_distance,
                                  smooth=True, finish=False)
                self.slot_free[k].record(side)
        for x in (dets, counts, frame_numbers):
            x.record_stream(side)

    def finish(self):
        """End of video: run end_processing() on every lane, bring results to the host.
        Returns dict(rows=f64[n,8] append order, phases={id: [Phase]}, path={id: float}).
        Synchronous: drains the whole pipeline.  `next_video()` is the overlapped form."""
        self._sync_streams()
        st = self.states[self.cur]
        st.lanes.update(st.tracker.rows, st.tracker.row_count, st.tracker.row_cap,
                        st.lane_table, self.lane_id, st.lane_begin, self.n_velocity_lanes,
                        self.plate_diameter, self.diff_threshold, self.min_distance,
                        smooth=True, finish=True)
        return self._collect(st)

    def next_video(self, fps=None):
        """End of the current video WITHOUT draining the pipeline: its end_processing() is queued
        behind its last tracker step on the side stream, the other state set is reset there too, and
        the caller may feed the next video's frames at once.  Returns a handle whose `result()` gives
        what `finish()` would have returned; a set's previous handle is collected before reuse."""
        t = self.torch
        st = self.states[self.cur]
        st.side2.wait_stream(st.side)
        with t.cuda.stream(st.side2):
            st.lanes.update(st.tracker.rows, st.tracker.row_count, st.tracker.row_cap,
                            st.lane_table, self.lane_id, st.lane_begin, self.n_velocity_lanes,
                            self.plate_diameter, self.diff_threshold, self.min_distance,
                            smooth=True, finish=True)
            done = t.cuda.Event()
            done.record(st.side2)
        st.pending = pending = _PendingVideo(self, st, done)
        self.cur ^= 1
        nxt = self.states[self.cur]
        if nxt.pending is not None:
            nxt.pending.result()
        nxt.side.wait_stream(t.cuda.current_stream())      # result() read the old tables on the current stream
        if fps is not None:
            self.fps = float(fps)
        with t.cuda.stream(nxt.side):
            nxt.tracker.reset()
            nxt.lanes.reset()
            nxt.lane_begin.zero_()
            nxt.d_fps.fill_(self.fps)
        nxt.frame_log.clear()
        self.frames_done = 0
        return pending

    def _collect(self, st):
        st.tracker.check_status()
        phases, count, state = st.lanes.read()
        results = []
        for v in range(self.V):
            rows = st.tracker.rows_host(v)
            # every id the tracker emitted needs a velocity lane (lane l follows id l + 1); an id past the
            # lanes would silently get no phases and no path length -- the reference's plot.py can
            # analyse any id (plot.py:88), so this is a capacity error, not a truncation
            if len(rows) and int(rows[:, 0].max()) > self.id_lanes:
                raise _lib.VbtError(_lib.ECAPACITY,
                                    f'track id {int(rows[:, 0].max())} exceeds the {self.id_lanes} velocity lanes of '
                                    f'this VideoPipeline; construct it with id_lanes >= the number of tracks born')
            out_ph, out_path = {}, {}
            for l in range(self.id_lanes):
                j = v * self.id_lanes + l
                if state[j, 2] > 0:
                    out_ph[l + 1] = _phases_from(phases[j], int(count[j]))
                    out_path[l + 1] = float(state[j, 3])
            out = dict(rows=rows, phases=out_ph, path=out_path)
            if self.keep_details:
                out['details'] = st.tracker.details[v, :len(rows)].cpu().numpy()
                t = self.torch
                if st.frame_log and self.V == 1:
                    nos = t.cat([a for a, _ in st.frame_log]).cpu().numpy()
                    cnt = t.cat([b for _, b in st.frame_log]).cpu().numpy()
                    out['frames_with_results'] = [int(f) for f, c in zip(nos, cnt) if c > 0]
                else:
                    out['frames_with_results'] = []
            results.append(out)
        return results[0] if self.V == 1 else results


def rows_to_data(rows):
    """Row table -> the dict-of-lists `track()` returns (track.py:144-145, 227-234)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    data = {'id': [int(v) for v in rows[:, 0]]}
    for j, c in enumerate(COLUMNS[1:], start=1):
        data[c] = [np.float64(v) for v in rows[:, j]]
    return data


def export_dataframe(data, src, model, df_dir=None):
    """track.py:103-126: DataFrame, (id,time) sort, per-id cumulative Euclidean path,
    file name from the id with the largest path, gzip pickle.  Returns (df, path)."""
    import os
    import pandas as pd
    df = pd.DataFrame.from_dict(data)
    df = df.sort_values(by=['id', 'time'])
    df2 = df.copy()
    same = df2['id'] == df2['id'].shift()
    step = ((df2['x'] - df2['x'].shift()) ** 2 + (df2['y'] - df2['y'].shift()) ** 2) ** 0.5
    df2['distance'] = np.where(same, step, np.nan)
    df2['cumulative_distance'] = df2.groupby('id')['distance'].cumsum()
    max_distance_id = df2.loc[df2['cumulative_distance'].idxmax(), 'id']
    model_name = os.path.basename(model).split('.')[0].replace(':', '_')
    name = f'{os.path.basename(src).split(".")[0]}_id{max_distance_id}_{model_name}.pkl.gz'
    path = name if df_dir is None else os.path.join(df_dir, name)
    df.to_pickle(path)
    return df, path
