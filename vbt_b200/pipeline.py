"""Per-video pipeline on one GPU: frame batches -> detections -> track rows -> phases.

This is the body of the reference's hot loop (track.py:159-247) re-cut for a GPU: instead
of one frame per iteration through four libraries, a batch of frames already in HBM goes
through K1 (preprocess), the network, K6 (post-process), the threshold/packing kernel,
K7 (tracker, sequential over the batch's frames inside one warp) and K8 (velocity lanes,
streaming) without leaving the device.  Only `finish()` copies the row table and the
phases to the host, where the DataFrame of track.py:103-126 is assembled with pandas.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .interpreter import Detector, score_to_q
from .ocsort import BatchedTracker
from .velocity import _Lanes, _phases_from

COLUMNS = ['id', 'time', 'x', 'y', 'dx', 'dy', 'norm_plate_height', 'norm_plate_width']


class VideoPipeline:
    def __init__(self, detector: Detector, fps, detection_threshold=0.5, plate_diameter=0.45,
                 row_cap=1 << 17, id_lanes=32, tracker_kw=None, diff_threshold=0.6,
                 min_distance=0.1):
        self.torch = t = _lib.require_cuda()
        self.det = detector
        self.fps = float(fps)
        self.threshold = float(detection_threshold)
        self.plate_diameter, self.diff_threshold, self.min_distance = plate_diameter, diff_threshold, min_distance
        self.F = detector.max_batch
        self.tracker = BatchedTracker(1, row_cap=row_cap, **(tracker_kw or {}))
        self.id_lanes = id_lanes
        self.lanes = _Lanes(id_lanes, path_cap=min(row_cap, 1 << 15))
        D = detector.max_det
        # Two slots of tracker inputs: the detector fills slot i while K7/K8 still read slot
        # i^1 on the side stream (the recurrence over frames is latency-bound on one warp and
        # must stay off the detector's critical path).
        self.dets = t.zeros((2, 1, self.F, D, 6), dtype=t.float64, device='cuda')
        self.det_count = t.zeros((2, 1, self.F), dtype=t.int32, device='cuda')
        self.frame_no = t.zeros((2, 1, self.F), dtype=t.int32, device='cuda')
        self.n_frames = t.zeros((2, 1), dtype=t.int32, device='cuda')
        self.d_fps = t.tensor([self.fps], dtype=t.float64, device='cuda')
        self.lane_table = t.zeros(id_lanes, dtype=t.int32, device='cuda')
        self.lane_id = t.arange(1, id_lanes + 1, dtype=t.int32, device='cuda')
        self.lane_begin = t.zeros(id_lanes, dtype=t.int32, device='cuda')
        self.side = t.cuda.Stream()
        self.det_ready = [t.cuda.Event() for _ in range(2)]
        self.slot_free = [t.cuda.Event() for _ in range(2)]
        self.slot = 0
        self.last_slot = 0
        self.frames_done = 0
        self.stage_events = None      # bench.py: list of per-step event tuples when profiling

    def _mark(self, marks, stream=None):
        if marks is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append(e)

    def reset(self, fps=None):
        t = self.torch
        t.cuda.current_stream().wait_stream(self.side)
        if fps is not None:
            self.fps = float(fps)
            self.d_fps.fill_(self.fps)
        self.tracker.reset()
        self.lanes.reset()
        self.lane_begin.zero_()
        self.frames_done = 0
        self.side.wait_stream(t.cuda.current_stream())

    def process(self, frames, frame_numbers, swap_rb=True):
        """frames: uint8 CUDA [n,H,W,3] (n <= detector.max_batch); frame_numbers: int32 CUDA
        tensor [n] with the 1-based frame_count of each (track.py:161).
        Detection (K1, network, K6, packing) is enqueued on the current stream; tracking and
        velocity (K7, K8) follow on `self.side`, overlapping the next batch's detection."""
        t = self.torch
        main = t.cuda.current_stream()
        n = frames.shape[0]
        marks = [] if self.stage_events is not None else None
        self._mark(marks)
        det = self.det
        images = det.preprocess(frames, swap_rb)
        self._mark(marks)
        det.network(images)
        self._mark(marks)
        boxes, _, scores, count, _ = det.postprocess(n, score_to_q(self.threshold))
        self._mark(marks)
        k = self.slot
        self.slot ^= 1
        self.last_slot = k
        main.wait_event(self.slot_free[k])           # K7 of two batches ago has read slot k
        _lib.check(_lib.lib().vbt_pack_detections(
            boxes.data_ptr(), scores.data_ptr(), count.data_ptr(), n, self.det.max_det,
            self.threshold, self.dets[k].data_ptr(), self.det_count[k].data_ptr(),
            _lib.stream_ptr(main)))
        self.frame_no[k, 0, :n].copy_(frame_numbers, non_blocking=True)
        self.n_frames[k].fill_(n)
        self._mark(marks)
        self.det_ready[k].record(main)
        side = self.side
        side.wait_event(self.det_ready[k])
        with t.cuda.stream(side):
            self._mark(marks, side)
            self.tracker.update(self.dets[k], self.det_count[k], self.frame_no[k], self.d_fps,
                                self.n_frames[k], stream=side)
            self._mark(marks, side)
            self.lanes.update(self.tracker.rows, self.tracker.row_count, self.tracker.row_cap,
                              self.lane_table, self.lane_id, self.lane_begin, self.id_lanes,
                              self.plate_diameter, self.diff_threshold, self.min_distance,
                              smooth=True, finish=False)
            self._mark(marks, side)
            self.slot_free[k].record(side)
        if marks is not None:
            self.stage_events.append(marks)
        self.frames_done += n

    def finish(self):
        """End of video: run end_processing() on every lane, bring results to the host.
        Returns dict(rows=f64[n,8] append order, phases={id: [Phase]}, path={id: float})."""
        self.torch.cuda.current_stream().wait_stream(self.side)
        self.tracker.check_status()
        self.lanes.update(self.tracker.rows, self.tracker.row_count, self.tracker.row_cap,
                          self.lane_table, self.lane_id, self.lane_begin, self.id_lanes,
                          self.plate_diameter, self.diff_threshold, self.min_distance,
                          smooth=True, finish=True)
        phases, count, state = self.lanes.read()
        rows = self.tracker.rows_host(0)
        out_ph, out_path = {}, {}
        for l in range(self.id_lanes):
            if state[l, 2] > 0:
                out_ph[l + 1] = _phases_from(phases[l], int(count[l]))
                out_path[l + 1] = float(state[l, 3])
        return dict(rows=rows, phases=out_ph, path=out_path)


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
