"""The whole reference hot path on the CPU, chained from the oracle pieces -- TEST
INFRASTRUCTURE and the timed CPU arm of bench.py (`cpu_baseline`, `--impl reference`).

One frame at a time, like track.py's loop (track.py:159-247): BGR->RGB + resize
(oracle/resize.py) -> int8 EfficientDet-Lite (oracle/effdet.py, batch 1, mirrors the
per-frame interpreter invoke) -> TFLite_Detection_PostProcess (oracle/postprocess.py) ->
threshold + packing -> OC-SORT (oracle/ocsort.py) -> rows; then per id the plot.py
smoothing + VelocityTracker (oracle/velocity.py).

This is a PORT (the reference's own TFLite/XNNPACK int8 path is not installable here and
its weights are absent), so bench.py reports it with "kind": "port".
"""
import numpy as np

from oracle import effdet as OE, ocsort as oo, postprocess as OP, resize as OR, velocity as ov


class CpuPipeline:
    def __init__(self, graph, fps, threshold=0.5, plate_diameter=0.45):
        self.g, self.fps, self.thr, self.plate = graph, float(fps), threshold, plate_diameter
        self.anchors = graph.anchors()
        self.reset()

    def reset(self):
        self.tracker = oo.OCSortOracle(max_age=30, iou_threshold=0.1)
        self.rows = []

    def step(self, frame_bgr, frame_no):
        g = self.g
        img = OR.resize_bilinear_u8(frame_bgr, g.S, swap_rb=True)[None]
        cls, box, _ = OE.run(g, img)
        ob, _, osc, cnt, _ = OP.detection_postprocess(cls[0], box[0], self.anchors, g.box_scale,
                                                      g.box_zp)
        dets = OP.tracker_inputs(OP.detect_results(ob, osc, cnt, self.thr)).reshape(-1, 6)
        if len(dets) == 0:
            return 0
        time = frame_no / self.fps
        out = self.tracker.update(dets)
        for r in out:
            xmin, ymin, xmax, ymax = r[:4]
            self.rows.append([int(r[4]), time, (xmin + xmax) / 2, (ymin + ymax) / 2, r[7], r[8],
                              abs(ymax - ymin), abs(xmax - xmin)])
        return len(out)

    def finish(self):
        rows = np.array(self.rows, dtype=np.float64).reshape(-1, 8)
        phases = {}
        for tid in np.unique(rows[:, 0]).astype(int):
            phases[int(tid)] = ov.analyze_series(rows[rows[:, 0] == tid][:, 1:], self.plate)
        return rows, phases
