"""CPU oracle for detection post-processing (row a5p of SURVEY.md section 8) -- TEST INFRASTRUCTURE.

Restates the TFLite_Detection_PostProcess custom op (tflite-runtime 2.14.0,
requirements.txt:381; source not in the reference checkout [3P-MEM]) in the mode the
EfficientDet-Lite export uses: fast class-agnostic NMS, max_detections 25,
nms_score_threshold -inf, IoU 0.5, scales 1, no clipping; then odt.py:64-75 (threshold
filter) and odt.py:102-118 (tracker input packing).

Parity status: WEAKLY PINNED -- dfs/eval_detections.pkl.gz confirms 25 outputs per image
at threshold 0 and scores on a 1/256 grid (tests/test_oracle_postprocess.py checks both
against the fixture statistics committed in tests/golden/eval_detections_stats.json); box
numerics are unpinned (no interpreter, no weights in this container).

float32 in the op's operation order (numpy keeps multiply and add separately rounded).
"""
import numpy as np


def exp_lut(box_scale, box_zp):
    """exp(dequant(q)) for q=-128..127 as float32: the 256 values exp() can see."""
    q = np.arange(-128, 128, dtype=np.int32)
    v = np.float32(box_scale) * (q - box_zp).astype(np.float32)
    return np.exp(v).astype(np.float32)


def decode_boxes(box_q, anchors, box_scale, box_zp, lut=None):
    """box_q int8 [N,4] (ty,tx,th,tw); anchors f32 [N,4] (yc,xc,h,w) -> f32 [N,4]
    (ymin,xmin,ymax,xmax).  DecodeCenterSizeBoxes with scale values 1."""
    lut = exp_lut(box_scale, box_zp) if lut is None else lut
    q = box_q.astype(np.int32)
    s = np.float32(box_scale)
    ty = s * (q[:, 0] - box_zp).astype(np.float32)
    tx = s * (q[:, 1] - box_zp).astype(np.float32)
    a = anchors.astype(np.float32)
    yc = ty * a[:, 2] + a[:, 0]
    xc = tx * a[:, 3] + a[:, 1]
    hh = np.float32(0.5) * lut[q[:, 2] + 128] * a[:, 2]
    hw = np.float32(0.5) * lut[q[:, 3] + 128] * a[:, 3]
    return np.stack([yc - hh, xc - hw, yc + hh, xc + hw], axis=1).astype(np.float32)


def _iou(bi, bj):
    ai = (bi[2] - bi[0]) * (bi[3] - bi[1])
    aj = (bj[2] - bj[0]) * (bj[3] - bj[1])
    if ai <= 0 or aj <= 0:
        return np.float32(0.0)
    ymin, xmin = max(bi[0], bj[0]), max(bi[1], bj[1])
    ymax, xmax = min(bi[2], bj[2]), min(bi[3], bj[3])
    inter = max(ymax - ymin, np.float32(0.0)) * max(xmax - xmin, np.float32(0.0))
    return inter / (ai + aj - inter)


def detection_postprocess(cls_q, box_q, anchors, box_scale, box_zp, iou_threshold=0.5,
                          max_det=25, min_score_q=-128):
    """One frame.  cls_q int8 [N] (LOGISTIC output, scale 1/256, zp -128).
    Returns boxes f32 [max_det,4], classes f32 [max_det], scores f32 [max_det],
    count (float), index i32 [max_det] (-1 padded)."""
    boxes = decode_boxes(box_q, anchors, box_scale, box_zp)
    level = cls_q.astype(np.int32) + 128
    cand = np.nonzero(cls_q.astype(np.int32) >= min_score_q)[0]
    order = cand[np.argsort(-level[cand], kind='stable')]       # stable: ties by index
    thr = np.float32(iou_threshold)
    sel = []
    for i in order:
        if len(sel) >= max_det:
            break
        bi = boxes[i]
        if all(not (_iou(boxes[j], bi) > thr) for j in sel):
            sel.append(int(i))
    ob = np.zeros((max_det, 4), np.float32)
    oc = np.zeros(max_det, np.float32)
    osc = np.zeros(max_det, np.float32)
    oi = np.full(max_det, -1, np.int32)
    for k, i in enumerate(sel):
        ob[k] = boxes[i]
        osc[k] = np.float32(0.00390625) * np.float32(level[i])
        oi[k] = i
    return ob, oc, osc, np.float32(len(sel)), oi


def detect_results(boxes, scores, count, threshold):
    """odt.detect_objects lines 64-75: list of {'bounding_box', 'score'}."""
    return [{'bounding_box': boxes[i], 'score': scores[i]}
            for i in range(int(count)) if scores[i] >= threshold]


def tracker_inputs(results):
    """odt.results_to_sorttracker_inputs (odt.py:102-118)."""
    rows = []
    for r in results:
        ymin, xmin, ymax, xmax = r['bounding_box']
        rows.append(np.array([xmin, ymin, xmax, ymax, r['score'], 0]))
    return np.empty((0, 6)) if not rows else np.array(rows)
