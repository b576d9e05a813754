#!/usr/bin/env python
"""Real-weight known-answer harness: the reference's eval.py detection pass on the GPU path.

    python scripts/eval_kat.py --img_dir data/test --annotations_dir data/test \
        --golden dfs/eval_detections.pkl.gz models/efficientdet_lite0_whole.tflite [more .tflite ...]

Restates eval.py:156-215 (`create_detections_df`) over vbt_b200's drop-in surfaces: one
`Interpreter(model_path, num_threads)` per model (eval.py:167-168), `cv2.imread` frames fed to
`run_odt(..., threshold=0)` as they are -- BGR, NO colour swap (the quirk at eval.py:173) --, boxes
scaled to integer pixels by truncation (`scaled_bbox`, eval.py:57-71), Hungarian matching against the
Pascal-VOC ground truth (`match_bboxes`, eval.py:96-153), one (Score, Model, IoU) row per matched
detection.  With `--golden` the rows are compared, per model and as multisets (the reference iterates
glob order), with the reference's own cached run `dfs/eval_detections.pkl.gz`, and the AP / AUC legends
of docs/precision_recall_iou_0.75.png / docs/roc_iou_0.75.png are recomputed from both
(eval.py:515, 228-240: sklearn average_precision_score / roc_auc_score on Label = IoU > 0.75).

The `.tflite` blobs are absent from the reference checkout (.MISSING_LARGE_BLOBS): until someone
supplies them this runs on exported synthetic models (tests/test_gpu_eval_kat.py), which exercises every
step but cannot reproduce the golden numbers.  Tolerances for real weights (BASELINE.json north star):
|dScore| <= 1e-2, box IoU >= 0.99 against the reference box -- on integer-pixel boxes that is
|dIoU-with-ground-truth| of a few 1e-2 at most; the report prints the observed maxima."""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LABEL = 'barbell'                                    # eval.py:32, train.py:34


def load_annotations(annotations_dir):
    """{image file name: int [n,4] (ymin,xmin,ymax,xmax)} from Pascal-VOC xml (eval.py:488-504)."""
    out = {}
    for path in sorted(glob.glob(os.path.join(annotations_dir, '*.xml'))):
        root = ET.parse(path).getroot()
        boxes = []
        for obj in root.findall('object'):
            if obj.find('name').text != LABEL:
                continue
            bb = obj.find('bndbox')
            boxes.append([int(bb.find(k).text) for k in ('ymin', 'xmin', 'ymax', 'xmax')])
        out[root.find('filename').text] = np.array(boxes, dtype=int).reshape(-1, 4)
    return out


def to_pixels(box, height, width):
    """Normalised (ymin,xmin,ymax,xmax) -> integer pixels, truncating like ndarray.astype(int)."""
    return (np.asarray(box, dtype=np.float64) * np.array([height, width, height, width], dtype=np.float64)).astype(int)


def iou(a, b):
    ih = max(0, min(a[2], b[2]) - max(a[0], b[0]))
    iw = max(0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = ih * iw
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union > 0 else 0


def match(gt, det):
    """Best one-to-one assignment of detections to ground-truth boxes by IoU (Hungarian on 1 - IoU over
    the matrix padded square with IoU 0); returns (detection indices, IoUs) of the real detections that
    were assigned a row."""
    from scipy.optimize import linear_sum_assignment
    n_gt, n_det = len(gt), len(det)
    n = max(n_gt, n_det)
    m = np.zeros((n, n))
    for i in range(n_gt):
        for j in range(n_det):
            m[i, j] = iou(det[j], gt[i])
    rows, cols = linear_sum_assignment(1 - m)
    keep = cols < n_det
    return cols[keep], m[rows[keep], cols[keep]]


def detections_table(models, img_dir, annotations, threads=4):
    """[(score, model name, IoU)] -- eval.py:156-215."""
    import cv2
    from vbt_b200.interpreter import Interpreter
    from vbt_b200.odt import run_odt
    files = sorted(glob.glob(os.path.join(img_dir, '*.jpg')))
    per_model = {}
    for m in models:
        interp = Interpreter(model_path=m, num_threads=threads)
        interp.allocate_tensors()
        dets = {}
        for f in files:
            img = cv2.imread(f)                              # BGR, fed unswapped (eval.py:173)
            h, w = img.shape[:2]
            res = run_odt(frame=img, interpreter=interp, threshold=0)
            dets[os.path.basename(f)] = [(to_pixels(r['bounding_box'], h, w), float(r['score'])) for r in res]
        per_model[os.path.basename(m).split('.')[0]] = dets
    rows = []
    for name, gt in annotations.items():
        for model, dets in per_model.items():
            if name not in dets:
                continue
            boxes = np.array([b for b, _ in dets[name]]).reshape(-1, 4)
            idx, ious = match(gt, boxes)
            for j, v in zip(idx, ious):
                rows.append((dets[name][j][1], model, float(v)))
    return rows


def legends(rows, iou_threshold=0.75):
    """{model: (AP, AUC)} as the reference's figures label them (eval.py:515, 240, plot_roc)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    out = {}
    for model in sorted({r[1] for r in rows}):
        s = np.array([r[0] for r in rows if r[1] == model])
        y = np.array([r[2] > iou_threshold for r in rows if r[1] == model])
        if y.all() or not y.any():
            out[model] = (float('nan'), float('nan'))
        else:
            out[model] = (float(average_precision_score(y, s)), float(roc_auc_score(y, s)))
    return out


def compare_with_golden(rows, golden_path):
    """Per model: multiset distance between our (Score, IoU) rows and the reference's cached ones."""
    import pandas as pd
    gold = pd.read_pickle(golden_path)
    report = {}
    for model in sorted({r[1] for r in rows}):
        g = gold[gold['Model'] == model]
        if len(g) == 0:
            report[model] = {'golden_rows': 0}
            continue
        ours_s = np.sort(np.array([r[0] for r in rows if r[1] == model]))
        ours_i = np.sort(np.array([r[2] for r in rows if r[1] == model]))
        gs, gi = np.sort(g['Score'].to_numpy()), np.sort(g['IoU'].to_numpy())
        rep = {'golden_rows': int(len(g)), 'our_rows': int(len(ours_s))}
        if len(gs) == len(ours_s):
            rep['max_abs_dscore_sorted'] = float(np.abs(gs - ours_s).max())
            rep['max_abs_diou_sorted'] = float(np.abs(gi - ours_i).max())
            rep['scores_identical'] = bool(np.array_equal(gs, ours_s))
        report[model] = rep
    gold_rows = [(float(s), m, float(i)) for s, m, i in zip(gold['Score'], gold['Model'], gold['IoU'])]
    return report, legends(gold_rows)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('models', nargs='+')
    ap.add_argument('--img_dir', default='data/test')
    ap.add_argument('--annotations_dir', default='data/test')
    ap.add_argument('--golden', default=None, help='dfs/eval_detections.pkl.gz of the reference')
    ap.add_argument('--iou_threshold', type=float, default=0.75)
    ap.add_argument('--threads', type=int, default=4)
    ap.add_argument('--export', default=None, help='write our (Score, Model, IoU) DataFrame here (pickle)')
    args = ap.parse_args()
    for m in args.models:
        if not os.path.isfile(m):
            raise FileNotFoundError(m)
    ann = load_annotations(args.annotations_dir)
    rows = detections_table(args.models, args.img_dir, ann, args.threads)
    out = {'rows': len(rows), 'images': len(ann),
           'legends': {m: {'AP': a, 'AUC': u} for m, (a, u) in legends(rows, args.iou_threshold).items()}}
    if args.export:
        import pandas as pd
        pd.DataFrame({'Score': [r[0] for r in rows], 'Model': [r[1] for r in rows], 'IoU': [r[2] for r in rows]}).to_pickle(args.export)
    if args.golden:
        rep, gold_leg = compare_with_golden(rows, args.golden)
        out['vs_golden'] = rep
        out['golden_legends'] = {m: {'AP': a, 'AUC': u} for m, (a, u) in gold_leg.items()}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
