"""scripts/eval_kat.py -- the reference's eval.py detection pass (eval.py:156-215) restated over the GPU
path, the harness that turns `dfs/eval_detections.pkl.gz` into a detector known-answer test the moment
the `.tflite` blobs are supplied.

CPU: matching, pixel truncation and the AP / AUC legends -- the latter reproduce, from the reference's own
cached run, the numbers printed in docs/precision_recall_iou_0.75.png and docs/roc_iou_0.75.png.
GPU: the whole harness on an exported synthetic model over generated jpgs + Pascal-VOC annotations."""
import importlib.util
import os

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location('eval_kat', os.path.join(ROOT, 'scripts', 'eval_kat.py'))
K = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(K)

# legends of the reference's figures (README.md:53-54)
PUBLISHED = {'efficientdet_lite0': (0.7156, 0.9833), 'efficientdet_lite0_whole': (0.9529, 0.9969),
             'efficientdet_lite1': (0.8333, 0.9753), 'efficientdet_lite1_whole': (0.9333, 0.9878),
             'efficientdet_lite2': (0.7836, 0.9865), 'efficientdet_lite2_whole': (0.9358, 0.9952)}


def test_pixel_boxes_truncate_and_iou():
    assert K.to_pixels([0.1, 0.2, 0.55, 0.999], 416, 416).tolist() == [41, 83, 228, 415]      # 41.6 -> 41, 415.6 -> 415
    assert K.to_pixels([-0.01, 0.0, 1.2, 1.0], 100, 200).tolist() == [-1, 0, 120, 200]        # not clipped (odt.py:66)
    assert K.iou([0, 0, 10, 10], [0, 0, 10, 10]) == 1.0
    assert K.iou([0, 0, 10, 10], [5, 5, 15, 15]) == 25 / 175
    assert K.iou([0, 0, 10, 10], [20, 20, 30, 30]) == 0 and K.iou([0, 0, 0, 0], [0, 0, 0, 0]) == 0


def test_matching_is_one_to_one_and_optimal():
    gt = np.array([[0, 0, 10, 10], [20, 20, 40, 40]])
    det = np.array([[21, 19, 41, 39], [100, 100, 110, 110], [1, 0, 10, 11], [0, 0, 10, 10]])
    idx, ious = K.match(gt, det)
    assert len(idx) == 4 and sorted(idx.tolist()) == [0, 1, 2, 3]       # every real detection gets a row (padded square)
    by_det = dict(zip(idx.tolist(), ious.tolist()))
    assert by_det[3] == 1.0 and by_det[0] > 0.8                         # the perfect box wins gt 0, det 0 takes gt 1
    assert by_det[1] == 0.0 and by_det[2] == 0.0                        # the rest land on padding rows: IoU 0
    idx, ious = K.match(np.zeros((0, 4), int), det[:2])                 # no ground truth: all unmatched, IoU 0
    assert len(idx) == 2 and not ious.any()


def test_legends_reproduce_the_published_figures_from_the_reference_cache():
    path = os.path.join(helpers.REFERENCE, 'dfs', 'eval_detections.pkl.gz')
    if not os.path.exists(path):
        pytest.skip('reference checkout not present (GPU box)')
    import pandas as pd
    g = pd.read_pickle(path)
    rows = [(float(s), m, float(i)) for s, m, i in zip(g['Score'], g['Model'], g['IoU'])]
    assert len(rows) == 9150                                            # 6 models x 61 images x 25 detections
    got = K.legends(rows, 0.75)
    assert {m: (round(a, 4), round(u, 4)) for m, (a, u) in got.items()} == PUBLISHED
    ann = K.load_annotations(os.path.join(helpers.REFERENCE, 'data', 'test'))
    assert len(ann) == 61 and sum(len(v) for v in ann.values()) == 105


@pytest.mark.gpu
def test_harness_end_to_end_on_a_synthetic_model(tmp_path):
    import cv2
    from vbt_b200 import effdet, tflite_writer
    g = effdet.build_synthetic('lite0')
    model = str(tmp_path / 'efficientdet_lite0_synthetic.tflite')
    tflite_writer.save(g, model)
    rng = np.random.default_rng(7)
    for i, (h, w) in enumerate([(416, 416), (416, 416), (480, 270)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cv2.imwrite(str(tmp_path / f'img{i}.jpg'), img)
        boxes = ''.join(f'<object><name>barbell</name><bndbox><xmin>{x0}</xmin><ymin>{y0}</ymin><xmax>{x0 + 80}</xmax>'
                        f'<ymax>{y0 + 90}</ymax></bndbox></object>' for x0, y0 in [(20 + 30 * i, 40), (150, 200)][:1 + i % 2])
        (tmp_path / f'img{i}.xml').write_text(f'<annotation><filename>img{i}.jpg</filename>{boxes}'
                                              f'<object><name>person</name><bndbox><xmin>1</xmin><ymin>1</ymin><xmax>5</xmax><ymax>5</ymax></bndbox></object></annotation>')
    ann = K.load_annotations(str(tmp_path))
    assert sorted(ann) == ['img0.jpg', 'img1.jpg', 'img2.jpg'] and [len(ann[f'img{i}.jpg']) for i in range(3)] == [1, 2, 1]
    rows = K.detections_table([model], str(tmp_path), ann)
    # threshold 0 keeps all 25 detections of every image (the property dfs/eval_detections.pkl.gz shows)
    assert len(rows) == 3 * 25 and {r[1] for r in rows} == {'efficientdet_lite0_synthetic'}
    scores = np.array([r[0] for r in rows])
    assert np.all(scores * 256 == np.round(scores * 256)) and scores.min() >= 0 and scores.max() <= 1
    assert all(0.0 <= r[2] <= 1.0 for r in rows)
    # the same table again, from the in-memory graph instead of the file: identical rows
    rows2 = K.detections_table(['synthetic:lite0'], str(tmp_path), ann)
    assert [(a[0], a[2]) for a in rows] == [(b[0], b[2]) for b in rows2]
    leg = K.legends(rows)
    assert list(leg) == ['efficientdet_lite0_synthetic']
