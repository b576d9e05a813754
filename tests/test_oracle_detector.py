"""CPU checks of the detector-side oracles (no GPU): what the reference's own artefacts pin.

* architecture: the restated EfficientDet-Lite0/1/2 graphs reproduce the converter's MAC
  estimate printed at models/*.log:110 (0.876 / 1.773 / 3.033 GMAC) to within 2 % -- the
  remainder is adds / pools / resizes the converter also counts;
* anchors: 19,206 / 27,621 / 37,629 boxes;
* post-process: dfs/eval_detections.pkl.gz shows 25 detections per image at threshold 0
  and scores on a 1/256 grid (tests/golden/eval_detections_stats.json); the oracle op
  reproduces both on random inputs, orders by score with ties to the lower anchor index,
  and never keeps two boxes with IoU > 0.5;
* resize: identity at equal size, exact values on a hand-computed 2x2 -> 4x4 case,
  truncation (not rounding) of the float result.
"""
import json
import os

import numpy as np
import pytest

import helpers
from oracle import postprocess as OP, resize as OR
from vbt_b200 import effdet as E

LOG_GMAC = {'lite0': 0.876, 'lite1': 1.773, 'lite2': 3.033}      # models/*.log:110
ANCHORS = {'lite0': 19206, 'lite1': 27621, 'lite2': 37629}


@pytest.mark.parametrize('variant', ['lite0', 'lite1', 'lite2'])
def test_architecture_mac_count_matches_reference_logs(variant):
    g = E.Graph(variant)
    assert g.n_anchors == ANCHORS[variant]
    assert len(g.anchors()) == g.n_anchors
    gmac = g.macs() / 1e9
    assert abs(gmac - LOG_GMAC[variant]) / LOG_GMAC[variant] < 0.02, gmac
    assert gmac <= LOG_GMAC[variant]          # the converter counts extra non-MAC ops, never fewer


def test_postprocess_matches_eval_detections_statistics():
    with open(os.path.join(helpers.GOLDEN, 'eval_detections_stats.json')) as f:
        stats = json.load(f)
    g = E.anchors_only('lite0', box_scale=0.02, box_zp=3)
    a = g.anchors()
    rng = np.random.default_rng(0)
    cls = rng.integers(-128, 128, g.n_anchors).astype(np.int8)
    box = rng.integers(-40, 40, (g.n_anchors, 4)).astype(np.int8)
    ob, oc, osc, cnt, oi = OP.detection_postprocess(cls, box, a, g.box_scale, g.box_zp, min_score_q=-128)
    assert cnt == stats['detections_per_image'] == 25
    assert stats['scores_on_1_256_grid'] and np.all(osc * 256 == np.round(osc * 256))
    assert np.all(np.diff(osc) <= 0)
    # ties resolve to the lower anchor index (stable descending order)
    for i in range(24):
        if osc[i] == osc[i + 1]:
            assert oi[i] < oi[i + 1]

    def iou(p, q):
        ih = max(0.0, min(p[2], q[2]) - max(p[0], q[0])); iw = max(0.0, min(p[3], q[3]) - max(p[1], q[1]))
        inter = ih * iw
        ua = (p[2] - p[0]) * (p[3] - p[1]) + (q[2] - q[0]) * (q[3] - q[1]) - inter
        return inter / ua if ua > 0 else 0.0
    for i in range(25):
        for j in range(i + 1, 25):
            assert iou(ob[i], ob[j]) <= 0.5


def test_threshold_filter_and_tracker_inputs():
    boxes = np.array([[0.1, 0.2, 0.5, 0.6], [0.0, 0.0, 1.0, 1.0], [0.3, 0.3, 0.4, 0.4]], np.float32)
    scores = np.array([0.9, 0.5, 0.25], np.float32)
    res = OP.detect_results(boxes, scores, 3.0, 0.5)
    assert len(res) == 2                                        # score >= threshold (odt.py:71)
    t = OP.tracker_inputs(res).reshape(-1, 6)
    assert t.dtype == np.float64
    assert np.allclose(t[0], [0.2, 0.1, 0.6, 0.5, 0.9, 0.0])    # xmin,ymin,xmax,ymax,score,0 (odt.py:116)
    assert OP.tracker_inputs([]).reshape(-1, 6).shape == (0, 6)


def test_resize_identity_known_values_and_truncation():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    assert np.array_equal(OR.resize_bilinear_u8(img, 16), img)
    src = np.zeros((2, 2, 3), np.uint8)
    src[0, 0] = 10; src[0, 1] = 20; src[1, 0] = 30; src[1, 1] = 41
    out = OR.resize_bilinear_u8(src, 4)[..., 0]
    # half-pixel centres: src = (dst + 0.5) * 0.5 - 0.5 -> -0.25, 0.25, 0.75, 1.25 (clamped ends)
    want_row0 = [10, 12, 17, 20]          # 12.5 and 17.5 truncate
    assert list(out[0]) == want_row0
    assert out[3, 3] == 41 and out[1, 0] == 15 and out[2, 0] == 25
    lo, hi, lerp = OR.interpolation_weights(320, 1080)
    assert lo[0] == 1 and hi[0] == 2 and abs(lerp[0] - 0.1875) < 1e-6     # (0.5 * 3.375) - 0.5 = 1.1875
    flipped = OR.resize_bilinear_u8(img, 8, swap_rb=True)
    assert np.array_equal(flipped, OR.resize_bilinear_u8(img[..., ::-1], 8))


def test_network_oracle_float32_fast_path_is_exact():
    """oracle/effdet.py evaluates integer convolutions in float32 where partial sums stay below
    2^24 and in float64 elsewhere; forcing float64 everywhere must not change a single value."""
    from oracle import effdet as OE
    from vbt_b200.synth import synthetic_model_inputs
    g = E.build_synthetic('lite0')
    x = synthetic_model_inputs(1, g.S, seed=3)
    cls, box, _ = OE.run(g, x)
    OE.FORCE_F64 = True
    try:
        cls64, box64, _ = OE.run(g, x)
    finally:
        OE.FORCE_F64 = False
    assert np.array_equal(cls, cls64) and np.array_equal(box, box64)
