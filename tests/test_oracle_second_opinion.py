"""Independent second opinions on the oracles that no reference vector can pin in this container
(tensorflow, tflite_runtime and the .tflite blobs are absent -- DESIGN.md section 2):

* oracle/resize.py (restated TF ResizeBilinear, half-pixel centres, truncating cast) against
  torch.nn.functional.interpolate(bilinear, align_corners=False, antialias=False): a separately
  written implementation of the same sampling rule (it blends as (1 - l) * a + l * b where TF and
  the oracle compute a + (b - a) * l, so single results may sit on the other side of an integer);
* oracle/effdet.py's convolution + requantisation against PyTorch's quantized CPU engine
  (oracle/effdet_q.py: fbgemm / oneDNN int8 kernels, fp32 requantisation -- the XNNPACK family);
* oracle/postprocess.py against a brute-force restatement written here (decode every anchor, full
  stable sort, greedy suppression) and hand-computed boxes.
None of these imports vbt_b200 for arithmetic; the graph is only data."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers
from oracle import effdet as OE, effdet_q as OQ, postprocess as OP, resize as OR


def _interp(frame, S):
    x = torch.from_numpy(frame.astype(np.float32)).permute(2, 0, 1)[None]
    y = F.interpolate(x, size=(S, S), mode='bilinear', align_corners=False, antialias=False)
    return y[0].permute(1, 2, 0).numpy()


@pytest.mark.parametrize('S', [320, 384, 448])
def test_resize_oracle_vs_torch_interpolate_1080p(S):
    rng = np.random.default_rng(S)
    f = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    a = OR.resize_bilinear_u8(f, S).astype(np.int16)
    b = _interp(f, S)
    d = np.abs(a - b.astype(np.uint8).astype(np.int16))
    assert d.max() <= 1                                   # never more than one step
    assert (d == 0).mean() >= 0.999                       # measured: 1.0 / 1.0 / 0.99999
    assert np.abs(b - a).max() < 1.01                     # the truncated value is the float's floor


def test_resize_oracle_vs_torch_interpolate_on_the_reference_test_images():
    """The 61 data/test jpgs the north star names as parity inputs (416x416 and 1080x1920 portrait)."""
    files = sorted(glob.glob(os.path.join(helpers.REFERENCE, 'data', 'test', '*.jpg')))
    if not files:
        pytest.skip('reference checkout not present (GPU box): data/test is read in place, never copied')
    import cv2
    tot = bad = 0
    for p in files:
        im = cv2.imread(p)
        a = OR.resize_bilinear_u8(im, 320).astype(np.int16)
        d = np.abs(a - _interp(im, 320).astype(np.uint8).astype(np.int16))
        assert d.max() <= 1
        tot += d.size
        bad += int((d > 0).sum())
    assert len(files) == 61 and bad / tot < 1e-3          # measured: 9.0e-5


def test_conv_requant_oracle_vs_quantized_engine():
    """Every STEM / PW / DW op of Lite0 (172 convs, 33 M outputs over two frames) on the oracle's own
    input tensors: the exact oracle and the production int8 engine agree on all but a few results in a
    million, and never differ by more than one quantisation step (the engine forms acc * M in a
    different order)."""
    from vbt_b200 import effdet as E            # data only: the synthetic quantised graph
    from vbt_b200.synth import synthetic_model_inputs
    g = E.build_synthetic('lite0')
    x = synthetic_model_inputs(2, g.S, seed=3)
    _, _, tensors = OE.run(g, x, keep=True)
    res = OQ.check_convs(g, x, tensors)
    assert len(res) == sum(1 for op in g.ops if op.type in (E.OP_STEM, E.OP_PW, E.OP_DW) and op.out >= 0) == 172
    total = sum(r[1] for r in res)
    bad = sum(r[2] for r in res)
    assert total > 30e6
    assert max(r[3] for r in res) <= 1
    assert bad / total < 1e-4                             # measured: 1.6e-6
    # every kind of op is covered, residual projections included
    names = {r[0] for r in res}
    assert {'stem', 'b1.0.dw', 'b2.1.project', 'b3.0.dw', 'fpn0.n3.pw', 'cls3.0.dw'} <= names


def _brute_postprocess(cls_q, box_q, anchors, box_scale, box_zp, max_det=25, iou_thr=0.5):
    """TFLite_Detection_PostProcess, fast NMS, num_classes = 1, written the long way."""
    score = (cls_q.astype(np.int32) + 128) / np.float32(256.0)
    t = (box_q.astype(np.float32) - np.float32(box_zp)) * np.float32(box_scale)
    ya, xa, ha, wa = [anchors[:, i].astype(np.float32) for i in range(4)]
    yc = t[:, 0] * ha + ya
    xc = t[:, 1] * wa + xa
    h = np.exp(t[:, 2]).astype(np.float32) * ha
    w = np.exp(t[:, 3]).astype(np.float32) * wa
    half = np.float32(0.5)
    boxes = np.stack([yc - half * h, xc - half * w, yc + half * h, xc + half * w], axis=1).astype(np.float32)
    order = np.argsort(-score, kind='stable')
    keep = []
    for i in order:
        ok = True
        for j in keep:
            a, b = boxes[i], boxes[j]
            area_a = (a[2] - a[0]) * (a[3] - a[1])
            area_b = (b[2] - b[0]) * (b[3] - b[1])
            if area_a <= 0 or area_b <= 0:
                continue
            ih = min(a[2], b[2]) - max(a[0], b[0])
            iw = min(a[3], b[3]) - max(a[1], b[1])
            inter = max(ih, np.float32(0)) * max(iw, np.float32(0))
            if inter / (area_a + area_b - inter) > iou_thr:
                ok = False
                break
        if ok:
            keep.append(int(i))
            if len(keep) == max_det:
                break
    return np.asarray(keep), boxes[keep], score[keep]


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_postprocess_oracle_vs_brute_force(seed):
    from oracle import arch
    a = arch.anchors('lite0')                              # the independent anchor table
    rng = np.random.default_rng(seed)
    n = len(a)
    # a few hundred candidates over a floor of low scores, clustered boxes: plenty of ties and overlaps
    cls = np.full(n, -128, np.int8)
    hot = rng.choice(n, 400, replace=False)
    cls[hot] = rng.integers(-100, 127, 400).astype(np.int8)
    box = rng.integers(-30, 30, (n, 4)).astype(np.int8)
    box_scale, box_zp = float(np.float32(0.03)), 2
    ob, _, osc, cnt, oi = OP.detection_postprocess(cls, box, a, box_scale, box_zp, min_score_q=-128)
    keep, kb, ks = _brute_postprocess(cls, box, a, box_scale, box_zp)
    assert int(cnt) == len(keep) == 25
    assert np.array_equal(np.asarray(oi[:25]), keep)       # detection indices: exact
    assert np.array_equal(osc[:25], ks.astype(np.float32))
    assert np.allclose(ob[:25], kb, rtol=0, atol=2e-6)     # same formulas, exp() from a table vs libm


def test_postprocess_decode_hand_computed():
    """One anchor, encodings chosen by hand: centre moves by ty * ha, size scales by exp(th)."""
    a = np.array([[0.5, 0.5, 0.2, 0.4]], np.float32)
    scale, zp = float(np.float32(0.1)), 0
    box = np.array([[10, -5, 0, 7]], np.int8)              # ty = 1.0, tx = -0.5, th = 0, tw = 0.7
    cls = np.array([64], np.int8)                          # (64 + 128) / 256 = 0.75
    ob, _, osc, cnt, oi = OP.detection_postprocess(cls, box, a, scale, zp, min_score_q=-128)
    yc, xc, h, w = 0.5 + 1.0 * 0.2, 0.5 - 0.5 * 0.4, 0.2, 0.4 * np.exp(np.float32(0.7))
    assert int(cnt) == 1 and osc[0] == np.float32(0.75) and oi[0] == 0
    assert np.allclose(ob[0], [yc - h / 2, xc - w / 2, yc + h / 2, xc + w / 2], atol=1e-6)
