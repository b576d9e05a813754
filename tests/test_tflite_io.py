"""`.tflite` import / export on the CPU: the FlatBuffer / FlexBuffer codec,
the writer and the reader against each other.  No file written by TensorFlow exists in this
container (the reference's six models are listed in .MISSING_LARGE_BLOBS), so what is pinned here
is self-consistency: a model exported and re-imported computes bit-identical outputs."""

import numpy as np
import pytest

from oracle import effdet as OE
from vbt_b200 import effdet as E, flatbuf as F, tflite_reader as R, tflite_schema as S, tflite_writer as W
from vbt_b200.synth import synthetic_model_inputs

_cache = {}


def lite0():
    if 'g' not in _cache:
        g = E.build_synthetic('lite0')
        _cache['g'] = (g, W.graph_to_tflite(g))
    return _cache['g']


def test_flexbuffer_map_round_trip():
    d = dict(max_detections=25, max_classes_per_detection=1, use_regular_nms=False, nms_iou_threshold=0.5,
             nms_score_threshold=-3.4028234663852886e38, num_classes=1, y_scale=1.0, w_scale=1.0)
    got = F.flex_map(F.flex_build_map(d))
    assert set(got) == set(d)
    for k, v in d.items():
        assert type(got[k]) is type(v)
        assert got[k] == (np.float32(v) if isinstance(v, float) else v)


def test_flatbuffer_builder_reader_round_trip():
    t = F.T(f0=('i32', -7), f2='hello', f3=F.V('i64', [1, -2, 3]), f5=[F.T(f0=('u8', 200)), F.T(f1=F.V('f32', [0.5]))],
            f6=('f32', 1.25), f7=('i8', -3))
    buf = F.build(t, b'TEST')
    r = F.root(buf, b'TEST')
    assert r.scalar(0, 'i32') == -7 and r.scalar(1, 'i32', 42) == 42 and r.string(2) == 'hello'
    assert list(r.vector(3, 'i64')) == [1, -2, 3] and r.vector(4, 'i32').size == 0
    subs = r.tables(5)
    assert subs[0].scalar(0, 'u8') == 200 and list(subs[1].vector(1, 'f32')) == [0.5]
    assert r.scalar(6, 'f32') == 1.25 and r.scalar(7, 'i8') == -3
    with pytest.raises(ValueError):
        F.root(buf, b'TFL3')


def test_exported_file_is_a_tflite_detection_graph():
    g, buf = lite0()
    assert buf[4:8] == b'TFL3'
    tensors, ops, sg_in, sg_out = R.parse(buf)
    assert tensors[sg_in[0]].type == S.UINT8 and tensors[sg_in[0]].shape == [1, 320, 320, 3]
    assert len(sg_out) == 4 and tensors[sg_out[0]].shape == [1, 25, 4]
    hist = {}
    for code, custom, *_ in ops:
        hist[S.OP_NAMES[code]] = hist.get(S.OP_NAMES[code], 0) + 1
    n_conv = sum(op.type in (E.OP_STEM, E.OP_PW) for op in g.ops)
    n_dw = sum(op.type == E.OP_DW for op in g.ops)
    assert hist['CONV_2D'] == n_conv and hist['DEPTHWISE_CONV_2D'] == n_dw
    assert hist['QUANTIZE'] == 1 and hist['LOGISTIC'] == 1 and hist['CUSTOM'] == 1 and hist['CONCATENATION'] == 2
    assert hist['RESHAPE'] == 10 and hist['DEQUANTIZE'] == 2
    assert ops[-1][1] == S.DETECTION_POSTPROCESS
    opts = F.flex_map(ops[-1][5])
    assert opts['max_detections'] == 25 and opts['nms_iou_threshold'] == 0.5 and opts['num_classes'] == 1
    # per-channel filter quantisation, int32 bias with scale s_in * s_w
    conv = next(o for o in ops if o[0] == S.CONV_2D)
    w, b = tensors[conv[2][1]], tensors[conv[2][2]]
    assert w.type == S.INT8 and w.scale.size == w.shape[0] and b.type == S.INT32
    assert np.allclose(b.scale, tensors[conv[2][0]].scale[0] * w.scale, rtol=1e-6)


def test_import_of_the_export_computes_the_same_network():
    g, buf = lite0()
    g2 = R.tflite_to_graph(buf)
    assert g2.S == g.S and g2.n_anchors == g.n_anchors and g2.level_sizes == g.level_sizes
    assert (g2.box_scale, g2.box_zp, g2.cls_scale, g2.cls_zp) == (g.box_scale, g.box_zp, g.cls_scale, g.cls_zp)
    assert np.array_equal(g2.anchors(), g.anchors())
    x = synthetic_model_inputs(1, g.S, seed=3)
    c1, b1, _ = OE.run(g, x)
    c2, b2, _ = OE.run(g2, x)
    assert np.array_equal(c1, c2) and np.array_equal(b1, b2)
    # the head chains come back as the ten independent branches
    assert sorted({op.branch for op in g2.ops}) == list(range(11))
    E.pack_blob(g2)                                    # and the library's model format accepts it


def test_reader_rejects_what_it_cannot_run():
    g, buf = lite0()
    with pytest.raises(ValueError):
        R.tflite_to_graph(b'\0' * 16)
    bad = bytearray(buf)
    bad[4:8] = b'XXXX'
    with pytest.raises(ValueError):
        R.tflite_to_graph(bytes(bad))
    # a graph with an operator outside the supported set
    t = F.T(f0=('u32', 3), f1=[F.T(f3=('i32', 9))],   # FULLY_CONNECTED
            f2=[F.T(f0=[F.T(f0=F.V('i32', [1, 320, 320, 3]), f1=('i8', S.UINT8), f3='in',
                            f4=F.T(f2=F.V('f32', [1 / 128]), f3=F.V('i64', [127])))],
                    f1=F.V('i32', [0]), f2=F.V('i32', [0]),
                    f3=[F.T(f0=('u32', 0), f1=F.V('i32', [0]), f2=F.V('i32', [0]))])],
            f4=[F.T()])
    with pytest.raises(R.TfliteError):
        R.tflite_to_graph(F.build(t, b'TFL3'))
