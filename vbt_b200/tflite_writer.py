"""Quantised layer program (`effdet.Graph`) -> `.tflite` FlatBuffer.

Why a writer: the reference's models are `.tflite` files (track.py:68 `--model`, track.py:93
`Interpreter(model_path=...)`) and none ships with the checkout.  Exporting OUR synthetic
post-training-quantised EfficientDet-Lite as a standard TFLite graph (QUANTIZE, CONV_2D,
DEPTHWISE_CONV_2D, ADD, MAX_POOL_2D, RESIZE_NEAREST_NEIGHBOR, RESHAPE, CONCATENATION, LOGISTIC,
DEQUANTIZE, TFLite_Detection_PostProcess) gives (1) a file the drop-in `Interpreter(model_path=
'x.tflite')` and `track.py --model x.tflite` accept, (2) a round-trip check of the reader, and
(3) a way for a maintainer who HAS tflite_runtime to run the same weights through the real
interpreter and pin this repo's int8 semantics.  Schema: tflite_schema.py [3P-MEM].
"""
from __future__ import annotations

import numpy as np

from . import effdet as E
from . import tflite_schema as S
from .flatbuf import T, V, build, flex_build_map


class _Model:
    def __init__(self):
        self.tensors, self.buffers, self.ops, self.codes = [], [T()], [], []

    def buffer(self, arr):
        self.buffers.append(T(f0=V('u8', np.frombuffer(np.ascontiguousarray(arr).tobytes(), np.uint8))))
        return len(self.buffers) - 1

    def tensor(self, name, shape, ttype, scale=None, zp=None, qdim=0, data=None):
        quant = None
        if scale is not None:
            quant = T(f2=V('f32', np.atleast_1d(scale)), f3=V('i64', np.atleast_1d(zp)),
                      f6=('i32', int(qdim)) if qdim else None)
        self.tensors.append(T(f0=V('i32', shape), f1=('i8', ttype) if ttype else None,
                              f2=('u32', self.buffer(data) if data is not None else 0), f3=name, f4=quant))
        return len(self.tensors) - 1

    def code(self, builtin, custom=None):
        key = (builtin, custom)
        if key not in [c[0] for c in self.codes]:
            self.codes.append((key, T(f0=('i8', min(builtin, 127)), f1=custom, f2=('i32', 1), f3=('i32', builtin))))
        return [c[0] for c in self.codes].index(key)

    def op(self, builtin, inputs, outputs, opt_type=0, options=None, custom=None, custom_options=None):
        self.ops.append(T(f0=('u32', self.code(builtin, custom)), f1=V('i32', inputs), f2=V('i32', outputs),
                          f3=('u8', opt_type) if opt_type else None, f4=options,
                          f5=V('u8', np.frombuffer(custom_options, np.uint8)) if custom_options else None))


def graph_to_tflite(g: E.Graph, description='vbt_b200 synthetic EfficientDet-Lite (int8 PTQ)'):
    assert getattr(g, 'quantized', False)
    m = _Model()
    tin = g.tensors[g.input]
    a_per = E.NUM_SCALES * len(E.ASPECTS)
    images = m.tensor('serving_default_images:0', [1, tin.h, tin.w, 3], S.UINT8, tin.scale, tin.zp)
    tid = {g.input: m.tensor('images_int8', [1, tin.h, tin.w, 3], S.INT8, tin.scale, tin.zp - 128)}
    m.op(S.QUANTIZE, [images], [tid[g.input]])

    def act_tensor(t, name=None):
        tt = g.tensors[t]
        return m.tensor(name or tt.name, [1, tt.h, tt.w, tt.c], S.INT8, tt.scale, tt.zp)

    cls_parts, box_parts = [], []
    for op in g.ops:
        q = op.q
        act = S.ACT_RELU6 if op.act else S.ACT_NONE
        if op.type in (E.OP_STEM, E.OP_PW, E.OP_DW):
            src = g.tensors[op.inputs[0]]
            w = q['w']
            if op.type == E.OP_DW:
                wt = m.tensor(op.name + '/w', [1, op.k, op.k, w.shape[0]], S.INT8, q['w_scale'],
                              np.zeros(w.shape[0], np.int64), qdim=3, data=np.transpose(w, (1, 2, 0))[None])
            else:
                w4 = w if op.type == E.OP_STEM else w[:, None, None, :]
                wt = m.tensor(op.name + '/w', list(w4.shape), S.INT8, q['w_scale'],
                              np.zeros(w.shape[0], np.int64), qdim=0, data=w4)
            bt = m.tensor(op.name + '/b', [w.shape[0]], S.INT32,
                          (np.float64(src.scale) * q['w_scale'].astype(np.float64)).astype(np.float32),
                          np.zeros(w.shape[0], np.int64), data=q['bias'].astype(np.int32))
            if op.out_kind:                                   # head output: conv -> reshape -> (concat)
                h, wd = src.h, src.w
                cout = w.shape[0]
                scale, zp = (g.cls_scale, g.cls_zp) if op.out_kind == 1 else (g.box_scale, g.box_zp)
                raw = m.tensor(op.name, [1, h, wd, cout], S.INT8, scale, zp)
                last = 1 if op.out_kind == 1 else 4
                shp = [1, h * wd * cout // last, last]
                flat = m.tensor(op.name + '/flat', shp, S.INT8, scale, zp)
                shape_t = m.tensor(op.name + '/shape', [3], S.INT32, data=np.asarray(shp, np.int32))
                out_t = raw
            elif op.type == E.OP_PW and op.residual >= 0:
                tt = g.tensors[op.out]
                out_t = m.tensor(op.name + '/pre', [1, tt.h, tt.w, tt.c], S.INT8, q['pre_scale'], q['pre_zp'])
            else:
                out_t = tid[op.out] = act_tensor(op.out)
            stride = op.stride
            if op.type == E.OP_DW:
                m.op(S.DEPTHWISE_CONV_2D, [tid[op.inputs[0]], wt, bt], [out_t], S.OPT_DEPTHWISE,
                     T(f0=('i8', S.PAD_SAME), f1=('i32', stride), f2=('i32', stride), f3=('i32', 1),
                       f4=('i8', act) if act else None))
            else:
                conv_act = act if not (op.type == E.OP_PW and op.residual >= 0) else S.ACT_NONE
                m.op(S.CONV_2D, [tid[op.inputs[0]], wt, bt], [out_t], S.OPT_CONV2D,
                     T(f0=('i8', S.PAD_SAME), f1=('i32', stride), f2=('i32', stride),
                       f3=('i8', conv_act) if conv_act else None))
            if op.out_kind:
                m.op(S.RESHAPE, [raw, shape_t], [flat], S.OPT_RESHAPE, T(f0=V('i32', shp)))
                (cls_parts if op.out_kind == 1 else box_parts).append((op.level_offset, flat))
            elif op.type == E.OP_PW and op.residual >= 0:
                tid[op.out] = act_tensor(op.out)
                m.op(S.ADD, [out_t, tid[op.residual]], [tid[op.out]], S.OPT_ADD,
                     T(f0=('i8', act) if act else None))
        elif op.type == E.OP_MAXPOOL:
            tid[op.out] = act_tensor(op.out)
            m.op(S.MAX_POOL_2D, [tid[op.inputs[0]]], [tid[op.out]], S.OPT_POOL2D,
                 T(f0=('i8', S.PAD_SAME), f1=('i32', 2), f2=('i32', 2), f3=('i32', 3), f4=('i32', 3)))
        elif op.type == E.OP_ADD:
            assert len(op.inputs) == 2, 'TFLite ADD is binary (3-input fusions are trees of ADDs)'
            tout = g.tensors[op.out]
            ins = []
            for i, (x, rs) in enumerate(zip(op.inputs, op.resample)):
                tx = g.tensors[x]
                if rs == E.RS_NONE:
                    ins.append(tid[x])
                    continue
                tmp = m.tensor(f'{op.name}/in{i}', [1, tout.h, tout.w, tx.c], S.INT8, tx.scale, tx.zp)
                if rs == E.RS_UP:
                    size = m.tensor(f'{op.name}/size{i}', [2], S.INT32, data=np.asarray([tout.h, tout.w], np.int32))
                    m.op(S.RESIZE_NEAREST_NEIGHBOR, [tid[x], size], [tmp])
                else:
                    m.op(S.MAX_POOL_2D, [tid[x]], [tmp], S.OPT_POOL2D,
                         T(f0=('i8', S.PAD_SAME), f1=('i32', 2), f2=('i32', 2), f3=('i32', 3), f4=('i32', 3)))
                ins.append(tmp)
            tid[op.out] = act_tensor(op.out)
            m.op(S.ADD, ins, [tid[op.out]], S.OPT_ADD, T(f0=('i8', act) if act else None))
        else:
            raise ValueError(f'op type {op.type} has no TFLite counterpart')
    n = g.n_anchors
    cls_parts.sort(); box_parts.sort()
    cls_cat = m.tensor('class_net/concat', [1, n, 1], S.INT8, g.cls_scale, g.cls_zp)
    m.op(S.CONCATENATION, [t for _, t in cls_parts], [cls_cat], S.OPT_CONCAT, T(f0=('i32', 1)))
    cls_sig = m.tensor('class_net/logistic', [1, n, 1], S.INT8, 1.0 / 256.0, -128)
    m.op(S.LOGISTIC, [cls_cat], [cls_sig])
    cls_f = m.tensor('class_predictions', [1, n, 1], S.FLOAT32)
    m.op(S.DEQUANTIZE, [cls_sig], [cls_f])
    box_cat = m.tensor('box_net/concat', [1, n, 4], S.INT8, g.box_scale, g.box_zp)
    m.op(S.CONCATENATION, [t for _, t in box_parts], [box_cat], S.OPT_CONCAT, T(f0=('i32', 1)))
    box_f = m.tensor('box_encodings', [1, n, 4], S.FLOAT32)
    m.op(S.DEQUANTIZE, [box_cat], [box_f])
    anchors = m.tensor('anchors', [n, 4], S.FLOAT32, data=g.anchors().astype(np.float32))
    outs = [m.tensor('StatefulPartitionedCall:3', [1, 25, 4], S.FLOAT32),      # boxes   -> output_3
            m.tensor('StatefulPartitionedCall:2', [1, 25], S.FLOAT32),         # classes -> output_2
            m.tensor('StatefulPartitionedCall:1', [1, 25], S.FLOAT32),         # scores  -> output_1
            m.tensor('StatefulPartitionedCall:0', [1], S.FLOAT32)]             # count   -> output_0
    opts = flex_build_map(dict(max_detections=25, max_classes_per_detection=1, detections_per_class=100,
                               use_regular_nms=False, nms_score_threshold=-3.4028234663852886e38,
                               nms_iou_threshold=0.5, num_classes=E.NUM_CLASSES, y_scale=1.0, x_scale=1.0,
                               h_scale=1.0, w_scale=1.0))
    m.op(S.CUSTOM, [box_f, cls_f, anchors], outs, custom=S.DETECTION_POSTPROCESS, custom_options=opts)
    sub = T(f0=m.tensors, f1=V('i32', [images]), f2=V('i32', outs), f3=m.ops, f4='main')
    model = T(f0=('u32', 3), f1=[c[1] for c in m.codes], f2=[sub], f3=description, f4=m.buffers)
    return build(model, b'TFL3')


def save(g, path):
    with open(path, 'wb') as f:
        f.write(graph_to_tflite(g))
