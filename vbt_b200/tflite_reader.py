"""`.tflite` EfficientDet-Lite detection model -> layer program (`effdet.Graph`).

replaces: tflite_runtime.Interpreter(model_path=...) + allocate_tensors() (track.py:93-94,
eval.py:167-168), i.e. the model loading half of the interpreter.  SURVEY.md section 8(f) rank 1.

The file is walked operator by operator and mapped onto the ops libvbt_b200.so executes:

  QUANTIZE (uint8 -> int8 at the input)         folded: the stem reads uint8 with zero point zp+128
  CONV_2D 3x3 stride 2 on the input             OP_STEM
  CONV_2D 1x1                                   OP_PW   (+ a following ADD with the block input
                                                         becomes its fused residual epilogue)
  DEPTHWISE_CONV_2D                             OP_DW
  MAX_POOL_2D 3x3 s2 / RESIZE_NEAREST_NEIGHBOR  folded into the consuming ADD as a resampled input
                                                (a max-pool with other consumers stays OP_MAXPOOL)
  ADD                                           OP_ADD (binary)
  RESHAPE + CONCATENATION (+ LOGISTIC) + DEQUANTIZE + TFLite_Detection_PostProcess
                                                head outputs written level-major into the raw class /
                                                box tensors; LOGISTIC as a 256-entry LUT; anchors and
                                                NMS options taken from the custom op

PARITY UNPINNED [3P-MEM]: no `.tflite` written by TensorFlow exists in this container (the
reference's blobs are listed in .MISSING_LARGE_BLOBS), so the reader is only proven against files
written by tflite_writer.py.  The integer semantics follow the XNNPACK delegate tflite_runtime
2.14 applies by default (fp32 requantisation of the convolutions, single multiply-shift qs8 ADD);
TFLite's built-in reference / optimised kernels use gemmlowp fixed-point multipliers and a
two-stage ADD instead and may differ by one quantisation step.  The north star's tolerance for
real weights (IoU >= 0.99, |dscore| <= 1e-2) is the bar for files from TensorFlow, not bit equality.
"""
from __future__ import annotations

import numpy as np

from . import effdet as E
from . import tflite_schema as S
from .flatbuf import flex_map, root


class TfliteError(ValueError):
    pass


class _Tensor:
    def __init__(self, t, buffers):
        self.shape = [int(v) for v in t.vector(0, 'i32')]
        self.type = t.scalar(1, 'i8', 0)
        self.buffer = t.scalar(2, 'u32', 0)
        self.name = t.string(3)
        q = t.table(4)
        self.scale = q.vector(2, 'f32').copy() if q is not None else np.zeros(0, np.float32)
        self.zp = q.vector(3, 'i64').copy() if q is not None else np.zeros(0, np.int64)
        self.qdim = q.scalar(6, 'i32', 0) if q is not None else 0
        self._buffers = buffers

    def data(self):
        b = self._buffers[self.buffer]
        raw = b.vector(0, 'u8') if b is not None else np.zeros(0, np.uint8)
        if raw.size == 0:
            return None
        return raw.view(np.dtype({'f32': np.float32, 'i32': np.int32, 'u8': np.uint8, 'i64': np.int64,
                                  'i8': np.int8}[S.NP_OF_TYPE[self.type]]).newbyteorder('<')).reshape(self.shape)

    def s(self):
        if self.scale.size != 1:
            raise TfliteError(f'tensor {self.name!r}: expected per-tensor quantisation')
        return float(self.scale[0]), int(self.zp[0])


def parse(buf):
    """-> (tensors[_Tensor], operators[(builtin code, custom name, inputs, outputs, options table,
    custom options bytes)], subgraph inputs, subgraph outputs)."""
    m = root(buf, b'TFL3')
    codes = []
    for c in m.tables(1):
        builtin = max(c.scalar(3, 'i32', 0), c.scalar(0, 'i8', 0))
        codes.append((builtin, c.string(1)))
    subs = m.tables(2)
    if len(subs) != 1:
        raise TfliteError(f'expected one subgraph, found {len(subs)}')
    buffers = m.tables(4)
    sg = subs[0]
    tensors = [_Tensor(t, buffers) for t in sg.tables(0)]
    ops = []
    for o in sg.tables(3):
        builtin, custom = codes[o.scalar(0, 'u32', 0)]
        ops.append((builtin, custom, [int(v) for v in o.vector(1, 'i32')], [int(v) for v in o.vector(2, 'i32')],
                    o.table(4), bytes(o.vector(5, 'u8'))))
    return tensors, ops, [int(v) for v in sg.vector(1, 'i32')], [int(v) for v in sg.vector(2, 'i32')]


def _act_range(act, scale, zp):
    if act == S.ACT_NONE:
        return -128, 127, False
    if act != S.ACT_RELU6:
        raise TfliteError(f'fused activation {act} is not used by EfficientDet-Lite')
    return int(max(-128, zp)), int(min(127, zp + round(6.0 / scale))), True


def tflite_to_graph(path_or_bytes):
    buf = path_or_bytes
    if isinstance(buf, str):
        with open(buf, 'rb') as f:
            buf = f.read()
    tensors, ops, sg_in, _ = parse(buf)
    consumers = {}
    for oi, (_, _, ins, _, _, _) in enumerate(ops):
        for t in ins:
            consumers.setdefault(t, []).append(oi)
    g = E.Graph.__new__(E.Graph)
    g.tensors, g.ops = [], []
    g.quantized = True
    g.head_dtype = 'int8'
    tin = tensors[sg_in[0]]
    if tin.type != S.UINT8 or len(tin.shape) != 4 or tin.shape[3] != 3:
        raise TfliteError('expected a uint8 [1,S,S,3] image input')
    g.S = tin.shape[1]
    g.variant = {320: 'lite0', 384: 'lite1', 448: 'lite2'}.get(g.S, 'tflite')
    g.input = g._t(tin.shape[1], tin.shape[2], 3, 'input')
    g.tensors[g.input].scale, g.tensors[g.input].zp = tin.s()
    gid = {sg_in[0]: g.input}             # file tensor -> graph tensor
    via = {}                               # file tensor produced by a folded resample -> (source, mode)
    pre = {}                               # conv output awaiting its residual ADD -> op
    heads = {}                             # flattened head output (file tensor) -> (op, kind)
    concat = {}                            # concatenated tensor -> [parts]
    chain = {}                             # alias: tensor after LOGISTIC / DEQUANTIZE -> source, flags
    post = None

    def new_tensor(ft, relu6=False):
        t = tensors[ft]
        scale, zp = t.s()
        i = g._t(t.shape[1], t.shape[2], t.shape[3], t.name)
        g.tensors[i].scale, g.tensors[i].zp, g.tensors[i].relu6 = scale, zp, relu6
        gid[ft] = i
        return i

    for oi, (code, custom, ins, outs, opt, copt) in enumerate(ops):
        if code == S.QUANTIZE:
            if ins[0] != sg_in[0]:
                raise TfliteError('QUANTIZE is only supported on the model input')
            s8, z8 = tensors[outs[0]].s()
            if tensors[outs[0]].type != S.INT8 or z8 + 128 != g.tensors[g.input].zp:
                raise TfliteError('input QUANTIZE must be the uint8 -> int8 re-centring')
            gid[outs[0]] = g.input
        elif code in (S.CONV_2D, S.DEPTHWISE_CONV_2D):
            x, wt, bt = ins[:3]
            w, b = tensors[wt].data(), tensors[bt].data()
            sw = tensors[wt].scale.astype(np.float32)
            stride = opt.scalar(1, 'i32', 1) if opt is not None else 1
            depthwise = code == S.DEPTHWISE_CONV_2D
            act = opt.scalar(4 if depthwise else 3, 'i8', 0) if opt is not None else 0
            if (opt.scalar(0, 'i8', 0) if opt is not None else 0) != S.PAD_SAME:
                raise TfliteError('only SAME padding is supported')
            src = gid[x]
            tsrc = g.tensors[src]
            tout = tensors[outs[0]]
            so, zo = tout.s()
            lo, hi, relu6 = _act_range(act, so, zo)
            if depthwise:
                k = w.shape[1]
                op = E.Op(E.OP_DW, [src], -1, k=k, stride=stride, act=relu6, name=tout.name)
                wq = np.ascontiguousarray(np.transpose(w[0], (2, 0, 1)))
            elif src == g.input:
                if w.shape[1:3] != (3, 3) or stride != 2:
                    raise TfliteError('the first convolution must be 3x3 stride 2')
                op = E.Op(E.OP_STEM, [src], -1, k=3, stride=2, act=relu6, name=tout.name)
                wq = np.ascontiguousarray(w)
            else:
                if w.shape[1:3] != (1, 1) or stride != 1:
                    raise TfliteError(f'{tout.name}: only 1x1 stride-1 convolutions after the stem')
                op = E.Op(E.OP_PW, [src], -1, act=relu6, name=tout.name)
                wq = np.ascontiguousarray(w[:, 0, 0, :])
            if sw.size == 1:
                sw = np.full(wq.shape[0], sw[0], np.float32)
            q = op.q
            q['w'], q['w_scale'] = wq.astype(np.int8), sw
            q['bias'] = b.astype(np.int32)
            q['zp_in'] = [tsrc.zp]
            q['mult'] = (np.float64(tsrc.scale) * sw / np.float64(so)).astype(np.float32)
            q['zp_out'] = q['conv_zp_out'] = zo
            q['act_lo'], q['act_hi'] = lo, hi
            q['pre_scale'] = so
            g.ops.append(op)
            cons = consumers.get(outs[0], [])
            if len(cons) == 1 and ops[cons[0]][0] == S.RESHAPE:
                cout = wq.shape[0]
                op.out, op.out_kind = -1, 0                   # kind is known once the concat is seen
                heads[ops[cons[0]][3][0]] = op
                op._head_hw = (tsrc.h, tsrc.w, cout)
            elif (op.type == E.OP_PW and len(cons) == 1 and ops[cons[0]][0] == S.ADD and act == S.ACT_NONE
                  and _is_residual(ops[cons[0]], outs[0], gid, g, tout)):
                pre[outs[0]] = op                              # finished by the ADD below
            else:
                op.out = new_tensor(outs[0], relu6)
        elif code == S.MAX_POOL_2D:
            if opt is None or opt.scalar(3, 'i32', 0) != 3 or opt.scalar(1, 'i32', 0) != 2:
                raise TfliteError('only 3x3 stride-2 max-pooling is supported')
            cons = consumers.get(outs[0], [])
            if len(cons) == 1 and ops[cons[0]][0] == S.ADD:
                via[outs[0]] = (gid[ins[0]], E.RS_DOWN)
            else:
                o = new_tensor(outs[0])
                g.ops.append(E.Op(E.OP_MAXPOOL, [gid[ins[0]]], o, k=3, stride=2, name=tensors[outs[0]].name))
                g.ops[-1].q.update(zp_in=[g.tensors[gid[ins[0]]].zp], zp_out=g.tensors[o].zp, act_lo=-128, act_hi=127)
        elif code == S.RESIZE_NEAREST_NEIGHBOR:
            if any(ops[c][0] != S.ADD for c in consumers.get(outs[0], [])):
                raise TfliteError('RESIZE_NEAREST_NEIGHBOR must feed ADD ops')
            via[outs[0]] = (gid[ins[0]], E.RS_UP)
        elif code == S.ADD:
            act = opt.scalar(0, 'i8', 0) if opt is not None else 0
            tout = tensors[outs[0]]
            so, zo = tout.s()
            lo, hi, relu6 = _act_range(act, so, zo)
            pre_in = [t for t in ins if t in pre]
            if pre_in:                                         # residual epilogue of a project conv
                op = pre.pop(pre_in[0])
                other = [t for t in ins if t != pre_in[0]][0]
                op.residual = gid[other]
                op.out = new_tensor(outs[0], relu6)
                q = op.q
                q['pre_zp'] = q['conv_zp_out']
                q['res_zp'] = g.tensors[op.residual].zp
                q['add_mult'], q['add_shift'] = E._add_params([q['pre_scale'], g.tensors[op.residual].scale], so)
                q['zp_out'] = zo
                q['act_lo'], q['act_hi'] = lo, hi
                op.act = relu6
                # program order: the conv must come after everything the residual needs -- it does,
                # the ADD follows the conv in the file
            else:
                srcs, modes = [], []
                for t in ins:
                    s_, m_ = via.get(t, (gid.get(t), E.RS_NONE))
                    srcs.append(s_); modes.append(m_)
                o = new_tensor(outs[0], relu6)
                op = E.Op(E.OP_ADD, srcs, o, act=relu6, resample=modes, name=tout.name)
                q = op.q
                q['zp_in'] = [g.tensors[s_].zp for s_ in srcs]
                q['zp_out'], q['act_lo'], q['act_hi'] = zo, lo, hi
                q['add_mult'], q['add_shift'] = E._add_params([g.tensors[s_].scale for s_ in srcs], so)
                g.ops.append(op)
        elif code == S.RESHAPE:
            pass                                               # resolved through `heads`
        elif code == S.CONCATENATION:
            concat[outs[0]] = list(ins)
        elif code in (S.LOGISTIC, S.DEQUANTIZE):
            chain[outs[0]] = (ins[0], code)
        elif code == S.CUSTOM and custom == S.DETECTION_POSTPROCESS:
            post = (ins, flex_map(copt))
        else:
            raise TfliteError(f'operator {S.OP_NAMES.get(code, code)} ({custom!r}) is not part of the '
                              'EfficientDet-Lite detection graphs this loader supports')
    if post is None:
        raise TfliteError('no TFLite_Detection_PostProcess op: not a detection model')
    if pre:
        raise TfliteError('a convolution was left waiting for its residual ADD')
    (box_in, cls_in, anchors_t), opts = post

    def resolve(t):
        logistic = False
        while t in chain:
            t, code = chain[t]
            logistic |= code == S.LOGISTIC
        return t, logistic

    box_cat, _ = resolve(box_in)
    cls_cat, has_logistic = resolve(cls_in)
    if not has_logistic:
        raise TfliteError('class predictions must pass through an int8 LOGISTIC')
    g.box_scale, g.box_zp = tensors[box_cat].s()
    g.cls_scale, g.cls_zp = tensors[cls_cat].s()
    a_per = E.NUM_SCALES * len(E.ASPECTS)
    g.level_sizes = []
    for kind, cat in ((1, cls_cat), (2, box_cat)):
        offs = 0
        for li, part in enumerate(concat[cat]):
            op = heads[part]
            h, w, cout = op._head_hw
            if cout != a_per * (E.NUM_CLASSES if kind == 1 else 4):
                raise TfliteError(f'head output with {cout} channels: expected {a_per} anchors, {E.NUM_CLASSES} class')
            if tensors[part].s() != tensors[cat].s():
                raise TfliteError('head outputs must share the quantisation of their concatenation')
            op.out, op.out_kind, op.level_offset = -kind, kind, offs
            op.q['zp_out'] = op.q['conv_zp_out'] = tensors[cat].s()[1]
            if kind == 1:
                g.level_sizes.append((h, w))
                qs = np.arange(-128, 128)
                real = (qs - g.cls_zp) * g.cls_scale
                op.q['lut'] = np.clip(np.rint(256.0 / (1.0 + np.exp(-real))) - 128, -128, 127).astype(np.int8)
            offs += h * w * a_per
        g.n_anchors = offs
    anchors = tensors[anchors_t].data()
    if anchors is None or anchors.shape != (g.n_anchors, 4):
        raise TfliteError('anchor tensor does not match the head outputs')
    g.anchor_table = np.asarray(anchors, np.float32)
    g.postprocess_options = opts
    for key, want in (('max_detections', 25), ('num_classes', 1)):
        if int(opts.get(key, want)) != want:
            raise TfliteError(f'{key}={opts[key]} is not what libvbt_b200.so is built for ({want})')
    for key in ('y_scale', 'x_scale', 'h_scale', 'w_scale'):
        if float(opts.get(key, 1.0)) != 1.0:
            raise TfliteError(f'{key}={opts[key]}: only unit box-coder scales are supported')
    _assign_branches(g)
    return g


def _is_residual(add_op, conv_out, gid, g, tout):
    """ADD(conv, block input): the other operand is an existing activation of the same shape."""
    other = [t for t in add_op[2] if t != conv_out]
    if len(other) != 1 or other[0] not in gid:
        return False
    t = g.tensors[gid[other[0]]]
    return [t.h, t.w, t.c] == tout.shape[1:4]


def _assign_branches(g):
    """Head chains (ops that lead to exactly one head output through single-consumer tensors) may
    run concurrently: give each its own branch id, like effdet.Graph._build."""
    users = {}
    for i, op in enumerate(g.ops):
        for t in op.inputs + ([op.residual] if op.residual >= 0 else []):
            users.setdefault(t, []).append(i)
    producer = {op.out: i for i, op in enumerate(g.ops) if op.out >= 0}
    b = 0
    for i, op in enumerate(g.ops):
        if op.out_kind == 0:
            continue
        b += 1
        j = i
        while True:
            g.ops[j].branch = b
            src = g.ops[j].inputs[0]
            if len(g.ops[j].inputs) != 1 or len(users.get(src, [])) != 1 or src not in producer:
                break
            j = producer[src]
    if b > 16:
        for op in g.ops:
            op.branch = 0
    # the library expects a branch's ops to be contiguous in program order
    order = sorted(range(len(g.ops)), key=lambda i: (g.ops[i].branch > 0, g.ops[i].branch, i))
    g.ops = [g.ops[i] for i in order]
