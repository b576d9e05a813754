"""EfficientDet-Lite0/1/2 layer program for libvbt_b200.so.

The reference runs the detector as an opaque `.tflite` flatbuffer inside
tflite_runtime (track.py:93-94, odt.py:58-61); the architecture is only named by
train.py:23,29 (`model_spec.get('efficientdet_lite0')`, tflite-model-maker 0.4.3).  This
module restates that architecture as a flat list of int8 ops --
the "layer program" -- that the CUDA library executes, and packs it into the binary
blob `vbt_model_create` parses (layout: vbt_b200/csrc/model.cuh).

Since the reference checkout ships no weights (.MISSING_LARGE_BLOBS), models are built
from seeded synthetic weights and post-training-quantised here (`build_synthetic`):
per-output-channel symmetric int8 weights, per-tensor asymmetric int8 activations,
int32 bias, fp32 requantisation multiplier -- the scheme of the exported graphs.

Integer op semantics (identical in oracle/effdet.py and the kernels):
  conv / depthwise: acc = sum((x - zp_in) * w) + bias            (int32)
                    y   = clamp(rne(float32(acc) * M[c]) + zp_out, act_lo, act_hi)
  add (n-ary)     : y   = clamp(((sum_i (x_i - zp_i) * mult_i + 2^(shift-1)) >> shift)
                          + zp_out, act_lo, act_hi)
  max-pool / nearest resize: on raw int8 (same scale both sides)
  logistic        : 256-entry int8 LUT, output scale 1/256, zero point -128
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

OP_STEM, OP_PW, OP_DW, OP_ADD, OP_MAXPOOL, OP_LOGISTIC = 1, 2, 3, 4, 5, 6
RS_NONE, RS_UP, RS_DOWN = 0, 1, 2
BLOB_MAGIC, BLOB_VERSION = 0x4d544256, 11

VARIANTS = {
    # name: (input size, width, depth, fpn channels, fpn cells, head repeats)
    'lite0': (320, 1.0, 1.0, 64, 3, 3),
    'lite1': (384, 1.0, 1.1, 88, 4, 3),
    'lite2': (448, 1.1, 1.2, 112, 5, 3),
}
# EfficientNet-B0 stages: kernel, repeats, out channels, expand ratio, stride
_STAGES = [(3, 1, 16, 1, 1), (3, 2, 24, 6, 2), (5, 2, 40, 6, 2), (3, 3, 80, 6, 2),
           (5, 3, 112, 6, 1), (5, 4, 192, 6, 2), (3, 1, 320, 6, 1)]
NUM_SCALES, ASPECTS, ANCHOR_SCALE = 3, (1.0, 2.0, 0.5), 3.0
NUM_CLASSES = 1
FUSE_MAX_HW = 64       # largest map side the fused node kernel takes (csrc/model.cu)


def pad16(c):
    return (c + 15) // 16 * 16


def round_filters(f, width):
    f = f * width
    new = max(8, int(f + 4) // 8 * 8)
    if new < 0.9 * f:
        new += 8
    return int(new)


def same_pad(size, k, stride):
    """TF SAME: (output size, pad before)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


@dataclass
class Tensor:
    h: int
    w: int
    c: int
    name: str = ''
    scale: float = 1.0
    zp: int = 0
    relu6: bool = False          # produced by a ReLU6-clamped op
    ws_offset: int = -1

    @property
    def c_p(self):
        return pad16(self.c)

    @property
    def bytes_per_frame(self):
        return self.h * self.w * self.c_p


@dataclass
class Op:
    type: int
    inputs: list
    out: int
    k: int = 1
    stride: int = 1
    act: bool = False                        # ReLU6 after this op
    resample: list = field(default_factory=list)
    weight: np.ndarray = None                # float: PW [co,ci]; DW [c,k,k]; STEM [co,k,k,ci]
    bias: np.ndarray = None
    residual: int = -1                       # PW: tensor added after the conv
    out_kind: int = 0                        # 0 workspace, 1 class output, 2 box output
    level_offset: int = 0                    # anchors before this level (head outputs)
    branch: int = 0                          # 0: trunk; k > 0: independent head chain k
    name: str = ''
    # filled by quantize()
    q: dict = field(default_factory=dict)


class Graph:
    def __init__(self, variant):
        self.variant = variant
        self.S, self.width, self.depth, self.C, self.cells, self.head_rep = VARIANTS[variant]
        self.tensors: list[Tensor] = []
        self.ops: list[Op] = []
        self.level_sizes = []
        self.n_anchors = 0
        self._build()

    # -- construction helpers ---------------------------------------------------------
    def _t(self, h, w, c, name):
        self.tensors.append(Tensor(h, w, c, name))
        return len(self.tensors) - 1

    def _pw(self, x, cout, act, name, residual=-1, out_kind=0, level_offset=0):
        t = self.tensors[x]
        o = self._t(t.h, t.w, cout, name) if out_kind == 0 else -out_kind
        self.ops.append(Op(OP_PW, [x], o, act=act, residual=residual, out_kind=out_kind,
                           level_offset=level_offset, name=name))
        return o

    def _dw(self, x, k, stride, act, name):
        t = self.tensors[x]
        ho, _ = same_pad(t.h, k, stride)
        wo, _ = same_pad(t.w, k, stride)
        o = self._t(ho, wo, t.c, name)
        self.ops.append(Op(OP_DW, [x], o, k=k, stride=stride, act=act, name=name))
        return o

    def _maxpool(self, x, name):
        t = self.tensors[x]
        ho, _ = same_pad(t.h, 3, 2)
        wo, _ = same_pad(t.w, 3, 2)
        o = self._t(ho, wo, t.c, name)
        self.ops.append(Op(OP_MAXPOOL, [x], o, k=3, stride=2, name=name))
        return o

    def _fuse(self, xs, level_hw, name, act=True):
        h, w = level_hw
        rs = []
        for x in xs:
            t = self.tensors[x]
            if (t.h, t.w) == (h, w):
                rs.append(RS_NONE)
            elif t.h < h:
                rs.append(RS_UP)
            else:
                assert same_pad(t.h, 3, 2)[0] == h, 'only one pooling step between levels'
                rs.append(RS_DOWN)
        o = self._t(h, w, self.tensors[xs[0]].c, name)
        self.ops.append(Op(OP_ADD, list(xs), o, act=act, resample=rs, name=name))
        return o

    # -- the network -------------------------------------------------------------------
    def _build(self):
        S, C = self.S, self.C
        self.input = self._t(S, S, 3, 'input')
        ho, _ = same_pad(S, 3, 2)
        x = self._t(ho, ho, 32, 'stem')
        self.ops.append(Op(OP_STEM, [self.input], x, k=3, stride=2, act=True, name='stem'))
        feats = {}
        n_stage = len(_STAGES)
        for si, (k, rep, cout, expand, stride) in enumerate(_STAGES):
            cout = round_filters(cout, self.width)
            if 0 < si < n_stage - 1:
                rep = int(math.ceil(self.depth * rep))
            for r in range(rep):
                s = stride if r == 0 else 1
                cin = self.tensors[x].c
                name = f'b{si + 1}.{r}'
                y = x
                if expand != 1:
                    y = self._pw(y, cin * expand, True, name + '.expand')
                y = self._dw(y, k, s, True, name + '.dw')
                skip = x if (s == 1 and cin == cout) else -1
                x = self._pw(y, cout, False, name + '.project', residual=skip)
            if si in (2, 4, 6):
                feats[3 + (si - 2) // 2] = x
        # BiFPN
        lvl_hw = {l: (self.tensors[feats[l]].h, self.tensors[feats[l]].w) for l in (3, 4, 5)}
        p6_in = self._pw(feats[5], C, False, 'p6.lateral')
        p6 = self._maxpool(p6_in, 'p6')
        p7 = self._maxpool(p6, 'p7')
        lvl_hw[6] = (self.tensors[p6].h, self.tensors[p6].w)
        lvl_hw[7] = (self.tensors[p7].h, self.tensors[p7].w)
        feats5 = [feats[3], feats[4], feats[5], p6, p7]          # P3..P7 entering a cell
        plan = [(6, [3, 4]), (5, [2, 5]), (4, [1, 6]), (3, [0, 7]),
                (4, [1, 7, 8]), (5, [2, 6, 9]), (6, [3, 5, 10]), (7, [4, 11])]
        for cell in range(self.cells):
            local = list(feats5)                     # indices 0..4, nodes append 5..12
            for ni, (level, ins) in enumerate(plan):
                xs = []
                for j in ins:
                    src = local[j]
                    if self.tensors[src].c != C:       # backbone feature: own lateral conv
                        src = self._pw(src, C, False, f'fpn{cell}.n{ni}.lat{j}')
                    xs.append(src)
                name = f'fpn{cell}.n{ni}'
                # tf.add_n of three inputs reaches the exported graph as a tree of binary ADDs
                # (the converter lowers AddN; int8 ADD_N does not exist), each with its own
                # output quantisation: ADD(ADD(a, b), c), the activation on the last one
                if len(xs) == 3:
                    xs = [self._fuse(xs[:2], lvl_hw[level], name + '.sum0', act=False), xs[2]]
                f = self._fuse(xs, lvl_hw[level], name + '.sum')
                d = self._dw(f, 3, 1, False, name + '.dw')
                local.append(self._pw(d, C, False, name + '.pw'))
            feats5 = local[8:13]                     # last output of levels 3,4,5,6,7
        self.fpn_out = feats5
        # heads
        a_per = NUM_SCALES * len(ASPECTS)
        self.level_sizes = [lvl_hw[l] for l in (3, 4, 5, 6, 7)]
        offs = 0
        for li, l in enumerate((3, 4, 5, 6, 7)):
            for ni, (net, cout, kind) in enumerate((('cls', a_per * NUM_CLASSES, 1), ('box', a_per * 4, 2))):
                first = len(self.ops)
                x = self.fpn_out[li]
                for r in range(self.head_rep):
                    x = self._dw(x, 3, 1, False, f'{net}{l}.{r}.dw')
                    x = self._pw(x, C, True, f'{net}{l}.{r}.pw')
                x = self._dw(x, 3, 1, False, f'{net}{l}.out.dw')
                self._pw(x, cout, False, f'{net}{l}.out.pw', out_kind=kind, level_offset=offs)
                # the ten (level, net) chains only read the BiFPN outputs and write disjoint
                # slices of the raw outputs: independent branches the GPU may run concurrently
                for op in self.ops[first:]:
                    op.branch = 1 + li * 2 + ni
            offs += lvl_hw[l][0] * lvl_hw[l][1] * a_per
        self.n_anchors = offs

    # -- anchors ---------------------------------------------------
    def anchors(self):
        """f32 [N,4] (ycentre, xcentre, h, w), normalised, level-major then y, x, then
        octave-major / aspect-minor.  A graph imported from a .tflite file carries the file's own
        anchor tensor (`anchor_table`)."""
        if getattr(self, 'anchor_table', None) is not None:
            return np.asarray(self.anchor_table, dtype=np.float32)
        out = []
        S = float(self.S)
        for (h, w) in self.level_sizes:
            sy, sx = S / h, S / w
            for y in range(h):
                for x in range(w):
                    for o in range(NUM_SCALES):
                        for a in ASPECTS:
                            base_y = ANCHOR_SCALE * sy * 2 ** (o / NUM_SCALES)
                            base_x = ANCHOR_SCALE * sx * 2 ** (o / NUM_SCALES)
                            ah = base_y / math.sqrt(a)
                            aw = base_x * math.sqrt(a)
                            out.append(((sy / 2 + y * sy) / S, (sx / 2 + x * sx) / S,
                                        ah / S, aw / S))
        return np.asarray(out, dtype=np.float32)

    def macs(self):
        tot = 0
        for op in self.ops:
            if op.type == OP_PW:
                t = self.tensors[op.inputs[0]]
                cout = self.tensors[op.out].c if op.out >= 0 else (9 if op.out_kind == 1 else 36)
                tot += t.h * t.w * t.c * cout
            elif op.type == OP_DW:
                t = self.tensors[op.out]
                tot += t.h * t.w * t.c * op.k * op.k
            elif op.type == OP_STEM:
                t = self.tensors[op.out]
                tot += t.h * t.w * t.c * 27
        return tot

    def out_channels(self, op):
        if op.out >= 0:
            return self.tensors[op.out].c
        return NUM_SCALES * len(ASPECTS) * (NUM_CLASSES if op.out_kind == 1 else 4)


# ---------------------------------------------------------------------------------------
# synthetic weights, float forward (calibration only), quantisation
# ---------------------------------------------------------------------------------------

def init_weights(g: Graph, seed=1234):
    """Seeded He-normal folded-BN weights."""
    rng = np.random.default_rng(seed)
    for op in g.ops:
        if op.type == OP_STEM:
            cout = g.tensors[op.out].c
            op.weight = rng.normal(0, math.sqrt(2.0 / 27), (cout, 3, 3, 3)).astype(np.float32)
            op.bias = rng.normal(0, 0.1, cout).astype(np.float32)
        elif op.type == OP_PW:
            cin = g.tensors[op.inputs[0]].c
            cout = g.out_channels(op)
            gain = 2.0 if op.act else 1.0
            op.weight = rng.normal(0, math.sqrt(gain / cin), (cout, cin)).astype(np.float32)
            if op.out_kind == 1:      # class prior: sigmoid(bias) ~ 1 % like the real head
                op.bias = np.full(cout, -math.log(99.0), np.float32) + \
                    rng.normal(0, 0.5, cout).astype(np.float32)
                op.weight *= 1.1      # a handful of anchors per frame clear score 0.5
            else:
                op.bias = rng.normal(0, 0.1, cout).astype(np.float32)
                if op.out_kind == 2:
                    op.weight *= 0.25  # keeps exp(th), exp(tw) near 1 like a trained head
        elif op.type == OP_DW:
            c = g.tensors[op.out].c
            gain = 2.0 if op.act else 1.0
            op.weight = rng.normal(0, math.sqrt(gain / (op.k * op.k)), (c, op.k, op.k)).astype(np.float32)
            op.bias = rng.normal(0, 0.1, c).astype(np.float32)


def nearest_index(dst, n_in, n_out):
    return (np.arange(dst) * n_in) // n_out


def float_forward(g: Graph, frames_u8, record=None):
    """fp32 forward of the un-quantised graph on uint8 [B,S,S,3] frames (torch CPU).
    Used ONLY to calibrate activation ranges when a synthetic model is built."""
    import torch
    import torch.nn.functional as F
    vals = {}
    x = torch.from_numpy(np.asarray(frames_u8)).float().permute(0, 3, 1, 2)
    vals[g.input] = (x - 127.0) / 128.0
    outs = {1: [], 2: []}

    def conv(x, w, b, k, stride, groups):
        _, pt = same_pad(x.shape[2], k, stride)
        _, pl = same_pad(x.shape[3], k, stride)
        ho, wo = same_pad(x.shape[2], k, stride)[0], same_pad(x.shape[3], k, stride)[0]
        pb = max((ho - 1) * stride + k - x.shape[2] - pt, 0)
        pr = max((wo - 1) * stride + k - x.shape[3] - pl, 0)
        x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, w, b, stride=stride, groups=groups)

    for op in g.ops:
        ins = [vals[i] for i in op.inputs]
        if op.type == OP_STEM:
            w = torch.from_numpy(op.weight).permute(0, 3, 1, 2).contiguous()
            y = conv(ins[0], w, torch.from_numpy(op.bias), 3, 2, 1)
        elif op.type == OP_PW:
            w = torch.from_numpy(op.weight)[:, :, None, None]
            y = F.conv2d(ins[0], w, torch.from_numpy(op.bias))
        elif op.type == OP_DW:
            w = torch.from_numpy(op.weight)[:, None]
            y = conv(ins[0], w, torch.from_numpy(op.bias), op.k, op.stride, w.shape[0])
        elif op.type == OP_MAXPOOL:
            xin = ins[0]
            ho, pt = same_pad(xin.shape[2], 3, 2)
            wo, pl = same_pad(xin.shape[3], 3, 2)
            pb = max((ho - 1) * 2 + 3 - xin.shape[2] - pt, 0)
            pr = max((wo - 1) * 2 + 3 - xin.shape[3] - pl, 0)
            y = F.max_pool2d(F.pad(xin, (pl, pr, pt, pb), value=-1e30), 3, 2)
        elif op.type == OP_ADD:
            t = g.tensors[op.out]
            y = 0
            for xin, rs in zip(ins, op.resample):
                if rs == RS_UP:
                    iy = torch.from_numpy(nearest_index(t.h, xin.shape[2], t.h))
                    ix = torch.from_numpy(nearest_index(t.w, xin.shape[3], t.w))
                    xin = xin[:, :, iy][:, :, :, ix]
                elif rs == RS_DOWN:
                    ho, pt = same_pad(xin.shape[2], 3, 2)
                    wo, pl = same_pad(xin.shape[3], 3, 2)
                    pb = max((ho - 1) * 2 + 3 - xin.shape[2] - pt, 0)
                    pr = max((wo - 1) * 2 + 3 - xin.shape[3] - pl, 0)
                    xin = F.max_pool2d(F.pad(xin, (pl, pr, pt, pb), value=-1e30), 3, 2)
                y = y + xin
        if op.act:
            y = y.clamp(0.0, 6.0)
        if op.type == OP_PW and op.residual >= 0:
            if record is not None:
                record.setdefault(('pre', id(op)), []).append((float(y.min()), float(y.max())))
            y = y + vals[op.residual]
        if op.out >= 0:
            vals[op.out] = y
        else:
            outs[op.out_kind].append(y)
        if record is not None:
            record.setdefault(id(op), []).append((float(y.min()), float(y.max())))
    return vals, outs


def _coarse(x):
    """Calibrated range end rounded to four significant digits: the float forward pass sums in a
    thread-count dependent order, and a model built under torchrun (OMP_NUM_THREADS=1) must be the
    model built anywhere else -- ranks compare results byte for byte."""
    return float(f'{float(x):.4g}')


def _qparams(lo, hi):
    lo, hi = min(_coarse(lo), 0.0), max(_coarse(hi), 0.0)
    if hi - lo < 1e-6:
        hi = lo + 1e-6
    scale = (hi - lo) / 255.0
    zp = int(round(-128 - lo / scale))
    return float(np.float32(scale)), int(min(127, max(-128, zp)))


def _add_params(scales, out_scale):
    ratios = [s / out_scale for s in scales]
    mx = max(ratios)
    shift = 20 - int(math.floor(math.log2(mx))) if mx > 0 else 20
    shift = max(1, min(30, shift))
    mults = [int(round(r * (1 << shift))) for r in ratios]
    return mults, shift


def quantize(g: Graph, calib_frames):
    """Post-training quantisation of a graph that already has float weights."""
    rec = {}
    float_forward(g, calib_frames, rec)
    rng_of = {k: (min(a for a, _ in v), max(b for _, b in v)) for k, v in rec.items()}
    tin = g.tensors[g.input]
    tin.scale, tin.zp = 1.0 / 128.0, 127          # uint8 input, mean 127 / std 128
    box_lo, box_hi = 0.0, 0.0
    for op in g.ops:
        if op.out_kind == 2:
            lo, hi = rng_of[id(op)]
            box_lo, box_hi = min(box_lo, lo), max(box_hi, hi)
    g.box_scale, g.box_zp = _qparams(box_lo, box_hi)
    cls_lo = min(rng_of[id(op)][0] for op in g.ops if op.out_kind == 1)
    cls_hi = max(rng_of[id(op)][1] for op in g.ops if op.out_kind == 1)
    g.cls_scale, g.cls_zp = _qparams(cls_lo, cls_hi)
    for op in g.ops:
        lo, hi = rng_of[id(op)]
        if op.out >= 0:
            t = g.tensors[op.out]
            if op.act and not (op.type == OP_PW and op.residual >= 0):
                lo, hi = 0.0, min(hi, 6.0)
            t.scale, t.zp = _qparams(lo, hi)
            t.relu6 = op.act
            so, zo = t.scale, t.zp
        elif op.out_kind == 1:
            so, zo = g.cls_scale, g.cls_zp
        else:
            so, zo = g.box_scale, g.box_zp
        q = op.q
        q['zp_out'] = zo
        q['act_lo'], q['act_hi'] = -128, 127
        if op.act:
            q['act_lo'] = zo
            q['act_hi'] = int(min(127, zo + round(6.0 / so)))
        ins = [g.tensors[i] for i in op.inputs]
        q['zp_in'] = [t.zp for t in ins]
        if op.type in (OP_STEM, OP_PW, OP_DW):
            si = ins[0].scale
            conv_so, conv_zo = so, zo
            if op.type == OP_PW and op.residual >= 0:
                plo, phi = rng_of[('pre', id(op))]
                conv_so, conv_zo = _qparams(plo, phi)
                r = g.tensors[op.residual]
                q['pre_zp'] = conv_zo
                q['res_zp'] = r.zp
                q['add_mult'], q['add_shift'] = _add_params([conv_so, r.scale], so)
            w = op.weight.reshape(op.weight.shape[0], -1)
            sw = np.maximum(np.abs(w).max(axis=1), 1e-8) / 127.0
            q['w'] = np.clip(np.rint(op.weight / sw.reshape((-1,) + (1,) * (op.weight.ndim - 1))),
                             -127, 127).astype(np.int8)
            q['bias'] = np.rint(op.bias / (si * sw)).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)
            q['mult'] = (np.float64(si) * sw / np.float64(conv_so)).astype(np.float32)
            q['conv_zp_out'] = conv_zo
            q['w_scale'] = sw.astype(np.float32)          # per-channel filter scales (.tflite export)
            q['pre_scale'] = float(np.float32(conv_so))   # scale of the conv's own int8 result
            if op.out_kind == 1:       # LOGISTIC fused behind the class conv
                qs = np.arange(-128, 128)
                real = (qs - zo) * so
                q['lut'] = np.clip(np.rint(256.0 / (1.0 + np.exp(-real))) - 128, -128, 127).astype(np.int8)
        elif op.type == OP_ADD:
            q['add_mult'], q['add_shift'] = _add_params([t.scale for t in ins], so)
        elif op.type == OP_MAXPOOL:
            t = g.tensors[op.out]
            t.scale, t.zp = ins[0].scale, ins[0].zp
            q['zp_out'] = t.zp
    g.quantized = True
    return g


def mbconv_runs(g):
    """{index of the first op: (expand op index or -1, depthwise index, project index)} for every
    MBConv block the library runs as ONE kernel (csrc/mbconv_umma.cu): 1x1 expand (ReLU6) ->
    depthwise 3x3 / 5x5, stride 1 / 2 (ReLU6) -> 1x1 project [+ the block input as residual], the
    expanded tensor and the depthwise output living only in shared memory / TMEM.  The first block
    of the backbone has no expand conv (expand ratio 1): depthwise -> project on a map too large
    for the fused node kernel."""
    n = len(g.ops)
    readers = {}
    for op in g.ops:
        for t in op.inputs + ([op.residual] if op.residual >= 0 else []):
            readers[t] = readers.get(t, 0) + 1
    runs = {}

    def dw_project(i, x):
        if i + 1 >= n:
            return False
        d, p = g.ops[i], g.ops[i + 1]
        return (d.type == OP_DW and d.k in (3, 5) and d.stride in (1, 2) and d.branch == 0 and
                readers.get(d.out, 0) == 1 and p.type == OP_PW and p.inputs == [d.out] and p.out_kind == 0 and
                p.branch == 0 and p.residual in (-1, x) and (p.residual < 0 or d.stride == 1) and
                pad16(g.out_channels(p)) <= 352)

    i = 0
    while i < n:
        e = g.ops[i]
        if (e.type == OP_STEM and e.k == 3 and e.stride == 2 and i + 2 < n and g.tensors[e.out].c_p <= 32 and
                e.inputs == [g.input] and readers.get(e.out, 0) == 1 and g.ops[i + 1].inputs == [e.out] and
                dw_project(i + 1, -2) and g.ops[i + 2].residual < 0):
            # the stem as the block's "expand" stage: a 3x3 s2 conv is a K = 27 GEMM over im2col rows
            runs[i] = (i, i + 1, i + 2)
            i += 3
            continue
        if (e.type == OP_PW and e.out_kind == 0 and e.residual < 0 and e.branch == 0 and i + 2 < n and
                readers.get(e.out, 0) == 1 and g.ops[i + 1].inputs == [e.out] and dw_project(i + 1, e.inputs[0]) and
                g.tensors[e.inputs[0]].c_p <= 256):
            runs[i] = (i, i + 1, i + 2)
            i += 3
        elif (e.type == OP_DW and g.tensors[e.out].c_p <= 32 and
              max(g.tensors[e.inputs[0]].h, g.tensors[e.inputs[0]].w) > FUSE_MAX_HW and
              dw_project(i, -2) and g.ops[i + 1].residual < 0):
            runs[i] = (-1, i, i + 1)
            i += 2
        else:
            i += 1
    return runs


def mbconv_image_layout(cin_p, k, cout_p, has_expand):
    """Byte offsets inside one 32-expanded-channel chunk image (mirrored by csrc/mbconv_umma.cu):
    (taps offset, project-weight offset, constants offset, image stride)."""
    ge_in = (cin_p // 16 + 1) // 2 * 2
    sz_wexp = 32 * ge_in * 16 if has_expand else 0
    sz_taps = k * (2 if k == 5 else 1) * 32 * 4   # [k rows][1 or 2 words][32 channels] u32
    off_taps = sz_wexp
    off_wproj = off_taps + sz_taps
    off_consts = off_wproj + cout_p * 32
    stride = (off_consts + 512 + 127) // 128 * 128
    return off_taps, off_wproj, off_consts, stride


def mbconv_images(cin_p, k, cout_p, e, d, p):
    """Per-chunk weight images of one MBConv block, each the exact shared-memory picture the fused
    kernel wants (one bulk copy per chunk): the expand weights of 32 expanded channels as a K-major
    no-swizzle core-matrix B operand, the 32 channels' depthwise taps packed four horizontal taps to a word [k][1 or 2][32], the project
    weights' 32-wide K slice as a core-matrix B operand [cout_p][32], and the chunk's expand /
    depthwise bias + multiplier vectors.  e / d / p: dict(w=int8 padded weights, bias=int32 folded,
    mult=float32) of the three convs (e = None: no expand conv).
    Core-matrix layout of a [N][K] operand: byte (n, kb) at ((n // 8) * (K // 16) + kb // 16) * 128 +
    (n % 8) * 16 + kb % 16."""
    off_taps, off_wproj, off_consts, stride = mbconv_image_layout(cin_p, k, cout_p, e is not None)
    ge_in = (cin_p // 16 + 1) // 2 * 2
    cexp = d['w'].shape[0]
    n_chunks = (cexp + 31) // 32
    out = np.zeros((n_chunks, stride), np.uint8)
    nn = np.arange(32)
    for c in range(n_chunks):
        lo, hi = 32 * c, min(32 * c + 32, cexp)
        m = hi - lo
        if e is not None:
            w = np.zeros((32, ge_in * 16), np.int8)
            ew = e['w'].reshape(e['w'].shape[0], -1)      # stem: [cout][ky][kx][c] -> K = (ky * 3 + kx) * 3 + c
            w[:m, :ew.shape[1]] = ew[lo:hi]
            kb = np.arange(ge_in * 16)
            off = ((nn[:, None] // 8) * ge_in + kb[None, :] // 16) * 128 + (nn[:, None] % 8) * 16 + kb[None, :] % 16
            out[c, off.reshape(-1)] = w.view(np.uint8).reshape(-1)
        # depthwise taps for the planar dp4a form: per window row ky, word 0 = taps kx 0..3 of the channel
        # in its four bytes (kx 3 = 0 for 3x3), word 1 (5x5 only) = tap kx 4 in byte 0
        nww = 2 if k == 5 else 1
        tw = np.zeros((k, nww * 4, 32), np.int64)
        tw[:, :k, :m] = (d['w'][lo:hi].astype(np.int64) & 0xff).transpose(1, 2, 0)       # [ky][kx][ch]
        words = tw.reshape(k, nww, 4, 32)
        taps = (words[:, :, 0] | (words[:, :, 1] << 8) | (words[:, :, 2] << 16) | (words[:, :, 3] << 24)).astype(np.uint32)
        out[c, off_taps:off_taps + k * nww * 128] = taps.view(np.uint8).reshape(-1)
        wp = np.zeros((cout_p, 32), np.int8)
        wp[:p['w'].shape[0], :m] = p['w'][:, lo:hi]
        n2, k2 = np.arange(cout_p)[:, None], np.arange(32)[None, :]
        off = ((n2 // 8) * 2 + k2 // 16) * 128 + (n2 % 8) * 16 + k2 % 16
        out[c, off_wproj + off.reshape(-1)] = wp.view(np.uint8).reshape(-1)
        consts = np.zeros(128, np.int32)
        if e is not None:
            consts[:m] = e['bias'][lo:hi]
            consts[32:32 + m] = e['mult'][lo:hi].view(np.int32)
        # the expanded tensor is held in shared memory as UNSIGNED bytes (value + 128, dp4a.u32.s32): the
        # depthwise bias carries the - 128 * sum(w) that undoes it
        consts[64:64 + m] = d['bias'][lo:hi].astype(np.int64) - 128 * d['w'][lo:hi].astype(np.int64).reshape(m, -1).sum(axis=1)
        consts[96:96 + m] = d['mult'][lo:hi].view(np.int32)
        out[c, off_consts:off_consts + 512] = consts.view(np.uint8)
    return out, stride, n_chunks


def fused_run_end(g):
    """run_end[i] = index of the last op of the [[ADD ->] ADD ->] DW3x3 s1 -> PW run op i may be executed
    in as one kernel (vbt_model_create's launch plan, csrc/model.cu), i itself otherwise.  A
    superset of what the library fuses is fine: it only delays memory reuse."""
    n = len(g.ops)
    readers = {}
    for op in g.ops:
        for t in op.inputs + ([op.residual] if op.residual >= 0 else []):
            readers[t] = readers.get(t, 0) + 1
    end = list(range(n))

    def dw_pw(i):
        if i + 1 >= n:
            return False
        d, p = g.ops[i], g.ops[i + 1]
        t = g.tensors[d.out]
        return (d.type == OP_DW and d.k == 3 and d.stride == 1 and readers.get(d.out, 0) == 1 and
                t.c_p <= 128 and max(t.h, t.w) <= FUSE_MAX_HW and pad16(g.out_channels(p)) <= 128 and
                p.type == OP_PW and p.inputs == [d.out] and p.residual < 0 and p.branch == d.branch)

    def add_dw_pw(i):
        o = g.ops[i]
        return (o.type == OP_ADD and readers.get(o.out, 0) == 1 and dw_pw(i + 1) and
                g.ops[i + 1].inputs == [o.out] and g.ops[i + 1].branch == o.branch)

    for first, (_, _, last) in mbconv_runs(g).items():
        for j in range(first, last + 1):
            end[j] = last
    i = 0
    while i < n:
        o = g.ops[i]
        if end[i] != i:
            i = end[i] + 1
            continue
        if (o.type == OP_ADD and len(o.inputs) == 2 and readers.get(o.out, 0) == 1 and i + 1 < n and
                g.ops[i + 1].type == OP_ADD and len(g.ops[i + 1].inputs) == 2 and o.out in g.ops[i + 1].inputs and
                g.ops[i + 1].branch == o.branch and add_dw_pw(i + 1)):
            end[i] = end[i + 1] = end[i + 2] = end[i + 3] = i + 3
            i += 4
        elif add_dw_pw(i):
            end[i] = end[i + 1] = end[i + 2] = i + 2
            i += 3
        elif dw_pw(i):
            end[i] = end[i + 1] = i + 1
            i += 2
        else:
            i += 1
    return end


def plan_workspace(g: Graph):
    """First-fit allocation of per-frame activation offsets with liveness reuse.

    Ops of branch 0 run in program order and recycle memory as tensors die.  Ops of a branch
    k > 0 (the head chains) may run concurrently with every other branch: their tensors come
    from a private region that is only recycled within the branch, and a trunk tensor read by
    any branch is never recycled."""
    last_use = {}
    keep = set()
    for i, op in enumerate(g.ops):
        for t in op.inputs + ([op.residual] if op.residual >= 0 else []):
            last_use[t] = i
            if op.branch != 0:
                keep.add(t)
    pools = {}                            # branch -> (free list, sizes)
    top = 0

    def alloc(free, n):
        nonlocal top
        n = (n + 255) // 256 * 256
        for j, (o, s) in enumerate(free):
            if s >= n:
                if s == n:
                    free.pop(j)
                else:
                    free[j] = (o + n, s - n)
                return o, n
        o = top
        top += n
        return o, n

    def release(free, o, n):
        free.append((o, n))
        free.sort()
        merged = []
        for o2, s2 in free:
            if merged and merged[-1][0] + merged[-1][1] == o2:
                merged[-1] = (merged[-1][0], merged[-1][1] + s2)
            else:
                merged.append((o2, s2))
        free[:] = merged

    # The library runs [ADD ->] DW3x3 -> PW runs as ONE kernel (csrc/node_umma.cu) whose CTAs
    # write the last op's output while other CTAs still read the first op's inputs: tensors
    # read inside such a run are recycled only after the run's last op has its output placed.
    run_end = fused_run_end(g)
    owner = {}
    pending = []                           # (op index after which the tensor is free, tensor id)
    g.tensors[g.input].ws_offset = -1      # the input lives in the caller's buffer
    for i, op in enumerate(g.ops):
        for item in [p for p in pending if p[0] < i]:
            pending.remove(item)
            tid = item[1]
            bfree, bsizes = pools[owner[tid]]
            release(bfree, g.tensors[tid].ws_offset, bsizes.pop(tid))
        free, sizes = pools.setdefault(op.branch, ([], {}))
        if op.out >= 0:
            t = g.tensors[op.out]
            t.ws_offset, sizes[op.out] = alloc(free, t.bytes_per_frame)
            owner[op.out] = op.branch
        for tid in set(op.inputs + ([op.residual] if op.residual >= 0 else [])):
            if tid == g.input or last_use.get(tid) != i or tid not in owner:
                continue
            b = owner[tid]
            if b == 0 and tid in keep:
                continue                   # read by a concurrent branch: stays until the end
            if b == op.branch:
                pending.append((run_end[i], tid))
    g.ws_bytes_per_frame = top
    return top


def dw_diag_blocks(w_q, c_p):
    """Depthwise weights as block-diagonal tensor-core operands (vbt_b200/csrc/dw_umma.cu).

    For every pair of 16-channel groups and every tap: a [32 out][32 in] int8 matrix whose
    diagonal holds the 32 tap weights, in the K-major no-swizzle core-matrix layout the MMA
    reads (byte (n, k) at (k // 16) * 512 + (n // 8) * 128 + (n % 8) * 16 + k % 16).
    w_q: int8 [c, k, k].  Returns int8 [pairs, k*k, 1024]."""
    c, k, _ = w_q.shape
    pairs = (c_p // 16 + 1) // 2
    wt = np.zeros((pairs * 32, k * k), np.int8)
    wt[:c] = w_q.reshape(c, k * k)
    out = np.zeros((pairs, k * k, 1024), np.int8)
    n = np.arange(32)
    off = (n // 16) * 512 + (n // 8) * 128 + (n % 8) * 16 + n % 16
    for p in range(pairs):
        out[p][:, off] = wt[p * 32:(p + 1) * 32].T
    return out


def _pack_op(rec):
    """OpRecord (vbt_b200/csrc/model.cuh), little endian."""
    f = rec
    out = struct.pack('<i3ii i ii ii ii iiii 3i i ii ii', f['type'], *f['in'], f['out'], f['n_in'],
                      f['k'], f['stride'], f['cin'], f['cout'], f['cin_p'], f['cout_p'],
                      f['h_in'], f['w_in'], f['h_out'], f['w_out'], *f['zp_in'], f['zp_out'],
                      f['act_lo'], f['act_hi'], f['pad_top'], f['pad_left'])
    out += struct.pack('<5q', f['w_off'], f['bias_off'], f['scale_off'], f['lut_off'],
                       f['out_elem_offset'])
    out += struct.pack('<3i i 3i 3i 3i ii 7i', *f['add_mult'], f['add_shift'], *f['resample'],
                       *f['in_h'], *f['in_w'], f['out_kind'], f['out_pix_stride'], f['branch'],
                       f.get('requant_fast', 0), f.get('pw_dtype', 0), *f.get('mb', [0, 0, 0, 0]))
    return out


OP_RECORD_BYTES = len(_pack_op(dict(
    type=0, **{'in': [0, 0, 0]}, out=0, n_in=0, k=0, stride=0, cin=0, cout=0, cin_p=0, cout_p=0,
    h_in=0, w_in=0, h_out=0, w_out=0, zp_in=[0, 0, 0], zp_out=0, act_lo=0, act_hi=0, pad_top=0,
    pad_left=0, w_off=0, bias_off=0, scale_off=0, lut_off=0, add_mult=[0, 0, 0], add_shift=0,
    resample=[0, 0, 0], in_h=[0, 0, 0], in_w=[0, 0, 0], out_kind=0, out_pix_stride=0,
    out_elem_offset=0, branch=0)))


def pack_blob(g: Graph):
    """Serialise a quantised graph for vbt_model_create."""
    assert getattr(g, 'quantized', False)
    plan_workspace(g)
    data = bytearray()

    def put(arr):
        while len(data) % 256:
            data.append(0)
        off = len(data)
        data.extend(np.ascontiguousarray(arr).tobytes())
        return off

    n = g.n_anchors
    n_pad = pad16(n)
    anchors_off = put(g.anchors())
    from_q = np.arange(-128, 128, dtype=np.int32)
    lut = np.exp(np.float32(g.box_scale) * (from_q - g.box_zp).astype(np.float32)).astype(np.float32)
    lut_off = put(lut)
    recs = []
    conv = {}                  # op index -> padded weights / folded bias / multiplier (MBConv chunk images)
    a_per = NUM_SCALES * len(ASPECTS)
    for oi, op in enumerate(g.ops):
        q = op.q
        ins = [g.tensors[i] for i in op.inputs]
        tout = g.tensors[op.out] if op.out >= 0 else None
        cout = g.out_channels(op)
        cout_p = tout.c_p if tout is not None else pad16(cout)    # head outputs: padded rows are zero
        r = dict(type=op.type, out=op.out if op.out >= 0 else -1, n_in=len(op.inputs), k=op.k,
                 stride=op.stride, cin=ins[0].c, cout=cout, cin_p=ins[0].c_p, cout_p=cout_p,
                 h_in=ins[0].h, w_in=ins[0].w,
                 h_out=tout.h if tout is not None else ins[0].h,
                 w_out=tout.w if tout is not None else ins[0].w,
                 zp_out=q['zp_out'], act_lo=q['act_lo'], act_hi=q['act_hi'],
                 pad_top=0, pad_left=0, w_off=-1, bias_off=-1, scale_off=-1, lut_off=-1,
                 add_mult=[0, 0, 0], add_shift=0, out_kind=op.out_kind,
                 out_pix_stride=cout_p, out_elem_offset=0, branch=op.branch)
        r['in'] = (op.inputs + [-1, -1, -1])[:3]
        r['zp_in'] = (q['zp_in'] + [0, 0, 0])[:3]
        r['resample'] = (list(op.resample) + [0, 0, 0])[:3]
        r['in_h'] = ([t.h for t in ins] + [0, 0, 0])[:3]
        r['in_w'] = ([t.w for t in ins] + [0, 0, 0])[:3]
        if op.type in (OP_DW, OP_MAXPOOL, OP_STEM):
            r['pad_top'] = same_pad(ins[0].h, op.k, op.stride)[1]
            r['pad_left'] = same_pad(ins[0].w, op.k, op.stride)[1]
        if op.type in (OP_STEM, OP_PW, OP_DW):
            w = q['w'].astype(np.int32)
            zin = q['zp_in'][0]
            if op.type == OP_PW:
                wp = np.zeros((cout_p, ins[0].c_p), np.int8)
                wp[:cout, :ins[0].c] = q['w']
                wsum = w.sum(axis=1)
            elif op.type == OP_DW:
                # [k*k, c_p] 32-bit words, weight of channel c in byte (c % 4) and zeros
                # elsewhere: dp4a(x_word, w_word) then multiplies exactly one channel of a
                # 4-channel activation word, no unpacking in the kernel
                wt = q['w'].reshape(cout, -1).T.astype(np.int64) & 0xff          # [k*k, cout]
                wp = np.zeros((op.k * op.k, cout_p), np.uint32)
                wp[:, :cout] = (wt << (8 * (np.arange(cout) % 4))[None, :]).astype(np.uint32)
                wsum = w.reshape(cout, -1).sum(axis=1)
            else:
                # stem: [9 pixel taps][cout_p] words = (w[.,0], w[.,1], w[.,2], 0): one dp4a
                # per RGB pixel of the 3x3 window
                w9 = q['w'].reshape(cout, 9, 3).astype(np.int64) & 0xff
                wp = np.zeros((9, cout_p), np.uint32)
                wp[:, :cout] = (w9[:, :, 0] | (w9[:, :, 1] << 8) | (w9[:, :, 2] << 16)).T.astype(np.uint32)
                wsum = w.reshape(cout, -1).sum(axis=1)
            bias = np.zeros(cout_p, np.int32)
            bias[:cout] = (q['bias'].astype(np.int64) - zin * wsum).astype(np.int32)
            mult = np.zeros(cout_p, np.float32)
            mult[:cout] = q['mult']
            r['w_off'], r['bias_off'], r['scale_off'] = put(wp), put(bias), put(mult)
            conv[oi] = dict(w=wp if op.type == OP_PW else q['w'], bias=bias, mult=mult)
            # the kernels' packed requantisation carries rint(acc * M) in 16-bit lanes: allowed only
            # where no input can push |acc * M| past 2^15 - 256 (csrc/requant.cuh)
            wabs = np.abs(w).reshape(cout, -1).sum(axis=1)
            bound = (255.0 * wabs + np.abs(q['bias'].astype(np.float64))) * np.abs(q['mult'].astype(np.float64))
            r['requant_fast'] = int(bound.max() < 32000.0)
            r['zp_out'] = q['conv_zp_out']
            if op.type == OP_PW and op.residual >= 0:
                r['in'][1] = op.residual
                r['n_in'] = 2
                r['zp_in'][1] = q['res_zp']
                r['add_mult'][:2] = q['add_mult']
                r['add_shift'] = q['add_shift']
                # zp_out = conv result zero point; the final one rides in zp_in[2]
                r['zp_in'][2] = q['zp_out']
            if 'lut' in q:
                r['lut_off'] = put(q['lut'])
            if op.type == OP_DW:
                r['lut_off'] = put(dw_diag_blocks(q['w'], cout_p))
        elif op.type == OP_ADD:
            r['add_mult'] = (q['add_mult'] + [0, 0, 0])[:3]
            r['add_shift'] = q['add_shift']
        if op.type == OP_PW and getattr(g, 'head_dtype', 'int8') == 'bf16' and \
                (op.branch > 0 or getattr(g, 'variant', '') == 'micro'):
            r['pw_dtype'] = 1          # head contraction on the bf16 tensor path (same integers)
        if op.out_kind == 1:
            r['out_pix_stride'] = a_per * NUM_CLASSES
            r['out_elem_offset'] = op.level_offset * NUM_CLASSES
        elif op.out_kind == 2:
            r['out_pix_stride'] = a_per * 4
            r['out_elem_offset'] = op.level_offset * 4
        recs.append(r)
    # MBConv blocks: per-chunk weight images for the fused kernel, announced on the depthwise op's
    # record (mb = [image offset / 256 + 1, image stride, chunks, first op of the run relative to it])
    for first, (ei, di, pi) in mbconv_runs(g).items():
        d_op = g.ops[di]
        cin_p = 32 if g.ops[first].type == OP_STEM else g.tensors[g.ops[first].inputs[0]].c_p   # stem: 27 im2col bytes -> 32
        img, stride, n_chunks = mbconv_images(cin_p, d_op.k, recs[pi]['cout_p'], conv[ei] if ei >= 0 else None,
                                              conv[di], conv[pi])
        off = put(img)
        assert off % 256 == 0
        recs[di]['mb'] = [off // 256 + 1, stride, n_chunks, first - di]
    recs = [_pack_op(r) for r in recs]
    tens = b''.join(struct.pack('<q4i', t.ws_offset, t.h, t.w, t.c, t.c_p) for t in g.tensors)
    header_bytes = 128
    table = header_bytes + OP_RECORD_BYTES * len(recs) + 24 * len(g.tensors)
    data_offset = (table + 255) // 256 * 256
    hdr = struct.pack('<I7i q q q q q f i i 11i', BLOB_MAGIC, BLOB_VERSION, g.S, n, n_pad,
                      NUM_CLASSES, len(recs), len(g.tensors), g.ws_bytes_per_frame, data_offset,
                      len(data), anchors_off, lut_off, g.box_scale, g.box_zp,
                      g.tensors[g.input].zp, *([0] * 11))
    assert len(hdr) == header_bytes, len(hdr)
    blob = bytearray(hdr) + b''.join(recs) + tens
    blob.extend(b'\0' * (data_offset - len(blob)))
    blob.extend(data)
    return bytes(blob)


def build_synthetic(variant='lite0', seed=1234, calib_frames=None, n_calib=4, head_dtype='int8'):
    """Seeded synthetic model: float init -> PTQ calibration -> quantised Graph.
    head_dtype: 'int8' or 'bf16' -- the tensor-core data type of the class / box nets' pointwise
    convs (BASELINE.json configs[3]); the quantised model and its outputs are the same."""
    assert head_dtype in ('int8', 'bf16')
    g = Graph(variant)
    g.head_dtype = head_dtype
    init_weights(g, seed)
    if calib_frames is None:
        from .synth import synthetic_model_inputs
        calib_frames = synthetic_model_inputs(n_calib, g.S, seed=seed + 1)
    if variant != 'lite0':
        calibrate_class_prior(g, calib_frames)
    quantize(g, calib_frames)
    return g


def calibrate_class_prior(g, frames, per_frame=4.0):
    """Shift the class head's bias so that about `per_frame` anchors per frame clear score 0.5 -- what a
    trained single-class detector shows on these videos (1-3 plates) and what the
    Lite0 initialisation gives by itself (3-5).  The wider / deeper Lite1 and Lite2 random heads would
    otherwise fire on hundreds of anchors, every frame would return all 25 detections and the tracker
    would carry ~30 live tracks: a workload no real clip produces (round-1 bench lines of configs[2]/[3]
    were tracker-bound for that reason).  Lite0 is left exactly as it was."""
    _, outs = float_forward(g, frames)
    logits = np.concatenate([o.reshape(o.shape[0], -1).numpy() for o in outs[1]], axis=1)
    t = float(np.quantile(logits, 1.0 - per_frame / logits.shape[1]))
    for op in g.ops:
        if op.out_kind == 1:
            op.bias = (op.bias - np.float32(_coarse(t))).astype(np.float32)


def anchors_only(variant='lite0', box_scale=0.05, box_zp=0):
    """A graph with no ops: enough for vbt_postprocess_q8 (anchors + box dequantisation)."""
    g = Graph(variant)
    g.ops = []
    g.box_scale, g.box_zp = float(np.float32(box_scale)), int(box_zp)
    g.quantized = True
    return g
