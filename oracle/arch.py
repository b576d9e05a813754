"""Independent description of the EfficientDet-Lite0/1/2 architecture -- TEST INFRASTRUCTURE.

Written from SURVEY.md appendix A (tflite-model-maker 0.4.3 / google automl `efficientdet`, the spec
train.py:23,29 names) WITHOUT importing vbt_b200: the product's layer program (vbt_b200/effdet.py) and
this file are two separate readings of the same appendix, and tests/test_oracle_arch.py requires them
to describe the same dataflow graph -- block order, kernel sizes, strides, channel widths, residuals,
BiFPN wiring and resampling modes, head layout, anchor table.  An architecture error in one of them
is then visible, which "GPU == oracle on the product's own graph" alone cannot show.

A graph is reduced to STRUCTURAL ids: every tensor gets the id of the tuple (producer op kind, its
parameters, ids of its inputs), interned in a table shared by both builders.  Equal ids = the same
computation, regardless of op order or naming."""
from __future__ import annotations

import math

import numpy as np

SPEC = {
    # input size, width coefficient, depth coefficient, BiFPN channels, BiFPN cells, head repeats
    'lite0': (320, 1.0, 1.0, 64, 3, 3),
    'lite1': (384, 1.0, 1.1, 88, 4, 3),
    'lite2': (448, 1.1, 1.2, 112, 5, 3),
}
# appendix A.2: (kernel, repeats at depth 1.0, output channels at width 1.0, expand ratio, stride)
STAGES = [(3, 1, 16, 1, 1), (3, 2, 24, 6, 2), (5, 2, 40, 6, 2), (3, 3, 80, 6, 2), (5, 3, 112, 6, 1),
          (5, 4, 192, 6, 2), (3, 1, 320, 6, 1)]
# appendix A.2, Lite0 @320: stage -> (pointwise MACs, depthwise MACs) in millions (rounded to 0.1)
LITE0_STAGE_MMACS = {'stem': (22.1, 0.0), 1: (13.1, 7.4), 2: (98.3, 13.8), 3: (62.1, 15.4), 4: (84.5, 4.3),
                     5: (157.3, 18.2), 6: (175.7, 10.3), 7: (59.0, 1.0)}
LEVEL_SIZES = {'lite0': (40, 20, 10, 5, 3), 'lite1': (48, 24, 12, 6, 3), 'lite2': (56, 28, 14, 7, 4)}
ANCHORS = {'lite0': 19206, 'lite1': 27621, 'lite2': 37629}


class Interner:
    def __init__(self):
        self.table = {}

    def __call__(self, *key):
        return self.table.setdefault(key, len(self.table))


def _filters(c, width):
    """EfficientNet round_filters with divisor 8."""
    c = c * width
    new = max(8, int(c + 4) // 8 * 8)
    if new < 0.9 * c:
        new += 8
    return int(new)


def _out(size, stride):
    return -(-size // stride)          # SAME padding


def build(variant, intern: Interner):
    """Returns dict(outputs=[(kind, level offset in anchors, id)], n_anchors, stage_macs) of the
    architecture as structural ids.  Op keys (shared vocabulary with the product-side walker in
    tests/test_oracle_arch.py):
      ('input', S)
      ('stem', cout, in)                                 3x3 s2 conv + ReLU6
      ('pw', cout, relu6, in, residual id or -1)         1x1 conv (+ residual add)
      ('dw', k, stride, relu6, in)                       depthwise, SAME
      ('maxpool', in)                                    3x3 s2 SAME
      ('sum', relu6, ((resample mode, id), ...))         modes: 'same', 'up' (nearest), 'down' (max-pool)
    """
    S, width, depth, C, cells, head_rep = SPEC[variant]
    x = intern('input', S)
    size = _out(S, 2)
    x = intern('stem', 32, x)
    cin = 32
    feats, sizes = {}, {}
    macs = {'stem': [size * size * 32 * 27, 0]}
    for si, (k, rep, cout, expand, stride) in enumerate(STAGES, start=1):
        cout = _filters(cout, width)
        if 1 < si < len(STAGES):
            rep = int(math.ceil(depth * rep))            # first and last stage are not depth-scaled
        m = macs.setdefault(si, [0, 0])
        for r in range(rep):
            s = stride if r == 0 else 1
            block_in = x
            if expand != 1:
                x = intern('pw', cin * expand, True, x, -1)
                m[0] += size * size * cin * cin * expand
            mid = cin * expand
            size = _out(size, s)
            x = intern('dw', k, s, True, x)
            m[1] += size * size * mid * k * k
            skip = block_in if (s == 1 and cin == cout) else -1
            x = intern('pw', cout, False, x, skip)
            m[0] += size * size * mid * cout
            cin = cout
        if si in (3, 5, 7):
            feats[3 + (si - 3) // 2] = x
            sizes[3 + (si - 3) // 2] = size
    # BiFPN
    p6 = intern('maxpool', intern('pw', C, False, feats[5], -1))
    sizes[6] = _out(sizes[5], 2)
    p7 = intern('maxpool', p6)
    sizes[7] = _out(sizes[6], 2)
    assert tuple(sizes[l] for l in (3, 4, 5, 6, 7)) == LEVEL_SIZES[variant]
    level_of = [3, 4, 5, 6, 7]
    entering = [feats[3], feats[4], feats[5], p6, p7]
    backbone = {feats[3], feats[4], feats[5]}
    nodes = [(6, (3, 4)), (5, (2, 5)), (4, (1, 6)), (3, (0, 7)), (4, (1, 7, 8)), (5, (2, 6, 9)), (6, (3, 5, 10)), (7, (4, 11))]
    for _ in range(cells):
        feat = list(entering)
        lvl = list(level_of)
        for level, ins in nodes:
            terms = []
            for j in ins:
                t = feat[j]
                if t in backbone:                        # backbone feature: its own 1x1 lateral conv per use
                    t = intern('pw', C, False, t, -1)
                mode = 'same' if lvl[j] == level else ('up' if lvl[j] > level else 'down')
                terms.append((mode, t))
            if len(terms) == 3:                          # add_n of three = ADD(ADD(a, b), c) in the exported graph
                inner = intern('sum', False, tuple(terms[:2]))
                terms = [('same', inner), terms[2]]
            f = intern('sum', True, tuple(terms))
            d = intern('dw', 3, 1, False, f)
            feat.append(intern('pw', C, False, d, -1))
            lvl.append(level)
        entering = feat[8:13]                            # newest output of levels 3..7
        backbone = set()
    # heads: class net then box net per level, 9 anchors per location
    outputs, off = [], 0
    for li, level in enumerate((3, 4, 5, 6, 7)):
        for kind, cout in (('cls', 9), ('box', 36)):
            h = entering[li]
            for _ in range(head_rep):
                h = intern('pw', C, True, intern('dw', 3, 1, False, h), -1)
            h = intern('pw', cout, False, intern('dw', 3, 1, False, h), -1)
            outputs.append((kind, off, h))
        off += sizes[level] * sizes[level] * 9
    assert off == ANCHORS[variant]
    return dict(outputs=outputs, n_anchors=off, stage_macs=macs, level_sizes=[sizes[l] for l in (3, 4, 5, 6, 7)])


def anchors(variant):
    """f32 [N,4] (ycentre, xcentre, h, w) normalised (appendix A, 'Anchors'): level-major, then y, x,
    then octave-major / aspect-minor; anchor_scale 3, 3 octaves, aspects 1, 2, 0.5."""
    S = float(SPEC[variant][0])
    out = []
    for f in LEVEL_SIZES[variant]:
        stride = S / f
        for y in range(f):
            for x in range(f):
                for o in range(3):
                    for a in (1.0, 2.0, 0.5):
                        base = 3.0 * stride * 2.0 ** (o / 3.0)
                        w, h = base * math.sqrt(a), base / math.sqrt(a)
                        out.append(((stride / 2 + y * stride) / S, (stride / 2 + x * stride) / S, h / S, w / S))
    return np.asarray(out, dtype=np.float32)
