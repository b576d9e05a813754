"""The product's layer program (vbt_b200/effdet.py) against an independently written description of
EfficientDet-Lite0/1/2 (oracle/arch.py; it does not import vbt_b200):
the two must be the same dataflow graph.  Every GPU-vs-oracle parity test interprets the PRODUCT's
graph, so this is the test that sees a wrong block order, kernel size, stride, channel width, residual,
BiFPN edge, resampling mode, head layout or anchor."""
import numpy as np
import pytest

from oracle import arch as A
from vbt_b200 import effdet as E


def product_ids(g, intern):
    """Structural ids of the product graph's tensors, in oracle/arch.py's vocabulary."""
    ids = {g.input: intern('input', g.S)}
    outputs = []
    for op in g.ops:
        ins = [ids[i] for i in op.inputs]
        if op.type == E.OP_STEM:
            assert (op.k, op.stride, op.act) == (3, 2, True)
            v = intern('stem', g.tensors[op.out].c, ins[0])
        elif op.type == E.OP_PW:
            v = intern('pw', g.out_channels(op), bool(op.act), ins[0], ids[op.residual] if op.residual >= 0 else -1)
        elif op.type == E.OP_DW:
            v = intern('dw', op.k, op.stride, bool(op.act), ins[0])
        elif op.type == E.OP_MAXPOOL:
            assert (op.k, op.stride) == (3, 2)
            v = intern('maxpool', ins[0])
        elif op.type == E.OP_ADD:
            mode = {E.RS_NONE: 'same', E.RS_UP: 'up', E.RS_DOWN: 'down'}
            v = intern('sum', bool(op.act), tuple((mode[r], i) for r, i in zip(op.resample, ins)))
        else:
            raise AssertionError(op.type)
        if op.out >= 0:
            ids[op.out] = v
        else:
            outputs.append(({1: 'cls', 2: 'box'}[op.out_kind], op.level_offset, v))
    return ids, outputs


@pytest.mark.parametrize('variant', ['lite0', 'lite1', 'lite2'])
def test_product_graph_equals_independent_description(variant):
    intern = A.Interner()
    want = A.build(variant, intern)
    g = E.Graph(variant)
    _, outputs = product_ids(g, intern)
    assert g.n_anchors == want['n_anchors'] == A.ANCHORS[variant]
    assert [s[0] for s in g.level_sizes] == want['level_sizes'] == [s[1] for s in g.level_sizes]
    # same ten output computations at the same anchor offsets (order-insensitive)
    assert sorted(outputs) == sorted(want['outputs'])
    # and nothing else in the program: every op of the product feeds an output
    used = set()
    for op in reversed(g.ops):
        if op.out < 0 or op.out in used:
            used.update(op.inputs + ([op.residual] if op.residual >= 0 else []))
    assert all(op.out < 0 or op.out in used for op in g.ops)


def test_shapes_follow_from_the_wiring():
    """Tensor shapes of the product graph obey SAME-padding arithmetic op by op (a wrong shape with the
    right wiring would otherwise pass the structural check)."""
    for variant in ('lite0', 'lite1', 'lite2'):
        g = E.Graph(variant)
        for op in g.ops:
            t_in = g.tensors[op.inputs[0]]
            if op.out < 0:
                continue
            t = g.tensors[op.out]
            if op.type in (E.OP_STEM, E.OP_DW, E.OP_MAXPOOL):
                assert (t.h, t.w) == (-(-t_in.h // op.stride), -(-t_in.w // op.stride)), op.name
            elif op.type == E.OP_PW:
                assert (t.h, t.w) == (t_in.h, t_in.w)
                if op.residual >= 0:
                    r = g.tensors[op.residual]
                    assert (r.h, r.w, r.c) == (t.h, t.w, t.c)
            elif op.type == E.OP_ADD:
                for i, rs in zip(op.inputs, op.resample):
                    ti = g.tensors[i]
                    assert ti.c == t.c
                    if rs == E.RS_NONE:
                        assert (ti.h, ti.w) == (t.h, t.w)
                    elif rs == E.RS_UP:
                        assert ti.h < t.h and t.h in (2 * ti.h, 2 * ti.h - 1)
                    else:
                        assert (-(-ti.h // 2), -(-ti.w // 2)) == (t.h, t.w)


def test_lite0_stage_macs_match_appendix_a2():
    """Per-stage pointwise / depthwise MACs of Lite0 @320 against the SURVEY appendix A.2 table, from
    both descriptions."""
    want = A.build('lite0', A.Interner())['stage_macs']
    g = E.Graph('lite0')
    got = {}
    for op in g.ops:
        name = op.name
        if name == 'stem':
            key = 'stem'
        elif name[0] == 'b' and name[1].isdigit():
            key = int(name[1:name.index('.')])
        else:
            continue
        m = got.setdefault(key, [0, 0])
        t = g.tensors[op.out]
        if op.type == E.OP_STEM:
            m[0] += t.h * t.w * t.c * 27
        elif op.type == E.OP_PW:
            m[0] += t.h * t.w * t.c * g.tensors[op.inputs[0]].c
        else:
            m[1] += t.h * t.w * t.c * op.k * op.k
    assert got == want
    for key, (pw, dw) in A.LITE0_STAGE_MMACS.items():
        assert abs(got[key][0] / 1e6 - pw) < 0.06 and abs(got[key][1] / 1e6 - dw) < 0.06, (key, got[key])


@pytest.mark.parametrize('variant', ['lite0', 'lite1', 'lite2'])
def test_anchor_table(variant):
    a, b = E.Graph(variant).anchors(), A.anchors(variant)
    assert a.shape == b.shape == (A.ANCHORS[variant], 4)
    assert np.allclose(a, b, rtol=0, atol=1e-6)
    assert np.array_equal(a, b)
