"""Host-side planning of the layer program (vbt_b200/effdet.py), on the CPU:
* the workspace plan never lets a tensor written by a fused [[ADD ->] ADD ->] DW3x3 -> PW run, or by a
  fused MBConv block ([expand ->] depthwise -> project), alias a tensor the run reads -- the fused kernel's CTAs write the output while other CTAs still read the
  inputs (the aliasing the library re-checks in vbt_model_create);
* liveness reuse stays safe for the single ops too (an op's output never overlaps its own inputs,
  nor any tensor that is still to be read later);
* head chains are independent branches with private memory."""
import re

import pytest

from vbt_b200 import effdet as E


def _span(g, t):
    T = g.tensors[t]
    return T.ws_offset, T.ws_offset + T.bytes_per_frame


def _inputs(op):
    return op.inputs + ([op.residual] if op.residual >= 0 else [])


@pytest.mark.parametrize('variant', ['lite0', 'lite1', 'lite2'])
def test_fused_runs_never_alias_their_inputs(variant):
    g = E.Graph(variant)
    E.plan_workspace(g)
    end = E.fused_run_end(g)
    runs = {}
    for i, e in enumerate(end):
        runs.setdefault(e, []).append(i)
    n_runs = 0
    for e, idx in runs.items():
        if len(idx) == 1:
            continue
        n_runs += 1
        assert idx == list(range(idx[0], e + 1)) and len(idx) in (2, 3, 4)
        produced = {g.ops[i].out for i in idx}
        for i in idx:
            out = g.ops[i].out
            if out < 0:
                continue
            o0, o1 = _span(g, out)
            for j in idx:
                for t in _inputs(g.ops[j]):
                    if t == g.input or t in produced:
                        continue
                    a0, a1 = _span(g, t)
                    assert o1 <= a0 or a1 <= o0, (variant, g.ops[i].name, g.ops[j].name)
    # every BiFPN node and every head stage is a run: cells * 8 nodes + 5 levels * 2 nets * 4 stages;
    # and every MBConv block of the backbone is one (16 / 21 / 21 blocks)
    n_blocks = sum(1 for op in g.ops if re.fullmatch(r'b\d+\.\d+\.dw', op.name))
    assert n_blocks == {'lite0': 16, 'lite1': 21, 'lite2': 21}[variant]
    assert len(E.mbconv_runs(g)) == n_blocks
    assert n_runs == g.cells * 8 + 40 + n_blocks


@pytest.mark.parametrize('variant', ['lite0', 'lite2'])
def test_liveness_reuse_is_safe(variant):
    g = E.Graph(variant)
    top = E.plan_workspace(g)
    last_use = {}
    for i, op in enumerate(g.ops):
        for t in _inputs(op):
            last_use[t] = i
    for i, op in enumerate(g.ops):
        if op.out < 0:
            continue
        o0, o1 = _span(g, op.out)
        assert 0 <= o0 and o1 <= top and o0 % 256 == 0
        for t, lu in last_use.items():
            if t == g.input or t == op.out or lu < i:
                continue                              # dead before this op writes
            producer = next((k for k, p in enumerate(g.ops) if p.out == t), -1)
            if producer > i:
                continue                              # not born yet
            a0, a1 = _span(g, t)
            assert o1 <= a0 or a1 <= o0, (variant, op.name, g.tensors[t].name)


def test_head_chains_are_independent_branches():
    g = E.Graph('lite0')
    E.plan_workspace(g)
    branches = sorted({op.branch for op in g.ops})
    assert branches == list(range(11))                # trunk + 5 levels x (class, box)
    spans = {}
    for op in g.ops:
        if op.branch > 0 and op.out >= 0:
            spans.setdefault(op.branch, []).append(_span(g, op.out))
    for b1 in spans:
        for b2 in spans:
            if b1 < b2:
                for a0, a1 in spans[b1]:
                    for c0, c1 in spans[b2]:
                        assert a1 <= c0 or c1 <= a0, (b1, b2)
    # a branch's ops are contiguous in program order (the library forks one stream per branch)
    order = [op.branch for op in g.ops if op.branch > 0]
    assert order == sorted(order)
