"""One- and two-op layer programs with hand-set quantisation, for testing a single kernel of
libvbt_b200.so against oracle/effdet.py at arbitrary shapes (the full networks only visit
the shapes of EfficientDet-Lite0/1/2).

The graph's input tensor is an int8 activation [B,h,w,c_p] supplied by the test (vbt_detect
lets any op read the model input); results are read back from the workspace."""
import math

import numpy as np

from vbt_b200 import effdet as E


def _empty_graph(h, w, c, zp):
    g = E.Graph.__new__(E.Graph)
    g.variant, g.S = 'micro', h
    g.tensors, g.ops = [], []
    g.level_sizes, g.n_anchors = [(1, 1)], 9
    g.box_scale, g.box_zp = float(np.float32(0.05)), 0
    g.input = g._t(h, w, c, 'input')
    g.tensors[g.input].zp = zp
    g.quantized = True
    return g


def _conv_q(rng, op, g, zp_in, zp_out, act, fan_in):
    cout = g.out_channels(op)
    q = op.q
    q['zp_in'] = [zp_in]
    q['zp_out'] = q['conv_zp_out'] = zp_out
    q['act_lo'], q['act_hi'] = (zp_out, min(127, zp_out + 200)) if act else (-128, 127)
    q['bias'] = rng.integers(-3000, 3000, cout).astype(np.int32)
    acc_std = math.sqrt(fan_in) * 74.0 * 73.0
    q['mult'] = (rng.uniform(0.5, 1.5, cout) * 45.0 / acc_std).astype(np.float32)


def pw_graph(h, w, cin, cout, act=False, residual=False, seed=0, zp_in=-7, zp_out=11):
    """input -> PW(cin->cout).  With residual: input -> PW(cin->cmid=cout) -> PW(cout->cin)
    + input (quantised add), the MBConv project shape."""
    rng = np.random.default_rng(seed)
    g = _empty_graph(h, w, cin, zp_in)
    o = g._pw(g.input, cout, act, 'pw0')
    op = g.ops[-1]
    op.q['w'] = rng.integers(-127, 128, (cout, cin)).astype(np.int8)
    _conv_q(rng, op, g, zp_in, zp_out, act, cin)
    g.tensors[o].zp = zp_out
    if residual:
        o2 = g._pw(o, cin, False, 'pw1', residual=g.input)
        op2 = g.ops[-1]
        op2.q['w'] = rng.integers(-127, 128, (cin, cout)).astype(np.int8)
        _conv_q(rng, op2, g, zp_out, 5, False, cout)
        q = op2.q
        q['conv_zp_out'] = 5                 # requant target of the conv itself
        q['zp_out'] = -3                     # zero point of the sum
        q['res_zp'] = zp_in
        q['add_mult'], q['add_shift'] = [int(0.61 * (1 << 20)), int(0.83 * (1 << 20))], 20
        q['act_lo'], q['act_hi'] = -128, 127
        g.tensors[o2].zp = -3
    return g


def dw_graph(h, w, c, k, stride, act=True, seed=0, zp_in=-9, zp_out=-20):
    rng = np.random.default_rng(seed)
    g = _empty_graph(h, w, c, zp_in)
    o = g._dw(g.input, k, stride, act, 'dw0')
    op = g.ops[-1]
    op.q['w'] = rng.integers(-127, 128, (c, k, k)).astype(np.int8)
    _conv_q(rng, op, g, zp_in, zp_out, act, k * k)
    g.tensors[o].zp = zp_out
    return g


def _set_dw(rng, g, op, zp_in, zp_out, act):
    c = g.tensors[op.out].c
    op.q['w'] = rng.integers(-127, 128, (c, op.k, op.k)).astype(np.int8)
    _conv_q(rng, op, g, zp_in, zp_out, act, op.k * op.k)
    g.tensors[op.out].zp = zp_out


def _set_pw(rng, g, op, cin, cout, zp_in, zp_out, act):
    op.q['w'] = rng.integers(-127, 128, (cout, cin)).astype(np.int8)
    _conv_q(rng, op, g, zp_in, zp_out, act, cin)
    if op.out >= 0:
        g.tensors[op.out].zp = zp_out


def head_graph(h, w, c, cout, act=True, seed=0, out_kind=0):
    """input -> DW3x3 s1 -> PW(c->cout): one stage of the class / box nets, the pair
    csrc/node_umma.cu runs as one kernel.  out_kind 1 / 2: packed raw head outputs."""
    rng = np.random.default_rng(seed)
    g = _empty_graph(h, w, c, -9)
    if out_kind:
        g.level_sizes, g.n_anchors = [(h, w)], h * w * 9
    d = g._dw(g.input, 3, 1, False, 'dw0')
    _set_dw(rng, g, g.ops[-1], -9, -20, False)
    g._pw(d, cout, act, 'pw0', out_kind=out_kind)
    _set_pw(rng, g, g.ops[-1], c, cout, -20, 11, act)
    if out_kind == 1:
        qs = np.arange(-128, 128)
        g.ops[-1].q['lut'] = np.clip(np.rint(256.0 / (1.0 + np.exp(-(qs - 11) * 0.06))) - 128,
                                     -128, 127).astype(np.int8)
    return g


def node_graph(h, w, c, n_in=3, seed=0, odd=False, tree=False):
    """A BiFPN node on a pyramid built from the input: input [2h(-1), 2w(-1)] -> DW s2 -> p1
    [h, w] -> DW s2 -> p2; then ADD(max-pooled input, p1, up-sampled p2) -> ReLU6 -> DW3x3 ->
    PW(c->c), the triple csrc/node_umma.cu runs as one kernel.  n_in = 2 drops the max-pooled
    input.  odd: the input is (2h-1, 2w-1), the 5 -> 3 / 7 -> 4 pooling case.  tree (n_in = 3):
    the sum as the exported graphs hold it, ADD(ADD(pooled input, p1), up-sampled p2), the inner
    ADD with its own output quantisation and no activation."""
    rng = np.random.default_rng(seed)
    g = _empty_graph(2 * h - (1 if odd else 0), 2 * w - (1 if odd else 0), c, 4)
    g.S = g.tensors[g.input].h
    p1 = g._dw(g.input, 3, 2, False, 'p1')
    _set_dw(rng, g, g.ops[-1], 4, -13, False)
    assert (g.tensors[p1].h, g.tensors[p1].w) == (h, w)
    p2 = g._dw(p1, 3, 2, False, 'p2')
    _set_dw(rng, g, g.ops[-1], -13, 21, False)
    xs = [g.input, p1, p2] if n_in == 3 else [p1, p2]
    if tree:
        assert n_in == 3
        t0 = g._fuse(xs[:2], (h, w), 'n.sum0', act=False)
        op0 = g.ops[-1]
        op0.q['zp_in'] = [g.tensors[x].zp for x in xs[:2]]
        op0.q['zp_out'], op0.q['act_lo'], op0.q['act_hi'] = 9, -128, 127
        op0.q['add_mult'], op0.q['add_shift'] = [int(0.37 * (1 << 20)), int(0.58 * (1 << 20))], 20
        g.tensors[t0].zp = 9
        xs, n_in = [t0, xs[2]], 2
    f = g._fuse(xs, (h, w), 'n.sum')
    op = g.ops[-1]
    zp_out = -128
    op.q['zp_in'] = [g.tensors[x].zp for x in xs]
    op.q['zp_out'] = zp_out
    op.q['act_lo'], op.q['act_hi'] = zp_out, 95
    op.q['add_mult'] = [int(m * (1 << 20)) for m in (0.43, 0.71, 0.52)][3 - n_in:]
    op.q['add_shift'] = 20
    g.tensors[f].zp = zp_out
    d = g._dw(f, 3, 1, False, 'n.dw')
    _set_dw(rng, g, g.ops[-1], zp_out, -6, False)
    g._pw(d, c, False, 'n.pw')
    _set_pw(rng, g, g.ops[-1], c, c, -6, 17, False)
    return g


def mbconv_graph(h, w, cin, cexp, cout, k, stride, residual=False, seed=0, expand=True):
    """input -> PW(cin->cexp, ReLU6) -> DW kxk stride s (ReLU6) -> PW(cexp->cout) [+ input]: one MBConv
    block, the run csrc/mbconv_umma.cu executes as one kernel.  expand=False: DW -> PW only (the
    backbone's first block, expand ratio 1)."""
    rng = np.random.default_rng(seed)
    zp_in = -7
    g = _empty_graph(h, w, cin, zp_in)
    x = g.input
    if expand:
        x = g._pw(x, cexp, True, 'mb.expand')
        _set_pw(rng, g, g.ops[-1], cin, cexp, zp_in, -128, True)
        zp_mid = -128
    else:
        assert cexp == cin
        zp_mid = zp_in
    d = g._dw(x, k, stride, True, 'mb.dw')
    _set_dw(rng, g, g.ops[-1], zp_mid, -120, True)
    if residual:
        assert stride == 1 and cin == cout and expand
    o = g._pw(d, cout, False, 'mb.project', residual=g.input if residual else -1)
    op = g.ops[-1]
    _set_pw(rng, g, op, cexp, cout, -120, 5 if residual else 9, False)
    if residual:
        q = op.q
        q['conv_zp_out'] = 5
        q['zp_out'] = -3
        q['res_zp'] = zp_in
        q['add_mult'], q['add_shift'] = [int(0.61 * (1 << 20)), int(0.83 * (1 << 20))], 20
        q['act_lo'], q['act_hi'] = -128, 127
        g.tensors[o].zp = -3
    return g


def stem_block_graph(h, w, cout, k, stride, cmid=32, seed=0):
    """uint8 frame [h,w,3] -> STEM 3x3 s2 (3 -> cmid, ReLU6) -> DW kxk (ReLU6) -> PW(cmid -> cout): the network's
    first block with the stem as its expand stage (csrc/mbconv_umma.cu, im2col rows of the frame)."""
    rng = np.random.default_rng(seed)
    g = _empty_graph(h, w, 3, 127)
    ho, wo = E.same_pad(h, 3, 2)[0], E.same_pad(w, 3, 2)[0]
    x = g._t(ho, wo, cmid, 'stem')
    g.ops.append(E.Op(E.OP_STEM, [g.input], x, k=3, stride=2, act=True, name='stem'))
    op = g.ops[-1]
    op.q['w'] = rng.integers(-127, 128, (cmid, 3, 3, 3)).astype(np.int8)
    _conv_q(rng, op, g, 127, -128, True, 27)
    g.tensors[x].zp = -128
    d = g._dw(x, k, stride, True, 'b1.dw')
    _set_dw(rng, g, g.ops[-1], -128, -120, True)
    g._pw(d, cout, False, 'b1.project')
    _set_pw(rng, g, g.ops[-1], cmid, cout, -120, 9, False)
    return g


def random_input(g, B, seed=1):
    """(logical int8 [B,h,w,c], padded int8 [B,h,w,c_p] with the zero point in the pad)."""
    t = g.tensors[g.input]
    rng = np.random.default_rng(seed)
    x = rng.integers(-128, 128, (B, t.h, t.w, t.c)).astype(np.int8)
    xp = np.full((B, t.h, t.w, t.c_p), t.zp, np.int8)
    xp[..., :t.c] = x
    return x, xp


def run_gpu(g, xp):
    """Run the layer program on the GPU; returns {tensor id: int8 [B,h,w,c] logical}."""
    import torch
    from vbt_b200.interpreter import Detector
    B = xp.shape[0]
    det = Detector(g, max_batch=B)
    dev = torch.as_tensor(np.ascontiguousarray(xp), device='cuda')
    det.network(dev.view(torch.uint8))
    torch.cuda.synchronize()
    ws = det.workspace.cpu().numpy().view(np.int8)
    out = {}
    plan = det.plan()
    run_gpu.last_plan = plan
    for i, op in enumerate(g.ops):
        # tensors inside a fused run never reach the workspace: only a launch's last op does
        j = i
        while plan[j] == 0:
            j -= 1
        if i != j + plan[j] - 1:
            continue
        if op.out < 0:
            n = g.n_anchors
            raw = (det.raw_cls if op.out_kind == 1 else det.raw_box)[:B].cpu().numpy()
            out[-op.out_kind] = raw[:, :n] if op.out_kind == 1 else raw[:, :n, :]
            continue
        t = g.tensors[op.out]
        off = B * t.ws_offset
        full = ws[off:off + B * t.h * t.w * t.c_p].reshape(B, t.h, t.w, t.c_p)
        out[op.out] = (full[..., :t.c].copy(), full[..., t.c:].copy())
    return out
