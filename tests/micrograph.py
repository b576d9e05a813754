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
    cout = g.tensors[op.out].c
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
    for op in g.ops:
        t = g.tensors[op.out]
        off = B * t.ws_offset
        full = ws[off:off + B * t.h * t.w * t.c_p].reshape(B, t.h, t.w, t.c_p)
        out[op.out] = (full[..., :t.c].copy(), full[..., t.c:].copy())
    return out
