"""Second, INDEPENDENT CPU implementation of the detector's int8 convolutions -- TEST INFRASTRUCTURE.

oracle/effdet.py evaluates every int8 convolution as an exact floating-point convolution followed
by its own restatement of the requantisation.  This module runs the same ops through PyTorch's
quantized CPU kernels instead (`torch.ao.nn.quantized.functional.conv2d`: fbgemm / oneDNN int8 GEMM
with int32 accumulation and fp32 requantisation, the engine family TFLite's XNNPACK path belongs
to): a production int8 inference engine that shares no code with this repository.  Two uses:

* `check_convs(g, frames)`: per-op agreement between oracle/effdet.py and the quantized engine on
  the oracle's own input tensors (tests/test_oracle_second_opinion.py pins it at >= 99.9 % equal,
  never more than one quantisation step apart);
* `run(g, frames)`: the whole network chained through the engine -- the realistic "int8 interpreter
  on the host cores" arm of bench.py's CPU baseline (`--impl reference`); the exact oracle stays the
  arbiter of parity.

Mapping: our activations are int8 with zero points in [-128, 127]; torch's are quint8.  x_u8 =
x + 128, zp_u8 = zp + 128 is the same real value.  The engine computes its requantisation
multiplier as s_in * s_w[c] / s_out: with s_in = s_out = 1 and s_w[c] = M[c] it is exactly the op's
fp32 multiplier M[c].  The fused ReLU6 clamp is applied to the engine's saturated output (the
clamp range lies inside [0, 255], so the order does not matter)."""
from __future__ import annotations

import numpy as np
import torch
import torch.ao.nn.quantized.functional as qF
import torch.nn.functional as F

OP_STEM, OP_PW, OP_DW, OP_ADD, OP_MAXPOOL = 1, 2, 3, 4, 5        # layer-program op codes
RS_UP, RS_DOWN = 1, 2


def _same_pad(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def qconv(op, x_s8, out_range=None):
    """One STEM / PW / DW op on the quantized engine.  x_s8: int64 / int8 [B,C,H,W] holding int8
    values (uint8 values for the stem, whose input zero point is 127 on the uint8 scale).
    Returns int64 [B,Cout,Ho,Wo] requantised to the conv's own int8 target (before any residual)."""
    q = op.q
    w = np.asarray(q['w'])
    mult = np.asarray(q['mult'], np.float32)
    if op.type == OP_STEM:
        xu = torch.as_tensor(np.asarray(x_s8), dtype=torch.uint8)
        zp_u = int(q['zp_in'][0])
        wt = torch.from_numpy(np.ascontiguousarray(w.transpose(0, 3, 1, 2)))
        k, stride, groups = 3, 2, 1
    else:
        xu = (torch.as_tensor(np.asarray(x_s8)).long() + 128).to(torch.uint8)
        zp_u = int(q['zp_in'][0]) + 128
        if op.type == OP_PW:
            wt = torch.from_numpy(np.ascontiguousarray(w[:, :, None, None]))
            k, stride, groups = 1, 1, 1
        else:
            wt = torch.from_numpy(np.ascontiguousarray(w[:, None]))
            k, stride, groups = op.k, op.stride, w.shape[0]
    if k > 1:      # TF SAME: pad with the zero point (real value 0)
        pt, pb = _same_pad(xu.shape[2], k, stride)
        pl, pr = _same_pad(xu.shape[3], k, stride)
        xu = F.pad(xu, (pl, pr, pt, pb), value=zp_u)
    xq = torch._make_per_tensor_quantized_tensor(xu.contiguous(), 1.0, zp_u)
    wq = torch._make_per_channel_quantized_tensor(wt.to(torch.int8), torch.from_numpy(mult.astype(np.float64)),
                                                  torch.zeros(len(mult), dtype=torch.int64), 0)
    bias = torch.from_numpy(np.asarray(q['bias'], np.float64) * mult.astype(np.float64)).float()
    zp_out = int(q['conv_zp_out'])
    y = qF.conv2d(xq, wq, bias, stride=stride, groups=groups, scale=1.0, zero_point=zp_out + 128, dtype=torch.quint8)
    y = y.int_repr().long() - 128
    lo, hi = out_range if out_range is not None else (q['act_lo'], q['act_hi'])
    return y.clamp(lo, hi)


def _maxpool(x):
    pt, pb = _same_pad(x.shape[2], 3, 2)
    pl, pr = _same_pad(x.shape[3], 3, 2)
    return F.max_pool2d(F.pad(x.double(), (pl, pr, pt, pb), value=-1e9), 3, 2).long()


def _add(xs, zps, mults, shift, zp_out, lo, hi):
    acc = torch.zeros_like(xs[0])
    for x, z, m in zip(xs, zps, mults):
        acc = acc + (x - z) * m
    return (((acc + (1 << (shift - 1))) >> shift) + zp_out).clamp(lo, hi)


def _conv_op(op, x, res):
    q = op.q
    if op.type == OP_PW and op.residual >= 0:
        y = qconv(op, x, out_range=(-128, 127))
        return _add([y, res], [q['conv_zp_out'], q['res_zp']], q['add_mult'], q['add_shift'], q['zp_out'],
                    q['act_lo'], q['act_hi'])
    return qconv(op, x)


def run(g, frames_u8):
    """The whole layer program with every convolution on the quantized engine.
    Returns (cls int8 [B,N], box int8 [B,N,4]) like oracle.effdet.run."""
    B = frames_u8.shape[0]
    vals = {g.input: torch.from_numpy(np.asarray(frames_u8)).long().permute(0, 3, 1, 2)}
    N = g.n_anchors
    cls, box = np.zeros((B, N), np.int8), np.zeros((B, N, 4), np.int8)
    with torch.no_grad():
        for op in g.ops:
            q = op.q
            ins = [vals[i] for i in op.inputs]
            if op.type in (OP_STEM, OP_PW, OP_DW):
                y = _conv_op(op, ins[0], vals[op.residual] if op.residual >= 0 else None)
            elif op.type == OP_MAXPOOL:
                y = _maxpool(ins[0])
            elif op.type == OP_ADD:
                t = g.tensors[op.out]
                xs = []
                for xin, rs in zip(ins, op.resample):
                    if rs == RS_UP:
                        iy = (torch.arange(t.h) * xin.shape[2]) // t.h
                        ix = (torch.arange(t.w) * xin.shape[3]) // t.w
                        xin = xin[:, :, iy][:, :, :, ix]
                    elif rs == RS_DOWN:
                        xin = _maxpool(xin)
                    xs.append(xin)
                y = _add(xs, q['zp_in'], q['add_mult'], q['add_shift'], q['zp_out'], q['act_lo'], q['act_hi'])
            else:
                raise ValueError(op.type)
            if op.out >= 0:
                vals[op.out] = y
            else:
                yv = y.permute(0, 2, 3, 1).numpy()
                n = yv.shape[1] * yv.shape[2] * 9
                if op.out_kind == 1:
                    cls[:, op.level_offset:op.level_offset + n] = q['lut'][(yv.reshape(B, n) + 128).astype(np.int64)]
                else:
                    box[:, op.level_offset:op.level_offset + n] = yv.reshape(B, n, 4).astype(np.int8)
    return cls, box


def check_convs(g, frames_u8, exact_tensors):
    """Per convolution op: run it on the quantized engine with the EXACT oracle's input tensor and
    compare with the exact oracle's output.  exact_tensors: the `keep=True` dict of
    oracle.effdet.run ({tensor id: int16 [B,h,w,c]}).  Returns [(op name, elements, mismatches,
    max |difference|)] for every conv op that writes a workspace tensor."""
    out = []
    frames = torch.from_numpy(np.asarray(frames_u8)).long().permute(0, 3, 1, 2)

    def tensor(i):
        if i == g.input:
            return frames
        return torch.from_numpy(exact_tensors[i].astype(np.int64)).permute(0, 3, 1, 2)

    with torch.no_grad():
        for op in g.ops:
            if op.type not in (OP_STEM, OP_PW, OP_DW) or op.out < 0:
                continue
            y = _conv_op(op, tensor(op.inputs[0]), tensor(op.residual) if op.residual >= 0 else None)
            want = tensor(op.out)
            d = (y - want).abs()
            out.append((op.name, int(d.numel()), int((d > 0).sum()), int(d.max())))
    return out


class QuantNet:
    """The layer program with prepacked weights on the quantized engine: what `run` computes, arranged
    for speed (weights packed once, activations kept as uint8 NCHW tensors between convolutions).
    bench.py's CPU arm times this as the host-core baseline: a production int8 conv engine (fbgemm /
    oneDNN, all host threads), the closest thing to the reference's TFLite/XNNPACK path that can run
    here.  Results follow the exact oracle to within the engine's one-step rounding differences; it is
    a speed baseline, never the parity arbiter."""

    def __init__(self, g):
        self.g = g
        self.packed = {}
        for i, op in enumerate(g.ops):
            if op.type not in (OP_STEM, OP_PW, OP_DW):
                continue
            q = op.q
            w = np.asarray(q['w'])
            mult = np.asarray(q['mult'], np.float32)
            if op.type == OP_STEM:
                wt, k, stride, groups = torch.from_numpy(np.ascontiguousarray(w.transpose(0, 3, 1, 2))), 3, 2, 1
            elif op.type == OP_PW:
                wt, k, stride, groups = torch.from_numpy(np.ascontiguousarray(w[:, :, None, None])), 1, 1, 1
            else:
                wt, k, stride, groups = torch.from_numpy(np.ascontiguousarray(w[:, None])), op.k, op.stride, w.shape[0]
            wq = torch._make_per_channel_quantized_tensor(wt.to(torch.int8), torch.from_numpy(mult.astype(np.float64)),
                                                          torch.zeros(len(mult), dtype=torch.int64), 0)
            bias = torch.from_numpy(np.asarray(q['bias'], np.float64) * mult.astype(np.float64)).float()
            pk = torch.ops.quantized.conv2d_prepack(wq, bias, [stride, stride], [0, 0], [1, 1], groups)
            self.packed[i] = (pk, k, stride)

    def _conv(self, i, op, xu, lo, hi):
        """xu: uint8 [B,C,H,W] (our int8 value + 128; the stem input as is).  Returns uint8."""
        pk, k, stride = self.packed[i]
        q = op.q
        zp_u = int(q['zp_in'][0]) + (0 if op.type == OP_STEM else 128)
        if k > 1:
            pt, pb = _same_pad(xu.shape[2], k, stride)
            pl, pr = _same_pad(xu.shape[3], k, stride)
            xu = F.pad(xu, (pl, pr, pt, pb), value=zp_u)
        xq = torch._make_per_tensor_quantized_tensor(xu.contiguous(memory_format=torch.channels_last), 1.0, zp_u)
        y = torch.ops.quantized.conv2d(xq, pk, 1.0, int(q['conv_zp_out']) + 128).int_repr()
        if lo > -128 or hi < 127:
            y = y.clamp(lo + 128, hi + 128)
        return y

    def run(self, frames_u8):
        g = self.g
        B = frames_u8.shape[0]
        vals = {g.input: torch.from_numpy(np.ascontiguousarray(frames_u8)).permute(0, 3, 1, 2)}
        N = g.n_anchors
        cls, box = np.zeros((B, N), np.int8), np.zeros((B, N, 4), np.int8)
        with torch.no_grad():
            for i, op in enumerate(g.ops):
                q = op.q
                ins = [vals[t] for t in op.inputs]
                if op.type in (OP_STEM, OP_PW, OP_DW):
                    if op.type == OP_PW and op.residual >= 0:
                        y = self._conv(i, op, ins[0], -128, 127).long() - 128
                        r = vals[op.residual].long() - 128
                        y = (_add([y, r], [q['conv_zp_out'], q['res_zp']], q['add_mult'], q['add_shift'], q['zp_out'],
                                  q['act_lo'], q['act_hi']) + 128).to(torch.uint8)
                    else:
                        y = self._conv(i, op, ins[0], q['act_lo'], q['act_hi'])
                elif op.type == OP_MAXPOOL:
                    y = (_maxpool(ins[0].long()) ).to(torch.uint8)
                elif op.type == OP_ADD:
                    t = g.tensors[op.out]
                    xs = []
                    for xin, rs in zip(ins, op.resample):
                        xin = xin.long() - 128
                        if rs == RS_UP:
                            iy = (torch.arange(t.h) * xin.shape[2]) // t.h
                            ix = (torch.arange(t.w) * xin.shape[3]) // t.w
                            xin = xin[:, :, iy][:, :, :, ix]
                        elif rs == RS_DOWN:
                            xin = _maxpool(xin)
                        xs.append(xin)
                    y = (_add(xs, q['zp_in'], q['add_mult'], q['add_shift'], q['zp_out'], q['act_lo'], q['act_hi']) + 128).to(torch.uint8)
                else:
                    raise ValueError(op.type)
                if op.out >= 0:
                    vals[op.out] = y
                else:
                    yv = (y.long() - 128).permute(0, 2, 3, 1).numpy()
                    n = yv.shape[1] * yv.shape[2] * 9
                    if op.out_kind == 1:
                        cls[:, op.level_offset:op.level_offset + n] = q['lut'][(yv.reshape(B, n) + 128).astype(np.int64)]
                    else:
                        box[:, op.level_offset:op.level_offset + n] = yv.reshape(B, n, 4).astype(np.int8)
        return cls, box
