"""CPU oracle for the detector network (row a5n of SURVEY.md section 8) -- TEST INFRASTRUCTURE.

Interprets a quantised `vbt_b200.effdet.Graph` (logical, un-padded tensors and weights)
with the int8 semantics of the exported EfficientDet-Lite graphs: int8 activations
(per-tensor scale / zero point), int8 per-channel weights, int32 accumulate and bias,
requantise, fused ReLU6 clamp, quantised ADD, raw-int8 max-pool / nearest resize, int8
LOGISTIC.  The arithmetic the reference runs lives in tflite-runtime==2.14.0
(requirements.txt:381) and in `.tflite` blobs that are ABSENT from the reference checkout
(.MISSING_LARGE_BLOBS): PARITY UNPINNED [3P-MEM] -- this restates the published TFLite
quantisation spec (XNNPACK-style fp32 requantisation, integer ADD); what IS pinned is the
architecture's MAC count against models/*.log:110 (tests/test_oracle_effdet.py).

Integer accumulations are evaluated with floating-point convolutions on the CPU -- float32
where every partial sum provably stays below 2^24, float64 otherwise -- so they are exact
integers, computed independently of the CUDA kernels' dp4a / tensor-core paths.  Results must match the GPU bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# Nothing is imported from the product: the graph to interpret is passed in as data (ops with their
# quantised weights), and the op codes / padding rules below are this file's own statement of them.
# That the product's graph IS EfficientDet-Lite is checked separately against oracle/arch.py
# (tests/test_oracle_arch.py); that the conv + requantisation arithmetic below agrees with an
# unrelated production int8 engine is checked by oracle/effdet_q.py.
OP_STEM, OP_PW, OP_DW, OP_ADD, OP_MAXPOOL = 1, 2, 3, 4, 5
RS_NONE, RS_UP, RS_DOWN = 0, 1, 2
ANCHORS_PER_LOCATION = 9            # 3 octaves x 3 aspect ratios


def same_pad(size, k, stride):
    """TF SAME: (output size, pad before)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


def nearest_index(dst, n_in, n_out):
    """Legacy nearest-neighbour resize index (no half-pixel centres): floor(dst * in / out)."""
    return (np.arange(dst) * n_in) // n_out


FORCE_F64 = False      # tests flip this to check the float32 fast path against float64


def _requant(acc, mult, zp_out, lo, hi):
    """acc int64 [B,C,H,W]; mult float32 [C] -> int64 clamp(rne(f32(acc)*M)+zp)."""
    a = acc.numpy().astype(np.float32)                     # int -> f32, round to nearest even
    y = a * mult.astype(np.float32)[None, :, None, None]   # single fp32 multiply
    y = np.rint(y).astype(np.int64) + zp_out
    return torch.from_numpy(np.clip(y, lo, hi))


def _same_pad(x, k, stride, value):
    ho, pt = same_pad(x.shape[2], k, stride)
    wo, pl = same_pad(x.shape[3], k, stride)
    pb = max((ho - 1) * stride + k - x.shape[2] - pt, 0)
    pr = max((wo - 1) * stride + k - x.shape[3] - pl, 0)
    return F.pad(x, (pl, pr, pt, pb), value=value)


def _maxpool(x):
    return F.max_pool2d(_same_pad(x.double(), 3, 2, -1e9), 3, 2).long()


def _add(xs, zps, mults, shift, zp_out, lo, hi):
    acc = torch.zeros_like(xs[0])
    for x, z, m in zip(xs, zps, mults):
        acc = acc + (x - z) * m
    y = ((acc + (1 << (shift - 1))) >> shift) + zp_out      # arithmetic shift (floor)
    return y.clamp(lo, hi)


def run(g, frames_u8, keep=False):
    """frames_u8: uint8 [B,S,S,3].  Returns (cls int8 [B,N], box int8 [B,N,4], tensors)
    where cls is the post-LOGISTIC score (scale 1/256, zp -128)."""
    B = frames_u8.shape[0]
    vals = {g.input: torch.from_numpy(np.asarray(frames_u8)).long().permute(0, 3, 1, 2)}
    N = g.n_anchors
    cls = np.zeros((B, N), np.int8)
    box = np.zeros((B, N, 4), np.int8)
    a_per = ANCHORS_PER_LOCATION
    with torch.no_grad():
        for op in g.ops:
            q = op.q
            ins = [vals[i] for i in op.inputs]
            if op.type in (OP_STEM, OP_PW, OP_DW):
                # float32 convolutions are exact while every partial sum stays below 2^24
                # (|x - zp| <= 255, |w| <= 127): true for every layer with fan-in <= 518;
                # wider ones (the 672 / 1152-channel projections) use float64
                fan_in = int(np.prod(q['w'].shape[1:]))
                ft, nt = (torch.float32, np.float32) if fan_in * 255 * 127 < (1 << 24) and not FORCE_F64 \
                    else (torch.float64, np.float64)
                x = (ins[0] - q['zp_in'][0]).to(ft)
                w = torch.from_numpy(q['w'].astype(nt))
                if op.type == OP_STEM:
                    acc = F.conv2d(_same_pad(x, 3, 2, 0.0), w.permute(0, 3, 1, 2).contiguous(),
                                   stride=2)
                elif op.type == OP_PW:
                    acc = F.conv2d(x, w[:, :, None, None])
                else:
                    acc = F.conv2d(_same_pad(x, op.k, op.stride, 0.0), w[:, None],
                                   stride=op.stride, groups=w.shape[0])
                acc = acc.round().long() + torch.from_numpy(q['bias'].astype(np.int64))[None, :, None, None]
                if op.type == OP_PW and op.residual >= 0:
                    y = _requant(acc, q['mult'], q['conv_zp_out'], -128, 127)
                    y = _add([y, vals[op.residual]], [q['conv_zp_out'], q['res_zp']],
                             q['add_mult'], q['add_shift'], q['zp_out'], q['act_lo'], q['act_hi'])
                else:
                    y = _requant(acc, q['mult'], q['conv_zp_out'], q['act_lo'], q['act_hi'])
            elif op.type == OP_MAXPOOL:
                y = _maxpool(ins[0])
            elif op.type == OP_ADD:
                t = g.tensors[op.out]
                xs = []
                for xin, rs in zip(ins, op.resample):
                    if rs == RS_UP:
                        iy = torch.from_numpy(nearest_index(t.h, xin.shape[2], t.h))
                        ix = torch.from_numpy(nearest_index(t.w, xin.shape[3], t.w))
                        xin = xin[:, :, iy][:, :, :, ix]
                    elif rs == RS_DOWN:
                        xin = _maxpool(xin)
                    xs.append(xin)
                y = _add(xs, q['zp_in'], q['add_mult'], q['add_shift'], q['zp_out'],
                         q['act_lo'], q['act_hi'])
            else:
                raise ValueError(op.type)
            if op.out >= 0:
                vals[op.out] = y
            else:
                yv = y.permute(0, 2, 3, 1).numpy()              # [B,h,w,c]
                h, w = yv.shape[1:3]
                n = h * w * a_per
                if op.out_kind == 1:
                    lut = q['lut']
                    flat = lut[(yv.reshape(B, n) + 128).astype(np.int64)]
                    cls[:, op.level_offset:op.level_offset + n] = flat
                else:
                    box[:, op.level_offset:op.level_offset + n] = yv.reshape(B, n, 4).astype(np.int8)
    tensors = {k: v.permute(0, 2, 3, 1).numpy().astype(np.int16) for k, v in vals.items()} if keep else None
    return cls, box, tensors
