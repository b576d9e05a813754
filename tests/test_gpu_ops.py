"""Single-kernel parity on a B200: one-op layer programs (tests/micrograph.py) through
vbt_detect vs oracle/effdet.py, bit-exact int8, at shapes the full networks do not visit
(K / N / M tails of the tcgen05 pointwise GEMM, multi-chunk K and N, residual epilogue;
depthwise 3x3 / 5x5 at stride 1 / 2 with odd sizes)."""
import numpy as np
import pytest

from oracle import effdet as OE
import micrograph as MG

pytestmark = pytest.mark.gpu


def check(g, B, seed=1):
    x, xp = MG.random_input(g, B, seed)
    _, _, want = OE.run(g, x, keep=True)
    got = MG.run_gpu(g, xp)
    for op in g.ops:
        if op.out not in got:
            continue                               # inside a fused run
        t = g.tensors[op.out]
        vals, pad = got[op.out]
        assert np.array_equal(vals.astype(np.int16), want[op.out]), f'{op.name}: values differ'
        assert np.all(pad == t.zp), f'{op.name}: pad channels must hold the zero point'


@pytest.mark.parametrize('h,w,cin,cout,B', [
    (8, 16, 16, 16, 1),          # one exact 128-row tile, smallest K and N
    (5, 7, 16, 96, 3),           # M tail (105 rows), b2.0.expand channel shape
    (9, 9, 24, 144, 2),          # cin pad 24->32
    (9, 9, 40, 240, 2),          # cin_p 48: odd number of 16-byte K chunks
    (6, 5, 80, 480, 5),          # N split in two chunks of 240
    (4, 4, 112, 672, 9),         # N = 3 x 224, K = 7 chunks
    (3, 3, 192, 1152, 15),       # N = 5 chunks, M tail
    (3, 3, 1152, 192, 15),       # K = 72 chunks -> 5 K iterations
    (5, 5, 672, 112, 6),         # K = 42 chunks (last iteration 10)
    (10, 10, 320, 64, 2),        # BiFPN lateral
    (7, 3, 272, 48, 7),          # K just over one stage (17 chunks)
    (40, 40, 64, 64, 3),         # many tiles
])
@pytest.mark.parametrize('act', [False, True])
def test_pointwise(h, w, cin, cout, B, act):
    check(MG.pw_graph(h, w, cin, cout, act=act, seed=h * 1000 + cin), B)


@pytest.mark.parametrize('h,w,c,cmid,B', [(8, 16, 16, 96, 1), (5, 7, 24, 144, 3), (10, 10, 112, 672, 2),
                                          (4, 4, 192, 1152, 9), (20, 20, 40, 240, 2)])
def test_pointwise_residual(h, w, c, cmid, B):
    check(MG.pw_graph(h, w, c, cmid, act=True, residual=True, seed=c), B)


@pytest.mark.parametrize('h,w,c,k,stride,B', [
    (16, 16, 32, 3, 1, 2), (17, 13, 96, 3, 2, 3), (16, 16, 144, 5, 2, 2), (9, 11, 240, 5, 1, 2),
    (5, 5, 64, 3, 1, 3), (3, 3, 64, 3, 1, 5), (40, 40, 40, 5, 1, 1), (7, 7, 1152, 3, 1, 2),
    (33, 31, 16, 3, 2, 1), (10, 10, 672, 5, 2, 2),
])
@pytest.mark.parametrize('act', [False, True])
def test_depthwise(h, w, c, k, stride, B, act):
    check(MG.dw_graph(h, w, c, k, stride, act=act, seed=h * 100 + c), B)


@pytest.mark.parametrize('impl', ['simt', 'umma'])
def test_depthwise_both_implementations(impl):
    """The two depthwise implementations -- tensor-pipe implicit GEMM with block-diagonal
    weights (dw_umma.cu) and channel-word-stationary dp4a (net.cu) -- forced in turn for every
    kernel size / stride against the oracle.  VBT_DW_IMPL is read once per process, hence the
    subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ('import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_ops as T; '
            'import micrograph as MG\n'
            'for (h, w, c, k, s, B) in [(17, 13, 96, 3, 2, 3), (16, 16, 144, 5, 2, 2), (9, 11, 240, 5, 1, 2), '
            '(40, 40, 40, 3, 1, 1), (33, 31, 16, 3, 2, 1), (7, 7, 1152, 3, 1, 2), (160, 160, 32, 3, 1, 1)]:\n'
            '    T.check(MG.dw_graph(h, w, c, k, s, act=True, seed=h), B)\n' % (os.path.dirname(here), here))
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, VBT_DW_IMPL=impl),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


# ---- fused [ADD ->] DW3x3 -> PW kernel (csrc/node_umma.cu) ---------------------------------

@pytest.mark.parametrize('h,w,c,cout,B', [
    (3, 3, 64, 64, 5), (5, 5, 64, 64, 3), (10, 10, 64, 64, 2), (20, 20, 64, 64, 2), (40, 40, 64, 64, 2),
    (6, 6, 88, 88, 3), (12, 12, 88, 88, 2), (24, 24, 88, 88, 2), (48, 48, 88, 88, 1),
    (4, 4, 112, 112, 3), (7, 7, 112, 112, 2), (28, 28, 112, 112, 1), (56, 56, 112, 112, 1),
    (9, 13, 16, 16, 2), (11, 6, 128, 128, 2), (17, 19, 40, 24, 1),
])
def test_fused_head_stage(h, w, c, cout, B):
    g = MG.head_graph(h, w, c, cout, act=True, seed=h * 100 + c)
    check(g, B)
    assert list(MG.run_gpu.last_plan) == [2, 0], 'the DW -> PW pair must run as one kernel'


@pytest.mark.parametrize('h,w,c,kind', [(5, 5, 64, 1), (20, 20, 64, 2), (12, 12, 88, 1), (24, 24, 88, 2),
                                        (48, 48, 88, 2), (7, 7, 112, 2), (56, 56, 112, 1)])
def test_fused_head_output(h, w, c, kind):
    """Packed 9 / 36-channel raw outputs (+ LOGISTIC LUT on the class head) from the fused kernel."""
    B = 2
    g = MG.head_graph(h, w, c, 9 if kind == 1 else 36, act=False, seed=h + c, out_kind=kind)
    x, xp = MG.random_input(g, B, 3)
    cls, box, _ = OE.run(g, x)
    got = MG.run_gpu(g, xp)
    assert list(MG.run_gpu.last_plan) == [2, 0]
    if kind == 1:
        assert np.array_equal(got[-1], cls)
    else:
        assert np.array_equal(got[-2], box)


@pytest.mark.parametrize('h,w,c,n_in,odd,B', [
    (3, 3, 64, 3, True, 4), (5, 5, 64, 3, False, 3), (10, 10, 64, 3, False, 2), (20, 20, 64, 3, False, 2),
    (40, 40, 64, 2, False, 1), (6, 6, 88, 3, False, 2), (12, 12, 88, 3, False, 2), (24, 24, 88, 3, False, 2),
    (48, 48, 88, 2, False, 1), (4, 4, 112, 3, True, 2), (14, 14, 112, 3, False, 2), (28, 28, 112, 3, False, 1),
    (56, 56, 112, 2, False, 1), (9, 7, 32, 3, True, 2),
])
def test_fused_bifpn_node(h, w, c, n_in, odd, B):
    g = MG.node_graph(h, w, c, n_in=n_in, seed=h * 10 + c, odd=odd)
    check(g, B)
    assert list(MG.run_gpu.last_plan)[-3:] == [3, 0, 0], 'ADD -> DW -> PW must run as one kernel'


# ---- requantisation: packed 16-bit path vs the plain one, exact ties --------------------------

@pytest.mark.parametrize('mult,fast', [(0.5, 1), (0.25, 1), (0.125, 1), (8.0, 0)])
@pytest.mark.parametrize('zp_out', [11, -20, -128])
def test_requant_ties_and_both_paths(mult, fast, zp_out):
    """Power-of-two multipliers put acc * M exactly on .5 for every odd accumulator: the result
    must go to the even integer BEFORE the zero point is added (odd and even zero points), on
    the packed path (requant_fast = 1) and on the plain one (bound too large: requant_fast = 0)."""
    from vbt_b200 import effdet as E
    for g in (MG.pw_graph(9, 9, 16, 32, act=True, seed=5, zp_out=zp_out),
              MG.dw_graph(9, 11, 32, 3, 1, act=True, seed=6, zp_out=zp_out),
              MG.dw_graph(9, 11, 32, 5, 1, act=False, seed=7, zp_out=zp_out),
              MG.head_graph(7, 7, 64, 64, act=True, seed=8)):
        for op in g.ops:
            n = op.q['mult'].shape[0]
            op.q['mult'] = np.full(n, mult, np.float32)
            op.q['w'] = np.clip(op.q['w'], -3, 3)               # small accumulators: many land in range
            op.q['bias'] = (op.q['bias'] % 7).astype(np.int32)
        E.pack_blob(g)
        check(g, 2)


# ---- persistent warp-specialised pointwise kernel (csrc/pw_persist.cu) -------------------------

def test_pointwise_persistent_kernel():
    """pw_persist.cu takes the large-M, K <= 256 layers; VBT_PW_PERSIST_MIN_TILES=1 forces it for
    every eligible shape here: M tails, one tile, many tiles per CTA (ring and accumulator phases
    wrap several times), odd K chunk counts, two N chunks, residual epilogue.  The env var is read
    once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ('import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_ops as T; '
            'import micrograph as MG\n'
            'for (h, w, cin, cout, B, act) in [(8, 16, 16, 16, 1, False), (5, 7, 16, 96, 3, True), (9, 9, 24, 144, 2, True), '
            '(9, 9, 40, 240, 2, False), (6, 5, 80, 480, 5, True), (40, 40, 240, 40, 3, False), (160, 160, 16, 96, 6, True), '
            '(80, 80, 144, 24, 7, False), (96, 96, 32, 16, 5, False), (31, 33, 256, 256, 9, True), (64, 64, 112, 64, 9, False)]:\n'
            '    T.check(MG.pw_graph(h, w, cin, cout, act=act, seed=h + cin), B)\n'
            'for (h, w, c, cmid, B) in [(8, 16, 16, 96, 1), (5, 7, 24, 144, 3), (20, 20, 40, 240, 2), (80, 80, 24, 144, 6)]:\n'
            '    T.check(MG.pw_graph(h, w, c, cmid, act=True, residual=True, seed=c), B)\n'
            % (os.path.dirname(here), here))
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, VBT_PW_PERSIST='1', VBT_PW_PERSIST_MIN_TILES='1'),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


# ---- bf16 head contraction (BASELINE configs[3]): same integers on the kind::f16 tensor path ----

@pytest.mark.parametrize('h,w,c,cout,kind', [
    (3, 3, 64, 64, 0), (20, 20, 64, 64, 0), (40, 40, 64, 36, 2), (12, 12, 88, 88, 0), (48, 48, 88, 9, 1),
    (7, 7, 112, 112, 0), (28, 28, 112, 112, 0), (56, 56, 112, 36, 2), (56, 56, 112, 9, 1), (9, 13, 16, 16, 0),
    (11, 6, 128, 128, 0),
])
def test_fused_head_stage_bf16(h, w, c, cout, kind):
    """DW3x3 -> PW with the pointwise stage as tcgen05.mma.kind::f16 (bf16 operands, fp32
    accumulators): bit-identical to the int8 oracle, because every operand, product and partial
    sum is an integer below 2^24."""
    B = 2
    g = MG.head_graph(h, w, c, cout, act=(kind == 0), seed=h * 100 + c, out_kind=kind)
    g.head_dtype = 'bf16'
    x, xp = MG.random_input(g, B, 4)
    # extremes: -128 activations against +-127 weights everywhere in one frame
    x[0], xp[0, ..., :c] = -128, -128
    cls, box, want = OE.run(g, x, keep=True)
    got = MG.run_gpu(g, xp)
    assert list(MG.run_gpu.last_plan) == [2, 0]
    if kind == 1:
        assert np.array_equal(got[-1], cls)
    elif kind == 2:
        assert np.array_equal(got[-2], box)
    else:
        assert np.array_equal(got[g.ops[-1].out][0].astype(np.int16), want[g.ops[-1].out])


@pytest.mark.parametrize('h,w,c,odd,B', [(3, 3, 64, True, 4), (5, 5, 64, False, 3), (20, 20, 64, False, 2),
                                          (40, 40, 64, False, 1), (12, 12, 88, False, 2), (7, 7, 112, True, 2),
                                          (28, 28, 112, False, 1)])
def test_fused_bifpn_node_with_add_tree(h, w, c, odd, B):
    """ADD(ADD(a, b), c) -> DW -> PW: four ops, one kernel; the inner sum is requantised to its own
    int8 tensor (and clamped) before the outer sum uses it."""
    g = MG.node_graph(h, w, c, n_in=3, seed=h * 10 + c + 1, odd=odd, tree=True)
    check(g, B)
    assert list(MG.run_gpu.last_plan)[-4:] == [4, 0, 0, 0]


# ---- fused MBConv block (csrc/mbconv_umma.cu) -----------------------------------------------------

# every block shape of EfficientDet-Lite0 / 1 / 2 (channels, kernel, stride; maps shrunk where the
# full size adds nothing but time) plus odd sizes, tile tails and two-tile cases
MBCONV_SHAPES = [
    # h, w, cin, cexp, cout, k, s, residual, B
    (24, 24, 16, 96, 24, 3, 2, False, 2),       # b2.0
    (160, 160, 16, 96, 24, 3, 2, False, 1),     # b2.0 at full size: 2-D tiles, right / bottom tails
    (20, 20, 24, 144, 24, 3, 1, True, 2),       # b2.1
    (80, 80, 24, 144, 24, 3, 1, True, 1),
    (22, 26, 24, 144, 40, 5, 2, False, 2),      # b3.0
    (80, 80, 24, 144, 40, 5, 2, False, 1),
    (40, 40, 40, 240, 40, 5, 1, True, 2),       # b3.1
    (40, 40, 40, 240, 80, 3, 2, False, 2),      # b4.0
    (20, 20, 80, 480, 80, 3, 1, True, 2),       # b4.1
    (20, 20, 80, 480, 112, 5, 1, False, 2),     # b5.0
    (20, 20, 112, 672, 112, 5, 1, True, 2),     # b5.1
    (20, 20, 112, 672, 192, 5, 2, False, 2),    # b6.0
    (10, 10, 192, 1152, 192, 5, 1, True, 3),    # b6.1
    (10, 10, 192, 1152, 320, 3, 1, False, 3),   # b7.0: Cout 320 = two N halves
    (24, 24, 48, 288, 48, 5, 1, True, 1),       # Lite2 widths
    (14, 14, 120, 720, 120, 5, 1, True, 2),
    (14, 14, 120, 720, 208, 5, 2, False, 2),
    (7, 7, 208, 1248, 208, 5, 1, True, 2),
    (7, 7, 208, 1248, 352, 3, 1, False, 2),     # Lite2 b7.0: Cout 352
    (12, 12, 88, 528, 88, 3, 1, True, 2),
    (17, 13, 24, 144, 24, 3, 1, True, 3),       # odd sizes
    (17, 13, 24, 144, 40, 5, 2, False, 3),
    (33, 31, 16, 96, 24, 3, 2, False, 1),
    (9, 11, 40, 240, 40, 5, 1, True, 2),
    (5, 5, 80, 480, 80, 3, 1, True, 4),
    (3, 3, 192, 1152, 192, 5, 1, True, 5),
    (1, 1, 16, 96, 16, 3, 1, True, 2),
]


@pytest.mark.parametrize('h,w,cin,cexp,cout,k,s,res,B', MBCONV_SHAPES)
def test_fused_mbconv_block(h, w, cin, cexp, cout, k, s, res, B):
    g = MG.mbconv_graph(h, w, cin, cexp, cout, k, s, residual=res, seed=h * 100 + cin + k)
    check(g, B)
    assert list(MG.run_gpu.last_plan) == [3, 0, 0], 'expand -> depthwise -> project must run as one kernel'


@pytest.mark.parametrize('h,w,c,cout,k,s,B', [(160, 160, 32, 16, 3, 1, 1), (96, 70, 32, 16, 3, 1, 2), (80, 80, 16, 24, 5, 2, 1),
                                              (67, 91, 32, 32, 3, 2, 2)])
def test_fused_mbconv_without_expand(h, w, c, cout, k, s, B):
    """The backbone's first block (expand ratio 1): depthwise -> project on a map too large for the
    fused node kernel."""
    g = MG.mbconv_graph(h, w, c, c, cout, k, s, seed=h + c, expand=False)
    check(g, B)
    assert list(MG.run_gpu.last_plan) == [2, 0]


@pytest.mark.parametrize('h,w,cout,k,s,B', [(64, 64, 16, 3, 1, 2), (97, 71, 16, 3, 1, 3), (40, 56, 24, 5, 2, 2), (320, 320, 16, 3, 1, 1),
                                            (33, 33, 16, 5, 1, 2)])
def test_fused_stem_block(h, w, cout, k, s, B):
    """The network's first block with the STEM as its expand stage: uint8 frame -> 3x3 s2 conv (a K = 27 GEMM
    over im2col rows, unsigned A operand, zero-point padding at the frame border) -> depthwise -> project, one
    kernel.  Even / odd frame sizes move the SAME padding between the two borders."""
    import torch
    from vbt_b200.interpreter import Detector
    g = MG.stem_block_graph(h, w, cout, k, s, seed=h + k)
    rng = np.random.default_rng(h)
    frames = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    frames[0, :2] = 255
    frames[0, -2:] = 0
    _, _, want = OE.run(g, frames, keep=True)
    det = Detector(g, max_batch=B)
    assert list(det.plan()) == [3, 0, 0] and list(det.plan_kinds()) == [1, 0, 0]
    det.network(torch.as_tensor(frames, device='cuda'))
    torch.cuda.synchronize()
    t = g.tensors[g.ops[-1].out]
    ws = det.workspace.cpu().numpy().view(np.int8)
    got = ws[B * t.ws_offset:B * t.ws_offset + B * t.h * t.w * t.c_p].reshape(B, t.h, t.w, t.c_p)
    assert np.array_equal(got[..., :t.c].astype(np.int16), want[g.ops[-1].out])
    assert np.all(got[..., t.c:] == t.zp)


def test_mbconv_unfused_path_still_matches():
    """VBT_MBCONV=0 runs the same blocks op by op (the kernels the fused one replaced)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ('import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_ops as T; '
            'import micrograph as MG\n'
            'for (h, w, cin, cexp, cout, k, s, res, B) in T.MBCONV_SHAPES[::4]:\n'
            '    T.check(MG.mbconv_graph(h, w, cin, cexp, cout, k, s, residual=res, seed=h), B)\n'
            '    assert list(MG.run_gpu.last_plan) == [1, 1, 1]\n' % (os.path.dirname(here), here))
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, VBT_MBCONV='0'), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
