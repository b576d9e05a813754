"""The packed requantisation of csrc/requant.cuh (Requant::pack4t<true>), restated in numpy float32
and checked against the plain definition  y = clamp(rne(float32(acc) * M) + zp, lo, hi)  -- the
arithmetic argument of the kernel comment, executed: magic-number rounding with the 2^15 bias,
unsigned 16-bit clamp, zero point added as an integer afterwards, low byte taken."""
import numpy as np
import pytest

MAGIC_BITS = 0x4B400000          # bits of 1.5 * 2^23


def plain(acc, mult, zp, lo, hi):
    p = acc.astype(np.float32) * mult.astype(np.float32)
    return np.clip(np.rint(p).astype(np.int64) + zp, lo, hi)


def packed(acc, mult, zp, lo, hi):
    p = acc.astype(np.float32) * mult.astype(np.float32)                 # FMUL (scalar: no FFMA2 contraction)
    q = (p + np.float32(12582912.0 + 32768.0)).astype(np.float32)        # FADD2: rounds p like rint()
    half = q.view(np.uint32) & 0xFFFF                                    # PRMT: low 16 bits = rint(p) + 2^15
    lo16, hi16 = lo - zp + 32768, hi - zp + 32768
    half = np.minimum(np.maximum(half, lo16), hi16)                      # VIMNMX.U16x2
    half = (half + (zp & 0xFFFF)) & 0xFFFF                               # 32-bit add of zp * 0x10001: no carry between halves
    return ((half & 0xFF).astype(np.uint8)).view(np.int8).astype(np.int64)   # PRMT 0x6420: the low byte is the int8


@pytest.mark.parametrize('zp,lo,hi', [(11, -128, 127), (-128, -128, 95), (-20, -20, 107), (127, 0, 127), (-3, -128, 127)])
def test_packed_equals_plain(zp, lo, hi):
    rng = np.random.default_rng(zp + 200)
    mult = rng.uniform(1e-5, 2e-2, 4096).astype(np.float32)
    acc = np.rint(rng.uniform(-1, 1, 4096) * 31000.0 / mult).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)
    assert np.abs(acc.astype(np.float32) * mult).max() < 32000             # the bound pack_blob proves per op
    assert np.array_equal(packed(acc, mult, zp, lo, hi), plain(acc, mult, zp, lo, hi))


@pytest.mark.parametrize('zp', [11, -20, -127, 0, 126])
def test_ties_go_to_even_before_the_zero_point(zp):
    """acc * M exactly on .5: rint() goes to the even integer of p, whatever the parity of zp --
    which is why the zero point is NOT folded into the magic constant."""
    acc = np.arange(-4001, 4001, 2, dtype=np.int32)                          # odd accumulators
    for m in (0.5, 0.25, 0.125):
        mult = np.full(acc.shape, m, np.float32)
        a = acc * int(1 / m) // 2 * 2 + 1 if m != 0.5 else acc
        a = a.astype(np.int32)
        assert np.array_equal(packed(a, mult, zp, -128, 127), plain(a, mult, zp, -128, 127))
    # folding zp into the float add would break odd zero points on ties
    p = np.float32(0.5)
    folded = int((np.float32(p + np.float32(12582912.0 + 1))).view(np.uint32)) - MAGIC_BITS   # zp = 1
    assert folded == 2 and int(np.rint(p)) + 1 == 1


def test_synthetic_lite1_head_fires_like_a_trained_detector():
    """effdet.calibrate_class_prior: the random Lite1 / Lite2 class heads are shifted so that a handful of
    anchors per frame clear score 0.5 (a trained single-class detector on these clips shows 1-3 plates),
    not hundreds; Lite0 is left as initialised."""
    import numpy as np
    from oracle import effdet as OE
    from vbt_b200 import effdet as E
    from vbt_b200.synth import synthetic_model_inputs
    g = E.build_synthetic('lite1')
    x = synthetic_model_inputs(2, g.S, seed=77)
    cls, _, _ = OE.run(g, x)
    above = (cls.astype(np.int32) + 128 >= 128).sum(axis=1)          # post-LOGISTIC score >= 0.5
    assert np.all(above >= 1) and np.all(above <= 60), above
    g0 = E.Graph('lite0')
    E.init_weights(g0, 1234)
    b_before = [op.bias.copy() for op in g0.ops if op.out_kind == 1]
    g0b = E.build_synthetic('lite0')
    assert all(np.array_equal(a, op.bias) for a, op in zip(b_before, [o for o in g0b.ops if o.out_kind == 1]))


def add_min_relu(acc, mult, zp, lo, hi):
    """The fused MBConv expand epilogue (csrc/mbconv_umma.cu: ee_tiles): plain magic add, then ONE add-min-relu on
    the float's bits; the result plus (lo + 128) is the UNSIGNED byte the expanded planes hold (value + 128)."""
    p = acc.astype(np.float32) * mult.astype(np.float32)
    bits = (p + np.float32(12582912.0)).astype(np.float32).view(np.int32).astype(np.int64)   # 0x4B400000 + rint(p)
    off = -(MAGIC_BITS + (lo - zp))
    t = np.maximum(np.minimum(bits + off, hi - lo), 0)            # VIADDMNMX.RELU
    return t + (lo + 128)


@pytest.mark.parametrize('zp,lo,hi', [(-128, -128, 95), (11, -128, 127), (-20, -20, 107), (127, 0, 127)])
def test_add_min_relu_is_the_plain_value_plus_128(zp, lo, hi):
    rng = np.random.default_rng(zp + 300)
    mult = rng.uniform(1e-5, 2e-2, 4096).astype(np.float32)
    acc = np.rint(rng.uniform(-1, 1, 4096) * 31000.0 / mult).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)
    acc[:64] = np.rint((np.arange(64) - 32 + 0.5) / mult[:64]).astype(np.int32)        # near .5 ties
    assert np.array_equal(add_min_relu(acc, mult, zp, lo, hi), plain(acc, mult, zp, lo, hi) + 128)


def test_mbconv_image_depthwise_bias_undoes_the_unsigned_storage():
    """effdet.mbconv_images: the depthwise bias of a chunk image carries -128 * sum(w), so that the kernel's
    dp4a.u32.s32 over (value + 128) bytes equals the signed sum."""
    from vbt_b200 import effdet
    rng = np.random.default_rng(5)
    cexp, k, cin_p, cout_p = 48, 5, 16, 32
    e = dict(w=rng.integers(-127, 128, (cexp, cin_p), dtype=np.int8), bias=rng.integers(-5000, 5000, cexp).astype(np.int32),
             mult=rng.uniform(1e-4, 1e-2, cexp).astype(np.float32))
    d = dict(w=rng.integers(-127, 128, (cexp, k, k), dtype=np.int8), bias=rng.integers(-5000, 5000, cexp).astype(np.int32),
             mult=rng.uniform(1e-4, 1e-2, cexp).astype(np.float32))
    p = dict(w=rng.integers(-127, 128, (24, cexp), dtype=np.int8), bias=np.zeros(24, np.int32), mult=np.ones(24, np.float32))
    img, stride, n_chunks = effdet.mbconv_images(cin_p, k, cout_p, e, d, p)
    _, _, off_consts, stride2 = effdet.mbconv_image_layout(cin_p, k, cout_p, True)
    assert stride == stride2 and n_chunks == 2
    x = rng.integers(-128, 128, (cexp, k, k)).astype(np.int64)                # one window per channel
    want = (d['w'].astype(np.int64) * x).sum(axis=(1, 2)) + d['bias']
    for c in range(n_chunks):
        consts = img[c, off_consts:off_consts + 512].view(np.int32)
        lo, hi = 32 * c, min(32 * c + 32, cexp)
        got = (d['w'][lo:hi].astype(np.int64) * (x[lo:hi] + 128)).sum(axis=(1, 2)) + consts[64:64 + hi - lo]
        assert np.array_equal(got, want[lo:hi])
        assert np.array_equal(consts[:hi - lo], e['bias'][lo:hi])              # the expand bias: pre-stored in TMEM as it is
