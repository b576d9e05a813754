"""The arg-min of csrc/tracker.cu:assign_min_cost_warp, restated in numpy: candidate keys (doubles below +inf)
as order-preserving unsigned 64-bit integers with -0.0 folded onto +0.0, the minimum taken over the high words,
then over the low words of the survivors, ties to the lowest lane -- against the serial procedure's strict `<`
scan (oracle/ocsort.py:assign_min_cost), which keeps the FIRST column of the smallest value."""
import numpy as np


def serial_argmin(keys):
    """columns 1..M scanned in order with `if key < best` -> the first column holding the minimum; 0: none"""
    best, j1 = np.inf, 0
    for j, k in enumerate(keys, start=1):
        if k < best:
            best, j1 = k, j
    return j1


def warp_argmin(keys):
    k = np.asarray(keys, dtype=np.float64) + 0.0                         # key + 0.0: -0.0 -> +0.0
    cand = k < np.inf
    bits = k.view(np.uint64)
    sign = (bits >> np.uint64(63)).astype(bool)
    ko = np.where(sign, ~bits, bits | np.uint64(1 << 63))
    ko = np.where(cand, ko, np.uint64(0xFFFFFFFFFFFFFFFF))
    hi, lo = (ko >> np.uint64(32)).astype(np.uint32), (ko & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    min_hi = hi.min()                                                     # REDUX.MIN over the warp
    min_lo = np.where(hi == min_hi, lo, np.uint32(0xFFFFFFFF)).min()      # second REDUX.MIN
    win = cand & (hi == min_hi) & (lo == min_lo)                          # ballot
    return int(np.argmax(win)) + 1 if win.any() else 0                    # ffs: the lowest lane


def test_warp_argmin_equals_the_serial_scan():
    rng = np.random.default_rng(7)
    specials = np.array([0.0, -0.0, np.inf, 1e-300, -1e-300, 5e-324, -5e-324, 1.0, -1.0, 1e300, -1e300])
    for trial in range(4000):
        m = int(rng.integers(1, 32))
        keys = rng.normal(size=m) * 10.0 ** rng.integers(-6, 6)
        pick = rng.random(m) < 0.4
        keys[pick] = rng.choice(specials, size=int(pick.sum()))
        if trial % 3 == 0:                                                # exact ties between several columns
            keys[rng.integers(0, m, size=max(1, m // 2))] = keys[rng.integers(0, m)]
        assert warp_argmin(keys) == serial_argmin(keys), keys
    assert warp_argmin([np.inf, np.inf]) == 0 and serial_argmin([np.inf, np.inf]) == 0
    assert warp_argmin([0.0, -0.0]) == 1 and warp_argmin([-0.0, 0.0]) == 1   # equal as doubles: the first column
