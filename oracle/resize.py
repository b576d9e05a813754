"""CPU oracle for frame preprocessing (rows a2-a3 of SURVEY.md section 8) -- TEST INFRASTRUCTURE.

Restates cv2.cvtColor(BGR2RGB) (track.py:171) + odt.preprocess_image (odt.py:10-19):
`tf.image.resize(img, (S, S))` (TF 2.8.4 ResizeBilinear, half_pixel_centers=True,
antialias=False, aspect ratio not preserved) followed by `tf.cast(..., tf.uint8)`.

The arithmetic lives in tensorflow==2.8.4 (requirements.txt:336), which is absent here:
PARITY UNPINNED [3P-MEM].  The restatement follows the published CPU kernel
(resize_bilinear_op.cc: compute_interpolation_weights + compute_lerp, fp32, no FMA) and
numpy float32 keeps every multiply/add separately rounded.
"""
import numpy as np


def interpolation_weights(out_size, in_size):
    scale = np.float32(in_size) / np.float32(out_size)
    o = np.arange(out_size, dtype=np.float32)
    src = (o + np.float32(0.5)) * scale - np.float32(0.5)
    src_f = np.floor(src)
    lo = np.maximum(src_f.astype(np.int64), 0)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    lerp = (src - src_f).astype(np.float32)
    return lo, hi, lerp


def resize_bilinear_u8(frame, size, swap_rb=False):
    """frame: uint8 [H,W,3] -> uint8 [S,S,3] (truncating cast).  swap_rb: frame is BGR."""
    frame = np.asarray(frame)
    if swap_rb:
        frame = frame[..., ::-1]
    h, w = frame.shape[:2]
    ylo, yhi, yl = interpolation_weights(size, h)
    xlo, xhi, xl = interpolation_weights(size, w)
    f = frame.astype(np.float32)
    tl = f[ylo][:, xlo]
    tr = f[ylo][:, xhi]
    bl = f[yhi][:, xlo]
    br = f[yhi][:, xhi]
    xl = xl[None, :, None]
    yl = yl[:, None, None]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    out = top + (bot - top) * yl
    return out.astype(np.uint8)          # values are within [0, 255]: truncation


def preprocess_batch(frames, size, swap_rb=False):
    return np.stack([resize_bilinear_u8(f, size, swap_rb) for f in frames])
