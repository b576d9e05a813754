"""ctypes binding of libvbt_b200.so (the C ABI declared in include/vbt_b200.h).

There is no CPU implementation behind this module: if the shared library is missing,
or no sm_100 device is usable, every call raises.  PyTorch is used only to own device
memory and streams (`tensor.data_ptr()`, `torch.cuda.current_stream().cuda_stream`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvbt_b200.so')

OK, EINVAL, ECUDA, ECAPACITY, EFORMAT = 0, -1, -2, -3, -4
MAX_DETECTIONS = 25
ROW_COLS = 8
PHASE_COLS = 6


class VbtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libvbt_b200 error {code}: {msg}')
        self.code = code


class TrackerParams(C.Structure):
    _fields_ = [('det_thresh', C.c_double), ('iou_threshold', C.c_double),
                ('inertia', C.c_double), ('max_age', C.c_int), ('min_hits', C.c_int),
                ('delta_t', C.c_int), ('vdc_uses_class_column', C.c_int)]


_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/vbt_b200.h declares
PROTOTYPES = {
    'vbt_abi_version': (_I, []),
    'vbt_last_error': (C.c_char_p, []),
    'vbt_launch_count': (C.c_longlong, []),
    'vbt_preprocess_u8': (_I, [_P, _I, _I, _I, _I, _P, _I, _P]),
    'vbt_copy_rows_h2d': (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P]),
    'vbt_preprocess_rows_u8': (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    'vbt_model_create': (_I, [_P, _SZ, C.POINTER(_P)]),
    'vbt_model_destroy': (None, [_P]),
    'vbt_model_info': (_I, [_P, C.POINTER(C.c_longlong)]),
    'vbt_model_plan': (_I, [_P, _P]),
    'vbt_model_plan_kinds': (_I, [_P, _P]),
    'vbt_detect': (_I, [_P, _P, _I, _P, _SZ, _P, _P, _P]),
    'vbt_model_profile': (_I, [_P, _I]),
    'vbt_model_op_times': (_I, [_P, _P, C.POINTER(C.c_longlong)]),
    'vbt_postprocess_q8': (_I, [_P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P]),
    'vbt_pack_detections': (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _P]),
    'vbt_tracker_create': (_I, [_I, _I, C.POINTER(TrackerParams), C.POINTER(_P)]),
    'vbt_tracker_destroy': (None, [_P]),
    'vbt_tracker_reset': (_I, [_P, _P]),
    'vbt_tracker_update': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P]),
    'vbt_tracker_row_details': (_I, [_P, _P]),
    'vbt_tracker_status': (_I, [_P, _P, _P]),
    'vbt_tracker_peek': (_I, [_P, _I, _P, _P]),
    'vbt_velocity_create': (_I, [_I, _I, _I, C.POINTER(_P)]),
    'vbt_velocity_destroy': (None, [_P]),
    'vbt_velocity_reset': (_I, [_P, _P]),
    'vbt_velocity_update': (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _D, _D, _D, _I, _I, _P]),
    'vbt_velocity_read': (_I, [_P, _P, _P, _P, _P]),
    'vbt_running_average': (_I, [_P, _I, _P, _I, _P, _P]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VbtError(ECUDA, f'{LIB_PATH} is missing: build it with '
                                  f'`make -C vbt_b200/csrc` (or __graft_entry__.build()); '
                                  f'vbt_b200 has no CPU path')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """Raise VbtError for a negative status; pass non-negative values through."""
    if rc is not None and rc < 0:
        raise VbtError(rc, lib().vbt_last_error().decode('utf-8', 'replace'))
    return rc


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise VbtError(ECUDA, 'no CUDA device is visible: vbt_b200 runs its hot path only as '
                              'sm_100a kernels and has no CPU path')
    return torch
