"""CPU-side checks of the drop-in boundary: the shared library loads here, exports every
symbol include/vbt_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import pytest

from vbt_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      'include', 'vbt_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vbt_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run `make -C vbt_b200/csrc` first'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f'{n} declared in vbt_b200.h but not exported'
    assert sorted(_lib.PROTOTYPES) == names, 'ctypes prototypes out of sync with the header'
    assert _lib.lib().vbt_abi_version() == 1


def test_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = _lib.lib().vbt_velocity_create(1, 16, 4, ctypes.byref(h))
    assert rc == _lib.ECUDA
    assert b'no CPU path' in _lib.lib().vbt_last_error()
    from vbt_b200.velocity import VelocityTracker
    with pytest.raises(_lib.VbtError):
        VelocityTracker(0.45)
