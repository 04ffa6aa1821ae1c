"""CPU checks of the C-ABI boundary: the library builds, loads and exports every symbol include/robustcap_b200.h
declares; the ctypes table binds every one of them.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

from robustcap_b200 import _lib


def test_library_builds_and_exports_every_header_symbol():
    so = _lib.build()
    lib = ctypes.CDLL(so)
    names = _lib.exported_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.rc_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.rc_version()


def test_ctypes_table_covers_header():
    bound = set(_lib._SIGS) | set(_lib._OPTIONAL_SIGS)
    assert set(_lib.exported_symbols()) <= bound, set(_lib.exported_symbols()) - bound


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.build()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_compute_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from robustcap_b200 import math as M
    with pytest.raises(RuntimeError):
        M.r6d_to_rotation_matrix(torch.randn(4, 6))


def test_constants_bit_exact(golden_dir):
    import numpy as np
    from robustcap_b200 import constants as C
    g = np.load(os.path.join(golden_dir, 'kinematics.npz'))
    assert C.MP_MASK == g['mp_mask'].tolist()
    assert C.JI_MASK == g['ji_mask'].tolist()
    assert C.VI_MASK == g['vi_mask'].tolist()
    assert C.SMPL_PARENT == g['parent'].tolist()
