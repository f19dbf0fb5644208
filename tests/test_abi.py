"""CPU: libfnnu.so builds, loads, and exports every symbol include/fnnu.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from fast_nnunet_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'fnnu.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(fnnu_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/fnnu.h but not exported by libfnnu.so'


def test_binding_covers_header(lib_path):
    from fast_nnunet_b200 import _lib
    assert set(_lib.exported_symbols()) == set(_declared())
    lib = _lib.load()
    assert lib.fnnu_abi_version() == 1


def test_struct_layout_matches_c(lib_path):
    from fast_nnunet_b200 import _lib
    # int32 x19 (76 B) padded to 80, then 4 pointers
    assert ctypes.sizeof(_lib.OpDesc) == 112
    assert _lib.OpDesc.weight.offset == 80
    assert ctypes.sizeof(_lib.BufferDesc) == 16


def test_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_invalid_arguments_return_error_codes(lib_path):
    from fast_nnunet_b200 import _lib
    lib = _lib.load()
    I3 = ctypes.c_int32 * 3
    rc = lib.fnnu_gather_tiles(None, 1, I3(8, 8, 8), None, 1, I3(4, 4, 4), b'\0', 1, None, 1, None)
    assert rc == -1 and b'null' in lib.fnnu_last_error()
    rc = lib.fnnu_engine_forward(None, 1, None)
    assert rc == -1
    with pytest.raises(_lib.FnnuError):
        _lib.check(rc)
