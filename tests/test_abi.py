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


def test_c_host_compiles_links_and_runs(lib_path, tmp_path):
    """include/fnnu.h is a C header (not only C++): a C99 host built with -Wall -Wextra -Werror links every declared entry
    point and runs the calls that need no GPU (tests/c/abi_host.c) — no Python, no torch on that side of the boundary."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    src = os.path.join(ROOT, 'tests', 'c', 'abi_host.c')
    taken = set(re.findall(r'\(void\*\)(fnnu_[a-z0-9_]+)', open(src).read()))
    assert taken == set(_declared()), sorted(set(_declared()) ^ taken)
    exe = str(tmp_path / 'abi_host')
    libdir = os.path.dirname(lib_path)
    cuda_lib = '/usr/local/cuda/lib64'
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(ROOT, 'include'), src,
                        '-L', libdir, '-lfnnu', f'-Wl,-rpath,{libdir}', f'-Wl,-rpath-link,{cuda_lib}', '-o', exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = cuda_lib + ':' + env.get('LD_LIBRARY_PATH', '')
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and 'entry points linked' in r.stdout, r.stdout + r.stderr
