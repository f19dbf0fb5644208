"""ctypes binding of libfnnu.so (include/fnnu.h).  No fallback: a missing library or a failing call
raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('FNNU_LIB') or os.path.join(_HERE, 'libfnnu.so')     # FNNU_LIB: instrumented profiling builds

OP_CONV, OP_TCONV, OP_ADD_ACT, OP_AVGPOOL = 0, 1, 2, 3
ACC_F32, ACC_F16 = 0, 1
IN_F16, IN_F32 = 0, 1
E_INF = -4


class BufferDesc(C.Structure):
    _fields_ = [('dims', C.c_int32 * 3), ('channels', C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [('op', C.c_int32),
                ('src', C.c_int32), ('src_coff', C.c_int32),
                ('src2', C.c_int32), ('src2_coff', C.c_int32),
                ('dst', C.c_int32), ('dst_coff', C.c_int32),
                ('cin', C.c_int32), ('cout', C.c_int32),
                ('kernel', C.c_int32 * 3), ('stride', C.c_int32 * 3),
                ('has_bias', C.c_int32), ('has_norm', C.c_int32),
                ('norm_eps', C.c_float), ('act_slope', C.c_float),
                ('weight', C.c_void_p), ('bias', C.c_void_p), ('gamma', C.c_void_p), ('beta', C.c_void_p)]


class FnnuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libfnnu error {code}: {msg}')
        self.code = code


_lib = None

_I3 = C.c_int32 * 3

_SIGNATURES = {
    'fnnu_abi_version': (C.c_int, []),
    'fnnu_last_error': (C.c_char_p, []),
    'fnnu_device_ok': (C.c_int, []),
    'fnnu_gather_tiles': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_int,
                                    C.POINTER(C.c_int32), C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'fnnu_accumulate_tiles': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int,
                                        C.POINTER(C.c_int32), C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_int32), C.c_void_p]),
    'fnnu_weight_sum': (C.c_int, [C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int,
                                  C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                  C.c_int, C.POINTER(C.c_int32), C.c_void_p]),
    'fnnu_finalize': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_size_t, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    'fnnu_mem_launches': (C.c_longlong, []),
    'fnnu_pre_nonzero_bbox': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    'fnnu_pre_filled_mask': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'fnnu_pre_channel_stats': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    'fnnu_pre_crop_normalize': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    'fnnu_pre_resample_workspace_bytes': (C.c_size_t, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'fnnu_pre_resample_channel': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                            C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    'fnnu_export_labels': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    'fnnu_scale_inplace_f32': (C.c_int, [C.c_void_p, C.c_float, C.c_size_t, C.c_void_p]),
    'fnnu_add_inplace_f32': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'fnnu_engine_sizes': (C.c_int, [C.POINTER(BufferDesc), C.c_int, C.POINTER(OpDesc), C.c_int, C.c_int,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    'fnnu_engine_create': (C.c_int, [C.POINTER(BufferDesc), C.c_int, C.POINTER(OpDesc), C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.POINTER(C.c_void_p)]),
    'fnnu_engine_destroy': (None, [C.c_void_p]),
    'fnnu_engine_buffer': (C.c_void_p, [C.c_void_p, C.c_int]),
    'fnnu_engine_stats': (C.c_void_p, [C.c_void_p, C.c_int]),
    'fnnu_engine_forward': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    'fnnu_engine_set_backend': (C.c_int, [C.c_void_p, C.c_int]),
    'fnnu_engine_launch_counts': (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'fnnu_engine_profile_op': (C.c_int, [C.c_void_p, C.c_int]),
    'fnnu_engine_profile_ms': (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Loads libfnnu.so and types every entry point.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f'{LIB_PATH} is missing: build it with `python -m fast_nnunet_b200.build` (there is no CPU or '
            f'PyTorch fallback for the inference path)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fnnu_abi_version() != 1:
        raise RuntimeError(f'libfnnu ABI version {lib.fnnu_abi_version()} != 1')
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise FnnuError(rc, load().fnnu_last_error().decode('utf-8', 'replace'))


def i3(v):
    return _I3(int(v[0]), int(v[1]), int(v[2]))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
