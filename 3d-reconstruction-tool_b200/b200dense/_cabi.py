"""ctypes binding of libb200dense.so (include/b200dense.h).

The product path: there is no CPU fallback.  If the shared library is missing or
cannot be loaded, importing a compute entry point raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_SRC = 8
PARITY, FAST, PARITY_CHAINS, PARITY_CHECK = 0, 1, 2, 3
SCHEDULE_REFERENCE, SCHEDULE_REDBLACK = 0, 1
STAGE_NO_INIT, STAGE_NO_CONF = 1, 2
ABI_VERSION = 3

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200DENSE_LIB selects another build of the same ABI (tuning experiments under profiles/microbench)
LIB_PATH = os.environ.get('B200DENSE_LIB') or os.path.join(os.path.dirname(_HERE), 'csrc', 'libb200dense.so')


class Cams(C.Structure):
    _fields_ = [('K', C.c_float * 9), ('Kinv', C.c_float * 9), ('R_ref', C.c_float * 9), ('t_ref', C.c_float * 3),
                ('n_src', C.c_int32), ('R_src', (C.c_float * 9) * MAX_SRC), ('t_src', (C.c_float * 3) * MAX_SRC)]


class Rng(C.Structure):
    _fields_ = [('seed', C.c_uint64), ('view_id', C.c_uint32), ('d_init_depth', C.c_void_p),
                ('d_init_normal', C.c_void_p), ('d_rand_u', C.c_void_p), ('d_rand_n', C.c_void_p)]


class PmParams(C.Structure):
    _fields_ = [('patch', C.c_int32), ('iters', C.c_int32), ('samples', C.c_int32), ('arith', C.c_int32),
                ('depth_min', C.c_double), ('depth_max', C.c_double), ('schedule', C.c_int32), ('stage', C.c_int32),
                ('it_begin', C.c_int32), ('it_end', C.c_int32)]


class B200DenseError(RuntimeError):
    pass


# every symbol include/b200dense.h declares: (name, restype, argtypes)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
_dp = C.POINTER(C.c_double)
PROTOTYPES = {
    'b200d_abi_version': (C.c_int, []),
    'b200d_last_error': (C.c_char_p, []),
    'b200d_device_count': (C.c_int, []),
    'b200d_scratch_bytes': (C.c_size_t, [_i, _i]),
    'b200d_fuse_scratch_bytes': (C.c_size_t, [_i, _i]),
    'b200d_ref_stats': (C.c_int, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'b200d_pm_cost': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(Cams), _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'b200d_pm_view': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(Cams), _i, _i, C.POINTER(PmParams), C.POINTER(Rng),
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200d_pm_viewprop': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(Cams), _i, _i, C.POINTER(PmParams), C.POINTER(_vp),
                                    C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200d_sweep_bound_check': (C.c_int, [C.POINTER(C.c_float), _i]),
    'b200d_sweep_fallbacks': (C.c_int, [C.POINTER(C.c_uint64), _i]),
    'b200d_plane_sweep': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(Cams), C.POINTER(C.c_float), _i, _i, _i, _i, _f, _i,
                                    _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200d_knn_mean_dist': (C.c_int, [_vp, C.c_int64, _i, _vp, _vp]),
    'b200d_selftest_fast_math': (C.c_int, [C.c_uint64, C.POINTER(C.c_uint64), _vp]),
    'b200d_measure_fp32_peak': (C.c_int, [_dp, _vp]),
    'b200d_prepare_image': (C.c_int, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'b200d_write_ply_ascii': (C.c_int, [C.c_char_p, _dp, C.POINTER(C.c_uint8), C.c_int64]),
    'b200d_geo_consistency': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(Cams), _i, _i, _f, _f, _vp, _vp]),
    'b200d_profile_begin': (C.c_int, [_i]),
    'b200d_profile_end': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), _i, C.POINTER(C.c_int)]),
    'b200d_mask_count': (C.c_int, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp]),
    'b200d_mask_backproject': (C.c_int, [_vp, _vp, _vp, _i, _i, _dp, _dp, _dp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises B200DenseError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200DenseError(
            f"{LIB_PATH} not found: build it with `python 3d-reconstruction-tool_b200/csrc/build.py` "
            "(there is no CPU fallback for the dense path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch; let it surface
        fn.restype = res
        fn.argtypes = args
    if lib.b200d_abi_version() != ABI_VERSION:
        raise B200DenseError(f"ABI version mismatch: library reports {lib.b200d_abi_version()}, "
                             f"binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200d_last_error().decode('utf-8', 'replace')
        raise B200DenseError(f"libb200dense error {rc}: {msg}")
