"""ctypes binding of libpax.so (include/pax.h).  No CPU fallback: a missing library is fatal."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpax.so")

PAX_AX_PLAIN, PAX_AX_RESIDUAL = 0, 1
PAX_ATB_STORE, PAX_ATB_ADD, PAX_ATB_SART = 0, 1, 2
PAX_PROJECTOR_AUTO, PAX_PROJECTOR_RAY, PAX_PROJECTOR_FAN = 0, 1, 2
PAX_ATB_SART_MAX = 64
PAX_MAX_PEERS = 16
ANGLE_STRIDE = 8

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)


class PaxGeo(ctypes.Structure):
    _fields_ = [("nVoxelZ", ctypes.c_int32), ("nVoxelY", ctypes.c_int32), ("nVoxelX", ctypes.c_int32),
                ("nDetecV", ctypes.c_int32), ("nDetecU", ctypes.c_int32),
                ("dVoxelZ", ctypes.c_double), ("dVoxelY", ctypes.c_double), ("dVoxelX", ctypes.c_double),
                ("dDetecV", ctypes.c_double), ("dDetecU", ctypes.c_double),
                ("accuracy", ctypes.c_double)]


class PaxAxEpilogue(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("a0", ctypes.c_float), ("a1", ctypes.c_float),
                ("c", ctypes.c_float), ("b", ctypes.c_void_p),
                ("w_half_z", ctypes.c_float), ("w_half_y", ctypes.c_float), ("w_half_x", ctypes.c_float),
                ("w_thr", ctypes.c_float), ("sumsq", ctypes.c_void_p), ("nsamples", ctypes.c_void_p),
                ("out1", ctypes.c_void_p), ("maxout", ctypes.c_void_p)]


class PaxAtbEpilogue(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("noneg", ctypes.c_int32), ("lambda_", ctypes.c_float),
                ("inv_v", ctypes.c_void_p), ("row_part", ctypes.c_int32), ("row_parts", ctypes.c_int32)]


# every symbol include/pax.h declares: (name, restype, argtypes)
_VP = ctypes.c_void_p
SYMBOLS = [
    ("pax_version", ctypes.c_int, []),
    ("pax_last_error", ctypes.c_char_p, []),
    ("pax_plan_create", ctypes.c_int, [ctypes.POINTER(PaxGeo), c_double_p, ctypes.c_int,
                                       ctypes.POINTER(_VP)]),
    ("pax_plan_destroy", ctypes.c_int, [_VP]),
    ("pax_plan_n_angles", ctypes.c_int, [_VP]),
    ("pax_plan_set_projector", ctypes.c_int, [_VP, ctypes.c_int]),
    ("pax_plan_projector", ctypes.c_int, [_VP]),
    ("pax_plan_set_column_share", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int]),
    ("pax_plan_fan_info", ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                         c_double_p]),
    ("pax_plan_fault", ctypes.c_int, [_VP]),
    ("pax_vol_elems", ctypes.c_size_t, [_VP]),
    ("pax_vol_pack", ctypes.c_int, [_VP, _VP, _VP, _VP]),
    ("pax_vol_unpack", ctypes.c_int, [_VP, _VP, _VP, _VP]),
    ("pax_vol_axpy", ctypes.c_int, [_VP, _VP, _VP, ctypes.c_float, _VP]),
    ("pax_max_decode", ctypes.c_int, [_VP, _VP]),
    ("pax_lincomb", ctypes.c_int, [_VP, _VP, _VP, ctypes.c_float, ctypes.c_float, ctypes.c_size_t, _VP]),
    ("pax_ax_workspace_bytes", ctypes.c_size_t, [_VP, ctypes.c_int]),
    ("pax_ax_fan_scratch_bytes", ctypes.c_size_t, [_VP, ctypes.c_int]),
    ("pax_ax", ctypes.c_int, [_VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP,
                              ctypes.POINTER(PaxAxEpilogue), _VP, _VP]),
    ("pax_atb", ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP,
                               ctypes.POINTER(PaxAtbEpilogue), _VP]),
    ("pax_sart_update", ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_float, ctypes.c_int, _VP]),
    ("pax_bcast_chunks", ctypes.c_int, [_VP, _VP, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_size_t,
                                        ctypes.c_long, ctypes.c_long, ctypes.c_long, _VP]),
    ("pax_bcast_chunks2", ctypes.c_int, [_VP, ctypes.c_long, _VP, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int,
                                         ctypes.c_size_t, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_int, _VP]),
    ("pax_bcast_xrows", ctypes.c_int, [_VP, _VP, _VP, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, _VP,
                                       ctypes.c_int, ctypes.c_int, _VP]),
    ("pax_sart_update_fused", ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.POINTER(ctypes.c_uint64),
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, _VP,
                                             ctypes.c_float, ctypes.c_int, _VP]),
    ("pax_compute_w", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_float, ctypes.c_float, _VP, _VP]),
    ("pax_compute_v", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_double, _VP, _VP]),
    ("pax_reduce_max", ctypes.c_int, [_VP, ctypes.c_size_t, _VP, _VP]),
    ("pax_resample_linear", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, _VP,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_int, _VP]),
    ("pax_minmax", ctypes.c_int, [_VP, ctypes.c_size_t, _VP, _VP]),
    ("pax_rescale_intensity", ctypes.c_int, [_VP, ctypes.c_size_t, _VP, ctypes.c_double, ctypes.c_double,
                                             _VP]),
    ("pax_ts_create", ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_void_p)]),
    ("pax_ts_destroy", ctypes.c_int, [_VP]),
    ("pax_ts_set_volume", ctypes.c_int, [_VP, _VP, _VP]),
    ("pax_ts_ax", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    ("pax_ts_atb", ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    ("pax_resize", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, _VP]),
    ("pax_pair_sums", ctypes.c_int, [_VP, _VP, ctypes.c_size_t, _VP, _VP]),
    ("pax_plahe", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_float, ctypes.c_float, _VP, _VP, _VP]),
    ("pax_ctnoise", ctypes.c_int, [_VP, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint64,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_double, _VP, ctypes.c_uint64,
                                   _VP]),
]

_lib = None


class PaxError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """dlopen csrc/libpax.so and bind every declared symbol; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PaxError(
            f"{LIB_PATH} is missing: build it with `python -m physics_arx_b200.csrc.build` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pax_version() != 100:
        raise PaxError(f"libpax.so version {lib.pax_version()} does not match this package (100)")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().pax_last_error().decode(errors="replace")
        if status == -1:
            raise ValueError(f"{what}: {msg}")
        raise PaxError(f"{what}: {msg} (status {status})")
