"""ctypes binding of include/radtxfr_b200.h.

There is no CPU fallback: if the shared library is missing or no B200 is visible, the first use raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RADTXFR_LIB", os.path.join(_PKG, "libradtxfr_b200.so"))

_lib = None

# name -> (argtypes); every function returns int except rtx_last_error
_SIGNATURES = {
    "rtx_version": [],
    "rtx_device_info": [c_void_p, POINTER(c_int), c_char_p, c_int],
    "rtx_create": [c_int, POINTER(c_void_p)],
    "rtx_destroy": [c_void_p],
    "rtx_malloc": [c_void_p, c_uint64, POINTER(c_void_p)],
    "rtx_free": [c_void_p, c_void_p],
    "rtx_memcpy_h2d": [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p],
    "rtx_memcpy_d2h": [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p],
    "rtx_stream_sync": [c_void_p, c_void_p],
    "rtx_set_iso_table": [c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "rtx_lines_upload": [c_void_p, c_int64] + [c_void_p] * 12,
    "rtx_lines_count": [c_void_p, POINTER(c_int64)],
    "rtx_lines_set_sd": [c_void_p, c_void_p, c_void_p],
    "rtx_lines_set_diluent": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "rtx_voigt_set_mix": [c_void_p, c_int, c_void_p],
    "rtx_par_read": [c_char_p, c_int64] + [c_void_p] * 10 + [POINTER(c_int64)],
    "rtx_tape3_write": [c_char_p, c_int64] + [c_void_p] * 9,
    "rtx_tape3_read": [c_char_p, c_double, c_double, c_int64] + [c_void_p] * 9 + [POINTER(c_int64)],
    "rtx_voigt_sum": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double,
                      c_double, c_void_p, c_int64, c_void_p, c_int, c_void_p],
    "rtx_voigt_count": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double,
                        c_double, c_void_p, c_int64, c_void_p, c_void_p],
    "rtx_tips": [c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "rtx_tud": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int,
                c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "rtx_set_iso_molecules": [c_void_p, c_int, c_void_p, c_void_p],
    "rtx_layer_weights": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                          c_void_p],
    "rtx_tud_from_lines": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                           c_double, c_double, c_double, c_double, c_double, c_int, c_void_p, c_int64, c_void_p,
                           c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "rtx_mako_bands": [c_int, c_double, c_double, c_double, c_double, c_double, c_double, c_int, c_void_p, c_void_p,
                       c_void_p, POINTER(c_int)],
    "rtx_planck": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p],
    "rtx_brightness": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_double, c_void_p, c_void_p],
    "rtx_apparent_radiance": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    "rtx_hsi_pixels": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                       c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "rtx_apparent_radiance_f32": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    "rtx_hsi_pixels_f32": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                           c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "rtx_ils_apply": [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                      c_void_p, c_void_p],
    "rtx_smooth_sym": [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "rtx_spline_eval": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p],
    "rtx_measure_fp64_peak": [c_void_p, POINTER(c_double)],
    "rtx_launch_count": [c_void_p, POINTER(c_uint64)],
    "rtx_voigt_last_tile": [c_void_p, POINTER(c_int)],
    "rtx_weideman_coeffs": [c_void_p],
}

EXPORTS = ["rtx_last_error"] + sorted(_SIGNATURES)


class RtxError(RuntimeError):
    pass


def lib():
    """Load libradtxfr_b200.so (built in-tree by `python -m radtxfr_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("radtxfr_b200: %s is missing. Build it with `python -m radtxfr_b200.build` "
                              "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.rtx_last_error.restype = c_char_p
        L.rtx_last_error.argtypes = []
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = c_int
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().rtx_last_error().decode("utf-8", "replace")
        raise RtxError("%s failed (%d): %s" % (what or "radtxfr_b200 call", rc, msg))


def call(name, *args):
    check(getattr(lib(), name)(*args), name)
