"""Build and load the C restatement of the Voigt profile sum (oracle/voigt_c.c) -- test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "voigt_c.c")
LIB = os.path.join(HERE, "_build", "liboracle_voigt.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -ffp-contract=off: no FMA contraction, like the numpy operations the reference executes
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.oracle_voigt_accumulate.restype = ctypes.c_int64
        L.oracle_voigt_accumulate.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 7 + [ctypes.c_int64, ctypes.c_void_p,
                                                                                        ctypes.c_void_p, ctypes.c_double,
                                                                                        ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def absorption_coefficient_voigt(cols, T=296.0, p=1.0, OmegaGrid=None, HITRAN_units=True, out=None, scale=1.0, **kw):
    """Same result as oracle.port.absorption_coefficient_voigt (per-line quantities from port.line_state, profile sum
    in C/OpenMP).  With `out` the scaled cross-section is accumulated into it (used to build optical depths)."""
    from . import port
    Omegas = np.ascontiguousarray(np.sort(np.asarray(OmegaGrid, dtype=float)))
    st = port.line_state(cols, T, p, kw.get("Diluent"), kw.get("GammaL", "gamma_air"), kw.get("Components"),
                         kw.get("IntensityThreshold", 0.0), kw.get("OmegaWing", 0.0), kw.get("OmegaWingHW", 50.0))
    factor = 1.0 if HITRAN_units else port.volume_concentration(p, T)
    sc = np.ascontiguousarray(scale * (factor * st["ampl"] * st["S"]))
    a, L = port.weideman_coeffs(24)
    if out is None:
        out = np.zeros(Omegas.size)
    arrs = [np.ascontiguousarray(st[k], dtype=float) for k in ("nu", "Shift0", "GamD", "Gam0", "wing")]
    keep = np.ascontiguousarray(st["keep"].astype(np.uint8))
    ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    ev = lib().oracle_voigt_accumulate(len(sc), *[ptr(x) for x in arrs], ptr(sc), ptr(keep), Omegas.size, ptr(Omegas),
                                       ptr(np.ascontiguousarray(a)), float(L), ptr(out), max(1, 4 * (os.cpu_count() or 1)))
    return Omegas, out, int(ev)


def layer_od(tables_by_mol, X, T, P_pa, PL_km, MF_VAL_ppmv, MF_ID, **kw):
    """oracle.port.layer_od with the profile sums in C."""
    from . import port
    p = P_pa / 101325.0
    od = np.zeros(X.size)
    for mid, val in zip(np.asarray(MF_ID).ravel(), np.asarray(MF_VAL_ppmv).ravel()):
        mid = int(mid)
        if mid not in tables_by_mol:
            continue
        xs = np.zeros(X.size)
        absorption_coefficient_voigt(tables_by_mol[mid], T=float(T), p=float(p), OmegaGrid=X, out=xs, **kw)
        od += (val * 1e-6) * port.volume_concentration(p, float(T)) * (float(PL_km) * 1e5) * xs
    return od
