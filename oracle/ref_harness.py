"""Container-only loader for the UNMODIFIED reference (westi024/RadTxfr at /root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It exists to
(1) mint golden vectors (tools/make_golden.py) and (2) cross-check the numpy restatement in
oracle/port.py while /root/reference is mounted (i.e. in the build container, never on the
GPU box).

Recipes follow SURVEY.md App. F:
  * misc/hapi.py imports as-is (prints a banner); lines are injected through
    hapi.LOCAL_TABLE_CACHE (misc/hapi.py:438-463) and the Voigt driver must be called with
    OmegaGrid= (the arange_ route raises on numpy>=1.18, misc/hapi.py:133-139).
  * radiative_transfer.py needs matplotlib stubbed (radiative_transfer.py:64) and two module
    globals patched: make_spectral_axis (float nX, :269-270) and compute_OD (LBLRTM, :395).
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("RADTXFR_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "misc", "hapi.py"))


_hapi = None
_rt = None
_ils = None


def hapi():
    """Import the reference's misc/hapi.py unmodified."""
    global _hapi
    if _hapi is None:
        sys.dont_write_bytecode = True
        p = os.path.join(REF_ROOT, "misc")
        if p not in sys.path:
            sys.path.insert(0, p)
        import warnings
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import hapi as _h
        _hapi = _h
    return _hapi


def rt():
    """Import the reference's radiative_transfer.py with matplotlib stubbed."""
    global _rt
    if _rt is None:
        sys.dont_write_bytecode = True
        for name in ("matplotlib", "matplotlib.pyplot"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        with contextlib.redirect_stdout(io.StringIO()):
            import radiative_transfer as _r
        _rt = _r
    return _rt


def ils_gauss():
    """Import the reference's standalone ILS_MAKO.py (Gaussian line shape)."""
    global _ils
    if _ils is None:
        sys.dont_write_bytecode = True
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        import importlib
        _ils = importlib.import_module("ILS_MAKO")
    return _ils


def inject_table(name, cols):
    """Put a dict-of-columns line table into hapi.LOCAL_TABLE_CACHE (misc/hapi.py:438-463)."""
    h = hapi()
    n = len(cols["nu"])
    data = {}
    for k, v in cols.items():
        if k in ("molec_id", "local_iso_id"):
            data[k] = [int(x) for x in v]
        else:
            data[k] = [float(x) for x in v]
    h.LOCAL_TABLE_CACHE[name] = {"header": {"number_of_rows": n, "table_name": name}, "data": data}


def ref_xsect(cols, T, p, grid, name="__oracle__", **kw):
    """sigma(omega) from the reference's absorptionCoefficient_Voigt (misc/hapi.py:10906)."""
    h = hapi()
    inject_table(name, cols)
    with contextlib.redirect_stdout(io.StringIO()):
        om, xs = h.absorptionCoefficient_Voigt(SourceTables=name, Environment={"T": float(T), "p": float(p)},
                                               OmegaGrid=grid, **kw)
    return om, xs


def ref_xsect_sd(cols, T, p, grid, name="__oracle_sd__", **kw):
    """sigma(omega) from the reference's absorptionCoefficient_SDVoigt (misc/hapi.py:10657)."""
    h = hapi()
    inject_table(name, cols)
    with contextlib.redirect_stdout(io.StringIO()):
        om, xs = h.absorptionCoefficient_SDVoigt(SourceTables=name, Environment={"T": float(T), "p": float(p)},
                                                 OmegaGrid=grid, **kw)
    return om, xs


def ref_tud(tables_by_mol, Xmin, Xmax, **kw):
    """Run the reference's compute_TUD body unmodified (radiative_transfer.py:274-392) with
    compute_OD replaced by the hapi-based wiring of SURVEY.md App. B.1."""
    import numpy as np
    r = rt()
    h = hapi()
    for m, cols in tables_by_mol.items():
        inject_table("__mol%d__" % m, cols)

    def make_axis(a, b, dv):
        return np.linspace(a, b, int(np.ceil((b - a) / dv)))

    def hapi_od(Xmin_in, Xmax_in, opts=None, **k):
        o = dict(r.options)
        o.update(k)
        X = make_axis(Xmin_in, Xmax_in, o["DVOUT"])
        T = float(o["T"]); p = float(o["P"]) / 101325.0
        od = np.zeros_like(X)
        for mid, val in zip(np.asarray(o["MF_ID"]).ravel(), np.asarray(o["MF_VAL"]).ravel()):
            mid = int(mid)
            if mid not in tables_by_mol:
                continue
            with contextlib.redirect_stdout(io.StringIO()):
                _, xs = h.absorptionCoefficient_Voigt(SourceTables="__mol%d__" % mid,
                                                      Environment={"T": T, "p": p}, OmegaGrid=X,
                                                      HITRAN_units=True)
            od += (val * 1e-6) * h.volumeConcentration(p, T) * (float(o["PL"]) * 1e5) * xs
        return X, od

    saved = dict(r.options)
    old_axis, old_od = r.make_spectral_axis, r.compute_OD
    r.make_spectral_axis, r.compute_OD = make_axis, hapi_od
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            out = r.compute_TUD(Xmin, Xmax, **kw)
    finally:
        r.make_spectral_axis, r.compute_OD = old_axis, old_od
        r.options.clear(); r.options.update(saved)
    return out
