"""Line-file readers (subsystem 1): HITRAN 160-column .par/.data files and LNFL TAPE3 binaries -> hapi-style
dict-of-columns tables that engine.LineDB / hapi_compat.storeTable accept.  Parsing is native (csrc/ingest.cu)."""
import ctypes

import numpy as np

from . import _lib

_PAR_COLS = ("molec_id", "local_iso_id", "nu", "sw", "a", "gamma_air", "gamma_self", "elower", "n_air", "delta_air")
_TP3_COLS = ("molec_id", "local_iso_id", "nu", "sw", "gamma_air", "gamma_self", "elower", "n_air", "delta_air")


def _alloc(cols, n):
    return {c: np.zeros(n, dtype=np.int32 if c in ("molec_id", "local_iso_id") else np.float64) for c in cols}


def _ptrs(arrs, cols):
    return [arrs[c].ctypes.data_as(ctypes.c_void_p) for c in cols]


def read_par(path):
    """Parse a HITRAN .par / hapi .data file (misc/hapi.py:468-538, 1535-1556) into a dict of numpy columns with
    hapi's column names."""
    n = ctypes.c_int64()
    _lib.call("rtx_par_read", str(path).encode(), 0, *([None] * 10), ctypes.byref(n))
    arrs = _alloc(_PAR_COLS, n.value)
    m = ctypes.c_int64()
    _lib.call("rtx_par_read", str(path).encode(), n.value, *_ptrs(arrs, _PAR_COLS), ctypes.byref(m))
    assert m.value == n.value
    arrs["molec_id"] = arrs["molec_id"].astype(np.int64)
    arrs["local_iso_id"] = arrs["local_iso_id"].astype(np.int64)
    return arrs


def write_tape3(path, cols):
    """Write a table in the LNFL TAPE3 layout documented in csrc/ingest.cu (lines must be sorted by nu)."""
    n = len(cols["nu"])
    a = {c: np.ascontiguousarray(np.asarray(cols[c], dtype=np.int32 if c in ("molec_id", "local_iso_id") else np.float64))
         for c in _TP3_COLS}
    _lib.call("rtx_tape3_write", str(path).encode(), n, *_ptrs(a, _TP3_COLS))


def read_tape3(path, vmin=-np.inf, vmax=np.inf):
    """Read the lines of a TAPE3 file with vmin <= nu <= vmax.  UNVERIFIED against a real AER line file: the
    reference ships AER-v3.6-0500-6000.tp3 only as a Git-LFS pointer (pointing this function at that pointer
    raises)."""
    n = ctypes.c_int64()
    _lib.call("rtx_tape3_read", str(path).encode(), float(max(vmin, -1e300)), float(min(vmax, 1e300)), 0,
              *([None] * 9), ctypes.byref(n))
    arrs = _alloc(_TP3_COLS, n.value)
    m = ctypes.c_int64()
    _lib.call("rtx_tape3_read", str(path).encode(), float(max(vmin, -1e300)), float(min(vmax, 1e300)), n.value,
              *_ptrs(arrs, _TP3_COLS), ctypes.byref(m))
    arrs["molec_id"] = arrs["molec_id"].astype(np.int64)
    arrs["local_iso_id"] = arrs["local_iso_id"].astype(np.int64)
    return arrs


def split_by_molecule(cols):
    """{molec_id: table} as radiative_transfer.set_line_database expects."""
    out = {}
    mol = np.asarray(cols["molec_id"])
    for m in np.unique(mol):
        sel = mol == m
        out[int(m)] = {k: np.asarray(v)[sel] for k, v in cols.items()}
    return out
