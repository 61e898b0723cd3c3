"""Extract the DATA tables the hot path needs from the reference into radtxfr_b200/data/.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tools/extract_reference_data.py

What is extracted (facts / physical constants tables, no code):
  * isotopologue metadata (M, I, natural abundance, molar mass)     misc/hapi.py:3372-3497 (ISO)
  * TIPS-2011 partition-sum tables: Tdat (119 pts) and float32 Q(T)  misc/hapi.py:5401-9564
  * the 1976 US Standard Atmosphere table (66 layers x 15 columns)   radiative_transfer.py:75-148
  * the 128 MAKO band centres in micrometres                         radiative_transfer.py:1092-1223,
                                                                     ILS_MAKO.py:8-18 (identical)
"""
import ast
import inspect
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as R  # noqa: E402


def _first_long_list(src, n):
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.List) and len(node.elts) == n:
            return np.array([ast.literal_eval(e) for e in node.elts], dtype=np.float64)
    raise RuntimeError("literal list of %d not found" % n)


def main():
    h = R.hapi()
    rt = R.rt()
    keys = sorted(k for k in h.ISO.keys())
    M = np.array([k[0] for k in keys], dtype=np.int32)
    I = np.array([k[1] for k in keys], dtype=np.int32)
    abun = np.array([h.ISO[k][h.ISO_INDEX["abundance"]] for k in keys], dtype=np.float64)
    mass = np.array([h.ISO[k][h.ISO_INDEX["mass"]] for k in keys], dtype=np.float64)
    npt = int(h.TIPS_NPT)
    Q = np.full((len(keys), npt), np.nan, dtype=np.float32)
    has_q = np.zeros(len(keys), dtype=np.bool_)
    gsi = np.zeros(len(keys), dtype=np.float64)
    for j, k in enumerate(keys):
        if k in h.TIPS_ISO_HASH:
            q = np.asarray(h.TIPS_ISO_HASH[k])
            assert q.dtype == np.float32
            if q.shape != (npt,):
                continue  # 12 placeholder entries of length 1: AtoB cannot interpolate them (IndexError upstream)
            Q[j] = q
            has_q[j] = True
            gsi[j] = h.TIPS_GSI_HASH[k]
    Tdat = np.asarray(h.Tdat, dtype=np.float64)
    mako_tri = _first_long_list(inspect.getsource(rt.ILS_MAKO), 128)
    with open(os.path.join(R.REF_ROOT, "ILS_MAKO.py")) as f:
        mako_gauss = _first_long_list(f.read(), 128)
    assert np.array_equal(mako_tri, mako_gauss)
    out = os.path.join(os.path.dirname(HERE), "radtxfr_b200", "data", "spectro_tables.npz")
    np.savez_compressed(out, iso_M=M, iso_I=I, iso_abundance=abun, iso_mass=mass, tips_T=Tdat, tips_Q=Q,
                        tips_has=has_q, tips_gsi=gsi, std_atmos=np.asarray(rt.StdAtmos, dtype=np.float64),
                        mako_um=mako_tri)
    print("wrote", out, os.path.getsize(out), "bytes;", len(keys), "isotopologues,", int(has_q.sum()), "with TIPS")


if __name__ == "__main__":
    main()


def write_mako_include():
    """csrc/mako_table.inc: the MAKO band centres as a C array for rtx_mako_bands (same values as the npz)."""
    z = np.load(os.path.join(os.path.dirname(HERE), "radtxfr_b200", "data", "spectro_tables.npz"))
    m = z["mako_um"]
    lines = ["// MAKO band centres in micrometres: the 128 literals of radiative_transfer.py:1092-1223 (a data table of the instrument),",
             "// written by tools/extract_reference_data.py from radtxfr_b200/data/spectro_tables.npz.  Do not edit by hand.",
             "static const double RTX_MAKO_UM[128] = {"]
    for i in range(0, 128, 4):
        lines.append("    " + ", ".join(repr(float(v)) for v in m[i:i + 4]) + ",")
    lines.append("};")
    with open(os.path.join(os.path.dirname(HERE), "radtxfr_b200", "csrc", "mako_table.inc"), "w") as f:
        f.write("\n".join(lines) + "\n")
