import sys, numpy as np
sys.path.insert(0, "/root/repo")
from radtxfr_b200 import engine, synthetic as syn
cols = syn.synthetic_lines(71, 3000, 990.0, 1010.0, 1)
db = engine.LineDB(cols)
for nS, grid in ((1, np.linspace(995.0, 1005.0, 10001)), (2, np.linspace(995.0, 1005.0, 120001))):
    T = np.linspace(296.0, 240.0, nS); p = np.geomspace(1.0, 0.1, nS)
    w = np.tile(db.iso_factors(None, 1.0), (nS, 1))
    ref = db.voigt_sum(T, p, w, grid, flags=65536 | 128 | 4096).cpu().numpy()
    for name, fl, tol in (("default", 0, 5e-12), ("ws", 4, 5e-12), ("fp32", 2, 1e-5), ("exact", 128, 5e-13), ("nofar", 4096, 5e-12), ("ws+fp32", 6, 1e-5)):
        a = db.voigt_sum(T, p, w, grid, flags=fl).cpu().numpy()
        b = db.voigt_sum(T, p, w, grid, flags=fl).cpu().numpy()
        dev = np.max(np.abs(a - ref) / np.maximum(ref, 1e-300))
        print(nS, grid.size, name, "max rel dev %.2e" % dev, "reproducible", np.array_equal(a, b), "OK" if dev < tol else "FAIL")
