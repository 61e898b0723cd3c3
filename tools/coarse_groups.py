"""Config-3 regime (0.01 cm^-1 grid, 80 000 points): time of rtx_voigt_sum per layer group, 16 copies of each layer so
that the launch fills the GPU.    python tools/coarse_groups.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn  # noqa: E402
from tools.voigt_microbench import timeit  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
NC = 16
atm = syn.standard_atmosphere()
tabs = syn.config2_tables()
db = engine.LineDB([tabs[m] for m in sorted(tabs)])
X = rt.make_spectral_axis(700.0, 1500.0, 0.01)
grid = torch.from_numpy(X).cuda()
p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
tot = 0.0
for lo, hi in ((0, 10), (10, 20), (20, 30), (30, 40), (40, 50), (50, 66)):
    sl = slice(lo, hi)
    T = torch.from_numpy(np.tile(atm["Ts"][sl], NC)).cuda()
    pd = torch.from_numpy(np.tile(np.ascontiguousarray(p)[sl], NC)).cuda()
    wd = torch.from_numpy(np.tile(w[sl], (NC, 1))).cuda().contiguous()
    out = torch.empty((T.numel(), X.size), dtype=torch.float64, device="cuda")
    ev, evw = db.count_evals(T[:hi - lo], pd[:hi - lo], wd[:hi - lo], grid)
    row = []
    for K in (0, 4, 2):
        row.append(timeit(lambda: db.voigt_sum(T, pd, wd, grid, out=out, flags=K << 8), reps) / NC)
    tot += row[0]
    print("layers %2d-%2d: %.3e evals/atm (weideman %.3f, %5.0f points per line window)  automatic %.3f ms/atm  K=4 %.3f  K=2 %.3f" %
          (lo, hi - 1, ev, evw / max(ev, 1), ev / (hi - lo) / db.n_lines, *row), flush=True)
print("sum %.3f ms per atmosphere" % tot)
