"""One state (a single cross-section) on config 2's grid and table: time of rtx_voigt_sum against the multi-state rate.
    python tools/single_state.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn  # noqa: E402
from tools.voigt_microbench import timeit  # noqa: E402

atm = syn.standard_atmosphere()
tabs = syn.config2_tables()
db = engine.LineDB([tabs[m] for m in sorted(tabs)])
p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
for dv in (0.001, 0.0025, 0.01):
    X = rt.make_spectral_axis(700.0, 1500.0, dv)
    grid = torch.from_numpy(X).cuda()
    for layers in ((0,), (20,), (45,), (0, 1, 2, 3), tuple(range(0, 16))):
        sl = list(layers)
        T = torch.from_numpy(atm["Ts"][sl].copy()).cuda(); pd = torch.from_numpy(np.ascontiguousarray(p)[sl]).cuda()
        wd = torch.from_numpy(w[sl]).cuda().contiguous()
        out = torch.empty((len(sl), X.size), dtype=torch.float64, device="cuda")
        ev, _ = db.count_evals(T, pd, wd, grid)
        row = []
        for fl in (0, 8 << 8, 4 << 8, 2 << 8):
            row.append(timeit(lambda: db.voigt_sum(T, pd, wd, grid, out=out, flags=fl), 5))
        print("dv %.4f (%7d pts) layers %-12s %.3e evals: automatic %.3f ms (%.2f Teval/s)  K=8 %.3f  K=4 %.3f  K=2 %.3f" %
              (dv, X.size, str(sl[:4]) + ("..." if len(sl) > 4 else ""), ev, row[0], ev / row[0] / 1e9, row[1], row[2], row[3]), flush=True)
