"""Run rtx_voigt_sum on one layer group of config 2 (for ncu captures of a single regime).
    python tools/layer_group_run.py LO HI [reps] [K]      (K pins the tile size: one grid launch instead of two)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
K = int(sys.argv[4]) if len(sys.argv) > 4 else 0
atm = syn.standard_atmosphere()
tabs = syn.config2_tables()
db = engine.LineDB([tabs[m] for m in sorted(tabs)])
X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
grid = torch.from_numpy(X).cuda()
p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
sl = slice(lo, hi)
T = torch.from_numpy(atm["Ts"].copy()).cuda()[sl]
pd = torch.from_numpy(np.ascontiguousarray(p)).cuda()[sl]
wd = torch.from_numpy(w).cuda()[sl].contiguous()
for _ in range(reps):
    out = db.voigt_sum(T, pd, wd, grid, flags=K << 8)
torch.cuda.synchronize()
print("ok", float(out.sum()))
