"""Break-even of the tile-level far-field blocks: rtx_voigt_sum on layers 0-7 of config 2 with every line's wing set to
wing_abs (OmegaWing), tile level on (flags 0) and off (RTX_VOIGT_NOTILEFAR).  python tools/tilefar_sweep.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn  # noqa: E402
from tools.voigt_microbench import timeit  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables()
    db = engine.LineDB([tabs[m] for m in sorted(tabs)])
    X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    grid = torch.from_numpy(X).cuda()
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    T = torch.from_numpy(atm["Ts"].copy()).cuda()[:8]
    pd = torch.from_numpy(np.ascontiguousarray(p)).cuda()[:8]
    wd = torch.from_numpy(w).cuda()[:8].contiguous()
    od = torch.empty((8, X.size), dtype=torch.float64, device="cuda")
    for wing in (1.5, 2.0, 3.0, 4.0, 5.0, 6.0, 8.0, 12.0, 16.0, 25.0):
        row = []
        for fl in (8 << 8, (8 << 8) | 16384, 0):
            row.append(timeit(lambda: db.voigt_sum(T, pd, wd, grid, out=od, flags=fl, wing_abs=wing, wing_hw=0.0), reps))
        print("wing %5.1f cm-1 (%6d points): tile level %.3f ms, without %.3f ms, automatic %.3f ms" %
              (wing, int(wing / 0.001), row[0], row[1], row[2]), flush=True)


if __name__ == "__main__":
    main()
