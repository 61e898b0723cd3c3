"""Timing breakdown of the reference cross-section workload (tools/run_configs.py xs): fraction of speed-dependent lines,
plain / extended kernel instantiation, far-field levels.    python tools/xs_breakdown.py [n_states]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, hapi_compat, synthetic as syn  # noqa: E402


def main():
    nS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n_lines = 60000
    nX = round((7100 - 400) / 0.0025) + 1
    X = torch.from_numpy(np.linspace(400, 7100, nX)).cuda()
    T = torch.linspace(275, 320, nS, dtype=torch.float64, device="cuda")
    P = torch.linspace(0.85, 1.05, nS, dtype=torch.float64, device="cuda")
    out = torch.empty((nS, nX), dtype=torch.float64, device="cuda")
    for frac in (0.0, 0.05, 1.0):
        cols = syn.synthetic_lines(77, n_lines, 300.0, 7200.0, 1)
        rng = np.random.default_rng(78)
        cols["SD_air"] = np.where(rng.random(n_lines) < frac, rng.uniform(0.05, 0.15, n_lines), 0.0)
        hapi_compat.storeTable("xsb", cols)
        hapi_compat.invalidate("xsb")
        db = hapi_compat.line_database(["xsb"])
        w = torch.from_numpy(np.tile(db.iso_factors(None, 1.0), (nS, 1))).cuda()
        for kern, base in (("voigt", 0), ("sdvoigt", engine.VOIGT_SDVOIGT)):
            if frac == 1.0 and kern == "voigt":
                continue
            row = []
            for fl in (0, engine.VOIGT_TILEFAR, engine.VOIGT_NOTILEFAR, engine.VOIGT_NOFAR):
                f = lambda: db.voigt_sum(T, P, w, X, wing_hw=350.0, out=out, flags=base | fl | (8 << 8))
                f(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                row.append(e0.elapsed_time(e1))
            print("SD fraction %.2f  %-8s  automatic %.2f ms   tile level forced %.2f ms   slice level only %.2f ms   no far field %.2f ms" % (frac, kern, *row), flush=True)


if __name__ == "__main__":
    main()
