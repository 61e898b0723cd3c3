"""Time rtx_voigt_sum alone under several settings (tile size K, reciprocal variant, wing) on config-2 data.
    python tools/voigt_microbench.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn  # noqa: E402


def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    only = sys.argv[2] if len(sys.argv) > 2 else None
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables()
    db = engine.LineDB([tabs[m] for m in sorted(tabs)])
    X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    grid = torch.from_numpy(X).cuda()
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    T = torch.from_numpy(atm["Ts"].copy()).cuda(); pd = torch.from_numpy(np.ascontiguousarray(p)).cuda()
    wd = torch.from_numpy(w).cuda()
    od = torch.empty((66, X.size), dtype=torch.float64, device="cuda")
    if only is None:
        a0 = db.voigt_sum(T, pd, wd, grid, flags=128 | 4096).clone()          # exact reciprocal, no far-field blocks
        for name, fl in (("default (one-step rcp + far-field)", 0), ("far-field only", 128), ("one-step rcp only", 4096),
                         ("slice-level far-field only", 128 | 16384)):
            a1 = db.voigt_sum(T, pd, wd, grid, flags=fl)
            rel = ((a1 - a0).abs() / a0.abs().clamp_min(1e-300)).max().item()
            print("max relative deviation of %s vs exact: %.3e" % (name, rel), flush=True)
        del a0, a1
    cases = []
    for K in (8,):
        for fast in (0, 16384, 128, 4096, 128 | 4096, 2, 4):
            cases.append(("cfg2 K=%d flags=%d" % (K, fast), dict(flags=fast | (K << 8)), T, pd, wd))
    # all-wide: every line +-25 cm^-1 -> almost every visit takes the test-free path
    cases.append(("wide25 K=8 fast=0", dict(flags=8 << 8, wing_abs=25.0), T[:8], pd[:8], wd[:8]))
    cases.append(("wide25 K=8 notile", dict(flags=16384 | (8 << 8), wing_abs=25.0), T[:8], pd[:8], wd[:8]))
    cases.append(("wide25 K=8 nofar", dict(flags=4096 | (8 << 8), wing_abs=25.0), T[:8], pd[:8], wd[:8]))
    cases.append(("wide25 K=8 fp32", dict(flags=2 | (8 << 8), wing_abs=25.0), T[:8], pd[:8], wd[:8]))
    cases.append(("wide25 K=4 fast=0", dict(flags=4 << 8, wing_abs=25.0), T[:8], pd[:8], wd[:8]))
    # per-layer groups of config 2
    for name, sl in (("layers 0-9", slice(0, 10)), ("layers 10-29", slice(10, 30)), ("layers 30-49", slice(30, 50)),
                     ("layers 50-65", slice(50, 66))):
        for K in (8, 4):
            cases.append((name + " K=%d" % K, dict(flags=K << 8), T[sl], pd[sl], wd[sl].contiguous()))
        cases.append((name + " K=8 notile", dict(flags=16384 | (8 << 8)), T[sl], pd[sl], wd[sl].contiguous()))
    for name, kw, Tt, pp, ww in cases:
        if only is not None and only not in name.replace(" ", "_"):
            continue
        n = Tt.numel()
        ev, evw = db.count_evals(Tt, pp, ww, grid, **{k: v for k, v in kw.items() if k != "flags"})
        ms = timeit(lambda: db.voigt_sum(Tt, pp, ww, grid, out=od[:n], **kw), reps)
        print("%-24s %8.3f ms  %10.3e evals  %6.3f Teval/s  weideman %.4f" % (name, ms, ev, ev / ms / 1e9, evw / max(ev, 1)),
              flush=True)


if __name__ == "__main__":
    main()
