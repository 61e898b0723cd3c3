"""Time rtx_tud alone on config-2-sized input (random optical depths)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
atm = syn.standard_atmosphere()
X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
grid = torch.from_numpy(X).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
od = torch.rand((66, X.size), dtype=torch.float64, device="cuda", generator=g) * 0.2
T = atm["Ts"][None, :]
for reps in (1, 5, 20):
    engine.tud(od, T, grid, atm["Zs"], [500.0], [1.0], 30); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.tud(od, T, grid, atm["Zs"], [500.0], [1.0], 30)
    e1.record(); torch.cuda.synchronize()
    print("tud x%d: %.3f ms each" % (reps, e0.elapsed_time(e1) / reps), flush=True)
