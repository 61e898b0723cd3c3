"""Where does the host time of engine.tud go when it is not hidden behind a running kernel?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
opts = dict(rt.options); rt.set_line_database(tabs, opts=opts); db = opts["LINE_DB"]
X_, grid = rt._device_axis(700.0, 1500.0, 0.001, db.device)
T = np.asarray(atm["Ts"], dtype=float)
p, w = rt.layer_weights(db, T, atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
Td, pd, wd = (engine.to_device(a, db.device) for a in (T, p, w))
od = db.voigt_sum(Td, pd, wd, grid)
torch.cuda.synchronize()
for i in range(4):
    t0 = time.perf_counter()
    tau, lu, ld = engine.tud(od, Td[None, :], grid, atm["Zs"], np.array([500.0]), np.array([1.0]), 30, False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("call returned after %.3f ms, device done after %.3f ms more" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); engine.tud(od, Td[None, :], grid, atm["Zs"], np.array([500.0]), np.array([1.0]), 30, False); e1.record()
torch.cuda.synchronize(); print("events: %.3f ms" % e0.elapsed_time(e1))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
engine.tud(od, Td[None, :], grid, atm["Zs"], np.array([500.0]), np.array([1.0]), 30, False)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
