"""Fine-grained wall-clock breakdown of radiative_transfer.compute_TUD on config 2 (host view)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
opts = dict(rt.options); rt.set_line_database(tabs, opts=opts); db = opts["LINE_DB"]
kw = dict(Zs=atm["Zs"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_ID=atm["MFs_ID"], DVOUT=0.001, N_angle=30, Ts=atm["Ts"], MFs_VAL=atm["MFs_VAL"])
for _ in range(3):
    rt.compute_TUD(700.0, 1500.0, opts=dict(opts), **kw)
torch.cuda.synchronize()
N = 8
t0 = time.perf_counter()
for _ in range(N):
    rt.compute_TUD(700.0, 1500.0, opts=dict(opts), **kw)
print("compute_TUD: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / N))

# ---- the same call in pieces (each piece followed by a device sync, so the pieces do not overlap) ----
P = np.asarray(atm["Ps"], dtype=float); PL = np.asarray(atm["PLs"], dtype=float)
T = np.asarray(atm["Ts"], dtype=float); MF = np.asarray(atm["MFs_VAL"], dtype=float)
acc = {}
def lap(name, t):
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t); return time.perf_counter()
for _ in range(N):
    t = time.perf_counter()
    X_, grid = rt._device_axis(700.0, 1500.0, 0.001, db.device)
    p, w = rt.layer_weights(db, T, P, PL, MF, atm["MFs_ID"])
    Td, pd, wd = (engine.to_device(a, db.device) for a in (T, p, w))
    t = lap("host prep + H2D", t)
    od = db.voigt_sum(Td, pd, wd, grid, out=rt._od_workspace(db.device, T.size, grid.numel()))
    t = lap("voigt", t)
    tau, lu, ld = engine.tud(od, Td[None, :], grid, atm["Zs"], np.array([500.0]), np.array([1.0]), 30, False)
    t = lap("tud", t)
    outs = rt._host_outputs([(grid.numel(), 1, 1), (grid.numel(), 1, 1), (grid.numel(),)])
    t = lap("host outputs", t)
    rt._to_host(tau[0], lu[0], ld[0], outs=outs)
    t = lap("D2H + copy-out", t)
print({k: "%.3f ms" % (1e3 * v / N) for k, v in acc.items()}, "sum %.3f ms" % (1e3 * sum(acc.values()) / N))
