"""Where does the host-side time of radiative_transfer.compute_TUD go? (config 2)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
opts = dict(rt.options); rt.set_line_database(tabs, opts=opts); db = opts["LINE_DB"]
kw = dict(Zs=atm["Zs"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_ID=atm["MFs_ID"], DVOUT=0.001, N_angle=30, Ts=atm["Ts"], MFs_VAL=atm["MFs_VAL"])
for _ in range(2):
    rt.compute_TUD(700.0, 1500.0, opts=dict(opts), **kw)
def T(label, f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); print("%-34s %7.3f ms" % (label, 1e3 * (time.perf_counter() - t0) / n), flush=True); return r
T("compute_TUD total", lambda: rt.compute_TUD(700.0, 1500.0, opts=dict(opts), **kw))
X = T("make_spectral_axis", lambda: rt.make_spectral_axis(700.0, 1500.0, 0.001))
pw = T("layer_weights", lambda: rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"]))
grid = T("grid H2D (pageable)", lambda: engine.to_device(X, db.device))
od = T("voigt_sum (incl. small H2D)", lambda: db.voigt_sum(atm["Ts"], pw[0], pw[1], grid, flags=0))
out = T("tud", lambda: engine.tud(od, atm["Ts"][None, :], grid, atm["Zs"], [500.0], [1.0], 30))
T("3x .cpu().numpy()", lambda: [o[0].cpu().numpy() for o in out])
pin = torch.empty((3, X.size), dtype=torch.float64, pin_memory=True)
def packed():
    pin[0].copy_(out[0].view(-1), non_blocking=True); pin[1].copy_(out[1].view(-1), non_blocking=True); pin[2].copy_(out[2].view(-1), non_blocking=True)
    torch.cuda.synchronize(); return pin.numpy()
T("3x D2H into pinned", packed)
T("pinned alloc+free (3,nX)", lambda: torch.empty((3, X.size), dtype=torch.float64, pin_memory=True))
T("X.copy()", lambda: X.copy())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): rt.compute_TUD(700.0, 1500.0, opts=dict(opts), **kw)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
