"""Config 3 style: batch of perturbed atmospheres at 0.01 cm^-1 (80 000 points), spectra/s on one GPU."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
n_atm = int(sys.argv[1]) if len(sys.argv) > 1 else 64
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
opts = dict(rt.options); rt.set_line_database(tabs, opts=opts); db = opts["LINE_DB"]
Ts, MF = syn.perturbed_atmospheres(2, n_atm)
kw = dict(Zs=atm["Zs"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_ID=atm["MFs_ID"], DVOUT=0.01, N_angle=30)
for k in (8, 8 + 1024):
    o = dict(opts); o["voigt"] = {"tile_k": k & 15, "ws": k > 15}
    rt.compute_TUD_batch(700.0, 1500.0, Ts[:8], MF[:8], opts=dict(o), **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rt.compute_TUD_batch(700.0, 1500.0, Ts[:32], MF[:32], opts=dict(o), **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("K=%d ws=%d: %.1f spectra/s" % (k & 15, k > 15, 32 / dt), flush=True)
for ils in (None, "triangle"):
    rt.compute_TUD_batch(700.0, 1500.0, Ts[:8], MF[:8], opts=dict(opts), ils=ils, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    X, tau, lu, ld = rt.compute_TUD_batch(700.0, 1500.0, Ts, MF, opts=dict(opts), ils=ils, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("batch %d atmospheres, ils=%s: %.1f ms -> %.1f spectra/s (out %s)" % (n_atm, ils, 1e3 * dt, n_atm / dt, tuple(tau.shape)), flush=True)
# eval count of one atmosphere at this resolution
p, w = rt.layer_weights(db, Ts[0], atm["Ps"], atm["PLs"], MF[0], atm["MFs_ID"])
grid = torch.from_numpy(rt.make_spectral_axis(700.0, 1500.0, 0.01)).cuda()
e, ew = db.count_evals(Ts[0], p, w, grid)
print("evals per atmosphere %.3e (weideman %.4f)" % (e, ew / e))
