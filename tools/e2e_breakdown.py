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
