import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn
from tools.voigt_microbench import timeit
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
db = engine.LineDB([tabs[m] for m in sorted(tabs)])
for dv in (0.001, 0.01):
    X = rt.make_spectral_axis(700.0, 1500.0, dv); grid = torch.from_numpy(X).cuda()
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    rep = 1 if dv < 0.005 else 8
    T = torch.from_numpy(np.tile(atm["Ts"], rep)).cuda(); pd = torch.from_numpy(np.tile(np.ascontiguousarray(p), rep)).cuda()
    wd = torch.from_numpy(np.tile(w, (rep, 1))).cuda()
    od = torch.empty((T.numel(), X.size), dtype=torch.float64, device="cuda")
    print("dv %.3f automatic: %.3f ms per atmosphere" % (dv, timeit(lambda: db.voigt_sum(T, pd, wd, grid, out=od), 5) / rep), flush=True)
