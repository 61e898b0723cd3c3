"""One coarse-grid batch launch (config-3 regime: N atmospheres x 66 layers at 0.01 cm^-1) for ncu captures.
    python tools/batch_ncu_run.py [n_atm]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, synthetic as syn, radiative_transfer as rt
n_atm = int(sys.argv[1]) if len(sys.argv) > 1 else 8
atm = syn.standard_atmosphere(); tabs = syn.config2_tables()
opts = dict(rt.options); rt.set_line_database(tabs, opts=opts); db = opts["LINE_DB"]
Ts, MF = syn.perturbed_atmospheres(2, n_atm)
p, w = rt.layer_weights(db, Ts, atm["Ps"][None, :], atm["PLs"][None, :], MF, atm["MFs_ID"])
grid = torch.from_numpy(rt.make_spectral_axis(700.0, 1500.0, 0.01)).cuda()
T = torch.from_numpy(Ts.reshape(-1).copy()).cuda()
pd = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p, Ts.shape)).reshape(-1)).cuda()
wd = torch.from_numpy(w.reshape(T.numel(), -1)).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev, evw = db.count_evals(T, pd, wd, grid)
for K in [int(k) for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]:
    for i in range(2):
        e0.record(); out = db.voigt_sum(T, pd, wd, grid, flags=K << 8); e1.record(); torch.cuda.synchronize()
    print("n_atm %d K=%d (used %d): %.3f ms, %.3e evals, %.3f Teval/s, weideman %.4f" % (n_atm, K, db.handle.last_tile(), e0.elapsed_time(e1), ev, ev / e0.elapsed_time(e1) / 1e9, evw / ev))
