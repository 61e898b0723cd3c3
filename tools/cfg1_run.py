"""Config 1 (single-layer cross-section, 50 001 points, 20 000 lines): GPU time of rtx_voigt_sum per tile size, by CUDA
events over back-to-back C-ABI calls.   python tools/cfg1_run.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radtxfr_b200 import engine, hapi_compat, synthetic as syn, _lib
cols = syn.config1_lines(); hapi_compat.storeTable("cfg1", cols)
grid = np.linspace(800.0, 1300.0, 50001)
db = hapi_compat.line_database(["cfg1"])
gd = torch.from_numpy(grid).cuda(); w = torch.from_numpy(db.iso_factors(None, 1.0)[None, :]).cuda()
Tt = torch.tensor([296.0], dtype=torch.float64, device="cuda"); pp = torch.tensor([1.0], dtype=torch.float64, device="cuda")
out = torch.empty((1, grid.size), dtype=torch.float64, device="cuda")
ev, _ = db.count_evals([296.0], [1.0], db.iso_factors(None, 1.0)[None, :], grid)
f = _lib.lib().rtx_voigt_sum
ref = None
for K in (0, 1, 2, 4, 8):
    args = (db.handle.h, 1, engine._ptr(Tt), engine._ptr(pp), engine._ptr(w), 1.0, 0.0, 0.0, 50.0, 0.0, engine._ptr(gd),
            grid.size, engine._ptr(out), K << 8, engine._stream(gd.device))
    for _ in range(10):
        f(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(300):
        f(*args)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 300 * 1e3
    o = out.clone()
    ref = o if ref is None else ref
    print("K=%d (used %d): %.1f us per call, %.3e evals/s, max rel dev vs first %.1e"
          % (K, db.handle.last_tile(), us, ev / us * 1e6, ((o - ref).abs() / ref).max().item()), flush=True)
