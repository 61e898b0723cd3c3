import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
from radtxfr_b200 import engine, synthetic as syn
from oracle import port
cols = syn.synthetic_lines(61, 120, 995.0, 1005.0, 1)
db = engine.LineDB(cols)
grid = np.linspace(996.0, 1004.0, 8001)
states = [(250.0, 1e-2), (240.0, 1e-3), (230.0, 2e-4), (220.0, 5e-5), (220.0, 1e-6)]
T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
out = db.voigt_sum(T, p, w, grid, wing_abs=1.0, flags=8 << 8).cpu().numpy()
for k, (Tk, pk) in enumerate(states):
    _, ref = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid, OmegaWing=1.0)
    rel = np.abs(out[k] - ref) / np.maximum(np.abs(ref), 1e-300)
    print("p=%g: max rel %.2e  (at value/peak %.1e)" % (pk, rel.max(), ref[rel.argmax()] / ref.max()))
