"""One launch of each sensor kernel at a representative size (for ncu captures):
    python tools/sensor_ncu_run.py [ils|hsi|hsi32|app|app32] [mode bits for ils]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt

what = sys.argv[1] if len(sys.argv) > 1 else "ils"
g = torch.Generator(device="cuda").manual_seed(0)
if what == "ils":
    nx, ns = 80000, 12288
    X = torch.linspace(700.0, 1500.0, nx, dtype=torch.float64, device="cuda")
    Y = torch.rand((nx, ns), dtype=torch.float64, device="cuda", generator=g)
    Xo, c, w = rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)
    v = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
    for _ in range(3):
        out = engine.ils_apply(v, X, Y, c, w)
elif what in ("hsi", "hsi32"):
    rng = np.random.default_rng(3)
    X = rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)[0]
    nB, nA, nE, nP = X.size, 16, 64, 512 * 512
    emis = torch.rand((nE, nB), dtype=torch.float64, device="cuda", generator=g)
    tau, La, Ld = (torch.rand((nA, nB), dtype=torch.float64, device="cuda", generator=g) for _ in range(3))
    atm = rng.integers(0, nA, nP).astype(np.int32)
    Tp = 290.0 + 3.0 * rng.standard_normal(nP)
    mat = rng.integers(0, nE, (nP, 2)).astype(np.int32)
    fr = rng.random((nP, 2)); fr /= fr.sum(axis=1)[:, None]
    for _ in range(3):
        out = engine.hsi_pixels(X, emis, tau, La, Ld, atm, Tp, mat, fr, out_dtype=torch.float32 if what == "hsi32" else None)
else:
    nx, nT, nE, nA = 128, 41, 1000, 100
    X = rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)[0]
    emis = torch.rand((nx, nE), dtype=torch.float64, device="cuda", generator=g)
    Ts = 280 + 30 * torch.rand(nA, dtype=torch.float64, device="cuda", generator=g)
    tau, La, Ld = (torch.rand((nx, nA), dtype=torch.float64, device="cuda", generator=g) for _ in range(3))
    dT = torch.linspace(-10, 10, nT, dtype=torch.float64, device="cuda")
    for _ in range(3):
        out = engine.apparent_radiance(X, emis, Ts, tau, La, Ld, dT=dT, out_dtype=torch.float32 if what == "app32" else None)
torch.cuda.synchronize()
print("ok", what, tuple(out.shape), out.dtype)
