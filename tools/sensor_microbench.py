"""HBM-bound kernels of subsystems 3b/4 at realistic sizes: achieved GB/s (algorithmic bytes / CUDA-event time)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps

g = torch.Generator(device="cuda").manual_seed(0)
for nx, ns in ((800000, 3), (800000, 192), (80000, 12288)):
    X = torch.linspace(700.0, 1500.0, nx, dtype=torch.float64, device="cuda")
    Y = torch.rand((nx, ns), dtype=torch.float64, device="cuda", generator=g)
    Xo, c, w = rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)
    cd, wd = torch.from_numpy(c).cuda(), torch.from_numpy(w).cuda()
    _, ms = timeit(lambda: engine.ils_apply(0, X, Y, cd, wd))
    dv = 800.0 / (nx - 1)
    support = float(np.sum(2 * w / dv))                 # samples inside the 128 triangle supports
    algo = 8.0 * (support * ns + support) + 8.0 * 128 * ns  # each covered sample read once per covering band
    uniq = 8.0 * nx * ns * (Xo.max() - Xo.min() + 2 * w.max()) / 800.0
    print("ILS triangle nx=%d ns=%d: %.3f ms, %.1f GB/s (bytes read per band support), unique bytes %.2f GB -> %.1f GB/s"
          % (nx, ns, ms, algo / ms / 1e6, uniq / 1e9, uniq / ms / 1e6), flush=True)
    _, msg = timeit(lambda: engine.ils_apply(1, X, Y, cd, wd), reps=5)
    print("ILS gauss    nx=%d ns=%d: %.3f ms (128 bands x all samples: %.2e exp/s)" % (nx, ns, msg, 128.0 * nx / msg * 1e3))
# apparent radiance: Compute_LWIR_Apparent_Radiance.py shape (128 bands, nE emissivities, nA atmospheres, 41 dT)
for nE, nA in ((64, 16), (1000, 100)):
    nx, nT = 128, 41
    X = torch.from_numpy(rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)[0]).cuda()
    emis = torch.rand((nx, nE), dtype=torch.float64, device="cuda", generator=g)
    Ts = 280 + 30 * torch.rand(nA, dtype=torch.float64, device="cuda", generator=g)
    tau, La, Ld = (torch.rand((nx, nA), dtype=torch.float64, device="cuda", generator=g) for _ in range(3))
    dT = torch.linspace(-10, 10, nT, dtype=torch.float64, device="cuda")
    L, ms = timeit(lambda: engine.apparent_radiance(X, emis, Ts, tau, La, Ld, dT=dT))
    print("apparent radiance nE=%d nA=%d nT=%d: %.3f ms, %.1f GB/s written (%.2f GB)" % (nE, nA, nT, ms, L.numel() * 8 / ms / 1e6, L.numel() * 8 / 1e9))
