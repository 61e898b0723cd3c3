"""HBM-bound kernels of subsystems 3b/4 at realistic sizes: achieved GB/s (algorithmic bytes / CUDA-event time).
    python tools/sensor_microbench.py [ils|app|hsi ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn

which = sys.argv[1:] or ["ils", "app", "hsi"]
try:
    import json
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    HBM = 6444.4


def timeit(fn, reps=20):
    for _ in range(3):                 # let the caching allocator settle (outputs are allocated per call)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


g = torch.Generator(device="cuda").manual_seed(0)
if "ils" in which:
    for nx, ns in ((800000, 3), (800000, 192), (80000, 12288), (800000, 1536)):
        X = torch.linspace(700.0, 1500.0, nx, dtype=torch.float64, device="cuda")
        Y = torch.rand((nx, ns), dtype=torch.float64, device="cuda", generator=g)
        Xo, c, w = rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)
        cd, wd = torch.from_numpy(c).cuda(), torch.from_numpy(w).cuda()
        dv = 800.0 / (nx - 1)
        support = float(np.sum(2 * w / dv))                 # samples inside the 128 triangle supports
        lo, hi = (c - w).min(), (c + w).max()
        uniq = 8.0 * ns * (hi - lo) / dv                    # every sample of the union support once
        r0 = None
        for name, v in (("default", 0), ("TMA, 2 spectra per lane", 0x400), ("TMA, 1 spectrum per lane", 0x800), ("TMA, 16 bands, 2 per lane", 0x1400), ("TMA, 16 bands, 1 per lane", 0x1800), ("groups of 8, no staging", 0x200),
                        ("band per CTA (round 1)", 0x100)):
            r1, ms = timeit(lambda: engine.ils_apply(v, X, Y, cd, wd))
            r0 = r1 if r0 is None else r0
            dev = ((r1 - r0).abs() / r0.abs()).max().item()
            print("ILS triangle nx=%d ns=%d %-26s %.3f ms: unique bytes %.2f GB -> %.0f GB/s = %.2f of HBM (%.0f GB/s of band-support "
                  "bytes); dev vs default %.1e" % (nx, ns, name, ms, uniq / 1e9, uniq / ms / 1e6, uniq / ms / 1e6 / HBM,
                                                   8.0 * support * ns / ms / 1e6, dev), flush=True)
        _, msg = timeit(lambda: engine.ils_apply(1, X, Y, cd, wd), reps=5)
        print("ILS gauss    nx=%d ns=%d: %.3f ms" % (nx, ns, msg), flush=True)
if "app" in which:
    # apparent radiance: Compute_LWIR_Apparent_Radiance.py shape (128 bands, nE emissivities, nA atmospheres, 41 dT)
    for nE, nA in ((64, 16), (1000, 100)):
        nx, nT = 128, 41
        X = torch.from_numpy(rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)[0]).cuda()
        emis = torch.rand((nx, nE), dtype=torch.float64, device="cuda", generator=g)
        Ts = 280 + 30 * torch.rand(nA, dtype=torch.float64, device="cuda", generator=g)
        tau, La, Ld = (torch.rand((nx, nA), dtype=torch.float64, device="cuda", generator=g) for _ in range(3))
        dT = torch.linspace(-10, 10, nT, dtype=torch.float64, device="cuda")
        from radtxfr_b200 import _lib
        h = engine.default_handle(X.device)
        for dt, fn in ((torch.float64, "rtx_apparent_radiance"), (torch.float32, "rtx_apparent_radiance_f32")):
            # preallocated output, C-ABI entry called directly (a 4 GB allocation per call would be timed otherwise)
            L = torch.empty((nx, nE, nA, nT), dtype=dt, device="cuda")
            args = (h.h, engine._ptr(X), nx, engine._ptr(emis), nE, engine._ptr(Ts), nA, engine._ptr(tau), engine._ptr(La),
                    engine._ptr(Ld), engine._ptr(dT), nT, engine._ptr(L), None, engine._stream(X.device))
            f = getattr(_lib.lib(), fn)
            _, ms = timeit(lambda: f(*args), reps=20)
            b = L.numel() * L.element_size()
            print("apparent radiance nE=%d nA=%d nT=%d %s: %.3f ms, %.0f GB/s written = %.2f of HBM (%.2f GB)"
                  % (nE, nA, nT, L.dtype, ms, b / ms / 1e6, b / ms / 1e6 / HBM, b / 1e9), flush=True)
            del L
if "hsi" in which:
    # config 4: 512 x 512 pixels x 128 bands, 16 atmospheres, 64 emissivities, two-material mixes
    rng = np.random.default_rng(3)
    X = torch.from_numpy(rt._mako_bands(np.array([700.0, 1500.0]), None, 1.0, 0.0, 1.0)[0]).cuda()
    nB, nA, nE, nP = X.numel(), 16, 64, 512 * 512
    emis = torch.rand((nE, nB), dtype=torch.float64, device="cuda", generator=g)
    tau, La, Ld = (torch.rand((nA, nB), dtype=torch.float64, device="cuda", generator=g) for _ in range(3))
    atm = torch.from_numpy(rng.integers(0, nA, nP).astype(np.int32)).cuda()
    Tp = torch.from_numpy(290.0 + 3.0 * rng.standard_normal(nP)).cuda()
    mat = torch.from_numpy(rng.integers(0, nE, (nP, 2)).astype(np.int32)).cuda()
    fr = rng.random((nP, 2)); fr /= fr.sum(axis=1)[:, None]
    fr = torch.from_numpy(fr).cuda()
    # the kernel takes tens of microseconds, less than the Python call path: time the C-ABI entry directly with
    # prebuilt arguments (a ctypes call is ~3 us), 200 launches back to back
    from radtxfr_b200 import _lib
    h = engine.default_handle(X.device)
    st = engine._stream(X.device)
    for dt, fn in ((torch.float64, "rtx_hsi_pixels"), (torch.float32, "rtx_hsi_pixels_f32")):
        L = torch.empty((nP, nB), dtype=dt, device="cuda")
        args = (h.h, engine._ptr(X), nB, engine._ptr(emis), nE, engine._ptr(tau), engine._ptr(La), engine._ptr(Ld), nA, nP,
                engine._ptr(atm), engine._ptr(Tp), 2, engine._ptr(mat), engine._ptr(fr), engine._ptr(L), st)
        f = getattr(_lib.lib(), fn)
        _, ms = timeit(lambda: f(*args), reps=200)
        b = L.numel() * L.element_size()
        print("hsi cube 512x512x%d %s: %.4f ms, %.0f GB/s written = %.2f of HBM (%.3f GB)"
              % (nB, L.dtype, ms, b / ms / 1e6, b / ms / 1e6 / HBM, b / 1e9), flush=True)
