"""Mint golden vectors by executing the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

Writes tests/golden/*.npz.  Each file stores the inputs (or the seeds that regenerate them through
radtxfr_b200.synthetic) and the outputs of the reference function named in the file's `source` entry.
The reference has no tests/golden vectors of its own; these files are the pins.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from radtxfr_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **kw):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **kw)
    print("wrote %-28s %8d bytes" % (name, os.path.getsize(path)))


def g_wofz():
    h = R.hapi()
    rng = np.random.default_rng(100)
    x = np.concatenate([[0, 0.5, 1, 3, 7.5, 14, -2, 20, 100, 2000], rng.uniform(-30, 30, 400),
                        rng.uniform(-16, 16, 200)])
    y = np.concatenate([[1e-3, 0.5, 1, 0.1, 2, 0.5, 1e-5, 0.3, 40, 40], 10 ** rng.uniform(-6, 1.7, 400),
                        10 ** rng.uniform(-3, 1.2, 200)])
    wr, wi = h.hum1_wei(x.copy(), y.copy())
    # Weideman coefficients as the reference derives them (misc/hapi.py:9816-9824)
    N = 24; M = 2 * N; M2 = 2 * M
    k = h.arange(-M + 1, M); L = h.sqrt(N / h.sqrt(2)); th = k * h.pi / M; t = L * h.tan(th / 2)
    f = h.zeros(len(t) + 1); f[1:] = h.exp(-t ** 2) * (L ** 2 + t ** 2)
    a = h.flipud((h.real(h.fft(h.fftshift(f))) / M2)[1:N + 1])
    save("wofz.npz", source="misc/hapi.py:9833-9844 hum1_wei", x=x, y=y, wr=wr, wi=wi, weideman_a=a, weideman_L=L)


def g_profile():
    h = R.hapi()
    cases = [(1000.0, 0.0014544, 0.07), (1000.0, 0.0014544, 0.0007), (800.3, 0.0011, 0.012),
             (1299.9, 0.002, 1e-5), (950.0, 0.0013, 0.035)]
    sgs, vals = [], []
    for sg0, gd, g0 in cases:
        sg = np.sort(np.concatenate([sg0 + np.linspace(-50 * max(gd, g0), 50 * max(gd, g0), 301),
                                     [sg0, sg0 + 1e-4, sg0 - 0.01]]))
        sgs.append(sg)
        vals.append(h.PROFILE_VOIGT(sg0, gd, g0, sg)[0])
    save("profile_voigt.npz", source="misc/hapi.py:10131-10140 PROFILE_VOIGT", cases=np.array(cases),
         sg=np.array(sgs), V=np.array(vals))


def g_tips():
    h = R.hapi()
    mi = [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (4, 1), (5, 1),
          (6, 1), (7, 1), (22, 1)]
    Ts = np.array([70.0, 84.9, 85.0, 100.0, 180.3, 216.7, 250.0, 287.87, 296.0, 330.0, 1000.0, 2985.0, 2990.0, 3000.0])
    Q = np.array([[h.PYTIPS(m, i, float(T)) for T in Ts] for (m, i) in mi])
    S = h.EnvironmentDependency_Intensity(1e-22, 250., 296., h.PYTIPS(1, 1, 250.), h.PYTIPS(1, 1, 296.), 1500., 1000.)
    m = h.molecularMass(1, 1) * 1.66053873e-27 * 1000
    gd = h.sqrt(2 * h.cBolts * 296. * h.log(2) / m / h.cc ** 2) * 1000.
    save("tips.npz", source="misc/hapi.py:10030 PYTIPS; :10169 EnvironmentDependency_Intensity", mi=np.array(mi),
         T=Ts, Q=Q, S_kat=S, gammaD_kat=gd, volconc_kat=h.volumeConcentration(1.0, 296.0))


XS_CASES = [
    # name, seed, nlines, grid(lo,hi,n), T, p, kwargs
    ("base", 10, 300, (800.0, 1300.0, 10001), 296.0, 1.0, {}),
    ("cold_low", 11, 300, (800.0, 1300.0, 10001), 220.0, 0.01, {}),
    ("mid", 12, 300, (900.0, 1000.0, 20001), 250.0, 0.3, {}),
    ("top", 13, 200, (990.0, 1010.0, 20001), 225.53, 7.30 / 101325.0, {}),
    ("self", 14, 200, (800.0, 1300.0, 5001), 296.0, 0.5, {"GammaL": "gamma_self"}),
    ("mix", 15, 200, (800.0, 1300.0, 5001), 280.0, 0.8, {"Diluent": {"air": 0.7, "self": 0.3}}),
    ("nonhitran", 16, 150, (800.0, 1300.0, 5001), 296.0, 1.0, {"HITRAN_units": False}),
    ("wing_abs", 17, 150, (800.0, 1300.0, 5001), 296.0, 0.2, {"OmegaWing": 10.0}),
    ("wing_hw", 18, 150, (800.0, 1300.0, 5001), 296.0, 1.0, {"OmegaWingHW": 5.0}),
    ("thresh", 19, 300, (800.0, 1300.0, 5001), 296.0, 1.0, {"IntensityThreshold": 1e-24}),
    ("components", 20, 300, (800.0, 1300.0, 5001), 296.0, 1.0, {"Components": [(1, 1), (1, 3, 0.5)]}),
]


def g_xsect():
    out = {"source": "misc/hapi.py:10906-11141 absorptionCoefficient_Voigt", "names": np.array([c[0] for c in XS_CASES])}
    for name, seed, n, (lo, hi, npts), T, p, kw in XS_CASES:
        cols = syn.synthetic_lines(seed, n, lo - 25.0, hi + 25.0, 1)
        grid = np.linspace(lo, hi, npts)
        _, xs = R.ref_xsect(cols, T, p, grid, **kw)
        out["xs_" + name] = xs
    save("xsect.npz", **out)


def tud_inputs():
    atm = syn.standard_atmosphere()
    tabs = {1: syn.synthetic_lines(30, 40, 690.0, 760.0, 1),
            2: syn.synthetic_lines(31, 60, 690.0, 760.0, 2),
            3: syn.synthetic_lines(32, 60, 690.0, 760.0, 3, (1, 2, 3), (0.8, 0.1, 0.1))}
    return atm, tabs


def g_tud():
    atm, tabs = tud_inputs()
    kw = dict(Zs=atm["Zs"], Ts=atm["Ts"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_VAL=atm["MFs_VAL"], MFs_ID=atm["MFs_ID"],
              DVOUT=0.05)
    X, tau, Lu, Ld = R.ref_tud(tabs, 700.0, 750.0, **kw)
    X2, tau2, Lu2, Ld2 = R.ref_tud(tabs, 700.0, 750.0, Altitudes=np.array([1.0, 10.0, 500.0, 3.0]), N_angle=7, **kw)
    X3, tau3, Lu3, Ld3 = R.ref_tud(tabs, 700.0, 750.0, returnOD=True, theta_r=0.6, N_angle=4, **kw)
    save("tud.npz", source="radiative_transfer.py:274-392 compute_TUD (+App. B.1 OD wiring via hapi)",
         X=X, tau=tau, Lu=Lu, Ld=Ld, tau_alt=tau2, Lu_alt=Lu2, Ld_alt=Ld2, tau_od=tau3, Lu_od=Lu3, Ld_od=Ld3)


def g_sensor():
    r = R.rt()
    rng = np.random.default_rng(40)
    X = np.linspace(700.0, 1400.0, 7001)
    Y1 = rng.uniform(0, 1, X.size)
    Y2 = rng.uniform(0, 50, (X.size, 5))
    Xo, t1 = r.ILS_MAKO(X, Y1)
    _, t2 = r.ILS_MAKO(X, Y2)
    Xo_r, t3 = r.ILS_MAKO(X, Y2, resFactor=2, fwhm_sf=0.8, shift=0.3, scale=1.001)
    Xc = np.linspace(800.0, 1200.0, 4001)
    Xo_c, t4 = r.ILS_MAKO(Xc, Y2[:4001])
    g = R.ils_gauss()
    Xg, g1 = g.ILS_MAKO(X, Y1)
    _, g2 = g.ILS_MAKO(X, Y2)
    nE, nA = 6, 4
    Xb = Xo
    emis = syn.synthetic_emissivities(41, nE, Xb)
    Ts = rng.uniform(270, 320, nA)
    tau = rng.uniform(0.2, 1.0, (Xb.size, nA)); La = rng.uniform(0, 300, (Xb.size, nA)); Ld = rng.uniform(0, 600, (Xb.size, nA))
    dT = np.arange(-10, 10.5, 2.5)
    L0 = r.compute_LWIR_apparent_radiance(Xb, emis, Ts, tau, La, Ld)
    L1, Ls1 = r.compute_LWIR_apparent_radiance(Xb, emis, Ts, tau, La, Ld, dT=dT, return_Ls=True)
    Pl = r.planckian(np.linspace(700, 1500, 8001), np.array([280.0, 300.0]))
    Pw = r.planckian(np.linspace(7.5, 13.5, 61), np.array([[280.0, 300.0], [310., 250.]]), wavelength=True)
    save("sensor.npz", source="radiative_transfer.py:1072-1263 ILS_MAKO; ILS_MAKO.py:2-35; :1017-1069; :792-848",
         X=X, Y1=Y1, Y2=Y2, Xo=Xo, tri1=t1, tri2=t2, Xo_res=Xo_r, tri_res=t3, Xc=Xc, Xo_clip=Xo_c, tri_clip=t4,
         Xg=Xg, gauss1=g1, gauss2=g2, emis=emis, Ts=Ts, tau=tau, La=La, Ld=Ld, dT=dT, L0=L0, L1=L1, Ls1=Ls1,
         planck_kat=r.planckian(500., 296.), planck_grid=Pl, planck_um=Pw)


def g_sensor_f32():
    """The float32 call of the production script (Compute_LWIR_Apparent_Radiance.py:9-25): every input cast to float32,
    dT = arange(-10, 10.5, 0.5); numpy then evaluates the whole expression in float32."""
    r = R.rt()
    g = np.load(os.path.join(OUT, "sensor.npz"))
    X, emis, Ts, tau, La, Ld = (g[k].astype(np.float32) for k in ("Xo", "emis", "Ts", "tau", "La", "Ld"))
    dT = np.arange(-10, 10.5, 0.5).astype(np.float32)
    L = r.compute_LWIR_apparent_radiance(X, emis, Ts, tau, La, Ld, dT)
    L0 = r.compute_LWIR_apparent_radiance(X, emis, Ts, tau, La, Ld)
    assert L.dtype == np.float32 and L0.dtype == np.float32
    save("sensor_f32.npz", source="radiative_transfer.py:1017-1069 called as Compute_LWIR_Apparent_Radiance.py:9-25 does (float32)",
         X=X, emis=emis, Ts=Ts, tau=tau, La=La, Ld=Ld, dT=dT, L=L, L0=L0)


def g_bt():
    r = R.rt()
    rng = np.random.default_rng(60)
    X = np.linspace(700.0, 1400.0, 141)
    T = rng.uniform(220, 330, (141, 3, 2))
    L = r.BT2L(X, T)
    Lb = L.copy(); Lb[3, 0, 0] = -1.0; Lb[4, 1, 1] = np.nan; Lb[5, 2, 0] = 0.0
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bt = r.brightnessTemperature(X, Lb, bad_value=-5.0)
        bt_vec = r.brightnessTemperature(X, L[:, 0, 0])
        Tb = T.copy(); Tb[7, 0, 1] = -3.0
        L_bad = r.BT2L(X, Tb)
        Xu = np.linspace(7.5, 13.5, 61)
        Lu = r.planckian(Xu, np.array([280.0, 310.0]), wavelength=True)
        bt_um = r.brightnessTemperature(Xu, Lu, wavelength=True)
        bt_sd = r.brightnessTemperature(X, np.swapaxes(L[:, :, 0], 0, 1), spectral_dim=1)
    save("bt.npz", source="radiative_transfer.py:851-1014 brightnessTemperature, BT2L", X=X, T=T, L=L, Lb=Lb, bt=bt,
         bt_vec=bt_vec, Tb=Tb, L_bad=L_bad, Xu=Xu, Lu=Lu, bt_um=bt_um, bt_sd=bt_sd)


def g_reduce():
    np.int = int      # the reference needs the alias numpy >= 1.24 removed (radiative_transfer.py:1329)
    r = R.rt()
    rng = np.random.default_rng(70)
    X = np.linspace(690.0, 1410.0, 28800)
    Y = np.cumsum(rng.normal(0, 1, (X.size, 3)), axis=0) * 0.01 + np.sin(X / 7.0)[:, None]
    Xo, Yo = r.reduceResolution(X, Y, 0.25)
    Xo5, Yo5 = r.reduceResolution(X, Y[:, 1], 0.1, N=2, window="hamming")
    s11 = r.smooth(Y[:, 0], 11)
    s10f = r.smooth(Y[:, 2], 10, "flat")
    save("reduce.npz", source="radiative_transfer.py:1266-1350 smooth, reduceResolution (with np.int = int)", seed=70,
         Xo=Xo, Yo=Yo, Xo5=Xo5, Yo5=Yo5, s11=s11, s10f=s10f)


def g_reduce2():
    """reduceResolution near and beyond the ends of the data, where the spline's not-a-knot end conditions matter:
    the smallest smoothing factor the reference smooths with (3), and a caller-supplied X_out that spans the whole
    input range and extrapolates beyond it."""
    np.int = int
    r = R.rt()
    rng = np.random.default_rng(71)
    X = np.linspace(700.0, 760.0, 2401)                       # dX_in = 0.025
    Y = np.cumsum(rng.normal(0, 1, (X.size, 2)), axis=0) * 0.02 + np.cos(X / 3.0)[:, None]
    Xo3, Yo3 = r.reduceResolution(X, Y, 0.075)                # smFactor = 3
    Xfull = np.concatenate([[699.2, 699.9, 700.0, 700.01, 700.03, 700.05, 700.3], np.linspace(701.0, 759.0, 117),
                            [759.7, 759.95, 759.99, 760.0, 760.4]])
    Yfull = r.reduceResolution(X, Y, 0.25, X_out=Xfull)       # smFactor = 10, X_out from outside to outside
    Xs = np.linspace(700.0, 701.0, 12)                        # fewer knots than one window: both ends not-a-knot
    Ys = np.sin(Xs * 9.0) + Xs
    Yshort = r.reduceResolution(Xs, Ys, 0.3, X_out=np.linspace(699.9, 701.1, 41))
    save("reduce2.npz", source="radiative_transfer.py:1326-1350 reduceResolution (with np.int = int)", seed=71,
         Xo3=Xo3, Yo3=Yo3, Xfull=Xfull, Yfull=Yfull, Yshort=Yshort)


def optcols_table():
    rng = np.random.default_rng(9)
    cols = syn.synthetic_lines(56, 300, 1180.0, 1320.0, 1)
    n = 300
    cols["n_self"] = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0.3, 0.9, n))
    cols["delta_self"] = rng.uniform(-0.02, 0.01, n)
    cols["deltap_air"] = rng.uniform(-2e-5, 2e-5, n)
    cols["deltap_self"] = rng.uniform(-5e-5, 5e-5, n)
    return cols


OPT_CASES = [(250.0, 0.4, {"air": 1.0}), (310.0, 0.9, {"self": 1.0}), (280.0, 0.05, {"air": 0.6, "self": 0.4})]


def g_optcols():
    cols = optcols_table()
    grid = np.linspace(1200.0, 1300.0, 20001)
    out = {}
    for k, (T, p, dil) in enumerate(OPT_CASES):
        _, out["xs%d" % k] = R.ref_xsect(cols, T, p, grid, Diluent=dil)
    save("optcols.npz", source="misc/hapi.py:10906-11141 with n_self/delta_self/deltap_* columns and diluent mixes", **out)


def sdv_table():
    """300 H2O lines with speed-dependence columns for air and self, and a third broadening species (H2)."""
    rng = np.random.default_rng(19)
    cols = syn.synthetic_lines(58, 300, 1180.0, 1320.0, 1)
    n = 300
    cols["SD_air"] = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(0.05, 0.2, n))      # a fifth of the lines: no SD
    cols["SD_self"] = rng.uniform(0.05, 0.15, n)
    cols["n_self"] = rng.uniform(0.3, 0.9, n)
    cols["delta_self"] = rng.uniform(-0.02, 0.01, n)
    cols["gamma_h2"] = rng.uniform(0.03, 0.12, n)
    cols["n_h2"] = rng.uniform(0.3, 0.7, n)
    cols["delta_h2"] = rng.uniform(-0.01, 0.01, n)
    cols["deltap_h2"] = rng.uniform(-2e-5, 2e-5, n)
    cols["gamma_he"] = rng.uniform(0.01, 0.05, n)                                       # no n_he column -> n_air
    return cols


SDV_CASES = [(296.0, 1.0, {"air": 1.0}), (250.0, 0.3, {"air": 0.7, "self": 0.3}), (230.0, 0.02, {"air": 1.0}),
             (280.0, 0.5, {"air": 0.5, "h2": 0.3, "he": 0.2})]
DIL_CASES = [(296.0, 1.0, {"h2": 1.0}), (260.0, 0.4, {"air": 0.5, "h2": 0.3, "he": 0.15, "co2": 0.05})]


def g_sdvoigt():
    cols = sdv_table()
    grid = np.linspace(1200.0, 1300.0, 20001)
    out = {}
    for k, (T, p, dil) in enumerate(SDV_CASES):
        _, out["sd%d" % k] = R.ref_xsect_sd(cols, T, p, grid, Diluent=dil)
    for k, (T, p, dil) in enumerate(DIL_CASES):
        _, out["dil%d" % k] = R.ref_xsect(cols, T, p, grid, Diluent=dil)
    save("sdvoigt.npz", source="misc/hapi.py:10657-10905 absorptionCoefficient_SDVoigt with SD_<diluent> columns (pcqsdhc "
         "PART2-4, :9923-10020); :10906-11141 with gamma_h2 / gamma_he diluents", **out)


def widewing_table():
    cols = syn.synthetic_lines(61, 300, 960.0, 1040.0, 1)
    cols["sw"][::41] *= 1e6
    return cols


WIDE_GRID = (990.0, 1010.0, 40001)            # 0.0005 cm^-1: 19 full 2048-point tiles and a partial one
WIDE_SD_GRID = (1240.0, 1260.0, 40001)
WIDE_SD_CASE = (280.0, 0.8, {"air": 0.7, "h2": 0.3})


def g_widewing():
    """Wings of thousands of grid points -- the regime of the tile-level far-field blocks: absorptionCoefficient_Voigt with
    OmegaWing = 8 cm^-1 and with the reference script's WavenumberWingHW = 350 (misc/RT_gen_AbsXS_files.py:18,90), and
    absorptionCoefficient_SDVoigt with that wing on the speed-dependent table."""
    cols = widewing_table()
    grid = np.linspace(*WIDE_GRID)
    out = {}
    _, out["voigt_w8"] = R.ref_xsect(cols, 296.0, 1.0, grid, OmegaWing=8.0)
    _, out["voigt_hw350"] = R.ref_xsect(cols, 250.0, 0.3, grid, WavenumberWingHW=350.0)
    T, p, dil = WIDE_SD_CASE
    _, out["sd_hw350"] = R.ref_xsect_sd(sdv_table(), T, p, np.linspace(*WIDE_SD_GRID), Diluent=dil, WavenumberWingHW=350.0)
    save("widewing.npz", source="misc/hapi.py:10906-11141 (OmegaWing=8, WavenumberWingHW=350) and :10657-10905 (SDVoigt, "
         "WavenumberWingHW=350) on fine grids: windows of 16 000 grid points and more", **out)


def g_lut():
    """SDVoigt on a table without SD columns (the reference's own call convention, misc/RT_gen_AbsXS_files.py:90)
    and the AFIT_XS writer, whose source is lifted out of the script with ast (the script itself fetches HITRAN
    from the network at import)."""
    import ast
    import tempfile
    h = R.hapi()
    cols = syn.synthetic_lines(80, 250, 1950.0, 2400.0, 2, (1, 2, 3), (0.8, 0.1, 0.1))
    R.inject_table("__lut__", cols)
    grid = np.linspace(2000.0, 2350.0, 7001)
    out = {}
    import contextlib, io
    for k, (T, p) in enumerate([(296.0, 1.0), (250.0, 0.1), (330.0, 0.5)]):
        with contextlib.redirect_stdout(io.StringIO()):
            _, xs = h.absorptionCoefficient_SDVoigt(SourceTables="__lut__", HITRAN_units=True, Environment={"T": T, "p": p},
                                                    OmegaGrid=grid, IntensityThreshold=0, WavenumberWingHW=350.0)
        out["xs%d" % k] = xs
    src = open(os.path.join(R.REF_ROOT, "misc", "RT_gen_AbsXS_files.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "AFIT_XS_write"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "AFIT_XS_write", "exec"), ns)
    with tempfile.TemporaryDirectory() as d:
        f = ns["AFIT_XS_write"](grid, out["xs0"], 296.0, 101325 * 1.0, 2, "HITRAN2016-synthetic", File=os.path.join(d, "x.bin"))
        out["afit_bytes"] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
    save("lut.npz", source="misc/hapi.py:10657 absorptionCoefficient_SDVoigt; misc/RT_gen_AbsXS_files.py:45-83", **out)


LORENTZ_CASES = [(296.0, 1.0, {}), (250.0, 0.2, {"OmegaWingHW": 25.0}), (310.0, 0.6, {"Diluent": {"air": 0.5, "self": 0.5}, "HITRAN_units": False})]


def g_lorentz():
    h = R.hapi()
    import contextlib, io
    cols = syn.synthetic_lines(90, 300, 1480.0, 1620.0, 1)
    R.inject_table("__lor__", cols)
    grid = np.linspace(1500.0, 1600.0, 10001)
    out = {}
    for k, (T, p, kw) in enumerate(LORENTZ_CASES):
        with contextlib.redirect_stdout(io.StringIO()):
            _, out["xs%d" % k] = h.absorptionCoefficient_Lorentz(SourceTables="__lor__", Environment={"T": T, "p": p}, OmegaGrid=grid, **kw)
    save("lorentz.npz", source="misc/hapi.py:11144-11381 absorptionCoefficient_Lorentz", **out)


DOPPLER_CASES = [(296.0, 1.0, {}), (220.0, 0.05, {"OmegaWingHW": 30.0, "LineShift": False}),
                 (310.0, 0.6, {"HITRAN_units": False, "OmegaWing": 0.02})]


def g_doppler():
    h = R.hapi()
    import contextlib, io
    cols = syn.synthetic_lines(91, 80, 1495.0, 1525.0, 1)
    R.inject_table("__dop__", cols)
    grid = np.linspace(1500.0, 1520.0, 20001)
    out = {}
    for k, (T, p, kw) in enumerate(DOPPLER_CASES):
        with contextlib.redirect_stdout(io.StringIO()):
            _, xs = h.absorptionCoefficient_Doppler(SourceTables="__dop__", Environment={"T": T, "p": p}, OmegaGrid=grid, **kw)
        nz = np.nonzero(xs)[0]
        out["idx%d" % k] = nz.astype(np.int32)
        out["xs%d" % k] = xs[nz]
    save("doppler.npz", source="misc/hapi.py:11384-11550 absorptionCoefficient_Doppler (non-zero samples of a 20001-point grid)", **out)


def g_spectra():
    """transmittanceSpectrum / absorptionSpectrum / radianceSpectrum of hapi on a fixed coefficient vector."""
    h = R.hapi()
    nu = np.linspace(900.0, 1100.0, 41)
    coef = 1e-3 * (1.0 + np.sin(nu / 7.0) ** 2)
    env = {"l": 250.0, "T": 283.0}
    out = {"nu": nu, "coef": coef}
    out["trans"] = h.transmittanceSpectrum(nu, coef, Environment=env)[1]
    out["absorp"] = h.absorptionSpectrum(nu, coef, Environment=env)[1]
    out["radi"] = h.radianceSpectrum(nu, coef, Environment=env)[1]
    save("spectra.npz", source="misc/hapi.py:11582-11681", **out)


def mbi_image():
    ii, jj = np.meshgrid(np.arange(40), np.arange(32), indexing="ij")
    img = np.zeros((3, 40, 32), dtype=np.uint8)
    img[0] = (ii ** 2 + jj ** 2) < 16 ** 2
    img[1] = (abs(ii - 20) > 0) & (abs(jj - 16) > 0)
    img[2] = (ii > 20) & (jj < 16)
    return img


def g_mbi():
    """AFIT_MBI_export of the reference (AFIT_MBI.py:45-156); its functions are lifted out with ast because the
    module runs a self-test (and imports matplotlib) at import time."""
    import ast
    import tempfile
    src = open(os.path.join(R.REF_ROOT, "AFIT_MBI.py")).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    import pickle, sys as _sys
    import scipy.io as sio
    ns = {"np": np, "os": os, "sio": sio, "pickle": pickle, "tempfile": tempfile, "sys": _sys}
    exec(compile(ast.Module(body=body, type_ignores=[]), "AFIT_MBI", "exec"), ns)
    img = mbi_image()
    cube = np.random.default_rng(5).uniform(0, 100, (4, 6, 5))
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            ns["AFIT_MBI_export"]("a.bsq", img, Verbose=False, Q=np.array([1, 2, 3]), Meta1=2, myString="hello")
            ns["AFIT_MBI_export"]("b.bip", img, Verbose=False)
            ns["AFIT_MBI_export"]("c.bip", cube, Verbose=False, Bands=np.array([7.5, 8.5, 9.5, 10.5]))
            for k in ("a.bsq", "b.bip", "c.bip"):
                out[k.replace(".", "_")] = np.frombuffer(open(k, "rb").read(), dtype=np.uint8)
        finally:
            os.chdir(cwd)
    save("mbi.npz", source="AFIT_MBI.py:45-156 AFIT_MBI_export", cube=cube, **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["wofz", "profile", "tips", "xsect", "tud", "sensor", "bt", "reduce", "lut", "optcols", "mbi", "lorentz", "doppler", "spectra", "sensor_f32", "reduce2", "sdvoigt", "widewing"]
    for w in which:
        globals()["g_" + w]()
