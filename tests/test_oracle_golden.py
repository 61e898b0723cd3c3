"""Pin the CPU oracle (oracle/port.py) to golden vectors minted from the unmodified reference
(tools/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, assert_parity

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
from oracle import port  # noqa: E402
from radtxfr_b200 import synthetic as syn  # noqa: E402

TIGHT = dict(rtol=1e-13, floor=1e-15)


def test_weideman_coefficients_and_wofz():
    g = golden("wofz.npz")
    a, L = port.weideman_coeffs(24)
    assert L == float(g["weideman_L"]) == 4.119534287814235
    assert np.array_equal(a, g["weideman_a"])
    wr, wi = port.hum1_wei(g["x"], g["y"])
    assert np.array_equal(wr, g["wr"]) and np.array_equal(wi, g["wi"])
    # SURVEY.md App. D spot values
    assert wr[1] == 0.5331567079123801 and wi[1] == 0.2304882313844082
    assert wr[7] == 0.0004246359021385162


def test_profile_voigt():
    g = golden("profile_voigt.npz")
    for (sg0, gd, g0), sg, V in zip(g["cases"], g["sg"], g["V"]):
        assert_parity(port.profile_voigt(sg0, gd, g0, sg), V, what="PROFILE_VOIGT", **TIGHT)
    v = port.profile_voigt(1000.0, 0.0014544, 0.07, np.array([999, 999.99, 1000, 1000.5]))
    assert_parity(v, [0.02217314446367562, 4.4550875777944, 4.5458685127508485, 0.0874149938033936], **TIGHT)


def test_tips_and_line_kats():
    g = golden("tips.npz")
    for (m, i), row in zip(g["mi"], g["Q"]):
        for T, q in zip(g["T"], row):
            assert port.tips_q(int(m), int(i), float(T)) == q
    assert port.tips_q(1, 1, 296.0) == 174.63820437890627
    with pytest.raises(Exception):
        port.tips_q(1, 1, 69.9)
    with pytest.raises(Exception):
        port.tips_q(1, 1, 3000.1)
    S = port.line_strength(1e-22, 250.0, port.tips_q(1, 1, 250.0), port.tips_q(1, 1, 296.0), 1500.0, 1000.0)
    assert S == float(g["S_kat"])
    assert port.doppler_hwhm(1000.0, 296.0, 18.010565) == float(g["gammaD_kat"])
    assert port.volume_concentration(1.0, 296.0) == float(g["volconc_kat"]) == 2.479371831717524e19


def test_xsect_cases():
    import make_golden as mg
    g = golden("xsect.npz")
    for name, seed, n, (lo, hi, npts), T, p, kw in mg.XS_CASES:
        cols = syn.synthetic_lines(seed, n, lo - 25.0, hi + 25.0, 1)
        grid = np.linspace(lo, hi, npts)
        om, xs = port.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=grid, **kw)
        assert np.array_equal(om, grid)
        assert_parity(xs, g["xs_" + name], what="xsect " + name, **TIGHT)


def _tud_kw():
    import make_golden as mg
    atm, tabs = mg.tud_inputs()
    return tabs, dict(Zs=atm["Zs"], Ts=atm["Ts"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_VAL=atm["MFs_VAL"],
                      MFs_ID=atm["MFs_ID"], DVOUT=0.05)


def test_tud():
    g = golden("tud.npz")
    tabs, kw = _tud_kw()
    X, tau, Lu, Ld, OD = port.compute_tud(tabs, 700.0, 750.0, return_od_matrix=True, **kw)
    assert np.array_equal(X, g["X"])
    for a, b, w in ((tau, g["tau"], "tau"), (Lu, g["Lu"], "Lu"), (Ld, g["Ld"], "Ld")):
        assert_parity(a, b, what=w, **TIGHT)
    assert tau.min() < 0.1 and 0.2 < np.median(tau) < 0.8 and tau.max() <= 1.0   # optically interesting column
    Z, T = kw["Zs"], kw["Ts"]
    t2, u2, d2 = port.tud_from_od(X, OD, Z, T, N_angle=7, Altitudes=np.array([1.0, 10.0, 500.0, 3.0]))
    assert t2.shape == (X.size, 4)
    for a, b, w in ((t2, g["tau_alt"], "tau_alt"), (u2, g["Lu_alt"], "Lu_alt"), (d2, g["Ld_alt"], "Ld_alt")):
        assert_parity(a, b, what=w, **TIGHT)
    t3, u3, d3 = port.tud_from_od(X, OD, Z, T, N_angle=4, theta_r=0.6, returnOD=True)
    for a, b, w in ((t3, g["tau_od"], "tau_od"), (u3, g["Lu_od"], "Lu_od"), (d3, g["Ld_od"], "Ld_od")):
        assert_parity(a, b, what=w, **TIGHT)


def test_sensor_model():
    g = golden("sensor.npz")
    X, Y1, Y2 = g["X"], g["Y1"], g["Y2"]
    Xo, t1 = port.ils_mako_triangle(X, Y1)
    assert np.array_equal(Xo, g["Xo"]) and Xo.size == 128
    assert_parity(t1, g["tri1"], what="tri 1-D", **TIGHT)
    assert_parity(port.ils_mako_triangle(X, Y2, returnX=False), g["tri2"], what="tri 2-D", **TIGHT)
    Xr, t3 = port.ils_mako_triangle(X, Y2, resFactor=2, fwhm_sf=0.8, shift=0.3, scale=1.001)
    assert np.array_equal(Xr, g["Xo_res"]) and Xr.size == 256
    assert_parity(t3, g["tri_res"], what="tri resFactor", **TIGHT)
    Xc, t4 = port.ils_mako_triangle(g["Xc"], Y2[:4001])
    assert np.array_equal(Xc, g["Xo_clip"]) and Xc.size < 128
    assert_parity(t4, g["tri_clip"], what="tri clipped", **TIGHT)
    Xg, g1 = port.ils_mako_gauss(X, Y1)
    assert np.array_equal(Xg, g["Xg"])
    assert_parity(g1, g["gauss1"], what="gauss 1-D", **TIGHT)
    assert_parity(port.ils_mako_gauss(X, Y2)[1], g["gauss2"], what="gauss 2-D", **TIGHT)
    args = (g["Xo"], g["emis"], g["Ts"], g["tau"], g["La"], g["Ld"])
    assert np.array_equal(port.apparent_radiance(*args), g["L0"])
    L1, Ls1 = port.apparent_radiance(*args, dT=g["dT"], return_Ls=True)
    assert np.array_equal(L1, g["L1"]) and np.array_equal(Ls1, g["Ls1"])
    assert abs(port.planckian(500.0, 296.0)[0] - 14.36645756) < 5e-9
    assert np.array_equal(port.planckian(np.linspace(700, 1500, 8001), np.array([280.0, 300.0])), g["planck_grid"])
    assert np.array_equal(port.planckian(np.linspace(7.5, 13.5, 61), np.array([[280.0, 300.0], [310., 250.]]),
                                         wavelength=True), g["planck_um"])


def test_brightness_temperature():
    g = golden("bt.npz")
    X, T, L = g["X"], g["T"], g["L"]
    assert np.array_equal(port.bt2l(X, T), L)
    bt = port.brightness_temperature(X, g["Lb"], bad_value=-5.0)
    assert_parity(bt, g["bt"], what="bt", **TIGHT)
    assert np.array_equal(port.brightness_temperature(X, L[:, 0, 0]), g["bt_vec"])
    lb = port.bt2l(X, g["Tb"])
    assert np.array_equal(np.isnan(lb), np.isnan(g["L_bad"]))
    assert np.array_equal(port.brightness_temperature(g["Xu"], g["Lu"], wavelength=True), g["bt_um"])
    assert np.array_equal(port.brightness_temperature(X, np.swapaxes(L[:, :, 0], 0, 1), spectral_dim=1), g["bt_sd"])


def _reduce_inputs():
    rng = np.random.default_rng(70)
    X = np.linspace(690.0, 1410.0, 28800)
    Y = np.cumsum(rng.normal(0, 1, (X.size, 3)), axis=0) * 0.01 + np.sin(X / 7.0)[:, None]
    return X, Y


def test_reduce_resolution():
    g = golden("reduce.npz")
    X, Y = _reduce_inputs()
    Xo, Yo = port.reduce_resolution(X, Y, 0.25)
    assert np.array_equal(Xo, g["Xo"]) and np.array_equal(Yo, g["Yo"])
    Xo5, Yo5 = port.reduce_resolution(X, Y[:, 1], 0.1, N=2, window="hamming")
    assert np.array_equal(Xo5, g["Xo5"]) and np.array_equal(Yo5, g["Yo5"])
    assert np.array_equal(port.smooth(Y[:, 0], 11), g["s11"])
    assert np.array_equal(port.smooth(Y[:, 2], 10, "flat"), g["s10f"])


def test_optional_columns():
    import make_golden as mg
    g = golden("optcols.npz")
    cols = mg.optcols_table()
    grid = np.linspace(1200.0, 1300.0, 20001)
    for k, (T, p, dil) in enumerate(mg.OPT_CASES):
        _, xs = port.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=grid, Diluent=dil)
        assert_parity(xs, g["xs%d" % k], what="optional columns %d" % k, **TIGHT)


def test_c_oracle_matches_numpy_port_and_golden():
    """The C/OpenMP restatement used for full-size checks (oracle/voigt_c.c) against the numpy port and, through
    tests/golden/xsect.npz, against the reference itself."""
    import make_golden as mg
    from oracle import cbuild
    g = golden("xsect.npz")
    for name, seed, n, (lo, hi, npts), T, p, kw in mg.XS_CASES:
        cols = syn.synthetic_lines(seed, n, lo - 25.0, hi + 25.0, 1)
        grid = np.linspace(lo, hi, npts)
        _, xs, ev = cbuild.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=grid, **kw)
        assert_parity(xs, g["xs_" + name], rtol=1e-11, floor=1e-14, what="C oracle " + name)
        _, _, ev_port, _ = port.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=grid, return_evals=True, **kw)
        assert ev == ev_port
    atm, tabs = mg.tud_inputs()
    X = port.make_spectral_axis(700.0, 750.0, 0.05)
    for L in (0, 30, 65):
        a = port.layer_od(tabs, X, atm["Ts"][L], atm["Ps"][L], atm["PLs"][L], atm["MFs_VAL"][L], atm["MFs_ID"])
        b = cbuild.layer_od(tabs, X, atm["Ts"][L], atm["Ps"][L], atm["PLs"][L], atm["MFs_VAL"][L], atm["MFs_ID"])
        assert_parity(b, a, rtol=1e-11, floor=1e-14, what="C oracle layer_od %d" % L)


def test_lorentz():
    import make_golden as mg
    g = golden("lorentz.npz")
    cols = syn.synthetic_lines(90, 300, 1480.0, 1620.0, 1)
    grid = np.linspace(1500.0, 1600.0, 10001)
    for k, (T, p, kw) in enumerate(mg.LORENTZ_CASES):
        _, xs = port.absorption_coefficient_lorentz(cols, T=T, p=p, OmegaGrid=grid, **kw)
        assert_parity(xs, g["xs%d" % k], what="Lorentz %d" % k, **TIGHT)


def test_doppler():
    import make_golden as mg
    g = golden("doppler.npz")
    cols = syn.synthetic_lines(91, 80, 1495.0, 1525.0, 1)
    grid = np.linspace(1500.0, 1520.0, 20001)
    for k, (T, p, kw) in enumerate(mg.DOPPLER_CASES):
        _, xs = port.absorption_coefficient_doppler(cols, T=T, p=p, OmegaGrid=grid, **kw)
        ref = np.zeros_like(xs); ref[g["idx%d" % k]] = g["xs%d" % k]
        assert_parity(xs, ref, what="Doppler %d" % k, **TIGHT)


def test_sdvoigt_and_diluents_port():
    """The port of pcqsdhc (all parts) / absorptionCoefficient_SDVoigt and the generic diluent loop against the reference's
    outputs for a table with SD_air / SD_self / gamma_h2 / gamma_he columns (tests/golden/sdvoigt.npz)."""
    import make_golden as mg
    g = golden("sdvoigt.npz")
    cols = mg.sdv_table()
    grid = np.linspace(1200.0, 1300.0, 20001)
    for k, (T, p, dil) in enumerate(mg.SDV_CASES[:2]):
        _, xs = port.absorption_coefficient_sdvoigt(cols, T=T, p=p, OmegaGrid=grid, Diluent=dil)
        assert_parity(xs, g["sd%d" % k], rtol=1e-11, floor=1e-15, what="SDVoigt port case %d" % k)
    T, p, dil = mg.DIL_CASES[1]
    _, xs = port.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=grid, Diluent=dil)
    assert_parity(xs, g["dil1"], what="diluent port", **TIGHT)
    # cpf3 against the reference's own series evaluated in complex arithmetic
    x, y = np.array([9.0, -7.5, 8.2]), np.array([1.0, 3.0, 0.05])
    z = x + 1j * y
    s, t = 1.0 + 0j, 1.0 + 0j
    for k in range(15):
        t = t * ((1 / z) ** 2 * (k + 0.5)); s = s + t
    ref = s * (1j / z * 0.564189583547756)
    wr, wi = port.cpf3(x, y)
    assert np.allclose(wr, ref.real, rtol=1e-14) and np.allclose(wi, ref.imag, rtol=1e-14)


def test_wide_wings_port():
    """The port on windows of 16 000 grid points and more (OmegaWing = 8 cm^-1, the reference script's
    WavenumberWingHW = 350) against the reference's outputs (tests/golden/widewing.npz)."""
    import make_golden as mg
    g = golden("widewing.npz")
    cols = mg.widewing_table()
    grid = np.linspace(*mg.WIDE_GRID)
    _, xs = port.absorption_coefficient_voigt(cols, T=296.0, p=1.0, OmegaGrid=grid, OmegaWing=8.0)
    assert_parity(xs, g["voigt_w8"], what="Voigt, OmegaWing=8", **TIGHT)
    _, xs = port.absorption_coefficient_voigt(cols, T=250.0, p=0.3, OmegaGrid=grid, OmegaWingHW=350.0)
    assert_parity(xs, g["voigt_hw350"], what="Voigt, WavenumberWingHW=350", **TIGHT)
    T, p, dil = mg.WIDE_SD_CASE
    _, xs = port.absorption_coefficient_sdvoigt(mg.sdv_table(), T=T, p=p, OmegaGrid=np.linspace(*mg.WIDE_SD_GRID), Diluent=dil,
                                                OmegaWingHW=350.0)
    assert_parity(xs, g["sd_hw350"], rtol=1e-10, floor=1e-15, what="SDVoigt, WavenumberWingHW=350")
