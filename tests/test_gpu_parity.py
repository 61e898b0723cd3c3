"""GPU parity tests: the CUDA path (through the Python shim -> ctypes -> C ABI) against the CPU oracle
(oracle/port.py) on the same seeded inputs, and against the golden vectors minted from the reference.

Tolerance (SURVEY.md 8c, FP64 path): |gpu-ref| <= 1e-9*|ref| + 1e-12*max|ref| per output array, and max relative
error <= 1e-9 over points with |ref| >= 1e-6*max|ref|.
"""
import os
import sys

import numpy as np
import pytest

from conftest import golden, assert_parity

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))

pytestmark = pytest.mark.gpu

RTOL, FLOOR = 1e-9, 1e-12


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from radtxfr_b200 import hapi_compat, radiative_transfer, ILS_MAKO, engine, synthetic
    from oracle import port
    return dict(torch=torch, hapi=hapi_compat, rt=radiative_transfer, ilsg=ILS_MAKO, engine=engine, syn=synthetic,
                port=port)


def test_library_is_native(mods):
    """The product path must be the in-tree CUDA library, on a B200."""
    from radtxfr_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH)
    h = mods["engine"].default_handle()
    sm, name = h.device_info()
    assert sm > 0 and name
    maps = open("/proc/self/maps").read()
    assert "libradtxfr_b200.so" in maps


def test_tips_device(mods):
    g = golden("tips.npz")
    for (m, i), row in zip(g["mi"], g["Q"]):
        for T, q in zip(g["T"], row):
            got = mods["hapi"].PYTIPS(int(m), int(i), float(T))
            assert abs(got - q) <= 1e-13 * abs(q), (m, i, T, got, q)
    with pytest.raises(Exception):
        mods["hapi"].PYTIPS(1, 1, 69.0)


def test_xsect_golden_cases(mods):
    """absorptionCoefficient_Voigt against the reference's own outputs (tests/golden/xsect.npz)."""
    import make_golden as mg
    hapi, syn = mods["hapi"], mods["syn"]
    g = golden("xsect.npz")
    for name, seed, n, (lo, hi, npts), T, p, kw in mg.XS_CASES:
        cols = syn.synthetic_lines(seed, n, lo - 25.0, hi + 25.0, 1)
        hapi.storeTable("t_" + name, cols)
        grid = np.linspace(lo, hi, npts)
        om, xs = hapi.absorptionCoefficient_Voigt(SourceTables="t_" + name, Environment={"T": T, "p": p},
                                                  OmegaGrid=grid, **kw)
        assert isinstance(xs, np.ndarray) and xs.dtype == np.float64 and np.array_equal(om, grid)
        assert_parity(xs, g["xs_" + name], RTOL, FLOOR, "xsect " + name)


@pytest.mark.parametrize("K", [1, 2, 4, 8])
@pytest.mark.parametrize("variant", [0, 128, 4096, 128 | 4096, 4, 4 | 128 | 4096])
def test_voigt_all_tile_sizes(mods, K, variant):
    """Every tile-size instantiation of the kernel; default arithmetic (one-Newton-step reciprocal + far-field blocks),
    RTX_VOIGT_EXACT_RCP (128), RTX_VOIGT_NOFAR (4096), both, plain and warp-specialised (flag 4), on a multi-state
    problem that mixes wide (1 atm) and narrow (top of atmosphere) lines."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(50, 400, 985.0, 1015.0, 1)
    db = engine.LineDB(cols)
    grid = np.linspace(990.0, 1010.0, 5003)
    states = [(296.0, 1.0), (250.0, 0.2), (225.53, 7.3 / 101325.0), (216.7, 0.02)]
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    out = db.voigt_sum(T, p, w, grid, flags=variant | (K << 8)).cpu().numpy()
    for k, (Tk, pk) in enumerate(states):
        _, ref = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid)
        assert_parity(out[k], ref, RTOL, FLOOR, "K=%d variant=%d state %d" % (K, variant, k))


@pytest.mark.parametrize("K", [4, 8])
def test_voigt_far_field_blocks(mods, K):
    """Far-field blocks (slices a window covers from >= 0.375 slice widths away are summed at 32 Chebyshev nodes and
    interpolated): on a fine grid where most evaluations take that path, the result agrees with the node-free
    evaluation (RTX_VOIGT_NOFAR | RTX_VOIGT_EXACT_RCP) to 5e-12 and with the oracle to the parity tolerance; worst
    cases: a strong isolated line next to weak ones (the window cut of the strong line must stay exact), a grid whose
    spacing changes by a factor of 3 half way (Clenshaw runs on the actual grid values), and a jittered grid."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(52, 300, 992.0, 1008.0, 1)
    cols["sw"][::37] *= 1e6                                   # a few lines 10^6 times stronger than their neighbours
    db = engine.LineDB(cols)
    states = [(296.0, 1.0), (260.0, 0.5), (230.0, 0.1)]
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    rng = np.random.default_rng(5)
    uniform = np.linspace(995.0, 1005.0, 20001)
    two_step = np.concatenate([np.linspace(995.0, 1000.0, 15001)[:-1], np.linspace(1000.0, 1005.0, 5001)])
    jitter = np.sort(uniform + rng.uniform(-2e-4, 2e-4, uniform.size))
    for name, grid in (("uniform", uniform), ("two_step", two_step), ("jitter", jitter)):
        far = db.voigt_sum(T, p, w, grid, flags=(K << 8)).cpu().numpy()
        fx = db.voigt_sum(T, p, w, grid, flags=128 | (K << 8)).cpu().numpy()
        exact = db.voigt_sum(T, p, w, grid, flags=128 | 4096 | (K << 8)).cpu().numpy()
        assert np.max(np.abs(fx - exact) / np.maximum(exact, 1e-300)) < 2e-13, name        # the interpolation alone
        assert np.max(np.abs(far - exact) / np.maximum(exact, 1e-300)) < 5e-12, name       # plus the one-step reciprocal
        assert not np.array_equal(far, db.voigt_sum(T, p, w, grid, flags=4096 | (K << 8)).cpu().numpy())  # path taken
        for k, (Tk, pk) in enumerate(states[:2]):
            _, ref = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid)
            assert_parity(far[k], ref, RTOL, FLOOR, "far K=%d %s state %d" % (K, name, k))


def test_voigt_fp32_option(mods):
    """Optional FP32 evaluation of the asymptote branch (north star: 1e-5 relative, floor 1e-8)."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(55, 600, 940.0, 1060.0, 1)
    db = engine.LineDB(cols)
    states = [(296.0, 1.0), (250.0, 0.2), (225.53, 7.3 / 101325.0), (216.7, 0.02)]
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    for grid in (np.linspace(950.0, 1050.0, 100001), np.linspace(950.0, 1050.0, 10001)):
        for K in (8, 1):
            out = db.voigt_sum(T, p, w, grid, flags=2 | (K << 8)).cpu().numpy()
            for k, (Tk, pk) in enumerate(states):
                _, ref = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid)
                assert_parity(out[k], ref, 1e-5, 1e-8, "fp32 K=%d n=%d state %d" % (K, grid.size, k))


def test_voigt_eval_counter(mods):
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(51, 300, 985.0, 1015.0, 1)
    db = engine.LineDB(cols)
    grid = np.linspace(990.0, 1010.0, 4001)
    tot = totw = 0
    states = [(296.0, 1.0), (230.0, 0.01), (225.53, 7.3 / 101325.0)]
    for Tk, pk in states:
        _, _, e, ew = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid, return_evals=True)
        tot += e; totw += ew
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    e, ew = db.count_evals(T, p, w, grid)
    assert (e, ew) == (tot, totw) and ew > 0


def test_voigt_edge_cases(mods):
    torch, port, syn, engine, hapi = (mods[k] for k in ("torch", "port", "syn", "engine", "hapi"))
    cols = syn.synthetic_lines(52, 100, 800.0, 900.0, 1)
    # unsorted table rows, lines far outside the grid, single-point and empty grids, non-uniform grid
    perm = np.random.default_rng(0).permutation(100)
    shuffled = {k: np.asarray(v)[perm] for k, v in cols.items()}
    hapi.storeTable("edge", shuffled)
    rng = np.random.default_rng(1)
    grid = np.sort(np.concatenate([rng.uniform(840, 860, 700), np.linspace(845, 846, 301)]))
    _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=grid, Environment={"T": 260.0, "p": 0.5})
    _, ref = port.absorption_coefficient_voigt(cols, T=260.0, p=0.5, OmegaGrid=grid)
    assert_parity(xs, ref, RTOL, FLOOR, "non-uniform grid")
    _, xs1 = hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=np.array([850.0]))
    _, ref1 = port.absorption_coefficient_voigt(cols, OmegaGrid=np.array([850.0]))
    assert_parity(xs1, ref1, RTOL, FLOOR, "one point")
    _, xs0 = hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=np.zeros(0))
    assert xs0.shape == (0,)
    far = np.linspace(2000.0, 2010.0, 101)
    _, xsf = hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=far)
    assert np.array_equal(xsf, np.zeros(101))
    # empty table
    empty = {k: np.asarray(v)[:0] for k, v in cols.items()}
    hapi.storeTable("empty", empty)
    _, xse = hapi.absorptionCoefficient_Voigt(SourceTables="empty", OmegaGrid=far)
    assert np.array_equal(xse, np.zeros(101))
    # missing table / bad GammaL
    with pytest.raises(Exception, match="no such table"):
        hapi.absorptionCoefficient_Voigt(SourceTables="nope", OmegaGrid=far)
    with pytest.raises(Exception, match="Unknown GammaL"):
        hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=far, GammaL="gamma_h2")
    with pytest.raises(Exception, match="TIPS"):
        hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=far, Environment={"T": 50.0, "p": 1.0})
    # misaligned device grid (8-byte but not 16-byte aligned) takes the non-TMA staging path
    db = engine.LineDB(cols)
    g = torch.from_numpy(np.concatenate([[0.0], np.linspace(840, 860, 3001)])).cuda()[1:]
    assert g.data_ptr() % 16 == 8
    w = db.iso_factors(None, 1.0)[None, :]
    o1 = db.voigt_sum(np.array([296.0]), np.array([1.0]), w, g).cpu().numpy()
    o2 = db.voigt_sum(np.array([296.0]), np.array([1.0]), w, g.clone()).cpu().numpy()
    assert np.array_equal(o1, o2)
    # tensors in -> tensors out
    om, xt = hapi.absorptionCoefficient_Voigt(SourceTables="edge", OmegaGrid=torch.from_numpy(grid).cuda(),
                                              Environment={"T": 260.0, "p": 0.5})
    assert xt.is_cuda and np.array_equal(xt.cpu().numpy(), xs)


def test_voigt_window_ties_and_range_route(mods):
    """bisect_right semantics on exact ties (grid point == nu-W is outside, == nu+W is inside, misc/hapi.py:11133-11134)
    and the OmegaRange/OmegaStep route with the default Environment."""
    port, hapi = mods["port"], mods["hapi"]
    cols = {"molec_id": np.array([1, 1, 1]), "local_iso_id": np.array([1, 1, 2]), "nu": np.array([1000.0, 1002.5, 1010.0]),
            "sw": np.array([1e-21, 3e-22, 2e-21]), "elower": np.array([100.0, 500.0, 0.0]),
            "gamma_air": np.array([0.07, 0.05, 0.09]), "gamma_self": np.array([0.3, 0.2, 0.4]),
            "n_air": np.array([0.7, 0.6, 0.5]), "delta_air": np.array([0.0, -0.003, 0.002])}
    hapi.storeTable("ties", cols)
    grid = np.arange(9900, 10200) / 10.0                       # 990.0 .. 1019.9, exact tenths where representable
    assert 997.5 in grid and 1002.5 in grid
    _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="ties", OmegaGrid=grid, OmegaWing=2.5, OmegaWingHW=0.0)
    _, ref = port.absorption_coefficient_voigt(cols, OmegaGrid=grid, OmegaWing=2.5, OmegaWingHW=0.0)
    assert_parity(xs, ref, RTOL, FLOOR, "ties")
    assert np.array_equal(xs == 0.0, ref == 0.0)               # identical window membership, point by point
    i_lo, i_hi = int(np.where(grid == 997.5)[0][0]), int(np.where(grid == 1002.5)[0][0])
    assert ref[i_lo] == 0.0 and ref[i_hi] > 0.0
    om, xs2 = hapi.absorptionCoefficient_Voigt(SourceTables="ties", OmegaRange=(995.0, 1015.0), OmegaStep=0.01)
    assert om.size == 2001 and om[0] == 995.0
    _, ref2 = port.absorption_coefficient_voigt(cols, OmegaGrid=om)
    assert_parity(xs2, ref2, RTOL, FLOOR, "range route")
    _, xs3 = hapi.absorptionCoefficient_Voigt(SourceTables="ties", OmegaGrid=om, Components=[(1, 2)])
    _, ref3 = port.absorption_coefficient_voigt(cols, OmegaGrid=om, Components=[(1, 2)])
    assert_parity(xs3, ref3, RTOL, FLOOR, "components filter")


def test_voigt_optional_columns(mods):
    """Optional hapi columns on the device path: n_self (0 -> n_air), delta_self, deltap_air, deltap_self
    (misc/hapi.py:11097-11128), with air/self diluent mixes, against the oracle (and the reference via the oracle's
    golden pins for the plain cases)."""
    import make_golden as mg
    port, syn, hapi = mods["port"], mods["syn"], mods["hapi"]
    cols = mg.optcols_table()
    g = golden("optcols.npz")
    hapi.storeTable("optcols", cols)
    grid = np.linspace(1200.0, 1300.0, 20001)
    for k, (T, p, dil) in enumerate(mg.OPT_CASES):
        _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="optcols", OmegaGrid=grid, Environment={"T": T, "p": p},
                                                 Diluent=dil)
        assert_parity(xs, g["xs%d" % k], RTOL, FLOOR, "optional columns %s" % dil)      # the reference's own output
    # a table with deltap but no self columns
    c2 = {k: v for k, v in cols.items() if k not in ("gamma_self", "n_self", "delta_self", "deltap_self")}
    hapi.storeTable("optcols2", c2)
    _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="optcols2", OmegaGrid=grid, Environment={"T": 230.0, "p": 0.2})
    _, ref = port.absorption_coefficient_voigt(c2, T=230.0, p=0.2, OmegaGrid=grid)
    assert_parity(xs, ref, RTOL, FLOOR, "deltap_air only")


def test_voigt_two_tables_and_deterministic(mods):
    port, syn, hapi = mods["port"], mods["syn"], mods["hapi"]
    a = syn.synthetic_lines(53, 150, 900.0, 1000.0, 1)
    b = syn.synthetic_lines(54, 150, 900.0, 1000.0, 2)
    hapi.storeTable("ta", a); hapi.storeTable("tb", b)
    grid = np.linspace(920.0, 980.0, 6001)
    _, xs = hapi.absorptionCoefficient_Voigt(SourceTables=["ta", "tb"], OmegaGrid=grid)
    _, ra = port.absorption_coefficient_voigt(a, OmegaGrid=grid)
    _, rb = port.absorption_coefficient_voigt(b, OmegaGrid=grid)
    assert_parity(xs, ra + rb, RTOL, FLOOR, "two tables")
    _, xs2 = hapi.absorptionCoefficient_Voigt(SourceTables=["ta", "tb"], OmegaGrid=grid)
    assert np.array_equal(xs, xs2)            # bit-reproducible run to run


def _tud_setup(mods):
    import make_golden as mg
    atm, tabs = mg.tud_inputs()
    rt = mods["rt"]
    opts = dict(rt.options)
    rt.set_line_database(tabs, opts=opts)
    kw = dict(Zs=atm["Zs"], Ts=atm["Ts"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_VAL=atm["MFs_VAL"], MFs_ID=atm["MFs_ID"],
              DVOUT=0.05)
    return atm, tabs, opts, kw


def test_tud_golden(mods):
    """compute_TUD against the reference's own outputs (tests/golden/tud.npz), incl. the nL-overwrite quirk."""
    rt = mods["rt"]
    g = golden("tud.npz")
    atm, tabs, opts, kw = _tud_setup(mods)
    X, tau, Lu, Ld = rt.compute_TUD(700.0, 750.0, opts=dict(opts), **kw)
    assert np.array_equal(X, g["X"]) and tau.shape == X.shape and Ld.shape == X.shape
    for a, b, w in ((tau, g["tau"], "tau"), (Lu, g["Lu"], "Lu"), (Ld, g["Ld"], "Ld")):
        assert_parity(a, b, RTOL, FLOOR, w)
    X, tau, Lu, Ld = rt.compute_TUD(700.0, 750.0, opts=dict(opts), Altitudes=np.array([1.0, 10.0, 500.0, 3.0]),
                                    N_angle=7, **kw)
    assert tau.shape == (X.size, 4)
    for a, b, w in ((tau, g["tau_alt"], "tau_alt"), (Lu, g["Lu_alt"], "Lu_alt"), (Ld, g["Ld_alt"], "Ld_alt")):
        assert_parity(a, b, RTOL, FLOOR, w)
    X, tau, Lu, Ld = rt.compute_TUD(700.0, 750.0, opts=dict(opts), returnOD=True, theta_r=0.6, N_angle=4, **kw)
    for a, b, w in ((tau, g["tau_od"], "tau_od"), (Lu, g["Lu_od"], "Lu_od"), (Ld, g["Ld_od"], "Ld_od")):
        assert_parity(a, b, RTOL, FLOOR, w)


def test_tud_many_layers_uncached_path(mods):
    """More layers than the shared-memory Planck cache holds (> 72 layers): the recompute variant of the kernel."""
    engine, port = mods["engine"], mods["port"]
    rng = np.random.default_rng(11)
    nL, nx = 90, 700
    X = np.linspace(700.0, 1400.0, nx)
    OD = rng.uniform(0, 0.3, (nx, nL)) * rng.uniform(0, 1, (nx, 1))
    Z = np.linspace(0.0, 89.0, nL); T = rng.uniform(200.0, 300.0, nL)
    tau, lu, ld = engine.tud(np.ascontiguousarray(OD.T), T[None, :], X, Z, [5.0, 500.0], [1.0, 1.3], 12)
    t_ref, u_ref, d_ref = port.tud_from_od(X, OD, Z, T, N_angle=12, Altitudes=np.array([5.0, 500.0]), theta_r=0.0)
    assert_parity(tau[0, :, :, 0].cpu().numpy(), t_ref, RTOL, FLOOR, "tau 90 layers")
    assert_parity(lu[0, :, :, 0].cpu().numpy(), u_ref, RTOL, FLOOR, "Lu 90 layers")
    assert_parity(ld[0].cpu().numpy(), d_ref, RTOL, FLOOR, "Ld 90 layers")


def test_tud_sticky_options_and_compute_od(mods):
    rt, port = mods["rt"], mods["port"]
    atm, tabs, opts, kw = _tud_setup(mods)
    o = dict(opts)
    rt.compute_TUD(700.0, 710.0, opts=o, returnOD=True, **kw)
    assert o["returnOD"] is True and o["DVOUT"] == 0.05       # opts.update(kwargs): settings stick (ref :303)
    X, od = rt.compute_OD(700.0, 750.0, opts=dict(opts), T=atm["Ts"][3], P=atm["Ps"][3], PL=atm["PLs"][3],
                          MF_VAL=atm["MFs_VAL"][3], MF_ID=atm["MFs_ID"], DVOUT=0.05)
    ref = port.layer_od(tabs, X, atm["Ts"][3], atm["Ps"][3], atm["PLs"][3], atm["MFs_VAL"][3], atm["MFs_ID"])
    assert_parity(od, ref, RTOL, FLOOR, "compute_OD")
    with pytest.raises(RuntimeError, match="line database"):
        rt.compute_TUD(700.0, 710.0, opts={k: v for k, v in rt.options.items() if k != "LINE_DB"})


def test_tud_batch_matches_single(mods):
    import torch
    torch_equal = torch.equal
    rt, syn = mods["rt"], mods["syn"]
    atm, tabs, opts, kw = _tud_setup(mods)
    Ts, MF = syn.perturbed_atmospheres(2, 5)
    o = dict(opts)
    X, tau, lu, ld = rt.compute_TUD_batch(700.0, 730.0, Ts, MF[:, :, :], opts=o, atm_chunk=2,
                                          **{k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL")})
    assert tau.shape == (5, X.size, 1, 1)
    # automatic kernel variants follow the problem size (a single atmosphere on 600 points is a "tiny" launch whose
    # lines are split over several CTAs, a chunk of two atmospheres is not): agreement to rounding ...
    for j in (0, 3, 4):
        Xs, t1, u1, d1 = rt.compute_TUD(700.0, 730.0, opts=dict(opts), **dict(kw, Ts=Ts[j], MFs_VAL=MF[j]))
        for a, b in ((tau[j, :, 0, 0], t1), (lu[j, :, 0, 0], u1), (ld[j], d1)):
            assert_parity(a.cpu().numpy(), b, 1e-10, 1e-13, "batch member %d vs single call" % j)
    # a batch does not depend on how it is chunked
    X2, tau2, lu2, ld2 = rt.compute_TUD_batch(700.0, 730.0, Ts, MF[:, :, :], opts=dict(opts), atm_chunk=5,
                                              **{k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL")})
    assert torch_equal(tau, tau2) and torch_equal(lu, lu2) and torch_equal(ld, ld2)
    # ... and with the tile size pinned (which also switches the split off) bit for bit
    o8 = dict(opts); o8["voigt"] = {"tile_k": 8}
    X, tau, lu, ld = rt.compute_TUD_batch(700.0, 730.0, Ts, MF[:, :, :], opts=dict(o8), atm_chunk=2,
                                          **{k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL")})
    for j in (0, 3, 4):
        Xs, t1, u1, d1 = rt.compute_TUD(700.0, 730.0, opts=dict(o8), **dict(kw, Ts=Ts[j], MFs_VAL=MF[j]))
        assert np.array_equal(tau[j, :, 0, 0].cpu().numpy(), t1)
        assert np.array_equal(lu[j, :, 0, 0].cpu().numpy(), u1)
        assert np.array_equal(ld[j].cpu().numpy(), d1)
    Xb, tb, ub, db_ = rt.compute_TUD_batch(700.0, 1400.0, Ts[:2], MF[:2], opts=dict(o8), ils="triangle",
                                           **dict({k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL")}, DVOUT=0.25))
    Xs, t1, u1, d1 = rt.compute_TUD(700.0, 1400.0, opts=dict(o8), **dict(kw, Ts=Ts[1], MFs_VAL=MF[1], DVOUT=0.25))
    xo, tt = rt.ILS_MAKO(Xs, t1)
    assert np.array_equal(xo, Xb) and np.array_equal(tb[1, :, 0, 0].cpu().numpy(), tt)


def test_sensor_golden(mods):
    rt, ilsg = mods["rt"], mods["ilsg"]
    g = golden("sensor.npz")
    X, Y1, Y2 = g["X"], g["Y1"], g["Y2"]
    Xo, t1 = rt.ILS_MAKO(X, Y1)
    assert np.array_equal(Xo, g["Xo"]) and t1.shape == (128,)
    assert_parity(t1, g["tri1"], RTOL, FLOOR, "tri 1-D")
    t2 = rt.ILS_MAKO(X, Y2, returnX=False)
    assert t2.shape == (128, 5)
    assert_parity(t2, g["tri2"], RTOL, FLOOR, "tri 2-D")
    Xr, t3 = rt.ILS_MAKO(X, Y2, resFactor=2, fwhm_sf=0.8, shift=0.3, scale=1.001)
    assert np.array_equal(Xr, g["Xo_res"])
    assert_parity(t3, g["tri_res"], RTOL, FLOOR, "tri resFactor")
    Xc, t4 = rt.ILS_MAKO(g["Xc"], Y2[:4001])
    assert np.array_equal(Xc, g["Xo_clip"])
    assert_parity(t4, g["tri_clip"], RTOL, FLOOR, "tri clipped")
    Xg, g1 = ilsg.ILS_MAKO(X, Y1)
    assert np.array_equal(Xg, g["Xg"])
    assert_parity(g1, g["gauss1"], RTOL, FLOOR, "gauss 1-D")
    assert_parity(ilsg.ILS_MAKO(X, Y2)[1], g["gauss2"], RTOL, FLOOR, "gauss 2-D")
    args = (g["Xo"], g["emis"], g["Ts"], g["tau"], g["La"], g["Ld"])
    assert_parity(rt.compute_LWIR_apparent_radiance(*args), g["L0"], RTOL, FLOOR, "apparent")
    L1, Ls1 = rt.compute_LWIR_apparent_radiance(*args, dT=g["dT"], return_Ls=True)
    assert L1.shape == g["L1"].shape
    assert_parity(L1, g["L1"], RTOL, FLOOR, "apparent dT")
    assert_parity(Ls1, g["Ls1"], RTOL, FLOOR, "apparent Ls")
    assert_parity(rt.planckian(500.0, 296.0), g["planck_kat"], RTOL, FLOOR, "planck kat")
    assert_parity(rt.planckian(np.linspace(700, 1500, 8001), np.array([280.0, 300.0])), g["planck_grid"], RTOL, FLOOR)
    Pw = rt.planckian(np.linspace(7.5, 13.5, 61), np.array([[280.0, 300.0], [310., 250.]]), wavelength=True)
    assert Pw.shape == (61, 2, 2)
    assert_parity(Pw, g["planck_um"], RTOL, FLOOR, "planck um")


def test_brightness_temperature_golden(mods):
    rt = mods["rt"]
    g = golden("bt.npz")
    X, T, L = g["X"], g["T"], g["L"]
    got = rt.BT2L(X, T)
    assert got.shape == L.shape
    assert_parity(got, L, RTOL, FLOOR, "BT2L")
    bt = rt.brightnessTemperature(X, g["Lb"], bad_value=-5.0)
    assert_parity(bt, g["bt"], RTOL, FLOOR, "brightnessTemperature")
    assert bt[3, 0, 0] == -5.0 and bt[4, 1, 1] == -5.0 and bt[5, 2, 0] == -5.0
    v = rt.brightnessTemperature(X, L[:, 0, 0])
    assert v.shape == g["bt_vec"].shape == (X.size, 1)
    assert_parity(v, g["bt_vec"], RTOL, FLOOR, "bt vector")
    Lbad = rt.BT2L(X, g["Tb"])
    assert np.isnan(Lbad[7, 0, 1]) and np.isnan(g["L_bad"][7, 0, 1])
    m = np.isfinite(g["L_bad"])
    assert_parity(Lbad[m], g["L_bad"][m], RTOL, FLOOR, "BT2L bad")
    assert_parity(rt.brightnessTemperature(g["Xu"], g["Lu"], wavelength=True), g["bt_um"], RTOL, FLOOR, "bt um")
    sd = rt.brightnessTemperature(X, np.swapaxes(L[:, :, 0], 0, 1), spectral_dim=1)
    assert sd.shape == g["bt_sd"].shape
    assert_parity(sd, g["bt_sd"], RTOL, FLOOR, "bt spectral_dim")
    # round trip
    assert_parity(rt.brightnessTemperature(X, rt.BT2L(X, T)), T, 1e-9, 1e-12, "BT round trip")


def test_reduce_resolution_golden(mods):
    """smooth / reduceResolution against the reference's outputs (np.int restored)."""
    rt = mods["rt"]
    g = golden("reduce.npz")
    rng = np.random.default_rng(70)
    X = np.linspace(690.0, 1410.0, 28800)
    Y = np.cumsum(rng.normal(0, 1, (X.size, 3)), axis=0) * 0.01 + np.sin(X / 7.0)[:, None]
    Xo, Yo = rt.reduceResolution(X, Y, 0.25)
    assert Xo.shape == g["Xo"].shape and Yo.shape == g["Yo"].shape
    assert_parity(Xo, g["Xo"], 1e-12, 0.0, "reduceResolution X_out")
    assert_parity(Yo, g["Yo"], RTOL, FLOOR, "reduceResolution Y_out")
    Xo5, Yo5 = rt.reduceResolution(X, Y[:, 1], 0.1, N=2, window="hamming")
    assert Yo5.shape == g["Yo5"].shape
    assert_parity(Yo5, g["Yo5"], RTOL, FLOOR, "reduceResolution 1-D hamming")
    y1 = rt.reduceResolution(X, Y[:, 0], 0.25, X_out=g["Xo"])
    assert_parity(y1, g["Yo"][:, 0], RTOL, FLOOR, "reduceResolution given X_out")
    assert_parity(rt.smooth(Y[:, 0], 11), g["s11"], RTOL, FLOOR, "smooth hanning 11")
    assert_parity(rt.smooth(Y[:, 2], 10, "flat"), g["s10f"], RTOL, FLOOR, "smooth flat 10")
    assert rt.smooth(Y[:, 0], 2) is Y[:, 0] or np.array_equal(rt.smooth(Y[:, 0], 2), Y[:, 0])


def test_reduce_resolution_ends_golden(mods):
    """reduceResolution where the spline's end conditions matter (scipy interp1d(kind='cubic') is a not-a-knot spline
    and extrapolates with the end polynomial): smFactor = 3 (X_out starts three knots inside the data), a caller-supplied
    X_out from outside the data range to outside it, and a 12-knot axis (both ends inside every window); against the
    reference's outputs (tests/golden/reduce2.npz)."""
    rt = mods["rt"]
    g = golden("reduce2.npz")
    rng = np.random.default_rng(71)
    X = np.linspace(700.0, 760.0, 2401)
    Y = np.cumsum(rng.normal(0, 1, (X.size, 2)), axis=0) * 0.02 + np.cos(X / 3.0)[:, None]
    Xo3, Yo3 = rt.reduceResolution(X, Y, 0.075)
    assert_parity(Xo3, g["Xo3"], 1e-12, 0.0, "smFactor 3 X_out")
    assert_parity(Yo3, g["Yo3"], RTOL, FLOOR, "smFactor 3 Y_out")
    Yfull = rt.reduceResolution(X, Y, 0.25, X_out=g["Xfull"])
    assert_parity(Yfull, g["Yfull"], RTOL, FLOOR, "X_out across and beyond the data")
    Xs = np.linspace(700.0, 701.0, 12)
    Yshort = rt.reduceResolution(Xs, np.sin(Xs * 9.0) + Xs, 0.3, X_out=np.linspace(699.9, 701.1, 41))
    assert_parity(Yshort, g["Yshort"], RTOL, FLOOR, "12-knot axis")


def test_ils_many_spectra(mods):
    """The coalesced many-spectra ILS kernels (ns >= 16), both line shapes, ragged ns, against the oracle."""
    rt, ilsg, port = mods["rt"], mods["ilsg"], mods["port"]
    rng = np.random.default_rng(8)
    X = np.linspace(700.0, 1400.0, 3501)
    for ns in (16, 45):
        Y = rng.uniform(0, 10, (X.size, ns))
        Xo, t = rt.ILS_MAKO(X, Y)
        assert_parity(t, port.ils_mako_triangle(X, Y)[1], RTOL, FLOOR, "tri ns=%d" % ns)
        Xg, g = ilsg.ILS_MAKO(X, Y)
        assert_parity(g, port.ils_mako_gauss(X, Y)[1], RTOL, FLOOR, "gauss ns=%d" % ns)


def test_ils_grouped_walks(mods):
    """Enough spectra for the grouped walks: TMA-staged groups of 8 adjacent bands with 2 (0x400) or 1 (0x800) spectra per
    lane (default for even ns), groups of 16 (0x1000), the grouped walk without staging (0x200, also the fallback for odd ns) and the
    band-per-CTA walk (0x100)
    agree with each other to rounding and with the oracle; ragged last spectra block; a Gaussian band whose centre lies
    outside the axis (empty 9.5-sigma support) falls back to the whole axis like the reference's sum."""
    torch, engine, rt, port = mods["torch"], mods["engine"], mods["rt"], mods["port"]
    rng = np.random.default_rng(18)
    X = np.linspace(700.0, 1400.0, 2801)
    tri = rt._mako_bands(X, None, 1.0, 0.0, 1.0)
    gau = rt._mako_bands_gauss()
    for ns in (2600, 1301):                                  # even -> TMA staging; odd -> plain grouped walk
        Y = rng.uniform(0, 10, (X.size, ns))
        Yd = torch.from_numpy(Y).cuda()
        for mode, bands in ((0, tri), (1, gau)):
            outs = {v: engine.ils_apply(mode | v, X, Yd, bands[1], bands[2]).cpu().numpy() for v in (0, 0x400, 0x800, 0x1400, 0x200, 0x100)}
            for v in (0x400, 0x800, 0x1400, 0x200, 0x100):
                ok = np.isfinite(outs[0x100])
                assert np.array_equal(np.isfinite(outs[v]), ok)
                assert np.max(np.abs(outs[v][ok] - outs[0][ok]) / np.abs(outs[0][ok])) < 1e-13, (ns, mode, hex(v))
        cols = [0, 31, 32, ns - 1]
        assert_parity(engine.ils_apply(0, X, Yd, tri[1], tri[2]).cpu().numpy()[:, cols],
                      port.ils_mako_triangle(X, Y[:, cols])[1], RTOL, FLOOR, "grouped tri ns=%d" % ns)
        assert_parity(engine.ils_apply(1, X, Yd, gau[1], gau[2]).cpu().numpy()[:, cols],
                      port.ils_mako_gauss(X, Y[:, cols])[1], RTOL, FLOOR, "grouped gauss ns=%d" % ns)


def test_sdvoigt_lut_and_afit_xs(mods, tmp_path):
    """The reference's one hapi call site (misc/RT_gen_AbsXS_files.py:90): SDVoigt on plain HITRAN-style tables with a
    350-halfwidth wing, batched over (T, p), and the AFIT_XS file it writes."""
    hapi, syn = mods["hapi"], mods["syn"]
    from radtxfr_b200 import formats
    g = golden("lut.npz")
    cols = syn.synthetic_lines(80, 250, 1950.0, 2400.0, 2, (1, 2, 3), (0.8, 0.1, 0.1))
    hapi.storeTable("lut", cols)
    grid = np.linspace(2000.0, 2350.0, 7001)
    states = [(296.0, 1.0), (250.0, 0.1), (330.0, 0.5)]
    for k, (T, p) in enumerate(states):
        _, xs = hapi.absorptionCoefficient_SDVoigt(SourceTables="lut", HITRAN_units=True, Environment={"T": T, "p": p},
                                                   OmegaGrid=grid, IntensityThreshold=0, WavenumberWingHW=350.0)
        assert_parity(xs, g["xs%d" % k], RTOL, FLOOR, "SDVoigt state %d" % k)
    lut = formats.cross_section_lut("lut", [296.0, 250.0, 330.0], [1.0, 0.1, 0.5], grid, WavenumberWingHW=350.0)
    assert lut.shape == (3, 3, grid.size)
    assert_parity(lut[0, 0], g["xs0"], RTOL, FLOOR, "lut[0,0]")
    assert_parity(lut[1, 1], g["xs1"], RTOL, FLOOR, "lut[1,1]")
    assert_parity(lut[2, 2], g["xs2"], RTOL, FLOOR, "lut[2,2]")
    f = formats.AFIT_XS_write(grid, g["xs0"], 296.0, 101325 * 1.0, 2, "HITRAN2016-synthetic", File=str(tmp_path / "x.bin"))
    assert np.array_equal(np.frombuffer(open(f, "rb").read(), dtype=np.uint8), g["afit_bytes"])
    X, Y, T, P, ID, name = formats.AFIT_XS_read(f)
    assert np.array_equal(Y, g["xs0"]) and (T, P, ID, name) == (296.0, 101325.0, 2, "HITRAN2016-synthetic")


def test_sdvoigt_speed_dependence_and_diluents(mods):
    """True speed-dependent Voigt (tables with SD_<diluent> columns: pcqsdhc PART2-4 with cpf3 where the reference selects
    it) and broadening species other than air / self (gamma_h2, n_h2, delta_h2, deltap_h2; gamma_he without an n column;
    a species the table has no columns for contributes nothing), against the reference's own outputs
    (tests/golden/sdvoigt.npz) and the oracle port of pcqsdhc."""
    import make_golden as mg
    hapi, port = mods["hapi"], mods["port"]
    g = golden("sdvoigt.npz")
    cols = mg.sdv_table()
    hapi.storeTable("sdv", cols)
    grid = np.linspace(1200.0, 1300.0, 20001)
    for k, (T, p, dil) in enumerate(mg.SDV_CASES):
        _, xs = hapi.absorptionCoefficient_SDVoigt(SourceTables="sdv", Environment={"T": T, "p": p}, OmegaGrid=grid,
                                                   Diluent=dil)
        assert_parity(xs, g["sd%d" % k], RTOL, FLOOR, "SDVoigt with SD columns, case %d" % k)
        _, plain = hapi.absorptionCoefficient_Voigt(SourceTables="sdv", Environment={"T": T, "p": p}, OmegaGrid=grid,
                                                    Diluent=dil)
        assert np.max(np.abs(plain - g["sd%d" % k])) > 1e-3 * np.max(g["sd%d" % k])      # the SD columns matter
    for k, (T, p, dil) in enumerate(mg.DIL_CASES):
        _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="sdv", Environment={"T": T, "p": p}, OmegaGrid=grid,
                                                 Diluent=dil)
        assert_parity(xs, g["dil%d" % k], RTOL, FLOOR, "Voigt with h2/he/co2 diluents, case %d" % k)
    # tiny Gamma2 (PART2 of pcqsdhc: |X| <= 3e-8 |Y|) and a fine grid, against the oracle port
    tiny = dict(cols, SD_air=np.full(300, 1e-13))
    hapi.storeTable("sdv_tiny", tiny)
    fine = np.linspace(1240.0, 1242.0, 4001)
    _, xs = hapi.absorptionCoefficient_SDVoigt(SourceTables="sdv_tiny", Environment={"T": 260.0, "p": 0.5}, OmegaGrid=fine)
    _, ref = port.absorption_coefficient_sdvoigt(tiny, T=260.0, p=0.5, OmegaGrid=fine)
    assert_parity(xs, ref, RTOL, FLOOR, "SDVoigt PART2")


def test_hsi_pixels(mods):
    """Per-pixel mixed-material radiance (LWIR_HSI_Generator.py:148-166): 128-band MAKO axis and a 300-point axis
    (more bands than threads), pixel counts that do not fill the last group, three-material mixes, and the float32
    cube (config 4's FP32 variant: same arithmetic, rounded at the store)."""
    torch, engine, port, syn = mods["torch"], mods["engine"], mods["port"], mods["syn"]
    rng = np.random.default_rng(7)
    for X, nP, nM in ((port.mako_axis(np.array([700.0, 1400.0])), 1000, 2), (np.linspace(750.0, 1300.0, 300), 77, 3),
                      (np.linspace(750.0, 1300.0, 5), 4097, 1), (np.linspace(750.0, 1300.0, 40), 130, 5)):
        nE, nA = 9, 4
        emis = syn.synthetic_emissivities(3, nE, X).T.copy()
        tau = rng.uniform(0.2, 1, (nA, X.size)); La = rng.uniform(0, 300, (nA, X.size)); Ld = rng.uniform(0, 600, (nA, X.size))
        atm = rng.integers(0, nA, nP); Tp = rng.uniform(270, 320, nP)
        mat = rng.integers(0, nE, (nP, nM)); fr = rng.uniform(0, 1, (nP, nM)); fr /= fr.sum(1)[:, None]
        L = engine.hsi_pixels(X, emis, tau, La, Ld, atm, Tp, mat, fr).cpu().numpy()
        em = np.array([np.dot(fr[i], emis[mat[i]]) for i in range(nP)])       # LWIR_HSI_Generator.py:159
        B = port.planckian(X, Tp).T
        ref = tau[atm] * (em * B + (1 - em) * Ld[atm]) + La[atm]
        assert_parity(L, ref, RTOL, FLOOR, "hsi pixels nx=%d" % X.size)
        L32 = engine.hsi_pixels(X, emis, tau, La, Ld, atm, Tp, mat, fr, out_dtype=torch.float32)
        assert L32.dtype == torch.float32
        assert np.array_equal(L32.cpu().numpy(), L.astype(np.float32))        # rounded once at the store


def test_apparent_radiance_float32(mods):
    """compute_LWIR_apparent_radiance called the way the production script calls it (every input float32, dT =
    arange(-10, 10.5, 0.5); Compute_LWIR_Apparent_Radiance.py:9-25) returns float32 like numpy does, within float32
    rounding of the reference's own float32 result (tests/golden/sensor_f32.npz, minted from the reference)."""
    rt, port = mods["rt"], mods["port"]
    g = golden("sensor_f32.npz")
    args = [g[k] for k in ("X", "emis", "Ts", "tau", "La", "Ld")]
    L = rt.compute_LWIR_apparent_radiance(*args, g["dT"])
    assert L.dtype == np.float32 and L.shape == g["L"].shape
    assert_parity(L.astype(float), g["L"].astype(float), 1e-5, 1e-8, "float32 apparent radiance vs reference")
    L0 = rt.compute_LWIR_apparent_radiance(*args)
    assert L0.dtype == np.float32
    assert_parity(L0.astype(float), g["L0"].astype(float), 1e-5, 1e-8, "float32 apparent radiance (no dT)")
    ref64 = port.apparent_radiance(*[a.astype(float) for a in args], dT=g["dT"].astype(float))
    assert np.array_equal(L, ref64.astype(np.float32)) or np.max(np.abs(L - ref64) / np.abs(ref64)) < 1.2e-7


def test_full_size_properties(mods):
    """Config-2-sized grid (800 000 points, 66 layers): size-independent properties instead of a CPU oracle run.
    (1) linearity in the column amounts, (2) bit-reproducibility, (3) tile-size independence, (4) the first 2 000
    points of three layers against the oracle, (5) TUD invariants: 0<tau<=1, Lu -> B where the column is opaque."""
    torch, rt, syn, engine, port = (mods[k] for k in ("torch", "rt", "syn", "engine", "port"))
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables(scale=0.02)
    db = engine.LineDB([tabs[m] for m in sorted(tabs)])
    X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    assert X.size == 800000
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    grid = torch.from_numpy(X).cuda()
    od = db.voigt_sum(atm["Ts"], p, w, grid)
    assert torch.isfinite(od).all() and (od >= 0).all()
    od2 = db.voigt_sum(atm["Ts"], p, 2.0 * w, grid)
    assert torch.equal(od2, 2.0 * od)                                   # exact: power-of-two scaling
    assert torch.equal(od, db.voigt_sum(atm["Ts"], p, w, grid))         # bit-reproducible
    # tile-size independence: K=2 has no far-field blocks and pairs lines differently -> default arithmetic agrees
    # to its own bound (one-step reciprocal 9e-13 per term), the exact arithmetic to rounding
    od_k2 = db.voigt_sum(atm["Ts"], p, w, grid, flags=2 << 8)
    err = ((od_k2 - od).abs() / od.abs().clamp_min(1e-300)).max().item()
    assert err < 5e-12, err
    ex8 = db.voigt_sum(atm["Ts"], p, w, grid, flags=128 | 4096 | (8 << 8))
    ex2 = db.voigt_sum(atm["Ts"], p, w, grid, flags=128 | 4096 | (2 << 8))
    assert ((ex8 - ex2).abs().max() / ex8.abs().max()).item() < 1e-13
    assert ((od - ex8).abs() / ex8.abs().clamp_min(1e-300)).max().item() < 5e-12
    del od_k2, ex8, ex2
    sub = slice(0, 2000)
    for layer in (0, 30, 65):
        ref = port.layer_od(tabs, X[sub], atm["Ts"][layer], atm["Ps"][layer], atm["PLs"][layer],
                            atm["MFs_VAL"][layer], atm["MFs_ID"])
        # the oracle sees only the first 2000 points, the windows are point-wise so the values must agree
        assert_parity(od[layer, sub].cpu().numpy(), ref, RTOL, FLOOR, "layer %d" % layer)
    tau, lu, ld = engine.tud(od, atm["Ts"][None, :], grid, atm["Zs"], [500.0], [1.0], 30)
    tau = tau[0, :, 0, 0]; lu = lu[0, :, 0, 0]; ld = ld[0]
    assert (tau >= 0).all() and (tau <= 1).all() and torch.isfinite(lu).all() and torch.isfinite(ld).all()
    B0 = torch.from_numpy(port.planckian(X, atm["Ts"][:1])[:, 0]).cuda()
    assert (ld <= B0 * (1 + 1e-12)).all()                               # downwelling never exceeds the warmest layer


def test_config2_full_density_window(mods):
    """Config 2 exactly as benchmarked (all 210 000 synthetic lines, 66 layers, 0.001 cm^-1), compared with the oracle on
    a 2 000-point window for EVERY layer (the oracle gets only the lines that can reach the window), then through the
    layer integration."""
    torch, rt, syn, engine, port = (mods[k] for k in ("torch", "rt", "syn", "engine", "port"))
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables()
    opts = dict(rt.options)
    db = rt.set_line_database(tabs, opts=opts)
    X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    i0, i1 = 401000, 403000
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    grid = torch.from_numpy(X).cuda()
    od = db.voigt_sum(atm["Ts"], p, w, grid)
    lo, hi = X[i0] - 12.0, X[i1 - 1] + 12.0          # 50 half-widths of the widest line is < 8 cm^-1
    sub = {m: {k: v[(t["nu"] > lo) & (t["nu"] < hi)] for k, v in t.items()} for m, t in tabs.items()}
    OD = np.zeros((i1 - i0, 66))
    for L in range(66):
        OD[:, L] = port.layer_od(sub, X[i0:i1], atm["Ts"][L], atm["Ps"][L], atm["PLs"][L], atm["MFs_VAL"][L], atm["MFs_ID"])
        assert_parity(od[L, i0:i1].cpu().numpy(), OD[:, L], RTOL, FLOOR, "layer %d" % L)
    tr, ur, dr = port.tud_from_od(X[i0:i1], OD, atm["Zs"], atm["Ts"])
    tau, lu, ld = engine.tud(od[:, i0:i1].contiguous(), atm["Ts"][None, :], grid[i0:i1].contiguous(), atm["Zs"], [500.0], [1.0], 30)
    assert_parity(tau[0, :, 0, 0].cpu().numpy(), tr, RTOL, FLOOR, "tau")
    assert_parity(lu[0, :, 0, 0].cpu().numpy(), ur, RTOL, FLOOR, "Lu")
    assert_parity(ld[0].cpu().numpy(), dr, RTOL, FLOOR, "Ld")


def test_config2_full_size_against_c_oracle(mods):
    """The benchmark configuration at FULL size -- 800 000 grid points x 66 layers x 210 000 lines, 3.3e10 evaluations
    -- compared point by point with the CPU oracle (C/OpenMP restatement, oracle/voigt_c.c, pinned to the numpy port and
    the reference's golden vectors), then through the layer integration against the numpy port."""
    torch, rt, syn, engine, port = (mods[k] for k in ("torch", "rt", "syn", "engine", "port"))
    from oracle import cbuild
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables()
    opts = dict(rt.options)
    db = rt.set_line_database(tabs, opts=opts)
    X = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    grid = torch.from_numpy(X).cuda()
    od = db.voigt_sum(atm["Ts"], p, w, grid)
    OD = np.zeros((X.size, 66))
    worst = 0.0
    for L in range(66):
        OD[:, L] = cbuild.layer_od(tabs, X, atm["Ts"][L], atm["Ps"][L], atm["PLs"][L], atm["MFs_VAL"][L], atm["MFs_ID"])
        got = od[L].cpu().numpy()
        assert_parity(got, OD[:, L], RTOL, FLOOR, "full-size layer %d" % L)
        big = OD[:, L] >= 1e-6 * OD[:, L].max()
        worst = max(worst, float(np.max(np.abs(got - OD[:, L])[big] / OD[:, L][big])))
    assert worst < 1e-9
    tr, ur, dr = port.tud_from_od(X[::40], OD[::40], atm["Zs"], atm["Ts"])          # numpy recurrences on every 40th point
    tau, lu, ld = engine.tud(od, atm["Ts"][None, :], grid, atm["Zs"], [500.0], [1.0], 30)
    assert_parity(tau[0, ::40, 0, 0].cpu().numpy(), tr, RTOL, FLOOR, "full-size tau")
    assert_parity(lu[0, ::40, 0, 0].cpu().numpy(), ur, RTOL, FLOOR, "full-size Lu")
    assert_parity(ld[0, ::40].cpu().numpy(), dr, RTOL, FLOOR, "full-size Ld")
    print("full-size config 2: max relative deviation of the optical depths %.2e" % worst)


def test_generate_lwir_tud_driver(mods):
    """The production sweep (Generate_LWIR_TUD.py:79-132) as one device pipeline, against a loop of oracle calls:
    compute_TUD(returnOD=True, nine altitudes) followed by reduceResolution of OD, La and Ld, per atmosphere."""
    rt, syn, port = mods["rt"], mods["syn"], mods["port"]
    from radtxfr_b200 import drivers
    import make_golden as mg
    atm, tabs = mg.tud_inputs()
    opts = dict(rt.options)
    rt.set_line_database({m: tabs[m] for m in (1, 2, 3)}, opts=opts)
    Ts, MF = syn.perturbed_atmospheres(6, 3)
    H2O, O3 = MF[:, :, 0] * 1e-6, MF[:, :, 2] * 1e-6
    CO2 = atm["MFs_VAL"][:, 1] * 1e-6
    Alt = drivers.default_altitudes(float(atm["Zs"].max()))
    X, OD, La, Ld = drivers.generate_LWIR_TUD(Ts, H2O, O3, CO2, Xmin=700.0, Xmax=750.0, DV=0.01, DV_out=0.25,
                                              Altitudes=Alt, opts=opts, atm_chunk=2)
    assert OD.shape == (X.size, 9, 3) and La.shape == OD.shape and Ld.shape == (X.size, 3)
    Xr = None
    for jj in range(3):
        MFj = 1e6 * np.stack([H2O[jj], CO2, O3[jj]], axis=1)
        XX, od_, la_, ld_ = port.compute_tud(tabs, 700.0, 750.0, Zs=atm["Zs"], Ts=Ts[jj], Ps=atm["Ps"], PLs=atm["PLs"],
                                             MFs_VAL=MFj, MFs_ID=[1, 2, 3], DVOUT=0.01, Altitudes=Alt, returnOD=True)
        if Xr is None:
            Xr, od_r = port.reduce_resolution(XX, od_, 0.25)
        else:
            od_r = port.reduce_resolution(XX, od_, 0.25, X_out=Xr)
        la_r = port.reduce_resolution(XX, la_, 0.25, X_out=Xr)
        ld_r = port.reduce_resolution(XX, ld_, 0.25, X_out=Xr)
        assert_parity(X, Xr, 1e-12, 0.0, "driver axis")
        assert_parity(OD[:, :, jj], od_r, RTOL, FLOOR, "driver OD atm %d" % jj)
        assert_parity(La[:, :, jj], la_r, RTOL, FLOOR, "driver La atm %d" % jj)
        assert_parity(Ld[:, jj], ld_r, RTOL, FLOOR, "driver Ld atm %d" % jj)


def test_lorentz_profile_golden(mods):
    """absorptionCoefficient_Lorentz against the reference's own outputs (tests/golden/lorentz.npz)."""
    import make_golden as mg
    hapi, syn = mods["hapi"], mods["syn"]
    g = golden("lorentz.npz")
    cols = syn.synthetic_lines(90, 300, 1480.0, 1620.0, 1)
    hapi.storeTable("lor", cols)
    grid = np.linspace(1500.0, 1600.0, 10001)
    for k, (T, p, kw) in enumerate(mg.LORENTZ_CASES):
        _, xs = hapi.absorptionCoefficient_Lorentz(SourceTables="lor", Environment={"T": T, "p": p}, OmegaGrid=grid, **kw)
        assert_parity(xs, g["xs%d" % k], RTOL, FLOOR, "Lorentz %d" % k)


def test_doppler_profile_golden(mods):
    """absorptionCoefficient_Doppler against the reference's own outputs (tests/golden/doppler.npz)."""
    import make_golden as mg
    hapi, syn = mods["hapi"], mods["syn"]
    g = golden("doppler.npz")
    cols = syn.synthetic_lines(91, 80, 1495.0, 1525.0, 1)
    hapi.storeTable("dop", cols)
    grid = np.linspace(1500.0, 1520.0, 20001)
    for k, (T, p, kw) in enumerate(mg.DOPPLER_CASES):
        _, xs = hapi.absorptionCoefficient_Doppler(SourceTables="dop", Environment={"T": T, "p": p}, OmegaGrid=grid, **kw)
        ref = np.zeros_like(xs); ref[g["idx%d" % k]] = g["xs%d" % k]
        assert_parity(xs, ref, RTOL, FLOOR, "Doppler %d" % k)
    assert hapi.absorptionCoefficient_Gauss is hapi.absorptionCoefficient_Doppler


def test_continuum_hook(mods):
    """opts['continuum_od'] (array or callable) is added to the layer optical depths before the layer integration:
    a per-layer constant c_l scales the transmittance by exp(-sum of c_l below the sensor), and the result equals the
    oracle's layer integration of OD + continuum."""
    rt, port = mods["rt"], mods["port"]
    atm, tabs, opts, kw = _tud_setup(mods)
    nL = atm["Ts"].size
    X, OD0, _, _ = rt.compute_TUD(700.0, 720.0, opts=dict(opts), returnOD=True, Altitudes=np.array([500.0]), **kw)
    c_layer = 1e-3 * (1.0 + np.arange(nL))
    Xc, ODc, Luc, Ldc = rt.compute_TUD(700.0, 720.0, opts=dict(opts), returnOD=True, Altitudes=np.array([500.0]),
                                       continuum_od=c_layer, **kw)
    assert_parity(ODc, OD0 + c_layer.sum(), RTOL, FLOOR, "OD with per-layer continuum")
    ramp = np.linspace(0.0, 0.01, X.size)
    calls = []

    def cont(Xg, Ts, Ps, PLs, MF, ID):
        calls.append(np.asarray(Ts).size)
        return np.outer(np.ones(nL), ramp)

    Xf, ODf, _, _ = rt.compute_TUD(700.0, 720.0, opts=dict(opts), returnOD=True, Altitudes=np.array([500.0]),
                                   continuum_od=cont, **kw)
    assert calls == [nL]
    assert_parity(ODf, OD0 + nL * ramp, RTOL, FLOOR, "OD with callable continuum")


def test_voigt_tile_level_far_field(mods):
    """Tile-level far-field blocks (a line whose window covers a whole 2048-point tile from >= 0.375 tile widths away is
    summed once at the tile's 32 Chebyshev nodes): with wings of thousands of grid points -- the regime of the
    reference's own cross-section script, WavenumberWingHW=350 (misc/RT_gen_AbsXS_files.py:18,90) -- the kernel variant
    that carries them is chosen on the device.  It agrees with the slice-level-only evaluation (RTX_VOIGT_NOTILEFAR) and
    with the node-free one to 2e-13, with the oracle to the parity tolerance, on uniform, two-spacing and jittered
    grids; the speed-dependent / extra-diluent instantiation takes the same path."""
    import make_golden as mg
    port, syn, engine, hapi = mods["port"], mods["syn"], mods["engine"], mods["hapi"]
    cols = syn.synthetic_lines(61, 300, 960.0, 1040.0, 1)
    cols["sw"][::41] *= 1e6
    db = engine.LineDB(cols)
    states = [(296.0, 1.0), (250.0, 0.3)]
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    rng = np.random.default_rng(9)
    uniform = np.linspace(975.0, 1025.0, 100001)                       # 0.0005 cm^-1: 48 full tiles and a partial one
    two_step = np.concatenate([np.linspace(975.0, 1000.0, 75001)[:-1], np.linspace(1000.0, 1025.0, 25001)])
    jitter = np.sort(uniform + rng.uniform(-1e-4, 1e-4, uniform.size))
    K8 = (8 << 8) | 32768          # RTX_VOIGT_TILEFAR: the table is sparse, the automatic choice would not take the variant
    for name, grid, wing in (("uniform", uniform, 8.0), ("two_step", two_step, 8.0), ("jitter", jitter, 6.0),
                             ("hw350", uniform, 0.0)):
        kw = dict(wing_abs=wing) if wing else dict(wing_hw=350.0)
        tile = db.voigt_sum(T, p, w, grid, flags=128 | K8, **kw).cpu().numpy()
        slice_only = db.voigt_sum(T, p, w, grid, flags=128 | 16384 | K8, **kw).cpu().numpy()
        exact = db.voigt_sum(T, p, w, grid, flags=128 | 4096 | K8, **kw).cpu().numpy()
        assert not np.array_equal(tile, slice_only), name                                   # path taken
        assert np.max(np.abs(tile - exact) / np.maximum(exact, 1e-300)) < 2e-13, name
        assert np.max(np.abs(slice_only - exact) / np.maximum(exact, 1e-300)) < 2e-13, name
        dflt = db.voigt_sum(T, p, w, grid, flags=K8, **kw).cpu().numpy()
        assert np.max(np.abs(dflt - exact) / np.maximum(exact, 1e-300)) < 5e-12, name       # plus the one-step reciprocal
        okw = dict(OmegaWing=wing) if wing else dict(OmegaWingHW=350.0)
        _, ref = port.absorption_coefficient_voigt(cols, T=states[0][0], p=states[0][1], OmegaGrid=grid, **okw)
        assert_parity(dflt[0], ref, RTOL, FLOOR, "tile-level far field, %s" % name)
    # Lorentz driver (closed form with h = 0, no Weideman region) through the same blocks
    lt = db.voigt_sum(T, p, w, uniform, flags=128 | K8 | 16, wing_abs=8.0).cpu().numpy()
    ls = db.voigt_sum(T, p, w, uniform, flags=128 | K8 | 16 | 16384, wing_abs=8.0).cpu().numpy()
    lx = db.voigt_sum(T, p, w, uniform, flags=128 | K8 | 16 | 4096, wing_abs=8.0).cpu().numpy()
    assert not np.array_equal(lt, ls)
    assert np.max(np.abs(lt - lx) / np.maximum(lx, 1e-300)) < 2e-13
    _, ref = port.absorption_coefficient_lorentz(cols, T=states[0][0], p=states[0][1], OmegaGrid=uniform, OmegaWing=8.0)
    assert_parity(lt[0], ref, RTOL, FLOOR, "tile-level far field, Lorentz")
    # narrow wings on the same grid: the statistic keeps the variant without tile-level blocks (identical bits)
    a = db.voigt_sum(T, p, w, uniform, flags=8 << 8, wing_abs=1.0, wing_hw=10.0).cpu().numpy()
    b = db.voigt_sum(T, p, w, uniform, flags=(8 << 8) | 16384, wing_abs=1.0, wing_hw=10.0).cpu().numpy()
    assert np.array_equal(a, b)
    # ... and so does a sparse table whatever the wing (few hundred lines per tile: the prologue bounds the CTA)
    a = db.voigt_sum(T, p, w, uniform, flags=8 << 8, wing_abs=8.0).cpu().numpy()
    b = db.voigt_sum(T, p, w, uniform, flags=(8 << 8) | 16384, wing_abs=8.0).cpu().numpy()
    assert np.array_equal(a, b)
    # speed-dependent lines and an extra diluent with the reference script's wing (extended instantiation)
    sd = dict(mg.sdv_table())
    sd["SD_air"] = np.where(np.arange(300) % 2, sd["SD_air"], 0.0)     # lines without speed dependence take the Voigt path
    hapi.storeTable("sdv_wide", sd)
    fine = np.linspace(1235.0, 1265.0, 60001)
    dil = {"air": 0.7, "h2": 0.3}
    _, xs = hapi.absorptionCoefficient_SDVoigt(SourceTables="sdv_wide", Environment={"T": 280.0, "p": 0.8},
                                               OmegaGrid=fine, Diluent=dil, WavenumberWingHW=350, _profile_flags=K8)
    _, xs0 = hapi.absorptionCoefficient_SDVoigt(SourceTables="sdv_wide", Environment={"T": 280.0, "p": 0.8},
                                                OmegaGrid=fine, Diluent=dil, WavenumberWingHW=350,
                                                _profile_flags=K8 | 16384)
    assert not np.array_equal(xs, xs0)
    _, ref = port.absorption_coefficient_sdvoigt(sd, T=280.0, p=0.8, OmegaGrid=fine, Diluent=dil, OmegaWingHW=350.0)
    assert_parity(xs, ref, RTOL, FLOOR, "SDVoigt, WavenumberWingHW=350")
    # the far wing of a speed-dependent line (difference of two asymptotes) is summed at the nodes as well: all lines
    # speed-dependent, node-free evaluation as the yardstick, then the oracle
    sd_all = dict(mg.sdv_table())
    hapi.storeTable("sdv_all", sd_all)
    res = {}
    for name, fl in (("tile", K8 | 128), ("slice", K8 | 128 | 16384), ("exact", K8 | 128 | 4096), ("default", K8)):
        _, res[name] = hapi.absorptionCoefficient_SDVoigt(SourceTables="sdv_all", Environment={"T": 280.0, "p": 0.8},
                                                          OmegaGrid=fine, Diluent=dil, WavenumberWingHW=350,
                                                          _profile_flags=fl)
    assert not np.array_equal(res["tile"], res["slice"]) and not np.array_equal(res["slice"], res["exact"])
    for name in ("tile", "slice"):
        # the difference w(Z1) - w(Z2) of the reference's formulation cancels ~2 digits more than the Voigt closed form,
        # at the nodes as at the grid points: rounding noise ~1e-12, not interpolation error
        assert np.max(np.abs(res[name] - res["exact"]) / np.maximum(res["exact"], 1e-300)) < 2e-11, name
    _, ref = port.absorption_coefficient_sdvoigt(sd_all, T=280.0, p=0.8, OmegaGrid=fine, Diluent=dil, OmegaWingHW=350.0)
    assert_parity(res["default"], ref, RTOL, FLOOR, "SDVoigt far wing at the nodes")


def test_wide_wings_against_reference_golden(mods):
    """Tile- and slice-level far-field blocks against outputs of the unmodified reference (tests/golden/widewing.npz):
    absorptionCoefficient_Voigt with OmegaWing = 8 cm^-1 and WavenumberWingHW = 350, absorptionCoefficient_SDVoigt with
    WavenumberWingHW = 350 (speed-dependent far wings at the nodes) -- through the shim, automatic choice, and with the
    tile-level variant forced (the 300-line table is too sparse for the automatic choice to take it)."""
    import make_golden as mg
    hapi = mods["hapi"]
    g = golden("widewing.npz")
    hapi.storeTable("wide", mg.widewing_table())
    hapi.storeTable("wide_sd", mg.sdv_table())
    grid = np.linspace(*mg.WIDE_GRID)
    T, p, dil = mg.WIDE_SD_CASE
    for fl in (0, (8 << 8) | 32768, (8 << 8) | 16384):
        _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="wide", Environment={"T": 296.0, "p": 1.0}, OmegaGrid=grid,
                                                 OmegaWing=8.0, _profile_flags=fl)
        assert_parity(xs, g["voigt_w8"], RTOL, FLOOR, "OmegaWing=8, flags %d" % fl)
        _, xs = hapi.absorptionCoefficient_Voigt(SourceTables="wide", Environment={"T": 250.0, "p": 0.3}, OmegaGrid=grid,
                                                 WavenumberWingHW=350.0, _profile_flags=fl)
        assert_parity(xs, g["voigt_hw350"], RTOL, FLOOR, "WavenumberWingHW=350, flags %d" % fl)
        _, xs = hapi.absorptionCoefficient_SDVoigt(SourceTables="wide_sd", Environment={"T": T, "p": p},
                                                   OmegaGrid=np.linspace(*mg.WIDE_SD_GRID), Diluent=dil,
                                                   WavenumberWingHW=350.0, _profile_flags=fl)
        assert_parity(xs, g["sd_hw350"], RTOL, FLOOR, "SDVoigt WavenumberWingHW=350, flags %d" % fl)


def test_voigt_tile_level_automatic_on_dense_table(mods):
    """The variant with tile-level blocks is picked on the device where it pays: config 2's dense table (262 lines per
    cm^-1) with 6 cm^-1 wings takes it by itself (different bits from RTX_VOIGT_NOTILEFAR), the default 50-half-width
    wings do not (identical bits); either way the result agrees with the node-free evaluation."""
    syn, engine, rt = mods["syn"], mods["engine"], mods["rt"]
    tabs = syn.config2_tables()
    db = engine.LineDB([tabs[m] for m in sorted(tabs)])
    atm = syn.standard_atmosphere()
    p, w = rt.layer_weights(db, atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    sl = slice(0, 3)
    T, pp, ww = atm["Ts"][sl], np.ascontiguousarray(p)[sl], w[sl]
    grid = rt.make_spectral_axis(700.0, 1500.0, 0.001)
    K8 = 8 << 8
    auto = db.voigt_sum(T, pp, ww, grid, flags=128 | K8, wing_abs=6.0).cpu().numpy()
    no_tile = db.voigt_sum(T, pp, ww, grid, flags=128 | 16384 | K8, wing_abs=6.0).cpu().numpy()
    exact = db.voigt_sum(T, pp, ww, grid, flags=128 | 4096 | K8, wing_abs=6.0).cpu().numpy()
    assert not np.array_equal(auto, no_tile)
    assert np.max(np.abs(auto - exact) / np.maximum(exact, 1e-300)) < 2e-13
    assert np.max(np.abs(no_tile - exact) / np.maximum(exact, 1e-300)) < 2e-13
    a = db.voigt_sum(T, pp, ww, grid, flags=K8).cpu().numpy()
    b = db.voigt_sum(T, pp, ww, grid, flags=K8 | 16384).cpu().numpy()
    assert np.array_equal(a, b)


def test_voigt_small_problem_line_split(mods):
    """Small problems (fewer (tile, state) pairs than three per SM -- a single cross-section): the candidate lines of a
    tile are split over up to eight CTAs whose partial spectra are added in part order.  Same result as the unsplit
    launch (RTX_VOIGT_NOSPLIT) to rounding, bit-reproducible, parity with the oracle; one-point and empty cases."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(71, 3000, 990.0, 1010.0, 1)
    db = engine.LineDB(cols)
    for nS, grid in ((1, np.linspace(995.0, 1005.0, 10001)), (3, np.linspace(999.0, 1001.0, 777)), (1, np.array([1000.0]))):
        T = np.linspace(296.0, 240.0, nS); p = np.geomspace(1.0, 0.1, nS)
        w = np.tile(db.iso_factors(None, 1.0), (nS, 1))
        a = db.voigt_sum(T, p, w, grid).cpu().numpy()
        b = db.voigt_sum(T, p, w, grid).cpu().numpy()
        c = db.voigt_sum(T, p, w, grid, flags=65536).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.max(np.abs(a - c) / np.maximum(c, 1e-300)) < 5e-12
        if grid.size > 1:
            assert not np.array_equal(a, c)                                   # the split path was taken
        for k in range(nS):
            _, ref = port.absorption_coefficient_voigt(cols, T=T[k], p=p[k], OmegaGrid=grid)
            assert_parity(a[k], ref, RTOL, FLOOR, "line split, %d states, %d points, state %d" % (nS, grid.size, k))


def test_voigt_automatic_tile_size(mods):
    """Automatic tile size: 2048-point tiles when the windows are thousands of grid points wide (fine grid, low
    layers), 1024-point tiles at four CTAs per SM when they are a few hundred (coarse grid) -- decided from the
    eval-weighted mean wing bound in grid points (rtx_voigt_last_tile reports the choice).  Either way parity holds."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(57, 300, 940.0, 1060.0, 1)
    db = engine.LineDB(cols)
    nS = 80                                      # enough (tile, state) pairs for 2048-point tiles on both grids
    T = np.linspace(296.0, 220.0, nS); p = np.geomspace(1.0, 0.05, nS)
    w = np.tile(db.iso_factors(None, 1.0), (nS, 1))
    picked = {}
    for name, grid in (("fine", np.linspace(950.0, 1050.0, 100001)), ("coarse", np.linspace(950.0, 1050.0, 20001))):
        out = db.voigt_sum(T, p, w, grid).cpu().numpy()
        picked[name] = db.handle.last_tile()
        for k in (0, nS - 1):
            _, ref = port.absorption_coefficient_voigt(cols, T=T[k], p=p[k], OmegaGrid=grid)
            assert_parity(out[k], ref, RTOL, FLOOR, "%s grid state %d" % (name, k))
    assert picked == {"fine": 8, "coarse": 4}, picked


def test_tud_negative_optical_depth(mods):
    """A continuum hook may hand in negative layer optical depths: every exp(-OD...) of the layer integration uses the
    signed value, like the reference (radiative_transfer.py:349-355, :371), in the upwelling, transmittance AND
    downwelling halves -- including the thin-layer fast path (|OD|/cos(th) <= 2^-9) and large negative values."""
    engine, port = mods["engine"], mods["port"]
    rng = np.random.default_rng(21)
    nL, nx = 20, 2100
    X = np.linspace(700.0, 1400.0, nx)
    OD = rng.uniform(-0.4, 1.0, (nx, nL)) * rng.uniform(0, 1, (nx, 1))
    OD[:700] *= 1e-3                                  # thin layers of both signs
    OD[700:710, 3] = -30.0                            # strongly negative: exp(+30/cos th)
    Z = np.linspace(0.0, 60.0, nL); T = rng.uniform(200.0, 300.0, nL)
    tau, lu, ld = engine.tud(np.ascontiguousarray(OD.T), T[None, :], X, Z, [7.0, 500.0], [1.0, 1.2], 30)
    t_ref, u_ref, d_ref = port.tud_from_od(X, OD, Z, T, N_angle=30, Altitudes=np.array([7.0, 500.0]), theta_r=0.0)
    assert_parity(tau[0, :, :, 0].cpu().numpy(), t_ref, RTOL, FLOOR, "tau negative OD")
    assert_parity(lu[0, :, :, 0].cpu().numpy(), u_ref, RTOL, FLOOR, "Lu negative OD")
    for sl in (slice(0, 700), slice(710, nx)):        # row-wise: the exp(+30) rows would swamp the floor term
        assert_parity(ld[0, sl].cpu().numpy(), d_ref[sl], RTOL, FLOOR, "Ld negative OD")
    assert_parity(ld[0, 700:710].cpu().numpy(), d_ref[700:710], RTOL, FLOOR, "Ld strongly negative OD")


def test_small_compute_tud_alternating(mods):
    """Small outputs take the single-staging-buffer path of _to_host with tau/Lu copied on a side stream: alternate two
    different atmospheres many times and require every call to return its own result (a missing wait on the side
    stream would hand back stale bytes of the previous call)."""
    rt, syn = mods["rt"], mods["syn"]
    atm, tabs, opts, kw = _tud_setup(mods)
    Ts, MF = syn.perturbed_atmospheres(7, 2)
    want = [rt.compute_TUD(700.0, 730.0, opts=dict(opts), **dict(kw, Ts=Ts[j], MFs_VAL=MF[j])) for j in (0, 1)]
    assert not np.array_equal(want[0][1], want[1][1])
    for it in range(40):
        j = it & 1
        got = rt.compute_TUD(700.0, 730.0, opts=dict(opts), **dict(kw, Ts=Ts[j], MFs_VAL=MF[j]))
        for a, b in zip(got, want[j]):
            assert np.array_equal(a, b), (it, j)


def test_tud_split_halves(mods):
    """rtx_tud launched as two halves (tau+Lu, then Ld: NULL output pointers select the half) gives bit-identical
    results to the single launch; asking for tau without lu is refused."""
    import torch
    engine = mods["engine"]
    rng = np.random.default_rng(5)
    nL, nx = 12, 3000
    od = rng.uniform(0, 2.0, (nL, nx)) * rng.uniform(0, 1, (1, nx))
    T = np.linspace(290.0, 215.0, nL)[None, :]
    Z = np.linspace(0.5, 30.0, nL)
    X = np.linspace(700.0, 1400.0, nx)
    one = engine.tud(od, T, X, Z, [3.0, 500.0], [1.0, 1.3], 9)
    two = engine.tud(od, T, X, Z, [3.0, 500.0], [1.0, 1.3], 9, split=True)
    two[3].synchronize()
    for a, b in zip(one, two[:3]):
        assert torch.equal(a, b)
    from radtxfr_b200 import _lib
    with pytest.raises(RuntimeError):
        h = engine.default_handle(one[0].device)
        _lib.call("rtx_tud", h.h, 1, nL, engine._ptr(engine.to_device(od, one[0].device)), engine._ptr(engine.to_device(T, one[0].device)),
                  engine._ptr(engine.to_device(X, one[0].device)), nx, engine._hptr(Z), engine._hptr(np.array([3.0])), 1,
                  engine._hptr(np.array([1.0])), 1, 9, 0, engine._ptr(one[0]), None, None, engine._stream(one[0].device))


def test_voigt_core_slice_cancellation(mods):
    """Wide windows at very low pressure (OmegaWing forces +-1 cm^-1 windows while y = Gamma0*cte is 1e-2 ... 1e-5):
    the slice holding the Weideman core takes the test-free path and pass 2 adds series - closed form there.  The
    closed form is up to 1/(2 sqrt(pi) y) times the true core value, so this is the worst case for the cancellation;
    below y = 1e-4 the kernel keeps the tested path.  Parity must hold throughout."""
    port, syn, engine = mods["port"], mods["syn"], mods["engine"]
    cols = syn.synthetic_lines(61, 120, 995.0, 1005.0, 1)
    db = engine.LineDB(cols)
    grid = np.linspace(996.0, 1004.0, 8001)
    states = [(250.0, 1e-2), (240.0, 1e-3), (230.0, 2e-4), (220.0, 5e-5), (220.0, 1e-6)]
    T = np.array([s[0] for s in states]); p = np.array([s[1] for s in states])
    w = np.tile(db.iso_factors(None, 1.0), (len(states), 1))
    for K in (8, 4):
        out = db.voigt_sum(T, p, w, grid, wing_abs=1.0, flags=K << 8).cpu().numpy()
        for k, (Tk, pk) in enumerate(states):
            _, ref = port.absorption_coefficient_voigt(cols, T=Tk, p=pk, OmegaGrid=grid, OmegaWing=1.0)
            assert_parity(out[k], ref, RTOL, FLOOR, "K=%d p=%g atm" % (K, pk))


def test_c_abi_from_plain_c(mods, tmp_path):
    """examples/voigt_from_c.c: the whole Voigt path through the C ABI from plain C (no Python objects, no torch),
    checked against the oracle on the same three lines."""
    import subprocess
    import re as _re
    from test_host_logic import _build_c_example
    port = mods["port"]
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    peaks = [float(x) for x in _re.findall(r"peak ([0-9.e+-]+) at", r.stdout)]
    assert len(peaks) == 2 and all(p > 0 for p in peaks)
    # same problem through the oracle (isotopologue tables differ: the example uses its own Q(T) ~ T^1.5, so compare
    # the 296 K state, where Q(Tref)/Q(T) = 1 for any table)
    cols = dict(molec_id=np.array([1, 1, 1]), local_iso_id=np.array([1, 1, 1]), nu=np.array([999.2, 1000.0, 1000.9]),
                sw=np.array([1e-21, 3e-21, 5e-22]), elower=np.array([100.0, 250.0, 900.0]),
                gamma_air=np.array([0.07, 0.08, 0.06]), n_air=np.array([0.7, 0.65, 0.75]),
                delta_air=np.array([-0.002, 0.0, 0.001]))
    grid = 999.0 + 2.0 * np.arange(20001) / 20000.0
    _, ref = port.absorption_coefficient_voigt(cols, T=296.0, p=1.0, OmegaGrid=grid, Components=[(1, 1, port.tables()["iso_abundance"][port.tables()["index"][(1, 1)]])])
    assert abs(peaks[0] / ref.max() - 1.0) < 1e-5, (peaks[0], ref.max())       # the example prints 7 digits


def test_tud_from_plain_c(mods, tmp_path):
    """examples/tud_from_c.c: the whole path -- line table + atmosphere table in, (tau, Lu, Ld) out through
    rtx_tud_from_lines, then rtx_mako_bands + rtx_ils_apply -- from plain C, compared with the CPU oracle on the same
    inputs (the example's synthetic partition functions are handed to the oracle, which interpolates them with the
    reference's own AtoB routine)."""
    import subprocess
    from test_host_logic import _build_c_example
    port = mods["port"]
    exe = _build_c_example(tmp_path, "tud_from_c")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    tud = np.array([[float(x) for x in ln.split()[1:]] for ln in r.stdout.splitlines() if ln.startswith("TUD ")])
    bands = np.array([[float(x) for x in ln.split()[1:]] for ln in r.stdout.splitlines() if ln.startswith("BAND ")])
    assert tud.shape[0] >= 7 and bands.shape[0] >= 4
    # the same inputs for the oracle
    i = np.arange(40)
    cols = {1: {}, 2: {}}
    full = dict(nu=700.0 + 17.5 * i + 3.0 * np.sin(1.7 * i), sw=1e-20 * (1.0 + 0.5 * np.cos(0.9 * i)), elower=100.0 + 60.0 * i,
                gamma_air=0.05 + 0.03 * np.sin(0.3 * i) ** 2, n_air=0.5 + 0.2 * np.cos(0.5 * i) ** 2, delta_air=-0.002 + 0.0001 * i)
    for m in (1, 2):
        sel = (i % 2) == (m - 1)
        cols[m] = {k: v[sel] for k, v in full.items()}
        cols[m]["molec_id"] = np.full(sel.sum(), m); cols[m]["local_iso_id"] = np.full(sel.sum(), 1)
    tipsT = 70.0 + (3000.0 - 70.0) * np.arange(119) / 118
    qtab = {1: 170.0 * (tipsT / 296.0) ** 1.5, 2: 280.0 * (tipsT / 296.0)}
    t = port.tables()
    saved = (port.tips_q, t["iso_mass"].copy(), t["iso_abundance"].copy())
    try:
        port.tips_q = lambda M, I, T: port._lagrange(T, tipsT, qtab[M], 119)
        t["iso_mass"][t["index"][(1, 1)]] = 18.010565; t["iso_mass"][t["index"][(2, 1)]] = 43.98983
        t["iso_abundance"][t["index"][(1, 1)]] = 1.0; t["iso_abundance"][t["index"][(2, 1)]] = 1.0
        Z = np.array([0.5, 1.5, 3.0, 6.0, 12.0, 25.0]); T = np.array([290.0, 283.0, 272.0, 252.0, 217.0, 222.0])
        P = np.array([95000.0, 84000.0, 70000.0, 47000.0, 19000.0, 2500.0]); PL = np.array([1.0, 1.0, 2.0, 4.0, 8.0, 18.0])
        MF = np.stack([9000.0 / (1 + np.arange(6) ** 2), np.full(6, 400.0), np.full(6, 2.0e5)], axis=1)
        grid = 700.0 + (1400.0 - 700.0) * np.arange(14000) / 13999
        OD = np.stack([port.layer_od(cols, grid, T[l], P[l], PL[l], MF[l], np.array([1, 2, 7])) for l in range(6)], axis=1)
        tau, Lu, Ld = port.tud_from_od(grid, OD, Z, T)
        Xo, R = port.ils_mako_triangle(grid, np.stack([tau, Lu, Ld], axis=1))
    finally:
        port.tips_q = saved[0]; t["iso_mass"][:] = saved[1]; t["iso_abundance"][:] = saved[2]
    idx = tud[:, 0].astype(int)
    assert np.array_equal(tud[:, 1], grid[idx])
    for col, ref, name in ((2, tau, "tau"), (3, Lu, "Lu"), (4, Ld, "Ld")):
        assert_parity(tud[:, col], ref[idx], RTOL, FLOOR, "plain C " + name)
    b = bands[:, 0].astype(int)
    assert np.array_equal(bands[:, 1], Xo[b])
    assert_parity(bands[:, 2:5], R[b], RTOL, FLOOR, "plain C MAKO bands")


def test_mako_database_and_apparent_radiance_flows(mods):
    """The two production flows downstream of the TUD sweep: Generate_LWIR_TUD_MAKO.py:34-52 (Gaussian MAKO line shape on
    tau / La / Ld, atmospheres sorted by band-mean transmittance) and Compute_LWIR_Apparent_Radiance.py:9-26 (float32
    sweep over 41 surface-temperature offsets), against the oracle's restatement of the same steps."""
    from radtxfr_b200 import drivers
    port = mods["port"]
    rng = np.random.default_rng(31)
    X = np.linspace(700.0, 1400.0, 3501)
    nA = 23
    tau = np.clip(np.cumsum(rng.normal(0, 0.01, (X.size, nA)), axis=0) + rng.uniform(0.3, 0.9, nA), 0.0, 1.0)
    La = rng.uniform(0, 300, (X.size, nA)); Ld = rng.uniform(0, 600, (X.size, nA)); Ts = rng.uniform(270, 310, nA)
    Xo, t, a, d, Tso = drivers.generate_LWIR_TUD_MAKO(X, tau, La, Ld, Ts)
    Xr, tr = port.ils_mako_gauss(X, tau); _, ar = port.ils_mako_gauss(X, La); _, dr = port.ils_mako_gauss(X, Ld)
    ix = np.argsort(np.mean(tr, axis=0)); ixx = np.argsort(Xr)
    assert np.array_equal(Xo, Xr[ixx]) and np.array_equal(Tso, Ts[ix])
    for got, ref, name in ((t, tr, "tau"), (a, ar, "La"), (d, dr, "Ld")):
        assert_parity(got, ref[:, ix][ixx, :], RTOL, FLOOR, "MAKO database " + name)
    emis = mods["syn"].synthetic_emissivities(5, 7, Xo)
    L, T = drivers.compute_LWIR_apparent_radiance_sweep(Xo, emis, Tso, t, a, d)
    assert L.dtype == np.float32 and L.shape == (Xo.size, 7, nA, 41) and T.shape == (nA, 41) and T.dtype == np.float32
    f32 = [np.asarray(v).astype(np.float32).astype(float) for v in (Xo, emis, Tso, t, a, d)]
    ref = port.apparent_radiance(*f32, dT=np.arange(-10, 10.5, 0.5).astype(np.float32).astype(float))
    assert_parity(L.astype(float), ref, 1e-5, 1e-8, "float32 apparent-radiance sweep")
