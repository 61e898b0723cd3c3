"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host-side shim logic, synthetic
workload generators.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from radtxfr_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "radtxfr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rtx_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
    from radtxfr_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding and header disagree"


def test_fails_loudly_without_gpu(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = L.rtx_create(0, ctypes.byref(h))
    assert rc < 0 and b"no CPU fallback" in L.rtx_last_error()
    from radtxfr_b200 import hapi_compat, _lib, synthetic
    hapi_compat.storeTable("x", synthetic.synthetic_lines(0, 10, 800, 900))
    with pytest.raises(_lib.RtxError):
        hapi_compat.absorptionCoefficient_Voigt(SourceTables="x", OmegaGrid=np.linspace(800, 900, 11))


def test_weideman_coefficients_match_fft_recipe(L):
    from oracle import port
    a = np.zeros(24)
    L.rtx_weideman_coeffs(a.ctypes.data_as(ctypes.c_void_p))
    ref, Lw = port.weideman_coeffs(24)
    assert np.array_equal(a, ref)
    src = open(os.path.join(ROOT, "radtxfr_b200", "csrc", "rtx_common.cuh")).read()
    assert repr(Lw) in src


def test_null_and_state_errors(L):
    assert L.rtx_voigt_sum(None, 1, None, None, None, 1.0, 0.0, 0.0, 50.0, 0.0, None, 1, None, 0, None) < 0
    assert b"null handle" in L.rtx_last_error()
    assert L.rtx_destroy(None) == 0


def test_grids_and_bands():
    from radtxfr_b200 import radiative_transfer as rt, hapi_compat
    from oracle import port
    X = rt.make_spectral_axis(700.0, 1500.0, 0.1)
    assert X.size == 8000 and abs((X[1] - X[0]) - 0.1000125) < 1e-6
    assert np.array_equal(X, port.make_spectral_axis(700.0, 1500.0, 0.1))
    g = hapi_compat._hapi_grid((800.0, 1300.0), 0.01)
    assert g.size == 50001 and g[0] == 800.0
    Xo, c, s = rt._mako_bands(np.array([700.0, 1400.0]), None, 1.0, 0.0, 1.0)
    assert Xo.size == 128 and abs(Xo[0] - 759.91306595) < 1e-7 and abs(Xo[-1] - 1320.81203524) < 1e-7
    assert 4.0 < s.min() < 4.01 and 12.40 < s.max() < 12.41
    assert np.array_equal(Xo, port.mako_axis(np.array([700.0, 1400.0])))
    Xo2, _, _ = rt._mako_bands(np.array([700.0, 1400.0]), 2, 1.0, 0.0, 1.0)
    assert Xo2.size == 256
    Xo3, _, _ = rt._mako_bands(np.array([800.0, 1200.0]), None, 1.0, 0.0, 1.0)
    assert 0 < Xo3.size < 128 and Xo3.min() > 800 and Xo3.max() < 1200


def test_synthetic_workloads_are_seeded():
    from radtxfr_b200 import synthetic as syn
    a, b = syn.config1_lines(500), syn.config1_lines(500)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert np.all(np.diff(a["nu"]) >= 0) and a["nu"].min() >= 775 and a["nu"].max() <= 1325
    atm = syn.standard_atmosphere()
    assert atm["Ts"].shape == (66,) and atm["MFs_VAL"].shape == (66, 8) and atm["Ts"][0] == 287.87
    Ts, MF = syn.perturbed_atmospheres(2, 7)
    assert Ts.shape == (7, 66) and MF.shape == (7, 66, 8) and Ts.min() >= 180 and Ts.max() <= 330
    assert np.array_equal(MF[:, :, 1], np.repeat(atm["MFs_VAL"][None, :, 1], 7, axis=0))   # CO2 fixed
    t = syn.config2_tables(scale=0.001)
    m = syn.merge_tables(t)
    assert len(m["nu"]) == sum(len(v["nu"]) for v in t.values()) and np.all(np.diff(m["nu"]) >= 0)
    e = syn.synthetic_emissivities(3, 5, np.linspace(760, 1320, 128))
    assert e.shape == (128, 5) and e.min() >= 1e-4 and e.max() <= 1 - 1e-4


def test_layer_weights_match_oracle_wiring():
    """Host-side column amounts (no GPU): compare with the factor the oracle applies per molecule."""
    from radtxfr_b200 import hapi_compat
    from oracle import port
    for p, T in ((1.0, 296.0), (0.3, 250.0)):
        assert hapi_compat.volumeConcentration(p, T) == port.volume_concentration(p, T)
    assert hapi_compat.abundance(1, 1) == 0.997317 and hapi_compat.molecularMass(2, 1) == 43.98983


def test_afit_xs_bytes_match_reference_writer(tmp_path):
    """AFIT_XS writer against bytes produced by the reference's own AFIT_XS_write (tests/golden/lut.npz)."""
    from conftest import golden
    from radtxfr_b200 import formats
    g = golden("lut.npz")
    grid = np.linspace(2000.0, 2350.0, 7001)
    f = formats.AFIT_XS_write(grid, g["xs0"], 296.0, 101325 * 1.0, 2, "HITRAN2016-synthetic", File=str(tmp_path / "x.bin"))
    assert np.array_equal(np.frombuffer(open(f, "rb").read(), dtype=np.uint8), g["afit_bytes"])
    assert formats.AFIT_XS_write(grid[:3], g["xs0"][:3], 296.4, 101325.0, 2, "db", File=str(tmp_path / "y.bin")).endswith("y.bin")
    import os
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        assert formats.AFIT_XS_write(grid[:3], g["xs0"][:3], 296.4, 101325.0, 2, "db") == "XS-02-0296K-101325Pa.bin"
    finally:
        os.chdir(cwd)


def test_afit_mbi_export_matches_reference_bytes(tmp_path):
    """BIP/BSQ cube writer against files written by the reference's AFIT_MBI_export (tests/golden/mbi.npz); the only
    bytes allowed to differ are the 116-byte text header of the embedded MATLAB blob (scipy's timestamp)."""
    import os, sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden as mg
    from conftest import golden
    from radtxfr_b200 import formats
    g = golden("mbi.npz")
    img = mg.mbi_image()
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        formats.AFIT_MBI_export("a.bsq", img, Verbose=False, Q=np.array([1, 2, 3]), Meta1=2, myString="hello")
        formats.AFIT_MBI_export("b.bip", img, Verbose=False)
        formats.AFIT_MBI_export("c.bip", g["cube"], Verbose=False, Bands=np.array([7.5, 8.5, 9.5, 10.5]))
        for name, nrcb in (("a.bsq", 40 + 32 + 3), ("b.bip", 40 + 32 + 3), ("c.bip", 6 + 5 + 4)):
            got = np.frombuffer(open(name, "rb").read(), dtype=np.uint8).copy()
            ref = g[name.replace(".", "_")].copy()
            assert got.size == ref.size, name
            blob0 = 12 + 6 + 3 + 4 + 8 * nrcb            # start of the MATLAB blob
            got[blob0:blob0 + 116] = 0; ref[blob0:blob0 + 116] = 0
            assert np.array_equal(got, ref), name
    finally:
        os.chdir(cwd)


def test_jacobian_inputs():
    """JacIn / Jacobian sweep inputs as Generate_LWIR_TUD.py:55-71 builds them."""
    from radtxfr_b200 import drivers
    x = np.array([1.0, -4.0, 2.0])
    J = drivers.JacIn(x, 0.1, rel=True)
    assert J.shape == (3, 3) and np.allclose(np.diag(J), x + 0.4) and np.array_equal(J - np.diag(np.diag(J)), np.tile(x, (3, 1)) - np.diag(x))
    J2 = drivers.JacIn(x, 0.5)
    assert np.allclose(np.diag(J2), x + 0.5)
    T, H, O = drivers.jacobian_inputs(np.array([280.0, 250.0]), np.array([1e-2, 1e-4]), np.array([1e-8, 1e-6]))
    assert T.shape == (7, 2) and np.array_equal(T[0], [280.0, 250.0]) and T[1, 0] == 280.0 + 0.28 and T[2, 1] == 250.0 + 0.28
    assert np.array_equal(T[3:], np.tile([280.0, 250.0], (4, 1))) and H[3, 0] == 1e-2 + 1e-5 and O[6, 1] == 1e-6 + 1e-9
    alt = drivers.default_altitudes(70.0)
    assert alt.size == 9 and abs(alt[0] - 0.06096) < 1e-12 and alt[-1] == 70.0


def test_lwir_tud_products(tmp_path):
    """npz checkpoint of the production sweep keeps the reference's keys (Generate_LWIR_TUD.py:150); the HDF5 writer
    carries the reference's attribute table and refuses loudly without h5py."""
    from radtxfr_b200 import drivers
    rng = np.random.default_rng(3)
    X = np.linspace(700, 1400, 50); OD = rng.random((4, 9, 50)); La = rng.random((4, 9, 50)); Ld = rng.random((4, 50))
    alt = drivers.default_altitudes(100.0)
    f = str(tmp_path / "LWIR_TUD.npz")
    drivers.save_LWIR_TUD_npz(f, X, OD, La, Ld, alt)
    back = drivers.load_LWIR_TUD_npz(f)
    assert sorted(back) == ["Altitudes", "La", "Ld", "OD", "X"]
    assert np.array_equal(back["OD"], OD) and np.array_equal(back["Altitudes"], alt)
    assert set(drivers.TUD_DATASET_ATTRS) == {"X", "OD", "La", "Ld", "SensorAltitude", "StdAtmos", "z", "T", "P", "H2O", "O3"}
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            drivers.save_LWIR_TUD_h5(str(tmp_path / "x.h5"), X, OD, La, Ld, alt, OD, X, X, X, X, X)


def test_hapi_spectrum_helpers_golden():
    """transmittanceSpectrum / absorptionSpectrum / radianceSpectrum against the reference's outputs."""
    from radtxfr_b200 import hapi_compat as hapi
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spectra.npz"))
    env = {"l": 250.0, "T": 283.0}
    for fn, key in ((hapi.transmittanceSpectrum, "trans"), (hapi.absorptionSpectrum, "absorp"), (hapi.radianceSpectrum, "radi")):
        nu, y = fn(g["nu"], g["coef"], Environment=env)
        assert np.array_equal(nu, g["nu"])
        np.testing.assert_allclose(y, g[key], rtol=1e-14, atol=0)


def _build_c_example(tmp_path, name="voigt_from_c"):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "radtxfr_b200")
    exe = str(tmp_path / name)
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", name + ".c"), "-L", libdir, "-lradtxfr_b200",
                    "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def test_c_abi_from_plain_c_without_gpu(tmp_path):
    """The header compiles as C11 and the library links from plain C; without a CUDA device rtx_create fails loudly
    (exit status 77 of the example, with the library's own error text) -- no CPU fallback."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test")
    for name in ("voigt_from_c", "tud_from_c"):
        exe = _build_c_example(tmp_path, name)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 77, (name, r.returncode, r.stdout, r.stderr)
        assert "no CPU fallback" in r.stdout


def test_mako_bands_c_abi_matches_reference_derivation():
    """rtx_mako_bands (host only, no GPU needed) against the reference's derivation restated in the oracle
    (radiative_transfer.py:1226-1241, ILS_MAKO.py:20-22): band axis bit-identical, incl. resFactor, clip, empty result."""
    from oracle import port
    from radtxfr_b200 import engine
    X = np.linspace(700.0, 1400.0, 701)
    for rf, fw, sh, sc in ((None, 1.0, 0.0, 1.0), (2, 0.8, 0.3, 1.001), (0.5, 1.3, -0.2, 0.999)):
        Xo, c, w = engine.mako_bands(0, X.min(), X.max(), rf, fw, sh, sc)
        ref = port.mako_axis(X, rf, clip=True)
        assert np.array_equal(Xo, ref) and np.array_equal(c, sc * ref + sh)
        assert np.array_equal(w, fw * np.abs(np.gradient(ref)) * 1.6)
    Xc = np.linspace(900.0, 1000.0, 11)
    Xo, c, w = engine.mako_bands(0, 900.0, 1000.0)
    assert np.array_equal(Xo, port.mako_axis(Xc, None, clip=True)) and 0 < Xo.size < 128
    assert engine.mako_bands(0, 10.0, 20.0)[0].size == 0
    Xg, cg, wg = engine.mako_bands(1, 0.0, 0.0)
    ref = port.mako_axis(X, None, clip=False)
    assert np.array_equal(Xg, ref) and np.array_equal(cg, ref) and np.array_equal(wg, np.abs(np.gradient(ref)))


def test_flag_literals_match_header():
    """The shim passes the RTX_VOIGT_* bits as literals; they must equal the header's values."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "radtxfr_b200.h")).read()
    val = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (RTX_VOIGT_[A-Z0-9_]+) (\d+)", hdr)}
    assert val["RTX_VOIGT_FAST_RCP"] == 1 and val["RTX_VOIGT_FP32"] == 2 and val["RTX_VOIGT_WARPSPEC"] == 4
    assert val["RTX_VOIGT_LORENTZ"] == 16 and val["RTX_VOIGT_DOPPLER"] == 32 and val["RTX_VOIGT_NOSHIFT"] == 64
    assert val["RTX_VOIGT_TILE_SHIFT"] == 8
    shim = open(os.path.join(root, "radtxfr_b200", "hapi_compat.py")).read()
    assert 'kwargs["_profile_flags"] = 16' in shim and "_profile_flags=32 | (0 if LineShift else 64)" in shim
    assert val["RTX_VOIGT_EXACT_RCP"] == 128 and val["RTX_VOIGT_NOFAR"] == 4096
    import types
    eng_src = open(os.path.join(root, "radtxfr_b200", "engine.py")).read()
    consts = dict(re.findall(r"^(VOIGT_[A-Z0-9_]+(?:, VOIGT_[A-Z0-9_]+)*) = (.+)$", eng_src, flags=re.M))
    ns = types.SimpleNamespace()
    for names, vals in consts.items():
        for n, v in zip(names.split(", "), vals.split(", ")):
            setattr(ns, n, int(v))
    for n in ("FP32", "WARPSPEC", "LORENTZ", "DOPPLER", "NOSHIFT", "EXACT_RCP", "NOFAR", "TILE_SHIFT", "SDVOIGT", "NOTILEFAR", "TILEFAR", "NOSPLIT"):
        assert getattr(ns, "VOIGT_" + n) == val["RTX_VOIGT_" + n], n
    assert "engine.PRECISION_FLAGS[" in shim
    rtm = open(os.path.join(root, "radtxfr_b200", "radiative_transfer.py")).read()
    assert "<< 8" in rtm and "flags |= 4 if" in rtm
