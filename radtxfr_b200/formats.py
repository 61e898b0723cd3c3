"""On-disk formats of the reference's cross-section look-up-table generator (misc/RT_gen_AbsXS_files.py) and a
batched, device-side version of its T x P x molecule loop."""
import numpy as np

from . import hapi_compat


def AFIT_XS_write(X, Y, T, P, ID, fnDB, File=[]):
    """AFIT_XS binary cross-section file, byte for byte as misc/RT_gen_AbsXS_files.py:45-83 writes it:
    'v1' (2 bytes), six little-endian doubles [X.min, X.max, X.size, ID, T, P], the database name padded to
    128 bytes, then Y as little-endian doubles.  Returns the file name
    (default XS-<ID:02d>-<T:04d>K-<P:06d>Pa.bin)."""
    X = np.asarray(X)
    Y = np.asarray(Y)
    version = np.array("v1", "<S2")
    params = np.array([X.min(), X.max(), X.size, ID, T, P], "<f8")
    db = np.array(fnDB, "<S128")
    if len(File) == 0:
        File = "XS-{0:02d}-{1:04d}K-{2:06d}Pa.bin".format(int(ID), int(T), int(P))
    with open(File, "wb") as f:
        version.tofile(f)
        params.tofile(f)
        db.tofile(f)
        Y.astype("<f8").tofile(f)
    return File


def AFIT_XS_read(File):
    """Inverse of AFIT_XS_write: returns (X, Y, T, P, ID, fnDB)."""
    raw = open(File, "rb").read()
    assert raw[:2] == b"v1", "not an AFIT_XS v1 file"
    xmin, xmax, n, ID, T, P = np.frombuffer(raw, "<f8", 6, 2)
    name = raw[50:178].rstrip(b"\x00").decode()
    Y = np.frombuffer(raw, "<f8", int(n), 178).copy()
    return np.linspace(xmin, xmax, int(n)), Y, float(T), float(P), int(ID), name


def _matlab_header_blob(**kwargs):
    """Header variables serialised as a MATLAB v5 file in memory (AFIT_MBI.py:18-28)."""
    import io
    import scipy.io as sio
    H = {key: val for key, val in kwargs.items()}
    f = io.BytesIO()
    sio.savemat(f, {"Header": H})
    return f.getvalue()


def AFIT_MBI_export(fname, Data, Rows=[], Columns=[], Bands=[], Sidecar=True, BandsFirstDim=True, Verbose=True,
                    **kwargs):
    """BIP / BSQ multiband image export with the AFIT_MBI header, same arguments and byte layout as the reference's
    AFIT_MBI_export (AFIT_MBI.py:45-156): uint32 dims (rows, columns, bands), 6-char dtype string, 3-char
    interleave, uint32 offset of the pixel data, float64 rows|columns|bands, the MATLAB-v5 header blob, then the
    pixels band-sequential ('.bsq') or band-interleaved-by-pixel ('.bip').  (The blob carries scipy's creation
    timestamp in its 116-byte text header, so two exports differ in those bytes only.)  The output cube of
    engine.hsi_pixels / config 4 goes through this writer."""
    import os
    import sys
    Data = np.array(Data)
    if BandsFirstDim:
        Data = Data.transpose([1, 2, 0]).copy(order="C")
    nR, nC, nB = Data.shape
    if len(Bands) == 0: Bands = np.arange(nB)
    if len(Rows) == 0: Rows = np.arange(nR)
    if len(Columns) == 0: Columns = np.arange(nC)
    if len(kwargs) == 0: kwargs = {"MetaData": "None"}
    fdir, base = os.path.split(fname)
    interleave = os.path.splitext(base)[1][1:].lower()
    s = str(Data.dtype)
    dtypeStr = "{0:<6}".format({"float64": "double", "float32": "single"}.get(s, s))
    rcb = np.concatenate((Rows, Columns, Bands)).astype("float64")
    blob = _matlab_header_blob(**kwargs)
    offsetBytes = np.array(12 + 6 + 3 + 4 + len(blob) + rcb.nbytes)
    if Data.dtype.byteorder == ">" or not (sys.byteorder == "little" and Data.dtype.byteorder in ("=", "|")):
        Data = Data.byteswap()
    with open(os.path.join(fdir, base) if fdir else base, "wb") as f:
        np.array(Data.shape).astype("uint32").tofile(f)
        f.write(bytes(dtypeStr, "utf-8"))
        f.write(bytes(interleave, "utf-8"))
        offsetBytes.astype("uint32").tofile(f)
        rcb.tofile(f)
        f.write(blob)
        if interleave == "bsq":
            np.ascontiguousarray(Data.transpose(2, 0, 1)).tofile(f)        # band by band, rows x columns each
        elif interleave == "bip":
            np.ascontiguousarray(Data).tofile(f)                           # pixel by pixel, all bands of a pixel
        else:
            print("Unknown file extension -- no data saved")
    if Verbose:
        print("Wrote {0} bytes ({1} header, {2} data) to {3}".format(Data.nbytes + offsetBytes, offsetBytes,
                                                                     Data.nbytes, base))


def cross_section_lut(SourceTables, temperatures, pressures_atm, OmegaGrid, **voigt_kwargs):
    """Cross-sections for every (T, p) pair in one launch: the loop of misc/RT_gen_AbsXS_files.py:87-93
    (`for t in T: for p in P: absorptionCoefficient_SDVoigt(...)`) batched over states.
    Returns sigma [nT, nP, nX] (numpy) on np.sort(OmegaGrid); keyword arguments as absorptionCoefficient_Voigt
    (OmegaWing, OmegaWingHW / WavenumberWingHW, IntensityThreshold, GammaL, Diluent, Components)."""
    names = SourceTables if type(SourceTables) in (list, tuple) else [SourceTables]
    db = hapi_compat.line_database(names)
    Ts = np.asarray(temperatures, dtype=float).ravel()
    Ps = np.asarray(pressures_atm, dtype=float).ravel()
    TT, PP = np.meshgrid(Ts, Ps, indexing="ij")
    grid = np.sort(np.asarray(OmegaGrid, dtype=float))
    comps = voigt_kwargs.get("Components")
    w = np.tile(db.iso_factors(comps, 1.0), (TT.size, 1))
    GammaL = voigt_kwargs.get("GammaL", "gamma_air").lower()
    dil = voigt_kwargs.get("Diluent") or ({"air": 1.0} if GammaL == "gamma_air" else {"self": 1.0})
    x_air = float(sum(v for k, v in dil.items() if k.lower() == "air"))
    x_self = float(sum(v for k, v in dil.items() if k.lower() == "self")) if db.has_self else 0.0
    hw = voigt_kwargs.get("WavenumberWingHW", voigt_kwargs.get("OmegaWingHW", 50.0))
    wing = voigt_kwargs.get("WavenumberWing", voigt_kwargs.get("OmegaWing", 0.0)) or 0.0
    db.check_temperatures(Ts)
    out = db.voigt_sum(TT.ravel(), PP.ravel(), w, grid, x_air=x_air, x_self=x_self, wing_abs=float(wing),
                       wing_hw=float(hw), s_threshold=float(voigt_kwargs.get("IntensityThreshold", 0.0)))
    return out.view(Ts.size, Ps.size, grid.size).cpu().numpy()
