"""Batch driver for the reference's production sweep (Generate_LWIR_TUD.py): TUD for many atmospheres at several
sensor altitudes, reduced to a coarser spectral resolution -- one device pipeline instead of a multiprocessing.Pool
of compute_TUD calls followed by three reduceResolution calls per atmosphere (Generate_LWIR_TUD.py:117-150)."""
import numpy as np

from . import engine, radiative_transfer as rt

# Altitudes of the reference sweep, [km] from [ft] plus the top of the profile (Generate_LWIR_TUD.py:74)
def default_altitudes(z_top_km):
    return np.concatenate((np.array([200, 500, 1000, 2000, 5000, 10000, 20000, 50000]) * 0.3048 / 1e3, [z_top_km]))


def JacIn(X, dX, rel=False):
    """One-at-a-time perturbations of a profile: row i is X with element i stepped by dX (scalar; relative to
    max|X| when rel) -- Generate_LWIR_TUD.py:55-64."""
    X = np.asarray(X, dtype=float)
    X_out = np.tile(X, (X.shape[-1], 1))
    if np.isscalar(dX):
        dX = dX * np.ones_like(X_out) * np.max(np.abs(X_out)) if rel else dX + np.zeros_like(X_out)
    for ii in np.arange(len(X_out)):
        X_out[ii, ii] = X_out[ii, ii] + dX[ii, ii]
    return X_out


def jacobian_inputs(Tm, H2Om, O3m, relStep=0.001):
    """Inputs of the TUD Jacobian about mean profiles: 1 + 3*nL atmospheres (Generate_LWIR_TUD.py:66-71)."""
    Tm, H2Om, O3m = (np.asarray(a, dtype=float) for a in (Tm, H2Om, O3m))
    nL = Tm.size
    T_J, H2O_J, O3_J = (np.tile(x, (nL * 3 + 1, 1)) for x in (Tm, H2Om, O3m))
    T_J[1 + np.arange(0, nL), :] = JacIn(Tm, relStep, rel=True)
    H2O_J[1 + np.arange(nL, 2 * nL), :] = JacIn(H2Om, relStep, rel=True)
    O3_J[1 + np.arange(2 * nL, 3 * nL), :] = JacIn(O3m, relStep, rel=True)
    return T_J, H2O_J, O3_J


def generate_LWIR_TUD(T, H2O, O3, CO2, Xmin=690, Xmax=1410, DV=0.0005, DV_out=0.25, Altitudes=None, opts=None,
                      atm_chunk=None):
    """For every atmosphere jj (rows of T, H2O, O3 [mixing fractions]; CO2 profile shared) evaluate
        compute_TUD(Xmin, Xmax, MFs_VAL=1e6*[H2O[jj], CO2, O3[jj]].T, MFs_ID=[1,2,3], Ts=T[jj], DVOUT=DV,
                    Altitudes=Altitudes, returnOD=True)
    and reduce OD, La, Ld to resolution DV_out on a common axis, as Generate_LWIR_TUD.py:79-132 does.
    Returns numpy arrays X [nXo], OD [nXo, nAlt, nAtm], La [nXo, nAlt, nAtm], Ld [nXo, nAtm]."""
    torch = engine._torch()
    opts = dict(rt.options if opts is None else opts)
    T = np.atleast_2d(np.asarray(T, dtype=float))
    H2O = np.atleast_2d(np.asarray(H2O, dtype=float))
    O3 = np.atleast_2d(np.asarray(O3, dtype=float))
    CO2 = np.asarray(CO2, dtype=float)
    n_atm, nL = T.shape
    if Altitudes is None:
        Altitudes = default_altitudes(float(np.max(opts["Zs"])))
    Altitudes = np.asarray(Altitudes, dtype=float).ravel()
    MF = 1e6 * np.stack([H2O, np.broadcast_to(CO2, (n_atm, nL)), O3], axis=2)        # [n_atm, nL, 3]
    db = rt._require_db(opts)
    XX = rt.make_spectral_axis(Xmin, Xmax, DV)
    nX = XX.size
    if atm_chunk is None:
        atm_chunk = max(1, min(n_atm, int(3e9 // (8 * nL * nX))))
    X_out = None
    ODs, Las, Lds = [], [], []
    for a0 in range(0, n_atm, atm_chunk):
        a1 = min(n_atm, a0 + atm_chunk)
        na = a1 - a0
        o = dict(opts)
        _, tau, lu, ld = rt.compute_TUD_batch(Xmin, Xmax, T[a0:a1], MF[a0:a1], opts=o, atm_chunk=na, MFs_ID=[1, 2, 3],
                                              DVOUT=DV, Altitudes=Altitudes, returnOD=True, theta_r=0)
        nAlt = Altitudes.size
        # spectral axis first: columns = (atmosphere, altitude) for OD and La, atmosphere for Ld
        Y = torch.cat([tau[:, :, :, 0].permute(1, 0, 2).reshape(nX, na * nAlt),
                       lu[:, :, :, 0].permute(1, 0, 2).reshape(nX, na * nAlt),
                       ld.permute(1, 0).reshape(nX, na)], dim=1).contiguous()
        if X_out is None:
            X_out, R = rt.reduceResolution(XX, Y, DV_out)
        else:
            R = rt.reduceResolution(XX, Y, DV_out, X_out=X_out)
        k = na * nAlt
        ODs.append(R[:, :k].reshape(-1, na, nAlt).permute(0, 2, 1).cpu().numpy())
        Las.append(R[:, k:2 * k].reshape(-1, na, nAlt).permute(0, 2, 1).cpu().numpy())
        Lds.append(R[:, 2 * k:].cpu().numpy())
    return X_out, np.concatenate(ODs, axis=2), np.concatenate(Las, axis=2), np.concatenate(Lds, axis=1)


def generate_LWIR_TUD_MAKO(X_HI, tau_HI, La_HI, Ld_HI, Ts):
    """The MAKO-resolution atmosphere database of Generate_LWIR_TUD_MAKO.py:34-52: the Gaussian MAKO line shape
    (ILS_MAKO.py:2-35) applied to the high-resolution tau / La / Ld [nX, nA] -- one device call on the stacked
    spectra instead of three per-spectrum Python loops --, atmospheres ordered by band-mean transmittance, bands by
    wavenumber.  Returns numpy arrays X [128], tau, La, Ld [128, nA], Ts [nA]."""
    torch = engine._torch()
    tau_HI, La_HI, Ld_HI = (np.asarray(a, dtype=float) for a in (tau_HI, La_HI, Ld_HI))
    nA = tau_HI.shape[1]
    bands = rt._mako_bands_gauss()
    dev = engine.resolve_device(None)
    Y = torch.cat([engine.to_device(a, dev) for a in (tau_HI, La_HI, Ld_HI)], dim=1).contiguous()      # [nX, 3 nA]
    R = engine.ils_apply(1, np.asarray(X_HI, dtype=float), Y, bands[1], bands[2]).cpu().numpy()
    X = bands[0]
    tau, La, Ld = R[:, :nA], R[:, nA:2 * nA], R[:, 2 * nA:]
    ixSrt = np.argsort(np.mean(tau, axis=0))                              # Generate_LWIR_TUD_MAKO.py:40-44
    tau, La, Ld, Ts = tau[:, ixSrt], La[:, ixSrt], Ld[:, ixSrt], np.asarray(Ts)[ixSrt]
    ixSrt = np.argsort(X)                                                 # :46-50
    return X[ixSrt], tau[ixSrt, :], La[ixSrt, :], Ld[ixSrt, :], Ts


def compute_LWIR_apparent_radiance_sweep(X, emis, Ts, tau, La, Ld, dT=None):
    """The apparent-radiance data set of Compute_LWIR_Apparent_Radiance.py:9-26: every input cast to float32,
    surface-temperature offsets dT = arange(-10, 10.5, 0.5), L [nX, nE, nA, nT] float32 from the float32-output kernel
    (half the bytes of this write-bound step) and the surface temperatures T = Ts[:, None] + dT[None, :]."""
    f32 = [np.asarray(a).astype(np.float32) for a in (X, emis, Ts, tau, La, Ld)]
    dT = np.arange(-10, 10.5, 0.5).astype(np.float32) if dT is None else np.asarray(dT).astype(np.float32)
    L = rt.compute_LWIR_apparent_radiance(*f32, dT)
    T = f32[2][:, np.newaxis] + dT[np.newaxis, :]
    return L, T


# Dataset attributes of the reference's HDF5 product (Generate_LWIR_TUD.py:157-221): units / name / info / label
TUD_DATASET_ATTRS = {
    "X": ("cm^{-1}", "Wavenumbers", "Spectral axis for tau, La, Ld", r"$\tilde{\nu} \,\, \left[\si{cm^{-1}} \right]$"),
    "OD": ("none", "Optical Depth", "For nadir-viewing path. tau = np.exp(-OD)", r"$\tau(\tilde{\nu})$"),
    "La": ("µW/(cm^2 sr cm^{-1})", "Atmospheric Path Spectral Radiance", "For nadir-viewing path, earth-to-space",
           r"$L_a(\tilde{\nu})\,\,\left[\si{\micro W/(cm^2.sr.cm^{-1})}\right]$"),
    "Ld": ("µW/(cm^2 sr cm^{-1})", "Atmospheric Downwelling Spectral Radiance", "Hemispherically-averaged, space-to-earth",
           r"$L_d(\tilde{\nu})\,\,\left[\si{\micro W/(cm^2.sr.cm^{-1})}\right]$"),
    "SensorAltitude": ("km", "Sensor Altitude", "Sensor height above surface", r"$z_s \,\, \left[ \si{km} \right]$"),
    "StdAtmos": ("N.A.", "Standard Atmosphere Table", "1976 US Std. Atmos. with CO2 scaled to modern value", r"N.A."),
    "z": ("km", "Altitude", "z=0 at sea level", r"$z \,\, \left[ \si{km} \right]$"),
    "T": ("K", "Temperature profile", "", r"$T(z) \,\, \left[ \si{K} \right]$"),
    "P": ("Pa", "Pressure profile", "", r"$P(z) \,\, \left[ \si{Pa} \right]$"),
    "H2O": ("ppmv", "Water vapor VMR profile", "VMR - volume mixing ratio", r"$\mathrm{H_2O}(z)\,\,\left[\mathrm{ppm}_v\right]$"),
    "O3": ("ppmv", "Ozone VMR profile", "VMR - volume mixing ratio", r"$\mathrm{O_3}(z)\,\,\left[\mathrm{ppm}_v\right]$"),
}


def save_LWIR_TUD_npz(fname, X, OD, La, Ld, Altitudes):
    """The checkpoint file the reference rewrites after every batch (Generate_LWIR_TUD.py:150): same keys."""
    np.savez(fname, X=X, OD=OD, La=La, Ld=Ld, Altitudes=Altitudes)


def load_LWIR_TUD_npz(fname):
    with np.load(fname) as f:
        return {k: f[k] for k in ("X", "OD", "La", "Ld", "Altitudes")}


def save_LWIR_TUD_h5(fname, X, OD, La, Ld, Altitudes, StdAtmos, z, T, P, H2O, O3):
    """The reference's HDF5 product (Generate_LWIR_TUD.py:152-223): one dataset per array with units / name / info /
    label attributes.  Needs h5py, which the reference also needs; there is no hand-written HDF5 encoder here."""
    try:
        import h5py
    except ImportError as e:                                            # pragma: no cover - h5py is not in this image
        raise ImportError("save_LWIR_TUD_h5 needs h5py (as Generate_LWIR_TUD.py does); use save_LWIR_TUD_npz") from e
    data = dict(X=X, OD=OD, La=La, Ld=Ld, SensorAltitude=Altitudes, StdAtmos=StdAtmos, z=z, T=T, P=P, H2O=H2O, O3=O3)
    with h5py.File(fname, "w") as hf:
        for key, arr in data.items():
            d = hf.create_dataset(key, data=arr)
            for attr, val in zip(("units", "name", "info", "label"), TUD_DATASET_ATTRS[key]):
                d.attrs[attr] = val
