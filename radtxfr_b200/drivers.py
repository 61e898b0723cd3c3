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
