"""CPU oracle: numpy restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is the checker, never the product path (radtxfr_b200 never imports it).

Parity status: PINNED against outputs of the reference itself.  The reference (pure Python,
westi024/RadTxfr) has no tests or golden vectors of its own, so the pins
are minted by executing the unmodified reference functions in the build container
(tools/make_golden.py -> tests/golden/*.npz) and tests/test_oracle_golden.py holds this port to
them (<=1e-13 relative; most entries are bit-identical because the same numpy primitives are
evaluated in the same order).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radtxfr_b200", "data",
                     "spectro_tables.npz")
_tables = None


def tables():
    global _tables
    if _tables is None:
        z = np.load(_DATA)
        t = {k: z[k] for k in z.files}
        t["index"] = {(int(m), int(i)): j for j, (m, i) in enumerate(zip(t["iso_M"], t["iso_I"]))}
        _tables = t
    return _tables


# ----------------------------------------------------------------------------- constants
C_BOLTZ = 1.380648813e-16      # erg/K        misc/hapi.py:84-92
C_LIGHT = 2.99792458e10        # cm/s
C_MASS = 1.66053873e-27        # misc/hapi.py:11085
C2_HAPI = 1.4388028496642257   # cm K         misc/hapi.py:10171
T_REF = 296.0                  # misc/hapi.py:10988
P_REF = 1.0                    # misc/hapi.py:10989
RT_C1 = 1.19104295315e-16      # radiative_transfer.py:71
RT_C2 = 1.43877736830e-02      # radiative_transfer.py:72


def volume_concentration(p, T):
    """misc/hapi.py:10163-10164."""
    return (p / 9.869233e-7) / (C_BOLTZ * T)


# ----------------------------------------------------------------------------- TIPS
def _lagrange(aa, A, B, npt):
    """3/4-point Lagrange interpolation, misc/hapi.py:5311-5388 (AtoB).  The reference guards zero
    node spacings; Tdat is strictly increasing so those guards never fire and are omitted."""
    I = None
    for cand in range(2, npt + 1):
        if A[cand - 1] >= aa:
            I = cand
            break
    if I is None:
        raise ValueError("temperature above the TIPS grid")
    if I < 3 or I == npt:
        J = (3 if I < 3 else npt) - 1
        a0, a1, a2 = A[J - 2], A[J - 1], A[J]
        w0 = (aa - a1) * (aa - a2) / ((a0 - a1) * (a0 - a2))
        w1 = (aa - a0) * (aa - a2) / ((a1 - a0) * (a1 - a2))
        w2 = (aa - a0) * (aa - a1) / ((a2 - a0) * (a2 - a1))
        return w0 * B[J - 2] + w1 * B[J - 1] + w2 * B[J]
    J = I - 1
    a0, a1, a2, a3 = A[J - 2], A[J - 1], A[J], A[J + 1]
    w0 = (aa - a1) * (aa - a2) * (aa - a3)
    w0 = w0 / ((a0 - a1) * (a0 - a2) * (a0 - a3))
    w1 = (aa - a0) * (aa - a2) * (aa - a3)
    w1 = w1 / ((a1 - a0) * (a1 - a2) * (a1 - a3))
    w2 = (aa - a0) * (aa - a1) * (aa - a3)
    w2 = w2 / ((a2 - a0) * (a2 - a1) * (a2 - a3))
    w3 = (aa - a0) * (aa - a1) * (aa - a2)
    w3 = w3 / ((a3 - a0) * (a3 - a1) * (a3 - a2))
    return w0 * B[J - 2] + w1 * B[J - 1] + w2 * B[J] + w3 * B[J + 1]


def tips_q(M, I, T):
    """Q(T) as PYTIPS does, misc/hapi.py:9568-9582, 10030 (tables are float32, arithmetic float64)."""
    if T < 70.0 or T > 3000.0:
        raise Exception("TIPS: T must be between 70K and 3000K.")
    t = tables()
    j = t["index"].get((int(M), int(I)))
    if j is None or not t["tips_has"][j]:
        raise Exception("TIPS: no data for M,I = %d,%d." % (M, I))
    return float(_lagrange(np.float64(T), t["tips_T"], t["tips_Q"][j], len(t["tips_T"])))


# ----------------------------------------------------------------------------- w(z)
_WEIDEMAN = {}


def weideman_coeffs(N=24):
    """Coefficients of Weideman's N-term rational expansion, highest power first, plus L.
    Recipe of misc/hapi.py:9812-9824 (cef)."""
    if N not in _WEIDEMAN:
        M = 2 * N
        M2 = 2 * M
        k = np.arange(-M + 1, M)
        L = np.sqrt(N / np.sqrt(2))
        theta = k * np.pi / M
        t = L * np.tan(theta / 2)
        f = np.zeros(len(t) + 1)
        f[1:] = np.exp(-t ** 2) * (L ** 2 + t ** 2)
        a = np.real(np.fft.fft(np.fft.fftshift(f))) / M2
        a = np.flipud(a[1:N + 1])
        _WEIDEMAN[N] = (a.copy(), float(L))
    return _WEIDEMAN[N]


def w_weideman(x, y, N=24):
    """misc/hapi.py:9812-9827: w(z) by the rational series (complex result)."""
    a, L = weideman_coeffs(N)
    z = x + 1.0j * y
    Z = (L + 1.0j * z) / (L - 1.0j * z)
    p = np.zeros(Z.shape, dtype=complex)
    for c in a:                      # numpy.polyval's Horner order
        p = p * Z + c
    return 2 * p / (L - 1.0j * z) ** 2 + (1 / np.sqrt(np.pi)) / (L - 1.0j * z)


def hum1_wei(x, y, n=24):
    """misc/hapi.py:9833-9844: 1-term Humlicek asymptote everywhere, Weideman-24 where |x|+y<15."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    t = y - 1.0j * x
    cerf = 1 / np.sqrt(np.pi) * t / (0.5 + t ** 2)
    mask = np.abs(x) + y < 15.0
    if np.any(mask):
        cerf = np.array(cerf, dtype=complex, copy=True)
        cerf[mask] = w_weideman(x[mask], y[mask], n)
    return cerf.real, cerf.imag


def profile_voigt(sg0, GamD, Gam0, sg):
    """PROFILE_VOIGT, misc/hapi.py:10131-10140 -> pcqsdhc PART1 with every HT extra zero
    (:9900-9915, :10022): V = (1/pi) * (sqrt(pi)*cte*Re w(x+iy)), x=(sg-sg0)*cte, y=Gam0*cte."""
    sg = np.asarray(sg, dtype=float)
    cte = np.sqrt(np.log(2.0)) / GamD
    rpi = np.sqrt(np.pi)
    x = -((sg0 - sg) * cte)
    y = np.full(sg.shape, Gam0 * cte)
    wr, _ = hum1_wei(x, y)
    return (1.0 / np.pi) * ((rpi * cte) * wr)


def cpf3(X, Y):
    """misc/hapi.py:9645-9670: asymptotic series of w(z) for "region 3" of pcqsdhc (15 terms)."""
    X = np.atleast_1d(np.asarray(X, dtype=float))
    Y = np.atleast_1d(np.asarray(Y, dtype=float))
    zone, zi = 1.0 + 0.0j, 0.0 + 1.0j
    tt = np.array([0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5, 9.5, 10.5, 11.5, 12.5, 13.5, 14.5])
    zm1 = zone / (X + zi * Y).astype(complex)
    zm2 = zm1 ** 2
    zsum = zone
    zterm = zone
    for tt_i in tt:
        zterm = zterm * (zm2 * tt_i)
        zsum = zsum + zterm
    zsum = zsum * (zi * zm1 * 0.564189583547756)
    return zsum.real, zsum.imag


def pcqsdhc(sg0, GamD, Gam0, Gam2, Shift0, Shift2, anuVC, eta, sg):
    """pCqSDHC (Hartmann-Tran) profile, misc/hapi.py:9846-10023, all four parts.  Returns (real, imag) of the
    normalised line shape.  PART3 of the reference indexes `sqrt(X)` with the whole X instead of the PART3 subset
    (:9994-9996, usable only when every point is in PART3); the evident intent (the subset) is restated here."""
    sg = np.atleast_1d(np.asarray(sg, dtype=float))
    npts = sg.size
    A = np.zeros(npts, dtype=complex)
    B = np.zeros(npts, dtype=complex)
    cte = np.sqrt(np.log(2.0)) / GamD
    rpi = np.sqrt(np.pi)
    iz = 0.0 + 1.0j
    c0 = complex(Gam0 + 1.0j * Shift0)
    c2 = complex(Gam2 + 1.0j * Shift2)
    c0t = complex((1.0 - eta) * (c0 - 1.5 * c2) + anuVC)
    c2t = complex((1.0 - eta) * c2)

    def W(x, y):
        wr, wi = hum1_wei(x, y)
        return wr + 1.0j * wi

    if abs(c2t) == 0.0:                                            # PART1
        Z1 = (iz * (sg0 - sg) + c0t) * cte
        w1 = W(-Z1.imag, Z1.real)
        A = rpi * cte * w1
        i1 = np.abs(Z1) <= 4.0e3
        if np.any(i1):
            B = rpi * cte * ((1.0 - Z1 ** 2) * w1 + Z1 / rpi)
        if np.any(~i1):
            B = cte * (rpi * w1 + 0.5 / Z1 - 0.75 / (Z1 ** 3))
    else:
        X = (iz * (sg0 - sg) + c0t) / c2t
        Y = complex(1.0 / ((2.0 * cte * c2t)) ** 2)
        csqrtY = (Gam2 - iz * Shift2) / (2.0 * cte * (1.0 - eta) * (Gam2 ** 2 + Shift2 ** 2))
        i2 = np.abs(X) <= 3.0e-8 * abs(Y)
        i3 = (abs(Y) <= 1.0e-15 * np.abs(X)) & ~i2
        i4 = ~(i2 | i3)
        if np.any(i4):                                             # PART4 (the main part)
            Xt = X[i4]
            Z1 = np.sqrt(Xt + Y) - csqrtY
            Z2 = Z1 + 2.0 * csqrtY
            x1, y1, x2, y2 = -Z1.imag, Z1.real, -Z2.imag, Z2.real
            S1 = np.sqrt(x1 ** 2 + y1 ** 2)
            S2 = np.sqrt(x2 ** 2 + y2 ** 2)
            D = np.abs(S1 - S2)
            mx, mn = np.maximum(S1, S2), np.minimum(S1, S2)
            w1 = np.zeros(Xt.size, dtype=complex)
            w2 = np.zeros(Xt.size, dtype=complex)
            i_c3 = (D <= 1.0) & (mx > 8.0) & (mn <= 8.0)
            if np.any(i_c3):
                a, b = cpf3(x1[i_c3], y1[i_c3]); w1[i_c3] = a + 1.0j * b
                a, b = cpf3(x2[i_c3], y2[i_c3]); w2[i_c3] = a + 1.0j * b
            if np.any(~i_c3):
                w1[~i_c3] = W(x1[~i_c3], y1[~i_c3])
                w2[~i_c3] = W(x2[~i_c3], y2[~i_c3])
            A[i4] = rpi * cte * (w1 - w2)
            B[i4] = (-1.0 + rpi / (2.0 * csqrtY) * (1.0 - Z1 ** 2) * w1 - rpi / (2.0 * csqrtY) * (1.0 - Z2 ** 2) * w2) / c2t
        if np.any(i2):                                             # PART2
            Xt = X[i2]
            Z1 = (iz * (sg0 - sg[i2]) + c0t) * cte
            Z2 = np.sqrt(Xt + Y) + csqrtY
            w1 = W(-Z1.imag, Z1.real)
            w2 = W(-Z2.imag, Z2.real)
            A[i2] = rpi * cte * (w1 - w2)
            B[i2] = (-1.0 + rpi / (2.0 * csqrtY) * (1.0 - Z1 ** 2) * w1 - rpi / (2.0 * csqrtY) * (1.0 - Z2 ** 2) * w2) / c2t
        if np.any(i3):                                             # PART3
            Xt = X[i3]
            sq = np.sqrt(Xt + Y)
            w1 = W(-sq.imag, sq.real)
            ia = np.abs(np.sqrt(Xt)) <= 4.0e3
            At = np.zeros(Xt.size, dtype=complex)
            Bt = np.zeros(Xt.size, dtype=complex)
            if np.any(ia):
                sx = np.sqrt(Xt[ia])
                wb = W(-sx.imag, sx.real)
                At[ia] = (2.0 * rpi / c2t) * (1.0 / rpi - sx * wb)
                Bt[ia] = (1.0 / c2t) * (-1.0 + 2.0 * rpi * (1.0 - Xt[ia] - 2.0 * Y) * (1.0 / rpi - sx * wb)
                                        + 2.0 * rpi * sq[ia] * w1[ia])
            if np.any(~ia):
                xn = Xt[~ia]
                At[~ia] = (1.0 / c2t) * (1.0 / xn - 1.5 / (xn ** 2))
                Bt[~ia] = (1.0 / c2t) * (-1.0 + (1.0 - xn - 2.0 * Y) * (1.0 / xn - 1.5 / (xn ** 2))
                                         + 2.0 * rpi * sq[~ia] * w1[~ia])
            A[i3] = At
            B[i3] = Bt
    LS = (1.0 / np.pi) * (A / (1.0 - (anuVC - eta * (c0 - 1.5 * c2)) * A + eta * c2 * B))
    return LS.real, LS.imag


def profile_sdvoigt(sg0, GamD, Gam0, Gam2, Shift0, Shift2, sg):
    """PROFILE_SDVOIGT, misc/hapi.py:10116-10128."""
    return pcqsdhc(sg0, GamD, Gam0, Gam2, Shift0, Shift2, 0.0, 0.0, sg)


# ----------------------------------------------------------------------------- line quantities
def line_strength(sw, T, Q_T, Q_ref, elower, nu):
    """EnvironmentDependency_Intensity, misc/hapi.py:10169-10175."""
    ch = np.exp(-C2_HAPI * elower / T) * (1 - np.exp(-C2_HAPI * nu / T))
    zn = np.exp(-C2_HAPI * elower / T_REF) * (1 - np.exp(-C2_HAPI * nu / T_REF))
    return sw * Q_ref / Q_T * ch / zn


def doppler_hwhm(nu, T, mass):
    """misc/hapi.py:11085-11087 (uses the unshifted line centre)."""
    m = mass * C_MASS * 1000
    return np.sqrt(2 * C_BOLTZ * T * np.log(2) / m / C_LIGHT ** 2) * nu


def _col(cols, name, n, default=None):
    if name in cols:
        return np.asarray(cols[name], dtype=float)
    if default is None:
        return np.zeros(n)
    return default


def line_state(cols, T, p, Diluent=None, GammaL="gamma_air", Components=None, IntensityThreshold=0.0,
               OmegaWing=0.0, OmegaWingHW=50.0):
    """Per-line S(T), GammaD, Gamma0, Shift0, wing and amplitude for one (T, p): the prologue of the
    line loop, misc/hapi.py:11050-11131, evaluated for all rows at once (elementwise, same op order).
    Returns a dict of arrays plus `keep` (rows that pass the (M,I) and intensity filters)."""
    t = tables()
    nu = np.asarray(cols["nu"], dtype=float)
    n = nu.size
    sw = np.asarray(cols["sw"], dtype=float)
    el = np.asarray(cols["elower"], dtype=float)
    mol = np.asarray(cols["molec_id"], dtype=int)
    iso = np.asarray(cols["local_iso_id"], dtype=int)
    if not Diluent:
        g = GammaL.lower()
        if g == "gamma_air":
            Diluent = {"air": 1.0}
        elif g == "gamma_self":
            Diluent = {"self": 1.0}
        else:
            raise Exception("Unknown GammaL value: %s" % GammaL)
    pairs = sorted(set(zip(mol.tolist(), iso.tolist())))
    if Components is None:
        comp = {mi: None for mi in pairs}                       # misc/hapi.py:10237-10251
    else:
        comp = {}
        for c in Components:
            comp[(c[0], c[1])] = c[2] if len(c) >= 3 else None
    keep = np.zeros(n, dtype=bool)
    S = np.zeros(n)
    GamD = np.zeros(n)
    ampl = np.zeros(n)
    for (M, I) in pairs:
        if (M, I) not in comp:
            continue
        j = t["index"].get((M, I))
        if j is None:
            raise Exception("cannot find component M,I = %d,%d." % (M, I))
        sel = (mol == M) & (iso == I)
        qT = tips_q(M, I, T)
        qR = tips_q(M, I, T_REF)
        S[sel] = line_strength(sw[sel], T, qT, qR, el[sel], nu[sel])
        GamD[sel] = doppler_hwhm(nu[sel], T, t["iso_mass"][j])
        nat = t["iso_abundance"][j]
        ab = nat if comp[(M, I)] is None else comp[(M, I)]
        ampl[sel] = 1.0 / nat * ab                               # factor applied by the caller
        keep |= sel
    keep &= ~(S < IntensityThreshold)                            # misc/hapi.py:11082
    Gam0 = np.zeros(n)
    Shift0 = np.zeros(n)
    Gam2 = np.zeros(n)
    n_air = _col(cols, "n_air", n)
    for species, abun in Diluent.items():                        # misc/hapi.py:11090-11128
        s = species.lower()
        gam = _col(cols, "gamma_" + s, n)
        if ("n_" + s) in cols:
            npow = np.asarray(cols["n_" + s], dtype=float)
            if s == "self":
                npow = np.where(npow == 0.0, n_air, npow)
        else:
            npow = n_air
        Gam0 = Gam0 + abun * (gam * p / P_REF * (T_REF / T) ** npow)
        dlt = _col(cols, "delta_" + s, n)
        dlp = _col(cols, "deltap_" + s, n)
        Shift0 = Shift0 + abun * ((dlt + dlp * (T - T_REF)) * p / P_REF)
        Gam2 = Gam2 + abun * (_col(cols, "SD_" + s, n) * p / P_REF) * gam      # SDVoigt driver only, misc/hapi.py:10883-10890
    wing = np.maximum(np.maximum(OmegaWing, OmegaWingHW * Gam0), OmegaWingHW * GamD)   # :11131
    return dict(nu=nu, S=S, GamD=GamD, Gam0=Gam0, Shift0=Shift0, Gam2=Gam2, wing=wing, ampl=ampl, keep=keep)


def absorption_coefficient_voigt(cols, T=296.0, p=1.0, OmegaGrid=None, OmegaWing=0.0, OmegaWingHW=50.0,
                                 IntensityThreshold=0.0, GammaL="gamma_air", HITRAN_units=True, Diluent=None,
                                 Components=None, return_evals=False):
    """absorptionCoefficient_Voigt, misc/hapi.py:10906-11141, for one dict-of-columns table.
    Accumulates in table row order like the reference (:11050)."""
    Omegas = np.sort(np.asarray(OmegaGrid, dtype=float))         # :10980
    Xsect = np.zeros(Omegas.size)
    factor = 1.0 if HITRAN_units else volume_concentration(p, T)
    st = line_state(cols, T, p, Diluent, GammaL, Components, IntensityThreshold, OmegaWing, OmegaWingHW)
    nu, S, GamD, Gam0, Sh, W, ampl, keep = (st[k] for k in ("nu", "S", "GamD", "Gam0", "Shift0", "wing", "ampl",
                                                               "keep"))
    lo_all = np.searchsorted(Omegas, nu - W, side="right")       # bisect == bisect_right, :11133-11134
    hi_all = np.searchsorted(Omegas, nu + W, side="right")
    evals = 0
    evals_w = 0
    for r in np.nonzero(keep)[0]:
        lo, hi = int(lo_all[r]), int(hi_all[r])
        if hi <= lo:
            continue
        sg = Omegas[lo:hi]
        V = profile_voigt(nu[r] + Sh[r], GamD[r], Gam0[r], sg)
        Xsect[lo:hi] += factor * ampl[r] * S[r] * V
        if return_evals:
            evals += hi - lo
            cte = np.sqrt(np.log(2.0)) / GamD[r]
            evals_w += int(np.count_nonzero(np.abs((sg - (nu[r] + Sh[r])) * cte) + Gam0[r] * cte < 15.0))
    if return_evals:
        return Omegas, Xsect, evals, evals_w
    return Omegas, Xsect


def absorption_coefficient_sdvoigt(cols, T=296.0, p=1.0, OmegaGrid=None, OmegaWing=0.0, OmegaWingHW=50.0,
                                   IntensityThreshold=0.0, GammaL="gamma_air", HITRAN_units=True, Diluent=None,
                                   Components=None):
    """absorptionCoefficient_SDVoigt, misc/hapi.py:10657-10905: Gamma2 = sum_d x_d * SD_d * p/pref * gamma_d(ref)
    (:10883-10890), Shift2 = 0, profile PROFILE_SDVOIGT(nu, GamD, Gam0, Gam2, Shift0, 0, ...) with the line shift passed
    into the profile (:10899)."""
    Omegas = np.sort(np.asarray(OmegaGrid, dtype=float))
    Xsect = np.zeros(Omegas.size)
    factor = 1.0 if HITRAN_units else volume_concentration(p, T)
    st = line_state(cols, T, p, Diluent, GammaL, Components, IntensityThreshold, OmegaWing, OmegaWingHW)
    nu, S, GamD, Gam0, Gam2, Sh, W, ampl, keep = (st[k] for k in ("nu", "S", "GamD", "Gam0", "Gam2", "Shift0", "wing",
                                                                  "ampl", "keep"))
    lo_all = np.searchsorted(Omegas, nu - W, side="right")
    hi_all = np.searchsorted(Omegas, nu + W, side="right")
    for r in np.nonzero(keep)[0]:
        lo, hi = int(lo_all[r]), int(hi_all[r])
        if hi <= lo:
            continue
        V = profile_sdvoigt(nu[r], GamD[r], Gam0[r], Gam2[r], Sh[r], 0.0, Omegas[lo:hi])[0]
        Xsect[lo:hi] += factor * ampl[r] * S[r] * V
    return Omegas, Xsect


def absorption_coefficient_lorentz(cols, T=296.0, p=1.0, OmegaGrid=None, OmegaWing=0.0, OmegaWingHW=50.0,
                                   IntensityThreshold=0.0, GammaL="gamma_air", HITRAN_units=True, Diluent=None,
                                   Components=None):
    """absorptionCoefficient_Lorentz, misc/hapi.py:11144-11381: wing = max(OmegaWing, HW*Gamma0) (:11375), profile
    PROFILE_LORENTZ = Gam0/(pi*(Gam0**2+(sg-sg0)**2)) (:10142-10150) centred on nu+Shift0."""
    Omegas = np.sort(np.asarray(OmegaGrid, dtype=float))
    Xsect = np.zeros(Omegas.size)
    factor = 1.0 if HITRAN_units else volume_concentration(p, T)
    st = line_state(cols, T, p, Diluent, GammaL, Components, IntensityThreshold, OmegaWing, OmegaWingHW)
    nu, S, Gam0, Sh, ampl, keep = (st[k] for k in ("nu", "S", "Gam0", "Shift0", "ampl", "keep"))
    W = np.maximum(OmegaWing, OmegaWingHW * Gam0)
    lo_all = np.searchsorted(Omegas, nu - W, side="right")
    hi_all = np.searchsorted(Omegas, nu + W, side="right")
    for r in np.nonzero(keep)[0]:
        lo, hi = int(lo_all[r]), int(hi_all[r])
        if hi <= lo:
            continue
        sg = Omegas[lo:hi]
        V = Gam0[r] / (np.pi * (Gam0[r] ** 2 + (sg - (nu[r] + Sh[r])) ** 2))
        Xsect[lo:hi] += factor * ampl[r] * S[r] * V
    return Omegas, Xsect


def absorption_coefficient_doppler(cols, T=296.0, p=1.0, OmegaGrid=None, OmegaWing=0.0, OmegaWingHW=50.0,
                                   IntensityThreshold=0.0, HITRAN_units=True, LineShift=True, Components=None):
    """absorptionCoefficient_Doppler, misc/hapi.py:11384-11550: GammaD from the driver's own SI constants
    (:11533-11537), wing = max(OmegaWing, HW*GammaD) (:11540), Shift0 = delta_air*p/pref when LineShift (:11543),
    PROFILE_DOPPLER (:10152-10160) with hapi's truncated cLn2 / cSqrtLn2divSqrtPi (:89-90)."""
    Omegas = np.sort(np.asarray(OmegaGrid, dtype=float))
    Xsect = np.zeros(Omegas.size)
    factor = 1.0 if HITRAN_units else volume_concentration(p, T)
    st = line_state(cols, T, p, None, "gamma_air", Components, IntensityThreshold, OmegaWing, OmegaWingHW)
    nu, S, ampl, keep = (st[k] for k in ("nu", "S", "ampl", "keep"))
    t = tables()
    mass = np.array([t["iso_mass"][t["index"][(int(m), int(i))]] for m, i in zip(cols["molec_id"], cols["local_iso_id"])])
    GammaD = (1.1774100225 / 2.99792458e8) * np.sqrt(1.3806503e-23 / 1.66053873e-27) * np.sqrt(T) * nu / np.sqrt(mass)
    Sh = np.asarray(cols["delta_air"], dtype=float) * p / 1.0 if LineShift else np.zeros_like(nu)
    W = np.maximum(OmegaWing, OmegaWingHW * GammaD)
    lo_all = np.searchsorted(Omegas, nu - W, side="right")
    hi_all = np.searchsorted(Omegas, nu + W, side="right")
    for r in np.nonzero(keep)[0]:
        lo, hi = int(lo_all[r]), int(hi_all[r])
        if hi <= lo:
            continue
        sg = Omegas[lo:hi]
        V = 0.469718639319144059835 * np.exp(-0.6931471805599 * ((sg - (nu[r] + Sh[r])) / GammaD[r]) ** 2) / GammaD[r]
        Xsect[lo:hi] += factor * ampl[r] * S[r] * V
    return Omegas, Xsect


def count_evals(cols, T, p, Omegas, **kw):
    """Exact number of (line, grid-point) pairs inside hapi's window for one (T, p)."""
    st = line_state(cols, T, p, **kw)
    W = st["wing"]
    lo = np.searchsorted(Omegas, st["nu"] - W, side="right")
    hi = np.searchsorted(Omegas, st["nu"] + W, side="right")
    return int(np.sum((hi - lo)[st["keep"]]))


# ----------------------------------------------------------------------------- TUD
def make_spectral_axis(Xmin, Xmax, DVOUT):
    """radiative_transfer.py:251-271 with the int() fix needed on numpy>=1.18."""
    return np.linspace(Xmin, Xmax, int(np.ceil((Xmax - Xmin) / DVOUT)))


def planckian(X_in, T_in, wavelength=False):
    """radiative_transfer.py:792-848."""
    X = np.asarray(np.copy(X_in), dtype=float).flatten()[:, np.newaxis]
    T = np.asarray(np.copy(T_in), dtype=float)
    dimsT = T.shape
    T = T.flatten()[np.newaxis, :]
    if wavelength or np.mean(X) < 50:
        X = X * 1e-6
        L = RT_C1 / (X ** 5 * (np.exp(RT_C2 / (X * T)) - 1))
        L = L * 1e-4
    else:
        X = X * 100
        L = RT_C1 * X ** 3 / (np.exp(RT_C2 * X / T) - 1)
        L = L * 1e4
    return np.reshape(L, (X.size, *dimsT))


def brightness_temperature(X_in, L_in, wavelength=False, bad_value=np.nan, spectral_dim=0):
    """brightnessTemperature, radiative_transfer.py:851-933."""
    X = np.asarray(np.copy(X_in), dtype=float).flatten()[:, None]
    L = np.asarray(np.copy(L_in), dtype=float)
    if spectral_dim != 0:
        L = np.swapaxes(L, 0, spectral_dim)
    dims = L.shape if L.ndim > 1 else (L.shape[0], 1)
    L = L.reshape(dims[0], -1)
    with np.errstate(all="ignore"):
        if wavelength or np.mean(X) < 50:
            X = X * 1e-6
            L = L * 1e4
            T = RT_C2 / (X * np.log(1 + RT_C1 / (X ** 5 * L)))
        else:
            X = X * 100
            L = L * 1e-4
            T = RT_C2 * X / np.log(RT_C1 * X ** 3 / L + 1)
    T[~np.isfinite(L) | (L <= 0)] = bad_value
    T = T.reshape(dims)
    return np.swapaxes(T, 0, spectral_dim) if spectral_dim != 0 else T


def bt2l(X_in, T_in, wavelength=False, bad_value=np.nan, spectral_dim=0):
    """BT2L, radiative_transfer.py:936-1014."""
    X = np.asarray(np.copy(X_in), dtype=float).flatten()[:, None]
    T = np.asarray(np.copy(T_in), dtype=float)
    if spectral_dim != 0:
        T = np.swapaxes(T, 0, spectral_dim)
    dims = T.shape if T.ndim > 1 else (T.shape[0], 1)
    T = T.reshape(dims[0], -1)
    with np.errstate(all="ignore"):
        if wavelength or np.mean(X) < 50:
            X = X * 1e-6
            L = RT_C1 / (X ** 5 * (np.exp(RT_C2 / (X * T)) - 1)) * 1e-4
        else:
            X = X * 100
            L = RT_C1 * X ** 3 / (np.exp(RT_C2 * X / T) - 1) * 1e4
    L[~np.isfinite(L) | (T <= 0)] = bad_value
    L = L.reshape(dims)
    return np.swapaxes(L, 0, spectral_dim) if spectral_dim != 0 else L


def layer_od(tables_by_mol, X, T, P_pa, PL_km, MF_VAL_ppmv, MF_ID, **voigt_kw):
    """Optical depth of one homogeneous layer: the wiring of SURVEY.md App. B.1 that stands in for the
    reference's LBLRTM call (radiative_transfer.py:395-456 contract: returns OD on the output axis)."""
    p = P_pa / 101325.0
    od = np.zeros(X.size)
    for mid, val in zip(np.asarray(MF_ID).ravel(), np.asarray(MF_VAL_ppmv).ravel()):
        mid = int(mid)
        if mid not in tables_by_mol:
            continue
        _, xs = absorption_coefficient_voigt(tables_by_mol[mid], T=float(T), p=float(p), OmegaGrid=X,
                                             HITRAN_units=True, **voigt_kw)
        od += (val * 1e-6) * volume_concentration(p, float(T)) * (float(PL_km) * 1e5) * xs
    return od


def tud_from_od(X, OD, Z, T, N_angle=30, Altitudes=(500,), theta_r=0, returnOD=False):
    """Body of compute_TUD after the per-layer OD loop, radiative_transfer.py:340-392, quirks included:
    nL is overwritten inside the altitude loop (:353), so the downwelling recurrence runs over the layer
    count of the LAST altitude (:370)."""
    Z = np.asarray(Z, dtype=float)
    T = np.asarray(T, dtype=float)
    nL = T.size
    Z_s = np.array([Altitudes]).ravel()
    mu_s = np.array([1.0 / np.cos(theta_r)]).ravel()
    B = planckian(X, T)
    Lu = np.zeros((X.size, Z_s.size, mu_s.size))
    tau = Lu.copy()
    Ld = np.zeros((X.size, N_angle))
    for ii, zs in enumerate(Z_s):
        for jj, mu in enumerate(mu_s):
            ix = Z <= zs
            s = np.sum(OD[:, ix] * mu, axis=1)
            tau[:, ii, jj] = s if returnOD else np.exp(-1.0 * s)
            nL = int(np.sum(ix))
            for kk in range(nL):
                t = np.exp(-OD[:, kk] * mu)
                Lu[:, ii, jj] = t * Lu[:, ii, jj] + (1 - t) * B[:, kk]
    if len(Z_s) == 1 and len(mu_s) == 1:
        tau, Lu = tau[:, 0, 0], Lu[:, 0, 0]
    elif len(Z_s) == 1:
        tau, Lu = tau[:, 0, :], Lu[:, 0, :]
    elif len(mu_s) == 1:
        tau, Lu = tau[:, :, 0], Lu[:, :, 0]
    angles = np.linspace(0, np.pi / 2.0, N_angle, endpoint=False)
    for ii, th in enumerate(angles):
        for jj in range(nL)[::-1]:
            t = np.exp(-OD[:, jj] / np.cos(th))
            Ld[:, ii] = t * Ld[:, ii] + (1 - t) * B[:, jj]
    cw = np.cos(angles) * np.sin(angles)
    Ld = (np.sum(Ld * cw, axis=1) / np.sum(cw)).flatten()
    return tau, Lu, Ld


def compute_tud(tables_by_mol, Xmin, Xmax, Zs, Ts, Ps, PLs, MFs_VAL, MFs_ID, DVOUT=0.0005, N_angle=30,
                Altitudes=(500,), theta_r=0, returnOD=False, return_od_matrix=False, **voigt_kw):
    """compute_TUD, radiative_transfer.py:274-392, with compute_OD -> layer_od."""
    X = make_spectral_axis(Xmin, Xmax, DVOUT)
    Ts = np.asarray(Ts, dtype=float)
    OD = np.zeros((X.size, Ts.size))
    for ii in range(Ts.size):
        OD[:, ii] = layer_od(tables_by_mol, X, Ts[ii], np.asarray(Ps)[ii], np.asarray(PLs)[ii],
                             np.asarray(MFs_VAL)[ii, :], MFs_ID, **voigt_kw)
    tau, Lu, Ld = tud_from_od(X, OD, Zs, Ts, N_angle, Altitudes, theta_r, returnOD)
    if return_od_matrix:
        return X, tau, Lu, Ld, OD
    return X, tau, Lu, Ld


# ----------------------------------------------------------------------------- sensor model
def apparent_radiance(X, emis, Ts, tau, La, Ld, dT=None, return_Ls=False):
    """compute_LWIR_apparent_radiance, radiative_transfer.py:1017-1069."""
    X = np.asarray(X, dtype=float)
    emis, Ts, tau, La, Ld = (np.asarray(a, dtype=float) for a in (emis, Ts, tau, La, Ld))
    if dT is not None:
        T_ = Ts.flatten()[:, None] + np.asarray(dT, dtype=float).flatten()[None, :]
        B_ = planckian(X, T_)[:, None, :, :]
        tau_, La_, Ld_ = tau[:, None, :, None], La[:, None, :, None], Ld[:, None, :, None]
        em_ = emis[:, :, None, None]
    else:
        B_ = planckian(X, Ts.flatten())[:, None, :]
        tau_, La_, Ld_ = tau[:, None, :], La[:, None, :], Ld[:, None, :]
        em_ = emis[:, :, None]
    Ls = em_ * B_ + (1 - em_) * Ld_
    L = tau_ * Ls + La_
    return (L, Ls) if return_Ls else L


def mako_axis(X, resFactor=None, clip=True):
    """Output band centres [cm^-1]: radiative_transfer.py:1226-1233 (clip=True) / ILS_MAKO.py:20."""
    lam = tables()["mako_um"]
    if resFactor is not None:
        x0 = np.linspace(0, 1, len(lam))
        x1 = np.linspace(0, 1, int(len(lam) * resFactor))
        lam = np.interp(x1, x0, lam)
    Xo = np.sort(10000.0 / lam)
    if clip:
        Xo = Xo[(Xo > X.min()) & (Xo < X.max())]
    return Xo


def ils_mako_triangle(X, Y, resFactor=None, returnX=True, fwhm_sf=1.0, shift=0.0, scale=1.0):
    """ILS_MAKO (triangle), radiative_transfer.py:1072-1263; band-at-a-time to avoid the dense
    (nX, nB) weight matrix -- per band the same sums over the same terms in the same x order."""
    X = np.asarray(X, dtype=float)
    Y = np.asarray(Y, dtype=float)
    Xo = mako_axis(X, resFactor, clip=True)
    sig = fwhm_sf * np.abs(np.gradient(Xo)) * 1.6
    Y2 = Y if Y.ndim == 2 else Y[:, None]
    out = np.zeros((Xo.size, Y2.shape[1]))
    for b in range(Xo.size):
        w = 1.0 - np.abs(X - (scale * Xo[b] + shift)) / sig[b]
        w[w < 0] = 0
        out[b] = np.sum(w[:, None] * Y2, axis=0) / np.sum(w)
    if Y.ndim == 1:
        out = out[:, 0]
    return (Xo, out) if returnX else out


def ils_mako_gauss(X, Y):
    """ILS_MAKO (Gaussian, no clipping, infinite support), ILS_MAKO.py:2-35."""
    X = np.asarray(X, dtype=float)
    Y = np.asarray(Y, dtype=float)
    Xo = mako_axis(X, None, clip=False)
    sig = np.abs(np.gradient(Xo))
    Y2 = Y if Y.ndim == 2 else Y[:, None]
    out = np.zeros((Xo.size, Y2.shape[1]))
    for b in range(Xo.size):
        g = np.exp(-0.5 * ((X - Xo[b]) / sig[b]) ** 2) / (sig[b] * np.sqrt(2.0 * np.pi))
        out[b] = np.sum(g[:, None] * Y2, axis=0) / np.sum(g)
    if Y.ndim == 1:
        out = out[:, 0]
    return Xo, out


# ----------------------------------------------------------------------------- resolution reduction
def smooth(x, window_len=11, window="hanning"):
    """smooth, radiative_transfer.py:1266-1323."""
    x = np.asarray(x, dtype=float)
    if x.ndim != 1 or x.size < window_len or window_len < 3:
        return x
    s = np.r_[x[window_len - 1:0:-1], x, x[-2:-window_len - 1:-1]]
    w = np.ones(window_len, "d") if window == "flat" else getattr(np, window)(window_len)
    y = np.convolve(w / w.sum(), s, mode="valid")
    return y[int(np.ceil(window_len / 2 - 1)):-int(np.floor(window_len / 2))]


def reduce_resolution(X, Y, dX, N=4, window="hanning", X_out=None):
    """reduceResolution, radiative_transfer.py:1326-1350 (np.int -> int)."""
    import scipy.interpolate
    X = np.asarray(X, dtype=float)
    Y = np.asarray(Y, dtype=float)
    smFactor = int(np.round(dX / np.mean(np.diff(X))))
    sm1 = lambda y: smooth(y, window_len=smFactor, window=window)
    sm = lambda y: 0.5 * (sm1(y) + sm1(y[::-1])[::-1])
    interp = lambda x, y, x0: scipy.interpolate.interp1d(x, y, kind="cubic", bounds_error=False,
                                                         fill_value="extrapolate")(x0)
    X_ = sm(X)
    ret_x = X_out is None
    if ret_x:
        nPts = int(np.ceil(N * (X_[-smFactor - 1] - X_[smFactor]) / dX)) + 1
        X_out = np.linspace(X_[smFactor], X_[-smFactor - 1], nPts)
    if Y.ndim > 1:
        Y_out = np.zeros((X_out.size, Y.shape[-1]))
        for ii in range(Y.shape[-1]):
            Y_out[:, ii] = interp(X_, sm(Y[:, ii]), X_out)
    else:
        Y_out = interp(X_, sm(Y), X_out)
    return (X_out, Y_out) if ret_x else Y_out
