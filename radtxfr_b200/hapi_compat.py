"""Drop-in for the part of the reference's vendored hapi (misc/hapi.py) that lies on the hot path.

Same names, argument meaning and errors as the reference for:
    LOCAL_TABLE_CACHE, absorptionCoefficient_Voigt, PYTIPS, volumeConcentration, molecularMass, abundance
The line sum itself runs in the sm_100a kernels of libradtxfr_b200.so (no CPU fallback).

Deliberate differences (documented in DESIGN.md):
  * EnvDependences callbacks, partitionFunction other than PYTIPS, and diluents other than 'air'/'self'
    raise NotImplementedError (they are arbitrary per-line Python in the reference);
  * the OmegaRange/OmegaStep route, which raises TypeError in the reference on numpy >= 1.18
    (misc/hapi.py:133-139), builds the grid the reference intends: linspace(lo, lo+step*(n-1), n),
    n = floor((hi-lo)/step)+1;
  * inputs that are CUDA tensors give CUDA tensors back; numpy in -> numpy out.
"""
from warnings import warn

import numpy as np

from . import engine

# hapi's in-memory table store: {name: {'header': {...}, 'data': {column: list-or-array}}}
# (misc/hapi.py:438-463).  Users fill it exactly as they would fill hapi's.
LOCAL_TABLE_CACHE = {}

DefaultIntensityThreshold = 0.0   # misc/hapi.py:10215
DefaultOmegaWingHW = 50.0         # misc/hapi.py:10218

cBolts = 1.380648813e-16          # misc/hapi.py:84-92
cc = 2.99792458e10

# Arithmetic of the Voigt sum: "fp64" (default; 1e-9 parity: one-Newton-step reciprocal ~1e-12 per term, far-field
# blocks < 1e-13), "fp64_exact" (every pair at its own grid point, fully refined reciprocal) or "fp32" (optional
# single-precision asymptote branch, 1e-5 parity); engine.PRECISION_FLAGS.
VARIABLES = {"precision": "fp64"}

_linedb_cache = {}


def volumeConcentration(p, T):
    """misc/hapi.py:10163-10164."""
    return (p / 9.869233e-7) / (cBolts * T)


def abundance(M, I):
    """misc/hapi.py:5088-5104."""
    t = engine.spectro_tables()
    return float(t["iso_abundance"][t["index"][(M, I)]])


def molecularMass(M, I):
    """misc/hapi.py:5109-5124."""
    t = engine.spectro_tables()
    return float(t["iso_mass"][t["index"][(M, I)]])


def PYTIPS(M, I, T):
    """Total internal partition sum Q(T), TIPS-2011 (misc/hapi.py:10030 -> :9568-9582), on the device."""
    if T < 70.0 or T > 3000.0:
        raise Exception("TIPS: T must be between 70K and 3000K.")
    t = engine.spectro_tables()
    j = t["index"].get((int(M), int(I)))
    if j is None or not t["tips_has"][j]:
        raise Exception("TIPS: no data for M,I = %d,%d." % (M, I))
    db = _tips_db((int(M), int(I)))
    return float(db.tips(np.array([float(T)]))[0, 0].item())


_tips_dbs = {}


def _tips_db(mi):
    if mi not in _tips_dbs:
        cols = {"molec_id": np.array([mi[0]]), "local_iso_id": np.array([mi[1]]), "nu": np.array([1000.0]),
                "sw": np.array([0.0]), "elower": np.array([0.0]), "gamma_air": np.array([0.0]),
                "n_air": np.array([0.0]), "delta_air": np.array([0.0])}
        _tips_dbs[mi] = engine.LineDB(cols)
    return _tips_dbs[mi]


def storeTable(name, cols):
    """Convenience: put a dict of columns into LOCAL_TABLE_CACHE the way hapi lays tables out."""
    LOCAL_TABLE_CACHE[name] = {"header": {"table_name": name, "number_of_rows": len(cols["nu"])}, "data": cols}
    invalidate(name)


_grid_cache = {}         # device -> (host copy of the last large axis, its device copy)


def invalidate(name=None):
    """Drop the device copies of table `name` (all tables when None): the next call uploads again.  Needed only after
    an in-place edit of column values that engine.table_fingerprint cannot see."""
    for key in [k for k in _linedb_cache if name is None or name in k[0]]:
        del _linedb_cache[key]


def line_database(SourceTables, device=None):
    """Device-resident LineDB for a tuple of table names, re-uploaded only when the table data changed."""
    names = tuple(SourceTables)
    tabs, fps = [], []
    for n in names:
        if n not in LOCAL_TABLE_CACHE:
            raise Exception("%s: no such table. Check tableList() for more info." % n)   # misc/hapi.py:10242
        data = LOCAL_TABLE_CACHE[n]["data"]
        nrow = LOCAL_TABLE_CACHE[n].get("header", {}).get("number_of_rows", len(data["nu"]))
        fps.append((nrow, engine.table_fingerprint(data)))
        if nrow != len(data["nu"]):
            data = {k: v[:nrow] for k, v in data.items()}      # the reference loops range(number_of_rows)
        if len(data["molec_id"]) != len(data["local_iso_id"]):
            raise Exception("Lengths if mol_ids and iso_ids differ!")                     # misc/hapi.py:10246
        tabs.append(data)
    dev = engine.resolve_device(device)
    fp = tuple(fps)
    key = (names, dev.index)
    hit = _linedb_cache.get(key)
    if hit is None or hit[0] != fp:
        _linedb_cache[key] = (fp, engine.LineDB(tabs, dev))
    return _linedb_cache[key][1]


def _hapi_grid(OmegaRange, OmegaStep):
    lo, hi = float(OmegaRange[0]), float(OmegaRange[1])
    n = int(np.floor((hi - lo) / OmegaStep)) + 1            # arange_, misc/hapi.py:133-139
    return np.linspace(lo, lo + OmegaStep * (n - 1), n)


def absorptionCoefficient_Voigt(Components=None, SourceTables=None, partitionFunction=PYTIPS, Environment=None,
                                OmegaRange=None, OmegaStep=None, OmegaWing=None,
                                IntensityThreshold=DefaultIntensityThreshold, OmegaWingHW=DefaultOmegaWingHW,
                                GammaL='gamma_air', HITRAN_units=True, LineShift=True, File=None, Format=None,
                                OmegaGrid=None, WavenumberRange=None, WavenumberStep=None, WavenumberWing=None,
                                WavenumberWingHW=None, WavenumberGrid=None, Diluent={}, EnvDependences=None,
                                _profile_flags=0):
    """Voigt absorption cross-section / coefficient on a wavenumber grid; signature, defaults and return value
    of the reference's absorptionCoefficient_Voigt (misc/hapi.py:10906-11141)."""
    torch = engine._torch()
    # deprecated aliases, misc/hapi.py:10957-10961 (the reference's `if WavenumberGrid:` truthiness test raises
    # for arrays; `is not None` is the evident intent)
    if WavenumberRange: OmegaRange = WavenumberRange
    if WavenumberStep: OmegaStep = WavenumberStep
    if WavenumberWing: OmegaWing = WavenumberWing
    if WavenumberWingHW: OmegaWingHW = WavenumberWingHW
    if WavenumberGrid is not None: OmegaGrid = WavenumberGrid
    if EnvDependences:
        raise NotImplementedError("EnvDependences callbacks are arbitrary per-line Python and are not supported")
    if partitionFunction is not PYTIPS:
        raise NotImplementedError("only the TIPS-2011 partition function (PYTIPS) is supported")
    if File:
        raise NotImplementedError("File output: write (Omegas, Xsect) on the host after the call")

    if type(SourceTables) not in (list, tuple):                     # listOfTuples, misc/hapi.py:10224-10227
        SourceTables = [SourceTables]
    if SourceTables[0] is None:
        SourceTables = ['__BUFFER__']
    if Environment is None:
        Environment = {'T': 296., 'p': 1.}
    db = line_database(SourceTables)
    if OmegaStep is None: OmegaStep_ = 0.01
    else: OmegaStep_ = OmegaStep
    if OmegaWing is None: OmegaWing = 0.0
    if OmegaStep_ > 0.1: warn('Big wavenumber step: possible accuracy decline')   # misc/hapi.py:10974

    as_tensor = isinstance(OmegaGrid, torch.Tensor)
    grid_dev = None
    if OmegaGrid is not None:
        if as_tensor:
            Omegas = torch.sort(OmegaGrid.to(device=db.device, dtype=torch.float64).reshape(-1))[0]
        else:
            Omegas = np.array(OmegaGrid, dtype=np.float64).ravel()                # a copy, like the reference's sort
            if Omegas.size > 1 and not np.all(Omegas[1:] >= Omegas[:-1]):
                Omegas.sort()                                                     # misc/hapi.py:10980
    else:
        if OmegaRange is None:
            OmegaRange = (db.nu_min, db.nu_max)                                   # misc/hapi.py:10252-10263
        Omegas = _hapi_grid(OmegaRange, OmegaStep_)
    if not as_tensor and Omegas.size >= (1 << 16):
        # loops over (T, p) on one large axis (misc/RT_gen_AbsXS_files.py:87-93): the device copy of the axis is kept
        # and reused while the host values are unchanged (compared in full: 20 MB/ms, a tenth of the upload)
        hit = _grid_cache.get(str(db.device))
        if hit is not None and hit[0].size == Omegas.size and np.array_equal(hit[0], Omegas):
            grid_dev = hit[1]
        else:
            grid_dev = engine.to_device(Omegas, db.device)
            _grid_cache[str(db.device)] = (Omegas.copy(), grid_dev)

    T = float(Environment['T'])
    p = float(Environment['p'])
    if type(Components) not in (list, tuple):
        Components = [Components]
    comps = None if list(Components) == [None] else Components
    factor = 1.0 if HITRAN_units else volumeConcentration(p, T)                    # misc/hapi.py:11011-11014
    w_iso = db.iso_factors(comps, factor)

    GammaL = GammaL.lower()
    if int(_profile_flags) & 32:
        Diluent = {'air': 1.}        # the Doppler driver has no pressure broadening; GammaL='dummy' is ignored
    if not Diluent:
        if GammaL == 'gamma_air': Diluent = {'air': 1.}
        elif GammaL == 'gamma_self': Diluent = {'self': 1.}
        else: raise Exception('Unknown GammaL value: %s' % GammaL)                 # misc/hapi.py:11032
    x_air = x_self = 0.0
    x_extra = {}
    for key, val in Diluent.items():          # any species: hapi reads gamma_<sp>, n_<sp>, delta_<sp>, ... and treats a
        k = key.lower()                        # missing gamma_<sp> column as zero broadening (misc/hapi.py:11090-11128)
        if k == 'air': x_air += float(val)
        elif k == 'self': x_self += float(val)
        else: x_extra[k] = x_extra.get(k, 0.0) + float(val)
    if x_self != 0.0 and not db.has_self:
        x_self = 0.0      # no gamma_self column: the reference's try/except gives Gamma0DB = 0 (:11097-11100)

    if db.n_lines:
        db.check_temperatures([T])
        db.check_tips(w_iso != 0.0)
    xs = db.voigt_sum(np.array([T]), np.array([p]), w_iso[None, :], Omegas if grid_dev is None else grid_dev, x_air=x_air, x_self=x_self,
                      wing_abs=float(OmegaWing), wing_hw=float(OmegaWingHW),
                      s_threshold=float(IntensityThreshold), x_extra=x_extra,
                      flags=engine.PRECISION_FLAGS[VARIABLES["precision"]] | int(_profile_flags))[0]
    if as_tensor:
        return Omegas, xs
    if xs.numel() >= (1 << 18):
        from . import radiative_transfer as _rt          # pinned staging buffer + threaded copy-out (lazy: it imports this module)
        return Omegas, _rt._to_host(xs)[0]
    return Omegas, xs.cpu().numpy()


def absorptionCoefficient_SDVoigt(*args, **kwargs):
    """Speed-dependent Voigt driver of hapi (misc/hapi.py:10657-10905).  Gamma2 = sum_d x_d * SD_d * p/pref * gamma_d
    over the diluents (:10883-10890), Shift2 = 0; lines with Gamma2 != 0 take the quadratic speed-dependent profile
    (pcqsdhc PART2-4, :9923-10020, with cpf3 :9645-9670 where the reference selects it), lines with Gamma2 == 0 --
    every line of a table without SD_<diluent> columns, the situation of the reference's only hapi call site,
    misc/RT_gen_AbsXS_files.py:90 -- its PART1 branch, i.e. the Voigt profile centred on nu + Shift0 (the reference
    forms that centre as (nu - omega) + Shift0 instead of (nu + Shift0) - omega: a last-bit difference)."""
    kwargs["_profile_flags"] = int(kwargs.get("_profile_flags", 0)) | engine.VOIGT_SDVOIGT
    return absorptionCoefficient_Voigt(*args, **kwargs)


def absorptionCoefficient_Lorentz(*args, **kwargs):
    """Lorentz-profile driver of hapi (misc/hapi.py:11144-11381): same arguments as absorptionCoefficient_Voigt; the
    profile is Gamma0/(pi*(Gamma0^2 + (omega-nu0-shift)^2)) (PROFILE_LORENTZ, :10142-10150) and the wing is
    max(OmegaWing, OmegaWingHW*Gamma0) (:11375)."""
    kwargs["_profile_flags"] = 16          # RTX_VOIGT_LORENTZ
    return absorptionCoefficient_Voigt(*args, **kwargs)


def absorptionCoefficient_Doppler(Components=None, SourceTables=None, partitionFunction=PYTIPS, Environment=None,
                                  OmegaRange=None, OmegaStep=None, OmegaWing=None,
                                  IntensityThreshold=DefaultIntensityThreshold, OmegaWingHW=DefaultOmegaWingHW,
                                  ParameterBindings=None, EnvironmentDependencyBindings=None, GammaL='dummy',
                                  HITRAN_units=True, LineShift=True, File=None, Format=None, OmegaGrid=None,
                                  WavenumberRange=None, WavenumberStep=None, WavenumberWing=None,
                                  WavenumberWingHW=None, WavenumberGrid=None):
    """Doppler (Gauss) profile driver of hapi (misc/hapi.py:11384-11550): profile
    cSqrtLn2divSqrtPi*exp(-cLn2*((omega-nu0-shift)/GammaD)**2)/GammaD (:10152-10160) with the driver's own SI constants
    for GammaD (:11533-11537), wing max(OmegaWing, OmegaWingHW*GammaD), shift delta_air*p only when LineShift."""
    return absorptionCoefficient_Voigt(Components=Components, SourceTables=SourceTables,
                                       partitionFunction=partitionFunction, Environment=Environment,
                                       OmegaRange=OmegaRange, OmegaStep=OmegaStep, OmegaWing=OmegaWing,
                                       IntensityThreshold=IntensityThreshold, OmegaWingHW=OmegaWingHW,
                                       HITRAN_units=HITRAN_units, File=File, Format=Format, OmegaGrid=OmegaGrid,
                                       WavenumberRange=WavenumberRange, WavenumberStep=WavenumberStep,
                                       WavenumberWing=WavenumberWing, WavenumberWingHW=WavenumberWingHW,
                                       WavenumberGrid=WavenumberGrid,
                                       _profile_flags=32 | (0 if LineShift else 64))   # RTX_VOIGT_DOPPLER [| NOSHIFT]


absorptionCoefficient_Gauss = absorptionCoefficient_Doppler       # misc/hapi.py:11557


hh = 6.626196e-27                 # erg*s, misc/hapi.py:86


def _xp(a):
    """numpy for host arrays, torch for tensors (these helpers are element-wise post-processing, not the hot path)."""
    torch = engine._torch()
    return torch if isinstance(a, torch.Tensor) else np


def transmittanceSpectrum(Omegas, AbsorptionCoefficient, Environment={'l': 100.}, File=None, Format='%e %e',
                          Wavenumber=None):
    """exp(-k*l) on the grid (misc/hapi.py:11582-11610)."""
    if Wavenumber is not None: Omegas = Wavenumber
    if File: raise NotImplementedError("File output: write (Omegas, Xsect) on the host after the call")
    return Omegas, _xp(AbsorptionCoefficient).exp(-AbsorptionCoefficient * Environment['l'])


def absorptionSpectrum(Omegas, AbsorptionCoefficient, Environment={'l': 100.}, File=None, Format='%e %e',
                       Wavenumber=None):
    """1 - exp(-k*l) on the grid (misc/hapi.py:11613-11641)."""
    if Wavenumber is not None: Omegas = Wavenumber
    if File: raise NotImplementedError("File output: write (Omegas, Xsect) on the host after the call")
    return Omegas, 1 - _xp(AbsorptionCoefficient).exp(-AbsorptionCoefficient * Environment['l'])


def radianceSpectrum(Omegas, AbsorptionCoefficient, Environment={'l': 100., 'T': 296.}, File=None, Format='%e %e',
                     Wavenumber=None):
    """(1 - exp(-k*l)) * Planck(Omegas, T) in W/sr/cm^2/cm^-1 (misc/hapi.py:11644-11681)."""
    if Wavenumber is not None: Omegas = Wavenumber
    if File: raise NotImplementedError("File output: write (Omegas, Xsect) on the host after the call")
    xp = _xp(AbsorptionCoefficient)
    l, T = Environment['l'], Environment['T']
    Alw = 1 - xp.exp(-AbsorptionCoefficient * l)
    Om = Omegas if xp is np or not isinstance(Omegas, np.ndarray) else xp.as_tensor(Omegas, device=AbsorptionCoefficient.device)
    LBBTw = 2 * hh * cc ** 2 * Om ** 3 / (xp.exp(hh * cc * Om / (cBolts * T)) - 1) * 1.0E-7
    return Omegas, Alw * LBBTw
