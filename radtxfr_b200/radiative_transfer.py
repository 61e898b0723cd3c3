"""Drop-in for the hot-path functions of the reference's radiative_transfer.py.

Same names, arguments, return values and quirks as the reference for
    make_spectral_axis, compute_TUD, compute_OD, planckian, compute_LWIR_apparent_radiance, ILS_MAKO
(radiative_transfer.py:251-271, 274-392, 395-456, 792-848, 1017-1069, 1072-1263).  The arithmetic runs in the
sm_100a kernels of libradtxfr_b200.so; numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.

What replaces LBLRTM: the reference obtains each layer's optical depth from an external LBLRTM binary reading a
TAPE3 line file (radiative_transfer.py:459-501), neither of which ships with it.  Here compute_OD is the hapi
Voigt line sum over a device-resident line database:
    OD = sum_m  (MF_VAL_m*1e-6) * volumeConcentration(p, T) * (PL*1e5) * sigma_m(omega; T, p)
Register the database once with set_line_database({molec_id: table}) (tables are hapi-style dicts of columns).
No continuum model (MT_CKD lives inside LBLRTM).
"""
import numpy as np

from . import engine, hapi_compat

c1 = 1.19104295315e-16  # [J m^2 / s]  radiative_transfer.py:71
c2 = 1.43877736830e-02  # [m K]        radiative_transfer.py:72

StdAtmos = engine.spectro_tables()["std_atmos"].copy()      # radiative_transfer.py:75-148

# Default options dictionary: the keys compute_TUD / compute_OD read (radiative_transfer.py:152-183).  As in the
# reference this dict is module-level and MUTABLE: every call does opts.update(kwargs), so settings are sticky
# across calls.  That behaviour is kept on purpose -- this is a drop-in.
options = {
    "T": 296.0,             # [K]
    "P": 101325.0,          # [Pa]
    "PL": 1.0,              # [km]
    "MF_ID": np.array([]),
    "MF_VAL": np.array([]),
    "DVOUT": 0.0005,        # [cm^-1]
    "Zs": StdAtmos[:, 1],   # [km]
    "Ts": StdAtmos[:, 5],   # [K]
    "Ps": StdAtmos[:, 4],   # [Pa]
    "PLs": StdAtmos[:, 3],  # [km]
    "MFs_VAL": StdAtmos[:, 6:14] * 1e6,  # [ppmv]
    "MFs_ID": np.array([1, 2, 3, 4, 5, 6, 7, 22]),
    "theta_r": 0,
    "N_angle": 30,
    "Altitudes": np.asarray([500]),
    "save": False,
    "returnOD": False,
    "continuum_od": None,          # optional additive layer optical depth (array or callable), see _add_continuum
    # --- additions of this implementation -----------------------------------------------------
    "LINE_DB": None,        # engine.LineDB, set through set_line_database()
    "voigt": {},            # extra Voigt settings: OmegaWing, OmegaWingHW, IntensityThreshold, Diluent, and
                            # precision: "fp64" (default, 1e-9), "fp64_exact" (no far-field blocks, fully refined
                            # reciprocal), "fp32" (1e-5) -- see engine.PRECISION_FLAGS
}


def set_line_database(tables_by_mol, device=None, opts=options):
    """Upload the line tables ({molec_id: dict of hapi columns}) once; stands in for the TAPE3 file of the
    reference (radiative_transfer.py:151,491).  Returns the engine.LineDB."""
    if isinstance(tables_by_mol, engine.LineDB):
        db = tables_by_mol
    else:
        tabs = [tables_by_mol[m] for m in sorted(tables_by_mol)]
        db = engine.LineDB(tabs, device)
    opts["LINE_DB"] = db
    return db


_axis_cache = {}
_workspace = {}


def _device_axis(Xmin, Xmax, DVOUT, device):
    """Device copy of the spectral axis, kept across calls (LRU of 4).  Returns (host axis or None, device axis):
    on a cache hit the host axis is NOT rebuilt here -- the caller builds it while the GPU is busy."""
    nX = int(np.ceil((Xmax - Xmin) / DVOUT))
    key = (float(Xmin), float(Xmax), nX, str(device))
    hit = _axis_cache.get(key)
    if hit is not None:
        return None, hit
    X_ = make_spectral_axis(Xmin, Xmax, DVOUT)
    if len(_axis_cache) >= 4:
        _axis_cache.pop(next(iter(_axis_cache)))
    _axis_cache[key] = engine.to_device(X_, device)
    return X_, _axis_cache[key]


def _od_workspace(device, n_states, nx):
    """Grow-only optical-depth buffer per device: avoids a 0.4 GB allocation per compute_TUD call."""
    torch = engine._torch()
    need = n_states * nx
    buf = _workspace.get(str(device))
    if buf is None or buf.numel() < need:
        buf = None
        _workspace.pop(str(device), None)
        buf = _workspace[str(device)] = torch.empty(need, dtype=torch.float64, device=device)
    return buf[:need].view(n_states, nx)


_copy_pool = None


def _host_outputs(shapes):
    """Fresh float64 numpy arrays whose pages are already touched (call this while the GPU is busy: first-touch
    page faults of 20 MB of results would otherwise add ~1.5 ms after the kernels)."""
    outs = [np.empty(sh, dtype=np.float64) for sh in shapes]
    for o in outs:
        o.fill(0.0)
    return outs


def _to_host(*tensors, outs=None, early=None):
    """Device tensors -> fresh numpy arrays through one persistent pinned staging buffer: async D2H, then the copies
    out of the staging buffer run on a small thread pool (numpy releases the GIL).  early = (event, n): the first n
    tensors are complete at `event` (recorded on the current stream while later kernels are still queued); their
    transfers go on a side stream so that they overlap those kernels.  The staging buffer and the side stream are
    module-level state: like the reference's module-level `options`, this is not meant for concurrent calls from
    several host threads."""
    global _copy_pool
    torch = engine._torch()
    dev = tensors[0].device
    side = None
    if early is not None:
        side = _workspace.get(("side_stream", str(dev)))
        if side is None:
            side = _workspace[("side_stream", str(dev))] = torch.cuda.Stream(device=dev)
        side.wait_event(early[0])
    total = sum(t.numel() for t in tensors)
    stage = _workspace.get("pinned")
    if stage is None or stage.numel() < total:
        stage = _workspace["pinned"] = torch.empty(max(total, 1), dtype=torch.float64, pin_memory=True)
    off = 0
    views, events = [], []
    for i, t in enumerate(tensors):
        v = stage[off:off + t.numel()]
        ev = torch.cuda.Event()
        if side is not None and i < early[1]:
            with torch.cuda.stream(side):
                v.copy_(t.reshape(-1), non_blocking=True)
                ev.record(side)
            t.record_stream(side)
        else:
            v.copy_(t.reshape(-1), non_blocking=True)
            ev.record()                               # this tensor's bytes have landed once the event completes
        events.append(ev)
        views.append(v.numpy().reshape(tuple(t.shape)))
        off += t.numel()
    if outs is None:
        outs = [np.empty(v.shape, dtype=np.float64) for v in views]
    if total < (1 << 18):
        for ev in events:                             # some transfers ran on the side stream: wait for each one
            ev.synchronize()
        for o, v in zip(outs, views):
            np.copyto(o, v)
    else:
        # copy each array out of the staging buffer as soon as its own transfer is done, while the next one is
        # still on the bus
        if _copy_pool is None:
            from concurrent.futures import ThreadPoolExecutor
            _copy_pool = ThreadPoolExecutor(max_workers=6)
        futs = []
        for o, v, ev in zip(outs, views, events):
            ev.synchronize()
            of, vf = o.reshape(-1), v.reshape(-1)
            h = of.size // 2
            futs.append(_copy_pool.submit(np.copyto, of[:h], vf[:h]))
            futs.append(_copy_pool.submit(np.copyto, of[h:], vf[h:]))
        for f in futs:
            f.result()
    return outs


def make_spectral_axis(Xmin, Xmax, DVOUT):
    """radiative_transfer.py:251-271 (with the int() the reference needs on numpy >= 1.18)."""
    nX = int(np.ceil((Xmax - Xmin) / DVOUT))
    return np.linspace(Xmin, Xmax, nX)


def _voigt_settings(opts):
    v = dict(opts.get("voigt") or {})
    dil = v.get("Diluent") or {"air": 1.0}
    x_air = float(sum(val for k, val in dil.items() if k.lower() == "air"))
    x_self = float(sum(val for k, val in dil.items() if k.lower() == "self"))
    flags = engine.PRECISION_FLAGS[v.get("precision", "fp64")]
    flags |= int(v.get("tile_k", 0)) << 8          # 0 = automatic tile size
    flags |= 4 if v.get("ws", False) else 0
    return dict(x_air=x_air, x_self=x_self, wing_abs=float(v.get("OmegaWing", 0.0)),
                wing_hw=float(v.get("OmegaWingHW", 50.0)), s_threshold=float(v.get("IntensityThreshold", 0.0)),
                flags=flags)


def layer_weights(db, T, P_pa, PL_km, MF_VAL, MF_ID):
    """Per-(layer, isotopologue) column amounts: (MF_VAL*1e-6)*volumeConcentration(p,T)*(PL*1e5) for the
    molecule of each isotopologue, times hapi's abundance factor (misc/hapi.py:11136).
    T, P_pa, PL_km: [...]; MF_VAL: [..., nMol]; returns p [atm] and weights [..., n_iso]."""
    T = np.asarray(T, dtype=float)
    p = np.asarray(P_pa, dtype=float) / 101325.0
    PL = np.asarray(PL_km, dtype=float)
    MF_VAL = np.asarray(MF_VAL, dtype=float)
    ids = [int(m) for m in np.asarray(MF_ID).ravel()]
    base = db.iso_factors(None, 1.0)
    w = np.zeros(T.shape + (db.n_iso,))
    nconc = hapi_compat.volumeConcentration(p, T)
    for j, mol in enumerate(db.iso_molecule.tolist()):
        if mol in ids:
            k = ids.index(mol)
            w[..., j] = (MF_VAL[..., k] * 1e-6) * nconc * (PL * 1e5) * base[j]
    return p, w


def _add_continuum(od, opts, X, T, P, PL, MF, ID):
    """Optional continuum hook (the reference gets its continuum, MT_CKD, from inside LBLRTM and only passes
    `continuum_factors` through, radiative_transfer.py:162-163, 530, 590-672; there is no continuum model here):
    opts["continuum_od"] is an array broadcastable to the layer optical depths [nL, nX], or a callable
    (X, Ts, Ps, PLs, MFs_VAL, MFs_ID) -> such an array (numpy or a torch tensor on the device); it is added to the
    line-by-line optical depth on the device before the layer integration."""
    c = opts.get("continuum_od")
    if c is None:
        return od
    if callable(c):
        c = c(X, T, P, PL, MF, ID)
    c = engine.to_device(c, od.device)
    if c.dim() == 1 and c.numel() == od.shape[0] and od.shape[0] != od.shape[-1]:
        c = c[:, None]                                   # one value per layer
    od += c                                              # broadcasts [nL, nX], [nX], [nL, 1] or a scalar
    return od


def _require_db(opts):
    db = opts.get("LINE_DB")
    if db is None:
        raise RuntimeError("no line database: call radtxfr_b200.radiative_transfer.set_line_database(...) first "
                           "(it replaces the TAPE3 file the reference hands to LBLRTM)")
    return db


def compute_OD(Xmin_in, Xmax_in, opts=options, **kwargs):
    """Optical depth of a single homogeneous layer: contract of radiative_transfer.py:395-456
    (returns X_out = linspace(Xmin, Xmax, ceil(range/DVOUT)) and OD on it)."""
    opts.update(kwargs)
    db = _require_db(opts)
    DVOUT = opts.get("DVOUT", 0.025)
    X = make_spectral_axis(Xmin_in, Xmax_in, DVOUT)
    T = np.array([float(opts["T"])])
    p, w = layer_weights(db, T, np.array([float(opts["P"])]), np.array([float(opts["PL"])]),
                         np.asarray(opts["MF_VAL"], dtype=float).reshape(1, -1), opts["MF_ID"])
    db.check_temperatures(T)
    od = db.voigt_sum(T, p, w, X, **_voigt_settings(opts))[0]
    return X, od.cpu().numpy()


def compute_TUD(Xmin, Xmax, opts=options, **kwargs):
    """Monochromatic transmittance, upwelling and downwelling radiance: radiative_transfer.py:274-392.
    Returns X, tau, Lu, Ld with the reference's shapes (squeeze rules :357-365)."""
    opts.update(kwargs)
    db = _require_db(opts)
    Z = np.asarray(opts["Zs"], dtype=float)
    T = np.asarray(opts["Ts"], dtype=float)
    P = np.asarray(opts["Ps"], dtype=float)
    PL = np.asarray(opts["PLs"], dtype=float)
    MF = np.asarray(opts["MFs_VAL"], dtype=float)
    ID = opts["MFs_ID"]
    nA = opts["N_angle"]
    Z_s = np.array([opts["Altitudes"]]).ravel()
    mu_s = np.array([1.0 / np.cos(opts["theta_r"])]).ravel()
    returnOD = opts["returnOD"]

    X_, grid = _device_axis(Xmin, Xmax, opts["DVOUT"], db.device)
    nX = grid.numel()
    db.check_temperatures(T)
    p, w = layer_weights(db, T, P, PL, MF, ID)
    # every host->device copy happens before the first launch (pageable copies synchronise the stream), so the
    # host is free while the kernels run
    Td, pd, wd = (engine.to_device(a, db.device) for a in (T, p, w))
    od = db.voigt_sum(Td, pd, wd, grid, out=_od_workspace(db.device, T.size, nX),
                      **_voigt_settings(opts))                                  # [nL, nX], layer-major
    if opts.get("continuum_od") is not None:
        if X_ is None and callable(opts["continuum_od"]):
            X_ = make_spectral_axis(Xmin, Xmax, opts["DVOUT"])     # a callable always receives the numpy axis
        od = _add_continuum(od, opts, X_ if X_ is not None else grid, T, P, PL, MF, ID)
    tau, lu, ld, up_ready = engine.tud(od, Td[None, :], grid, Z, Z_s, mu_s, nA, returnOD, split=True)
    # the kernels are in flight: build the host axis and pre-touch the result arrays meanwhile
    if X_ is None:
        X_ = make_spectral_axis(Xmin, Xmax, opts["DVOUT"])
    outs = _host_outputs([(nX, Z_s.size, mu_s.size), (nX, Z_s.size, mu_s.size), (nX,)])
    if opts["save"]:
        B = engine.planck(grid, T).cpu().numpy()
        np.savez("ComputeTUD.npz", OD=od.cpu().numpy().T, B=B, tau=tau[0].cpu().numpy(), Lu=lu[0].cpu().numpy(),
                 X=X_, angles=np.linspace(0, np.pi / 2.0, nA, endpoint=False), Z_s=Z_s, mu_s=mu_s)
    tau_, Lu_, Ld_ = _to_host(tau[0], lu[0], ld[0], outs=outs, early=(up_ready, 2))
    if (len(Z_s) == 1) and (len(mu_s) == 1):
        tau_, Lu_ = tau_[:, 0, 0], Lu_[:, 0, 0]
    if (len(Z_s) == 1) and not (len(mu_s) == 1):
        tau_, Lu_ = tau_[:, 0, :], Lu_[:, 0, :]
    if not (len(Z_s) == 1) and (len(mu_s) == 1):
        tau_, Lu_ = tau_[:, :, 0], Lu_[:, :, 0]
    return X_, tau_, Lu_, Ld_


def compute_TUD_batch(Xmin, Xmax, Ts, MFs_VAL, opts=options, atm_chunk=None, ils=None, **kwargs):
    """Batched form of the production loop `for jj: compute_TUD(..., MFs_VAL=MFs_VAL[jj], Ts=Ts[jj])`
    (Generate_LWIR_TUD.py:117-122): Ts [n_atm, nL], MFs_VAL [n_atm, nL, nMol]; everything else as compute_TUD.
    Returns X and CUDA tensors tau, Lu [n_atm, nX, nAlt, nMu], Ld [n_atm, nX] -- or, with ils='triangle'|'gauss',
    the MAKO-band products (X_out, tau, Lu, Ld) with nX replaced by the band count."""
    torch = engine._torch()
    opts.update(kwargs)
    db = _require_db(opts)
    Z = np.asarray(opts["Zs"], dtype=float)
    P = np.asarray(opts["Ps"], dtype=float)
    PL = np.asarray(opts["PLs"], dtype=float)
    Ts = np.asarray(Ts, dtype=float)
    MF = np.asarray(MFs_VAL, dtype=float)
    Ts = Ts.reshape(-1, Z.size)
    MF = MF.reshape(Ts.shape[0], Z.size, MF.shape[-1])
    n_atm, nL = Ts.shape
    Z_s = np.array([opts["Altitudes"]]).ravel()
    mu_s = np.array([1.0 / np.cos(opts["theta_r"])]).ravel()
    X_, grid = _device_axis(Xmin, Xmax, opts["DVOUT"], db.device)      # on a cache hit the host axis is built after the launches
    nX = grid.numel()
    db.check_temperatures(Ts)
    if atm_chunk is None:
        atm_chunk = max(1, min(max(n_atm, 1), int(4e9 // (8 * nL * nX))))   # <= 4 GB of optical depth at a time
    atm_chunk = min(int(atm_chunk), 32768)                                  # rtx_tud: n_atm <= 65535 per launch
    vs = _voigt_settings(opts)
    # batches never split the lines of a tile over CTAs (engine.VOIGT_NOSPLIT): the kernel variant of a state then depends on
    # the grid, the table and the state alone, so the result does not depend on how the batch is chunked or sharded
    vs["flags"] |= engine.VOIGT_NOSPLIT
    outs = []
    band = None
    if ils is not None:
        # linspace(Xmin, Xmax, n) starts at Xmin and ends at Xmax exactly: the clip interval of the triangle bands
        band = _mako_bands(np.array([float(Xmin), float(Xmax)]), None, 1.0, 0.0, 1.0) if ils == "triangle" else _mako_bands_gauss()
    if n_atm == 0:                                   # a rank that owns no atmosphere (distributed.compute_TUD_batch_sharded)
        nout = nX if band is None else band[0].size
        z = torch.zeros((0, nout, Z_s.size, mu_s.size), dtype=torch.float64, device=db.device)
        if X_ is None and band is None:
            X_ = make_spectral_axis(Xmin, Xmax, opts["DVOUT"])
        return (band[0] if band is not None else X_), z, z.clone(), torch.zeros((0, nout), dtype=torch.float64, device=db.device)
    Pb = np.ascontiguousarray(np.broadcast_to(P, (n_atm, nL)))
    PLb = np.ascontiguousarray(np.broadcast_to(PL, (n_atm, nL)))
    for a0 in range(0, n_atm, atm_chunk):
        a1 = min(n_atm, a0 + atm_chunk)
        na = a1 - a0
        # column amounts, Voigt sums and layer integration back to back inside the library (rtx_tud_from_lines); the
        # optical depth of the chunk lives in the handle's scratch
        tau, lu, ld = db.tud_from_lines(Ts[a0:a1], Pb[a0:a1], PLb[a0:a1], MF[a0:a1], opts["MFs_ID"], grid, Z, Z_s, mu_s,
                                        opts["N_angle"], opts["returnOD"], **vs)
        if band is not None:
            mode = 0 if ils == "triangle" else 1
            na_, nx_, nal, nmu = tau.shape
            Y = torch.cat([tau.permute(1, 0, 2, 3).reshape(nX, -1), lu.permute(1, 0, 2, 3).reshape(nX, -1),
                           ld.permute(1, 0).reshape(nX, -1)], dim=1).contiguous()
            R = engine.ils_apply(mode, grid, Y, band[1], band[2])
            nb = R.shape[0]
            k = na * nal * nmu
            tau = R[:, :k].reshape(nb, na, nal, nmu).permute(1, 0, 2, 3).contiguous()
            lu = R[:, k:2 * k].reshape(nb, na, nal, nmu).permute(1, 0, 2, 3).contiguous()
            ld = R[:, 2 * k:].reshape(nb, na).permute(1, 0).contiguous()
        outs.append((tau, lu, ld))
    if len(outs) == 1:
        tau, lu, ld = outs[0]
    else:
        tau = torch.cat([o[0] for o in outs], dim=0)
        lu = torch.cat([o[1] for o in outs], dim=0)
        ld = torch.cat([o[2] for o in outs], dim=0)
    if X_ is None and band is None:
        X_ = make_spectral_axis(Xmin, Xmax, opts["DVOUT"])        # the kernels are in flight
    return (band[0] if band is not None else X_), tau, lu, ld


def planckian(X_in, T_in, wavelength=False):
    """Planck spectral radiance, radiative_transfer.py:792-848: output shape (X.size, *T.shape), in
    uW/(cm^2 sr cm^-1), or uW/(cm^2 sr um) for wavelength input (also chosen when mean(X) < 50, :834)."""
    torch = engine._torch()
    as_tensor = isinstance(X_in, torch.Tensor) or isinstance(T_in, torch.Tensor)
    if as_tensor:
        X = X_in if isinstance(X_in, torch.Tensor) else torch.as_tensor(np.asarray(X_in, dtype=float))
        T = T_in if isinstance(T_in, torch.Tensor) else torch.as_tensor(np.asarray(T_in, dtype=float))
        X = X.reshape(-1).double()
        dimsT = tuple(T.shape)
        mean_small = bool(X.mean().item() < 50) if X.numel() else False
    else:
        X = np.asarray(np.copy(X_in), dtype=float).flatten()
        T = np.asarray(np.copy(T_in), dtype=float)
        dimsT = T.shape
        mean_small = bool(np.mean(X) < 50) if X.size else False
    use_wl = bool(wavelength or mean_small)
    if use_wl and not wavelength:
        print("Assumes X given in µm; returning L in µF")
    L = engine.planck(X, T.reshape(-1) if as_tensor else T.flatten(), use_wl)
    n = L.shape[0]
    if as_tensor:
        return L.reshape((n,) + dimsT)
    return np.reshape(L.cpu().numpy(), (n, *dimsT))


def _bt(X_in, V_in, wavelength, bad_value, spectral_dim, inverse):
    torch = engine._torch()
    as_tensor = isinstance(V_in, torch.Tensor)
    if as_tensor:
        V = V_in.double()
        X = (X_in if isinstance(X_in, torch.Tensor) else torch.as_tensor(np.asarray(X_in, dtype=float))).reshape(-1).double()
        if spectral_dim != 0:
            V = V.swapaxes(0, spectral_dim)
        mean_small = bool(X.mean().item() < 50)
    else:
        X = np.asarray(np.copy(X_in), dtype=float).flatten()
        V = np.asarray(np.copy(V_in), dtype=float)
        if spectral_dim != 0:
            V = np.swapaxes(V, 0, spectral_dim)
        mean_small = bool(np.mean(X) < 50)
    dims = tuple(V.shape)
    use_wl = bool(wavelength or mean_small)
    if use_wl and not wavelength:
        print("Assumes X given in µm and L given in µF")
    V2 = V.reshape(dims[0], -1) if as_tensor else np.ascontiguousarray(V.reshape(dims[0], -1))
    out = engine.brightness(X, V2.contiguous() if as_tensor else V2, use_wl, inverse, bad_value)
    out = out if as_tensor else out.cpu().numpy()
    if V.ndim == 1:                       # the reference returns (nX, 1) for vector input (both functions)
        out = out.reshape(dims[0], 1)
    else:
        out = out.reshape(dims)
    if spectral_dim != 0:
        out = out.swapaxes(0, spectral_dim) if as_tensor else np.swapaxes(out, 0, spectral_dim)
    return out


def brightnessTemperature(X_in, L_in, wavelength=False, bad_value=np.nan, spectral_dim=0):
    """Brightness temperature of a spectral radiance, radiative_transfer.py:851-933 (spectral axis first, or at
    `spectral_dim`; unphysical inputs give `bad_value`)."""
    return _bt(X_in, L_in, wavelength, bad_value, spectral_dim, False)


def BT2L(X_in, T_in, wavelength=False, bad_value=np.nan, spectral_dim=0):
    """Spectral radiance of a brightness temperature, radiative_transfer.py:936-1014."""
    return _bt(X_in, T_in, wavelength, bad_value, spectral_dim, True)


def compute_LWIR_apparent_radiance(X, emis, Ts, tau, La, Ld, dT=None, return_Ls=False):
    """L = tau*(emis*B(T) + (1-emis)*Ld) + La for every (emissivity, atmosphere[, dT]) combination:
    radiative_transfer.py:1017-1069.  Shapes: X (nX,), emis (nX,nE), Ts (nA,), tau/La/Ld (nX,nA), dT (nT,) ->
    L (nX,nE,nA) or (nX,nE,nA,nT).  When every array argument is float32 -- the reference's production script casts
    them so (Compute_LWIR_Apparent_Radiance.py:9-25) and numpy then returns float32 -- the result is float32 too:
    evaluated in FP64 from the float32 values and rounded once, i.e. within float32 rounding of the reference."""
    torch = engine._torch()
    as_tensor = any(isinstance(a, torch.Tensor) for a in (emis, tau, La, Ld))
    arrs = [a for a in (X, emis, Ts, tau, La, Ld, dT) if a is not None]

    def is_f32(a):
        return (a.dtype == torch.float32) if isinstance(a, torch.Tensor) else (getattr(a, "dtype", None) == np.float32)
    f32 = all(is_f32(a) for a in arrs)
    out = engine.apparent_radiance(X, emis, Ts, tau, La, Ld, dT=dT, return_Ls=return_Ls,
                                   out_dtype=torch.float32 if f32 else None)
    if as_tensor:
        return out
    if return_Ls:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def _window_taps(window_len, window):
    if window == "flat":
        w = np.ones(window_len, "d")
    else:
        w = getattr(np, window)(window_len)                  # eval("np." + window + "(window_len)"), :1316
    return w / w.sum()


def smooth(x, window_len=11, window="hanning"):
    """Window smoothing with reflected ends, radiative_transfer.py:1266-1323 (same guards and return length)."""
    torch = engine._torch()
    as_tensor = isinstance(x, torch.Tensor)
    nd = x.dim() if as_tensor else np.ndim(x)
    size = x.numel() if as_tensor else np.size(x)
    if nd != 1:
        print("smooth only accepts 1 dimension arrays.")
        return x
    if size < window_len:
        print("Input vector needs to be bigger than window size.")
        return x
    if window_len < 3:
        return x
    if window not in ["flat", "hanning", "hamming", "bartlett", "blackman"]:
        print("Window is on of 'flat', 'hanning', 'hamming', 'bartlett', 'blackman'")
        return x
    xin = x.reshape(-1, 1) if as_tensor else np.asarray(x, dtype=float)[:, None]
    out = engine.smooth(xin, _window_taps(window_len, window), False)[:, 0]
    return out if as_tensor else out.cpu().numpy()


def reduceResolution(X, Y, dX, N=4, window="hanning", X_out=None):
    """Smooth to resolution dX and resample with a cubic spline on N points per dX: radiative_transfer.py:1326-1350
    (the step that follows compute_TUD in Generate_LWIR_TUD.py:82-85, 124-126).  Returns (X_out, Y_out), or Y_out
    alone when X_out is given.  The reference needs np.int restored to run on numpy >= 1.24; int() is used here."""
    torch = engine._torch()
    as_tensor = isinstance(Y, torch.Tensor)
    Xh = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X, dtype=float)
    dX_in = np.mean(np.diff(Xh))
    smFactor = int(np.round(dX / dX_in))
    one_d = (Y.dim() if as_tensor else np.ndim(Y)) == 1
    Y2 = (Y.reshape(-1, 1) if one_d else Y) if as_tensor else (np.asarray(Y, dtype=float)[:, None] if one_d
                                                                  else np.asarray(Y, dtype=float))
    if smFactor < 3 or Xh.size < smFactor:                  # smooth() returns its input unchanged (:1301-1306)
        X_s = engine.to_device(Xh, None)[:, None]
        Y_s = engine.to_device(Y2, X_s.device)
    else:
        taps = _window_taps(smFactor, window)
        X_s = engine.smooth(Xh[:, None], taps, True)        # smFcn(X)
        Y_s = engine.smooth(Y2, taps, True)                 # smFcn(Y[:, ii])
    returnX_out = False
    if X_out is None:
        a = float(X_s[smFactor, 0].item())
        b = float(X_s[-smFactor - 1, 0].item())
        nPts = int(np.ceil(N * (b - a) / dX)) + 1
        X_out = np.linspace(a, b, nPts)
        returnX_out = True
    Xo = X_out.detach().cpu().numpy() if isinstance(X_out, torch.Tensor) else np.asarray(X_out, dtype=float)
    out = engine.spline_eval(X_s[:, 0].contiguous(), Y_s, Xo)
    if one_d:
        out = out[:, 0]
    if not as_tensor:
        out = out.cpu().numpy()
    if returnX_out:
        return X_out, out
    return out


def _mako_um():
    return engine.spectro_tables()["mako_um"]          # radiative_transfer.py:1092-1223


def _mako_bands(X, resFactor, fwhm_sf, shift, scale):
    """Band centres / triangle half-bases exactly as radiative_transfer.py:1226-1241 derives them (rtx_mako_bands)."""
    X = np.asarray(X)
    return engine.mako_bands(0, float(np.min(X)), float(np.max(X)), resFactor, fwhm_sf, shift, scale)


def _mako_bands_gauss():
    """ILS_MAKO.py:20-22: all 128 bands, sigma = |gradient| (rtx_mako_bands, mode 1)."""
    return engine.mako_bands(1, 0.0, 0.0)


def _ils(mode, X, Y, bands, returnX):
    torch = engine._torch()
    X_out, centres, widths = bands
    as_tensor = isinstance(Y, torch.Tensor)
    one_d = (Y.dim() if as_tensor else np.ndim(Y)) == 1
    if as_tensor:
        Y2 = Y.reshape(-1, 1) if one_d else Y
        Xd = X
    else:
        Y2 = np.asarray(Y, dtype=float)
        Y2 = Y2[:, None] if one_d else Y2
        Xd = np.asarray(X, dtype=float)
    out = engine.ils_apply(mode, Xd, Y2, centres, widths)
    if one_d:
        out = out[:, 0]
    if not as_tensor:
        out = out.cpu().numpy()
    if returnX:
        return X_out, out
    return out


def ILS_MAKO(X, Y, resFactor=None, returnX=True, fwhm_sf=1.0, shift=0.0, scale=1.0):
    """Apply the MAKO triangle instrument line shape: radiative_transfer.py:1072-1263."""
    torch = engine._torch()
    Xh = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X, dtype=float)
    bands = _mako_bands(Xh, resFactor, fwhm_sf, shift, scale)
    return _ils(0, X, Y, bands, returnX)
