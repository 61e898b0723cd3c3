"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the
single exchange the path has.

The path shards two ways, both exchange-free until the end:
  * by wavenumber band (config 5): every output grid point depends only on the lines within its wing window and on
    the atmosphere table, so rank r integrates the contiguous band [i0, i1) of the spectral axis with the full line
    table resident (the kernel's device-side binning picks the lines that reach the band -- a line halo, not a
    data halo) and the ranks all-gather (tau, Lu, Ld).  Bands start on multiples of the kernel's 2048-point tile
    and the tile size is pinned, so the gathered result is bit-identical to a single-GPU run with the same tile size.
  * by atmosphere (config 3): rank r takes atmospheres r, r+world, ...; results are all-gathered (optionally after
    the MAKO ILS, which shrinks the exchange from nX to 128 values per spectrum).
Per-layer optical depth is never exchanged.
"""
import numpy as np

TILE = 2048          # grid points per CTA of the Voigt kernel at K=8
FORCE_K8 = 8 << 8    # RTX_VOIGT_TILE_SHIFT: pin the tile size so that sharded == unsharded bit for bit


def plan_bands(nx, world, tile=TILE, weights=None):
    """Split [0, nx) into `world` contiguous bands whose starts are multiples of `tile`.  With `weights` (one
    non-negative cost per tile, e.g. estimated evals) the bands are balanced by cost, otherwise by size.
    Returns a list of (i0, i1); trailing bands may be empty when nx is small."""
    ntiles = (nx + tile - 1) // tile
    if weights is None:
        weights = np.ones(ntiles)
    weights = np.asarray(weights, dtype=float)
    assert weights.size == ntiles
    cum = np.concatenate([[0.0], np.cumsum(weights)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        t = int(np.searchsorted(cum, target, side="left"))
        t = max(cuts[-1], min(ntiles, t))
        cuts.append(t)
    cuts.append(ntiles)
    return [(min(nx, cuts[r] * tile), min(nx, cuts[r + 1] * tile)) for r in range(world)]


def plan_atmospheres(n_atm, world):
    """Round-robin: rank r owns atmospheres r, r+world, ... (SURVEY.md 8e partitioning B)."""
    return [np.arange(r, n_atm, world) for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist


_band_feedback = {}      # (axis, table, world) -> per-tile cost scale learnt from measured per-rank times
_edge_cache = {}         # axis -> first wavenumber of every tile (host copy)
_gather_bufs = {}        # device -> (send shard, receive buffer), grow-only


def _all_gather_flat(out, inp, group=None):
    """all_gather_into_tensor of flat tensors.  NCCL takes device tensors directly; under the gloo backend (CPU tests,
    or several ranks sharing one GPU while debugging) device tensors are staged through the host."""
    import torch
    dist = _dist()
    if inp.is_cuda and dist.get_backend(group) == "gloo":
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(host, inp.cpu(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


def all_gather_bands(local, bands, group=None):
    """local: tensor [n_fields, i1-i0] holding this rank's band; returns [n_fields, nx] on every rank.
    One all_gather of equal-size (padded) shards -- the path's single collective."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    i0, i1 = bands[rank]
    assert local.shape[-1] == i1 - i0
    nf = local.shape[0]
    width = max(1, max(b - a for a, b in bands))
    pad = torch.zeros((nf, width), dtype=local.dtype, device=local.device)
    pad[:, : i1 - i0] = local
    flat = torch.empty(world * pad.numel(), dtype=local.dtype, device=local.device)
    _all_gather_flat(flat, pad.reshape(-1), group)                       # 1-D in/out works on nccl and gloo
    out = flat.view(world, nf, width)
    nx = bands[-1][1]
    full = torch.empty((nf, nx), dtype=local.dtype, device=local.device)
    for r, (a, b) in enumerate(bands):
        full[:, a:b] = out[r, :, : b - a]
    return full


def all_gather_atmospheres(local, owners, n_atm, group=None):
    """local: tensor [len(owners[rank]), ...]; returns [n_atm, ...] in global atmosphere order on every rank."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert local.shape[0] == len(owners[rank])
    rest = tuple(local.shape[1:])
    width = max(1, max(len(o) for o in owners))
    pad = torch.zeros((width,) + rest, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    flat = torch.empty(world * pad.numel(), dtype=local.dtype, device=local.device)
    _all_gather_flat(flat, pad.reshape(-1), group)
    out = flat.view((world, width) + rest)
    full = torch.empty((n_atm,) + rest, dtype=local.dtype, device=local.device)
    for r, o in enumerate(owners):
        if len(o):
            full[torch.as_tensor(o, device=local.device)] = out[r, : len(o)]
    return full


def estimate_tile_costs(db, T, p_atm, X, tile=TILE, wing_hw=50.0, wing_abs=0.0):
    """Relative cost of every `tile`-point tile of the axis X for the Voigt kernel ("balance bands by
    estimated evals, not by width" -- and since the far-field blocks made wing evaluations cheap, by estimated
    INSTRUCTIONS): see _tile_costs_from_edges.  Host-side, from the sorted line centres and the column statistics the
    LineDB keeps; an estimate for cutting bands, not a count."""
    X = np.asarray(X, dtype=float)
    ntiles = (X.size + tile - 1) // tile
    lo = X[np.arange(ntiles) * tile]
    hi = X[np.minimum((np.arange(ntiles) + 1) * tile, X.size) - 1]
    return _tile_costs_from_edges(db, T, p_atm, lo, hi, wing_hw, wing_abs, tile, X.size)


def _tile_costs_from_edges(db, T, p_atm, lo, hi, wing_hw, wing_abs, tile=TILE, nx=None, inside=None):
    """Cost model per tile = (lines centred in the tile) x sum over states of the cost of one (line, state) pair with
    the mean broadening of the table and the Doppler width AT THE TILE'S WAVENUMBER, in units of one exact wing
    evaluation: prologue ~60 per tile the window reaches; wing points 1 each up to the 2 x 350-point near zone and 1/5
    beyond it (32 nodes per 256 points plus visit overhead); windows of at most 128 points (pass 2) 3 per point;
    Weideman-core points (|x| + y < 15, half-width (15 - y) GammaD / sqrt(ln 2)) 35 each.  The constants come from the
    per-layer-group timings of config 2 (profiles/r02_voigt_microbench.txt); what matters for cutting bands is the trend
    they capture: Doppler widths grow with wavenumber, so the expensive core points pile up at the top of a wide axis
    (config 5: the upper half of 500-6000 cm^-1 holds half of the evaluations but ~3/4 of the work)."""
    T = np.asarray(T, dtype=float).ravel()
    p_atm = np.asarray(p_atm, dtype=float).ravel()
    nu = db.nu_sorted
    if nu.size == 0 or lo.size == 0:
        return np.ones(lo.size)
    if inside is None:
        inside = (np.searchsorted(nu, hi, side="right") - np.searchsorted(nu, lo, side="left")).astype(float)
    npts = np.full(lo.size, float(tile))
    if nx is not None:
        npts[-1] = nx - (lo.size - 1) * tile                                  # the last tile may be partial
    # the per-line cost varies smoothly along the axis: evaluate the model on at most 64 sample tiles, interpolate
    n_all = lo.size
    samp = np.unique(np.linspace(0, n_all - 1, min(n_all, 64)).round().astype(int))
    dv = np.maximum((hi[samp] - lo[samp]) / np.maximum(npts[samp] - 1.0, 1.0), 1e-300)   # grid spacing of the sample tiles
    mid = 0.5 * (lo[samp] + hi[samp])
    Ts, ps = T[:, None], np.abs(p_atm)[:, None]                                # [n_states, 1] against [1, n_tiles]
    g0 = ps * db.gamma_air_mean * (296.0 / Ts) ** db.n_air_mean
    gd = mid[None, :] * np.sqrt(2.0 * 1.380648813e-16 * Ts * np.log(2.0) / (db.mass_mean * 1.66053873e-24)) / 2.99792458e10
    w = np.maximum(np.maximum(wing_abs, wing_hw * g0), wing_hw * gd) / dv[None, :]    # half-window, grid points
    y = g0 / gd * 0.8325546111576977
    core = np.minimum(np.where(y < 15.0, 2.0 * (15.0 - y) * gd / 0.8325546111576977 / dv[None, :], 0.0), 2.0 * w)
    cheap = 2.0 * w - core
    wide = 60.0 * (1.0 + 2.0 * w / tile) + np.minimum(cheap, 700.0) + np.maximum(cheap - 700.0, 0.0) / 5.0
    per_line = (np.where(2.0 * w <= 128.0, 60.0 + 3.0 * cheap, wide) + 35.0 * core).sum(axis=0)
    return inside * np.interp(np.arange(n_all), samp, per_line) + 1.0


def compute_TUD_banded(Xmin, Xmax, opts, group=None, balance=True, to_host=True, **kwargs):
    """compute_TUD with the spectral axis sharded by band across the ranks of `group`.  Same arguments and return
    value as radiative_transfer.compute_TUD (single altitude / view angle shapes follow the same squeeze rules);
    every rank returns the full spectra -- numpy arrays, or CUDA tensors with to_host=False.
    Bands start on multiples of the kernel's 2048-point tile and the tile size is pinned (K = 8), so the gathered
    result is bit-identical to a single-GPU compute_TUD with opts["voigt"] = {"tile_k": 8} (the automatic tile size of
    the single-GPU call may pick 1024-point tiles, whose summation order differs).  balance=True cuts the bands by
    estimated cost (estimate_tile_costs) instead of by size, corrected from call to call by the measured per-rank
    compute times (the band plan does not affect the result: any tile-aligned plan gives the same bits).  The three spectra of every rank travel in ONE
    all_gather of equal-size shards; the host copy goes through the pinned, threaded path of compute_TUD."""
    import torch
    from . import engine, radiative_transfer as rt
    import time as _t0
    t_enter = _t0.perf_counter()
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    opts.update(kwargs)
    db = rt._require_db(opts)
    Z = np.asarray(opts["Zs"], dtype=float); T = np.asarray(opts["Ts"], dtype=float)
    Z_s = np.array([opts["Altitudes"]]).ravel()
    mu_s = np.array([1.0 / np.cos(opts["theta_r"])]).ravel()
    # the full axis lives on the device across calls (radiative_transfer._device_axis, LRU); on a cache hit the host
    # copy is rebuilt only after the kernels are in flight
    X_, grid_full = rt._device_axis(Xmin, Xmax, opts["DVOUT"], db.device)
    nX = grid_full.numel()
    axis_job = None
    if X_ is None:                            # the host axis (88 MB for config 5) is built on a side thread meanwhile
        import threading
        box = {}
        axis_job = threading.Thread(target=lambda: box.setdefault("X", rt.make_spectral_axis(Xmin, Xmax, opts["DVOUT"])))
        axis_job.start()
    db.check_temperatures(T)
    p, w = rt.layer_weights(db, T, opts["Ps"], opts["PLs"], opts["MFs_VAL"], opts["MFs_ID"])
    vs = rt._voigt_settings(opts)
    vs["flags"] = (vs["flags"] & ~(0xF << 8)) | FORCE_K8
    costs = None
    fb_key = None
    if balance:
        ntiles = (nX + TILE - 1) // TILE
        ekey = (float(Xmin), float(Xmax), nX, str(db.device))
        edges = _edge_cache.get(ekey)
        if edges is None:
            if len(_edge_cache) >= 4:
                _edge_cache.pop(next(iter(_edge_cache)))
            edges = torch.cat([grid_full[::TILE], grid_full[-1:]]).cpu().numpy()   # tile starts + last point
            lo_e, hi_e = edges[:-1], np.append(edges[1:-1], edges[-1])
            inside = (np.searchsorted(db.nu_sorted, hi_e, side="right") - np.searchsorted(db.nu_sorted, lo_e, side="left")).astype(float)
            edges = _edge_cache[ekey] = (lo_e, hi_e, inside, id(db))
        if edges[3] != id(db):
            edges = (edges[0], edges[1], None, id(db))
        costs = _tile_costs_from_edges(db, T, p, edges[0], edges[1], vs["wing_hw"], vs["wing_abs"], TILE, nX, edges[2])
        assert costs.size == ntiles
        # feedback: the model above is a first guess; the per-rank compute times of the previous call on the same axis
        # and table (CUDA events, all-gathered here: one 8-byte-per-rank exchange) rescale the cost of every rank's
        # tiles, so repeated calls -- the production pattern: one axis, many atmospheres -- converge to equal times
        fb_key = (float(Xmin), float(Xmax), nX, id(db), world)
        fb = _band_feedback.get(fb_key)
        if fb is not None and world > 1:
            e0, e1, prev_bands = fb["events"]
            e1.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=db.device)
            allms = torch.empty(world, dtype=torch.float64, device=db.device)
            _all_gather_flat(allms, ms, group)
            allms = allms.cpu().numpy()
            if np.all(allms > 0):
                # bounded step: one call disturbed from outside (a rank stalled by its host, another job on the box)
                # must not throw a balanced plan away -- a 2x outlier moves the cuts like a 15 % imbalance would
                ratio = np.clip(allms / allms.mean(), 0.87, 1.15)
                for r, (a, b) in enumerate(prev_bands):
                    fb["scale"][a // TILE:(b + TILE - 1) // TILE] *= ratio[r] ** 0.8
            fb["last_ms"] = allms
            fb.setdefault("history", []).append(([int(b - a) for a, b in prev_bands], [float(x) for x in allms]))
            del fb["history"][:-12]
        if fb is not None:
            costs = costs * fb["scale"]
    bands = plan_bands(nX, world, weights=costs)
    i0, i1 = bands[rank]
    k = Z_s.size * mu_s.size
    nf = 2 * k + 1
    # equal-size shards: the widest band rounded up to 64 tiles, so that the persistent send / receive buffers (grow-only,
    # per device) keep their size while the feedback nudges the cuts -- a fresh allocation would be a cudaMalloc, which
    # blocks the host until the queued kernels are done and serialises everything the host still has to do
    width = max(1, max(b - a for a, b in bands))
    width = (width + 64 * TILE - 1) // (64 * TILE) * (64 * TILE)
    bufs = _gather_bufs.get(str(db.device))
    if bufs is None or bufs[0].numel() < nf * width or bufs[1].numel() < world * nf * width or bufs[2][0].numel() < k * width:
        bufs = _gather_bufs[str(db.device)] = (torch.zeros(nf * width, dtype=torch.float64, device=db.device),
                                               torch.empty(world * nf * width, dtype=torch.float64, device=db.device),
                                               [torch.empty(k * width if j < 2 else width, dtype=torch.float64, device=db.device)
                                                for j in range(3)])
    shard = bufs[0][: nf * width].view(nf, width)                                  # padding beyond the band is never read
    import time as _time
    t_launch = _time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if i1 > i0:
        n = i1 - i0
        grid = grid_full[i0:i1]
        od = db.voigt_sum(T, p, w, grid, out=rt._od_workspace(db.device, T.size, n), **vs)
        tau, lu, ld = engine.tud(od, T[None, :], grid, Z, Z_s, mu_s, opts["N_angle"], opts["returnOD"], out=bufs[2])
        if k == 1:
            shard[0, :n] = tau.reshape(n); shard[1, :n] = lu.reshape(n)
        else:
            shard[:k, :n] = tau[0].reshape(n, k).T; shard[k:2 * k, :n] = lu[0].reshape(n, k).T
        shard[2 * k, :n] = ld[0]
    ev1.record()
    if fb_key is not None:
        fb = _band_feedback.setdefault(fb_key, {"scale": np.ones((nX + TILE - 1) // TILE)})
        fb["events"] = (ev0, ev1, bands)
        fb["host_prep_ms"] = 1e3 * (t_launch - t_enter)
        while len(_band_feedback) > 8:
            _band_feedback.pop(next(iter(_band_feedback)))
    flat = bufs[1][: world * nf * width].view(world, nf, width)
    _all_gather_flat(flat.view(-1), shard.reshape(-1), group)                   # the path's single collective
    ev2 = torch.cuda.Event(enable_timing=True); ev2.record()
    full = torch.cat([flat[r, :, : b - a] for r, (a, b) in enumerate(bands) if b > a], dim=1)     # [nf, nX]
    ev3 = torch.cuda.Event(enable_timing=True); ev3.record()
    if fb_key is not None:
        _band_feedback[fb_key]["tail_events"] = (ev1, ev2, ev3)
    if k == 1:
        tau_d, lu_d = full[0].reshape(nX, 1, 1), full[1].reshape(nX, 1, 1)
    else:
        tau_d = full[:k].T.contiguous().reshape(nX, Z_s.size, mu_s.size)
        lu_d = full[k:2 * k].T.contiguous().reshape(nX, Z_s.size, mu_s.size)
    ld_d = full[2 * k]
    if axis_job is not None:
        axis_job.join()
        X_ = box["X"]
    if to_host:
        tau_, Lu_, Ld_ = rt._to_host(tau_d, lu_d, ld_d)
    else:
        tau_, Lu_, Ld_ = tau_d, lu_d, ld_d
    if (len(Z_s) == 1) and (len(mu_s) == 1):
        tau_, Lu_ = tau_[:, 0, 0], Lu_[:, 0, 0]
    elif len(Z_s) == 1:
        tau_, Lu_ = tau_[:, 0, :], Lu_[:, 0, :]
    elif len(mu_s) == 1:
        tau_, Lu_ = tau_[:, :, 0], Lu_[:, :, 0]
    return X_, tau_, Lu_, Ld_


def compute_TUD_batch_sharded(Xmin, Xmax, Ts, MFs_VAL, opts, group=None, ils=None, gather=True, **kwargs):
    """radiative_transfer.compute_TUD_batch with the atmospheres dealt round-robin to the ranks of `group`.
    Returns (X, tau, Lu, Ld) with all n_atm atmospheres on every rank when gather=True (one all_gather per
    product), else only this rank's atmospheres plus their global indices."""
    from . import radiative_transfer as rt
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    Ts = np.asarray(Ts, dtype=float)
    MF = np.asarray(MFs_VAL, dtype=float)
    owners = plan_atmospheres(Ts.shape[0], world)
    mine = owners[rank]
    X, tau, lu, ld = rt.compute_TUD_batch(Xmin, Xmax, Ts[mine], MF[mine], opts=opts, ils=ils, **kwargs)
    if not gather:
        return X, tau, lu, ld, mine
    n = Ts.shape[0]
    return (X, all_gather_atmospheres(tau, owners, n, group), all_gather_atmospheres(lu, owners, n, group),
            all_gather_atmospheres(ld, owners, n, group))
