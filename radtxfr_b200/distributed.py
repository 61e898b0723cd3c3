"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the
single exchange the path has.

The path shards two ways, both exchange-free until the end:
  * by wavenumber band (config 5): every output grid point depends only on the lines within its wing window and on
    the atmosphere table, so rank r integrates the contiguous band [i0, i1) of the spectral axis with the full line
    table resident (the kernel's device-side binning picks the lines that reach the band -- a line halo, not a
    data halo) and the ranks all-gather (tau, Lu, Ld).  Bands start on multiples of the kernel's 2048-point tile
    and the tile size is pinned, so the gathered result is bit-identical to a single-GPU run.
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
    dist.all_gather_into_tensor(flat, pad.reshape(-1), group=group)      # 1-D in/out works on nccl and gloo
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
    dist.all_gather_into_tensor(flat, pad.reshape(-1), group=group)
    out = flat.view((world, width) + rest)
    full = torch.empty((n_atm,) + rest, dtype=local.dtype, device=local.device)
    for r, o in enumerate(owners):
        if len(o):
            full[torch.as_tensor(o, device=local.device)] = out[r, : len(o)]
    return full


def compute_TUD_banded(Xmin, Xmax, opts, group=None, **kwargs):
    """compute_TUD with the spectral axis sharded by band across the ranks of `group`.  Same arguments and return
    value as radiative_transfer.compute_TUD (single altitude / view angle shapes follow the same squeeze rules);
    every rank returns the full spectra.  Bit-identical to the single-GPU call."""
    import torch
    from . import engine, radiative_transfer as rt
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    opts.update(kwargs)
    db = rt._require_db(opts)
    Z = np.asarray(opts["Zs"], dtype=float); T = np.asarray(opts["Ts"], dtype=float)
    Z_s = np.array([opts["Altitudes"]]).ravel()
    mu_s = np.array([1.0 / np.cos(opts["theta_r"])]).ravel()
    X_ = rt.make_spectral_axis(Xmin, Xmax, opts["DVOUT"])
    bands = plan_bands(X_.size, world)
    i0, i1 = bands[rank]
    db.check_temperatures(T)
    p, w = rt.layer_weights(db, T, opts["Ps"], opts["PLs"], opts["MFs_VAL"], opts["MFs_ID"])
    nf = 2 * Z_s.size * mu_s.size + 1
    if i1 > i0:
        grid = engine.to_device(X_[i0:i1], db.device)
        vs = rt._voigt_settings(opts)
        vs["flags"] |= FORCE_K8
        od = db.voigt_sum(T, p, w, grid, **vs)
        tau, lu, ld = engine.tud(od, T[None, :], grid, Z, Z_s, mu_s, opts["N_angle"], opts["returnOD"])
        n = i1 - i0
        local = torch.cat([tau[0].reshape(n, -1).T, lu[0].reshape(n, -1).T, ld[0].reshape(1, n)], dim=0).contiguous()
    else:
        local = torch.zeros((nf, 0), dtype=torch.float64, device=db.device)
    full = all_gather_bands(local, bands, group)
    k = Z_s.size * mu_s.size
    nX = X_.size
    tau_ = full[:k].T.reshape(nX, Z_s.size, mu_s.size).cpu().numpy()
    Lu_ = full[k:2 * k].T.reshape(nX, Z_s.size, mu_s.size).cpu().numpy()
    Ld_ = full[2 * k].cpu().numpy()
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
