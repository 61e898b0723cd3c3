"""Host-side engine: owns rtx handles and device buffers, and turns reference-style arguments into C-ABI calls.

PyTorch is used for device memory, streams and (in distributed.py) NCCL plumbing only; every computation on
the hot path is a kernel of libradtxfr_b200.so.  There is no CPU fallback.
"""
import ctypes
import os
import threading

import numpy as np

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spectro_tables.npz")
_tables = None
_lock = threading.Lock()


def spectro_tables():
    """Isotopologue metadata and TIPS-2011 tables (data extracted by tools/extract_reference_data.py from
    misc/hapi.py:3372-3497 and :5401-9564)."""
    global _tables
    if _tables is None:
        z = np.load(_DATA)
        t = {k: z[k] for k in z.files}
        t["index"] = {(int(m), int(i)): j for j, (m, i) in enumerate(zip(t["iso_M"], t["iso_I"]))}
        _tables = t
    return _tables


def _torch():
    import torch
    return torch


# RTX_VOIGT_* flag bits (include/radtxfr_b200.h)
VOIGT_FP32, VOIGT_WARPSPEC, VOIGT_LORENTZ, VOIGT_DOPPLER, VOIGT_NOSHIFT = 2, 4, 16, 32, 64
VOIGT_EXACT_RCP, VOIGT_NOFAR, VOIGT_TILE_SHIFT = 128, 4096, 8
VOIGT_SDVOIGT = 8192
VOIGT_NOTILEFAR = 16384
VOIGT_TILEFAR = 32768
VOIGT_NOSPLIT = 65536
MAX_EXTRA_DILUENTS = 4
# precision names of the shims -> flags.  "fp64" (default): FP64 with the one-Newton-step reciprocal (<= 9e-13 per
# term) and far-field blocks (< 1e-13); "fp64_fast" is the same thing (kept for round-1 callers); "fp64_exact":
# every (line, grid point) pair evaluated at the grid point with the fully refined reciprocal; "fp32": optional
# single-precision asymptote branch (1e-5 parity).
PRECISION_FLAGS = {"fp64": 0, "fp64_fast": 0, "fp64_exact": VOIGT_EXACT_RCP | VOIGT_NOFAR, "fp32": VOIGT_FP32}


def resolve_device(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.RtxError("radtxfr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.RtxError("radtxfr_b200 runs on CUDA devices only, got %s" % device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _hptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(_torch().cuda.current_stream(device).cuda_stream)


def to_device(x, device, dtype=None):
    """numpy / sequence / torch tensor -> contiguous CUDA tensor (float64 unless dtype is given)."""
    torch = _torch()
    dtype = dtype or torch.float64
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    npdt = {torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[dtype]
    a = np.ascontiguousarray(np.asarray(x, dtype=npdt))
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a).to(device)


class Handle:
    """An rtx_handle bound to one CUDA device."""

    def __init__(self, device=None):
        self.device = resolve_device(device)
        h = ctypes.c_void_p()
        _lib.call("rtx_create", self.device.index, ctypes.byref(h))
        self._h = h

    @property
    def h(self):
        return self._h

    def close(self):
        if self._h:
            _lib.lib().rtx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self):
        n = ctypes.c_uint64()
        _lib.call("rtx_launch_count", self._h, ctypes.byref(n))
        return int(n.value)

    def last_tile(self):
        """K (grid points per CTA / 256) of the last voigt_sum on this handle."""
        k = ctypes.c_int()
        _lib.call("rtx_voigt_last_tile", self._h, ctypes.byref(k))
        return int(k.value)

    def device_info(self):
        sm = ctypes.c_int()
        buf = ctypes.create_string_buffer(256)
        _lib.call("rtx_device_info", self._h, ctypes.byref(sm), buf, 256)
        return int(sm.value), buf.value.decode()

    def fp64_peak(self):
        v = ctypes.c_double()
        _lib.call("rtx_measure_fp64_peak", self._h, ctypes.byref(v))
        return float(v.value)


_handles = {}


def default_handle(device=None):
    """Shared line-less handle of a device (TUD, Planck, ILS, apparent radiance)."""
    dev = resolve_device(device)
    with _lock:
        if dev.index not in _handles:
            _handles[dev.index] = Handle(dev)
        return _handles[dev.index]


# --------------------------------------------------------------------------------------------- lines
_COLS_OPTIONAL = ("gamma_self", "n_self", "delta_self", "deltap_air", "deltap_self")


def table_fingerprint(cols, samples=32):
    """Cheap change detector for a hapi-style table (dict of columns): per column its identity, length and `samples`
    evenly spaced elements -- O(columns), not O(rows), so a cached LineDB can be validated on every call (a CRC of every
    column cost ~1 ms per call for 20 000 lines).  Appending or replacing rows / columns changes it; an in-place edit
    that leaves every sampled element unchanged does not -- call hapi_compat.invalidate(name) (or storeTable again)
    after such an edit."""
    parts = []
    for k in sorted(cols):
        v = cols[k]
        n = len(v)
        if n == 0:
            parts.append((k, id(v), 0))
            continue
        idx = range(0, n, max(1, n // samples))
        if isinstance(v, np.ndarray):
            sample = v[:: max(1, n // samples)].tobytes()
        else:
            sample = tuple(v[i] for i in idx)
        parts.append((k, id(v), n, hash(sample)))
    return tuple(parts)


class LineDB:
    """Device-resident line table (subsystem 1): SoA buffers sorted by line centre plus the compact
    isotopologue table of the (M, I) pairs that occur in it.  Built from hapi-style dict-of-columns tables
    (misc/hapi.py:438-463)."""

    def __init__(self, tables, device=None):
        """tables: one dict of columns or a list of them (hapi SourceTables)."""
        if isinstance(tables, dict):
            tables = [tables]
        self.handle = Handle(device)
        self.device = self.handle.device
        T = spectro_tables()
        cat = {}
        names = ["molec_id", "local_iso_id", "nu", "sw", "elower", "gamma_air", "n_air", "delta_air"]
        for req in names:
            for t in tables:
                if req not in t:
                    if req == "delta_air":      # missing shift column -> 0 (misc/hapi.py:11115-11119)
                        continue
                    raise KeyError("line table lacks column %r" % req)
        n_each = [len(t["nu"]) for t in tables]

        def column(name, required, dtype=np.float64):
            if not required and not any(name in t for t in tables):
                return None
            parts = []
            for t, n in zip(tables, n_each):
                parts.append(np.asarray(t[name], dtype=dtype) if name in t else np.zeros(n, dtype=dtype))
            return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, dtype=dtype)

        mol = column("molec_id", True, np.int64)
        iso = column("local_iso_id", True, np.int64)
        for k in ("nu", "sw", "elower", "gamma_air", "n_air", "delta_air"):
            cat[k] = column(k, True)
        for k in _COLS_OPTIONAL:
            cat[k] = column(k, False)
        self.n_lines = int(mol.size)
        key = mol * 1000 + iso                               # vectorised (M, I) grouping
        ukey, inverse = np.unique(key, return_inverse=True) if key.size else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        pairs = [(int(k // 1000), int(k % 1000)) for k in ukey]  # sorted like sorted(set(zip(M, I)))
        self.pairs = pairs
        self.pair_index = {mi: j for j, mi in enumerate(pairs)}
        idx = np.zeros(len(pairs), dtype=np.int64)
        for j, mi in enumerate(pairs):
            if mi not in T["index"]:
                raise Exception("cannot find component M,I = %d,%d." % mi)       # misc/hapi.py:11007
            idx[j] = T["index"][mi]
        n_iso = max(1, len(pairs))
        self.abundance = T["iso_abundance"][idx] if len(pairs) else np.ones(1)
        self.has_tips = T["tips_has"][idx] if len(pairs) else np.ones(1, dtype=bool)
        mass = np.ascontiguousarray(T["iso_mass"][idx]) if len(pairs) else np.ones(1)
        q = np.ascontiguousarray(T["tips_Q"][idx].astype(np.float64)) if len(pairs) else np.ones((1, 119))
        tT = np.ascontiguousarray(T["tips_T"].astype(np.float64))
        _lib.call("rtx_set_iso_table", self.handle.h, n_iso, _hptr(mass), _hptr(tT), _hptr(q))
        self.n_iso = n_iso
        iso_idx = np.ascontiguousarray(inverse.astype(np.int32))
        self.molecules = sorted(set(m for m, _ in pairs))
        self.iso_molecule = np.array([m for m, _ in pairs], dtype=np.int64)
        _lib.call("rtx_lines_upload", self.handle.h, self.n_lines, _hptr(cat["nu"]), _hptr(cat["sw"]),
                  _hptr(cat["elower"]), _hptr(cat["gamma_air"]), _hptr(cat["gamma_self"]), _hptr(cat["n_air"]),
                  _hptr(cat["n_self"]), _hptr(cat["delta_air"]), _hptr(cat["deltap_air"]),
                  _hptr(cat["delta_self"]), _hptr(cat["deltap_self"]), _hptr(iso_idx))
        _lib.call("rtx_set_iso_molecules", self.handle.h, n_iso,
                  _hptr(np.ascontiguousarray(self.iso_molecule.astype(np.int32)) if len(pairs) else np.zeros(1, np.int32)),
                  _hptr(np.ascontiguousarray(self.iso_factors(None, 1.0))))
        self.has_self = cat["gamma_self"] is not None
        # speed-dependence columns and broadening species other than air / self (hapi reads whatever gamma_<sp>,
        # n_<sp>, delta_<sp>, deltap_<sp>, SD_<sp> columns the table has: misc/hapi.py:11090-11128, 10883-10890)
        low = [{k.lower(): k for k in t} for t in tables]

        def column_ci(name):
            if not any(name in lo for lo in low):
                return None
            parts = [np.asarray(t[lo[name]], dtype=np.float64) if name in lo else np.zeros(n, dtype=np.float64)
                     for t, lo, n in zip(tables, low, n_each)]
            return np.ascontiguousarray(np.concatenate(parts))
        sd_air, sd_self = column_ci("sd_air"), column_ci("sd_self")
        self.has_sd = sd_air is not None or sd_self is not None
        if self.has_sd and self.n_lines:
            _lib.call("rtx_lines_set_sd", self.handle.h, _hptr(sd_air), _hptr(sd_self))
        species = sorted({k[6:] for lo in low for k in lo if k.startswith("gamma_")} - {"air", "self"})
        if len(species) > MAX_EXTRA_DILUENTS:
            raise ValueError("line table has %d broadening species besides air and self; at most %d are supported"
                             % (len(species), MAX_EXTRA_DILUENTS))
        self.extra_species = species
        for slot, sp in enumerate(species):
            cols = [column_ci(p + sp) for p in ("gamma_", "n_", "delta_", "deltap_", "sd_")]
            self.has_sd = self.has_sd or cols[4] is not None
            if self.n_lines:
                _lib.call("rtx_lines_set_diluent", self.handle.h, slot, *[_hptr(c) for c in cols])
        self.nu_min = float(cat["nu"].min()) if self.n_lines else 0.0
        self.nu_max = float(cat["nu"].max()) if self.n_lines else 0.0
        # host-side summary for cost estimates (distributed.estimate_tile_costs): sorted line centres and the extrema
        # that bound a line's wing
        self.nu_sorted = np.sort(cat["nu"])
        self.gamma_air_max = float(cat["gamma_air"].max()) if self.n_lines else 0.0
        self.n_air_range = (float(cat["n_air"].min()), float(cat["n_air"].max())) if self.n_lines else (0.0, 0.0)
        self.mass_min = float(np.min(mass))
        self.gamma_air_mean = float(cat["gamma_air"].mean()) if self.n_lines else 0.0
        self.n_air_mean = float(cat["n_air"].mean()) if self.n_lines else 0.0
        self.mass_mean = float(np.mean(np.asarray(mass)[inverse])) if self.n_lines else float(np.mean(mass))

    # -- abundance factors ------------------------------------------------------------------------
    def iso_factors(self, Components=None, factor=1.0):
        """factor / natural_abundance * abundance per isotopologue (misc/hapi.py:11136-11138); 0 for
        isotopologues excluded by Components (:11066)."""
        out = np.zeros(self.n_iso)
        if Components is None:
            for j in range(len(self.pairs)):
                out[j] = factor / self.abundance[j] * self.abundance[j]
            return out
        T = spectro_tables()
        for c in Components:
            mi = (int(c[0]), int(c[1]))
            if mi not in T["index"]:
                raise Exception("cannot find component M,I = %d,%d." % mi)
            if mi in self.pair_index:
                j = self.pair_index[mi]
                ab = c[2] if len(c) >= 3 else self.abundance[j]
                out[j] = factor / self.abundance[j] * ab
        return out

    def check_temperatures(self, T):
        T = np.asarray(T, dtype=float)
        if T.size and (T.min() < 70.0 or T.max() > 3000.0):
            raise Exception("TIPS: T must be between 70K and 3000K.")                # misc/hapi.py:9572

    def check_tips(self, weights_nonzero):
        for j, (mi, ok) in enumerate(zip(self.pairs, self.has_tips)):
            if not ok and weights_nonzero[j]:
                raise Exception("TIPS: no data for M,I = %d,%d." % mi)               # misc/hapi.py:9581

    # -- kernels ------------------------------------------------------------------------------------
    def _set_mix(self, x_extra):
        """Fractions of the extra broadening species ({species: x}) for the next kernel call (rtx_voigt_set_mix)."""
        x = np.zeros(MAX_EXTRA_DILUENTS)
        for sp, val in (x_extra or {}).items():
            if sp.lower() in self.extra_species:
                x[self.extra_species.index(sp.lower())] += float(val)
        _lib.call("rtx_voigt_set_mix", self.handle.h, MAX_EXTRA_DILUENTS, _hptr(x))

    def voigt_sum(self, T, p, weight, grid, x_air=1.0, x_self=0.0, wing_abs=0.0, wing_hw=50.0, s_threshold=0.0,
                  out=None, flags=0, x_extra=None):
        """out[s, i] = sum_lines weight[s, iso]*S(T_s)*V(grid[i]) (rtx_voigt_sum).  All arguments may be numpy
        or CUDA tensors; returns a CUDA float64 tensor [n_states, nx].  x_extra: {species: fraction} for broadening
        species other than air / self that the table has columns for (others contribute nothing, as in hapi)."""
        torch = _torch()
        dev = self.device
        if not any(isinstance(a, torch.Tensor) for a in (T, p, weight)):
            # host arrays (the shims' single-state calls): one staging array and one copy instead of three
            Th, ph = np.asarray(T, dtype=np.float64).reshape(-1), np.asarray(p, dtype=np.float64).reshape(-1)
            n_states = Th.size
            wh = np.asarray(weight, dtype=np.float64).reshape(n_states, self.n_iso)
            st = torch.from_numpy(np.concatenate([Th, ph, wh.reshape(-1)])).to(dev)
            T, p, weight = st[:n_states], st[n_states:2 * n_states], st[2 * n_states:].view(n_states, self.n_iso)
        T = to_device(T, dev).reshape(-1)
        p = to_device(p, dev).reshape(-1)
        n_states = T.numel()
        weight = to_device(weight, dev).reshape(n_states, self.n_iso)
        grid = to_device(grid, dev).reshape(-1)
        nx = grid.numel()
        if out is None:
            out = torch.empty((n_states, nx), dtype=torch.float64, device=dev)
        assert out.is_contiguous() and out.numel() == n_states * nx and out.dtype == torch.float64
        if nx == 0 or n_states == 0:
            return out
        if self.n_lines == 0:
            return out.zero_()
        if self.extra_species:
            self._set_mix(x_extra)
        _lib.call("rtx_voigt_sum", self.handle.h, n_states, _ptr(T), _ptr(p), _ptr(weight), float(x_air),
                  float(x_self), float(wing_abs), float(wing_hw), float(s_threshold), _ptr(grid), nx, _ptr(out),
                  int(flags), _stream(dev))
        return out

    def count_evals(self, T, p, weight, grid, x_air=1.0, x_self=0.0, wing_abs=0.0, wing_hw=50.0, s_threshold=0.0,
                    x_extra=None):
        """(evals, evals in the Weideman region) exactly as SURVEY.md 8d defines them (rtx_voigt_count)."""
        if self.extra_species:
            self._set_mix(x_extra)
        torch = _torch()
        dev = self.device
        T = to_device(T, dev).reshape(-1)
        p = to_device(p, dev).reshape(-1)
        n_states = T.numel()
        weight = to_device(weight, dev).reshape(n_states, self.n_iso)
        grid = to_device(grid, dev).reshape(-1)
        if self.n_lines == 0 or grid.numel() == 0:
            return 0, 0
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        # chunk the states so one launch stays below 2^31 (line, state) pairs
        per = max(1, int(2_000_000_000 // max(1, self.n_lines)))
        tot = np.zeros(2, dtype=np.int64)
        for s0 in range(0, n_states, per):
            s1 = min(n_states, s0 + per)
            _lib.call("rtx_voigt_count", self.handle.h, s1 - s0, _ptr(T[s0:s1]), _ptr(p[s0:s1]),
                      _ptr(weight[s0:s1].contiguous()), float(x_air), float(x_self), float(wing_abs),
                      float(wing_hw), float(s_threshold), _ptr(grid), grid.numel(), _ptr(counts), _stream(dev))
            tot += counts.cpu().numpy()
        return int(tot[0]), int(tot[1])

    def tips(self, T):
        """Q(T) for every isotopologue of this table: tensor [n, n_iso] (rtx_tips)."""
        torch = _torch()
        T = to_device(T, self.device).reshape(-1)
        out = torch.empty((T.numel(), self.n_iso), dtype=torch.float64, device=self.device)
        _lib.call("rtx_tips", self.handle.h, T.numel(), _ptr(T), _ptr(out), _stream(self.device))
        return out


    def layer_weights(self, T, P_pa, PL_km, MF_VAL, MF_ID):
        """rtx_layer_weights: (p_atm [n], weights [n, n_iso]) as CUDA tensors; T, P_pa, PL_km [n], MF_VAL [n, n_mol]."""
        torch = _torch()
        dev = self.device
        T = to_device(T, dev).reshape(-1)
        n = T.numel()
        P = to_device(P_pa, dev).reshape(n)
        PL = to_device(PL_km, dev).reshape(n)
        MF = to_device(MF_VAL, dev).reshape(n, -1)
        ids = np.ascontiguousarray(np.asarray(MF_ID).ravel().astype(np.int32))
        p_atm = torch.empty(n, dtype=torch.float64, device=dev)
        w = torch.empty((n, self.n_iso), dtype=torch.float64, device=dev)
        _lib.call("rtx_layer_weights", self.handle.h, n, _ptr(T), _ptr(P), _ptr(PL), _ptr(MF), ids.size, _hptr(ids),
                  _ptr(p_atm), _ptr(w), _stream(dev))
        return p_atm, w

    def tud_from_lines(self, T, P_pa, PL_km, MF_VAL, MF_ID, grid, Z, altitudes, mu, n_angle, return_od=False,
                       x_air=1.0, x_self=0.0, wing_abs=0.0, wing_hw=50.0, s_threshold=0.0, flags=0):
        """rtx_tud_from_lines: T, P_pa, PL_km [n_atm, nL], MF_VAL [n_atm, nL, n_mol] -> (tau, lu, ld) CUDA tensors of
        shapes [n_atm, nx, n_alt, n_mu], same, [n_atm, nx]."""
        torch = _torch()
        dev = self.device
        T = to_device(T, dev)
        if T.dim() == 1:
            T = T.unsqueeze(0)
        n_atm, nL = T.shape
        P = to_device(np.broadcast_to(np.asarray(P_pa, dtype=float), (n_atm, nL)) if not isinstance(P_pa, torch.Tensor) else P_pa, dev).reshape(n_atm, nL)
        PL = to_device(np.broadcast_to(np.asarray(PL_km, dtype=float), (n_atm, nL)) if not isinstance(PL_km, torch.Tensor) else PL_km, dev).reshape(n_atm, nL)
        MF = to_device(MF_VAL, dev).reshape(n_atm, nL, -1)
        ids = np.ascontiguousarray(np.asarray(MF_ID).ravel().astype(np.int32))
        grid = to_device(grid, dev).reshape(-1)
        nx = grid.numel()
        Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64).ravel())
        alt = np.ascontiguousarray(np.asarray(altitudes, dtype=np.float64).ravel())
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).ravel())
        tau = torch.empty((n_atm, nx, alt.size, mu.size), dtype=torch.float64, device=dev)
        lu = torch.empty_like(tau)
        ld = torch.empty((n_atm, nx), dtype=torch.float64, device=dev)
        _lib.call("rtx_tud_from_lines", self.handle.h, n_atm, nL, _ptr(T), _ptr(P), _ptr(PL), _ptr(MF), ids.size, _hptr(ids),
                  float(x_air), float(x_self), float(wing_abs), float(wing_hw), float(s_threshold), int(flags), _ptr(grid), nx,
                  _hptr(Z), _hptr(alt), alt.size, _hptr(mu), mu.size, int(n_angle), int(bool(return_od)), _ptr(tau), _ptr(lu),
                  _ptr(ld), _stream(dev))
        return tau, lu, ld


def mako_bands(mode, xmin, xmax, res_factor=None, fwhm_sf=1.0, shift=0.0, scale=1.0):
    """rtx_mako_bands (host only): (X_out, centres, widths) of the triangle (mode 0) or Gaussian (mode 1) MAKO bands."""
    n = ctypes.c_int()
    rf = float(res_factor) if res_factor else 0.0
    _lib.call("rtx_mako_bands", int(mode), float(xmin), float(xmax), rf, float(fwhm_sf), float(shift), float(scale), 0,
              None, None, None, ctypes.byref(n))
    X, c, w = (np.empty(max(n.value, 0)) for _ in range(3))
    if n.value:
        _lib.call("rtx_mako_bands", int(mode), float(xmin), float(xmax), rf, float(fwhm_sf), float(shift), float(scale),
                  n.value, _hptr(X), _hptr(c), _hptr(w), ctypes.byref(n))
    return X, c, w


# --------------------------------------------------------------------------------------------- line-less ops
def tud(od, T_layers, grid, Z, altitudes, mu, n_angle, return_od=False, device=None, split=False, out=None):
    """rtx_tud.  od [n_atm, nL, nx] (or [nL, nx]), T_layers [n_atm, nL]; returns (tau, lu, ld) CUDA tensors of
    shapes [n_atm, nx, n_alt, n_mu], same, [n_atm, nx].  split=True launches the two halves separately and returns
    (tau, lu, ld, ready): tau and lu are complete once the CUDA event `ready` has fired, while the downwelling half
    is still running."""
    torch = _torch()
    dev = od.device if isinstance(od, torch.Tensor) and od.is_cuda else resolve_device(device)
    h = default_handle(dev)
    od = to_device(od, dev)
    if od.dim() == 2:
        od = od.unsqueeze(0)
    n_atm, nL, nx = od.shape
    T_layers = to_device(T_layers, dev).reshape(n_atm, nL)
    grid = to_device(grid, dev).reshape(-1)
    assert grid.numel() == nx
    Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64).ravel())
    alt = np.ascontiguousarray(np.asarray(altitudes, dtype=np.float64).ravel())
    mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).ravel())
    assert Z.size == nL
    if out is not None:                       # caller-owned result buffers (flat tensors of sufficient size)
        k = n_atm * nx * alt.size * mu.size
        tau = out[0][:k].view(n_atm, nx, alt.size, mu.size)
        lu = out[1][:k].view(n_atm, nx, alt.size, mu.size)
        ld = out[2][: n_atm * nx].view(n_atm, nx)
    else:
        tau = torch.empty((n_atm, nx, alt.size, mu.size), dtype=torch.float64, device=dev)
        lu = torch.empty_like(tau)
        ld = torch.empty((n_atm, nx), dtype=torch.float64, device=dev)
    args = (h.h, n_atm, nL, _ptr(od), _ptr(T_layers), _ptr(grid), nx, _hptr(Z), _hptr(alt), alt.size, _hptr(mu), mu.size,
            int(n_angle), int(bool(return_od)))
    if not split:
        _lib.call("rtx_tud", *args, _ptr(tau), _ptr(lu), _ptr(ld), _stream(dev))
        return tau, lu, ld
    # two launches: tau and lu are complete at `ready` while the (longer) downwelling half still runs
    _lib.call("rtx_tud", *args, _ptr(tau), _ptr(lu), None, _stream(dev))
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    _lib.call("rtx_tud", *args, None, None, _ptr(ld), _stream(dev))
    return tau, lu, ld, ready


def planck(X, T, wavelength=False, device=None):
    torch = _torch()
    dev = resolve_device(device if not (isinstance(X, torch.Tensor) and X.is_cuda) else X.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    T = to_device(T, dev).reshape(-1)
    out = torch.empty((X.numel(), T.numel()), dtype=torch.float64, device=dev)
    if out.numel():
        _lib.call("rtx_planck", h.h, _ptr(X), X.numel(), _ptr(T), T.numel(), int(bool(wavelength)), _ptr(out),
                  _stream(dev))
    return out


def brightness(X, V, wavelength=False, inverse=False, bad_value=float("nan"), device=None):
    """rtx_brightness: V [nx, m] radiance -> brightness temperature (inverse=False) or back (inverse=True)."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(V, torch.Tensor) and V.is_cuda) else V.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    V = to_device(V, dev).reshape(X.numel(), -1)
    out = torch.empty_like(V)
    if out.numel():
        _lib.call("rtx_brightness", h.h, _ptr(X), X.numel(), _ptr(V), V.shape[1], int(bool(wavelength)),
                  int(bool(inverse)), float(bad_value), _ptr(out), _stream(dev))
    return out


def smooth(Y, window, symmetric, device=None):
    """rtx_smooth_sym: Y [n, m], window = normalised taps."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(Y, torch.Tensor) and Y.is_cuda) else Y.device)
    h = default_handle(dev)
    Y = to_device(Y, dev)
    assert Y.dim() == 2
    w = to_device(window, dev).reshape(-1)
    out = torch.empty_like(Y)
    _lib.call("rtx_smooth_sym", h.h, _ptr(Y), Y.shape[0], Y.shape[1], _ptr(w), w.numel(), int(bool(symmetric)), _ptr(out),
              _stream(dev))
    return out


def spline_eval(X, Y, Xo, device=None):
    """rtx_spline_eval: cubic spline through (X [n], Y [n, m]) evaluated at Xo -> [n_out, m]."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(Y, torch.Tensor) and Y.is_cuda) else Y.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    Y = to_device(Y, dev)
    Xo = to_device(Xo, dev).reshape(-1)
    out = torch.empty((Xo.numel(), Y.shape[1]), dtype=torch.float64, device=dev)
    _lib.call("rtx_spline_eval", h.h, _ptr(X), _ptr(Y), X.numel(), Y.shape[1], _ptr(Xo), Xo.numel(), _ptr(out),
              _stream(dev))
    return out


def ils_apply(mode, X, Y, centres, widths, device=None):
    """rtx_ils_apply: Y [nx, ns] -> [n_bands, ns]."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(Y, torch.Tensor) and Y.is_cuda) else Y.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    Y = to_device(Y, dev)
    assert Y.dim() == 2 and Y.shape[0] == X.numel()
    c = to_device(centres, dev).reshape(-1)
    w = to_device(widths, dev).reshape(-1)
    out = torch.empty((c.numel(), Y.shape[1]), dtype=torch.float64, device=dev)
    if out.numel():
        _lib.call("rtx_ils_apply", h.h, int(mode), _ptr(X), X.numel(), _ptr(Y), Y.shape[1], c.numel(), _ptr(c),
                  _ptr(w), _ptr(out), _stream(dev))
    return out


def apparent_radiance(X, emis, Ts, tau, La, Ld, dT=None, return_Ls=False, device=None, out_dtype=None):
    """rtx_apparent_radiance; out_dtype=torch.float32 selects the float32-output variant (rtx_apparent_radiance_f32)."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(tau, torch.Tensor) and tau.is_cuda) else tau.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    nx = X.numel()
    emis = to_device(emis, dev).reshape(nx, -1)
    Ts = to_device(Ts, dev).reshape(-1)
    nA = Ts.numel()
    tau, La, Ld = (to_device(a, dev).reshape(nx, nA) for a in (tau, La, Ld))
    dTt = to_device(dT, dev).reshape(-1) if dT is not None else None
    nT = dTt.numel() if dTt is not None else 1
    shape = (nx, emis.shape[1], nA) + ((nT,) if dTt is not None else ())
    f32 = out_dtype == torch.float32
    L = torch.empty(shape, dtype=torch.float32 if f32 else torch.float64, device=dev)
    Ls = torch.empty_like(L) if return_Ls else None
    if L.numel():
        _lib.call("rtx_apparent_radiance_f32" if f32 else "rtx_apparent_radiance", h.h, _ptr(X), nx, _ptr(emis),
                  emis.shape[1], _ptr(Ts), nA, _ptr(tau), _ptr(La), _ptr(Ld), _ptr(dTt), nT, _ptr(L), _ptr(Ls), _stream(dev))
    return (L, Ls) if return_Ls else L


def hsi_pixels(X, emis, tau, La, Ld, atm_idx, T_pix, mat_idx, frac, device=None, out_dtype=None):
    """rtx_hsi_pixels: emis [n_emis, nx], tau/La/Ld [n_atm, nx], per-pixel atm_idx [n_pix], T_pix [n_pix],
    mat_idx/frac [n_pix, n_mix] -> L [n_pix, nx] (float64, or float32 with out_dtype=torch.float32)."""
    torch = _torch()
    dev = resolve_device(device if not (isinstance(tau, torch.Tensor) and tau.is_cuda) else tau.device)
    h = default_handle(dev)
    X = to_device(X, dev).reshape(-1)
    nx = X.numel()
    emis = to_device(emis, dev).reshape(-1, nx)
    tau, La, Ld = (to_device(a, dev).reshape(-1, nx) for a in (tau, La, Ld))
    atm_idx = to_device(atm_idx, dev, torch.int32).reshape(-1)
    n_pix = atm_idx.numel()
    T_pix = to_device(T_pix, dev).reshape(n_pix)
    mat_idx = to_device(mat_idx, dev, torch.int32).reshape(n_pix, -1)
    frac = to_device(frac, dev).reshape(n_pix, mat_idx.shape[1])
    f32 = out_dtype == torch.float32
    L = torch.empty((n_pix, nx), dtype=torch.float32 if f32 else torch.float64, device=dev)
    if L.numel():
        _lib.call("rtx_hsi_pixels_f32" if f32 else "rtx_hsi_pixels", h.h, _ptr(X), nx, _ptr(emis), emis.shape[0], _ptr(tau),
                  _ptr(La), _ptr(Ld), tau.shape[0], n_pix, _ptr(atm_idx), _ptr(T_pix), mat_idx.shape[1], _ptr(mat_idx),
                  _ptr(frac), _ptr(L), _stream(dev))
    return L
