"""Seeded synthetic workloads.

No real spectral lines exist anywhere near this repo: the reference ships its AER-v3.6 TAPE3 file as a
Git-LFS pointer.  Every test and benchmark therefore uses seeded synthetic line lists
with HITRAN-like statistics, generated here so that the CUDA path, the oracle and the CPU baseline all see
byte-identical inputs.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spectro_tables.npz")


def standard_atmosphere():
    """The reference's 66-layer 1976 US Standard Atmosphere table (radiative_transfer.py:75-148) as a dict
    with the keys compute_TUD consumes (radiative_transfer.py:171-176)."""
    sa = np.load(_DATA)["std_atmos"]
    return {
        "Zs": sa[:, 1].copy(), "Ts": sa[:, 5].copy(), "Ps": sa[:, 4].copy(), "PLs": sa[:, 3].copy(),
        "MFs_VAL": sa[:, 6:14] * 1e6, "MFs_ID": np.array([1, 2, 3, 4, 5, 6, 7, 22]),
    }


def synthetic_lines(seed, n, nu_lo, nu_hi, molec_id=1, iso_ids=(1, 2, 3, 4), iso_p=(0.7, 0.1, 0.1, 0.1)):
    """HITRAN-like random line table (dict of numpy columns, hapi column names misc/hapi.py:497-517),
    sorted by line centre.  Distributions are the ones SURVEY.md 8d fixes for config 1."""
    rng = np.random.default_rng(seed)
    nu = np.sort(rng.uniform(nu_lo, nu_hi, n))
    iso = rng.choice(np.asarray(iso_ids), size=n, p=np.asarray(iso_p))
    return {
        "molec_id": np.full(n, molec_id, dtype=np.int64),
        "local_iso_id": iso.astype(np.int64),
        "nu": nu,
        "sw": 10.0 ** rng.uniform(-27.0, -20.0, n),
        "elower": rng.uniform(0.0, 3000.0, n),
        "gamma_air": rng.uniform(0.02, 0.10, n),
        "gamma_self": rng.uniform(0.1, 0.5, n),
        "n_air": rng.uniform(0.4, 0.8, n),
        "delta_air": rng.uniform(-0.01, 0.005, n),
    }


def config1_lines(n=20000):
    """Config 1: synthetic H2O, seed 0, 775-1325 cm^-1."""
    return synthetic_lines(0, n, 775.0, 1325.0, 1)


def config2_tables(scale=1.0, nu_lo=690.0, nu_hi=1510.0, seed=1):
    """Config 2: H2O 3e4, CO2 8e4, O3 1e5 lines over 690-1510 cm^-1 (seed 1) -> {molec_id: table}."""
    out = {}
    for k, (mol, n, isos, ps) in enumerate([(1, 30000, (1, 2, 3, 4), (0.7, 0.1, 0.1, 0.1)),
                                            (2, 80000, (1, 2, 3, 4), (0.7, 0.1, 0.1, 0.1)),
                                            (3, 100000, (1, 2, 3), (0.8, 0.1, 0.1))]):
        out[mol] = synthetic_lines(seed * 1000 + k, max(1, int(round(n * scale))), nu_lo, nu_hi, mol, isos, ps)
    return out


CONFIG5_MOLECULES = (1, 2, 3, 4, 5, 6, 7, 22)             # options["MFs_ID"], radiative_transfer.py:177
CONFIG5_SHARES = (0.25, 0.25, 0.2, 0.08, 0.05, 0.12, 0.03, 0.02)


def config5_tables(n_lines=5_000_000, nu_lo=475.0, nu_hi=6025.0, seed=4):
    """Config 5: all eight gases of the reference's default MFs_ID over 500-6000 cm^-1: synthetic lines,
    seed 4, HITRAN-like shares per molecule -> {molec_id: table}."""
    return {m: synthetic_lines(seed * 1000 + k, max(1, int(n_lines * s)), nu_lo, nu_hi, m, (1,), (1.0,))
            for k, (m, s) in enumerate(zip(CONFIG5_MOLECULES, CONFIG5_SHARES))}


def merge_tables(tables_by_mol):
    """Concatenate per-molecule tables into one table sorted by line centre."""
    keys = None
    for t in tables_by_mol.values():
        keys = list(t.keys()) if keys is None else [k for k in keys if k in t]
    cat = {k: np.concatenate([np.asarray(t[k]) for t in tables_by_mol.values()]) for k in keys}
    order = np.argsort(cat["nu"], kind="stable")
    return {k: v[order] for k, v in cat.items()}


def perturbed_atmospheres(seed, n_atm, base=None):
    """Config 3: (T, H2O, O3) draws around the standard atmosphere, mirroring the outputs of the reference's
    generative model (GenerativeModel_AtmosInputs.py:225-242) and the call shape at
    Generate_LWIR_TUD.py:117-122.  Returns Ts [n_atm, nL] and MFs_VAL [n_atm, nL, nMol] (ppmv)."""
    base = standard_atmosphere() if base is None else base
    rng = np.random.default_rng(seed)
    z = base["Zs"]
    nL = z.size
    zz = (z - z.min()) / (z.max() - z.min())
    modes = np.stack([np.ones(nL), np.cos(np.pi * zz), np.cos(2 * np.pi * zz)], axis=0)
    dT = rng.normal(0.0, 8.0, (n_atm, 1)) + rng.normal(0.0, 3.0, (n_atm, 3)) @ modes / np.sqrt(3.0)
    Ts = np.clip(base["Ts"][None, :] + dT, 180.0, 330.0)
    MF = np.repeat(base["MFs_VAL"][None, :, :], n_atm, axis=0)
    h2o = np.exp(rng.normal(0.0, 0.5, (n_atm, 1)) + rng.normal(0.0, 0.2, (n_atm, 1)) * np.cos(np.pi * zz)[None, :])
    o3 = np.exp(rng.normal(0.0, 0.25, (n_atm, 1))) * np.ones((1, nL))
    MF[:, :, 0] *= h2o
    MF[:, :, 2] *= o3
    return Ts, MF


def synthetic_emissivities(seed, n_emis, X):
    """Config 4: smooth emissivity spectra clipped to [1e-4, 1-1e-4] (Generate_Emissivity_DB.py:44-46)."""
    rng = np.random.default_rng(seed)
    X = np.asarray(X, dtype=float)
    u = (X - X.min()) / max(X.max() - X.min(), 1e-300)
    e = np.full((X.size, n_emis), 0.0)
    for j in range(n_emis):
        base = rng.uniform(0.6, 0.98)
        y = np.full(X.size, base)
        for _ in range(4):
            c, w, a = rng.uniform(0, 1), rng.uniform(0.03, 0.3), rng.uniform(-0.25, 0.1)
            y += a * np.exp(-0.5 * ((u - c) / w) ** 2)
        e[:, j] = y
    return np.clip(e, 1e-4, 1 - 1e-4)
