"""radtxfr_b200 -- B200-native line-by-line radiative transfer for RadTxfr's hot path.

Submodules mirror the reference's modules for that path:
    radtxfr_b200.hapi_compat         absorptionCoefficient_Voigt, PYTIPS, ... (misc/hapi.py)
    radtxfr_b200.radiative_transfer  compute_TUD, compute_OD, planckian, compute_LWIR_apparent_radiance, ILS_MAKO
    radtxfr_b200.ILS_MAKO            Gaussian ILS_MAKO (ILS_MAKO.py)
    radtxfr_b200.engine              device-resident line database and kernel launches (C ABI via ctypes)
Importing the package needs neither a GPU nor the built library; the first computation does, and raises if
either is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
