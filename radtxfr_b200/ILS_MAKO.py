"""Drop-in for the reference's standalone ILS_MAKO.py (Gaussian MAKO line shape, ILS_MAKO.py:2-35)."""
from . import radiative_transfer as _rt


def ILS_MAKO(X, Y):
    """Gaussian instrument line shape, sigma_b = |gradient(X_out)|, all 128 bands, infinite support.
    Returns (X_out, Y_out) like the reference."""
    return _rt._ils(1, X, Y, _rt._mako_bands_gauss(), True)
