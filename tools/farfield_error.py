"""Interpolation error of the Voigt kernel's far-field blocks (csrc/voigt.cu, RTX_FAR_FRAC).

A block of B grid points is represented by p Chebyshev nodes; a line whose centre lies `frac`*B points beyond the
block's nearest point is evaluated at the nodes only.  Prints the worst relative error of the degree p-1 interpolant
of the closed form a*(u+h)/(u^2+c) over the block, for several (Gamma0, GammaD) pairs.  The error depends on
frac only (scale-free), which is how RTX_FAR_FRAC = 0.375 (p = 32) was chosen: < 1e-13.

    python tools/farfield_error.py
"""
import numpy as np
from numpy.polynomial import chebyshev as C


def err(p, B, dmin, g0=0.05, gd=0.001, dv=0.001):
    half = B / 2.0
    k = np.arange(p)
    xn = np.cos(np.pi * (2 * k + 1) / (2 * p))
    xf = (np.arange(B) - (B - 1) / 2.0) / half
    h = gd * gd / np.log(2)
    m = g0 * g0 - 0.5 * h
    c = 2 * h * g0 * g0

    def f(x):
        d = (x * half + (B - 1) / 2.0 + dmin) * dv      # line centre dmin points before the block's first point
        u = d * d + m
        return (u + h) / (u * u + c)

    coef = C.chebfit(xn, f(xn), p - 1)
    return np.max(np.abs(C.chebval(xf, coef) - f(xf)) / f(xf))


if __name__ == "__main__":
    widths = [(0.05, 0.001), (0.1, 0.002), (0.002, 0.002), (0.0001, 0.001), (0.02, 0.0005), (0.5, 0.001)]
    fracs = [0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.375, 0.4, 0.5]
    print("p    B   " + "  ".join("%7.3f" % f for f in fracs))
    for p, B in [(32, 128), (32, 256), (32, 2048), (16, 128), (24, 256), (64, 2048)]:
        print("%-4d %-4d" % (p, B), "  ".join("%7.1e" % max(err(p, B, f * B, g0, gd) for g0, gd in widths) for f in fracs))
