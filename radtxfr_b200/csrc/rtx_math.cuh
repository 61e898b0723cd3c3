// Table-driven FP64 exp and Planck radiance shared by the layer-integration and sensor kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "rtx_common.cuh"

// exp constants: 256/ln2, 1.5*2^52, -(ln2/256) high and low parts, 1/24, 1/6
static __constant__ double c_exp[6] = {369.3299304675746, 6755399441055744.0, -0.00270760617331689, -7.453964567463233e-13,
                                4.1666666666666664e-02, 1.6666666666666666e-01};

// exp(x) for |x| < 700 by the same table-driven scheme as the angle loop below (1-2 ulp): k = rint(256 x/ln2),
// r = x - k ln2/256, degree-4 polynomial, 2^(j/256) from the shared table, 2^m through the exponent field.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  const double tt = fma(x, c_exp[0], c_exp[1]);
  const double nf = tt - c_exp[1];
  double r = fma(nf, c_exp[2], x);
  r = fma(nf, c_exp[3], r);
  double pz = fma(r, c_exp[4], c_exp[5]);
  pz = fma(pz, r, 0.5);
  pz = fma(pz, r, 1.0);
  const int k = __double2loint(tt);
  const double tj = tab[k & 255];
  const double t = fma(tj, pz * r, tj);
  return __hiloint2double(__double2hiint(t) + ((k >> 8) << 20), __double2loint(t));
}
// Planck radiance (radiative_transfer.py:841-843) from per-point and per-layer factors: c1X3 = c1*X^3, c2X = c2*X
// (X = 100 nu), invT = 1/T; the quotient is a Newton-refined reciprocal (full double precision).
__device__ __forceinline__ double planck_fast(double c1X3, double c2X, double invT, const double* __restrict__ tab) {
  const double d = exp_tab(fmin(c2X * invT, 700.0), tab) - 1.0;
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
  const double e = fma(-d, r0, 1.0);
  const double q = fma(r0, fma(e, e, e), r0);
  return c1X3 * q * 1e4;
}


// fill the 256-entry table 2^(j/256) (correctly rounded by exp2()) that exp_tab reads; call by the whole block, then sync
__device__ __forceinline__ void exp_tab_init(double* tab) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = exp2((double)i * 0.00390625);
}
