/* CPU oracle, C restatement of the profile/accumulate part of hapi's Voigt line loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/ (and bench.py's CPU legs, if ever wanted) may load this; the product never does.  It exists so that the
 * GPU path can be compared with a CPU computation at the benchmark's FULL size (800 000 grid points x 66 layers x
 * 210 000 lines = 3.3e10 evaluations), which the numpy port (oracle/port.py, ~1e7 evals/s/core) cannot do in test time.
 * Pinned to oracle/port.py (itself pinned to the reference's golden vectors) by tests/test_oracle_golden.py.
 *
 * What it follows (paths relative to /root/reference):
 *   misc/hapi.py:11131-11138  window by bisect on the unshifted centre, accumulate factor*S*V in table row order
 *   misc/hapi.py:10131-10140, 9900-9915, 10022  PROFILE_VOIGT -> pcqsdhc PART1: V = (1/pi)*(sqrt(pi)*cte*Re w)
 *   misc/hapi.py:9833-9844    hum1_wei: w = (1/sqrt(pi)) t/(0.5+t^2), t = y-ix, replaced where |x|+y<15 by
 *   misc/hapi.py:9812-9827    cef: Weideman N=24 rational series (coefficients passed in by the caller)
 * The per-line quantities (S(T), GammaD, Gamma0, Shift0, wing, amplitude) are computed by oracle/port.py:line_state.
 *
 * Parallelism: OpenMP over contiguous slices of the grid; every slice walks all lines in row order, so the result
 * does not depend on the thread count.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>

static int64_t bisect_right(const double* a, int64_t n, double v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (v < a[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

static double re_w(double x, double y, const double* wa, double L) {
  const double rpi_inv = 1.0 / sqrt(M_PI);
  if (fabs(x) + y < 15.0) {
    const double complex z = x + y * I;
    const double complex den = L - I * z;
    const double complex Z = (L + I * z) / den;
    double complex p = 0.0;
    for (int k = 0; k < 24; ++k) p = p * Z + wa[k];          /* numpy.polyval's Horner order */
    const double complex w = 2.0 * p / (den * den) + rpi_inv / den;
    return creal(w);
  }
  const double complex t = y - I * x;
  return creal(rpi_inv * t / (0.5 + t * t));
}

/* out[i] += sum over kept lines of scale[l] * V_l(grid[i]);  scale = factor*ampl*S.  Returns the eval count. */
int64_t oracle_voigt_accumulate(int64_t n_lines, const double* nu, const double* shift, const double* gamD,
                                const double* gam0, const double* wing, const double* scale, const uint8_t* keep,
                                int64_t nx, const double* grid, const double* weideman_a, double weideman_L,
                                double* out, int n_slices) {
  int64_t evals = 0;
  if (n_slices < 1) n_slices = 1;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
  for (int s = 0; s < n_slices; ++s) {
    const int64_t a = nx * s / n_slices, b = nx * (s + 1) / n_slices;
    for (int64_t l = 0; l < n_lines; ++l) {
      if (!keep[l]) continue;
      int64_t lo = bisect_right(grid, nx, nu[l] - wing[l]);
      int64_t hi = bisect_right(grid, nx, nu[l] + wing[l]);
      if (lo < a) lo = a;
      if (hi > b) hi = b;
      if (hi <= lo) continue;
      const double cte = sqrt(log(2.0)) / gamD[l];
      const double sg0 = nu[l] + shift[l];
      const double y = gam0[l] * cte;
      const double pref = sqrt(M_PI) * cte;
      for (int64_t i = lo; i < hi; ++i) {
        const double x = -((sg0 - grid[i]) * cte);
        out[i] += scale[l] * ((1.0 / M_PI) * (pref * re_w(x, y, weideman_a, weideman_L)));
      }
      evals += hi - lo;
    }
  }
  return evals;
}
