/* The C ABI of libradtxfr_b200 driven from plain C -- no Python, no torch: what a cgo / JNI / ctypes binding of the
 * reference's line loop (absorptionCoefficient_Voigt, misc/hapi.py:10906-11141) would call.
 *
 *   gcc -std=c11 -I include examples/voigt_from_c.c -L radtxfr_b200 -lradtxfr_b200 -Wl,-rpath,$PWD/radtxfr_b200 -lm -o voigt_from_c
 *
 * Exit status: 0 = ran on a GPU, 77 = the library loaded but found no usable CUDA device (its error text is
 * printed), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "radtxfr_b200.h"

#define CK(call)                                                              \
  do {                                                                        \
    int rc_ = (call);                                                         \
    if (rc_ != RTX_OK) {                                                      \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, rtx_last_error()); \
      return 1;                                                               \
    }                                                                         \
  } while (0)

int main(void) {
  printf("radtxfr_b200 C ABI version %d\n", rtx_version());
  rtx_handle* h = NULL;
  if (rtx_create(0, &h) != RTX_OK) {
    printf("no usable device: %s\n", rtx_last_error());
    return 77;
  }
  /* one isotopologue (H2O-like), partition function Q(T) ~ T^1.5 on the library's temperature nodes */
  double mass = 18.010565, tipsT[RTX_TIPS_NPT], q[RTX_TIPS_NPT];
  for (int i = 0; i < RTX_TIPS_NPT; ++i) {
    tipsT[i] = 70.0 + (3000.0 - 70.0) * i / (RTX_TIPS_NPT - 1);
    q[i] = 170.0 * pow(tipsT[i] / 296.0, 1.5);
  }
  CK(rtx_set_iso_table(h, 1, &mass, tipsT, q));
  /* three lines around 1000 cm^-1 */
  enum { NL = 3, NX = 20001 };
  double nu[NL] = {999.2, 1000.0, 1000.9}, sw[NL] = {1e-21, 3e-21, 5e-22}, el[NL] = {100.0, 250.0, 900.0};
  double ga[NL] = {0.07, 0.08, 0.06}, na[NL] = {0.7, 0.65, 0.75}, da[NL] = {-0.002, 0.0, 0.001};
  int32_t iso[NL] = {0, 0, 0};
  CK(rtx_lines_upload(h, NL, nu, sw, el, ga, NULL, na, NULL, da, NULL, NULL, NULL, iso));
  /* two states on a 0.0001 cm^-1 grid */
  double T[2] = {296.0, 230.0}, p[2] = {1.0, 0.05}, w[2] = {1.0, 1.0};
  double* grid = (double*)malloc(sizeof(double) * NX);
  double* out = (double*)malloc(sizeof(double) * 2 * NX);
  for (int i = 0; i < NX; ++i) grid[i] = 999.0 + 2.0 * i / (NX - 1);
  void *dT, *dp, *dw, *dg, *dout;
  CK(rtx_malloc(h, sizeof T, &dT)); CK(rtx_malloc(h, sizeof p, &dp)); CK(rtx_malloc(h, sizeof w, &dw));
  CK(rtx_malloc(h, sizeof(double) * NX, &dg)); CK(rtx_malloc(h, sizeof(double) * 2 * NX, &dout));
  CK(rtx_memcpy_h2d(h, dT, T, sizeof T, NULL)); CK(rtx_memcpy_h2d(h, dp, p, sizeof p, NULL));
  CK(rtx_memcpy_h2d(h, dw, w, sizeof w, NULL)); CK(rtx_memcpy_h2d(h, dg, grid, sizeof(double) * NX, NULL));
  CK(rtx_voigt_sum(h, 2, (const double*)dT, (const double*)dp, (const double*)dw, 1.0, 0.0, 0.0, 50.0, 0.0,
                   (const double*)dg, NX, (double*)dout, RTX_VOIGT_DEFAULT, NULL));
  CK(rtx_memcpy_d2h(h, out, dout, sizeof(double) * 2 * NX, NULL));
  CK(rtx_stream_sync(h, NULL));
  for (int s = 0; s < 2; ++s) {
    double mx = 0.0, area = 0.0;
    int at = 0;
    for (int i = 0; i < NX; ++i) {
      const double v = out[(size_t)s * NX + i];
      if (v > mx) { mx = v; at = i; }
      area += v * (2.0 / (NX - 1));
    }
    printf("state %d (T=%.0f K, p=%.2f atm): peak %.6e at %.4f cm^-1, integral %.6e\n", s, T[s], p[s], mx, grid[at], area);
    if (!(mx > 0.0) || !isfinite(area)) return 2;
  }
  int K = 0;
  CK(rtx_voigt_last_tile(h, &K));
  printf("tile size used: %d grid points per CTA\n", 256 * K);
  CK(rtx_free(h, dT)); CK(rtx_free(h, dp)); CK(rtx_free(h, dw)); CK(rtx_free(h, dg)); CK(rtx_free(h, dout));
  CK(rtx_destroy(h));
  free(grid); free(out);
  return 0;
}
