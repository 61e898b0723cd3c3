/* The whole hot path through the C ABI of libradtxfr_b200 from plain C -- no Python, no torch: line table and
 * atmosphere table in; transmittance, upwelling and downwelling radiance out (compute_TUD,
 * radiative_transfer.py:274-392); then the MAKO triangle line shape on the three spectra (ILS_MAKO, :1072-1263).
 * This is the call sequence a cgo / JNI / ctypes host would make.
 *
 *   gcc -std=c11 -I include examples/tud_from_c.c -L radtxfr_b200 -lradtxfr_b200 -Wl,-rpath,$PWD/radtxfr_b200 -lm -o tud_from_c
 *
 * Prints "TUD i nu tau Lu Ld" for a few grid points and "BAND b X tau Lu Ld" for a few bands with 17 significant
 * digits; tests/test_gpu_parity.py::test_tud_from_plain_c compares them with the CPU oracle.
 * Exit status: 0 = ran on a GPU, 77 = no usable CUDA device, anything else = failure. */
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

static void* to_dev(rtx_handle* h, const void* host, size_t bytes) {
  void* d = NULL;
  if (rtx_malloc(h, bytes, &d) != RTX_OK) return NULL;
  if (host && rtx_memcpy_h2d(h, d, host, bytes, NULL) != RTX_OK) return NULL;
  return d;
}

int main(void) {
  rtx_handle* h = NULL;
  if (rtx_create(0, &h) != RTX_OK) {
    printf("no usable device: %s\n", rtx_last_error());
    return 77;
  }
  /* two isotopologues of two molecules (ids 1 and 2), partition functions Q(T) ~ T^1.5 and T^1.0 on 119 nodes */
  enum { NISO = 2, NLINES = 40, NL = 6, NMOL = 3, NX = 14000 };
  double mass[NISO] = {18.010565, 43.98983}, tipsT[RTX_TIPS_NPT], q[NISO * RTX_TIPS_NPT];
  for (int i = 0; i < RTX_TIPS_NPT; ++i) {
    tipsT[i] = 70.0 + (3000.0 - 70.0) * i / (RTX_TIPS_NPT - 1);
    q[i] = 170.0 * pow(tipsT[i] / 296.0, 1.5);
    q[RTX_TIPS_NPT + i] = 280.0 * (tipsT[i] / 296.0);
  }
  CK(rtx_set_iso_table(h, NISO, mass, tipsT, q));
  int32_t iso_mol[NISO] = {1, 2};
  double iso_factor[NISO] = {1.0, 1.0};
  CK(rtx_set_iso_molecules(h, NISO, iso_mol, iso_factor));
  /* 40 lines between 690 and 1410 cm^-1 from a fixed recurrence (reproduced by the test) */
  double nu[NLINES], sw[NLINES], el[NLINES], ga[NLINES], na[NLINES], da[NLINES];
  int32_t iso[NLINES];
  for (int i = 0; i < NLINES; ++i) {
    nu[i] = 700.0 + 17.5 * i + 3.0 * sin(1.7 * i);
    sw[i] = 1e-20 * (1.0 + 0.5 * cos(0.9 * i));
    el[i] = 100.0 + 60.0 * i;
    ga[i] = 0.05 + 0.03 * sin(0.3 * i) * sin(0.3 * i);
    na[i] = 0.5 + 0.2 * cos(0.5 * i) * cos(0.5 * i);
    da[i] = -0.002 + 0.0001 * i;
    iso[i] = i % 2;
  }
  CK(rtx_lines_upload(h, NLINES, nu, sw, el, ga, NULL, na, NULL, da, NULL, NULL, NULL, iso));
  /* a six-layer atmosphere: altitude [km], T [K], P [Pa], path length [km], mixing ratios [ppmv] of molecules 1, 2, 7 */
  double Z[NL] = {0.5, 1.5, 3.0, 6.0, 12.0, 25.0}, T[NL] = {290.0, 283.0, 272.0, 252.0, 217.0, 222.0};
  double P[NL] = {95000.0, 84000.0, 70000.0, 47000.0, 19000.0, 2500.0}, PL[NL] = {1.0, 1.0, 2.0, 4.0, 8.0, 18.0};
  double MF[NL * NMOL];
  int32_t mf_id[NMOL] = {1, 2, 7};                         /* molecule 7 has no lines: contributes nothing */
  for (int l = 0; l < NL; ++l) { MF[l * NMOL] = 9000.0 / (1 + l * l); MF[l * NMOL + 1] = 400.0; MF[l * NMOL + 2] = 2.0e5; }
  double* grid = (double*)malloc(sizeof(double) * NX);
  for (int i = 0; i < NX; ++i) grid[i] = 700.0 + (1400.0 - 700.0) * i / (NX - 1);
  double *dT = to_dev(h, T, sizeof T), *dP = to_dev(h, P, sizeof P), *dPL = to_dev(h, PL, sizeof PL), *dMF = to_dev(h, MF, sizeof MF);
  double* dgrid = to_dev(h, grid, sizeof(double) * NX);
  double* dY = to_dev(h, NULL, sizeof(double) * 3 * NX);   /* tau | Lu | Ld, one spectrum after the other */
  if (!dT || !dP || !dPL || !dMF || !dgrid || !dY) { fprintf(stderr, "device allocation failed: %s\n", rtx_last_error()); return 1; }
  double alt = 500.0, mu = 1.0;
  CK(rtx_tud_from_lines(h, 1, NL, dT, dP, dPL, dMF, NMOL, mf_id, 1.0, 0.0, 0.0, 50.0, 0.0, RTX_VOIGT_DEFAULT, dgrid, NX, Z,
                        &alt, 1, &mu, 1, 30, 0, dY, dY + NX, dY + 2 * NX, NULL));
  double* Y = (double*)malloc(sizeof(double) * 3 * NX);
  CK(rtx_memcpy_d2h(h, Y, dY, sizeof(double) * 3 * NX, NULL));
  CK(rtx_stream_sync(h, NULL));
  for (int i = 0; i < NX; i += 1999)
    printf("TUD %d %.17g %.17g %.17g %.17g\n", i, grid[i], Y[i], Y[NX + i], Y[2 * NX + i]);
  /* MAKO triangle line shape: bands from the library, spectra transposed to [nx][3] as rtx_ils_apply reads them */
  int nb = 0;
  CK(rtx_mako_bands(0, grid[0], grid[NX - 1], 0.0, 1.0, 0.0, 1.0, 0, NULL, NULL, NULL, &nb));
  double *Xb = (double*)malloc(sizeof(double) * nb), *cb = (double*)malloc(sizeof(double) * nb), *wb = (double*)malloc(sizeof(double) * nb);
  CK(rtx_mako_bands(0, grid[0], grid[NX - 1], 0.0, 1.0, 0.0, 1.0, nb, Xb, cb, wb, &nb));
  double* Yt = (double*)malloc(sizeof(double) * 3 * NX);
  for (int i = 0; i < NX; ++i) for (int s = 0; s < 3; ++s) Yt[3 * i + s] = Y[(size_t)s * NX + i];
  double *dYt = to_dev(h, Yt, sizeof(double) * 3 * NX), *dc = to_dev(h, cb, sizeof(double) * nb), *dw = to_dev(h, wb, sizeof(double) * nb);
  double* dR = to_dev(h, NULL, sizeof(double) * 3 * nb);
  if (!dYt || !dc || !dw || !dR) return 1;
  CK(rtx_ils_apply(h, 0, dgrid, NX, dYt, 3, nb, dc, dw, dR, NULL));
  double* R = (double*)malloc(sizeof(double) * 3 * nb);
  CK(rtx_memcpy_d2h(h, R, dR, sizeof(double) * 3 * nb, NULL));
  CK(rtx_stream_sync(h, NULL));
  printf("NBANDS %d\n", nb);
  for (int b = 0; b < nb; b += 25) printf("BAND %d %.17g %.17g %.17g %.17g\n", b, Xb[b], R[3 * b], R[3 * b + 1], R[3 * b + 2]);
  uint64_t launches = 0;
  CK(rtx_launch_count(h, &launches));
  printf("kernels launched: %llu\n", (unsigned long long)launches);
  void* bufs[] = {dT, dP, dPL, dMF, dgrid, dY, dYt, dc, dw, dR};
  for (size_t i = 0; i < sizeof bufs / sizeof bufs[0]; ++i) CK(rtx_free(h, bufs[i]));
  CK(rtx_destroy(h));
  free(grid); free(Y); free(Yt); free(Xb); free(cb); free(wb); free(R);
  return 0;
}
