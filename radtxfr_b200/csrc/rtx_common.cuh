// Shared definitions for the radtxfr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/radtxfr_b200.h"

// ---- physical constants exactly as the reference spells them ----------
#define RTX_CBOLTZ 1.380648813e-16        // misc/hapi.py:84-92
#define RTX_CLIGHT 2.99792458e10
#define RTX_CMASS 1.66053873e-27          // misc/hapi.py:11085
#define RTX_C2_HAPI 1.4388028496642257    // misc/hapi.py:10171
#define RTX_TREF 296.0
#define RTX_LN2 0.6931471805599453        // numpy log(2.0)
#define RTX_SQRT_LN2 0.8325546111576977   // numpy sqrt(log(2.0))
#define RTX_PI 3.141592653589793
#define RTX_SQRT_PI 1.7724538509055159
#define RTX_INV_SQRT_PI 0.5641895835477563  // 1/sqrt(pi) as numpy evaluates it
#define RTX_WEIDEMAN_L 4.119534287814235    // sqrt(24/sqrt(2)), misc/hapi.py:9817
#define RTX_RT_C1 1.19104295315e-16         // radiative_transfer.py:71
#define RTX_RT_C2 1.43877736830e-02         // radiative_transfer.py:72

struct LineTable {
  int64_t n = 0;      // real lines
  int64_t n_pad = 0;  // padded length of the device arrays
  double *nu = nullptr, *s0 = nullptr, *elower = nullptr, *gamma_air = nullptr, *gamma_self = nullptr,
         *n_air = nullptr, *n_self = nullptr, *delta_air = nullptr, *deltap_air = nullptr,
         *delta_self = nullptr, *deltap_self = nullptr;
  int32_t* iso = nullptr;
  bool has_deltap = false;
  // speed-dependence columns SD_air / SD_self (absorptionCoefficient_SDVoigt) and extra broadening species
  // (gamma_<sp>, n_<sp>, delta_<sp>, deltap_<sp>, SD_<sp>), all optional
  double *sd_air = nullptr, *sd_self = nullptr;
  double *xg[RTX_MAX_EXTRA_DILUENTS] = {}, *xn[RTX_MAX_EXTRA_DILUENTS] = {}, *xd[RTX_MAX_EXTRA_DILUENTS] = {},
         *xdp[RTX_MAX_EXTRA_DILUENTS] = {}, *xsd[RTX_MAX_EXTRA_DILUENTS] = {};
  double xg_max[RTX_MAX_EXTRA_DILUENTS] = {}, xn_min[RTX_MAX_EXTRA_DILUENTS] = {}, xn_max[RTX_MAX_EXTRA_DILUENTS] = {};
  std::vector<int64_t> order;   // sorted position -> row of the caller's table (for columns uploaded later)
  // extrema used for the conservative per-state wing bound
  double nu_min = 0, nu_max = 0, gair_max = 0, gself_max = 0, nair_min = 0, nair_max = 0, nself_min = 0,
         nself_max = 0;
  // typical values (eval-weighted mean width sum(g^2)/sum(g), mean exponent) for the window statistic that picks the
  // kernel variant
  double gair_typ = 0, gself_typ = 0, nair_typ = 0, nself_typ = 0;
};

struct rtx_handle {
  int device = 0;
  int sm_count = 0;
  std::string name;
  int n_iso = 0;
  double *d_mass = nullptr, *d_tipsT = nullptr, *d_q = nullptr, *d_qref = nullptr;
  std::vector<double> h_mass;
  LineTable lines;
  // grow-only scratch
  double* d_state = nullptr;  size_t state_cap = 0;   // per-state tables
  double* d_part = nullptr;   size_t part_cap = 0;    // partial spectra of line-split launches (small problems)
  int32_t* d_bins = nullptr;  size_t bins_cap = 0;    // [n_states][n_tiles][2]
  double* h_pin = nullptr;                             // 2 pinned doubles for small read-backs
  double* d_cheb = nullptr;                            // Chebyshev tables of the Voigt far-field blocks
  uint64_t launches = 0;
  int last_K = 0;                                      // tile size of the last rtx_voigt_sum (-1: chosen on the device)
  const double* d_last_stat = nullptr;                 // window statistic that decided it
  void* last_stream = nullptr;
  double x_extra[RTX_MAX_EXTRA_DILUENTS] = {};
  int32_t* d_iso_mol = nullptr;  double* d_iso_factor = nullptr;     // rtx_set_iso_molecules
  double* d_pipe = nullptr;  size_t pipe_cap = 0;                    // scratch of rtx_tud_from_lines (doubles)         // fractions of the extra broadening species (rtx_voigt_set_mix)
};

void rtx_set_error(const std::string& msg);
int rtx_cuda_fail(cudaError_t e, const char* what);
#define RTX_CUDA(call)                                         \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return rtx_cuda_fail(_e, #call);    \
  } while (0)
#define RTX_REQUIRE(cond, msg)                                 \
  do {                                                         \
    if (!(cond)) { rtx_set_error(msg); return RTX_E_INVALID; } \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- small device helpers ----------------------------------------------------------------------
// bisect_right: number of elements <= v in the ascending array a[0..n)   (Python bisect.bisect,
// which the reference uses for both window ends, misc/hapi.py:11133-11134)
template <typename Ptr>
__device__ __forceinline__ int bisect_right_dev(Ptr a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (v < a[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}
// bisect_left: number of elements < v
template <typename Ptr>
__device__ __forceinline__ int bisect_left_dev(Ptr a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int64_t bisect_right_dev64(const double* a, int64_t n, double v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (v < a[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__device__ __forceinline__ int64_t bisect_left_dev64(const double* a, int64_t n, double v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Planck radiance in wavenumber form, radiative_transfer.py:841-843:
//   X*=100; L = c1*X**3/(exp(c2*X/T)-1); L*=1e4
__device__ __forceinline__ double planck_wn(double nu, double T) {
  double X = nu * 100.0;
  return RTX_RT_C1 * (X * X * X) / (exp(RTX_RT_C2 * X / T) - 1.0) * 1e4;
}
// wavelength form, radiative_transfer.py:836-839: X*=1e-6; L=c1/(X**5*(exp(c2/(X*T))-1)); L*=1e-4
__device__ __forceinline__ double planck_um(double lam, double T) {
  double X = lam * 1e-6;
  double X2 = X * X;
  return RTX_RT_C1 / ((X2 * X2 * X) * (exp(RTX_RT_C2 / (X * T)) - 1.0)) * 1e-4;
}
