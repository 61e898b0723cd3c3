// Subsystem 3: layer integration -- transmittance, upwelling and downwelling radiance (TUD).
//
// Replaces the numpy body of compute_TUD after the optical-depth loop (radiative_transfer.py:340-392).
// One thread owns one (grid point, atmosphere); optical depths are read layer-major so that a warp reads
// 256 contiguous bytes per layer.  The work is FP64-exp bound (2113 exp per grid point at 66 layers and 30
// angles), so the kernel keeps a block of angle accumulators in registers and reads each
// layer's optical depth once per angle block; Planck radiances are recomputed in registers instead of
// being materialised as the reference's (nX, nL) array.
#include <math.h>

#include "rtx_common.cuh"
#include "rtx_math.cuh"

namespace {

#ifndef RTX_TUD_ANGLE_BLOCK
#define RTX_TUD_ANGLE_BLOCK 10
#endif
#ifndef RTX_TUD_MINBLOCKS
#define RTX_TUD_MINBLOCKS 4
#endif
constexpr int TUD_THREADS = 128;
constexpr int ANGLE_BLOCK = RTX_TUD_ANGLE_BLOCK;
struct TudParams {
  int n_atm, n_layers, n_alt, n_mu, n_angle, n_down, return_od, max_cnt;
  int64_t nx;
  const double *od, *T, *grid;
  double *tau, *lu, *ld;
  double mu[RTX_MAX_MU];
  int cnt[RTX_MAX_ALT];
  uint32_t mask[RTX_MAX_ALT][RTX_MAX_LAYERS / 32];
  double inv_cos[RTX_MAX_ANGLES];   // only angles with a non-zero quadrature weight
  double wgt[RTX_MAX_ANGLES];       // cos(th)*sin(th)
  double wsum;                      // sum of cos(th)*sin(th) over ALL angles
};

__global__ void __launch_bounds__(TUD_THREADS, RTX_TUD_MINBLOCKS) tud_kernel(const TudParams P) {
  __shared__ double s_exp2[256];                      // 2^(j/256), correctly rounded by exp2()
  __shared__ double s_invT[RTX_MAX_LAYERS];           // 1/T of this block's atmosphere
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp2[i] = exp2((double)i * 0.00390625);
  for (int i = threadIdx.x; i < P.n_layers; i += blockDim.x) s_invT[i] = 1.0 / P.T[(int64_t)blockIdx.y * P.n_layers + i];
  __syncthreads();
  const int64_t x = (int64_t)blockIdx.x * TUD_THREADS + threadIdx.x;
  if (x >= P.nx) return;
  const unsigned lanes = __activemask();   // the last block of the grid may be partial
  const int atm = blockIdx.y;
  const double nu = P.grid[x];
  const double* __restrict__ od = P.od + ((int64_t)atm * P.n_layers) * P.nx + x;
  const double X100 = nu * 100.0;
  const double c1X3 = RTX_RT_C1 * (X100 * X100 * X100), c2X = RTX_RT_C2 * X100;
  const int64_t obase = ((int64_t)atm * P.nx + x) * P.n_alt * P.n_mu;

  // ---- upwelling: Lu <- t*Lu + (1-t)*B_k bottom-up, snapshot at every altitude's layer count
  //      (radiative_transfer.py:353-356)
  for (int m = 0; P.tau != nullptr && m < P.n_mu; ++m) {
    const double mu = P.mu[m];
    double lu = 0.0;
    for (int a = 0; a < P.n_alt; ++a)
      if (P.cnt[a] == 0) P.lu[obase + (int64_t)a * P.n_mu + m] = 0.0;
    for (int k = 0; k < P.max_cnt; ++k) {
      const double o = od[(int64_t)k * P.nx];
      const double B = planck_fast(c1X3, c2X, s_invT[k], s_exp2);
      const double t = exp_tab(fmin(fmax(-o * mu, -708.0), 708.0), s_exp2);
      lu = fma(t, lu - B, B);                                  // t*Lu + (1-t)*B
      for (int a = 0; a < P.n_alt; ++a)
        if (P.cnt[a] == k + 1) P.lu[obase + (int64_t)a * P.n_mu + m] = lu;
    }
    // ---- transmittance: exp(-sum_{Z<=zs} OD*mu) (radiative_transfer.py:348-352)
    for (int a = 0; a < P.n_alt; ++a) {
      double s = 0.0;
      for (int k = 0; k < P.n_layers; ++k)
        if ((P.mask[a][k >> 5] >> (k & 31)) & 1u) s += od[(int64_t)k * P.nx] * mu;
      P.tau[obase + (int64_t)a * P.n_mu + m] = P.return_od ? s : exp(-1.0 * s);
    }
  }

  // ---- downwelling: per angle Ld <- t*Ld + (1-t)*B_j top-down, then the cos-weighted hemispheric mean
  //      (radiative_transfer.py:368-373, 387-389).  The transmittances t = exp(-OD/cos th) are >99 % of the
  //      kernel's work, so exp is evaluated for a block of angles at once with the polynomial constants shared:
  //      table-driven range reduction (256 entries), degree-4 polynomial, scaling through the exponent field;
  //      10 FP64 operations per value, 1-2 ulp.
  if (P.ld == nullptr) return;              // upwelling / transmittance half only
  double ldsum = 0.0;
  for (int a0 = 0; a0 < P.n_angle; a0 += ANGLE_BLOCK) {
    double acc[ANGLE_BLOCK], sec[ANGLE_BLOCK];
    double secmax = 1.0;
#pragma unroll
    for (int a = 0; a < ANGLE_BLOCK; ++a) {
      acc[a] = 0.0;
      sec[a] = (a0 + a < P.n_angle) ? P.inv_cos[a0 + a] : 1.0;   // padding angles get weight 0 below
      secmax = fmax(secmax, sec[a]);
    }
    const double ao_max = 708.0 / secmax;
    for (int j = P.n_down - 1; j >= 0; --j) {
      // signed, like the reference's exp(-OD/cos th) (a continuum hook may hand in negative optical depths); bounded
      // above once per layer so that the per-angle exponent stays below the overflow threshold of the scaling trick
      // (matters only for optical depths below -708 cos(th_max), where the reference overflows as well)
      const double ao = fmin(-od[(int64_t)j * P.nx], ao_max);
      const double B = planck_fast(c1X3, c2X, s_invT[j], s_exp2);
      // optically thin layer for every lane of the warp and every angle of the block (|x| <= 2^-9): the degree-5
      // Taylor polynomial is exact to < 1e-19 and needs no range reduction
      if (__all_sync(lanes, fabs(ao) * secmax <= 1.953125e-3)) {
#pragma unroll
        for (int a = 0; a < ANGLE_BLOCK; ++a) {
          const double x = ao * sec[a];
          double t = fma(x, 8.333333333333333e-03, 4.1666666666666664e-02);
          t = fma(t, x, 1.6666666666666666e-01);
          t = fma(t, x, 0.5);
          t = fma(t, x, 1.0);
          t = fma(t, x, 1.0);
          acc[a] = fma(t, acc[a] - B, B);
        }
        continue;
      }
      // general case: exp(x) = 2^m * 2^(j/256) * exp(r),  k = rint(256 x / ln2) = 256 m + j,  r = x - k ln2/256
      // (two-piece), |r| <= ln2/512: degree-4 polynomial (truncation 3.8e-17), 2^(j/256) from a shared-memory table.
      // The constants sit in constant memory so the FP64 instructions read them as operands.
      double xa[ANGLE_BLOCK], tt[ANGLE_BLOCK], r[ANGLE_BLOCK], pz[ANGLE_BLOCK];
#pragma unroll
      for (int a = 0; a < ANGLE_BLOCK; ++a) {
        xa[a] = fmax(ao * sec[a], -708.0);                               // exp(-708) = 3e-308: nothing below matters
        tt[a] = fma(xa[a], c_exp[0], c_exp[1]);                          // 256/ln2; 1.5*2^52 puts k in the low word
      }
#pragma unroll
      for (int a = 0; a < ANGLE_BLOCK; ++a) {
        const double nf = tt[a] - c_exp[1];
        r[a] = fma(nf, c_exp[2], xa[a]);                                 // -ln2/256, high part (21 trailing zero bits)
        r[a] = fma(nf, c_exp[3], r[a]);                                  // -ln2/256, low part
        pz[a] = fma(r[a], c_exp[4], c_exp[5]);                           // 1/24, 1/6
      }
#pragma unroll
      for (int a = 0; a < ANGLE_BLOCK; ++a) pz[a] = fma(pz[a], r[a], 0.5);
#pragma unroll
      for (int a = 0; a < ANGLE_BLOCK; ++a) pz[a] = fma(pz[a], r[a], 1.0);
#pragma unroll
      for (int a = 0; a < ANGLE_BLOCK; ++a) {
        const int k = __double2loint(tt[a]);
        const double tj = s_exp2[k & 255];
        double t = fma(tj, pz[a] * r[a], tj);                            // 2^(j/256) * (1 + (exp(r)-1))
        t = __hiloint2double(__double2hiint(t) + ((k >> 8) << 20), __double2loint(t));
        acc[a] = fma(t, acc[a] - B, B);                                  // t*Ld + (1-t)*B
      }
    }
#pragma unroll
    for (int a = 0; a < ANGLE_BLOCK; ++a)
      if (a0 + a < P.n_angle) ldsum += acc[a] * P.wgt[a0 + a];
  }
  P.ld[(int64_t)atm * P.nx + x] = ldsum / P.wsum;
}

__global__ void planck_kernel(const double* __restrict__ X, int64_t nx, const double* __restrict__ T, int64_t nT,
                              int wavelength, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * nT) return;
  const double x = X[i / nT], t = T[i % nT];
  out[i] = wavelength ? planck_um(x, t) : planck_wn(x, t);
}

// brightnessTemperature (radiative_transfer.py:851-933) and BT2L (:936-1014), element-wise over [nx][m]
__global__ void bt_kernel(const double* __restrict__ X, int64_t nx, const double* __restrict__ V, int64_t m,
                          int wavelength, int inverse, double bad, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * m) return;
  const double x = X[i / m], v = V[i];
  double r;
  if (!inverse) {                                   // radiance -> brightness temperature
    if (wavelength) {
      const double Xm = x * 1e-6, L = v * 1e4;      // X *= 1e-6; L *= 1e4
      const double X2 = Xm * Xm;
      r = RTX_RT_C2 / (Xm * log(1.0 + RTX_RT_C1 / ((X2 * X2 * Xm) * L)));
    } else {
      const double Xm = x * 100.0, L = v * 1e-4;    // X *= 100; L *= 1e-4
      r = RTX_RT_C2 * Xm / log(RTX_RT_C1 * (Xm * Xm * Xm) / L + 1.0);
    }
    if (!isfinite(v) || v <= 0.0) r = bad;          // ixBad = ~isfinite(L) | (L <= 0)
  } else {                                          // brightness temperature -> radiance
    r = wavelength ? planck_um(x, v) : planck_wn(x, v);
    if (!isfinite(r) || v <= 0.0) r = bad;          // ixBad = ~isfinite(L) | (T <= 0)
  }
  out[i] = r;
}

}  // namespace

extern "C" int rtx_brightness(rtx_handle* h, const double* X_dev, int64_t nx, const double* V_dev, int64_t m,
                              int wavelength, int inverse, double bad_value, double* out_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && m > 0 && X_dev && V_dev && out_dev, "rtx_brightness: bad arguments");
  DeviceGuard guard(h->device);
  const int64_t total = nx * m;
  bt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X_dev, nx, V_dev, m, wavelength, inverse, bad_value, out_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_tud(rtx_handle* h, int n_atm, int n_layers, const double* od_dev, const double* T_dev,
                       const double* grid_dev, int64_t nx, const double* Z_host, const double* alt_host, int n_alt,
                       const double* mu_host, int n_mu, int n_angle, int return_od, double* tau_dev, double* lu_dev,
                       double* ld_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n_atm > 0 && n_atm <= 65535 && nx > 0, "rtx_tud: n_atm must be in 1..65535 and nx positive");
  RTX_REQUIRE(n_layers > 0 && n_layers <= RTX_MAX_LAYERS, "rtx_tud: n_layers out of range");
  RTX_REQUIRE(n_alt > 0 && n_alt <= RTX_MAX_ALT, "rtx_tud: n_alt out of range");
  RTX_REQUIRE(n_mu > 0 && n_mu <= RTX_MAX_MU, "rtx_tud: n_mu out of range");
  RTX_REQUIRE(n_angle > 0 && n_angle <= RTX_MAX_ANGLES, "rtx_tud: N_angle out of range");
  RTX_REQUIRE(od_dev && T_dev && grid_dev && Z_host && alt_host && mu_host, "rtx_tud: null pointer");
  RTX_REQUIRE((tau_dev != nullptr) == (lu_dev != nullptr) && (tau_dev || ld_dev),
              "rtx_tud: pass tau and lu together, or ld, or all three");
  DeviceGuard guard(h->device);
  TudParams P{};
  P.n_atm = n_atm; P.n_layers = n_layers; P.n_alt = n_alt; P.n_mu = n_mu; P.return_od = return_od ? 1 : 0;
  P.nx = nx; P.od = od_dev; P.T = T_dev; P.grid = grid_dev; P.tau = tau_dev; P.lu = lu_dev; P.ld = ld_dev;
  for (int m = 0; m < n_mu; ++m) P.mu[m] = mu_host[m];
  P.max_cnt = 0;
  for (int a = 0; a < n_alt; ++a) {
    int c = 0;
    for (int k = 0; k < n_layers; ++k)
      if (Z_host[k] <= alt_host[a]) { P.mask[a][k >> 5] |= (1u << (k & 31)); ++c; }   // ix = Z <= zs
    P.cnt[a] = c;                                                                     // nL = np.sum(ix)
    if (c > P.max_cnt) P.max_cnt = c;
  }
  P.n_down = P.cnt[n_alt - 1];   // the reference's nL after the altitude loop (radiative_transfer.py:353, 370)
  // angles = linspace(0, pi/2, nA, endpoint=False) (radiative_transfer.py:368)
  const double step = (M_PI / 2.0) / n_angle;
  P.wsum = 0.0;
  int kept = 0;
  for (int i = 0; i < n_angle; ++i) {
    const double th = i * step;
    const double w = cos(th) * sin(th);
    P.wsum += w;
    if (w != 0.0) { P.inv_cos[kept] = 1.0 / cos(th); P.wgt[kept] = w; ++kept; }
  }
  P.n_angle = kept;
  if (P.wsum == 0.0) P.wsum = nan("");   // N_angle=1: the reference divides 0 by 0
  dim3 grid((unsigned)((nx + TUD_THREADS - 1) / TUD_THREADS), (unsigned)n_atm);
  tud_kernel<<<grid, TUD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(P);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_planck(rtx_handle* h, const double* X_dev, int64_t nx, const double* T_dev, int64_t nT,
                          int wavelength, double* out_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && nT > 0 && X_dev && T_dev && out_dev, "rtx_planck: bad arguments");
  DeviceGuard guard(h->device);
  const int64_t total = nx * nT;
  planck_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(X_dev, nx, T_dev, nT,
                                                                                              wavelength, out_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}
