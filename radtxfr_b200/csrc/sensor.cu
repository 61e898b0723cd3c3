// Subsystems 3b and 4: apparent radiance and the MAKO instrument line shape.
//
//   rtx_ils_apply          ILS_MAKO weighted-mean resampling (radiative_transfer.py:1236-1256 triangle,
//                          ILS_MAKO.py:24-34 Gaussian) as a banded reduction: the reference materialises a
//                          dense (nX, nB) weight matrix of which <=3 % is non-zero; here one
//                          CTA owns one (band, block of spectra), walks only the band's support and forms the
//                          weights on the fly.  HBM-bound: each spectrum sample in the support is read once
//                          per band that covers it.
//   rtx_apparent_radiance  compute_LWIR_apparent_radiance (radiative_transfer.py:1017-1069)
//   rtx_hsi_pixels         per-pixel mixed-material radiance (LWIR_HSI_Generator.py:148-166)
#include <math.h>

#include "rtx_common.cuh"

namespace {

constexpr int ILS_THREADS = 256;
constexpr int ILS_SBLK = 8;   // spectra per CTA

template <int MODE>
__global__ void __launch_bounds__(ILS_THREADS) ils_kernel(const double* __restrict__ X, int64_t nx,
                                                          const double* __restrict__ Y, int64_t ns,
                                                          const double* __restrict__ centre,
                                                          const double* __restrict__ width, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int64_t s0 = (int64_t)blockIdx.y * ILS_SBLK;
  const int nsb = (int)min((int64_t)ILS_SBLK, ns - s0);
  const double c = centre[b], sg = width[b];
  __shared__ int64_t s_range[2];
  if (threadIdx.x == 0) {
    if (MODE == 0) {   // triangle: w>0 only for |x-c|<s
      s_range[0] = bisect_left_dev64(X, nx, c - sg);
      s_range[1] = bisect_right_dev64(X, nx, c + sg);
    } else {           // Gaussian: below 3e-18 of the peak beyond 9 sigma
      s_range[0] = bisect_left_dev64(X, nx, c - 9.5 * sg);
      s_range[1] = bisect_right_dev64(X, nx, c + 9.5 * sg);
      if (s_range[1] <= s_range[0]) { s_range[0] = 0; s_range[1] = nx; }   // centre outside the axis: sum it all
    }
  }
  __syncthreads();
  const int64_t lo = s_range[0], hi = s_range[1];
  double acc[ILS_SBLK];
#pragma unroll
  for (int k = 0; k < ILS_SBLK; ++k) acc[k] = 0.0;
  double wsum = 0.0;
  const double gnorm = sg * 2.5066282746310002;   // s*sqrt(2*pi) as numpy evaluates it
  for (int64_t i = lo + threadIdx.x; i < hi; i += ILS_THREADS) {
    const double x = X[i];
    double w;
    if (MODE == 0) {
      w = 1.0 - fabs(x - c) / sg;                 // tri(), radiative_transfer.py:1236-1239
      if (w < 0.0) w = 0.0;
    } else {
      const double u = (x - c) / sg;
      w = exp(-0.5 * (u * u)) / gnorm;            // g(), ILS_MAKO.py:24
    }
    wsum += w;
    const double* __restrict__ yr = Y + i * ns + s0;
#pragma unroll
    for (int k = 0; k < ILS_SBLK; ++k)
      if (k < nsb) acc[k] = fma(w, yr[k], acc[k]);
  }
  // block reduction (fixed order -> deterministic)
  __shared__ double s_red[ILS_THREADS / 32][ILS_SBLK + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wsum += __shfl_down_sync(0xffffffffu, wsum, o);
#pragma unroll
    for (int k = 0; k < ILS_SBLK; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < ILS_SBLK; ++k) s_red[warp][k] = acc[k];
    s_red[warp][ILS_SBLK] = wsum;
  }
  __syncthreads();
  if (threadIdx.x < nsb) {
    double a = 0.0, wt = 0.0;
    for (int q = 0; q < ILS_THREADS / 32; ++q) { a += s_red[q][threadIdx.x]; wt += s_red[q][ILS_SBLK]; }
    out[(int64_t)b * ns + s0 + threadIdx.x] = a / wt;   // Y_out = sum(ILS*Y)/N
  }
}

// Many-spectra variant (ns >= 16): lanes run along the spectrum index so that every warp load is one contiguous
// 256-byte row segment of Y[x][s0..s0+31]; the eight warps of a CTA take every eighth sample of the band's support.
// The Gaussian weight is below 3e-18 of its peak beyond 9 sigma, so its support is cut at 9.5 sigma (the
// reference sums the whole axis; the difference is far below one ulp of the result).
template <int MODE>
__global__ void __launch_bounds__(ILS_THREADS) ils_wide_kernel(const double* __restrict__ X, int64_t nx,
                                                               const double* __restrict__ Y, int64_t ns,
                                                               const double* __restrict__ centre,
                                                               const double* __restrict__ width,
                                                               double* __restrict__ out) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t sidx = (int64_t)blockIdx.y * 32 + lane;
  const bool live = sidx < ns;
  const double c = centre[b], sg = width[b];
  __shared__ int64_t s_range[2];
  if (threadIdx.x == 0) {
    const double half = (MODE == 0) ? sg : 9.5 * sg;
    s_range[0] = bisect_left_dev64(X, nx, c - half);
    s_range[1] = bisect_right_dev64(X, nx, c + half);
    if (MODE == 1 && s_range[1] <= s_range[0]) { s_range[0] = 0; s_range[1] = nx; }
  }
  __syncthreads();
  const int64_t lo = s_range[0], hi = s_range[1];
  const double gnorm = sg * 2.5066282746310002;
  double acc = 0.0, wsum = 0.0;
  const double* __restrict__ ycol = Y + (live ? sidx : 0);
  const double inv_sg = 1.0 / sg;
  // a warp takes 32 consecutive samples at a time: every lane forms ONE weight (coalesced read of X), then the 32
  // weights are broadcast in turn while all lanes stream the matching rows of Y (32 independent 256-byte loads)
  for (int64_t base = lo + 32 * warp; base < hi; base += 32 * (ILS_THREADS / 32)) {
    const int64_t ix = base + lane;
    double w = 0.0;
    if (ix < hi) {
      const double x = X[ix];
      if (MODE == 0) {
        w = 1.0 - fabs(x - c) * inv_sg;
        if (w < 0.0) w = 0.0;
      } else {
        const double u = (x - c) * inv_sg;
        w = exp(-0.5 * (u * u)) / gnorm;
      }
    }
    const int cnt = (int)min((int64_t)32, hi - base);
#pragma unroll 8
    for (int k = 0; k < cnt; ++k) {
      const double wk = __shfl_sync(0xffffffffu, w, k);
      wsum += wk;
      acc = fma(wk, ycol[(base + k) * ns], acc);
    }
  }
  __shared__ double s_acc[ILS_THREADS / 32][32], s_w[ILS_THREADS / 32];
  s_acc[warp][lane] = acc;
  if (lane == 0) s_w[warp] = wsum;
  __syncthreads();
  if (warp == 0 && live) {
    double a = 0.0, wt = 0.0;
#pragma unroll
    for (int q = 0; q < ILS_THREADS / 32; ++q) { a += s_acc[q][lane]; wt += s_w[q]; }
    out[(int64_t)b * ns + sidx] = a / wt;
  }
}

// Many-spectra variant, grouped (ns >= 16): one CTA owns G ADJACENT bands and 32 spectra and walks the UNION of the
// bands' supports once, so a sample row Y[x][s0..s0+31] (one coalesced 256-byte load per warp) feeds every band of
// the group that covers it -- MAKO bands overlap ~3.1-fold, so a band-per-CTA walk reads every sample 3.1 times,
// a group of 8 reads (8 + 2.1)/8 = 1.26 times the unique bytes and the rest of the overlap is between neighbouring
// CTAs that run at the same time (L2).  Lanes run along the spectrum index; per 32-sample step every lane forms the G
// weights of ONE sample (coalesced read of X) into the warp's shared-memory tile, then all lanes stream the 32 rows
// and take the weights as shared-memory broadcasts; bands whose support misses the step are skipped warp-uniformly.
// The eight warps interleave the steps; the final sum over warps is in fixed order (deterministic).
// Gaussian bands whose 9.5-sigma support holds no sample at all (centre outside the axis) fall back to the whole
// axis, which is what the reference sums (ILS_MAKO.py:24-34).
constexpr int ILS_G = 8;
template <int MODE>
__global__ void __launch_bounds__(ILS_THREADS) ils_group_kernel(const double* __restrict__ X, int64_t nx,
                                                                const double* __restrict__ Y, int64_t ns,
                                                                int n_bands, const double* __restrict__ centre,
                                                                const double* __restrict__ width,
                                                                double* __restrict__ out) {
  constexpr int NW = ILS_THREADS / 32;
  const int b0 = blockIdx.x * ILS_G;
  const int ng = min(ILS_G, n_bands - b0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t sidx = (int64_t)blockIdx.y * 32 + lane;
  const bool live = sidx < ns;
  __shared__ int64_t s_lo[ILS_G], s_hi[ILS_G];
  __shared__ double s_c[ILS_G], s_isg[ILS_G], s_gn[ILS_G];
  __shared__ double s_w[NW][32][ILS_G];
  __shared__ double s_acc[NW][ILS_G][32], s_ws[NW][ILS_G];
  if (threadIdx.x < ILS_G) {
    const int g = threadIdx.x;
    int64_t lo = 0, hi = 0;
    double c = 0.0, sg = 1.0;
    if (g < ng) {
      c = centre[b0 + g]; sg = width[b0 + g];
      const double half = (MODE == 0) ? sg : 9.5 * sg;
      lo = bisect_left_dev64(X, nx, c - half);
      hi = bisect_right_dev64(X, nx, c + half);
      if (MODE == 1 && hi <= lo) { lo = 0; hi = nx; }
    }
    s_lo[g] = lo; s_hi[g] = hi; s_c[g] = c; s_isg[g] = 1.0 / sg; s_gn[g] = sg * 2.5066282746310002;
  }
  __syncthreads();
  int64_t ulo = nx, uhi = 0;
#pragma unroll
  for (int g = 0; g < ILS_G; ++g)
    if (g < ng && s_hi[g] > s_lo[g]) { ulo = min(ulo, s_lo[g]); uhi = max(uhi, s_hi[g]); }
  double acc[ILS_G], wl[ILS_G];
#pragma unroll
  for (int g = 0; g < ILS_G; ++g) { acc[g] = 0.0; wl[g] = 0.0; }
  const double* __restrict__ ycol = Y + (live ? sidx : 0);
  for (int64_t base = ulo + 32 * warp; base < uhi; base += 32 * NW) {
    const int64_t ix = base + lane;
    const double x = ix < uhi ? X[ix] : 0.0;
    unsigned act = 0;                                   // bands whose support meets this step (warp-uniform)
#pragma unroll
    for (int g = 0; g < ILS_G; ++g) {
      double w = 0.0;
      if (g < ng && ix >= s_lo[g] && ix < s_hi[g]) {
        if (MODE == 0) {
          w = 1.0 - fabs(x - s_c[g]) * s_isg[g];        // tri(), radiative_transfer.py:1236-1239
          if (w < 0.0) w = 0.0;
        } else {
          const double u = (x - s_c[g]) * s_isg[g];
          w = exp(-0.5 * (u * u)) / s_gn[g];            // g(), ILS_MAKO.py:24
        }
      }
      s_w[warp][lane][g] = w;
      wl[g] += w;
      if (g < ng && base < s_hi[g] && base + 32 > s_lo[g]) act |= 1u << g;
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, uhi - base);
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const double y = ycol[(base + k) * ns];
#pragma unroll
      for (int g = 0; g < ILS_G; ++g)
        if (act & (1u << g)) acc[g] = fma(s_w[warp][k][g], y, acc[g]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int g = 0; g < ILS_G; ++g) {
    double t = wl[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    s_acc[warp][g][lane] = acc[g];
    if (lane == 0) s_ws[warp][g] = t;
  }
  __syncthreads();
  if (warp < ng && live) {
    const int g = warp;                                 // warp g finishes band g (ILS_G == warps per CTA)
    double a = 0.0, wt = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) { a += s_acc[q][g][lane]; wt += s_ws[q][g]; }
    out[(int64_t)(b0 + g) * ns + sidx] = a / wt;        // Y_out = sum(ILS*Y)/N
  }
}
static_assert(ILS_G == ILS_THREADS / 32, "one finishing warp per band of the group");

// thread per (x, a, t): one Planck evaluation shared by all emissivities
__global__ void apparent_kernel(const double* __restrict__ X, int64_t nx, const double* __restrict__ emis,
                                int64_t nE, const double* __restrict__ Ts, int64_t nA,
                                const double* __restrict__ tau, const double* __restrict__ la,
                                const double* __restrict__ ld, const double* __restrict__ dT, int64_t nT,
                                double* __restrict__ L, double* __restrict__ Ls) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * nA * nT) return;
  const int64_t t = i % nT, a = (i / nT) % nA, x = i / (nT * nA);
  const double T = dT ? Ts[a] + dT[t] : Ts[a];          // T_ = Ts[:,None] + dT[None,:]
  const double B = planck_wn(X[x], T);
  const double tu = tau[x * nA + a], up = la[x * nA + a], dn = ld[x * nA + a];
  for (int64_t e = 0; e < nE; ++e) {
    const double em = emis[x * nE + e];
    const double ls = em * B + (1.0 - em) * dn;         // Ls = em*B + (1-em)*Ld
    const int64_t o = ((x * nE + e) * nA + a) * nT + t;
    L[o] = tu * ls + up;                                // L = tau*Ls + La
    if (Ls) Ls[o] = ls;
  }
}

__global__ void hsi_kernel(const double* __restrict__ X, int64_t nx, const double* __restrict__ emis,
                           const double* __restrict__ tau, const double* __restrict__ la,
                           const double* __restrict__ ld, int64_t n_pix, const int32_t* __restrict__ atm,
                           const double* __restrict__ Tp, int n_mix, const int32_t* __restrict__ mat,
                           const double* __restrict__ frac, double* __restrict__ L) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix * nx) return;
  const int64_t k = i / nx, x = i % nx;
  double em = 0.0;
  for (int m = 0; m < n_mix; ++m) em += frac[k * n_mix + m] * emis[(int64_t)mat[k * n_mix + m] * nx + x];
  const int64_t ar = (int64_t)atm[k] * nx + x;
  const double B = planck_wn(X[x], Tp[k]);
  const double ls = em * B + (1.0 - em) * ld[ar];
  L[i] = tau[ar] * ls + la[ar];
}

// ---- reduceResolution / smooth (radiative_transfer.py:1266-1350) ------------------------------------------
// Symmetrised window smoothing: smFcn(y) = 0.5*(smooth(y) + smooth(y[::-1])[::-1]) with smooth() = reflect-pad,
// convolve with the normalised window, trim (ix0 = ceil(wl/2-1)).  In index form, with xr() the reflected extension,
//   smFcn(y)[i] = 0.5 * sum_j wn[j] * (xr(i + ix0 - j) + xr(i - ix0 + j)).
__global__ void smooth_sym_kernel(const double* __restrict__ Y, int64_t n, int64_t m, const double* __restrict__ wn,
                                  int wl, int ix0, int symmetric, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t i = idx / m, c = idx % m;
  double acc = 0.0;
  for (int j = 0; j < wl; ++j) {
    int64_t q1 = i + ix0 - j, q2 = i - ix0 + j;
    q1 = q1 < 0 ? -q1 : (q1 >= n ? 2 * (n - 1) - q1 : q1);
    q2 = q2 < 0 ? -q2 : (q2 >= n ? 2 * (n - 1) - q2 : q2);
    const double v = symmetric ? Y[q1 * m + c] + Y[q2 * m + c] : Y[q1 * m + c];
    acc = fma(wn[j], v, acc);
  }
  out[idx] = symmetric ? 0.5 * acc : acc;
}

// Cubic-spline interpolation (scipy interp1d(kind="cubic"): C2 cubic spline through all knots).  The spline's
// second derivatives decay like (2-sqrt(3))^k away from a knot, so each output point solves the tridiagonal system
// only on a window of HW knots either side of its interval (truncation < 1e-20 for HW = 40); the window ends use the
// natural condition, which is exact to rounding whenever the point is >= HW knots from the ends of the data -- always
// the case for reduceResolution's own output axis, which starts smFactor knots inside.
constexpr int SPL_HW = 40;
__global__ void spline_eval_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t n, int64_t m,
                                   const double* __restrict__ Xo, int64_t no, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= no * m) return;
  const int64_t t = idx / m, c = idx % m;
  const double xo = Xo[t];
  int64_t i = bisect_right_dev64(X, n, xo) - 1;          // X[i] <= xo < X[i+1]
  i = max((int64_t)0, min(n - 2, i));                    // ends: extrapolate with the end polynomial
  const int64_t k0 = max((int64_t)0, i - SPL_HW), k1 = min(n - 1, i + 1 + SPL_HW);
  const int nk = (int)(k1 - k0 + 1);                     // knots in the window, <= 2*HW+2
  double cp[2 * SPL_HW + 2], dp[2 * SPL_HW + 2];         // Thomas algorithm workspace (second derivatives)
  cp[0] = 0.0; dp[0] = 0.0;                              // natural: M_0 = 0
  for (int k = 1; k < nk - 1; ++k) {
    const int64_t g = k0 + k;
    const double h0 = X[g] - X[g - 1], h1 = X[g + 1] - X[g];
    const double rhs = 6.0 * ((Y[(g + 1) * m + c] - Y[g * m + c]) / h1 - (Y[g * m + c] - Y[(g - 1) * m + c]) / h0);
    const double den = 2.0 * (h0 + h1) - h0 * cp[k - 1];
    cp[k] = h1 / den;
    dp[k] = (rhs - h0 * dp[k - 1]) / den;
  }
  // back substitution down to the interval's two knots
  double Mn = 0.0;                                        // M_{nk-1} = 0
  const int ki = (int)(i - k0);
  double Mi1 = 0.0, Mi = 0.0;
  for (int k = nk - 2; k >= ki; --k) {
    const double Mk = (k >= 1) ? dp[k] - cp[k] * Mn : 0.0;
    if (k == ki + 1) Mi1 = Mk;
    if (k == ki) Mi = Mk;
    Mn = Mk;
  }
  if (ki + 1 == nk - 1) Mi1 = 0.0;
  const double h = X[i + 1] - X[i];
  const double A = (X[i + 1] - xo) / h, B = (xo - X[i]) / h;
  out[idx] = A * Y[i * m + c] + B * Y[(i + 1) * m + c] +
             ((A * A * A - A) * Mi + (B * B * B - B) * Mi1) * (h * h) / 6.0;
}

}  // namespace

extern "C" int rtx_smooth_sym(rtx_handle* h, const double* Y_dev, int64_t n, int64_t m, const double* window_dev,
                              int window_len, int symmetric, double* out_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n > 0 && m > 0 && window_len >= 1 && window_len <= n && Y_dev && window_dev && out_dev,
              "rtx_smooth_sym: bad arguments");
  DeviceGuard guard(h->device);
  const int ix0 = (window_len + 1) / 2 - 1;               // int(ceil(window_len/2 - 1)), radiative_transfer.py:1321
  const int64_t total = n * m;
  smooth_sym_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      Y_dev, n, m, window_dev, window_len, ix0, symmetric, out_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_spline_eval(rtx_handle* h, const double* X_dev, const double* Y_dev, int64_t n, int64_t m,
                               const double* Xo_dev, int64_t n_out, double* out_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n >= 4 && m > 0 && n_out > 0 && X_dev && Y_dev && Xo_dev && out_dev, "rtx_spline_eval: bad arguments");
  DeviceGuard guard(h->device);
  const int64_t total = n_out * m;
  spline_eval_kernel<<<(unsigned)((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      X_dev, Y_dev, n, m, Xo_dev, n_out, out_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_ils_apply(rtx_handle* h, int mode, const double* X_dev, int64_t nx, const double* Y_dev, int64_t ns,
                             int n_bands, const double* centre_dev, const double* width_dev, double* out_dev,
                             void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE((mode & ~0x101) == 0, "rtx_ils_apply: mode must be 0 (triangle) or 1 (Gaussian) [| 0x100: band-per-CTA walk]");
  RTX_REQUIRE(nx > 0 && ns > 0 && n_bands >= 0, "rtx_ils_apply: bad sizes");
  if (n_bands == 0) return RTX_OK;
  RTX_REQUIRE(X_dev && Y_dev && centre_dev && width_dev && out_dev, "rtx_ils_apply: null pointer");
  const bool wide = ns >= 16;
  const int64_t sblocks = wide ? (ns + 31) / 32 : (ns + ILS_SBLK - 1) / ILS_SBLK;
  RTX_REQUIRE(sblocks <= 65535, "rtx_ils_apply: too many spectra for one call (max 2097120)");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // grouped walk (each sample read once per 8 adjacent bands) when it still fills the machine; band-per-CTA otherwise
  const int64_t groups = (n_bands + ILS_G - 1) / ILS_G;
  if (wide && groups * sblocks >= 2LL * h->sm_count && !(mode & 0x100)) {
    dim3 grid((unsigned)groups, (unsigned)sblocks);
    if ((mode & 1) == 0) ils_group_kernel<0><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev);
    else ils_group_kernel<1><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev);
  } else {
    mode &= 1;
    dim3 grid((unsigned)n_bands, (unsigned)sblocks);
    if (wide) {
      if (mode == 0) ils_wide_kernel<0><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, centre_dev, width_dev, out_dev);
      else ils_wide_kernel<1><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, centre_dev, width_dev, out_dev);
    } else {
      if (mode == 0) ils_kernel<0><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, centre_dev, width_dev, out_dev);
      else ils_kernel<1><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, centre_dev, width_dev, out_dev);
    }
  }
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_apparent_radiance(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev,
                                     int64_t n_emis, const double* Ts_dev, int64_t n_atm, const double* tau_dev,
                                     const double* la_dev, const double* ld_dev, const double* dT_dev, int64_t n_dT,
                                     double* L_dev, double* ls_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && n_emis > 0 && n_atm > 0, "rtx_apparent_radiance: bad sizes");
  RTX_REQUIRE(X_dev && emis_dev && Ts_dev && tau_dev && la_dev && ld_dev && L_dev, "rtx_apparent_radiance: null pointer");
  if (!dT_dev) n_dT = 1;
  RTX_REQUIRE(n_dT > 0, "rtx_apparent_radiance: n_dT must be positive");
  DeviceGuard guard(h->device);
  const int64_t total = nx * n_atm * n_dT;
  const int64_t blocks = (total + 255) / 256;
  RTX_REQUIRE(blocks <= 2147483647LL, "rtx_apparent_radiance: problem too large for one launch");
  apparent_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X_dev, nx, emis_dev, n_emis, Ts_dev, n_atm, tau_dev, la_dev, ld_dev, dT_dev, n_dT, L_dev, ls_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_hsi_pixels(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                              const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm,
                              int64_t n_pix, const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix,
                              const int32_t* mat_idx_dev, const double* frac_dev, double* L_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && n_emis > 0 && n_atm > 0 && n_pix > 0 && n_mix > 0, "rtx_hsi_pixels: bad sizes");
  RTX_REQUIRE(X_dev && emis_dev && tau_dev && la_dev && ld_dev && atm_idx_dev && T_pix_dev && mat_idx_dev &&
                  frac_dev && L_dev, "rtx_hsi_pixels: null pointer");
  DeviceGuard guard(h->device);
  const int64_t total = n_pix * nx;
  const int64_t blocks = (total + 255) / 256;
  RTX_REQUIRE(blocks <= 2147483647LL, "rtx_hsi_pixels: problem too large for one launch");
  hsi_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X_dev, nx, emis_dev, tau_dev, la_dev, ld_dev, n_pix, atm_idx_dev, T_pix_dev, n_mix, mat_idx_dev, frac_dev, L_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}
