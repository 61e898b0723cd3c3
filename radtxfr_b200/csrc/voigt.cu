// Subsystem 2: Voigt line sum (sm_100a).
//
// Replaces the per-line Python loop of absorptionCoefficient_Voigt (misc/hapi.py:11050-11138) and the
// profile it calls (PROFILE_VOIGT -> pcqsdhc PART1 -> hum1_wei, misc/hapi.py:10131-10140, 9910-9915,
// 9833-9844) for many homogeneous (T, p) states at once.
//
// Design (B200-first, not a translation):
//   * one CTA (256 threads) per (wavenumber tile, state): 2048-point tiles at 80 registers and 72 KB of shared memory
//     (three CTAs per SM) when the windows are thousands of grid points wide, 1024-point tiles at 64 registers
//     (four CTAs per SM) when they are a few hundred (each state by its own wing, on the device).  The tile of the spectral grid is
//     staged in shared memory by one TMA bulk copy (cp.async.bulk + mbarrier) and kept there: the prologue searches
//     it with exact comparisons, so window membership is decided on the same doubles the reference bisects, and the
//     evaluation passes read it back.
//   * the CTA streams only the lines that can reach its tile (device-side binning on the sorted line centres
//     with a conservative per-state wing bound), 512 (256) lines per chunk, dealt round-robin so that every chunk
//     spans the whole candidate range.  One thread derives one line's state (S(T), Gamma_D, Gamma_0, shift, wing,
//     window indices, Weideman core indices) into ONE 80-byte record; records stay at fixed slots and
//     per-consumer-warp ballot masks say which records a warp must visit (no atomics, fixed summation order).
//   * pass 1 ("wide" lines): each warp owns 32*K contiguous grid points (K per lane, lane-strided), accumulators
//     in registers.  Outside the Weideman core the reference's w(z) is the one-term Humlicek asymptote, for which
//       S*V = a*(u+h)/(u^2+c),  u=(omega-nu0')^2+Gamma0^2-h/2, h=GammaD^2/ln2, c=2*h*Gamma0^2, a=amp*Gamma0/pi
//     (closed form of SURVEY.md App. A.3 with the square completed) = 8 FP64 pipe operations per (line, grid
//     point) incl. a Newton-refined MUFU.RCP64H reciprocal; on slices a window covers completely two lines share
//     one reciprocal (7.5 operations per evaluation) -- including the slice that holds the line's core, where pass 2
//     adds (series - closed form) on the core points.  Slices holding a window end use edge-specialised code.
//   * pass 2: the Weideman cores of wide lines and the whole window of narrow lines (<= 128 points) are evaluated
//     lane-per-point by the warp that owns that part of the tile (shared-memory accumulator, no atomics): wings
//     record by record, two rows at a time; cores flattened across record boundaries, two steps at a time, so the
//     24-term rational series never diverges against the cheap branch inside a warp.
//   * far field: on a slice (32*K points) or a whole tile that a window covers from >= 0.375 block widths away the
//     closed form is summed at the block's 32 Chebyshev nodes only and interpolated once per (tile, state) -- K times
//     (slice) or 8 K times (tile) fewer evaluations; the tile level lives in its own instantiation (TFP).  The far wing of a
//     speed-dependent line (difference of two asymptotes) goes the same way.
//   * which instantiation a state runs in (K = 8, K = 8 with tile-level blocks, K = 4) is decided on the device from
//     the state's own wing: the grids of all candidates are launched and the CTAs of the others return at once.  Small
//     launches split the candidate lines of a tile over several CTAs (partial spectra added in part order).
//   * flags select the Lorentz and Doppler drivers of hapi (same skeleton, different wing rule and profile).
//   * FP64 throughout (an optional FP32 mode exists for the asymptote branch); summation order is fixed, so
//     results are bit-reproducible run to run and, with the tile size pinned, independent of how the grid is
//     sharded across GPUs.
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "rtx_common.cuh"

namespace {

constexpr int NT = 256;         // threads per CTA
// 3 CTAs (24 warps) per SM: 80 registers per thread and 72 KB of shared memory per CTA; measured 7 % faster on
// config 2 than 2 CTAs with 126 registers and 512-line chunks.
#ifndef RTX_ROUNDS
#define RTX_ROUNDS 1
#endif
#ifndef RTX_P2_ILP
#define RTX_P2_ILP 2
#endif
// K = 4 (1024-point tiles, 36 KB of shared memory) is compiled for four CTAs per SM (64 registers): the regime it is
// chosen for -- windows of a few hundred grid points -- is latency- and issue-bound, not register-bound.
#ifndef RTX_MINBLOCKS_K4
#define RTX_MINBLOCKS_K4 4
#endif
// eval-weighted mean window half-width (in grid points) below which K = 4 beats K = 8 (measured: coarse-grid batch
// of config 3, ~350 points: K = 4 is 9 % faster; config 2, ~3000 points: K = 8 is 9 % faster)
#ifndef RTX_K4_WINDOW
#define RTX_K4_WINDOW 1200.0
#endif
#ifndef RTX_MINBLOCKS
#define RTX_MINBLOCKS 3
#endif
// K <= 2 (small problems): the CTAs are latency-bound, more of them per SM
#ifndef RTX_MINBLOCKS_SMALL
#define RTX_MINBLOCKS_SMALL 4
#endif

// Lines per chunk (one line per thread per prologue round).  The FP64 kernels keep ONE 80-byte record per line, so
// the 2048-point tile affords 512-line chunks at three CTAs per SM: half as many block-wide barriers and prologue
// phases per tile as 256-line chunks (measured gain 0.3 %; the single record itself 0.5 %).  The FP32 mode carries
// an extra single-precision record per line and the warp-specialised variant double-buffers, so they stay at 256.
#ifndef RTX_CORE_SUB
#define RTX_CORE_SUB 1
#endif
#ifndef RTX_LCH_BIG
#define RTX_LCH_BIG 512
#endif
template <int K, int MODE, bool WS>
struct ChunkCfg {
  static constexpr int LCH = (K == 8 && MODE != 2 && !WS) ? RTX_LCH_BIG : NT * RTX_ROUNDS;
};
// tile-level far-field blocks compiled in (a whole tile at 32 nodes when a window covers it from far away)
#ifndef RTX_TILEFAR
#define RTX_TILEFAR 1
#endif
// typical window half-width (grid points) from which the variant with tile-level blocks is faster (measured break-even on
// uniform wings: ~3700 points; tools/tilefar_sweep.py)
#ifndef RTX_TF_WINDOW
#define RTX_TF_WINDOW 4000.0
#endif
// ... and the estimated number of lines per tile that qualify (2 * line density * (typical wing - 1.375 tile widths)) from
// which it pays: with a few hundred lines per tile the CTA is bound by the latency of its prologue, not by its visits
// (measured: 1150 qualifying lines per tile 11 % faster, 220 lines 5 % slower; tools/xs_breakdown.py)
#ifndef RTX_TF_MINLINES
#define RTX_TF_MINLINES 600.0
#endif
// line split of small problems: at most this many CTAs per (tile, state), aiming at this many CTAs per SM in total
#ifndef RTX_SPLIT_MAX
#define RTX_SPLIT_MAX 8
#endif
#ifndef RTX_SPLIT_CTAS_PER_SM
#define RTX_SPLIT_CTAS_PER_SM 3
#endif
#ifndef RTX_NARROW_MAX
#define RTX_NARROW_MAX 128
#endif
constexpr int NARROW_MAX = RTX_NARROW_MAX;  // windows of at most this many tile points go to pass 2 entirely

__constant__ double c_weideman[24] = {
    // Weideman N=24 coefficients, highest power first.  tests/test_host_logic.py regenerates them from the
    // FFT recipe (misc/hapi.py:9816-9824) and compares with the values exported by rtx_weideman_coeffs().
    -1.5137469056014652e-10, 4.904820416137823e-09,  1.3310453109352295e-09, -3.0082823610602605e-08,
    -1.9122258957660688e-08, 1.87383434459447e-07,   2.5682641329013596e-07, -1.085647579377311e-06,
    -3.038893184278043e-06,  4.1394617242916025e-06, 3.0471066082970273e-05, 2.4331415462208058e-05,
    -0.00020748431511423976, -0.0007816642995623942, -0.000493642690128576,  0.006215006362949158,
    0.03372336685531593,     0.10838723484566715,    0.26549639598807695,    0.5361139535729114,
    0.9257087138588673,      1.394819673379119,      1.85628649920554,       2.1978589365315417};
const double h_weideman[24] = {
    -1.5137469056014652e-10, 4.904820416137823e-09,  1.3310453109352295e-09, -3.0082823610602605e-08,
    -1.9122258957660688e-08, 1.87383434459447e-07,   2.5682641329013596e-07, -1.085647579377311e-06,
    -3.038893184278043e-06,  4.1394617242916025e-06, 3.0471066082970273e-05, 2.4331415462208058e-05,
    -0.00020748431511423976, -0.0007816642995623942, -0.000493642690128576,  0.006215006362949158,
    0.03372336685531593,     0.10838723484566715,    0.26549639598807695,    0.5361139535729114,
    0.9257087138588673,      1.394819673379119,      1.85628649920554,       2.1978589365315417};

struct VoigtParams {
  const double *nu, *s0, *elower, *gamma_air, *gamma_self, *n_air, *n_self, *delta_air, *deltap_air,
      *delta_self, *deltap_self;
  const int32_t* iso;
  int64_t n_lines;
  const double *T, *p, *weight;  // [n_states], [n_states], [n_states][n_iso]
  const double *st_q, *st_d;     // [n_states][n_iso]: Q(Tref)/Q(T), Doppler factor
  int n_iso, n_states;
  double x_air, x_self, wing_abs, wing_hw, s_thr;
  const double* grid;
  int64_t nx;
  int n_tiles;
  const int32_t* bins;  // [n_states][n_tiles][2]
  double* out;
  int has_self, has_deltap, use_tma;
  // speed dependence (absorptionCoefficient_SDVoigt) and extra broadening species
  const double *sd_air, *sd_self;
  int n_extra, sdv;
  const double *xg[RTX_MAX_EXTRA_DILUENTS], *xn[RTX_MAX_EXTRA_DILUENTS], *xd[RTX_MAX_EXTRA_DILUENTS],
      *xdp[RTX_MAX_EXTRA_DILUENTS], *xsd[RTX_MAX_EXTRA_DILUENTS];
  double xx[RTX_MAX_EXTRA_DILUENTS];
  const double* cheb;   // [32] Chebyshev nodes cos(theta_k), then [32][32] node-values -> coefficients matrix [k][m]
  int far_ok;    // far-field blocks enabled
  int tfar_ok;   // tile-level far-field blocks enabled (whole tiles at 32 nodes)
  // device-side choice of the kernel variant, per state (null = run): sel[2] = 1 / mean grid spacing; st_wb / st_wt =
  // the state's wing bound / typical wing.  Variant 0 (K = 4): wing bound < RTX_K4_WINDOW grid points unless sel_k8;
  // variant 2 (K = 8 with the tile-level far-field blocks): typical wing >= RTX_TF_WINDOW points and sel_tf; variant 1
  // (K = 8 without) otherwise.  The grids of all candidate variants are launched; a CTA of a variant its state does
  // not pick returns at once.  sel[0], sel[1]: the launch-wide eval-weighted means (reporting only).
  const double* sel;
  const double *st_wb, *st_wt;
  int sel_k8, sel_tf;
  // small problems: the candidate lines of a (tile, state) are split over `split` CTAs that write partial sums to
  // part_out[part][state][nx]; sum_parts_kernel adds them in part order (fixed summation order)
  int split;
  double* part_out;
  double dens;   // lines per cm^-1 of the table (the tile-level blocks pay only when many lines reach over whole tiles)
  int lorentz;   // PROFILE_LORENTZ instead of PROFILE_VOIGT: no Doppler width, wing = max(OmegaWing, HW*Gamma0)
  int doppler;   // PROFILE_DOPPLER (absorptionCoefficient_Doppler): 1 = with line shift, 2 = LineShift=False
};

// Everything the reference derives for one line at one (T, p): misc/hapi.py:11058-11131.
struct LineState {
  double nu, sg0, W, gamD, gam0, gam2, amp;
  bool keep;
};

// EXT = false compiles the speed-dependence / extra-species terms out (the common case: air and self only)
template <bool EXT>
__device__ __forceinline__ LineState line_state(const VoigtParams& P, int64_t li, int s, double T, double p,
                                                double lnr) {
  LineState L;
  const double nu = P.nu[li];
  const int iso = P.iso[li];
  L.keep = true;
  L.amp = 0.0;                                                   // set by line_strength() once the line is known to reach
  L.gamD = P.st_d[(int64_t)s * P.n_iso + iso] * nu;              // :11087 (unshifted centre)
  // Gamma0 and Shift0 over the diluents, :11090-11128
  double g0 = 0.0, sh = 0.0;
  {
    const double rp = exp(P.n_air[li] * lnr);                    // (Tref/T)**n_air
    g0 += P.x_air * (P.gamma_air[li] * p * rp);
    double d = P.delta_air[li];
    if (P.has_deltap) d += P.deltap_air[li] * (T - RTX_TREF);
    sh += P.x_air * (d * p);
  }
  if (P.has_self) {
    const double rp = exp(P.n_self[li] * lnr);
    g0 += P.x_self * (P.gamma_self[li] * p * rp);
    double d = P.delta_self[li];
    if (P.has_deltap) d += P.deltap_self[li] * (T - RTX_TREF);
    sh += P.x_self * (d * p);
  }
  double g2 = 0.0;                                               // Gamma2 of the SDVoigt driver, misc/hapi.py:10883-10890
  if (EXT && P.sdv) {
    if (P.sd_air) g2 += P.x_air * (P.sd_air[li] * p) * P.gamma_air[li];
    if (P.has_self && P.sd_self) g2 += P.x_self * (P.sd_self[li] * p) * P.gamma_self[li];
  }
  for (int e = 0; EXT && e < P.n_extra; ++e) {                   // any other species of Diluent={...}, :11090-11128
    if (P.xx[e] == 0.0 || !P.xg[e]) continue;
    const double gam = P.xg[e][li];
    const double rp = exp((P.xn[e] ? P.xn[e][li] : P.n_air[li]) * lnr);
    g0 += P.xx[e] * (gam * p * rp);
    double d = P.xd[e] ? P.xd[e][li] : 0.0;
    if (P.xdp[e]) d += P.xdp[e][li] * (T - RTX_TREF);
    sh += P.xx[e] * (d * p);
    if (P.sdv && P.xsd[e]) g2 += P.xx[e] * (P.xsd[e][li] * p) * gam;
  }
  L.gam2 = g2;
  if (P.doppler) {                                               // misc/hapi.py:11510-11546 (Doppler driver)
    g0 = 0.0;
    sh = (P.doppler == 1) ? P.delta_air[li] * p : 0.0;           // Shift0DB*p/pref, air only, no T dependence
  }
  L.gam0 = g0;
  L.nu = nu;
  L.sg0 = nu + sh;
  L.W = P.doppler ? fmax(P.wing_abs, P.wing_hw * L.gamD)         // :11540
      : P.lorentz ? fmax(P.wing_abs, P.wing_hw * g0)                  // misc/hapi.py:11375 (Lorentz driver)
                  : fmax(fmax(P.wing_abs, P.wing_hw * g0), P.wing_hw * L.gamD);  // :11131
  return L;
}
// S(T) = sw*Qref/Q * ch/zn, EnvironmentDependency_Intensity (misc/hapi.py:10169-10175; zn is folded into s0 at ingest),
// times the caller's weight; and the reference's filters (:11066 component / weight, :11082 intensity threshold).
// Kept apart from line_state so that the two exponentials are spent only on lines whose window reaches the tile.
__device__ __forceinline__ void line_strength(const VoigtParams& P, int64_t li, int s, double T, LineState& L) {
  const int iso = P.iso[li];
  const double wgt = P.weight[(int64_t)s * P.n_iso + iso];
  const double ch = exp(-RTX_C2_HAPI * P.elower[li] / T) * (1.0 - exp(-RTX_C2_HAPI * L.nu / T));
  const double S = P.s0[li] * P.st_q[(int64_t)s * P.n_iso + iso] * ch;
  L.keep = !(S < P.s_thr) && (wgt != 0.0);
  L.amp = wgt * S;
}

// |x|+y < 15 exactly as hum1_wei tests it (misc/hapi.py:9841) on x = -((sg0-sg)*cte), y = Gam0*cte
// (misc/hapi.py:9911-9913); explicit _rn intrinsics keep ptxas from contracting into an FMA.
__device__ __forceinline__ bool in_weideman(double omega, double sg0, double cte, double y) {
  const double x = __dmul_rn(omega - sg0, cte);
  return __dadd_rn(fabs(x), y) < 15.0;
}

template <bool FAST>
__device__ __forceinline__ double rcp_nr(double x) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));   // MUFU.RCP64H
  const double e = fma(-x, r0, 1.0);
  if (FAST) return fma(r0, e, r0);
  const double t = fma(e, e, e);
  return fma(r0, t, r0);                                   // r0*(1+e+e^2): cubic convergence
}

// Humlicek one-term branch of hum1_wei folded with the Voigt prefactor (closed form):
//   S*V = a*(q+e)/(q^2+f*q+e^2),  q=(omega-nu0')^2, e=h/2+G0^2, f=2*G0^2-h, h=GD^2/ln2, a=amp*G0/pi.
// Completing the square, with u = q + f/2 = d^2 + (G0^2 - h/2):  S*V = a*(u+h)/(u^2+c),  c = e^2-f^2/4 = 2*h*G0^2
// -> 8 FP64-pipe operations: DADD, 3 DFMA, MUFU.RCP64H + 3 DFMA (cubic Newton), DFMA.  In the branch's domain
// (|x|+y >= 15) u is positive and free of cancellation; c is formed from its closed form, not by subtraction.
template <bool FAST>
__device__ __forceinline__ double cheap_d(double d, double a, double ah, double m, double c, double acc) {
  const double u = fma(d, d, m);
  const double den = fma(u, u, c);
  const double num = fma(a, u, ah);
  return fma(num, rcp_nr<FAST>(den), acc);
}
template <bool FAST>
__device__ __forceinline__ double cheap_term(double omega, double sg0, double a, double ah, double m, double c,
                                             double acc) {
  return cheap_d<FAST>(omega - sg0, a, ah, m, c, acc);
}

// Two lines through ONE reciprocal:  n1/d1 + n2/d2 = (n1*d2 + n2*d1)/(d1*d2)  -- 15 FP64 operations for two
// evaluations instead of 16, and half the MUFU traffic (micro-benchmark: +8 % on the test-free path).  All factors
// are positive and of moderate magnitude in the asymptote branch's domain, so nothing cancels or overflows.
template <bool FAST>
__device__ __forceinline__ double cheap_pair_d(double d1, double a1, double ah1, double m1, double c1, double d2,
                                               double a2, double ah2, double m2, double c2, double acc) {
  const double u1 = fma(d1, d1, m1), u2 = fma(d2, d2, m2);
  const double den1 = fma(u1, u1, c1), den2 = fma(u2, u2, c2);
  const double n1 = fma(a1, u1, ah1), n2 = fma(a2, u2, ah2);
  const double P = den1 * den2;
  const double X = fma(n2, den1, n1 * den2);
  return fma(X, rcp_nr<FAST>(P), acc);
}
template <bool FAST>
__device__ __forceinline__ double cheap_pair(double omega, double sg1, double a1, double ah1, double m1, double c1,
                                             double sg2, double a2, double ah2, double m2, double c2, double acc) {
  return cheap_pair_d<FAST>(omega - sg1, a1, ah1, m1, c1, omega - sg2, a2, ah2, m2, c2, acc);
}

// Re w(x+iy) by Weideman's 24-term rational series, misc/hapi.py:9812-9827 (cef):
//   Z = (L+iz)/(L-iz),  w = 2 p(Z)/(L-iz)^2 + (1/sqrt(pi))/(L-iz),  p = polyval(a, Z).
// p has real coefficients, so it is evaluated at the complex point Z with the real two-term recurrence
//   u_k = a_k + s u_{k+1} - t u_{k+2},  s = 2 Re Z, t = |Z|^2,   p(Z) = Z u_1 + a_0 - t u_2
// (2 FMAs per coefficient instead of the 4 of a complex Horner step; degree 23, so rounding-error growth is
// bounded by a few hundred ulp -- far inside the 1e-9 parity tolerance, and tested against the oracle).
__device__ __forceinline__ double weideman_re(double x, double y) {
  const double L = RTX_WEIDEMAN_L;
  const double dr = L + y, nr = L - y;          // L-iz = dr - i x ;  L+iz = nr + i x
  const double m2 = fma(dr, dr, x * x);
  const double inv = rcp_nr<false>(m2);
  const double Zr = fma(nr, dr, -(x * x)) * inv;
  const double Zi = (2.0 * L * x) * inv;
  const double s = 2.0 * Zr, t = -fma(Zr, Zr, Zi * Zi);   // t holds -|Z|^2
  double u2 = c_weideman[0];
  double u1 = fma(s, u2, c_weideman[1]);
#pragma unroll
  for (int k = 2; k < 23; ++k) {
    const double u0 = fma(s, u1, fma(t, u2, c_weideman[k]));
    u2 = u1;
    u1 = u0;
  }
  const double pr = fma(Zr, u1, fma(t, u2, c_weideman[23]));
  const double pi = Zi * u1;
  const double ir = dr * inv, ii = x * inv;     // 1/(L-iz)
  const double qr = fma(ir, ir, -(ii * ii)), qi = 2.0 * ir * ii;
  return fma(2.0, fma(pr, qr, -(pi * qi)), RTX_INV_SQRT_PI * ir);
}

// ---- complex w(z) and the quadratic speed-dependent profile (absorptionCoefficient_SDVoigt) -------------------------
// w(x+iy) as hum1_wei returns it (misc/hapi.py:9833-9844): the one-term asymptote t/(0.5+t^2)/sqrt(pi), t = y - ix,
// or, where |x|+y < 15, Weideman's 24-term series -- real and imaginary part.
__device__ __forceinline__ void w_hum1_wei(double x, double y, double& wr, double& wi) {
  if (fabs(x) + y < 15.0) {
    const double L = RTX_WEIDEMAN_L;
    const double dr = L + y, nr = L - y;
    const double m2 = fma(dr, dr, x * x);
    const double inv = 1.0 / m2;
    const double Zr = fma(nr, dr, -(x * x)) * inv, Zi = (2.0 * L * x) * inv;
    const double s = 2.0 * Zr, t = -fma(Zr, Zr, Zi * Zi);
    double u2 = c_weideman[0];
    double u1 = fma(s, u2, c_weideman[1]);
#pragma unroll
    for (int k = 2; k < 23; ++k) {
      const double u0 = fma(s, u1, fma(t, u2, c_weideman[k]));
      u2 = u1;
      u1 = u0;
    }
    const double pr = fma(Zr, u1, fma(t, u2, c_weideman[23])), pi = Zi * u1;
    const double ir = dr * inv, ii = x * inv;                     // 1/(L-iz)
    const double qr = fma(ir, ir, -(ii * ii)), qi = 2.0 * ir * ii;
    wr = fma(2.0, fma(pr, qr, -(pi * qi)), RTX_INV_SQRT_PI * ir);
    wi = fma(2.0, fma(pr, qi, pi * qr), RTX_INV_SQRT_PI * ii);
  } else {
    const double dr = 0.5 + y * y - x * x, di = -2.0 * x * y;     // 0.5 + t^2
    const double inv = RTX_INV_SQRT_PI / fma(dr, dr, di * di);
    wr = (y * dr - x * di) * inv;                                 // t * conj(den) / |den|^2
    wi = (-x * dr - y * di) * inv;
  }
}
// cpf3 (misc/hapi.py:9645-9670): 15-term asymptotic series, real part and imaginary part
__device__ __forceinline__ void w_cpf3(double x, double y, double& wr, double& wi) {
  const double inv = 1.0 / fma(x, x, y * y);
  const double ar = x * inv, ai = -y * inv;                       // zm1 = 1/(x+iy)
  const double br = ar * ar - ai * ai, bi = 2.0 * ar * ai;        // zm2
  double tr = 1.0, ti = 0.0, sr = 1.0, si = 0.0;
#pragma unroll
  for (int k = 0; k < 15; ++k) {
    const double f = 0.5 + (double)k;
    const double nr = (tr * br - ti * bi) * f, ni = (tr * bi + ti * br) * f;
    tr = nr; ti = ni;
    sr += tr; si += ti;
  }
  // zsum *= i*zm1*pipwoeronehalf
  const double c = 0.564189583547756;
  const double mr = -ai * c, mi = ar * c;                         // i*zm1*c
  wr = sr * mr - si * mi;
  wi = sr * mi + si * mr;
}
__device__ __forceinline__ void csqrt_dev(double a, double b, double& re, double& im) {
  const double r = hypot(a, b);
  if (r == 0.0) { re = 0.0; im = 0.0; return; }
  if (a >= 0.0) { re = sqrt(0.5 * (r + a)); im = b / (2.0 * re); }
  else { const double t = sqrt(0.5 * (r - a)); im = b >= 0.0 ? t : -t; re = b / (2.0 * im); }
}
// Real part of the pCqSDHC line shape (misc/hapi.py:9846-10023) for eta = anuVC = Shift2 = 0 and Gam2 != 0 -- what
// PROFILE_SDVOIGT asks of it from the SDVoigt driver (misc/hapi.py:10116-10128, 10899).  d = sg - (sg0 + Shift0),
// c0tr = Gam0 - 1.5 Gam2.  With these arguments the common part reduces to LS = Aterm / pi (:10022).
__device__ __noinline__ double sdvoigt_re(double d, double c0tr, double gam2, double cte) {
  const double xr = c0tr / gam2, xi = -d / gam2;                  // X = (i(sg0 - sg) + c0t)/c2t
  const double h = 1.0 / (2.0 * cte * gam2);                      // csqrtY (real: Shift2 = 0);  Y = h^2
  const double Y = h * h;
  const double ax = hypot(xr, xi);
  if (ax <= 3.0e-8 * Y) {                                         // PART2
    double sr, si, w1r, w1i, w2r, w2i;
    csqrt_dev(xr + Y, xi, sr, si);
    w_hum1_wei(d * cte, c0tr * cte, w1r, w1i);                    // Z1 = (i(sg0 - sg) + c0t)*cte
    w_hum1_wei(-si, sr + h, w2r, w2i);                            // Z2 = sqrt(X+Y) + csqrtY
    return (cte * RTX_INV_SQRT_PI) * (w1r - w2r);
  }
  if (Y <= 1.0e-15 * ax) {                                        // PART3
    double qr, qi;
    csqrt_dev(xr, xi, qr, qi);                                    // sqrt(X)
    double ar, ai;                                                // Aterm * c2t
    if (hypot(qr, qi) <= 4.0e3) {
      double wr, wi;
      w_hum1_wei(-qi, qr, wr, wi);
      ar = 2.0 * RTX_SQRT_PI * (RTX_INV_SQRT_PI - (qr * wr - qi * wi));
      ai = 2.0 * RTX_SQRT_PI * (-(qr * wi + qi * wr));
    } else {
      const double n2 = fma(xr, xr, xi * xi);
      const double ir = xr / n2, ii = -xi / n2;                   // 1/X
      ar = ir - 1.5 * (ir * ir - ii * ii);
      ai = ii - 1.5 * (2.0 * ir * ii);
    }
    (void)ai;
    return ar / gam2 / RTX_PI;
  }
  double sr, si;                                                  // PART4
  csqrt_dev(xr + Y, xi, sr, si);
  const double x1 = -si, y1 = sr - h, y2 = sr + h;                // Z1 = sqrt(X+Y) - csqrtY, Z2 = Z1 + 2 csqrtY
  const double S1 = sqrt(x1 * x1 + y1 * y1), S2 = sqrt(x1 * x1 + y2 * y2);
  double w1r, w1i, w2r, w2i;
  if (fabs(S1 - S2) <= 1.0 && fmax(S1, S2) > 8.0 && fmin(S1, S2) <= 8.0) {
    w_cpf3(x1, y1, w1r, w1i);
    w_cpf3(x1, y2, w2r, w2i);
  } else {
    w_hum1_wei(x1, y1, w1r, w1i);
    w_hum1_wei(x1, y2, w2r, w2i);
  }
  return (cte * RTX_INV_SQRT_PI) * (w1r - w2r);
}

// Far wing of the speed-dependent profile.  Where pcqsdhc takes PART4 (neither |X| <= 3e-8 |Y| nor |Y| <= 1e-15 |X|,
// misc/hapi.py:9923, 9968) and both w(Z) come from the one-term asymptote of hum1_wei (|x|+y >= 15 for Z1 and Z2, which
// also rules cpf3 out: |Z| >= 15/sqrt(2) > 8), the profile is  cte/sqrt(pi) * Re[w1(Z1) - w1(Z2)],  w1 = t/(0.5+t^2)/sqrt(pi),
// an analytic function of omega whose singularities sit within a few widths of the line centre -- so it can be summed at
// the Chebyshev nodes of a far slice or tile like the Voigt closed form.  |X|, |Im sqrt(X+Y)| and Re sqrt(X+Y) grow with
// |omega - nu0'|, so the branch conditions hold on a whole block when they hold at its nearest point (asymptotes, not
// PART2) and at its farthest point (not PART3).
// rec: a = Re c0t = Gam0 - 1.5 Gam2, g2 = Gam2, cte;  precomputed ig2 = 1/Gam2, h = 1/(2 cte Gam2).
__device__ __forceinline__ bool sdv_far_ok(double d_near, double d_far, double c0tr, double gam2, double cte) {
  const double xr = c0tr / gam2, h = 1.0 / (2.0 * cte * gam2), Y = h * h;
  const double xin = -d_near / gam2, xif = -d_far / gam2;
  const double axn = hypot(xr, xin), axf = hypot(xr, xif);
  if (axn <= 3.0e-8 * Y || Y <= 1.0e-15 * axf || Y <= 1.0e-15 * axn || !(h > 0.0)) return false;
  double sr, si;
  csqrt_dev(xr + Y, xin, sr, si);
  return fabs(si) + (sr - h) >= 15.0 + 1e-9 && (sr - h) > 0.0;   // Z1 = (-si) + i (sr - h); Z2 has the larger y
}
// value at distance d from the (shifted) centre, without the amplitude
__device__ __forceinline__ double sdv_far_term(double d, double c0tr, double gam2, double cte) {
  const double xr = c0tr / gam2, xi = -d / gam2, h = 1.0 / (2.0 * cte * gam2);
  double sr, si;
  csqrt_dev(xr + h * h, xi, sr, si);
  const double x = -si, x2 = x * x, y1 = sr - h, y2 = sr + h;
  // Re w1(x + i y) = y (0.5 + x^2 + y^2) / ((0.5 + y^2 - x^2)^2 + 4 x^2 y^2) / sqrt(pi)
  const double e1 = 0.5 + y1 * y1 - x2, e2 = 0.5 + y2 * y2 - x2;
  const double w1 = y1 * (0.5 + x2 + y1 * y1) / fma(e1, e1, 4.0 * x2 * y1 * y1);
  const double w2 = y2 * (0.5 + x2 + y2 * y2) / fma(e2, e2, 4.0 * x2 * y2 * y2);
  return (cte * (RTX_INV_SQRT_PI * RTX_INV_SQRT_PI)) * (w1 - w2);
}

struct __align__(16) WideRec {
  double sg0, a, ah, m, c;
  short lo, hi, clo, chi;
};
// FP32 mode reuses the WideRec slots with this layout: the line centre relative to the tile's first grid point as
// a float pair (hi + lo), so that omega - nu0' is formed in single precision without cancellation error.
struct __align__(16) WideRec32 {
  float chi, clo, af, hf, mf, cf, pad0, pad1;
  short lo, hi, core_lo, core_hi;
};
static_assert(sizeof(WideRec32) <= sizeof(WideRec), "WideRec32 must fit a WideRec slot");
// Far-field blocks.  Away from its core a line's contribution is the closed form a*(u+h)/(u^2+c): analytic in omega with
// poles only at |omega-nu0'| ~ Gamma0, GammaD.  On a warp's slice of 32*K grid points whose nearest point is at least
// FAR_FRAC slice widths away from the line centre, the degree-31 Chebyshev interpolant through 32 nodes reproduces it
// to < 1e-13 relative (scale-free in the distance/width ratio: 0.25 -> 7e-12, 0.3 -> 4e-13, 0.375 -> 2e-14, measured
// for Gamma0 from 1e-4 to 100 grid spacings; tools/farfield_error.py).  So a line that covers such a slice is evaluated
// at the slice's 32 nodes only -- one node per lane, K times fewer evaluations -- into per-lane node accumulators, and
// once per (tile, state) the node sums are turned into Chebyshev coefficients and evaluated at the slice's own grid
// points (Clenshaw at the actual grid values, so nothing assumes a uniform grid).  Window membership stays exact:
// a slice is far only if the reference's window covers all of it (misc/hapi.py:11133-11134).
#ifndef RTX_FAR_FRAC
#define RTX_FAR_FRAC 0.375
#endif
// per wide line: bit q = the window touches warp q's slice, bit 8+q = it covers the slice with no Weideman point
// inside (test-free path), bit 16+q = it covers the slice from far away (far-field nodes)
// With core_ok a covered slice is test-free even when the core lies inside it: pass 1 then evaluates the closed form
// on the core points too and pass 2 adds (series - closed form) there.
// s_far[q] / s_far[8 + q]: a line centre at or below / at or above these values is "far" from slice q (set per tile).
__device__ __forceinline__ unsigned range_mask(int a, int b) {        // bits a..b inclusive, empty when a > b
  return a > b ? 0u : (((2u << b) - 1u) & ~((1u << a) - 1u));
}
template <int SLICE>
__device__ __forceinline__ unsigned slice_masks(int lo, int hi, int clo, int chi, bool core_ok, bool far_ok, double sg0,
                                                const double* __restrict__ s_far) {
  constexpr int NQ = NT / 32;
  const unsigned touch = range_mask(lo / SLICE, (hi - 1) / SLICE);                      // hi > lo
  const unsigned cover = range_mask((lo + SLICE - 1) / SLICE, hi / SLICE - 1) & ((1u << NQ) - 1u);
  const unsigned core = chi > clo ? range_mask(clo / SLICE, (chi - 1) / SLICE) : 0u;
  unsigned far = 0;
  if (far_ok && (cover & ~core)) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) far |= ((sg0 <= s_far[q] || sg0 >= s_far[NQ + q]) ? 1u : 0u) << q;
    far &= cover & ~core;
  }
  const unsigned full = cover & ~far & (core_ok ? ~0u : ~core);
  return (touch & ((1u << NQ) - 1u)) | (full << 8) | (far << 16);
}
// per-line record: [s, e) is evaluated by pass 2, of which [clo, chi) is the Weideman core (empty: clo == chi)
// It is also the pass-1 record of the FP64 kernels: [lo, hi) is the line's window, [clo, chi) its Weideman core.
struct __align__(16) SegRec {
  double sg0, a, ah, m, c, cte, y, cw;
  short lo, hi, clo, chi, s, e;
  int sub;   // slices the window covers were evaluated with the closed form on the core points too (pass 2 subtracts)
};
static_assert(sizeof(SegRec) == 80, "one 80-byte record per line");

// ---------------------------------------------------------------------------------------------
// per-state tables: Q(Tref)/Q(T) by TIPS-2011 Lagrange interpolation (misc/hapi.py:5311-5388, 9568-9582),
// Doppler factor sqrt(2kT ln2/(m c^2)) (misc/hapi.py:11085-11087), conservative wing bound.
__device__ __forceinline__ double tips_lagrange(double aa, const double* __restrict__ A,
                                                const double* __restrict__ B, int npt) {
  // I = the first c in [2, npt] with A[c-1] >= aa (the reference scans linearly; the nodes ascend, so bisect)
  int lo = 2, hi = npt + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (A[mid - 1] >= aa) hi = mid; else lo = mid + 1;
  }
  if (lo > npt) return nan("");
  const int I = lo;
  if (I < 3 || I == npt) {
    const int J = (I < 3 ? 3 : npt) - 1;
    const double a0 = A[J - 2], a1 = A[J - 1], a2 = A[J];
    const double w0 = (aa - a1) * (aa - a2) / ((a0 - a1) * (a0 - a2));
    const double w1 = (aa - a0) * (aa - a2) / ((a1 - a0) * (a1 - a2));
    const double w2 = (aa - a0) * (aa - a1) / ((a2 - a0) * (a2 - a1));
    return w0 * B[J - 2] + w1 * B[J - 1] + w2 * B[J];
  }
  const int J = I - 1;
  const double a0 = A[J - 2], a1 = A[J - 1], a2 = A[J], a3 = A[J + 1];
  const double w0 = (aa - a1) * (aa - a2) * (aa - a3) / ((a0 - a1) * (a0 - a2) * (a0 - a3));
  const double w1 = (aa - a0) * (aa - a2) * (aa - a3) / ((a1 - a0) * (a1 - a2) * (a1 - a3));
  const double w2 = (aa - a0) * (aa - a1) * (aa - a3) / ((a2 - a0) * (a2 - a1) * (a2 - a3));
  const double w3 = (aa - a0) * (aa - a1) * (aa - a2) / ((a3 - a0) * (a3 - a1) * (a3 - a2));
  return w0 * B[J - 2] + w1 * B[J - 1] + w2 * B[J] + w3 * B[J + 1];
}

struct PrepParams {
  int n_states, n_iso;
  const double *T, *p, *mass, *tipsT, *q, *qref;
  double *st_q, *st_d, *st_wb;
  double x_air, x_self, wing_abs, wing_hw;
  double nu_max, gair_max, gself_max, nair_min, nair_max, nself_min, nself_max;
  double gair_typ, gself_typ, nair_typ, nself_typ;
  double* st_wt;          // [n_states] typical wing (eval-weighted mean width instead of the largest)
  int doppler, n_extra;
  double xx[RTX_MAX_EXTRA_DILUENTS], xg_max[RTX_MAX_EXTRA_DILUENTS], xn_min[RTX_MAX_EXTRA_DILUENTS], xn_max[RTX_MAX_EXTRA_DILUENTS];
};

// one warp per state: the lanes share the isotopologues (a single-state call is bound by this kernel's latency)
__global__ void state_prep_kernel(PrepParams P) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= P.n_states) return;                    // whole warps leave together
  const double T = P.T[s], p = P.p[s];
  const bool ok = (T >= 70.0) && (T <= 3000.0);   // misc/hapi.py:9571
  double dmax = 0.0;
  for (int j = lane; j < P.n_iso; j += 32) {
    const double Q = ok ? tips_lagrange(T, P.tipsT, P.q + (int64_t)j * RTX_TIPS_NPT, RTX_TIPS_NPT) : nan("");
    P.st_q[(int64_t)s * P.n_iso + j] = P.qref[j] / Q;
    const double m = P.mass[j] * RTX_CMASS * 1000.0;
    // the Doppler driver has its own constants (SI, cSqrt2Ln2 = 1.1774100225): misc/hapi.py:11533-11537
    const double d = P.doppler ? (1.1774100225 / 2.99792458e8) * sqrt(1.3806503e-23 / RTX_CMASS) * sqrt(T) / sqrt(P.mass[j])
                               : sqrt(2.0 * RTX_CBOLTZ * T * RTX_LN2 / m / (RTX_CLIGHT * RTX_CLIGHT));
    P.st_d[(int64_t)s * P.n_iso + j] = d;
    dmax = fmax(dmax, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane != 0) return;
  const double lnr = log(RTX_TREF / T);
  const double ra = fmax(exp(P.nair_min * lnr), exp(P.nair_max * lnr));
  const double rs = fmax(exp(P.nself_min * lnr), exp(P.nself_max * lnr));
  double g0 = fabs(p) * (fabs(P.x_air) * P.gair_max * ra + fabs(P.x_self) * P.gself_max * rs);
  for (int e = 0; e < P.n_extra; ++e)
    g0 += fabs(p) * fabs(P.xx[e]) * P.xg_max[e] * fmax(exp(P.xn_min[e] * lnr), exp(P.xn_max[e] * lnr));
  double wb = fmax(fmax(P.wing_abs, P.wing_hw * g0), P.wing_hw * dmax * P.nu_max);
  P.st_wb[s] = wb * (1.0 + 1e-6) + 1e-9;
  double gt = fabs(p) * (fabs(P.x_air) * P.gair_typ * exp(P.nair_typ * lnr) + fabs(P.x_self) * P.gself_typ * exp(P.nself_typ * lnr));
  for (int e = 0; e < P.n_extra; ++e)
    gt += fabs(p) * fabs(P.xx[e]) * P.xg_max[e] * fmax(exp(P.xn_min[e] * lnr), exp(P.xn_max[e] * lnr));
  if (P.doppler) gt = 0.0;
  P.st_wt[s] = fmin(wb, fmax(fmax(P.wing_abs, P.wing_hw * gt), P.wing_hw * dmax * 0.5 * P.nu_max));
}

// eval-weighted mean of the per-state wing in grid points, sum(w^2)/sum(w) / mean grid spacing: out[0] from the wing
// bound (largest width of the table), out[1] from the typical wing
__global__ void window_stats_kernel(int n_states, const double* __restrict__ st_wb, const double* __restrict__ st_wt,
                                    const double* __restrict__ grid, int64_t nx, double* __restrict__ out) {
  __shared__ double s1[256], s2[256], s3[256], s4[256];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int s = threadIdx.x; s < n_states; s += 256) {
    const double w = st_wb[s], t = st_wt[s];
    a += w * w; b += w; c += t * t; d += t;
  }
  s1[threadIdx.x] = a; s2[threadIdx.x] = b; s3[threadIdx.x] = c; s4[threadIdx.x] = d;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      s1[threadIdx.x] += s1[threadIdx.x + k]; s2[threadIdx.x] += s2[threadIdx.x + k];
      s3[threadIdx.x] += s3[threadIdx.x + k]; s4[threadIdx.x] += s4[threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double dv = nx > 1 ? (grid[nx - 1] - grid[0]) / (double)(nx - 1) : 0.0;
    out[0] = (s2[0] > 0.0 && dv > 0.0) ? (s1[0] / s2[0]) / dv : 1e300;
    out[1] = (s4[0] > 0.0 && dv > 0.0) ? (s3[0] / s4[0]) / dv : 1e300;
    out[2] = dv > 0.0 ? 1.0 / dv : 1e300;
  }
}

__global__ void tips_kernel(int n, int n_iso, const double* __restrict__ T, const double* __restrict__ tipsT,
                            const double* __restrict__ q, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * n_iso) return;
  const int s = (int)(i / n_iso), j = (int)(i % n_iso);
  const double t = T[s];
  out[i] = (t >= 70.0 && t <= 3000.0) ? tips_lagrange(t, tipsT, q + (int64_t)j * RTX_TIPS_NPT, RTX_TIPS_NPT)
                                      : nan("");
}

// Warp-cooperative bisection on the sorted line centres: every round the 32 lanes probe 32 positions of the current
// interval (3 rounds for 30 000 lines, 5 for 30 million) -- a thread-per-search bisection is a chain of ~20 dependent
// global loads, which a single-state call waits for.  RIGHT: number of elements <= v, else number of elements < v.
template <bool RIGHT>
__device__ __forceinline__ int64_t warp_bisect(const double* __restrict__ a, int64_t n, double v, int lane) {
  int64_t lo = 0, hi = n;                                    // the answer lies in [lo, hi]
  while (hi - lo > 32) {
    const int64_t len = hi - lo;
    const int64_t pk = lo + ((int64_t)(lane + 1) * len) / 33;                // lo < pk < hi, ascending with the lane
    const double x = a[pk];
    const unsigned m = __ballot_sync(0xffffffffu, RIGHT ? (x <= v) : (x < v));   // 1...10...0: the array is sorted
    const int cnt = __popc(m);
    const int64_t p_last = lo + ((int64_t)cnt * len) / 33, p_next = lo + ((int64_t)(cnt + 1) * len) / 33;
    if (cnt > 0) lo = p_last + 1;
    if (cnt < 32) hi = p_next;
  }
  const int64_t i = lo + lane;
  const bool in = i < hi && (RIGHT ? (a[i] <= v) : (a[i] < v));
  return lo + __popc(__ballot_sync(0xffffffffu, in));
}

// candidate line range of every (state, tile): lines with nu in [tile_lo - Wb, tile_hi + Wb]; one warp per pair
__global__ void bin_kernel(int n_states, int n_tiles, int tile_pts, const double* __restrict__ grid, int64_t nx,
                           const double* __restrict__ nu, int64_t n_lines, const double* __restrict__ st_wb,
                           int32_t* __restrict__ bins) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= (int64_t)n_states * n_tiles) return;
  const int s = (int)(i / n_tiles), t = (int)(i % n_tiles);
  const int64_t p0 = (int64_t)t * tile_pts;
  const int64_t p1 = min(nx, p0 + tile_pts) - 1;
  const double wb = st_wb[s];
  const int64_t first = warp_bisect<false>(nu, n_lines, grid[p0] - wb, lane);
  const int64_t last = warp_bisect<true>(nu, n_lines, grid[p1] + wb, lane);
  if (lane == 0) {
    bins[2 * i] = (int32_t)first;
    bins[2 * i + 1] = (int32_t)last;
  }
}

// Searches on the shared-memory tile with an arithmetic first guess.  The answers are defined by exact comparisons
// on the grid values (identical to bisect); the guess only shortens the walk: on a uniform grid it is off by at
// most one, on an irregular grid the walk is bounded and falls back to bisection.
struct TileScale { double g0, inv_dv; int n; };

__device__ __forceinline__ int guess_index(const TileScale& ts, double v) {
  const double t = (v - ts.g0) * ts.inv_dv;
  return (int)fmin(fmax(t, -1.0), (double)ts.n) + 1;          // ~ number of elements <= v
}
// number of elements <= v among a[0..n)   (== bisect_right)
__device__ __forceinline__ int count_le(const double* a, const TileScale& ts, double v) {
  int g = min(max(guess_index(ts, v), 0), ts.n);
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    if (g > 0 && v < a[g - 1]) { --g; continue; }
    if (g < ts.n && !(v < a[g])) { ++g; continue; }
    return g;
  }
  return bisect_right_dev(a, ts.n, v);
}
// number of elements < v among a[0..n)   (== bisect_left)
__device__ __forceinline__ int count_lt(const double* a, const TileScale& ts, double v) {
  int g = min(max(guess_index(ts, v), 0), ts.n);
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    if (g > 0 && !(a[g - 1] < v)) { --g; continue; }
    if (g < ts.n && a[g] < v) { ++g; continue; }
    return g;
  }
  return bisect_left_dev(a, ts.n, v);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifndef RTX_NPW
#define RTX_NPW 2
#endif
constexpr int NPW = RTX_NPW;   // producer warps of the warp-specialised variant
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}

// MODE 0: FP64, full-precision reciprocal; 1: FP64, one Newton step fewer; 2: FP32 evaluation of the asymptote branch
// (the optional 1e-5 path of the north star): per-slot origins keep omega-nu0' accurate in single precision, partial
// sums are kept in FP32 for at most one chunk of lines and then folded into the FP64 accumulators; the Weideman
// cores stay in FP64.
// WS = warp-specialised variant: NPW extra producer warps run the prologue one chunk ahead (double-buffered
// records, mbarrier full/empty handshake); the eight consumer warps never meet at a block-wide barrier.
// TFP = tile-level far-field blocks compiled in (K = 8 only; its own instantiation, so that the variant for windows of
// a few thousand points carries none of its code)
template <int K, int MODE, bool WS, bool EXT = false, bool TFP = false>
__global__ void __launch_bounds__(WS ? NT + 32 * NPW : NT, WS ? 2 : (K == 4 ? RTX_MINBLOCKS_K4 : K <= 2 ? RTX_MINBLOCKS_SMALL : RTX_MINBLOCKS)) voigt_tile_kernel(const VoigtParams P) {
  constexpr bool FAST = (MODE == 1);
  constexpr bool F32 = (MODE == 2);
  constexpr int TILE = K * NT;
  constexpr int NBUF = WS ? 2 : 1;
  constexpr int NTH = WS ? NT + 32 * NPW : NT;
  constexpr int LCHUNK = ChunkCfg<K, MODE, WS>::LCH;   // lines per chunk
  constexpr int NBLK = LCHUNK / 32;                    // 32-line blocks per chunk
  constexpr int ROUNDS = LCHUNK / NT;                  // prologue rounds per chunk
  using P1Rec = typename std::conditional<F32, WideRec, SegRec>::type;   // what pass 1 reads
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_grid = reinterpret_cast<double*>(smem_raw);                 // TILE
  double* s_acc2 = s_grid + TILE;                                       // TILE
  WideRec* s_wide = reinterpret_cast<WideRec*>(s_acc2 + TILE);          // NBUF * LCHUNK, FP32 mode only
  SegRec* s_seg = reinterpret_cast<SegRec*>(s_wide + (F32 ? NBUF * LCHUNK : 0));   // NBUF * LCHUNK
  // per block of 32 lines (one warp, one prologue round) and per consumer warp: which of the 32 records that warp
  // must visit -- windows covering its whole slice, windows touching it, segments touching it
  __shared__ unsigned s_mfull[NBUF][NBLK][NT / 32], s_mtouch[NBUF][NBLK][NT / 32], s_mseg[NBUF][NBLK][NT / 32],
      s_mfar[NBUF][NBLK][NT / 32];
  __shared__ __align__(8) unsigned long long s_bar, s_full[2], s_empty[2];

  // automatic tile size without a host round trip: the K = 8 and the K = 4 grid are both launched and the one the
  // window statistic (window_stats_kernel) does not pick retires at once
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x % P.n_tiles;
  const int s = (blockIdx.x / P.n_tiles) % P.n_states;
  const int part = (blockIdx.x / P.n_tiles) / P.n_states;      // 0 unless the lines are split
  if (P.sel != nullptr) {
    const double inv_dv = P.sel[2];
    const double wt = P.st_wt[s];
    const bool tf = P.sel_tf == 2 ||                      // RTX_VOIGT_TILEFAR: wherever a line qualifies
                    (P.sel_tf && wt * inv_dv >= RTX_TF_WINDOW &&
                     2.0 * P.dens * (wt * inv_dv - 1.375 * 2048.0) >= RTX_TF_MINLINES * inv_dv);   // est. lines that qualify per tile
    const int want = (!P.sel_k8 && P.st_wb[s] * inv_dv < RTX_K4_WINDOW) ? 0 : (tf ? 2 : 1);
    if (want != (K == 4 ? 0 : TFP ? 2 : 1)) return;
  }
  const int64_t p0 = (int64_t)tile * TILE;
  const int npts = (int)min((int64_t)TILE, P.nx - p0);

  // ---- stage the grid tile: one TMA bulk copy signalled through an mbarrier --------------------
  const int tma_pts = P.use_tma ? (npts & ~1) : 0;
  if (tma_pts > 0) {
    if (tid == 0) {
      const uint32_t bar = smem_u32(&s_bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = (uint32_t)tma_pts * 8u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(s_grid)),
          "l"(P.grid + p0), "r"(bytes), "r"(bar)
          : "memory");
    }
  }
  if (WS && tid == 0) {
    mbar_init(&s_full[0], NPW); mbar_init(&s_full[1], NPW);
    mbar_init(&s_empty[0], NT / 32); mbar_init(&s_empty[1], NT / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tma_pts + tid; i < TILE; i += NTH) s_grid[i] = (i < npts) ? P.grid[p0 + i] : 1.7976931348623157e308;
  for (int i = tid; i < TILE; i += NTH) s_acc2[i] = 0.0;
  __syncthreads();   // barrier init visible to all waiters; padding + acc2 written
  if (tma_pts > 0) {
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(bar), "r"(0u)
          : "memory");
    }
  }

  // far-field reach of every slice: a line centred at or below s_far[q] (left of the slice) or at or above
  // s_far[8 + q] (right of it) is at least RTX_FAR_FRAC slice widths away from every point of slice q
  __shared__ double s_far[2 * (NT / 32)];
  // tile-level far field: a line centred at or below s_tfar[0] / at or above s_tfar[1] is at least RTX_FAR_FRAC tile
  // widths away from every point of the tile; s_tfar[2] = tile centre, s_tfar[3] = node half-width
  __shared__ double s_tfar[4];
  __shared__ unsigned s_mtile[NBUF][NBLK];
  __shared__ unsigned s_msdf[NBUF][EXT ? NBLK : 1][NT / 32];   // speed-dependent lines: slices summed at their nodes
  constexpr bool TF = TFP && (RTX_TILEFAR != 0) && !F32 && !WS;
  const bool tfar_on = TF && P.tfar_ok != 0 && npts == TILE && !P.doppler;
  if (tid < NT / 32) {
    const double gf = s_grid[tid * (32 * K)], gl = s_grid[tid * (32 * K) + 32 * K - 1];
    const double D = RTX_FAR_FRAC * (gl - gf);
    s_far[tid] = gf - D;
    s_far[NT / 32 + tid] = gl + D;
  }
  if (tid == 32) {
    const double gf = s_grid[0], gl = s_grid[TILE - 1];
    const double D = RTX_FAR_FRAC * (gl - gf);
    s_tfar[0] = gf - D;
    s_tfar[1] = gl + D;
    s_tfar[2] = 0.5 * (gf + gl);
    s_tfar[3] = 0.5 * (gl - gf) * ((double)TILE / (double)(TILE - 1));
  }
  __syncthreads();

  double acc[K];
  // this lane's K grid values: kept in shared memory and re-read per visit (RTX_W_SMEM, default) -- most visits are
  // far-field node visits that never touch them, and the 2*K registers are worth more to the prologue
#ifndef RTX_W_SMEM
#define RTX_W_SMEM 1
#endif
  const double* __restrict__ wp = s_grid + (warp & (NT / 32 - 1)) * (32 * K) + lane;
#if RTX_W_SMEM
#define RTX_W(j) wp[(j) * 32]
#else
  double w[K];
#define RTX_W(j) w[j]
#endif
  float w32[K], w32l[K], a32[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
#if !RTX_W_SMEM
    w[j] = wp[j * 32];
#endif
    const double off = wp[j * 32] - s_grid[0];                          // offset from the tile's first point, as hi + lo
    w32[j] = (float)off;
    w32l[j] = (float)(off - (double)w32[j]);
    a32[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = 0.0;

  // far-field nodes of this warp's slice: omega_k = centre + halfwidth*cos(theta_k), one per lane
  // The node offset from the slice centre and the centre's offset from the line are formed separately: a node
  // position rounded to a double near 1000 cm^-1 would be off by ~1e-13, i.e. 2*1e-13/d ~ 1e-11 relative in a term
  // that varies like 1/d^2, while (centre - nu0') + offset carries only the relative rounding of d itself.
  const int own_w = warp & (NT / 32 - 1);
  double xoff, sl_c;
  {
    const double sl_first = s_grid[own_w * (32 * K)], sl_last = s_grid[own_w * (32 * K) + 32 * K - 1];
    sl_c = 0.5 * (sl_first + sl_last);
    xoff = (0.5 * (sl_last - sl_first) * ((double)(32 * K) / (double)(32 * K - 1))) * P.cheb[lane];
  }
  double accn = 0.0;
  bool far_used = false;
  double acct = 0.0;          // this warp's share of the tile-level node sums (node = lane)

  TileScale ts;
  ts.g0 = s_grid[0];
  ts.n = npts;
  ts.inv_dv = (npts > 1 && s_grid[npts - 1] > s_grid[0]) ? (double)(npts - 1) / (s_grid[npts - 1] - s_grid[0]) : 0.0;
  const double T = P.T[s], p = P.p[s];
  const double lnr = log(RTX_TREF / T);
  const int64_t bi = 2 * ((int64_t)s * P.n_tiles + tile);
  int first = P.bins[bi], last = P.bins[bi + 1];
  if (P.split > 1) {
    const int64_t len = last - first;
    last = first + (int)(len * (part + 1) / P.split);
    first = first + (int)(len * part / P.split);
  }
  const int own_lo = warp * (TILE / (NT / 32)), own_hi = own_lo + TILE / (NT / 32);

  // Lines are dealt to chunks round-robin (chunk c takes candidates c, c+nchunks, ...), so every chunk spans
  // the tile's whole candidate range and the eight warps' slices receive equal shares of every chunk.
  const int ncand = last - first;
  const int nchunks = (ncand + LCHUNK - 1) / LCHUNK;
  // one block of 32 candidate lines -> records + visit masks in buffer `buf` (executed by one full warp)
  auto produce = [&](int c, int blk, int buf) {
      WideRec* wide_b = s_wide + buf * LCHUNK;
      SegRec* seg_b = s_seg + buf * LCHUNK;
      const int k = blk * 32 + lane;
      const int li = first + c + k * nchunks;
      bool is_wide = false, has_seg = false;
      SegRec sr;
      sr.sub = 0;
      double h_keep = 0.0;
      bool is_tfar = false;
      unsigned sdf = 0;
      if (li < last) {
        LineState L = line_state<EXT>(P, li, s, T, p, lnr);
        // Tile-level far field: the window covers the whole tile (the same comparisons the two searches below would
        // make, misc/hapi.py:11133-11134), the line centre is >= RTX_FAR_FRAC tile widths away and no point of the tile
        // is in the Weideman region (the predicate is monotone in |omega - nu0'|, so the nearest point decides).
        // Such a line is evaluated once per tile at the tile's 32 Chebyshev nodes instead of once per slice.
        double cte = 0.0, y = 0.0;
        bool tf = false;
        const bool sdline = EXT && P.sdv && L.gam2 != 0.0;   // speed-dependent line: pcqsdhc instead of the Voigt profile
        if (TF && tfar_on && (L.sg0 <= s_tfar[0] || L.sg0 >= s_tfar[1]) && (s_grid[0] > L.nu - L.W) &&
            (s_grid[TILE - 1] <= L.nu + L.W)) {
          cte = RTX_SQRT_LN2 / L.gamD;
          y = P.lorentz ? 1e300 : L.gam0 * cte;
          const bool left = L.sg0 <= s_tfar[0];
          const double gn = left ? s_grid[0] : s_grid[TILE - 1], gf = left ? s_grid[TILE - 1] : s_grid[0];
          tf = sdline ? sdv_far_ok(gn - L.sg0, gf - L.sg0, L.gam0 - 1.5 * L.gam2, L.gam2, cte)
                      : (!(y < 15.0) || !in_weideman(gn, L.sg0, cte, y));
        }
        {
          int lo = 0, hi = TILE;
          if (!tf) {
            lo = count_le(s_grid, ts, L.nu - L.W);                // bisect(Omegas, nu-W), misc/hapi.py:11133
            hi = count_le(s_grid, ts, L.nu + L.W);                // bisect(Omegas, nu+W), misc/hapi.py:11134
          }
          // the bins are cut with a per-state bound on the wing, so a good part of the candidates does not reach the
          // tile: their strength is never computed (blocks of 32 candidates are neighbours in nu, so whole warps skip)
          if (hi > lo) line_strength(P, li, s, T, L);
          if (hi > lo && L.keep) {
            // Lorentz profile Gam0/(pi*(Gam0^2+d^2)) (misc/hapi.py:10142-10150) = the closed form below with h = 0 and
            // no Weideman region
            cte = RTX_SQRT_LN2 / L.gamD;                                   // misc/hapi.py:9900
            y = P.lorentz ? 1e300 : L.gam0 * cte;                          // misc/hapi.py:9913
            const double h = P.lorentz ? 0.0 : L.gamD * L.gamD / RTX_LN2;  // 1/cte^2
            const double g2 = L.gam0 * L.gam0;
            const double a = L.amp * L.gam0 / RTX_PI;
            sr.sg0 = L.sg0;
            sr.a = a;
            sr.ah = a * h;
            sr.m = g2 - 0.5 * h;
            sr.c = 2.0 * h * g2;
            h_keep = h;
            sr.cte = cte;
            sr.y = y;
            sr.cw = L.amp * (cte / RTX_SQRT_PI);
            int clo = lo, chi = lo;                                        // empty Weideman core
            if (y < 15.0 && !P.doppler && !tf) {
              // the mask |x|+y<15 is a contiguous index range around sg0; find it with the exact predicate
              // (exact predicate of hum1_wei; it is monotone on either side of sg0, so each end is found by a
              // short walk from the arithmetic estimate sg0 -+ (15-y)/cte, with bisection as the fallback)
              int m = count_lt(s_grid, ts, L.sg0);
              m = max(lo, min(hi, m));
              const double hw = (15.0 - y) / cte;
              {   // first true in [lo, m)
                int g = max(lo, min(m, count_lt(s_grid, ts, L.sg0 - hw)));
                bool done = false;
#pragma unroll 1
                for (int it = 0; it < 4 && !done; ++it) {
                  if (g > lo && in_weideman(s_grid[g - 1], L.sg0, cte, y)) --g;
                  else if (g < m && !in_weideman(s_grid[g], L.sg0, cte, y)) ++g;
                  else done = true;
                }
                if (!done) {
                  int a0 = lo, a1 = m;
                  while (a0 < a1) {
                    const int mid = (a0 + a1) >> 1;
                    if (in_weideman(s_grid[mid], L.sg0, cte, y)) a1 = mid; else a0 = mid + 1;
                  }
                  g = a0;
                }
                clo = g;
              }
              {   // first false in [m, hi)
                int g = max(m, min(hi, count_le(s_grid, ts, L.sg0 + hw)));
                bool done = false;
#pragma unroll 1
                for (int it = 0; it < 4 && !done; ++it) {
                  if (g > m && !in_weideman(s_grid[g - 1], L.sg0, cte, y)) --g;
                  else if (g < hi && in_weideman(s_grid[g], L.sg0, cte, y)) ++g;
                  else done = true;
                }
                if (!done) {
                  int a0 = m, a1 = hi;
                  while (a0 < a1) {
                    const int mid = (a0 + a1) >> 1;
                    if (in_weideman(s_grid[mid], L.sg0, cte, y)) a0 = mid + 1; else a1 = mid;
                  }
                  g = a0;
                }
                chi = g;
              }
            }
            sr.lo = (short)lo; sr.hi = (short)hi; sr.clo = (short)clo; sr.chi = (short)chi;
            if (sdline) {
              sr.a = L.gam0 - 1.5 * L.gam2;             // Re c0t
              sr.ah = L.gam2;
              sr.cw = L.amp;
              sr.sub = 2;
            }
            if (tf) {
              is_tfar = true;                           // the whole tile at its 32 nodes; no other visit
            } else if (sdline) {
              // speed-dependent line (Gamma2 != 0): the window goes to pass 2, evaluated with pcqsdhc PART2-4 -- except
              // the slices it covers from far away, where the profile is the smooth difference of two asymptotes
              has_seg = true; sr.s = sr.clo = (short)lo; sr.e = sr.chi = (short)hi;
              if (K >= 4 && P.far_ok != 0) {
#pragma unroll 1
                for (int q = 0; q < NT / 32; ++q) {
                  if (lo > q * 32 * K || hi < (q + 1) * 32 * K) continue;             // not covered completely
                  const bool left = sr.sg0 <= s_far[q];
                  if (!left && !(sr.sg0 >= s_far[NT / 32 + q])) continue;             // too close
                  const double g0 = s_grid[q * 32 * K], g1 = s_grid[q * 32 * K + 32 * K - 1];
                  if (sdv_far_ok((left ? g0 : g1) - sr.sg0, (left ? g1 : g0) - sr.sg0, sr.a, sr.ah, cte)) sdf |= 1u << q;
                }
              }
            } else if (P.doppler) {
              // cSqrtLn2divSqrtPi*exp(-cLn2*((sg-sg0)/GamD)**2)/GamD (misc/hapi.py:10160, hapi's truncated constants):
              // the whole window is a "core" of pass 2, evaluated with exp instead of the rational series
              sr.cte = 0.8325546111576705 / L.gamD;     // sqrt(cLn2), cLn2 = 0.6931471805599
              sr.cw = L.amp * 0.469718639319144059835 / L.gamD;
              has_seg = true; sr.s = sr.clo = (short)lo; sr.e = sr.chi = (short)hi;
            } else if (hi - lo <= NARROW_MAX) {
              has_seg = true; sr.s = (short)lo; sr.e = (short)hi;
            } else {
              is_wide = true;
              sr.s = (short)clo; sr.e = (short)chi;                          // pass 2 takes the core only
              // The slice that holds the core of a wide line would need per-point tests in pass 1.  Instead pass 1
              // treats it like any covered slice (closed form on every point, finite for Gamma0 > 0) and pass 2 adds
              // series - closed form on the core points.  The closed form peaks at amp/(2 pi Gamma0), 1/(2 sqrt(pi) y)
              // times the true peak, so the cancellation costs < 3e-13 relative for y >= 1e-4 with the fully refined
              // reciprocal; with the one-step reciprocal (9e-13 per term, and pass 1 pairs lines differently from pass 2)
              // it costs 9e-13/(2 sqrt(pi) y), hence y >= 0.1 there (< 3e-12).  Below that (and in the FP32 mode) the
              // slice keeps the tested path.
              sr.sub = (!F32 && RTX_CORE_SUB && y >= (FAST ? 1e-1 : 1e-4) && y < 15.0) ? 1 : 0;
              has_seg = chi > clo;
            }
          }
        }
      }
      // records live at fixed slots; the per-consumer-warp ballots below are the (deterministic) compaction
      unsigned wm = 0, sm = 0;
      if (is_wide) {
        if (F32) {
          WideRec32 w3;
          const double c = sr.sg0 - s_grid[0];
          w3.chi = (float)c; w3.clo = (float)(c - (double)w3.chi);
          w3.af = (float)sr.a; w3.hf = (float)h_keep; w3.mf = (float)sr.m; w3.cf = (float)sr.c;
          w3.pad0 = w3.pad1 = 0.f;
          w3.lo = sr.lo; w3.hi = sr.hi; w3.core_lo = sr.clo; w3.core_hi = sr.chi;
          *reinterpret_cast<WideRec32*>(&wide_b[k]) = w3;
        }
        wm = slice_masks<32 * K>(sr.lo, sr.hi, sr.clo, sr.chi, sr.sub != 0, K >= 4 && P.far_ok != 0, sr.sg0, s_far);   // K < 4: too few points per node to pay
      }
      if (is_wide || has_seg || is_tfar) seg_b[k] = sr;
      if (TF) {
        const unsigned bt = __ballot_sync(0xffffffffu, is_tfar);
        if (lane == 0) s_mtile[buf][blk] = bt;
      }
      if (has_seg) {
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) sm |= ((sr.e > q * 32 * K) && (sr.s < (q + 1) * 32 * K) ? 1u : 0u) << q;
        sm &= ~sdf;
      }
      if (EXT) {
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) {
          const unsigned bq = __ballot_sync(0xffffffffu, (sdf >> q) & 1u);
          if (lane == q) s_msdf[buf][blk][q] = bq;
        }
      }
#pragma unroll
      for (int q = 0; q < NT / 32; ++q) {
        const unsigned bt = __ballot_sync(0xffffffffu, (wm >> q) & 1u);
        const unsigned bf = __ballot_sync(0xffffffffu, (wm >> (8 + q)) & 1u);
        const unsigned bs = __ballot_sync(0xffffffffu, (sm >> q) & 1u);
        const unsigned bn = __ballot_sync(0xffffffffu, (wm >> (16 + q)) & 1u);
        if (lane == q) {
          s_mtouch[buf][blk][q] = bt & ~bf & ~bn; s_mfull[buf][blk][q] = bf; s_mseg[buf][blk][q] = bs;
          s_mfar[buf][blk][q] = bn;
        }
      }
  };
  // passes 1 and 2 of one consumer warp over the records of buffer `buf`
  auto consume = [&](int buf) {
    // ---- pass 1: wide lines, register accumulators ------------------------------------------------
#define RTX_EVAL(j)                                                                                   \
  do {                                                                                                \
    if (F32) {                                                                                        \
      const float d_ = (w32[j] - chf) + (w32l[j] - clf);                                              \
      const float u_ = fmaf(d_, d_, mf);                                                              \
      const float den_ = fmaf(u_, u_, cf);                                                            \
      float r_;                                                                                       \
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r_) : "f"(den_));                                       \
      a32[j] = fmaf(af, (u_ + hf) * r_, a32[j]);                                                      \
    } else {                                                                                          \
      acc[j] = cheap_term<FAST>(RTX_W(j), sg0, a, ah, m, c, acc[j]);                                      \
    }                                                                                                 \
  } while (0)
#define RTX_VISIT_SETUP()                                                                             \
  const WideRec32& r3 = *reinterpret_cast<const WideRec32*>(&rr);                                     \
  double sg0 = 0, a = 0, ah = 0, m = 0, c = 0;                                                        \
  float chf = 0.f, clf = 0.f, af = 0.f, hf = 0.f, mf = 0.f, cf = 0.f;                                 \
  int wlo, whi, wclo, wchi;                                                                           \
  if (F32) {                                                                                          \
    chf = r3.chi; clf = r3.clo; af = r3.af; hf = r3.hf; mf = r3.mf; cf = r3.cf;                       \
    wlo = r3.lo; whi = r3.hi; wclo = r3.core_lo; wchi = r3.core_hi;                                   \
  } else {                                                                                            \
    sg0 = rr.sg0; a = rr.a; ah = rr.ah; m = rr.m; c = rr.c;                                     \
    wlo = rr.lo; whi = rr.hi; wclo = rr.clo; wchi = rr.chi;                                           \
  }
    // A warp owns 32*K contiguous points (lane-strided).  Windows that span the warp's slice take the test-free
    // path; only the slices holding a window end (or a Weideman core) pay for per-point predicates.
    if (TF && tfar_on) {
      // tile-level far-field visits: warp q takes records 4q..4q+3 of every block of 32 at the tile's 32 nodes
      const double tc = s_tfar[2];
      const double tx = s_tfar[3] * P.cheb[lane];
#pragma unroll 1
      for (int blk = 0; blk < NBLK; ++blk) {
        unsigned mt = s_mtile[buf][blk] & (0xFu << (4 * warp));
        const SegRec* __restrict__ trecs = s_seg + buf * LCHUNK + blk * 32;
        if (EXT) {                                  // speed-dependent records one by one, the others below
          unsigned ms = 0;
          if ((mt >> lane) & 1u) ms = trecs[lane].sub == 2 ? 1u : 0u;
          ms = __ballot_sync(0xffffffffu, ms != 0u);
          mt &= ~ms;
          while (ms) {
            const SegRec& ra = trecs[__ffs(ms) - 1];
            ms &= ms - 1;
            acct = fma(ra.cw, sdv_far_term((tc - ra.sg0) + tx, ra.a, ra.ah, ra.cte), acct);
          }
        }
        while (mt & (mt - 1)) {
          const SegRec& ra = trecs[__ffs(mt) - 1];
          mt &= mt - 1;
          const SegRec& rb = trecs[__ffs(mt) - 1];
          mt &= mt - 1;
          acct = cheap_pair_d<FAST>((tc - ra.sg0) + tx, ra.a, ra.ah, ra.m, ra.c, (tc - rb.sg0) + tx, rb.a, rb.ah, rb.m, rb.c,
                                    acct);
        }
        if (mt) {
          const SegRec& ra = trecs[__ffs(mt) - 1];
          acct = cheap_d<FAST>((tc - ra.sg0) + tx, ra.a, ra.ah, ra.m, ra.c, acct);
        }
      }
    }
    // far-field visits: the line is evaluated at this slice's 32 Chebyshev nodes only (one per lane)
#pragma unroll 1
    for (int blk = 0; blk < NBLK; ++blk) {
      if (EXT) {
        const SegRec* __restrict__ frecs = s_seg + buf * LCHUNK + blk * 32;
        unsigned ms = s_msdf[buf][blk][warp];
        if (ms) far_used = true;
        while (ms) {
          const SegRec& ra = frecs[__ffs(ms) - 1];
          ms &= ms - 1;
          accn = fma(ra.cw, sdv_far_term((sl_c - ra.sg0) + xoff, ra.a, ra.ah, ra.cte), accn);
        }
      }
      {
        const SegRec* __restrict__ frecs = s_seg + buf * LCHUNK + blk * 32;
        unsigned mf = s_mfar[buf][blk][warp];
        if (mf) far_used = true;
        while (mf & (mf - 1)) {
          const SegRec& ra = frecs[__ffs(mf) - 1];
          mf &= mf - 1;
          const SegRec& rb = frecs[__ffs(mf) - 1];
          mf &= mf - 1;
          accn = cheap_pair_d<FAST>((sl_c - ra.sg0) + xoff, ra.a, ra.ah, ra.m, ra.c, (sl_c - rb.sg0) + xoff, rb.a, rb.ah, rb.m,
                                    rb.c, accn);
        }
        if (mf) {
          const SegRec& ra = frecs[__ffs(mf) - 1];
          accn = cheap_d<FAST>((sl_c - ra.sg0) + xoff, ra.a, ra.ah, ra.m, ra.c, accn);
        }
      }
    }
#pragma unroll 1
    for (int blk = 0; blk < NBLK; ++blk) {
      const P1Rec* __restrict__ recs = (F32 ? reinterpret_cast<const P1Rec*>(s_wide) : reinterpret_cast<const P1Rec*>(s_seg)) +
                                       buf * LCHUNK + blk * 32;
      unsigned m = s_mfull[buf][blk][warp];
#ifdef RTX_DBG_SKIP_FULL
      m = 0;
#endif
      if (!F32) {
        // FP64 paths: records that cover the slice are taken two at a time through one reciprocal
        while (m & (m - 1)) {
          const P1Rec& ra = recs[__ffs(m) - 1];
          m &= m - 1;
          const P1Rec& rb = recs[__ffs(m) - 1];
          m &= m - 1;
          const double sg1 = ra.sg0, a1 = ra.a, ah1 = ra.ah, m1 = ra.m, c1 = ra.c;
          const double sg2 = rb.sg0, a2 = rb.a, ah2 = rb.ah, m2 = rb.m, c2 = rb.c;
#pragma unroll
          for (int j = 0; j < K; ++j) acc[j] = cheap_pair<FAST>(RTX_W(j), sg1, a1, ah1, m1, c1, sg2, a2, ah2, m2, c2, acc[j]);
        }
      }
      while (m) {
        const P1Rec& rr = recs[__ffs(m) - 1];
        m &= m - 1;
        RTX_VISIT_SETUP();
        (void)wlo; (void)whi; (void)wclo; (void)wchi;
#pragma unroll
        for (int j = 0; j < K; ++j) RTX_EVAL(j);
      }
      m = s_mtouch[buf][blk][warp];
#ifdef RTX_DBG_SKIP_TOUCH
      m = 0;
#endif
      while (m) {
        const P1Rec& rr = recs[__ffs(m) - 1];
        m &= m - 1;
        RTX_VISIT_SETUP();
        const int lo = wlo - own_lo, hi = whi - own_lo;              // window in slice coordinates
        const int clo = wclo - own_lo, chi = wchi - own_lo;
        const bool core_here = (chi > clo) && (chi > 0) && (clo < 32 * K);
        if (!core_here && lo > 0 && hi >= 32 * K) {
          // the window starts inside this slice and runs past its end
          const int j0 = lo >> 5, l0 = lo & 31;
#pragma unroll
          for (int j = K - 1; j >= 0; --j) {
            if (j < j0) break;
            if (j > j0 || lane >= l0) RTX_EVAL(j);
          }
        } else if (!core_here && lo <= 0 && hi < 32 * K) {
          // the window ends inside this slice
          const int j1 = (hi - 1) >> 5, l1 = hi - (j1 << 5);
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (j > j1) break;
            if (j < j1 || lane < l1) RTX_EVAL(j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const int i = j * 32 + lane;
            const bool in = (i >= lo) && (i < hi) && !((i >= clo) && (i < chi));
            if (in) RTX_EVAL(j);
          }
        }
      }
    }
    if (F32) {
#pragma unroll
      for (int j = 0; j < K; ++j) { acc[j] += (double)a32[j]; a32[j] = 0.f; }
    }
    // ---- pass 2: segments (Weideman cores, narrow windows) inside the warp's slice -------------------------
    // Wings and cores are taken separately so that no warp step mixes the 8-operation closed form with the
    // 24-term series: (a) the wing points of each record (left and right of its core, concatenated) are dealt to
    // the lanes record by record; (b) the cores of the block's records are flattened -- (record, point) pairs
    // dealt to the lanes across record boundaries -- and lanes that hit the same grid point in one step are
    // found with match.any and summed in lane order by the lowest lane: no atomics, summation order fixed.
#pragma unroll 1
    for (int blk = 0; blk < NBLK; ++blk) {
      unsigned msk = s_mseg[buf][blk][warp];
#ifdef RTX_DBG_SKIP_PASS2
      msk = 0;
#endif
      if (!msk) continue;
      const SegRec* __restrict__ recs = s_seg + buf * LCHUNK + blk * 32;
      int cs = 0, n = 0, nw = 0;
      if ((msk >> lane) & 1u) {
        const SegRec& q = recs[lane];
        cs = max((int)q.clo, own_lo);
        n = max(min((int)q.chi, own_hi) - cs, 0);
        nw = max(min((int)q.clo, own_hi) - max((int)q.s, own_lo), 0) + max(min((int)q.e, own_hi) - max((int)q.chi, own_lo), 0);
      }
      // (a) wings
      unsigned m2 = __ballot_sync(0xffffffffu, nw > 0);
      while (m2) {
        const SegRec& r = recs[__ffs(m2) - 1];
        m2 &= m2 - 1;
        const double sg0 = r.sg0, a = r.a, ah = r.ah, mm = r.m, c = r.c;
        const int ls = max((int)r.s, own_lo), nl = max(min((int)r.clo, own_hi) - ls, 0);
        const int rs = max((int)r.chi, own_lo), nr = max(min((int)r.e, own_hi) - rs, 0);
        const int nt = nl + nr;
        int t = lane;
        for (; t + 32 < nt; t += 64) {               // two rows per step: independent chains, disjoint points
          const int i0 = t < nl ? ls + t : rs + (t - nl);
          const int i1 = t + 32 < nl ? ls + t + 32 : rs + (t + 32 - nl);
          const double v0 = cheap_term<false>(s_grid[i0], sg0, a, ah, mm, c, s_acc2[i0]);
          const double v1 = cheap_term<false>(s_grid[i1], sg0, a, ah, mm, c, s_acc2[i1]);
          s_acc2[i0] = v0;
          s_acc2[i1] = v1;
        }
        if (t < nt) {
          const int i = t < nl ? ls + t : rs + (t - nl);
          s_acc2[i] = cheap_term<false>(s_grid[i], sg0, a, ah, mm, c, s_acc2[i]);
        }
        __syncwarp();
      }
      // (b) cores
      int incl = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      const int excl = incl - n;
      auto core_steps = [&](int t0, auto NHc) {
        // NH steps of 32 (record, point) pairs at once: NH independent series per lane
        constexpr int NH = decltype(NHc)::value;
        int ii[NH];
        double vv[NH];
        bool ok[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int t = t0 + 32 * h + lane;
          ok[h] = t < total;
          int r = 0;                                 // number of records whose inclusive count is <= t
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, r + sft - 1);
            if (v <= t) r += sft;
          }
          r = min(r, 31);
          ii[h] = __shfl_sync(0xffffffffu, cs, r) + (t - __shfl_sync(0xffffffffu, excl, r));
          const SegRec& rec = recs[r];
          const int ic = ok[h] ? ii[h] : own_lo;
          const double om = s_grid[ic];
          const double x = __dmul_rn(om - rec.sg0, rec.cte);
          vv[h] = rec.cw * ((EXT && rec.sub == 2) ? sdvoigt_re(om - rec.sg0, rec.a, rec.ah, rec.cte)
                            : P.doppler  ? exp(-(x * x))
                                         : weideman_re(x, rec.y));
          if (rec.sub == 1 && rec.lo <= own_lo && rec.hi >= own_hi)   // pass 1 took this slice test-free
            vv[h] = cheap_term<FAST>(om, rec.sg0, -rec.a, -rec.ah, rec.m, rec.c, vv[h]);   // same reciprocal as pass 1
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int i = ii[h];
          const bool valid = ok[h];
          const double v = valid ? vv[h] : 0.0;
          const unsigned peers = __match_any_sync(0xffffffffu, valid ? i : TILE + lane);
          const int gmax = __reduce_max_sync(0xffffffffu, __popc(peers));
          double sum = 0.0;
          unsigned pm = peers;
#pragma unroll 1
          for (int k = 0; k < gmax; ++k) {
            const int src = pm ? __ffs(pm) - 1 : lane;
            const double val = __shfl_sync(0xffffffffu, v, src);
            if (pm) sum += val;
            pm &= pm - 1;
          }
          if (valid && lane == __ffs(peers) - 1) s_acc2[i] += sum;
          __syncwarp();
        }
      };
      int t0 = 0;
#pragma unroll 1
      for (; RTX_P2_ILP == 2 && t0 + 32 < total; t0 += 64) core_steps(t0, std::integral_constant<int, RTX_P2_ILP>{});
#pragma unroll 1
      for (; t0 < total; t0 += 32) core_steps(t0, std::integral_constant<int, 1>{});
    }
  };

  if (!WS) {
    for (int c = 0; c < nchunks; ++c) {
#pragma unroll 1
      for (int r = 0; r < ROUNDS; ++r) produce(c, r * (NT / 32) + warp, 0);
      __syncthreads();
      consume(0);
      __syncthreads();
    }
  } else if (warp >= NT / 32) {
    // ---- producer warps: derive line states one chunk ahead of the consumers (mbarrier full/empty handshake) ----
    const int pw = warp - NT / 32;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      if (c >= 2) mbar_wait(&s_empty[buf], ((c >> 1) - 1) & 1);
#pragma unroll 1
      for (int blk = pw; blk < NBLK; blk += NPW) produce(c, blk, buf);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[buf]);
    }
    return;
  } else {
    // ---- consumer warps: each runs at its own pace, at most one chunk apart from the slowest ----
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&s_full[buf], (c >> 1) & 1);
      consume(buf);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[buf]);
    }
  }

  // ---- far-field epilogue: node sums -> Chebyshev coefficients (lane m holds c_m) -> Clenshaw at the slice's points
  // (slice level first, then the tile level: the eight warps' shares are summed in warp order through shared memory)
  double tsum = 0.0;
  if (TF && tfar_on) {                             // block-uniform; the chunk loop ended with a barrier
    double* s_t = reinterpret_cast<double*>(s_seg);
    s_t[warp * 32 + lane] = acct;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NT / 32; ++q) tsum += s_t[q * 32 + lane];
  }
  const bool tile_used = TF && __any_sync(0xffffffffu, tsum != 0.0);
#pragma unroll 1
  for (int level = 0; level < (TF ? 2 : 1); ++level) {
    if (level == 0 ? !far_used : !tile_used) continue;   // warp-uniform
    const double nodes = level == 0 ? accn : tsum;
    const double* __restrict__ dct = P.cheb + 32;  // [k][m] = (2 - [m==0])/32 * cos(m*theta_k)
    double cm = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) cm = fma(__shfl_sync(0xffffffffu, nodes, k), dct[k * 32 + lane], cm);
    const double sl_first = s_grid[own_w * (32 * K)], sl_last = s_grid[own_w * (32 * K) + 32 * K - 1];
    const double inv_hw = level == 0 ? 1.0 / (0.5 * (sl_last - sl_first) * ((double)(32 * K) / (double)(32 * K - 1)))
                                     : 1.0 / s_tfar[3];
    const double ctr = level == 0 ? 0.5 * (sl_first + sl_last) : s_tfar[2];
    const double c0 = __shfl_sync(0xffffffffu, cm, 0);
    constexpr int G = K < 4 ? K : 4;               // rows per Clenshaw sweep (register budget)
#pragma unroll
    for (int j0 = 0; j0 < K; j0 += G) {
      double x2[G], b1[G], b2[G];
#pragma unroll
      for (int j = 0; j < G; ++j) { x2[j] = 2.0 * ((RTX_W(j0 + j) - ctr) * inv_hw); b1[j] = 0.0; b2[j] = 0.0; }
#pragma unroll 1
      for (int mm = 31; mm >= 1; --mm) {
        const double cc = __shfl_sync(0xffffffffu, cm, mm);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const double b0 = fma(x2[j], b1[j], cc - b2[j]);
          b2[j] = b1[j];
          b1[j] = b0;
        }
      }
#pragma unroll
      for (int j = 0; j < G; ++j) acc[j0 + j] += fma(0.5 * x2[j], b1[j], c0 - b2[j]);
    }
  }

  double* out = (P.split > 1 ? P.part_out + (int64_t)part * P.n_states * P.nx : P.out) + (int64_t)s * P.nx + p0;   // only consumer warps get here in the WS variant
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int i = warp * (32 * K) + j * 32 + lane;
    if (i < npts) out[i] = acc[j] + s_acc2[i];
  }
}

// partial sums of a line-split launch, added in part order
__global__ void sum_parts_kernel(const double* __restrict__ part, int split, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = part[i];
  for (int k = 1; k < split; ++k) a += part[(int64_t)k * n + i];
  out[i] = a;
}

// ---------------------------------------------------------------------------------------------
// exact eval counter: one thread per (line, state), bisects the whole grid.
__global__ void count_kernel(const VoigtParams P, unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long ev = 0, evw = 0;
  if (i < P.n_lines * (int64_t)P.n_states) {
    const int s = (int)(i / P.n_lines);
    const int64_t li = i % P.n_lines;
    const double T = P.T[s], p = P.p[s];
    LineState L = line_state<true>(P, li, s, T, p, log(RTX_TREF / T));
    line_strength(P, li, s, T, L);
    if (L.keep) {
      const int64_t lo = bisect_right_dev64(P.grid, P.nx, L.nu - L.W);
      const int64_t hi = bisect_right_dev64(P.grid, P.nx, L.nu + L.W);
      if (hi > lo) {
        ev = (unsigned long long)(hi - lo);
        const double cte = RTX_SQRT_LN2 / L.gamD, y = (P.lorentz || P.doppler) ? 1e300 : L.gam0 * cte;
        if (y < 15.0) {
          int64_t m = bisect_left_dev64(P.grid, P.nx, L.sg0);
          m = max(lo, min(hi, m));
          int64_t a0 = lo, a1 = m;
          while (a0 < a1) {
            const int64_t mid = (a0 + a1) >> 1;
            if (in_weideman(P.grid[mid], L.sg0, cte, y)) a1 = mid; else a0 = mid + 1;
          }
          const int64_t clo = a0;
          a0 = m; a1 = hi;
          while (a0 < a1) {
            const int64_t mid = (a0 + a1) >> 1;
            if (in_weideman(P.grid[mid], L.sg0, cte, y)) a0 = mid + 1; else a1 = mid;
          }
          evw = (unsigned long long)(a0 - clo);
        }
      }
    }
  }
  // warp reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    ev += __shfl_down_sync(0xffffffffu, ev, o);
    evw += __shfl_down_sync(0xffffffffu, evw, o);
  }
  if ((threadIdx.x & 31) == 0 && (ev | evw)) {
    atomicAdd(&counts[0], ev);
    atomicAdd(&counts[1], evw);
  }
}

template <int K, int MODE, bool WS>
size_t tile_smem_bytes() {
  return (size_t)K * NT * 8 * 2 + (size_t)(WS ? 2 : 1) * ChunkCfg<K, MODE, WS>::LCH *
                                      (sizeof(SegRec) + (MODE == 2 ? sizeof(WideRec) : 0));
}

template <int K, int MODE, bool WS, bool EXT = false, bool TFP = false>
int launch_tile(const VoigtParams& P, cudaStream_t st) {
  const size_t smem = tile_smem_bytes<K, MODE, WS>();
  RTX_CUDA(cudaFuncSetAttribute(voigt_tile_kernel<K, MODE, WS, EXT, TFP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (int64_t)P.n_tiles * P.n_states * (P.split > 1 ? P.split : 1);
  if (blocks > 2147483647LL) { rtx_set_error("rtx_voigt_sum: too many (tile, state) pairs for one launch"); return RTX_E_INVALID; }
  voigt_tile_kernel<K, MODE, WS, EXT, TFP><<<(unsigned)blocks, WS ? NT + 32 * NPW : NT, smem, st>>>(P);
  RTX_CUDA(cudaGetLastError());
  return RTX_OK;
}

int ensure_scratch(rtx_handle* h, size_t state_doubles, size_t bin_ints) {
  if (state_doubles > h->state_cap) {
    if (h->d_state) cudaFree(h->d_state);
    h->d_state = nullptr; h->state_cap = 0;
    RTX_CUDA(cudaMalloc(&h->d_state, state_doubles * sizeof(double)));
    h->state_cap = state_doubles;
  }
  if (bin_ints > h->bins_cap) {
    if (h->d_bins) cudaFree(h->d_bins);
    h->d_bins = nullptr; h->bins_cap = 0;
    RTX_CUDA(cudaMalloc(&h->d_bins, bin_ints * sizeof(int32_t)));
    h->bins_cap = bin_ints;
  }
  return RTX_OK;
}

int fill_params(rtx_handle* h, VoigtParams& P, int n_states, const double* T, const double* p, const double* weight,
                double x_air, double x_self, double wing_abs, double wing_hw, double s_thr, const double* grid,
                int64_t nx) {
  const LineTable& L = h->lines;
  if (h->n_iso <= 0) { rtx_set_error("isotopologue table not set (rtx_set_iso_table)"); return RTX_E_STATE; }
  if (L.n_pad == 0) { rtx_set_error("no line table uploaded (rtx_lines_upload)"); return RTX_E_STATE; }
  RTX_REQUIRE(n_states > 0 && nx > 0, "n_states and nx must be positive");
  RTX_REQUIRE(T && p && weight && grid, "null device pointer");
  RTX_REQUIRE(x_self == 0.0 || L.gamma_self != nullptr, "Diluent 'self' requested but the table has no gamma_self column");
  P.nu = L.nu; P.s0 = L.s0; P.elower = L.elower; P.gamma_air = L.gamma_air; P.gamma_self = L.gamma_self;
  P.n_air = L.n_air; P.n_self = L.n_self; P.delta_air = L.delta_air; P.deltap_air = L.deltap_air;
  P.delta_self = L.delta_self; P.deltap_self = L.deltap_self; P.iso = L.iso; P.n_lines = L.n;
  P.T = T; P.p = p; P.weight = weight; P.n_iso = h->n_iso; P.n_states = n_states;
  P.x_air = x_air; P.x_self = x_self; P.wing_abs = wing_abs; P.wing_hw = wing_hw; P.s_thr = s_thr;
  P.grid = grid; P.nx = nx;
  P.sd_air = L.sd_air; P.sd_self = L.sd_self;
  P.n_extra = 0; P.sdv = 0;
  for (int e = 0; e < RTX_MAX_EXTRA_DILUENTS; ++e) {
    P.xg[e] = L.xg[e]; P.xn[e] = L.xn[e]; P.xd[e] = L.xd[e]; P.xdp[e] = L.xdp[e]; P.xsd[e] = L.xsd[e];
    P.xx[e] = L.xg[e] ? h->x_extra[e] : 0.0;
    if (P.xx[e] != 0.0) P.n_extra = e + 1;
  }
  P.has_self = (x_self != 0.0) ? 1 : 0;
  P.has_deltap = L.has_deltap ? 1 : 0;
  P.use_tma = ((reinterpret_cast<uintptr_t>(grid) & 15u) == 0) ? 1 : 0;
  P.lorentz = 0;
  P.doppler = 0;
  P.sel = nullptr;
  P.split = 1; P.part_out = nullptr;
  return RTX_OK;
}

int run_prep(rtx_handle* h, VoigtParams& P, cudaStream_t st, double** st_wb_out) {
  const size_t per = (size_t)P.n_states * P.n_iso;
  int rc = ensure_scratch(h, 2 * per + 2 * (size_t)P.n_states + 4, h->bins_cap);
  if (rc) return rc;
  double* st_q = h->d_state;
  double* st_d = st_q + per;
  double* st_wb = st_d + per;
  const LineTable& L = h->lines;
  PrepParams Q;
  Q.n_states = P.n_states; Q.n_iso = P.n_iso; Q.T = P.T; Q.p = P.p; Q.mass = h->d_mass; Q.tipsT = h->d_tipsT;
  Q.q = h->d_q; Q.qref = h->d_qref; Q.st_q = st_q; Q.st_d = st_d; Q.st_wb = st_wb;
  Q.x_air = P.x_air; Q.x_self = P.x_self; Q.wing_abs = P.wing_abs; Q.wing_hw = P.wing_hw;
  Q.nu_max = L.nu_max; Q.gair_max = L.gair_max; Q.gself_max = L.gself_max; Q.nair_min = L.nair_min;
  Q.nair_max = L.nair_max; Q.nself_min = L.nself_min; Q.nself_max = L.nself_max;
  Q.gair_typ = L.gair_typ; Q.gself_typ = L.gself_typ; Q.nair_typ = L.nair_typ; Q.nself_typ = L.nself_typ;
  Q.st_wt = st_wb + P.n_states;
  Q.doppler = P.doppler;
  Q.n_extra = P.n_extra;
  for (int e = 0; e < RTX_MAX_EXTRA_DILUENTS; ++e) {
    Q.xx[e] = P.xx[e]; Q.xg_max[e] = L.xg_max[e]; Q.xn_min[e] = L.xn_min[e]; Q.xn_max[e] = L.xn_max[e];
  }
  state_prep_kernel<<<(unsigned)(((int64_t)P.n_states * 32 + 127) / 128), 128, 0, st>>>(Q);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  P.st_q = st_q; P.st_d = st_d;
  P.st_wb = st_wb; P.st_wt = st_wb + P.n_states;
  *st_wb_out = st_wb;
  return RTX_OK;
}

}  // namespace

extern "C" int rtx_weideman_coeffs(double* out24) {
  for (int i = 0; i < 24; ++i) out24[i] = h_weideman[i];
  return RTX_OK;
}

extern "C" int rtx_voigt_sum(rtx_handle* h, int n_states, const double* T_dev, const double* p_dev,
                             const double* weight_dev, double x_air, double x_self, double wing_abs,
                             double wing_hw, double s_threshold, const double* grid_dev, int64_t nx,
                             double* out_dev, int flags, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(out_dev != nullptr, "null output pointer");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VoigtParams P{};
  int rc = fill_params(h, P, n_states, T_dev, p_dev, weight_dev, x_air, x_self, wing_abs, wing_hw, s_threshold,
                       grid_dev, nx);
  if (rc) return rc;
  P.out = out_dev;
  P.lorentz = (flags & RTX_VOIGT_LORENTZ) ? 1 : 0;
  P.doppler = (flags & RTX_VOIGT_DOPPLER) ? ((flags & RTX_VOIGT_NOSHIFT) ? 2 : 1) : 0;
  RTX_REQUIRE(!(P.lorentz && P.doppler), "RTX_VOIGT_LORENTZ and RTX_VOIGT_DOPPLER are exclusive");
  P.sdv = (flags & RTX_VOIGT_SDVOIGT) ? 1 : 0;
  RTX_REQUIRE(!(P.sdv && (P.lorentz || P.doppler)), "RTX_VOIGT_SDVOIGT excludes the Lorentz and Doppler drivers");
  double* st_wb = nullptr;
  rc = run_prep(h, P, st, &st_wb);
  if (rc) return rc;
  int K = 8;                                                  // automatic: K = 8 / K = 4 per state, see below
  const int mode = (flags & RTX_VOIGT_FP32) ? 2 : ((flags & RTX_VOIGT_EXACT_RCP) ? 0 : 1);
  const bool ws = (flags & RTX_VOIGT_WARPSPEC) != 0;
  // speed dependence / extra broadening species: the extended instantiation (FP64, not warp-specialised)
  const bool ext = P.sdv || P.n_extra > 0;
  P.far_ok = (flags & RTX_VOIGT_NOFAR) ? 0 : 1;
  P.tfar_ok = (flags & (RTX_VOIGT_NOFAR | RTX_VOIGT_NOTILEFAR)) ? 0 : 1;
  // the variant with tile-level far-field blocks exists for K = 8, FP64, not warp-specialised
  const bool tf_variant = P.tfar_ok && !P.doppler && (ext || (!ws && mode != 2));
  const int forced = (flags >> RTX_VOIGT_TILE_SHIFT) & 0xF;   // test hook: force the tile size
  bool with4 = false;                                         // K = 8 and K = 4 grids, chosen on the device
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) {
    K = forced;
  } else if ((int64_t)nx * n_states <= 65536 && !(flags & RTX_VOIGT_NOSPLIT)) {
    // tiny problems (one cross-section on some ten thousand points): launch overhead counts, so one grid of 512-point
    // tiles with the lines split (below) instead of the K = 8 / K = 4 pair
    K = 2;
  } else {
    // 2048-point tiles where the windows are thousands of grid points wide; 1024-point tiles at four CTAs per SM where
    // they are narrow (coarse grids, low pressures): smaller slices turn partial visits into test-free ones and hide the
    // per-visit latency.  Each state's own wing decides -- on the device: every candidate grid is launched and the
    // CTAs of the variants a state does not pick return at once, so the call never waits for the GPU.
    with4 = true;
  }
  // Small launches (fewer (tile, state) pairs than the GPU holds CTAs -- a single cross-section): every CTA is bound by
  // the latency of its own chunk loop, so the candidate lines of each (tile, state) are split over up to eight CTAs and
  // the partial spectra added afterwards in part order (bit-reproducible; with a forced tile size there is no split, so
  // that differently sharded grids keep identical bits).
  if (!(flags & RTX_VOIGT_NOSPLIT) && !forced) {              // a forced tile size pins the summation order: no split
    const int64_t nt0 = (nx + (int64_t)K * NT - 1) / ((int64_t)K * NT);
    if (nt0 * n_states < 4LL * h->sm_count) {
      // K <= 2 (4 CTAs per SM, CTAs of a few microseconds): aim at 4 CTAs per SM; K >= 4: fill the slots of one wave
      // but do not start a second one (391 tiles of 2048 points stay unsplit, 157 are split in two)
      P.split = K <= 2 ? (int)std::min<int64_t>(RTX_SPLIT_MAX, (4LL * h->sm_count + nt0 - 1) / nt0)
                       : (int)std::max<int64_t>(1, std::min<int64_t>(RTX_SPLIT_MAX, ((int64_t)RTX_SPLIT_CTAS_PER_SM * h->sm_count) / nt0));
      size_t need = P.split > 1 ? (size_t)P.split * n_states * (size_t)nx : 0;
      if (need > ((size_t)32 << 20)) { P.split = 1; need = 0; }      // 256 MB of partial spectra: not a small problem
      if (need > h->part_cap) {
        if (h->d_part) cudaFree(h->d_part);
        h->d_part = nullptr; h->part_cap = 0;
        RTX_CUDA(cudaMalloc(&h->d_part, need * sizeof(double)));
        h->part_cap = need;
      }
      P.part_out = P.split > 1 ? h->d_part : nullptr;
    }
  }
  const bool with_tf = (K == 8) && tf_variant && P.split == 1;   // small launches: the tile-level variant would be one more idle grid
  P.sel = nullptr; P.sel_k8 = with4 ? 0 : 1; P.sel_tf = with_tf ? ((flags & RTX_VOIGT_TILEFAR) ? 2 : 1) : 0;
  P.dens = (h->lines.nu_max > h->lines.nu_min) ? (double)h->lines.n / (h->lines.nu_max - h->lines.nu_min) : 0.0;
  if (with4 || with_tf) {
    double* stat = st_wb + 2 * (size_t)n_states;
    window_stats_kernel<<<1, 256, 0, st>>>(n_states, st_wb, st_wb + n_states, grid_dev, nx, stat);
    RTX_CUDA(cudaGetLastError());
    h->launches++;
    P.sel = stat;
  }
  h->last_K = with4 ? -1 : K;
  h->d_last_stat = with4 ? P.sel : nullptr;
  h->last_stream = stream;
  const int tile_pts = K * NT;
  P.n_tiles = (int)((nx + tile_pts - 1) / tile_pts);
  const size_t nb = (size_t)P.n_tiles * n_states;
  const int n_tiles4 = with4 ? (int)((nx + 4 * NT - 1) / (4 * NT)) : 0;
  const size_t nb4 = (size_t)n_tiles4 * n_states;
  rc = ensure_scratch(h, h->state_cap, 2 * (nb + nb4));
  if (rc) return rc;
  P.bins = h->d_bins;
  bin_kernel<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, st>>>(n_states, P.n_tiles, tile_pts, grid_dev, nx, P.nu, P.n_lines,
                                                          st_wb, h->d_bins);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  if (with4) {
    bin_kernel<<<(unsigned)((nb4 * 32 + 255) / 256), 256, 0, st>>>(n_states, n_tiles4, 4 * NT, grid_dev, nx, P.nu, P.n_lines,
                                                             st_wb, h->d_bins + 2 * nb);
    RTX_CUDA(cudaGetLastError());
    h->launches++;
  }
  // Chebyshev tables of the far-field blocks (once per handle): cos(theta_k) and the node-values -> coefficients matrix
  if (!h->d_cheb) {
    std::vector<double> tab(32 + 32 * 32);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int k = 0; k < 32; ++k) {
      tab[k] = (double)cosl(pi * (2 * k + 1) / 64.0L);
      for (int m = 0; m < 32; ++m)
        tab[32 + k * 32 + m] = (double)((m == 0 ? 1.0L : 2.0L) / 32.0L * cosl(pi * m * (2 * k + 1) / 64.0L));
    }
    RTX_CUDA(cudaMalloc(&h->d_cheb, tab.size() * sizeof(double)));
    RTX_CUDA(cudaMemcpyAsync(h->d_cheb, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    RTX_CUDA(cudaStreamSynchronize(st));   // tab is a local
  }
  P.cheb = h->d_cheb;
  // default: paired evaluations through a reciprocal with ONE Newton step (relative error <= 2^-40 ~ 9e-13 per term);
  // RTX_VOIGT_EXACT_RCP adds the second step (full double precision)
#define RTX_LAUNCH(KK)                                                                                        \
  (ext ? (mode == 0 ? launch_tile<KK, 0, false, true>(P, st) : launch_tile<KK, 1, false, true>(P, st))          \
   : ws ? (mode == 2 ? launch_tile<KK, 2, true>(P, st) : mode == 1 ? launch_tile<KK, 1, true>(P, st) : launch_tile<KK, 0, true>(P, st)) \
      : (mode == 2 ? launch_tile<KK, 2, false>(P, st) : mode == 1 ? launch_tile<KK, 1, false>(P, st) : launch_tile<KK, 0, false>(P, st)))
  switch (K) {
    case 8: rc = RTX_LAUNCH(8); break;
    case 4: rc = RTX_LAUNCH(4); break;
    case 2: rc = RTX_LAUNCH(2); break;
    default: rc = RTX_LAUNCH(1); break;
  }
  if (rc) return rc;
  h->launches++;
  if (with_tf) {                                // K = 8 with the tile-level blocks
    rc = ext ? (mode == 0 ? launch_tile<8, 0, false, true, true>(P, st) : launch_tile<8, 1, false, true, true>(P, st))
             : (mode == 0 ? launch_tile<8, 0, false, false, true>(P, st) : launch_tile<8, 1, false, false, true>(P, st));
    if (rc) return rc;
    h->launches++;
  }
  if (with4) {                                  // the K = 4 grid of the same call
    P.n_tiles = n_tiles4;
    P.bins = h->d_bins + 2 * nb;
    K = 4;
    rc = RTX_LAUNCH(4);
    if (rc) return rc;
    h->launches++;
  }
#undef RTX_LAUNCH
  if (P.split > 1) {
    const int64_t n = (int64_t)n_states * nx;
    sum_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.part_out, P.split, n, out_dev);
    RTX_CUDA(cudaGetLastError());
    h->launches++;
  }
  return RTX_OK;
}

extern "C" int rtx_voigt_count(rtx_handle* h, int n_states, const double* T_dev, const double* p_dev,
                               const double* weight_dev, double x_air, double x_self, double wing_abs,
                               double wing_hw, double s_threshold, const double* grid_dev, int64_t nx,
                               uint64_t* counts_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(counts_dev != nullptr, "null counts pointer");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VoigtParams P{};
  int rc = fill_params(h, P, n_states, T_dev, p_dev, weight_dev, x_air, x_self, wing_abs, wing_hw, s_threshold,
                       grid_dev, nx);
  if (rc) return rc;
  double* st_wb = nullptr;
  rc = run_prep(h, P, st, &st_wb);
  if (rc) return rc;
  RTX_CUDA(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(uint64_t), st));
  const int64_t total = P.n_lines * (int64_t)n_states;
  const int64_t blocks = (total + 255) / 256;
  RTX_REQUIRE(blocks <= 2147483647LL, "rtx_voigt_count: too many (line, state) pairs");
  count_kernel<<<(unsigned)blocks, 256, 0, st>>>(P, reinterpret_cast<unsigned long long*>(counts_dev));
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_tips(rtx_handle* h, int n, const double* T_dev, double* q_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  if (h->n_iso <= 0) { rtx_set_error("isotopologue table not set (rtx_set_iso_table)"); return RTX_E_STATE; }
  RTX_REQUIRE(n > 0 && T_dev && q_dev, "bad arguments");
  DeviceGuard guard(h->device);
  const int64_t total = (int64_t)n * h->n_iso;
  tips_kernel<<<(unsigned)((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(n, h->n_iso, T_dev,
                                                                                            h->d_tipsT, h->d_q, q_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}
