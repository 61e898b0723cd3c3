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
#include <cuda.h>
#include <math.h>

#include <algorithm>

#include "rtx_common.cuh"
#include "rtx_math.cuh"

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

// Grouped walk with the sample rows staged in shared memory by TMA bulk copies (the "banded shared-memory stencil" of
// the north star): a producer thread streams the union support of TG adjacent bands, TR rows of 32*SPL spectra per
// stage, through a TS-deep ring of mbarrier-guarded stages (tens of KB in
// flight per CTA, no registers held by loads in flight); eight consumer warps take TR/8 rows of every stage each:
// lanes form the (row, band) weights of the warp's rows, then all lanes -- SPL adjacent spectra each -- accumulate the
// rows into TG x SPL register accumulators.  Every sample row is read from global memory once per group of TG bands.
// The stage is ONE 2-D tensor-map copy (cp.async.bulk.tensor.2d; per-row bulk copies of 256-1024 B were TMA-issue
// bound: 2.6 / 1.7 ms for SPL = 1 / 2 on 12 288 spectra x 80 000 samples, against 1.05 ms now = 0.83 of the HBM copy rate
// on unique bytes).  Requires ns even and Y 16-byte
// aligned (tensor-map stride granularity).
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(s_u32(b)), "r"(parity) : "memory");
  }
}
template <int SPL> struct IlsTmaCfg {
  static constexpr int TR = SPL == 1 ? 64 : 32;      // rows per stage (16 KB stages)
  static constexpr int TS = 4;                       // stages
};
template <int MODE, int TG, int SPL>
__global__ void __launch_bounds__(ILS_THREADS + 32, TG == 16 ? 2 : 3) ils_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                  const double* __restrict__ X, int64_t nx,
                                                                  const double* __restrict__ Y, int64_t ns,
                                                                  int n_bands, const double* __restrict__ centre,
                                                                  const double* __restrict__ width,
                                                                  double* __restrict__ out) {
  constexpr int TR = IlsTmaCfg<SPL>::TR, TS = IlsTmaCfg<SPL>::TS;
  constexpr int NW = ILS_THREADS / 32;              // consumer warps
  constexpr int RW = TR / NW;                       // rows per consumer warp per stage
  constexpr int WPL = (RW * TG + 31) / 32;          // (row, band) weights per lane per stage
  constexpr int NS = 32 * SPL;                      // spectra per CTA
  static_assert(32 % TG == 0 || TG % 32 == 0, "a lane's weights all belong to one band");
  extern __shared__ __align__(128) unsigned char ils_smem[];
  double (*s_y)[TR][NS] = reinterpret_cast<double (*)[TR][NS]>(ils_smem);                    // [TS][TR][NS]
  double (*s_wt)[RW][TG] = reinterpret_cast<double (*)[RW][TG]>(ils_smem + sizeof(double) * TS * TR * NS);   // [NW]
  double (*s_acc)[TG][NS] = reinterpret_cast<double (*)[TG][NS]>(ils_smem);                  // reuses the ring at the end
  __shared__ int64_t s_lo[TG], s_hi[TG];
  __shared__ double s_c[TG], s_isg[TG], s_gn[TG], s_ws[NW][TG];
  __shared__ __align__(8) unsigned long long s_full[TS], s_empty[TS];
  const int b0 = blockIdx.x * TG;
  const int ng = min(TG, n_bands - b0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = (int64_t)blockIdx.y * NS;
  const int nlive = (int)min((int64_t)NS, ns - s0);
  if (threadIdx.x < TG) {
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
  if (threadIdx.x == 0) {
    for (int i = 0; i < TS; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&s_full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&s_empty[i])), "r"(NW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t ulo = nx, uhi = 0;
  for (int g = 0; g < ng; ++g)
    if (s_hi[g] > s_lo[g]) { ulo = min(ulo, s_lo[g]); uhi = max(uhi, s_hi[g]); }
  const int64_t nrows = uhi > ulo ? uhi - ulo : 0;
  const int nit = (int)((nrows + TR - 1) / TR);
  // Neighbouring groups share the samples between their outermost band centres.  Even groups walk their support
  // upwards, odd groups downwards, so both neighbours touch a shared stretch at the same time (both at the start or
  // both at the end of their walks) and the second reader finds it in L2 instead of HBM.
  const bool rev = (blockIdx.x & 1) != 0;

  if (warp == NW) {
    // ---- producer: ONE 2-D tensor copy (TR rows x NS spectra) per stage; out-of-range rows / spectra are zero-filled
    if (lane == 0) {
      for (int it = 0; it < nit; ++it) {
        const int slot = it % TS;
        if (it >= TS) mb_wait(&s_empty[slot], ((it / TS) - 1) & 1);
        const int c0 = (int)s0, c1 = (int)(ulo + (int64_t)(rev ? nit - 1 - it : it) * TR);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&s_full[slot])),
                     "r"((uint32_t)(TR * NS * 8)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         s_u32(&s_y[slot][0][0])),
                     "l"(&tmap), "r"(c0), "r"(c1), "r"(s_u32(&s_full[slot]))
                     : "memory");
      }
    }
    return;
  }

  // ---- consumer warps.  Row indices are kept relative to the union support (32-bit); a lane's (row, band) pairs all
  // belong to band lane % TG, whose support, centre and width sit in registers.
  double acc[TG][SPL], wpart = 0.0;
#pragma unroll
  for (int g = 0; g < TG; ++g)
#pragma unroll
    for (int q = 0; q < SPL; ++q) acc[g][q] = 0.0;
  const int gl = lane % TG;
  const int rlo = (gl < ng && s_hi[gl] > s_lo[gl]) ? (int)(s_lo[gl] - ulo) : 0;
  const int rhi = (gl < ng && s_hi[gl] > s_lo[gl]) ? (int)(s_hi[gl] - ulo) : 0;
  const double cg = s_c[gl], isg = s_isg[gl], gng = s_gn[gl];
  const int nr = (int)nrows;
  const double* __restrict__ Xu = X + ulo;
  for (int it = 0; it < nit; ++it) {
    const int slot = it % TS;
    const int r0 = (rev ? nit - 1 - it : it) * TR + warp * RW;      // this warp's first row of the stage
#pragma unroll
    for (int q = 0; q < WPL; ++q) {
      const int pr = lane + 32 * q, j = pr / TG;
      if (RW * TG >= 32 || pr < RW * TG) {
        const int ri = r0 + j;
        double w = 0.0;
        if (ri >= rlo && ri < rhi) {
          const double x = Xu[ri];
          if (MODE == 0) {
            w = 1.0 - fabs(x - cg) * isg;                            // tri(), radiative_transfer.py:1236-1239
            if (w < 0.0) w = 0.0;
          } else {
            const double u = (x - cg) * isg;
            w = exp(-0.5 * (u * u)) / gng;                           // g(), ILS_MAKO.py:24
          }
        }
        s_wt[warp][j][gl] = w;
        wpart += w;
      }
    }
    // bands of the group that have any weight on this warp's rows (warp-uniform)
    unsigned act = __ballot_sync(0xffffffffu, r0 < rhi && r0 + RW > rlo);
    if (TG < 32) {
      unsigned m = 0;
#pragma unroll
      for (int sft = 0; sft < 32; sft += TG) m |= act >> sft;
      act = m & ((1u << TG) - 1u);
    }
    __syncwarp();
    mb_wait(&s_full[slot], (it / TS) & 1);
    if (act) {
      const int nv = min(RW, nr - r0);                              // rows past the support may hold anything: skip them
      double y[RW][SPL];
#pragma unroll
      for (int j = 0; j < RW; ++j)
#pragma unroll
        for (int q = 0; q < SPL; ++q) y[j][q] = (j < nv) ? s_y[slot][warp * RW + j][lane * SPL + q] : 0.0;
#pragma unroll
      for (int g = 0; g < TG; ++g) {
        if (act & (1u << g)) {
#pragma unroll
          for (int j = 0; j < RW; ++j) {
            const double w = s_wt[warp][j][g];
#pragma unroll
            for (int q = 0; q < SPL; ++q) acc[g][q] = fma(w, y[j][q], acc[g][q]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&s_empty[slot])) : "memory");
  }
  // weight sums: a lane's weights all belong to band lane % TG; reduce over the lanes that share a band
  for (int o = 16; o >= TG; o >>= 1) wpart += __shfl_xor_sync(0xffffffffu, wpart, o);
  // the ring is free once every consumer warp has passed its last stage
  asm volatile("bar.sync 1, %0;" ::"n"(ILS_THREADS));
#pragma unroll
  for (int g = 0; g < TG; ++g)
#pragma unroll
    for (int q = 0; q < SPL; ++q) s_acc[warp][g][lane * SPL + q] = acc[g][q];
  if (lane < TG) s_ws[warp][lane] = wpart;
  asm volatile("bar.sync 1, %0;" ::"n"(ILS_THREADS));
  for (int g = warp; g < ng; g += NW) {
    for (int i = lane; i < nlive; i += 32) {
      double a = 0.0, wt = 0.0;
#pragma unroll
      for (int q = 0; q < NW; ++q) { a += s_acc[q][g][i]; wt += s_ws[q][g]; }
      out[(int64_t)(b0 + g) * ns + s0 + i] = a / wt;                // Y_out = sum(ILS*Y)/N
    }
  }
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*rtx_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static rtx_tmap_encode_fn tmap_encoder() {
  static rtx_tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<rtx_tmap_encode_fn>(p);
    cudaGetLastError();
  }
  return fn;
}
// Y [nx][ns] doubles as a 2-D tensor (inner dimension = spectra), box = TR rows x NS spectra
static bool make_y_tmap(CUtensorMap* tm, const double* Y, int64_t nx, int64_t ns, int box_rows, int box_cols) {
  rtx_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)ns, (cuuint64_t)nx};
  const cuuint64_t gstride[1] = {(cuuint64_t)ns * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(Y), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// returns 1 when the tensor map could not be built (caller falls back to the unstaged walk)
template <int MODE, int TG, int SPL>
int launch_ils_tma(const double* X, int64_t nx, const double* Y, int64_t ns, int n_bands, const double* c, const double* w,
                   double* out, cudaStream_t st) {
  constexpr int TR = IlsTmaCfg<SPL>::TR, TS = IlsTmaCfg<SPL>::TS, NW = ILS_THREADS / 32;
  CUtensorMap tm;
  if (!make_y_tmap(&tm, Y, nx, ns, TR, 32 * SPL)) return 1;
  const size_t smem = sizeof(double) * (TS * TR * 32 * SPL + NW * (TR / NW) * TG);
  static_assert(NW * TG * 32 * SPL <= TS * TR * 32 * SPL, "final sums fit the ring");
  RTX_CUDA(cudaFuncSetAttribute(ils_tma_kernel<MODE, TG, SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n_bands + TG - 1) / TG), (unsigned)((ns + 32 * SPL - 1) / (32 * SPL)));
  ils_tma_kernel<MODE, TG, SPL><<<grid, ILS_THREADS + 32, smem, st>>>(tm, X, nx, Y, ns, n_bands, c, w, out);
  return RTX_OK;
}

// thread per (x, a, t): one Planck evaluation shared by all emissivities.  OutT = float is the float32 product of the
// reference's production script (Compute_LWIR_Apparent_Radiance.py:9-25 casts every input to float32): same FP64
// arithmetic, rounded once at the store, half the bytes of this purely write-bound kernel.
template <typename OutT>
__global__ void apparent_kernel(const double* __restrict__ X, int64_t nx, const double* __restrict__ emis,
                                int64_t nE, const double* __restrict__ Ts, int64_t nA,
                                const double* __restrict__ tau, const double* __restrict__ la,
                                const double* __restrict__ ld, const double* __restrict__ dT, int64_t nT,
                                OutT* __restrict__ L, OutT* __restrict__ Ls) {
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
    L[o] = (OutT)(tu * ls + up);                        // L = tau*Ls + La
    if (Ls) Ls[o] = (OutT)ls;
  }
}

// Per-pixel mixed-material radiance (LWIR_HSI_Generator.py:148-166), write-bound: 8 (or 4) bytes per (pixel, band)
// against ~18 FP64 operations, so the arithmetic has to stay out of the way of the stores.  A persistent CTA takes
// groups of HSI_PB pixels; a thread owns one band (stride blockDim.x for wider axes) and keeps the band's Planck
// factors c1 X^3 and c2 X in registers; the per-pixel 1/T is formed once per pixel (not per output) into shared
// memory; exp is the table-driven one of the layer integration (no libm call, no FP64 divide per output), and the
// 64-bit div/mod of a flat (pixel, band) index is gone.  Stores are 256 contiguous bytes per warp.
#ifndef RTX_HSI_UNROLL
#define RTX_HSI_UNROLL 4
#endif
#ifndef RTX_HSI_PB
#define RTX_HSI_PB 64
#endif
constexpr int HSI_THREADS = 128, HSI_PB = RTX_HSI_PB, HSI_MAXMIX = 4, HSI_UNROLL = RTX_HSI_UNROLL;
// NMIX > 0: materials per pixel known at compile time (per-pixel offsets and fractions staged in shared memory, 32-bit
// offsets); NMIX = 0: any n_mix, read from global memory.
template <typename OutT, int NMIX>
__global__ void __launch_bounds__(HSI_THREADS) hsi_kernel(const double* __restrict__ X, int nx, const double* __restrict__ emis,
                                                          const double* __restrict__ tau, const double* __restrict__ la,
                                                          const double* __restrict__ ld, int64_t n_pix,
                                                          const int32_t* __restrict__ atm, const double* __restrict__ Tp,
                                                          int n_mix, const int32_t* __restrict__ mat,
                                                          const double* __restrict__ frac, OutT* __restrict__ L) {
  __shared__ double s_tab[256], s_invT[HSI_PB], s_frac[HSI_PB][NMIX > 0 ? NMIX : 1];
  __shared__ int s_atm[HSI_PB], s_mat[HSI_PB][NMIX > 0 ? NMIX : 1];
  exp_tab_init(s_tab);
  const int64_t ngroups = (n_pix + HSI_PB - 1) / HSI_PB;
  for (int64_t pg = blockIdx.x; pg < ngroups; pg += gridDim.x) {
    const int64_t k0 = pg * HSI_PB;
    const int np = (int)min((int64_t)HSI_PB, n_pix - k0);
    __syncthreads();                                    // the previous group is done with the staged values (and the table is filled)
    if (threadIdx.x < np) {
      const int64_t k = k0 + threadIdx.x;
      s_invT[threadIdx.x] = 1.0 / Tp[k];
      s_atm[threadIdx.x] = atm[k] * nx;                 // element offsets fit 32 bits (checked by the launcher)
      if (NMIX > 0) {
#pragma unroll
        for (int m = 0; m < NMIX; ++m) { s_mat[threadIdx.x][m] = mat[k * NMIX + m] * nx; s_frac[threadIdx.x][m] = frac[k * NMIX + m]; }
      }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < nx; x += HSI_THREADS) {
      const double X100 = X[x] * 100.0;
      const double c1X3 = RTX_RT_C1 * (X100 * X100 * X100), c2X = RTX_RT_C2 * X100;
      const double* __restrict__ emis_x = emis + x;
      const double* __restrict__ tau_x = tau + x;
      const double* __restrict__ la_x = la + x;
      const double* __restrict__ ld_x = ld + x;
      OutT* __restrict__ out = L + k0 * nx + x;
#pragma unroll(HSI_UNROLL)
      for (int p = 0; p < np; ++p) {
        double em = 0.0;
        if (NMIX > 0) {
#pragma unroll
          for (int m = 0; m < NMIX; ++m) em = fma(s_frac[p][m], emis_x[s_mat[p][m]], em);
        } else {
          const int64_t k = k0 + p;
          for (int m = 0; m < n_mix; ++m) em += frac[k * n_mix + m] * emis_x[(int64_t)mat[k * n_mix + m] * nx];
        }
        const int oa = s_atm[p];
        const double B = planck_fast(c1X3, c2X, s_invT[p], s_tab);
        const double dn = ld_x[oa];
        const double ls = fma(em, B - dn, dn);          // em*B + (1-em)*Ld
        out[(int64_t)p * nx] = (OutT)fma(tau_x[oa], ls, la_x[oa]);
      }
    }
  }
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

// Cubic-spline interpolation (scipy interp1d(kind="cubic"): the C2 cubic spline through all knots with NOT-A-KNOT end
// conditions).  The spline's second derivatives decay like (2-sqrt(3))^k away from a knot, so each output point solves
// the tridiagonal system only on a window of HW knots either side of its interval (truncation < 1e-20 for HW = 40).
// A window end that lies inside the data uses the natural condition (M = 0 there, exact to rounding); a window end
// that IS an end of the data uses the reference's not-a-knot condition (third derivative continuous across the second
// knot: M_0 = M_1 (1 + h0/h1) - M_2 h0/h1, eliminated into the first row), so points near the ends of the data -- and
// beyond them (fill_value="extrapolate" continues the end polynomial) -- agree with scipy as well.
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
  const int nk = (int)(k1 - k0 + 1);                     // knots in the window, 4 <= nk <= 2*HW+2
  const bool left_end = (k0 == 0), right_end = (k1 == n - 1);
  double cp[2 * SPL_HW + 2], dp[2 * SPL_HW + 2];         // Thomas algorithm workspace; dp ends up holding M_k
  cp[0] = 0.0; dp[0] = 0.0;
  for (int k = 1; k < nk - 1; ++k) {
    const int64_t g = k0 + k;
    const double h0 = X[g] - X[g - 1], h1 = X[g + 1] - X[g];
    const double rhs = 6.0 * ((Y[(g + 1) * m + c] - Y[g * m + c]) / h1 - (Y[g * m + c] - Y[(g - 1) * m + c]) / h0);
    double a = h0, b = 2.0 * (h0 + h1), cc = h1;
    if (k == 1 && left_end) { a = 0.0; b = 3.0 * h0 + 2.0 * h1 + h0 * h0 / h1; cc = h1 - h0 * h0 / h1; }
    if (k == nk - 2 && right_end) { a = h0 - h1 * h1 / h0; b = 2.0 * h0 + 3.0 * h1 + h1 * h1 / h0; cc = 0.0; }   // nk >= 4: never row 1
    const double den = b - a * cp[k - 1];
    cp[k] = cc / den;
    dp[k] = (rhs - a * dp[k - 1]) / den;
  }
  // back substitution: dp[k] <- M_k  (M at a natural window end is 0)
  dp[nk - 1] = 0.0;
  for (int k = nk - 2; k >= 1; --k) dp[k] = dp[k] - cp[k] * dp[k + 1];
  if (right_end) {
    const double r = (X[k1] - X[k1 - 1]) / (X[k1 - 1] - X[k1 - 2]);
    dp[nk - 1] = dp[nk - 2] * (1.0 + r) - dp[nk - 3] * r;
  }
  if (left_end) {
    const double r = (X[1] - X[0]) / (X[2] - X[1]);
    dp[0] = dp[1] * (1.0 + r) - dp[2] * r;
  }
  const int ki = (int)(i - k0);
  const double Mi = dp[ki], Mi1 = dp[ki + 1];
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
  RTX_REQUIRE((mode & ~0x1F01) == 0, "rtx_ils_apply: mode must be 0 (triangle) or 1 (Gaussian) [| 0x100: band-per-CTA walk, "
                                    "0x200: grouped walk without TMA staging, 0x400 / 0x800: 2 / 1 spectra per lane]");
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
  const bool aligned = (ns % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y_dev) & 15u) == 0);
  if (wide && aligned && groups * sblocks >= 2LL * h->sm_count && !(mode & 0x300)) {
    // TMA-staged grouped walk: two spectra per lane when that still gives two full waves of three CTAs per SM, else
    // one; bits 0x400 / 0x800 force 2 / 1 spectra per lane, 0x1000 groups of 16 bands (experiment hooks)
    int spl = (groups * ((ns + 63) / 64) >= 6LL * h->sm_count) ? 2 : 1;
    if (mode & 0x400) spl = 2;
    if (mode & 0x800) spl = 1;
    int rc;
#define RTX_ILS_TMA(M, S) ((mode & 0x1000) ? launch_ils_tma<M, 16, S>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev, st) \
                                           : launch_ils_tma<M, 8, S>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev, st))
    if (mode & 1) rc = spl == 2 ? RTX_ILS_TMA(1, 2) : RTX_ILS_TMA(1, 1);
    else rc = spl == 2 ? RTX_ILS_TMA(0, 2) : RTX_ILS_TMA(0, 1);
#undef RTX_ILS_TMA
    if (rc < 0) return rc;
    if (rc == 1) {                        // no tensor-map encoder in this driver: unstaged grouped walk
      dim3 grid((unsigned)groups, (unsigned)sblocks);
      if ((mode & 1) == 0) ils_group_kernel<0><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev);
      else ils_group_kernel<1><<<grid, ILS_THREADS, 0, st>>>(X_dev, nx, Y_dev, ns, n_bands, centre_dev, width_dev, out_dev);
    }
  } else if (wide && groups * sblocks >= 2LL * h->sm_count && !(mode & 0x100)) {
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

template <typename OutT>
static int apparent_launch(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                           const double* Ts_dev, int64_t n_atm, const double* tau_dev, const double* la_dev,
                           const double* ld_dev, const double* dT_dev, int64_t n_dT, OutT* L_dev, OutT* ls_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && n_emis > 0 && n_atm > 0, "rtx_apparent_radiance: bad sizes");
  RTX_REQUIRE(X_dev && emis_dev && Ts_dev && tau_dev && la_dev && ld_dev && L_dev, "rtx_apparent_radiance: null pointer");
  if (!dT_dev) n_dT = 1;
  RTX_REQUIRE(n_dT > 0, "rtx_apparent_radiance: n_dT must be positive");
  DeviceGuard guard(h->device);
  const int64_t total = nx * n_atm * n_dT;
  const int64_t blocks = (total + 255) / 256;
  RTX_REQUIRE(blocks <= 2147483647LL, "rtx_apparent_radiance: problem too large for one launch");
  apparent_kernel<OutT><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X_dev, nx, emis_dev, n_emis, Ts_dev, n_atm, tau_dev, la_dev, ld_dev, dT_dev, n_dT, L_dev, ls_dev);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_apparent_radiance(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev,
                                     int64_t n_emis, const double* Ts_dev, int64_t n_atm, const double* tau_dev,
                                     const double* la_dev, const double* ld_dev, const double* dT_dev, int64_t n_dT,
                                     double* L_dev, double* ls_dev, void* stream) {
  return apparent_launch<double>(h, X_dev, nx, emis_dev, n_emis, Ts_dev, n_atm, tau_dev, la_dev, ld_dev, dT_dev, n_dT, L_dev,
                                 ls_dev, stream);
}
extern "C" int rtx_apparent_radiance_f32(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev,
                                         int64_t n_emis, const double* Ts_dev, int64_t n_atm, const double* tau_dev,
                                         const double* la_dev, const double* ld_dev, const double* dT_dev, int64_t n_dT,
                                         float* L_dev, float* ls_dev, void* stream) {
  return apparent_launch<float>(h, X_dev, nx, emis_dev, n_emis, Ts_dev, n_atm, tau_dev, la_dev, ld_dev, dT_dev, n_dT, L_dev,
                                ls_dev, stream);
}

template <typename OutT>
static int hsi_launch(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                      const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm, int64_t n_pix,
                      const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix, const int32_t* mat_idx_dev,
                      const double* frac_dev, OutT* L_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(nx > 0 && nx <= 2147483647LL && n_emis > 0 && n_atm > 0 && n_pix > 0 && n_mix > 0, "rtx_hsi_pixels: bad sizes");
  RTX_REQUIRE(X_dev && emis_dev && tau_dev && la_dev && ld_dev && atm_idx_dev && T_pix_dev && mat_idx_dev &&
                  frac_dev && L_dev, "rtx_hsi_pixels: null pointer");
  DeviceGuard guard(h->device);
  RTX_REQUIRE(n_emis * nx <= 2147483647LL && n_atm * nx <= 2147483647LL, "rtx_hsi_pixels: emissivity / atmosphere tables too large");
  const int64_t ngroups = (n_pix + HSI_PB - 1) / HSI_PB;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define RTX_HSI(NM)                                                                                                   \
  do {                                                                                                                \
    int per_sm = 0;                                                                                                   \
    RTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hsi_kernel<OutT, NM>, HSI_THREADS, 0));            \
    const int64_t blocks = std::min<int64_t>(ngroups, (int64_t)h->sm_count * std::max(per_sm, 1));   /* persistent CTAs */ \
    hsi_kernel<OutT, NM><<<(unsigned)blocks, HSI_THREADS, 0, st>>>(X_dev, (int)nx, emis_dev, tau_dev, la_dev, ld_dev, n_pix, \
                                                                   atm_idx_dev, T_pix_dev, n_mix, mat_idx_dev, frac_dev, L_dev); \
  } while (0)
  switch (n_mix) {
    case 1: RTX_HSI(1); break;
    case 2: RTX_HSI(2); break;
    case 3: RTX_HSI(3); break;
    case 4: RTX_HSI(4); break;
    default: RTX_HSI(0); break;
  }
#undef RTX_HSI
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_hsi_pixels(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                              const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm,
                              int64_t n_pix, const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix,
                              const int32_t* mat_idx_dev, const double* frac_dev, double* L_dev, void* stream) {
  return hsi_launch<double>(h, X_dev, nx, emis_dev, n_emis, tau_dev, la_dev, ld_dev, n_atm, n_pix, atm_idx_dev, T_pix_dev, n_mix,
                            mat_idx_dev, frac_dev, L_dev, stream);
}
extern "C" int rtx_hsi_pixels_f32(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                                  const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm,
                                  int64_t n_pix, const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix,
                                  const int32_t* mat_idx_dev, const double* frac_dev, float* L_dev, void* stream) {
  return hsi_launch<float>(h, X_dev, nx, emis_dev, n_emis, tau_dev, la_dev, ld_dev, n_atm, n_pix, atm_idx_dev, T_pix_dev, n_mix,
                           mat_idx_dev, frac_dev, L_dev, stream);
}
