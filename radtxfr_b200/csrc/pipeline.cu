// Entry points a non-Python host needs to run the whole path through the C ABI:
//   rtx_set_iso_molecules / rtx_layer_weights   column amounts of compute_OD's optical depth
//   rtx_tud_from_lines                          atmosphere table in -> (tau, Lu, Ld) out, batched over atmospheres:
//                                               the loop body of Generate_LWIR_TUD.py:117-122 around compute_TUD
//                                               (radiative_transfer.py:274-392)
//   rtx_mako_bands                              band centres / widths of both MAKO line shapes
//                                               (radiative_transfer.py:1226-1241, ILS_MAKO.py:20-22), host only
#include <math.h>

#include <algorithm>
#include <vector>

#include "rtx_common.cuh"
#include "mako_table.inc"

namespace {

constexpr int MAX_MOL = 64;
struct WeightParams {
  int n_states, n_iso, n_mol;
  const double *T, *P_pa, *PL_km, *mf;
  const int32_t* iso_mol;
  const double* iso_factor;
  int32_t mf_id[MAX_MOL];
  double *p_atm, *weight;
};

// One thread per (state, isotopologue): w = (MF*1e-6) * volumeConcentration(p, T) * (PL*1e5) * factor, in the operation
// order of the Python shim (radiative_transfer.layer_weights), so that both routes give the same bits.
__global__ void layer_weights_kernel(const WeightParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)P.n_states * P.n_iso) return;
  const int s = (int)(i / P.n_iso), j = (int)(i % P.n_iso);
  const double T = P.T[s];
  const double p = __ddiv_rn(P.P_pa[s], 101325.0);
  if (j == 0) P.p_atm[s] = p;
  int k = -1;
  const int mol = P.iso_mol[j];
  for (int q = 0; q < P.n_mol; ++q)
    if (P.mf_id[q] == mol) { k = q; break; }                                     // ids.index(mol): first match
  double w = 0.0;
  if (k >= 0) {
    const double nconc = __ddiv_rn(__ddiv_rn(p, 9.869233e-7), __dmul_rn(RTX_CBOLTZ, T));   // misc/hapi.py:10163-10164
    const double a = __dmul_rn(P.mf[(int64_t)s * P.n_mol + k], 1e-6);
    const double b = __dmul_rn(a, nconc);
    const double c = __dmul_rn(P.PL_km[s], 1e5);
    w = __dmul_rn(__dmul_rn(b, c), P.iso_factor[j]);
  }
  P.weight[i] = w;
}

int ensure_pipe(rtx_handle* h, size_t doubles) {
  if (doubles > h->pipe_cap) {
    if (h->d_pipe) cudaFree(h->d_pipe);
    h->d_pipe = nullptr; h->pipe_cap = 0;
    RTX_CUDA(cudaMalloc(&h->d_pipe, doubles * sizeof(double)));
    h->pipe_cap = doubles;
  }
  return RTX_OK;
}

}  // namespace

extern "C" int rtx_set_iso_molecules(rtx_handle* h, int n_iso, const int32_t* molec_id, const double* factor) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n_iso > 0 && n_iso == h->n_iso && molec_id && factor,
              "rtx_set_iso_molecules: set the isotopologue table first and pass one entry per isotopologue");
  DeviceGuard guard(h->device);
  if (h->d_iso_mol) { cudaFree(h->d_iso_mol); h->d_iso_mol = nullptr; }
  if (h->d_iso_factor) { cudaFree(h->d_iso_factor); h->d_iso_factor = nullptr; }
  RTX_CUDA(cudaMalloc(&h->d_iso_mol, n_iso * sizeof(int32_t)));
  RTX_CUDA(cudaMalloc(&h->d_iso_factor, n_iso * sizeof(double)));
  RTX_CUDA(cudaMemcpy(h->d_iso_mol, molec_id, n_iso * sizeof(int32_t), cudaMemcpyHostToDevice));
  RTX_CUDA(cudaMemcpy(h->d_iso_factor, factor, n_iso * sizeof(double), cudaMemcpyHostToDevice));
  return RTX_OK;
}

extern "C" int rtx_layer_weights(rtx_handle* h, int n_states, const double* T_dev, const double* P_pa_dev,
                                 const double* PL_km_dev, const double* mf_ppmv_dev, int n_mol, const int32_t* mf_id_host,
                                 double* p_atm_dev, double* weight_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n_states > 0 && n_mol > 0 && n_mol <= MAX_MOL, "rtx_layer_weights: bad sizes (at most 64 molecules)");
  RTX_REQUIRE(T_dev && P_pa_dev && PL_km_dev && mf_ppmv_dev && mf_id_host && p_atm_dev && weight_dev, "rtx_layer_weights: null pointer");
  if (!h->d_iso_mol) { rtx_set_error("rtx_layer_weights: call rtx_set_iso_molecules first"); return RTX_E_STATE; }
  DeviceGuard guard(h->device);
  WeightParams P{};
  P.n_states = n_states; P.n_iso = h->n_iso; P.n_mol = n_mol;
  P.T = T_dev; P.P_pa = P_pa_dev; P.PL_km = PL_km_dev; P.mf = mf_ppmv_dev;
  P.iso_mol = h->d_iso_mol; P.iso_factor = h->d_iso_factor;
  for (int q = 0; q < n_mol; ++q) P.mf_id[q] = mf_id_host[q];
  P.p_atm = p_atm_dev; P.weight = weight_dev;
  const int64_t total = (int64_t)n_states * h->n_iso;
  layer_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  RTX_CUDA(cudaGetLastError());
  h->launches++;
  return RTX_OK;
}

extern "C" int rtx_tud_from_lines(rtx_handle* h, int n_atm, int n_layers, const double* T_dev, const double* P_pa_dev,
                                  const double* PL_km_dev, const double* mf_ppmv_dev, int n_mol, const int32_t* mf_id_host,
                                  double x_air, double x_self, double wing_abs, double wing_hw, double s_threshold,
                                  int voigt_flags, const double* grid_dev, int64_t nx, const double* Z_host,
                                  const double* alt_host, int n_alt, const double* mu_host, int n_mu, int n_angle,
                                  int return_od, double* tau_dev, double* lu_dev, double* ld_dev, void* stream) {
  RTX_REQUIRE(h != nullptr, "null handle");
  RTX_REQUIRE(n_atm > 0 && n_layers > 0 && n_layers <= RTX_MAX_LAYERS && nx > 0, "rtx_tud_from_lines: bad sizes");
  RTX_REQUIRE(tau_dev && lu_dev && ld_dev, "rtx_tud_from_lines: null output pointer");
  DeviceGuard guard(h->device);
  // atmospheres per pass: at most 4 GB of optical depth (layer-major [atm][layer][nx]) in the handle's scratch
  const int64_t per_atm = (int64_t)n_layers * nx;
  int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(n_atm, (int64_t)(4e9 / 8.0) / per_atm));
  chunk = std::min(chunk, 32768);
  const size_t n_st = (size_t)chunk * n_layers;
  int rc = ensure_pipe(h, n_st * (size_t)nx + n_st * (1 + (size_t)h->n_iso));
  if (rc) return rc;
  double* od = h->d_pipe;
  double* p_atm = od + n_st * (size_t)nx;
  double* wgt = p_atm + n_st;
  const int64_t out_stride = nx * n_alt * n_mu;
  for (int a0 = 0; a0 < n_atm; a0 += chunk) {
    const int na = std::min(chunk, n_atm - a0);
    const int ns = na * n_layers;
    const int64_t o = (int64_t)a0 * n_layers;
    if ((rc = rtx_layer_weights(h, ns, T_dev + o, P_pa_dev + o, PL_km_dev + o, mf_ppmv_dev + o * n_mol, n_mol, mf_id_host,
                                p_atm, wgt, stream)))
      return rc;
    // a batch never splits the lines of a tile over CTAs: the result must not depend on the chunking below
    if ((rc = rtx_voigt_sum(h, ns, T_dev + o, p_atm, wgt, x_air, x_self, wing_abs, wing_hw, s_threshold, grid_dev, nx, od,
                            voigt_flags | RTX_VOIGT_NOSPLIT, stream)))
      return rc;
    if ((rc = rtx_tud(h, na, n_layers, od, T_dev + o, grid_dev, nx, Z_host, alt_host, n_alt, mu_host, n_mu, n_angle,
                      return_od, tau_dev + (int64_t)a0 * out_stride, lu_dev + (int64_t)a0 * out_stride,
                      ld_dev + (int64_t)a0 * nx, stream)))
      return rc;
  }
  return RTX_OK;
}

// np.interp(x, xp, fp) for ascending xp
static double interp1(double x, const double* xp, const double* fp, int n) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[n - 1]) return fp[n - 1];
  int j = (int)(std::upper_bound(xp, xp + n, x) - xp) - 1;
  j = std::max(0, std::min(n - 2, j));
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  return slope * (x - xp[j]) + fp[j];
}

extern "C" int rtx_mako_bands(int mode, double xmin, double xmax, double res_factor, double fwhm_sf, double shift,
                              double scale, int capacity, double* X_out, double* centre, double* width, int* n_out) {
  RTX_REQUIRE(mode == 0 || mode == 1, "rtx_mako_bands: mode must be 0 (triangle) or 1 (Gaussian)");
  RTX_REQUIRE(n_out != nullptr, "rtx_mako_bands: null n_out");
  std::vector<double> um(RTX_MAKO_UM, RTX_MAKO_UM + 128);
  if (mode == 0 && res_factor > 0.0) {                       // radiative_transfer.py:1226-1229
    const int n1 = (int)(128 * res_factor);
    RTX_REQUIRE(n1 >= 1, "rtx_mako_bands: res_factor too small");
    std::vector<double> x0(128), out(n1);
    for (int i = 0; i < 128; ++i) x0[i] = 0.0 + i * (1.0 / 127.0);           // np.linspace(0, 1, 128)
    x0[127] = 1.0;
    for (int i = 0; i < n1; ++i) {
      double x1 = n1 > 1 ? 0.0 + i * (1.0 / (n1 - 1)) : 0.0;
      if (i == n1 - 1 && n1 > 1) x1 = 1.0;
      out[i] = interp1(x1, x0.data(), um.data(), 128);
    }
    um.swap(out);
  }
  std::vector<double> X(um.size());
  for (size_t i = 0; i < um.size(); ++i) X[i] = 10000.0 / um[i];
  std::sort(X.begin(), X.end());                            // X_out = np.sort(10000.0 / X_out)
  if (mode == 0) {                                          // clip to (X.min, X.max), :1232-1233
    std::vector<double> kept;
    for (double v : X) if (v > xmin && v < xmax) kept.push_back(v);
    X.swap(kept);
  }
  const int n = (int)X.size();
  *n_out = n;
  if (capacity <= 0 || !X_out || !centre || !width) return RTX_OK;      // count only
  RTX_REQUIRE(capacity >= n, "rtx_mako_bands: capacity too small");
  for (int i = 0; i < n; ++i) {
    // np.gradient: central differences inside, one-sided at the ends (uneven spacing formula for the interior)
    double g;
    if (n == 1) g = NAN;
    else if (i == 0) g = X[1] - X[0];
    else if (i == n - 1) g = X[n - 1] - X[n - 2];
    else g = (X[i + 1] - X[i - 1]) / 2.0;
    X_out[i] = X[i];
    if (mode == 0) { centre[i] = scale * X[i] + shift; width[i] = fwhm_sf * fabs(g) * 1.6; }
    else { centre[i] = X[i]; width[i] = fabs(g); }
  }
  return RTX_OK;
}
