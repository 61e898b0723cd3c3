/*
 * radtxfr_b200 -- C ABI of the B200-native line-by-line radiative-transfer hot path.
 *
 * The reference (westi024/RadTxfr) is pure Python and has no FFI of its own; its boundary for this
 * path is the Python function surface.  Each entry point below names the reference
 * code it replaces.  The Python shim in radtxfr_b200/ binds these with ctypes and mirrors the
 * reference's function names and arguments; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success and a negative RTX_E_* code on failure;
 *     rtx_last_error() returns a thread-local message for the last failure in this thread;
 *   - "dev" pointers are device memory owned by the caller (for example torch tensor storage),
 *     "host" pointers are ordinary host memory; all arrays are dense, C order, double unless noted;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); work is
 *     enqueued on it and the call returns without synchronising unless stated otherwise;
 *   - no hidden global state beyond the opaque handle; a handle is bound to one device and may be
 *     used from one thread at a time.
 */
#ifndef RADTXFR_B200_H
#define RADTXFR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTX_OK 0
#define RTX_E_INVALID -1   /* bad argument */
#define RTX_E_CUDA -2      /* CUDA runtime error, see rtx_last_error() */
#define RTX_E_STATE -3     /* call order (e.g. no lines uploaded) */
#define RTX_E_RANGE -4     /* value outside the reference's domain (e.g. TIPS 70..3000 K) */

#define RTX_TIPS_NPT 119   /* TIPS-2011 temperature nodes, misc/hapi.py:5401-5415 */
#define RTX_MAX_ANGLES 64  /* compute_TUD N_angle limit of this build (reference default 30) */
#define RTX_MAX_LAYERS 256 /* layers per atmosphere (reference table: 66) */
#define RTX_MAX_ALT 32     /* sensor altitudes per call */
#define RTX_MAX_MU 8       /* view angles per call */
#define RTX_MAX_EXTRA_DILUENTS 4 /* broadening species besides air and self per line table */

typedef struct rtx_handle rtx_handle;

const char* rtx_last_error(void);
int rtx_version(void);
/* Number of SMs / device name of the handle's device (for grid sizing reports). */
int rtx_device_info(rtx_handle* h, int* sm_count, char* name, int name_len);

int rtx_create(int device, rtx_handle** out);
int rtx_destroy(rtx_handle* h);

/* ---- plain device-memory helpers so that a caller without torch can drive the library ---- */
int rtx_malloc(rtx_handle* h, uint64_t bytes, void** dev);
int rtx_free(rtx_handle* h, void* dev);
int rtx_memcpy_h2d(rtx_handle* h, void* dev, const void* host, uint64_t bytes, void* stream);
int rtx_memcpy_d2h(rtx_handle* h, void* host, const void* dev, uint64_t bytes, void* stream);
int rtx_stream_sync(rtx_handle* h, void* stream);

/* ---- subsystem 1: ingest ------------------------------------------------------------------
 * Isotopologue table: replaces the ISO / TIPS_ISO_HASH lookups of misc/hapi.py:3372-3497,
 * 5088-5124, 9568-9582.  mass [g/mol], q_table[n_iso][RTX_TIPS_NPT] (float32 values promoted to
 * double) on the temperature nodes tips_T[RTX_TIPS_NPT].  Host pointers, copied. */
int rtx_set_iso_table(rtx_handle* h, int n_iso, const double* mass, const double* tips_T,
                      const double* q_table);

/* Line table: replaces hapi's LOCAL_TABLE_CACHE[table]['data'] dict-of-lists (misc/hapi.py:438-463;
 * column meaning :497-517).  Host pointers, n rows in any order; the library sorts by line centre,
 * precomputes the reference-temperature part of S(T) (misc/hapi.py:10173), pads and uploads SoA
 * device buffers.  iso_idx[i] indexes the isotopologue table.  Optional columns may be NULL:
 * gamma_self, n_self (NULL or 0 -> n_air, misc/hapi.py:11104-11109), delta_self, deltap_air,
 * deltap_self (NULL -> 0, :11097-11125). */
int rtx_lines_upload(rtx_handle* h, int64_t n, const double* nu, const double* sw, const double* elower,
                     const double* gamma_air, const double* gamma_self, const double* n_air,
                     const double* n_self, const double* delta_air, const double* deltap_air,
                     const double* delta_self, const double* deltap_self, const int32_t* iso_idx);
int rtx_lines_count(rtx_handle* h, int64_t* n);

/* Optional columns of the table uploaded last, host pointers in that table's ORIGINAL row order, copied; NULL = absent.
 * rtx_lines_set_sd: the speed-dependence columns SD_air / SD_self read by absorptionCoefficient_SDVoigt
 * (misc/hapi.py:10883-10890).
 * rtx_lines_set_diluent: one more broadening species in slot 0..RTX_MAX_EXTRA_DILUENTS-1 -- the columns gamma_<sp>,
 * n_<sp> (NULL -> n_air, misc/hapi.py:11101-11109), delta_<sp>, deltap_<sp>, SD_<sp> that hapi's Diluent={'<sp>': x}
 * loop reads (misc/hapi.py:11090-11128).
 * rtx_voigt_set_mix: the fractions x of the extra species for the following rtx_voigt_sum / rtx_voigt_count calls on
 * this handle (default: all zero). */
int rtx_lines_set_sd(rtx_handle* h, const double* sd_air, const double* sd_self);
int rtx_lines_set_diluent(rtx_handle* h, int slot, const double* gamma, const double* n, const double* delta,
                          const double* deltap, const double* sd);
int rtx_voigt_set_mix(rtx_handle* h, int n_extra, const double* x_extra);

/* Line-file readers (host only, no handle, no GPU needed).  Call with capacity 0 and NULL columns to count rows,
 * then again with arrays of that size; any column pointer may be NULL.
 * rtx_par_read: HITRAN 160-character rows, the format hapi stores tables in (layout misc/hapi.py:468-538; parse
 * rules misc/hapi.py:1535-1556 -- fixed-width int()/float(), unparsable rows skipped :1650-1655).
 * rtx_tape3_read / rtx_tape3_write: LNFL TAPE3 binary line file, the format of the reference's
 * AER-v3.6-0500-6000.tp3 (radiative_transfer.py:151,491).  The reference has no reader for it and ships the file
 * as a Git-LFS pointer; the layout follows LNFL conventions (see csrc/ingest.cu) and is verified only by a
 * write/read round trip.  vmin/vmax select lines by wavenumber. */
int rtx_par_read(const char* path, int64_t capacity, int32_t* molec_id, int32_t* local_iso_id, double* nu, double* sw,
                 double* a, double* gamma_air, double* gamma_self, double* elower, double* n_air, double* delta_air,
                 int64_t* n_rows);
int rtx_tape3_write(const char* path, int64_t n, const int32_t* molec_id, const int32_t* local_iso_id,
                    const double* nu, const double* sw, const double* gamma_air, const double* gamma_self,
                    const double* elower, const double* n_air, const double* delta_air);
int rtx_tape3_read(const char* path, double vmin, double vmax, int64_t capacity, int32_t* molec_id,
                   int32_t* local_iso_id, double* nu, double* sw, double* gamma_air, double* gamma_self,
                   double* elower, double* n_air, double* delta_air, int64_t* n_rows);

/* ---- subsystem 2: Voigt line sum ----------------------------------------------------------
 * Replaces the line loop of absorptionCoefficient_Voigt (misc/hapi.py:11050-11138) together with
 * PROFILE_VOIGT/pcqsdhc PART1/hum1_wei (misc/hapi.py:10131-10140, 9910-9915, 9833-9844), for
 * n_states homogeneous (T, p) states at once (layers of one or many atmospheres):
 *
 *   out[s][i] = sum over lines  weight[s][iso(line)] * S_line(T_s) * V(grid[i]; line, T_s, p_s)
 *
 * restricted to the reference's window  nu-W < grid[i] <= nu+W,  W = max(wing_abs, wing_hw*Gamma0,
 * wing_hw*GammaD)  (misc/hapi.py:11131-11134).  Lines with S(T) < s_threshold or weight 0 are
 * skipped (:11066, :11082).
 *   T_dev, p_dev     [n_states]          K, atm
 *   weight_dev       [n_states][n_iso]   factor*abundance/natural_abundance (:11136-11138), times the
 *                                        column amount when optical depth is wanted
 *   x_air, x_self    diluent fractions (Diluent={'air':..,'self':..}, :11090-11128)
 *   grid_dev         [nx] sorted ascending (any spacing);   out_dev [n_states][nx], overwritten
 *   flags            RTX_VOIGT_* bits */
#define RTX_VOIGT_DEFAULT 0
#define RTX_VOIGT_FAST_RCP 1 /* accepted for compatibility: the one-Newton-step reciprocal is the default now */
#define RTX_VOIGT_FP32 2 /* optional FP32 evaluation of the asymptote branch (tolerance 1e-5 instead of 1e-9) */
#define RTX_VOIGT_WARPSPEC 4 /* warp-specialised kernel variant: producer warps run the line prologue one chunk ahead */
#define RTX_VOIGT_LORENTZ 16 /* Lorentz profile and wing rule of absorptionCoefficient_Lorentz (misc/hapi.py:11144-11381) */
#define RTX_VOIGT_DOPPLER 32    /* PROFILE_DOPPLER, absorptionCoefficient_Doppler (misc/hapi.py:11384-11550) */
#define RTX_VOIGT_NOSHIFT 64    /* with RTX_VOIGT_DOPPLER: LineShift=False */
#define RTX_VOIGT_EXACT_RCP 128 /* second Newton step in the reciprocal: every asymptote-branch term to full double
                                 * precision (default: one step, relative error <= 2^-40 ~ 9e-13 per term) */
#define RTX_VOIGT_NOFAR 4096    /* disable the far-field blocks: evaluate every (line, grid point) pair at the grid point
                                 * itself (default: slices a window covers from >= 0.375 slice widths away are summed
                                 * at 32 Chebyshev nodes and interpolated, relative error < 1e-13) */
#define RTX_VOIGT_NOTILEFAR 16384 /* keep the slice-level far-field blocks but not the tile-level ones (a line whose window
                                 * covers a whole tile from >= 0.375 tile widths away is summed at the tile's 32 nodes) */
#define RTX_VOIGT_TILEFAR 32768  /* use the tile-level far-field blocks wherever a line qualifies (default: only for states
                                 * whose typical wing is >= 4000 grid points and whose tiles see >= 600 such lines -- below
                                 * that the kernel variant without them is faster) */
#define RTX_VOIGT_NOSPLIT 65536  /* small launches (fewer (tile, state) pairs than four per SM -- a single cross-section): do
                                 * not split the candidate lines of a tile over several CTAs (default: up to 8 CTAs per
                                 * tile, partial spectra added in part order; never with a forced tile size).  Batch
                                 * callers set it so that results do not depend on how the batch is chunked. */
#define RTX_VOIGT_SDVOIGT 8192  /* absorptionCoefficient_SDVoigt (misc/hapi.py:10657-10905): lines whose Gamma2 = sum_d x_d
                                 * SD_d p gamma_d is non-zero take the quadratic speed-dependent profile (pcqsdhc PART2-4,
                                 * misc/hapi.py:9923-10020); the rest is the Voigt driver */
#define RTX_VOIGT_TILE_SHIFT 8 /* bits 8..11: force 256*K grid points per CTA, K in {1,2,4,8} (0 = automatic).
                                * The automatic choice between K = 8 and K = 4, and between the K = 8 kernel with and
                                * without tile-level far-field blocks, is made on the device, per state, from the state's
                                * wing in grid points (the grids of all candidates are launched; a CTA of a variant its
                                * state does not pick retires at once), so the call never synchronises; problems of at
                                * most 65 536 points x states take one grid of 512-point tiles instead.
                                * Force K to pin the summation order (e.g. for bit-identical sharded results). */
int rtx_voigt_sum(rtx_handle* h, int n_states, const double* T_dev, const double* p_dev,
                  const double* weight_dev, double x_air, double x_self, double wing_abs, double wing_hw,
                  double s_threshold, const double* grid_dev, int64_t nx, double* out_dev, int flags,
                  void* stream);

/* Exact count of (line, grid point) pairs inside the reference's window ("evals") and
 * of those in the Weideman region |x|+y<15, for the same arguments; counts_dev[2] (uint64, device)
 * is overwritten. */
int rtx_voigt_count(rtx_handle* h, int n_states, const double* T_dev, const double* p_dev,
                    const double* weight_dev, double x_air, double x_self, double wing_abs, double wing_hw,
                    double s_threshold, const double* grid_dev, int64_t nx, uint64_t* counts_dev,
                    void* stream);

/* Q(T) for every isotopologue: replaces PYTIPS (misc/hapi.py:10030, 5311-5388).  T_dev[n],
 * q_dev[n][n_iso].  Temperatures outside 70..3000 K give NaN (the shim raises like hapi :9571). */
int rtx_tips(rtx_handle* h, int n, const double* T_dev, double* q_dev, void* stream);

/* ---- subsystem 3: layer integration (TUD) ------------------------------------------------
 * Replaces the numpy body of compute_TUD after the OD loop (radiative_transfer.py:340-392) for
 * n_atm atmospheres that share the layer altitudes Z:
 *   od_dev   [n_atm][n_layers][nx]  optical depth per layer (layer-major, as rtx_voigt_sum writes it)
 *   T_dev    [n_atm][n_layers]      layer temperatures;   grid_dev [nx]
 *   Z_host   [n_layers] layer altitudes; alt_host [n_alt] sensor altitudes; mu_host [n_mu] = 1/cos(theta_r)
 *   tau_dev, lu_dev [n_atm][nx][n_alt][n_mu];  ld_dev [n_atm][nx]
 * Quirks of the reference are kept: layer mask Z<=alt (:348), upwelling over the first count(mask)
 * layers (:353-356), downwelling over the layer count of the LAST altitude (:353,:370);
 * return_od!=0 stores the optical-depth sum instead of the transmittance (:349-350).
 * The two halves can be launched separately: ld_dev == NULL computes only tau and lu, tau_dev == lu_dev == NULL
 * only ld (a caller can then start moving tau / lu off the device while the downwelling half still runs). */
int rtx_tud(rtx_handle* h, int n_atm, int n_layers, const double* od_dev, const double* T_dev,
            const double* grid_dev, int64_t nx, const double* Z_host, const double* alt_host, int n_alt,
            const double* mu_host, int n_mu, int n_angle, int return_od, double* tau_dev, double* lu_dev,
            double* ld_dev, void* stream);

/* ---- the whole path for a host without Python ------------------------------------------------------------------
 * rtx_set_iso_molecules: molecule number and abundance factor (factor * abundance / natural_abundance,
 * misc/hapi.py:11136-11138) of every isotopologue of the table set by rtx_set_iso_table.  Host pointers, copied.
 * rtx_layer_weights: pressures in atm and the per-(state, isotopologue) column amounts that turn the Voigt sum into a
 * layer's optical depth -- what the reference gets from LBLRTM per layer (compute_OD, radiative_transfer.py:395-456)
 * and SURVEY.md App. B.1 wires through hapi:
 *     weight[s][iso] = (MF[s][m(iso)] * 1e-6) * volumeConcentration(p_s, T_s) * (PL_s * 1e5) * factor[iso]
 *   T_dev [K], P_pa_dev [Pa], PL_km_dev [km]: [n_states];  mf_ppmv_dev [n_states][n_mol] with the molecule numbers
 *   mf_id_host[n_mol] (MFs_VAL / MFs_ID of the reference's options, radiative_transfer.py:171-177);
 *   p_atm_dev [n_states], weight_dev [n_states][n_iso] are written.
 * rtx_tud_from_lines: atmosphere tables in, (tau, Lu, Ld) out, for n_atm atmospheres of n_layers layers each -- the body
 * of the production loop `for jj: compute_TUD(..., MFs_VAL=MFs_VAL[jj], Ts=Ts[jj])` (Generate_LWIR_TUD.py:117-122) with
 * compute_TUD = radiative_transfer.py:274-392.  Column amounts, Voigt sums (rtx_voigt_sum arguments x_air .. voigt_flags)
 * and layer integration (rtx_tud arguments Z_host .. return_od) run back to back on `stream`; optical depth lives in
 * the handle's scratch (at most 4 GB at a time).  T / P / PL: [n_atm][n_layers], mf: [n_atm][n_layers][n_mol]; outputs
 * as rtx_tud. */
int rtx_set_iso_molecules(rtx_handle* h, int n_iso, const int32_t* molec_id, const double* factor);
int rtx_layer_weights(rtx_handle* h, int n_states, const double* T_dev, const double* P_pa_dev, const double* PL_km_dev,
                      const double* mf_ppmv_dev, int n_mol, const int32_t* mf_id_host, double* p_atm_dev,
                      double* weight_dev, void* stream);
int rtx_tud_from_lines(rtx_handle* h, int n_atm, int n_layers, const double* T_dev, const double* P_pa_dev,
                       const double* PL_km_dev, const double* mf_ppmv_dev, int n_mol, const int32_t* mf_id_host,
                       double x_air, double x_self, double wing_abs, double wing_hw, double s_threshold, int voigt_flags,
                       const double* grid_dev, int64_t nx, const double* Z_host, const double* alt_host, int n_alt,
                       const double* mu_host, int n_mu, int n_angle, int return_od, double* tau_dev, double* lu_dev,
                       double* ld_dev, void* stream);

/* planckian(X, T) in wavenumber or wavelength form (radiative_transfer.py:792-848):
 * out_dev[nx][nT].  wavelength!=0 selects the micrometre branch (the mean(X)<50 heuristic lives in
 * the shim). */
int rtx_planck(rtx_handle* h, const double* X_dev, int64_t nx, const double* T_dev, int64_t nT,
               int wavelength, double* out_dev, void* stream);

/* brightnessTemperature (inverse=0; radiative_transfer.py:851-933) and BT2L (inverse=1; :936-1014), element-wise:
 * X_dev[nx], V_dev[nx][m] (radiance or temperature), out_dev[nx][m].  Non-finite or non-positive inputs give
 * bad_value, as the reference's ixBad mask does. */
int rtx_brightness(rtx_handle* h, const double* X_dev, int64_t nx, const double* V_dev, int64_t m, int wavelength,
                   int inverse, double bad_value, double* out_dev, void* stream);

/* ---- subsystem 3b: apparent radiance -----------------------------------------------------
 * compute_LWIR_apparent_radiance (radiative_transfer.py:1017-1069):
 *   L[x][e][a][t] = tau[x][a]*(emis[x][e]*B(X[x], Ts[a]+dT[t]) + (1-emis[x][e])*Ld[x][a]) + La[x][a]
 * dT_dev may be NULL (n_dT is then 1 and the t axis is absent).  ls_dev (surface-leaving radiance,
 * return_Ls=True) may be NULL. */
int rtx_apparent_radiance(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                          const double* Ts_dev, int64_t n_atm, const double* tau_dev, const double* la_dev,
                          const double* ld_dev, const double* dT_dev, int64_t n_dT, double* L_dev,
                          double* ls_dev, void* stream);

/* float32-output variant (the reference's production script casts every input to float32 and stores a float32 cube,
 * Compute_LWIR_Apparent_Radiance.py:9-25): same FP64 arithmetic on the same double inputs, rounded once at the store. */
int rtx_apparent_radiance_f32(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                              const double* Ts_dev, int64_t n_atm, const double* tau_dev, const double* la_dev,
                              const double* ld_dev, const double* dT_dev, int64_t n_dT, float* L_dev,
                              float* ls_dev, void* stream);

/* Per-pixel mixed-material variant: the inner loop of LWIR_HSI_gen (LWIR_HSI_Generator.py:148-166).
 * For pixel k: em = sum_m frac[k][m]*emis[mat[k][m]][:],  L[k][x] = tau[atm[k]][x]*(em*B(X[x],T[k]) +
 * (1-em)*Ld[atm[k]][x]) + La[atm[k]][x].   emis_dev [n_emis][nx], tau/la/ld [n_atm][nx] (band-minor, the
 * layout of that script), L_dev [n_pix][nx]. */
int rtx_hsi_pixels(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                   const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm,
                   int64_t n_pix, const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix,
                   const int32_t* mat_idx_dev, const double* frac_dev, double* L_dev, void* stream);

/* float32 cube (config 4's "FP32 variant"): L_dev [n_pix][nx] float. */
int rtx_hsi_pixels_f32(rtx_handle* h, const double* X_dev, int64_t nx, const double* emis_dev, int64_t n_emis,
                       const double* tau_dev, const double* la_dev, const double* ld_dev, int64_t n_atm,
                       int64_t n_pix, const int32_t* atm_idx_dev, const double* T_pix_dev, int n_mix,
                       const int32_t* mat_idx_dev, const double* frac_dev, float* L_dev, void* stream);

/* ---- subsystem 4: MAKO instrument line shape ---------------------------------------------
 * The weighted-mean resampling at the heart of ILS_MAKO (triangle: radiative_transfer.py:1236-1256;
 * Gaussian: ILS_MAKO.py:24-34).  Band centres and widths are prepared by the shim exactly as the
 * reference does (sort, clip, gradient; :1226-1241).
 *   mode 0: w = max(0, 1-|x-c_b|/s_b);  mode 1: w = exp(-0.5((x-c_b)/s_b)^2)/(s_b*sqrt(2pi))
 *   (experiment bits: | 0x100 band-per-CTA walk, | 0x200 grouped walk without TMA staging, | 0x400 / 0x800 two / one
 *    spectra per lane, | 0x1000 groups of 16 bands; default for ns >= 16: groups of 8 adjacent bands, tiles of 32-64 rows
 *    x 64-32 spectra staged in shared memory by 2-D tensor-map TMA copies)
 *   X_dev [nx] ascending, Y_dev [nx][ns], centre_dev/width_dev [n_bands], out_dev [n_bands][ns]
 *   out[b][s] = sum_x w*Y / sum_x w */
int rtx_ils_apply(rtx_handle* h, int mode, const double* X_dev, int64_t nx, const double* Y_dev, int64_t ns,
                  int n_bands, const double* centre_dev, const double* width_dev, double* out_dev,
                  void* stream);

/* Band centres and widths of the MAKO line shapes, host only (no handle): mode 0 = triangle, exactly as
 * radiative_transfer.py:1226-1241 derives them (optional res_factor interpolation of the 128 band wavelengths, sort of
 * 1e4/lambda, clip to the open interval (xmin, xmax) = (X.min(), X.max()), half-base fwhm_sf*|gradient|*1.6, centres
 * scale*X_out + shift); mode 1 = Gaussian of ILS_MAKO.py:20-22 (no clip, sigma = |gradient|; res_factor, fwhm_sf, shift,
 * scale ignored).  res_factor <= 0 means none.  Call with capacity 0 to get the band count in *n_out. */
int rtx_mako_bands(int mode, double xmin, double xmax, double res_factor, double fwhm_sf, double shift, double scale,
                   int capacity, double* X_out, double* centre, double* width, int* n_out);

/* ---- reduceResolution / smooth (radiative_transfer.py:1266-1350), the step after TUD in the production driver
 * (Generate_LWIR_TUD.py:82-85,124-126).
 * rtx_smooth_sym: column-wise for Y_dev[n][m]: symmetric=0 gives smooth(y) (the reference's reflect-pad / convolve
 * with the NORMALISED window window_dev[window_len] / trim, :1266-1323), symmetric=1 gives
 * 0.5*(smooth(y) + smooth(y[::-1])[::-1]) as reduceResolution uses it (:1330).
 * rtx_spline_eval: C2 cubic-spline interpolation of the columns of Y_dev[n][m] on the knots X_dev[n] (ascending, n >= 4)
 * at Xo_dev[n_out] -> out_dev[n_out][m]; not-a-knot end conditions and end-polynomial extrapolation, i.e. scipy
 * interp1d(kind="cubic", fill_value="extrapolate") as reduceResolution calls it (:1331-1333), to rounding. */
int rtx_smooth_sym(rtx_handle* h, const double* Y_dev, int64_t n, int64_t m, const double* window_dev, int window_len,
                   int symmetric, double* out_dev, void* stream);
int rtx_spline_eval(rtx_handle* h, const double* X_dev, const double* Y_dev, int64_t n, int64_t m,
                    const double* Xo_dev, int64_t n_out, double* out_dev, void* stream);

/* ---- measurement helpers -------------------------------------------------------------------
 * Dependent-free DFMA loop: writes the achieved FP64 FMA rate [flop/s] of the handle's device.
 * Used by bench.py as the measured denominator of the Voigt kernel's FP64 roofline. */
int rtx_measure_fp64_peak(rtx_handle* h, double* flops_per_s);
/* The 24 Weideman coefficients the kernels use (highest power first), for the unit test that regenerates them
 * from the FFT recipe of misc/hapi.py:9816-9824.  Host pointer, 24 doubles. */
int rtx_weideman_coeffs(double* out24);
/* Number of kernels this library has launched on this handle since creation. */
int rtx_launch_count(rtx_handle* h, uint64_t* n);
/* Tile size the last rtx_voigt_sum on this handle used: K = grid points per CTA / 256 (0 before the first call).  When
 * the size was chosen on the device (per state) this reports the choice the launch-wide eval-weighted mean wing bound
 * would make -- the size most of the work ran at; it reads 8 bytes back and synchronises the stream of that call. */
int rtx_voigt_last_tile(rtx_handle* h, int* K);

#ifdef __cplusplus
}
#endif
#endif /* RADTXFR_B200_H */
