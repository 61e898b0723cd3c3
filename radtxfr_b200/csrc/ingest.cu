// Subsystem 1 (host side): line-file readers feeding rtx_lines_upload.
//
//   rtx_par_read     HITRAN 160-character ".par"/".data" rows, the transport format hapi stores tables in
//                    (column layout misc/hapi.py:468-538; row parsing misc/hapi.py:1535-1556: int()/float() of
//                    fixed-width slices, rows that fail to parse are skipped, :1650-1655).
//   rtx_tape3_read / rtx_tape3_write
//                    LNFL "TAPE3" binary line file, the format of the reference's bundled
//                    AER-v3.6-0500-6000.tp3 (radiative_transfer.py:151, 491).  The reference contains no TAPE3
//                    reader (LBLRTM reads the file itself) and ships the file only as a Git-LFS pointer, so this
//                    layout follows the LNFL/LBLRTM conventions as recalled in SURVEY.md App. E and is
//                    UNVERIFIED AGAINST A REAL AER FILE; it is validated by a writer/reader round trip only.
//
// TAPE3 layout implemented here (gfortran sequential unformatted, little endian: every record is
// [int32 nbytes][payload][int32 nbytes]):
//   record 1   file header: 10 x char[8] HID, 64 x char[8] molecule names, int32 MOLCNT[64], int32 MCNTLC[64],
//              int32 MCNTNL[64], float32 SUMSTR[64], int32 LINMOL, float64 FLINLO, float64 FLINHI, int32 LINCNT,
//              int32 ILINLC, int32 ILINNL, int32 IREC, int32 IRECTL, 2 x char[8] HID1
//   then per block of <= 250 lines two records:
//     panel header   float64 VMIN, float64 VMAX, int32 NREC (lines in the block), int32 NWDS
//     panel data     float64 VNU[250], float32 SP[250], ALFA0[250], EPP[250], int32 MOL[250], float32 HWHMS[250],
//                    TMPALF[250], PSHIFT[250], int32 IFLG[250]                                   (10 000 bytes)
//   SP = S / (nu * (1 - exp(-c2 nu / 296)))  with LBLRTM's c2 = 1.4387752 ;  MOL = M + 100 * ISO ;
//   ALFA0 = gamma_air, HWHMS = gamma_self, TMPALF = n_air, PSHIFT = delta_air, EPP = elower ;
//   IFLG != 0 marks a line-coupling entry whose following slot holds coefficients, not a line (skipped on read).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "rtx_common.cuh"

namespace {

constexpr double LBL_C2 = 1.4387752;   // RADCN2 of LBLRTM/LNFL
constexpr double LBL_T0 = 296.0;
constexpr int PANEL = 250;

// float()/int() of a fixed-width slice the way Python parses it: surrounding blanks allowed, anything else
// left over is an error.
bool parse_double(const char* s, int len, double* out) {
  char buf[32];
  if (len >= (int)sizeof(buf)) return false;
  memcpy(buf, s, len);
  buf[len] = 0;
  char* p = buf;
  while (*p == ' ' || *p == '\t') ++p;
  if (!*p) return false;
  char* end = nullptr;
  const double v = strtod(p, &end);
  if (end == p) return false;
  while (*end == ' ' || *end == '\t' || *end == '\r' || *end == '\n') ++end;
  if (*end) return false;
  *out = v;
  return true;
}
bool parse_int(const char* s, int len, int* out) {
  char buf[16];
  if (len >= (int)sizeof(buf)) return false;
  memcpy(buf, s, len);
  buf[len] = 0;
  char* p = buf;
  while (*p == ' ') ++p;
  if (!*p) return false;
  char* end = nullptr;
  const long v = strtol(p, &end, 10);
  if (end == p) return false;
  while (*end == ' ') ++end;
  if (*end) return false;
  *out = (int)v;
  return true;
}

struct Rec {
  FILE* f;
  bool write_rec(const void* payload, int32_t n) {
    return fwrite(&n, 4, 1, f) == 1 && fwrite(payload, 1, n, f) == (size_t)n && fwrite(&n, 4, 1, f) == 1;
  }
  // reads one record into buf; returns payload size or -1 at EOF / -2 on a framing error
  int64_t read_rec(std::vector<char>& buf) {
    int32_t n = 0, m = 0;
    if (fread(&n, 4, 1, f) != 1) return -1;
    if (n < 0 || n > (1 << 26)) return -2;
    buf.resize(n);
    if (n && fread(buf.data(), 1, n, f) != (size_t)n) return -2;
    if (fread(&m, 4, 1, f) != 1 || m != n) return -2;
    return n;
  }
};

constexpr int HDR_BYTES = 10 * 8 + 64 * 8 + 64 * 4 * 3 + 64 * 4 + 4 + 8 + 8 + 4 * 5 + 2 * 8;

}  // namespace

extern "C" int rtx_par_read(const char* path, int64_t capacity, int32_t* molec_id, int32_t* local_iso_id, double* nu,
                            double* sw, double* a, double* gamma_air, double* gamma_self, double* elower,
                            double* n_air, double* delta_air, int64_t* n_rows) {
  RTX_REQUIRE(path && n_rows, "rtx_par_read: null pointer");
  FILE* f = fopen(path, "r");
  if (!f) { rtx_set_error(std::string("rtx_par_read: cannot open ") + path); return RTX_E_INVALID; }
  char line[1024];
  int64_t n = 0;
  while (fgets(line, sizeof(line), f)) {
    const int len = (int)strlen(line);
    if (len < 67) continue;   // too short to hold the numeric columns: hapi's slice parse fails -> row skipped
    int M, I;
    double v[8];
    // molec_id %2d, local_iso_id %1d, nu %12, sw %10, a %10, gamma_air %5, gamma_self %5, elower %10, n_air %4,
    // delta_air %8 (misc/hapi.py:468-480)
    static const int pos[8] = {3, 15, 25, 35, 40, 45, 55, 59}, wid[8] = {12, 10, 10, 5, 5, 10, 4, 8};
    if (!parse_int(line, 2, &M) || !parse_int(line + 2, 1, &I)) continue;
    bool ok = true;
    for (int k = 0; k < 8 && ok; ++k) ok = parse_double(line + pos[k], wid[k], &v[k]);
    if (!ok) continue;
    if (n < capacity) {
      if (molec_id) molec_id[n] = M;
      if (local_iso_id) local_iso_id[n] = I;
      if (nu) nu[n] = v[0];
      if (sw) sw[n] = v[1];
      if (a) a[n] = v[2];
      if (gamma_air) gamma_air[n] = v[3];
      if (gamma_self) gamma_self[n] = v[4];
      if (elower) elower[n] = v[5];
      if (n_air) n_air[n] = v[6];
      if (delta_air) delta_air[n] = v[7];
    }
    ++n;
  }
  fclose(f);
  *n_rows = n;
  return RTX_OK;
}

extern "C" int rtx_tape3_write(const char* path, int64_t n, const int32_t* molec_id, const int32_t* local_iso_id,
                               const double* nu, const double* sw, const double* gamma_air, const double* gamma_self,
                               const double* elower, const double* n_air, const double* delta_air) {
  RTX_REQUIRE(path && n >= 0, "rtx_tape3_write: bad arguments");
  RTX_REQUIRE(n == 0 || (molec_id && local_iso_id && nu && sw && gamma_air && gamma_self && elower && n_air && delta_air),
              "rtx_tape3_write: null column");
  for (int64_t i = 1; i < n; ++i)
    RTX_REQUIRE(nu[i] >= nu[i - 1], "rtx_tape3_write: lines must be sorted by wavenumber");
  FILE* f = fopen(path, "wb");
  if (!f) { rtx_set_error(std::string("rtx_tape3_write: cannot open ") + path); return RTX_E_INVALID; }
  Rec R{f};
  std::vector<char> hdr(HDR_BYTES, ' ');
  {
    char* p = hdr.data();
    const char* id = "radtxfr_b200 TAPE3 writer (LNFL layout, unverified)";
    memcpy(p, id, std::min<size_t>(strlen(id), 80));
    p += 80 + 64 * 8;
    std::vector<int32_t> cnt(64, 0);
    std::vector<float> sums(64, 0.f);
    for (int64_t i = 0; i < n; ++i)
      if (molec_id[i] >= 1 && molec_id[i] <= 64) { cnt[molec_id[i] - 1]++; sums[molec_id[i] - 1] += (float)sw[i]; }
    memcpy(p, cnt.data(), 256); p += 256;
    memset(p, 0, 512); p += 512;                      // MCNTLC, MCNTNL
    memcpy(p, sums.data(), 256); p += 256;
    int32_t linmol = 0;
    for (int m = 0; m < 64; ++m) if (cnt[m]) linmol = m + 1;
    memcpy(p, &linmol, 4); p += 4;
    const double lo = n ? nu[0] : 0.0, hi = n ? nu[n - 1] : 0.0;
    memcpy(p, &lo, 8); p += 8;
    memcpy(p, &hi, 8); p += 8;
    const int32_t lincnt = (int32_t)n, zero = 0, irec = (int32_t)((n + PANEL - 1) / PANEL);
    memcpy(p, &lincnt, 4); p += 4;
    memcpy(p, &zero, 4); p += 4;
    memcpy(p, &zero, 4); p += 4;
    memcpy(p, &irec, 4); p += 4;
    memcpy(p, &irec, 4); p += 4;
  }
  bool ok = R.write_rec(hdr.data(), HDR_BYTES);
  std::vector<char> blk(PANEL * 40);
  for (int64_t b0 = 0; b0 < n && ok; b0 += PANEL) {
    const int cntb = (int)std::min<int64_t>(PANEL, n - b0);
    char ph[24];
    const double vmin = nu[b0], vmax = nu[b0 + cntb - 1];
    const int32_t nrec = cntb, nwds = PANEL * 10;
    memcpy(ph, &vmin, 8); memcpy(ph + 8, &vmax, 8); memcpy(ph + 16, &nrec, 4); memcpy(ph + 20, &nwds, 4);
    ok = ok && R.write_rec(ph, 24);
    memset(blk.data(), 0, blk.size());
    double* vnu = reinterpret_cast<double*>(blk.data());
    float* sp = reinterpret_cast<float*>(blk.data() + PANEL * 8);
    float* alfa = sp + PANEL; float* epp = alfa + PANEL;
    int32_t* mol = reinterpret_cast<int32_t*>(epp + PANEL);
    float* hwhms = reinterpret_cast<float*>(mol + PANEL);
    float* tmpalf = hwhms + PANEL; float* pshift = tmpalf + PANEL;
    int32_t* iflg = reinterpret_cast<int32_t*>(pshift + PANEL);
    for (int k = 0; k < cntb; ++k) {
      const int64_t i = b0 + k;
      vnu[k] = nu[i];
      sp[k] = (float)(sw[i] / (nu[i] * (1.0 - exp(-LBL_C2 * nu[i] / LBL_T0))));
      alfa[k] = (float)gamma_air[i]; epp[k] = (float)elower[i];
      mol[k] = molec_id[i] + 100 * local_iso_id[i];
      hwhms[k] = (float)gamma_self[i]; tmpalf[k] = (float)n_air[i]; pshift[k] = (float)delta_air[i];
      iflg[k] = 0;
    }
    ok = ok && R.write_rec(blk.data(), PANEL * 40);
  }
  fclose(f);
  if (!ok) { rtx_set_error("rtx_tape3_write: short write"); return RTX_E_INVALID; }
  return RTX_OK;
}

extern "C" int rtx_tape3_read(const char* path, double vmin, double vmax, int64_t capacity, int32_t* molec_id,
                              int32_t* local_iso_id, double* nu, double* sw, double* gamma_air, double* gamma_self,
                              double* elower, double* n_air, double* delta_air, int64_t* n_rows) {
  RTX_REQUIRE(path && n_rows, "rtx_tape3_read: null pointer");
  FILE* f = fopen(path, "rb");
  if (!f) { rtx_set_error(std::string("rtx_tape3_read: cannot open ") + path); return RTX_E_INVALID; }
  Rec R{f};
  std::vector<char> buf;
  int64_t got = R.read_rec(buf);
  if (got < 80) {
    fclose(f);
    rtx_set_error("rtx_tape3_read: not a Fortran-unformatted TAPE3 file (a Git-LFS pointer to the AER file is text, "
                  "not line data)");
    return RTX_E_INVALID;
  }
  int64_t n = 0;
  bool skip_next = false;
  for (;;) {
    got = R.read_rec(buf);
    if (got == -1) break;
    if (got != 24) { fclose(f); rtx_set_error("rtx_tape3_read: bad panel header"); return RTX_E_INVALID; }
    double pv0, pv1; int32_t nrec;
    memcpy(&pv0, buf.data(), 8); memcpy(&pv1, buf.data() + 8, 8); memcpy(&nrec, buf.data() + 16, 4);
    got = R.read_rec(buf);
    if (got != PANEL * 40 || nrec < 0 || nrec > PANEL) { fclose(f); rtx_set_error("rtx_tape3_read: bad panel"); return RTX_E_INVALID; }
    if (pv1 < vmin || pv0 > vmax) { skip_next = false; continue; }
    const double* vnu = reinterpret_cast<const double*>(buf.data());
    const float* sp = reinterpret_cast<const float*>(buf.data() + PANEL * 8);
    const float* alfa = sp + PANEL; const float* epp = alfa + PANEL;
    const int32_t* mol = reinterpret_cast<const int32_t*>(epp + PANEL);
    const float* hwhms = reinterpret_cast<const float*>(mol + PANEL);
    const float* tmpalf = hwhms + PANEL; const float* pshift = tmpalf + PANEL;
    const int32_t* iflg = reinterpret_cast<const int32_t*>(pshift + PANEL);
    for (int k = 0; k < nrec; ++k) {
      if (skip_next) { skip_next = false; continue; }      // slot holding line-coupling coefficients
      if (iflg[k] != 0) skip_next = true;                   // this is a line; its coefficients follow
      const double v = vnu[k];
      if (v < vmin || v > vmax) continue;
      if (n < capacity) {
        const int m = mol[k] % 100, iso = (mol[k] / 100) % 10;
        if (molec_id) molec_id[n] = m;
        if (local_iso_id) local_iso_id[n] = iso;
        if (nu) nu[n] = v;
        if (sw) sw[n] = (double)sp[k] * (v * (1.0 - exp(-LBL_C2 * v / LBL_T0)));
        if (gamma_air) gamma_air[n] = alfa[k];
        if (gamma_self) gamma_self[n] = hwhms[k];
        if (elower) elower[n] = fabs((double)epp[k]);       // negative EPP = "unknown" flag in LNFL
        if (n_air) n_air[n] = tmpalf[k];
        if (delta_air) delta_air[n] = pshift[k];
      }
      ++n;
    }
  }
  fclose(f);
  *n_rows = n;
  return RTX_OK;
}
