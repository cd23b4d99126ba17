// c2d_io.cu -- the device-owned state in the reference's own file formats.  Host code only.
//
//   census file  p###_census.dat   write_cens / read_cens (census2d.f:1-76): per photon one line `6e14.7`
//                                  (rpre zpre wmu phi ew xnu) and one line `6i5` (jgpsp jgplc jgpmu jph kph seed),
//                                  written by write_record (write_record.f:432-437) and read back by read_record
//   event file   p###_<eventfile>  one line per escaping photon, format 105 `6(e14.7,1x),e14.7` (imcleak2d.f:171,181):
//                                  t_bound xnu ew rpre zpre wmu phi -- what postprocessing/plcm.c, pspt.c read
// The E14.7 edit descriptor is restated here (fortran_e14_7): 0.dddddddE+xx, mantissa in [0.1, 1), the sign slot blank for
// positive values, `0.ddddddd+xxx` when the exponent needs three digits; pinned byte for byte against an event file the
// reference binary wrote (tests/golden/evb_dense.txt).  Like the reference's restart, a census written this way keeps 7
// digits per double; the full-precision path is c2d_census_export / _import.
#include "c2d_ctx.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace c2d {

// Fortran E14.7 into out[14] (no terminator)
static void fortran_e14_7(double v, char *out)
{
    char buf[40];
    if (std::isnan(v)) { memcpy(out, "           NaN", 14); return; }
    if (std::isinf(v)) { memcpy(out, v > 0 ? "      Infinity" : "     -Infinity", 14); return; }
    if (v == 0.0) { memcpy(out, std::signbit(v) ? "-0.0000000E+00" : " 0.0000000E+00", 14); return; }
    snprintf(buf, sizeof buf, "%.6E", fabs(v));                // d.ddddddE+xx, correctly rounded to 7 significant digits
    int ex = atoi(buf + 9) + 1;                                  // 0.ddddddd form: exponent one higher
    out[0] = v < 0 ? '-' : ' ';
    out[1] = '0'; out[2] = '.';
    out[3] = buf[0];
    memcpy(out + 4, buf + 2, 6);
    const int ax = ex < 0 ? -ex : ex;
    if (ax <= 99) snprintf(buf, sizeof buf, "E%c%02d", ex < 0 ? '-' : '+', ax);
    else snprintf(buf, sizeof buf, "%c%03d", ex < 0 ? '-' : '+', ax);
    memcpy(out + 10, buf, 4);
}

// Fortran I5 into out[5]: right-justified, `*****` when it does not fit
static void fortran_i5(int v, char *out)
{
    char buf[16];
    const int n = snprintf(buf, sizeof buf, "%5d", v);
    if (n > 5) memcpy(out, "*****", 5); else memcpy(out, buf, 5);
}

// a 14-column E field as Fortran reads it: blanks ignored, `0.1234567+100` (no exponent letter) accepted
static bool parse_e_field(const char *p, int width, double *v)
{
    char buf[48];
    int n = 0;
    for (int i = 0; i < width && p[i] && p[i] != '\n' && p[i] != '\r'; i++)
        if (p[i] != ' ') buf[n++] = p[i];
    buf[n] = 0;
    if (n == 0) { *v = 0.0; return true; }                       // an all-blank field reads as zero
    for (int i = 1; i < n; i++)
        if ((buf[i] == '+' || buf[i] == '-') && buf[i - 1] != 'E' && buf[i - 1] != 'e' && buf[i - 1] != 'D' && buf[i - 1] != 'd') {
            memmove(buf + i + 1, buf + i, (size_t)(n - i + 1));
            buf[i] = 'E';
            n++;
            break;
        }
    for (int i = 0; i < n; i++) if (buf[i] == 'D' || buf[i] == 'd') buf[i] = 'E';
    char *end = nullptr;
    *v = strtod(buf, &end);
    return end && *end == 0;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

void c2d_format_e14_7(double v, char *out14)
{
    fortran_e14_7(v, out14);
}

int c2d_events_write(c2d_ctx *h, const char *path, int append, int64_t *n_written)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !path) return fail(C2D_ESTATE, "context not initialised / NULL path");
    int64_t n = 0;
    std::vector<c2d_event> ev((size_t)std::max<int64_t>(c->event_cap, 1));
    const int rc = c2d_events_drain(h, ev.data(), c->event_cap, &n);
    if (rc && rc != C2D_EEVENTS) return rc;
    const int64_t stored = std::min<int64_t>(n, c->event_cap);
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return fail(C2D_EINVAL, "cannot open %s", path);
    std::string line(7 * 15, ' ');
    line[104] = '\n';
    for (int64_t i = 0; i < stored; i++) {
        const double *r = &ev[(size_t)i].t_bound;
        for (int k = 0; k < 7; k++) fortran_e14_7(r[k], &line[(size_t)15 * k]);      // 6(e14.7,1x),e14.7
        fwrite(line.data(), 1, 105, f);
    }
    if (fclose(f) != 0) return fail(C2D_EINVAL, "write to %s failed", path);
    if (n_written) *n_written = stored;
    return rc;                                                    // C2D_EEVENTS: records beyond event_capacity were lost
}

int c2d_census_write(c2d_ctx *h, const char *path, int64_t *n_written)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !path) return fail(C2D_ESTATE, "context not initialised / NULL path");
    int64_t n = 0;
    int rc = c2d_census_count(h, &n);
    if (rc) return rc;
    std::vector<double> d((size_t)std::max<int64_t>(6 * n, 1));
    std::vector<int32_t> iv((size_t)std::max<int64_t>(6 * n, 1));
    int64_t got = 0;
    rc = c2d_census_export(h, d.data(), iv.data(), nullptr, n, &got);
    if (rc) return rc;
    FILE *f = fopen(path, "w");
    if (!f) return fail(C2D_EINVAL, "cannot open %s", path);
    char l1[6 * 14 + 1], l2[6 * 5 + 1];
    l1[84] = '\n'; l2[30] = '\n';
    for (int64_t p = 0; p < got; p++) {                           // census2d.f:21-28, formats 200 / 210
        for (int k = 0; k < 6; k++) fortran_e14_7(d[(size_t)(6 * p + k)], l1 + 14 * k);
        for (int k = 0; k < 6; k++) fortran_i5(iv[(size_t)(6 * p + k)], l2 + 5 * k);
        fwrite(l1, 1, 85, f);
        fwrite(l2, 1, 31, f);
    }
    if (fclose(f) != 0) return fail(C2D_EINVAL, "write to %s failed", path);
    if (n_written) *n_written = got;
    return 0;
}

int c2d_census_read(c2d_ctx *h, const char *path, int64_t entries)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !path) return fail(C2D_ESTATE, "context not initialised / NULL path");
    if (entries < 0) return fail(C2D_EINVAL, "entries = %lld", (long long)entries);
    if (entries > c->census_cap) return fail(C2D_ECENSUS, "census file holds %lld photons, census_capacity %lld", (long long)entries,
                                             (long long)c->census_cap);
    FILE *f = fopen(path, "r");
    if (!f) return fail(C2D_EINVAL, "cannot open %s", path);
    std::vector<double> d((size_t)std::max<int64_t>(6 * entries, 1));
    std::vector<int32_t> iv((size_t)std::max<int64_t>(6 * entries, 1));
    char line[256];
    for (int64_t p = 0; p < entries; p++) {                       // census2d.f:59-66: ndxout of p###_misc.dat says how many
        if (!fgets(line, sizeof line, f)) { fclose(f); return fail(C2D_EINVAL, "%s: end of file at photon %lld of %lld", path, (long long)p, (long long)entries); }
        const size_t len = strcspn(line, "\r\n");
        memset(line + len, ' ', sizeof line - len - 1);
        line[sizeof line - 1] = 0;
        for (int k = 0; k < 6; k++)
            if (!parse_e_field(line + 14 * k, 14, &d[(size_t)(6 * p + k)])) { fclose(f); return fail(C2D_EINVAL, "%s: bad number in photon %lld", path, (long long)p); }
        if (!fgets(line, sizeof line, f)) { fclose(f); return fail(C2D_EINVAL, "%s: end of file at photon %lld of %lld", path, (long long)p, (long long)entries); }
        const size_t len2 = strcspn(line, "\r\n");
        memset(line + len2, ' ', sizeof line - len2 - 1);
        for (int k = 0; k < 6; k++) {
            char fld[6];
            memcpy(fld, line + 5 * k, 5); fld[5] = 0;
            iv[(size_t)(6 * p + k)] = (int32_t)atoi(fld);
        }
    }
    fclose(f);
    return c2d_census_import(h, d.data(), iv.data(), nullptr, entries);
}

}  // extern "C"
