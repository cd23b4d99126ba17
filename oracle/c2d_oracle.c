/*
 * c2d_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY, see c2d_oracle.h).
 *
 * Part 1: state, RNGs, comtot/dilog, nth2d, compb2d, get_bin, imcleak, imctrk2d,
 * vol_calc, field_calc.  Every function follows the FORTRAN it cites line by line
 * (paths relative to /root/reference/src_20121026/); the compile must not
 * contract a*b+c into FMAs (-ffp-contract=off) so that the arithmetic is the
 * reference's.
 *
 * REAL*4 literals in the FORTRAN (e.g. `1.00001*rmin`, `0.999999*r(k)`, `.98`)
 * are single precision promoted to double; they are spelled (double)1.00001f here.
 */
#include "c2d_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.1415926536e0          /* general.pa:25 */
#define RAD_CP 3.333564097e-11     /* general.pa:24 */
#define C_LIGHT 2.9979245620e10    /* general.pa:26 */

#define ZI(j, k) (((j) - 1) * s->nr + ((k) - 1))
#define ZZ(j) (s->z[(j) - 1])
#define RR(k) (s->r[(k) - 1])

/* ------------------------------------------------------------------ life cycle */
orc_state *orc_create(int nz, int nr, int n_vol, int num_nt, int nphfield, int n_gg, int nphtotal,
                      int nph_lc, int nmu, int64_t cens_cap, int64_t ev_cap)
{
    orc_state *s = (orc_state *)calloc(1, sizeof(orc_state));
    s->nz = nz; s->nr = nr; s->nzone = nz * nr; s->n_vol = n_vol; s->num_nt = num_nt;
    s->nphfield = nphfield; s->n_gg = n_gg; s->nphtotal = nphtotal; s->nph_lc = nph_lc; s->nmu = nmu;
    int nzone = s->nzone;
    (void)nzone;
#define X(name, type, count) s->name = (type *)calloc((size_t)(count) + 1, sizeof(type));
    ORC_FIELDS(X)
#undef X
    s->cens_cap = cens_cap; s->ev_cap = ev_cap;
    s->dbufout = (double *)calloc((size_t)6 * cens_cap + 6, sizeof(double));
    s->dbufin = (double *)calloc((size_t)6 * cens_cap + 6, sizeof(double));
    s->ibufout = (int *)calloc((size_t)6 * cens_cap + 6, sizeof(int));
    s->ibufin = (int *)calloc((size_t)6 * cens_cap + 6, sizeof(int));
    s->sidout = (uint64_t *)calloc((size_t)cens_cap + 1, sizeof(uint64_t));
    s->sidin = (uint64_t *)calloc((size_t)cens_cap + 1, sizeof(uint64_t));
    s->events = (double *)calloc((size_t)7 * ev_cap + 7, sizeof(double));
    s->split1 = s->split2 = s->split3 = 1; s->spl3_trg = 10;
    s->rng.mode = 2; s->rand_switch = 2;
    s->lnL = 20.0;
    return s;
}

void orc_destroy(orc_state *s)
{
    if (!s) return;
#define X(name, type, count) free(s->name);
    ORC_FIELDS(X)
#undef X
    free(s->dbufout); free(s->dbufin); free(s->ibufout); free(s->ibufin);
    free(s->sidout); free(s->sidin); free(s->events);
    free(s);
}

void *orc_array(orc_state *s, const char *name, int64_t *count, int *elem_size)
{
    int nzone = s->nzone, nz = s->nz, nr = s->nr, n_vol = s->n_vol, num_nt = s->num_nt;
    int nphfield = s->nphfield, n_gg = s->n_gg, nphtotal = s->nphtotal;
    (void)nzone; (void)nz; (void)nr; (void)n_vol; (void)num_nt; (void)nphfield; (void)n_gg; (void)nphtotal;
#define X(nm, type, cnt) if (!strcmp(name, #nm)) { *count = (cnt); *elem_size = (int)sizeof(type); return s->nm; }
    ORC_FIELDS(X)
#undef X
    if (!strcmp(name, "dbufout")) { *count = 6 * s->cens_cap; *elem_size = 8; return s->dbufout; }
    if (!strcmp(name, "dbufin")) { *count = 6 * s->cens_cap; *elem_size = 8; return s->dbufin; }
    if (!strcmp(name, "ibufout")) { *count = 6 * s->cens_cap; *elem_size = 4; return s->ibufout; }
    if (!strcmp(name, "ibufin")) { *count = 6 * s->cens_cap; *elem_size = 4; return s->ibufin; }
    if (!strcmp(name, "sidout")) { *count = s->cens_cap; *elem_size = 8; return s->sidout; }
    if (!strcmp(name, "sidin")) { *count = s->cens_cap; *elem_size = 8; return s->sidin; }
    if (!strcmp(name, "events")) { *count = 7 * s->ev_cap; *elem_size = 8; return s->events; }
    return NULL;
}

#define ORC_SCALARS_I(X) \
    X(pair_switch) X(cr_sent) X(spec_switch) X(rand_switch) X(ncycle) X(split1) X(split2) X(split3) \
    X(spl3_trg) X(inj_switch) X(inj_dis) X(g2var_switch) X(pick_sw) X(cf_sentinel) X(fp_dead_terms) \
    X(nph_lc) X(nmu) X(nphtotal) X(fatal) X(idead) X(star_switch)
#define ORC_SCALARS_L(X) \
    X(incounter) X(outcounter) X(cmcount1) X(cmcount2) X(lkcount) X(n_multiple) X(n_killed) \
    X(n_segments) X(n_trials) X(n_comtot) X(ndxout) X(ndxin) X(nev)
#define ORC_SCALARS_D(X) \
    X(rmin) X(zmin) X(time) X(dt) X(inj_g1) X(inj_g2) X(inj_p) X(inj_t) X(inj_gg) X(inj_sigma) \
    X(inj_L) X(inj_v) X(pick_rate) X(r_esc) X(r_acc) X(lnL) X(r_flare) X(z_flare) X(t_flare) \
    X(sigma_r) X(sigma_z) X(sigma_t) X(flare_amp) X(E_tot_old) X(E_tot_new) X(hr_total) \
    X(hr_st_total) X(f_t_implicit) X(dT_max) X(dr) X(dz) X(t_bound)

int orc_set_scalar(orc_state *s, const char *name, double v)
{
#define X(nm) if (!strcmp(name, #nm)) { s->nm = (int)v; if (!strcmp(#nm, "rand_switch")) s->rng.mode = (int)v; return 0; }
    ORC_SCALARS_I(X)
#undef X
#define X(nm) if (!strcmp(name, #nm)) { s->nm = (int64_t)v; return 0; }
    ORC_SCALARS_L(X)
#undef X
#define X(nm) if (!strcmp(name, #nm)) { s->nm = v; return 0; }
    ORC_SCALARS_D(X)
#undef X
    return -1;
}

double orc_get_scalar(orc_state *s, const char *name)
{
#define X(nm) if (!strcmp(name, #nm)) return (double)s->nm;
    ORC_SCALARS_I(X)
    ORC_SCALARS_L(X)
    ORC_SCALARS_D(X)
#undef X
    if (!strcmp(name, "ndraw")) return (double)s->rng.ndraw;
    return NAN;
}

/* ------------------------------------------------------------------ RNG */
void orc_ran1_seed(orc_state *s, int idum, const int *iv, int iy)
{
    s->rng.idum = idum; s->rng.iy = iy;
    memcpy(s->rng.iv, iv, sizeof(int) * 32);
}
void orc_ran1_get(orc_state *s, int *idum, int *iv, int *iy)
{
    *idum = s->rng.idum; *iy = s->rng.iy;
    memcpy(iv, s->rng.iv, sizeof(int) * 32);
}

/* rand.f:322-359 -- NR ran1 with REAL*4 arithmetic and the 7-digit truncation of l.354 */
static double ran1(orc_rng *g)
{
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836, NTAB = 32;
    const int NDIV = 1 + (IM - 1) / NTAB;
    const float AM = 1.0f / (float)IM;
    const float RNMX = 1.0f - 1.2e-7f;
    int j, k;
    if (g->idum <= 0 || g->iy == 0) {
        g->idum = (-g->idum > 1) ? -g->idum : 1;
        for (j = NTAB + 8; j >= 1; j--) {
            k = g->idum / IQ;
            g->idum = IA * (g->idum - k * IQ) - IR * k;
            if (g->idum < 0) g->idum += IM;
            if (j <= NTAB) g->iv[j - 1] = g->idum;
        }
        g->iy = g->iv[0];
    }
    k = g->idum / IQ;
    g->idum = IA * (g->idum - k * IQ) - IR * k;
    if (g->idum < 0) g->idum += IM;
    j = 1 + g->iy / NDIV;
    g->iy = g->iv[j - 1];
    g->iv[j - 1] = g->idum;
    volatile float t = AM * (float)g->iy;
    if (t > RNMX) t = RNMX;
    volatile float t7 = t * 1.e7f;
    return (double)(int)t7 / 1.e7;
}

void orc_philox_key(orc_state *s, uint32_t seed, uint32_t cycle)
{
    s->rng.key[0] = seed; s->rng.key[1] = cycle;
}

/* Philox-4x32-10 (Salmon et al. 2011), the published constants. */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream-id derivation: child = mix64(parent, index) (murmur3 finaliser over a combine) */
uint64_t orc_mix64(uint64_t a, uint64_t b)
{
    uint64_t x = a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}

/* draw d of stream sid: word pair (d&1) of philox(ctr = (sid_lo, sid_hi, d>>1, 0)); 52-bit uniform in (0,1) */
double orc_philox_uniform(const uint32_t key[2], uint64_t sid, uint32_t d)
{
    uint32_t ctr[4] = {(uint32_t)sid, (uint32_t)(sid >> 32), d >> 1, 0u}, o[4];
    orc_philox4x32(ctr, key, o);
    uint64_t w = (d & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    return ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

/* Philox replay mode: the xi of the trial flights of a history with trial stream S.  u_t = (K_t + 0.5) 2^-52 with
 * K_t = coarse8 << 44 | fine44, three independent Philox outputs:
 *   which trials have coarse = 0xFF ("candidates": 1 trial in 256, independently) is decided by the GAPS between them,
 *       gap_j = floor(ln U_j * (1 / ln(255/256))),  U_j = word pair (j & 1) of philox(ctr = (S_lo, S_hi, j >> 1, 3)),
 *       candidate_j = candidate_{j-1} + 1 + gap_j  (candidate_{-1} = -1): a Bernoulli(1/256) process has geometric gaps;
 *   every other trial has coarse = (w * 255) >> 32 in [0, 254], w = word (t & 3) of philox(ctr = (S_lo, S_hi, t >> 2, 0));
 *   fine = word 0 and the low 12 bits of word 1 of philox(ctr = (S_lo, S_hi, t, 1)).
 * A trial collides only if u exceeds a threshold close to 1, so for an optically thin ray (tau < 1/256) only the
 * candidates need to be looked at: the CUDA path jumps from candidate to candidate instead of drawing split1 numbers.
 * The generator must be asked for non-decreasing t. */
typedef struct { uint64_t S; int64_t next_c; uint32_t j; } orc_trialgen;
#define ORC_INV_LN_Q (-255.49967384180653)     /* 1 / ln(255/256) */
static void trialgen_advance(const uint32_t key[2], orc_trialgen *g)
{
    uint32_t c[4] = {(uint32_t)g->S, (uint32_t)(g->S >> 32), g->j >> 1, 3u}, o[4];
    orc_philox4x32(c, key, o);
    const uint64_t w = (g->j & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    const double U = ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0);
    const double gap = floor(log(U) * ORC_INV_LN_Q);
    g->j++;
    g->next_c = (gap > 1.e9) ? (int64_t)1 << 40 : g->next_c + 1 + (int64_t)gap;
}
static void trialgen_init(const uint32_t key[2], orc_trialgen *g, uint64_t S)
{
    g->S = S; g->j = 0; g->next_c = -1;
    trialgen_advance(key, g);
}
static double trialgen_uniform(const uint32_t key[2], orc_trialgen *g, uint32_t t)
{
    uint32_t c1[4] = {(uint32_t)g->S, (uint32_t)(g->S >> 32), t, 1u}, f[4];
    uint64_t coarse;
    while (g->next_c < (int64_t)t) trialgen_advance(key, g);
    if (g->next_c == (int64_t)t) coarse = 0xFFu;
    else {
        uint32_t c0[4] = {(uint32_t)g->S, (uint32_t)(g->S >> 32), t >> 2, 0u}, o[4];
        orc_philox4x32(c0, key, o);
        coarse = ((uint64_t)o[t & 3u] * 255u) >> 32;
    }
    orc_philox4x32(c1, key, f);
    const uint64_t fine = (((uint64_t)(f[1] & 0xFFFu)) << 32) | f[0];
    return ((double)((coarse << 44) | fine) + 0.5) * (1.0 / 4503599627370496.0);
}
/* stand-alone form (tests): u_t for t = 0 .. n-1 */
void orc_trial_uniforms(const uint32_t key[2], uint64_t S, int n, double *u)
{
    orc_trialgen g;
    trialgen_init(key, &g, S);
    for (int t = 0; t < n; t++) u[t] = trialgen_uniform(key, &g, (uint32_t)t);
}

/* Philox replay mode: the uniform of the k-th absorption-point draw (imctrk2d.f:451-459) of trial flight t:
 * word pair (t & 1) of philox(ctr = (S_lo, S_hi, t >> 1, 2 + k)) -- two trials share a call. */
double orc_abs_uniform(const uint32_t key[2], uint64_t S, uint32_t t, uint32_t k)
{
    uint32_t c[4] = {(uint32_t)S, (uint32_t)(S >> 32), t >> 1, 2u + k}, o[4];
    orc_philox4x32(c, key, o);
    const uint64_t w = (t & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    return ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

/* ---- rand_switch = 1: Knuth's lagged-Fibonacci generator (The Art of Computer Programming vol. 2, 3rd ed., 3.6: ranf_array /
 * ranf_start with lags 100 and 37, including the 2002 modification), as rand.f:230-316 has it, and the reference's integer
 * variant with lags 19937 / 7083 that deals out one seed per zone and per boundary ring (seed_zone, rand.f:9-88). ---- */
#define KN_KK 100
#define KN_LL 37
static double frac1(double y) { return y - (double)(int)y; }          /* Y - IDINT(Y), Y >= 0 */

/* rand.f:235-257 RNFARR: n >= 100 numbers into aa, state advanced */
static void kn_array(orc_rng *g, double *aa, int n)
{
    int j;
    for (j = 0; j < KN_KK; j++) aa[j] = g->ranx[j];
    for (j = KN_KK; j < n; j++) aa[j] = frac1(aa[j - KN_KK] + aa[j - KN_LL]);
    for (j = 0; j < KN_LL; j++) g->ranx[j] = frac1(aa[n + j - KN_KK] + aa[n + j - KN_LL]);
    for (j = KN_LL; j < KN_KK; j++) g->ranx[j] = frac1(aa[n + j - KN_KK] + g->ranx[j - KN_LL]);
}

/* rand.f:262-316 RNFSTR */
static void kn_start(orc_rng *g, int seed)
{
    const int MM = 1 << 30, KKK = 2 * KN_KK - 1;
    const double ULP = 1.0 / 4503599627370496.0;
    double u[2 * KN_KK - 1 + KN_KK], ss;            /* the 10 warm-up calls write KKK numbers into u */
    int j, s, t, sseed;
    sseed = (seed < 0) ? MM - 1 - ((-1 - seed) % MM) : seed % MM;
    ss = 2.0 * ULP * (double)(sseed + 2);
    for (j = 0; j < KN_KK; j++) {
        u[j] = ss;
        ss = ss + ss;
        if (ss >= 1.0) ss = ss - 1.0 + 2 * ULP;
    }
    for (j = KN_KK; j < KKK; j++) u[j] = 0.0;
    u[1] = u[1] + ULP;
    s = sseed; t = 70 - 1;
    while (t > 0) {
        for (j = KN_KK - 1; j >= 1; j--) { u[2 * j] = u[j]; u[2 * j - 1] = 0.0; }
        for (j = KKK - 1; j >= KN_KK; j--) {
            u[j - (KN_KK - KN_LL)] = frac1(u[j - (KN_KK - KN_LL)] + u[j]);
            u[j - KN_KK] = frac1(u[j - KN_KK] + u[j]);
        }
        if (s % 2 == 1) {
            for (j = KN_KK - 1; j >= 0; j--) u[j + 1] = u[j];
            u[0] = u[KN_KK];
            u[KN_LL] = frac1(u[KN_LL] + u[KN_KK]);
        }
        if (s != 0) s = s / 2; else t = t - 1;
    }
    for (j = 0; j < KN_LL; j++) g->ranx[j + KN_KK - KN_LL] = u[j];
    for (j = KN_LL; j < KN_KK; j++) g->ranx[j - KN_LL] = u[j];
    for (j = 0; j < 10; j++) kn_array(g, u, KKK);
}

/* rand.f:99-186 initialize_rand / _zrand / _rrand: a fresh generator per zone job, per boundary ring, per census photon */
static void kn_initialize(orc_state *s, int seed)
{
    if (s->rng.mode != 1) return;
    kn_start(&s->rng, seed);
    kn_array(&s->rng, s->rng.randlist, 10000);
    s->rng.randcounter = 1;
}

/* rand.f:9-88 seed_zone: the integer generator (lags 19937 / 7083, mod 2^30) deals seeds(j,k), zseeds(j); rseeds(k) all get
 * the same value nz nr + nz + 1 (l.81-84 never advance i); *seed_io leaves as seeds(1,1) (l.85), so the next step differs. */
void orc_seed_zone(orc_state *s, int *seed_io)
{
    enum { KK = 19937, LL = 7083, KKK = 2 * KK - 1 };
    const int MM = 1 << 30;
    int *x = (int *)malloc(sizeof(int) * (size_t)(KKK + 1));
    int j, k, i, ss, t, sseed, seed = *seed_io;
    sseed = (seed < 0) ? MM - 1 - ((-1 - seed) % MM) : seed % MM;
    ss = sseed - sseed % 2 + 2;
    for (j = 0; j < KK; j++) {
        x[j] = ss;
        ss = ss + ss;
        if (ss >= MM) ss = ss - MM + 2;
    }
    for (j = KK; j <= KKK; j++) x[j] = 0;
    x[1] = x[1] + 1;
    ss = sseed; t = 70 - 1;
    while (t > 0) {
        for (j = KK - 1; j >= 1; j--) { x[2 * j] = x[j]; x[2 * j - 1] = 0; }
        for (j = KKK - 1; j >= KK; j--) {
            x[j - (KK - LL)] -= x[j]; if (x[j - (KK - LL)] < 0) x[j - (KK - LL)] += MM;
            x[j - KK] -= x[j];        if (x[j - KK] < 0) x[j - KK] += MM;
        }
        if (ss % 2 == 1) {
            for (j = KK - 1; j >= 0; j--) x[j + 1] = x[j];
            x[0] = x[KK];
            x[LL] -= x[KK]; if (x[LL] < 0) x[LL] += MM;
        }
        if (ss != 0) ss = ss / 2; else t = t - 1;
    }
    i = 0;
    for (j = 1; j <= s->nz; j++)
        for (k = 1; k <= s->nr; k++) s->seeds[ZI(j, k)] = x[i++];
    for (j = 1; j <= s->nz; j++) s->zseeds[j - 1] = x[i++];
    for (k = 1; k <= s->nr; k++) s->rseeds[k - 1] = i + 1;
    *seed_io = s->seeds[ZI(1, 1)];
    free(x);
}
void orc_knuth_get(orc_state *s, double *ranx, double *randlist, int *randcounter)
{
    memcpy(ranx, s->rng.ranx, sizeof s->rng.ranx);
    memcpy(randlist, s->rng.randlist, sizeof s->rng.randlist);
    *randcounter = s->rng.randcounter;
}
void orc_knuth_set(orc_state *s, const double *ranx, const double *randlist, int randcounter)
{
    memcpy(s->rng.ranx, ranx, sizeof s->rng.ranx);
    memcpy(s->rng.randlist, randlist, sizeof s->rng.randlist);
    s->rng.randcounter = randcounter;
}

/* rand.f:195-223 fibran / Philox extension */
double orc_fibran(orc_state *s)
{
    orc_rng *g = &s->rng;
    g->ndraw++;
    if (g->mode == 2) return ran1(g);
    if (g->mode == 1) {                              /* l.211-221: refill when the list is used up */
        if (g->randcounter >= 10000) { kn_array(g, g->randlist, 10000); g->randcounter = 1; }
        else g->randcounter++;
        return g->randlist[g->randcounter - 1];
    }
    if (g->has_pending) { g->has_pending = 0; return g->pending; }
    return orc_philox_uniform(g->key, g->sid, g->d++);
}

static void rng_stream(orc_state *s, uint64_t sid)
{
    s->rng.sid = sid; s->rng.d = 0; s->rng.has_pending = 0;
}

/* ------------------------------------------------------------------ comtot */
/* comtot2d.f:356-433, CERNLIB C332 dilogarithm */
double orc_dilog(double x)
{
    static const double C[20] = {0.42996693560813697e0, 0.40975987533077105e0,
        -0.01858843665014592e0, 0.00145751084062268e0, -0.00014304184442340e0,
        0.1588415541880e-4, -0.190784959387e-5, 0.024195180854e-5,
        -0.003193341274e-5, 0.000434545063e-5, -0.000060578480e-5,
        0.000008612098e-5, -0.000001244332e-5, 0.000000182256e-5,
        -0.000000027007e-5, 0.000000004042e-5, -0.000000000610e-5,
        0.000000000093e-5, -0.000000000014e-5, 0.000000000002e-5};
    const double HF = 0.5, PI2 = PI * PI, PI3 = PI2 / 3, PI6 = PI2 / 6, PI12 = PI2 / 12;
    double T, H, Y, S, A, ALFA, B1, B2, B0 = 0;
    if (x == 1) return PI6;
    if (x == -1) return -PI12;
    T = -x;
    if (T <= -2) {
        Y = -1 / (1 + T); S = 1;
        B1 = log(-T); B2 = log(1 + 1 / T);
        A = -PI3 + HF * (B1 * B1 - B2 * B2);
    } else if (T < -1) {
        Y = -1 - T; S = -1;
        A = log(-T);
        A = -PI6 + A * (A + log(1 + 1 / T));
    } else if (T <= -0.5) {
        Y = -(1 + T) / T; S = 1;
        A = log(-T);
        A = -PI6 + A * (-HF * A + log(1 + T));
    } else if (T < 0) {
        Y = -T / (1 + T); S = -1;
        B1 = log(1 + T);
        A = HF * B1 * B1;
    } else if (T <= 1) {
        Y = T; S = 1; A = 0;
    } else {
        Y = 1 / T; S = -1;
        B1 = log(T);
        A = PI6 + HF * B1 * B1;
    }
    H = Y + Y - 1;
    ALFA = H + H;
    B1 = 0; B2 = 0;
    for (int i = 19; i >= 0; i--) {
        B0 = C[i] + ALFA * B1 - B2;
        B2 = B1; B1 = B0;
    }
    return -(S * (B0 - H * B2) + A);
}

/* comtot2d.f:337-352 */
double orc_intg_v(double x)
{
    double intg_1 = -x / 2.0 + 0.5 / (1.0 + x);
    double intg_2 = 4.0 * orc_dilog(-x);
    double intg_3 = (9.0 + x + 8.0 / x) * log(1.0 + x);
    return (intg_1 + intg_2 + intg_3);
}

/* comtot2d.f:219-247 (icoms = 6, the only live branch; enexc is numerically dead: SURVEY 0.6) */
double orc_comtot(orc_state *s, int j, int k, double xnuc)
{
    const double emasskev = 5.11e2, sigthom = 6.6516e-25;
    const double *f = s->f_nt + (size_t)ZI(j, k) * s->num_nt;
    double cosig = 0.0, x = xnuc / emasskev;
    s->n_comtot++;
    for (int i = 1; i <= s->num_nt - 1; i++) {
        double gamma0 = s->gnt[i - 1] + 1.0;
        double betta = sqrt(1.0 - 1.0 / (gamma0 * gamma0));
        double sigma_E;
        if (x * gamma0 * (1 + betta) < 1.e-2)
            sigma_E = sigthom * (1.0 - 2.0 * x * gamma0);
        else
            sigma_E = 9.375e-2 * sigthom / (gamma0 * gamma0) / betta / (x * x) *
                      (orc_intg_v(2 * gamma0 * (1 + betta) * x) - orc_intg_v(2 * gamma0 * (1 - betta) * x));
        cosig = cosig + sigma_E * f[i - 1] * (s->gnt[i] - s->gnt[i - 1]);
    }
    if (cosig < 1.e-40) return 1.e-40;
    return s->n_e[ZI(j, k)] * cosig;
}

/* ------------------------------------------------------------------ nth2d / compb2d */
/* nontherm2d.f:159-183 */
void orc_nth2d(orc_state *s, int j, int k, double *gamm, double *betb, int *ip)
{
    const double *P = s->Pnt + (size_t)ZI(j, k) * s->num_nt;
    double rnum = orc_fibran(s);
    rnum = (double)(int)(rnum * 1.e6) / 1.e6 + 1.e-6 * orc_fibran(s);
    int i;
    for (i = 2; i <= s->num_nt; i++)
        if (P[i - 1] > rnum) break;
    if (i > s->num_nt) i = s->num_nt;     /* the FORTRAN would read gnt(num_nt+1); never hit with Pnt(num_nt)=1 */
    *gamm = sqrt(s->gnt[i - 1] * s->gnt[i - 2]) + 1.0;
    *betb = sqrt(1.0 - 1.0 / (*gamm * *gamm));
    s->nelectron[i]++;
    *ip = i;
}

/* spectrum-group binary search shared by compb_2d.f:249-274, imcleak2d.f:345-370,
 * imcvol2d_para.f:334-359; `hubot_f`/`hutop_f` are the edge fudge factors, which are
 * REAL*8 literals in compb2d and REAL*4 literals in the other two. */
static int sp_group(const orc_state *s, double x, double hubot_f, double hutop_f, int top_value)
{
    int jbot = 1, jtop = s->nphtotal + 1, jmid;
    double hubot = hubot_f * s->hu[0], hutop = hutop_f * s->hu[jtop - 1];
    if (x >= hutop) return top_value;
    if (x <= hubot) return 0;
    for (;;) {
        jmid = (jbot + jtop) / 2;
        if (jmid == jbot) break;
        if (x == s->hu[jmid - 1]) break;
        if (x < s->hu[jmid - 1]) jtop = jmid; else jbot = jmid;
    }
    return jmid;
}
static int lc_group(const orc_state *s, double x)
{
    for (int m = 1; m <= s->nph_lc; m++)
        if (x > s->Elcmin[m - 1] && x <= s->Elcmax[m - 1]) return m;
    return 0;
}
static int mu_group(const orc_state *s, double w)
{
    for (int n = 1; n <= s->nmu; n++)
        if (w <= s->mu[n - 1]) return n;
    return s->nmu;
}

#define CLAMP8(v) do { if ((v) > 9.9999999e-1) (v) = 9.9999999e-1; if ((v) < -9.9999999e-1) (v) = -9.9999999e-1; } while (0)

/* compb_2d.f:1-318 */
void orc_compb2d(orc_state *s, int *i_gam)
{
    const double fuzz = 1.e-10, emasskev = 5.11e2;
    double znu = s->xnu / emasskev;
    double gamm, betb, omeg, tl, tr, znue, xknot, znue3, betz, gamz, xxx, sz, games, phat, znues;
    double wa, wb, swa, cazes, omege, omeges, omegs, znus, gams, cazs, wmus, xnus, cosdphi, dphi, phis;
    (void)omegs;
L100:
    orc_nth2d(s, s->jph, s->kph, &gamm, &betb, i_gam);
    omeg = 2.0 * orc_fibran(s) - 1.0;
    CLAMP8(omeg);
    tl = orc_fibran(s);
    tr = 0.5 * (1.0 - betb * omeg);
    if (tl > tr) omeg = -omeg;
    CLAMP8(omeg);
    znue = (1. - betb * omeg) * znu * gamm;
    if (znue < 1.e-10) goto L100;
    if (znue <= 1.e-2) {
        xknot = 1.0 - znue * (2.0 - znue * (5.2 - znue * (13.3 - 1.144e3 * znue / 3.5e1)));
    } else {
        znue3 = znue * znue * znue;
        betz = 1.0 + 2.0 * znue;
        gamz = znue * (znue - 2.0) - 2.0;
        xxx = 4.0 * znue + 2.0 * znue3 * (1.0 + znue) / (betz * betz) + gamz * log(betz);
        xknot = 3.75e-1 * xxx / znue3;
    }
    if (orc_fibran(s) > xknot) goto L100;
    betz = 1.0 + 2.0 * znue;
L200:
    sz = (1.0 + 2.0 * znue * orc_fibran(s)) / betz;
    games = 1.0 + (1.0 - 1.0 / sz) / znue;
    if ((1.0 - games * games) < 0.) goto L200;       /* goto 202 == goto 200 */
    tr = games * games - 1.0 + sz + 1.0 / sz;
    phat = betz + 1.0 / betz;
    if (orc_fibran(s) * phat > tr) goto L200;
    znues = znue * sz;
L210:
    wa = orc_fibran(s);
    wb = 2.0 * orc_fibran(s) - 1.0;
    swa = wa * wa + wb * wb;
    if (swa >= 1.0 || swa <= 1.e-20) goto L210;
    cazes = (wa * wa - wb * wb) / swa;
    omege = (omeg - betb) / (1. - betb * omeg);
    CLAMP8(omege);
    omeges = games * omege + cazes * sqrt((1. - omege * omege + fuzz) * (1. - games * games));
    CLAMP8(omeges);
    omegs = (omeges + betb) / (1.0 + omeges * betb);
    CLAMP8(omegs);
    znus = (1.0 + betb * omeges) * gamm * znues;
    gams = 1.0 - (znue - znues) / (znu * znus);
    CLAMP8(gams);
L220:
    wa = orc_fibran(s);
    wb = 2.0 * orc_fibran(s) - 1.0;
    swa = wa * wa + wb * wb;
    if (swa >= 1.0 || swa <= 1.e-20) goto L220;
    cazs = (wa * wa - wb * wb) / swa;
    CLAMP8(cazs);
    wmus = s->wmu * gams + cazs * sqrt((1.0 - gams * gams) * (1.0 - s->wmu * s->wmu + fuzz));
    CLAMP8(wmus);
    xnus = znus * emasskev;
    cosdphi = (gams - s->wmu * wmus) / sqrt((1.0 - s->wmu * s->wmu) * (1.0 - wmus * wmus));
    CLAMP8(cosdphi);
    dphi = acos(cosdphi);
    phis = s->phi + dphi;
    /* re-binning, l.249-302 */
    s->jgpsp = sp_group(s, xnus, 1.000001e0, .999999e0, 0);
    s->jgplc = lc_group(s, xnus);
    s->jgpmu = mu_group(s, wmus);
    s->ew = s->ew * xnus / s->xnu;
    s->xnu = xnus;
    s->wmu = wmus;
    s->phi = phis;
}

/* ------------------------------------------------------------------ imcleak */
/* imcleak2d.f:329-405 */
void orc_get_bin(orc_state *s)
{
    s->jgpsp = sp_group(s, s->xnu, (double)1.000001f, (double).999999f, 0);
    s->jgplc = lc_group(s, s->xnu);
    s->jgpmu = mu_group(s, s->wmu);
}

static void event_write(orc_state *s)
{   /* imcleak2d.f:171,181: write(nunit_evt,105) t_bound,xnu,ew,rpre,zpre,wmu,phi */
    if (s->nev >= s->ev_cap) { s->nev++; return; }
    double *e = s->events + 7 * s->nev++;
    e[0] = s->t_bound; e[1] = s->xnu; e[2] = s->ew; e[3] = s->rpre; e[4] = s->zpre; e[5] = s->wmu; e[6] = s->phi;
}
static void escape_tally(orc_state *s)
{   /* imcleak2d.f:171-175 (and 205-209, 304-310) */
    event_write(s);
    if (s->jgplc > 0)
        s->edout[(s->jgpmu - 1) * ORC_NLCMAX + (s->jgplc - 1)] += s->ew / s->dt;
    if (s->jgpsp > 0 && s->spec_switch == 0)
        s->fout[(s->jgpmu - 1) * s->nphtotal + (s->jgpsp - 1)] += s->ew;
}

/* imcleak2d.f:126-146 == 224-241: energy of the Compton-reflected photon from the Green's-function tables of ref_matrix.f
 * (P_ref = cumulative distribution of the outgoing bin per incoming bin, W_abs = surviving weight fraction; inputs here, the
 * host computes them once in setup, setup2d.f:40-41).  Two uniforms.  Returns W_abs(n_out, n_in); s->xnu is replaced. */
static double reflect_energy(orc_state *s)
{
    const int n_ref = ORC_NREF;
    int n_in = 1, n_out = 1;
    while (s->E_ref[n_in - 1] < s->xnu && n_in < n_ref) n_in++;
    {
        const double rnum = orc_fibran(s);
        while (s->P_ref[(size_t)(n_in - 1) * n_ref + (n_out - 1)] < rnum && n_out < n_ref) n_out++;
    }
    if (n_out > 1) s->xnu = s->E_ref[n_out - 2] + orc_fibran(s) * (s->E_ref[n_out - 1] - s->E_ref[n_out - 2]);
    else s->xnu = s->E_ref[n_out - 1];
    return s->W_abs[(size_t)(n_in - 1) * n_ref + (n_out - 1)];
}

/* imcleak2d.f:2-320 with its reflection options (cr_sent): specular mirror at the lower surface (l.107-122:
 * cr_sent = 4, or cr_sent = 1 / 3 where the boundary has no temperature -- every shipped input), Compton reflection off a hot
 * lower boundary (l.124-165) and off the outer cylinder for photons heading down (cr_sent = 2 / 3, l.216-273). */
void orc_imcleak(orc_state *s)
{
    const double xnu_sv = s->xnu;
    if (s->kph == 0) {                       /* inner r boundary, l.71-88 */
        if (s->rmin > 1.e-10) {
            s->idead = 1;
            s->erlki[s->jph - 1] += s->ew;
        } else {
            s->idead = 0;
            s->phi = 1.e-6;
            s->kph = 1;
        }
        return;
    }
    if (!(s->jph > 0)) {                     /* lower surface, l.92-179 */
        if (s->tbbl[s->kph - 1] > 0.0) {
            s->Ed_in[s->kph - 1] += s->ew;
            s->erlkl[s->kph - 1] += s->ew;
        }
        if (s->cr_sent == 1 || s->cr_sent == 3 || s->cr_sent == 4) {       /* l.107-165 */
            if (s->tbbl[s->kph - 1] <= 0.0 || s->cr_sent == 4) {           /* specular reflection, l.113-122 */
                s->wmu = -s->wmu;
                if (s->jgpsp > 0 && s->spec_switch == 1)
                    s->fout[(s->jgpmu - 1) * s->nphtotal + (s->jgpsp - 1)] += s->ew;
                orc_get_bin(s);                                            /* label 60, l.160-165 */
                s->idead = 0;
                s->jph = 1;
                return;
            }
            {                                                               /* l.124-165 */
                const double w = reflect_energy(s);
                const double ewnew = s->ew * w * s->xnu / xnu_sv;
                s->Ed_ref[s->kph - 1] += ewnew;
                s->wmu = fabs(s->wmu);
                s->ew = ewnew;
            }
            orc_get_bin(s);
            s->idead = 0;
            s->jph = 1;
            return;
        }
        /* no reflection: photon exits.  The FORTRAN does not update t_bound here
         * (stale COMMON value of the previous escape); order-free Philox mode uses the
         * formula of the other faces. */
        if (s->rng.mode == 3) s->t_bound = s->time + s->dt - RAD_CP * s->dcen;
        if (s->ncycle > 0) escape_tally(s);
        s->idead = 1;
        return;
    }
    if (s->jph != s->nz + 1) {               /* outer r boundary, l.194-212 */
        s->erlko[s->jph - 1] += s->ew;
        if (!(s->wmu > 0. || s->cr_sent == 0 || s->cr_sent == 1 || s->cr_sent == 4)) {    /* l.197-198: else l.216-273 */
            /* Compton reflection at the outer disk: the photon is taken down to z = 0 along its direction, re-emitted with
             * the reflected energy and weight in a random upward direction, and recorded as escaped */
            const double w = reflect_energy(s);
            s->ew = s->ew * w * s->xnu / xnu_sv;
            if (fabs(s->wmu) > 1.e-6) {
                double f;
                s->t_bound = s->time + s->dt - RAD_CP * (s->dcen + s->zpre / s->wmu);
                s->zpre = 0.;
                f = s->zpre * sqrt(1. - s->wmu * s->wmu) / fabs(s->wmu);     /* zpre already zeroed (l.249-250): f = 0 */
                s->rpre = sqrt(s->rpre * s->rpre + f * f + 2. * s->rpre * f * cos(s->phi));
            } else {
                s->t_bound = 1.e20;
            }
            s->wmu = orc_fibran(s);
            orc_get_bin(s);
            if (s->ncycle > 0) escape_tally(s);
            s->idead = 1;
            return;
        }
        s->t_bound = s->time + s->dt - RAD_CP * s->dcen;
        if (s->ncycle > 0) escape_tally(s);
        s->idead = 1;
        return;
    }
    /* upper z surface, l.282-313 */
    s->erlku[s->kph - 1] += s->ew;
    s->t_bound = s->time + s->dt - RAD_CP * s->dcen;
    if (s->ncycle > 0 && s->wmu < (double)0.98f) escape_tally(s);
    s->idead = 1;
}

/* ------------------------------------------------------------------ imctrk2d */
typedef struct { double ew, dcen, xnu, zpre, rpre, wmu, phi; int jph, kph, jgpsp, jgplc, jgpmu; } photon_sv;
static void ph_save(const orc_state *s, photon_sv *p)
{
    p->ew = s->ew; p->dcen = s->dcen; p->xnu = s->xnu; p->zpre = s->zpre; p->rpre = s->rpre;
    p->wmu = s->wmu; p->phi = s->phi; p->jph = s->jph; p->kph = s->kph;
    p->jgpsp = s->jgpsp; p->jgplc = s->jgplc; p->jgpmu = s->jgpmu;
}
static void ph_restore(orc_state *s, const photon_sv *p)
{
    s->ew = p->ew; s->dcen = p->dcen; s->xnu = p->xnu; s->zpre = p->zpre; s->rpre = p->rpre;
    s->wmu = p->wmu; s->phi = p->phi; s->jph = p->jph; s->kph = p->kph;
    s->jgpsp = p->jgpsp; s->jgplc = p->jgplc; s->jgpmu = p->jgpmu;
}

#define TAG_TRIALS 0xA1ull
#define TAG_F0 0xF0ull

/* imctrk2d.f:8-716 */
void orc_imctrk2d(orc_state *s, int scat_flag)
{
    const int philox = (s->rng.mode == 3);
    const uint64_t S_entry = s->rng.sid;
    const uint64_t S_trials = orc_mix64(S_entry, TAG_TRIALS);
    int i_gam = 0, i, i_gg = 0, nscat, iitrk, ii, ii2;
    int knew, jnew, kbnd, inout, ikind, Eta_switch;
    double ewold, disp, psq, disbr, dpbsq, Zbnd, rbnd, xqsqleft, trld, trldb, dcol;
    double wtmin, Zr, Rr, kgg, pair_enhance, Eta, comac, sigabs, colmfp, mb_ran, denom, f;
    double sigsc, deleabs, ewnew, xabs, sstar, delpr, wmustar, rnew, znew;
    uint32_t kabs = 0;
    orc_trialgen tgen;
    photon_sv tsv, csv;
    /* comac_ar(jmax,kmax) cache of l.103,175-183 (zero = not yet computed) */
    double *comac_ar = NULL;

    if (s->fatal) return;
    s->idead = 0;
    wtmin = 1.e-10 * s->ew;                                  /* l.82,92 */
    if (scat_flag == -1) comac_ar = (double *)calloc((size_t)s->nzone, sizeof(double));
    nscat = 0;
    ph_save(s, &tsv);
    if (scat_flag == -1) tsv.ew = s->ew / s->split1;         /* l.109-113 */
    if (scat_flag == -1 && philox) trialgen_init(s->rng.key, &tgen, S_trials);

    for (iitrk = 1; iitrk <= s->split1; iitrk++) {           /* do 910 */
        ph_restore(s, &tsv);
        if (scat_flag == -1) {
            s->n_trials++;
            if (philox) {          /* trial stream: first draw comes from the history's trial-xi stream */
                rng_stream(s, orc_mix64(S_trials, (uint64_t)iitrk));
                s->rng.has_pending = 1;
                s->rng.pending = trialgen_uniform(s->rng.key, &tgen, (uint32_t)(iitrk - 1));
                kabs = 0;
            }
        }
    L100:
        if (scat_flag == 0) {
            mb_ran = 1.e-10;
        } else {
            mb_ran = orc_fibran(s);
            if (s->rand_switch == 2) mb_ran = (double)(int)(mb_ran * 1.e6) / 1.e6 + 1.e-6 * orc_fibran(s);
        }
        if (mb_ran > 0.) colmfp = -log(mb_ran); else goto L100;
    L110:
        sigabs = 1.e-40;
        if (s->ew < 1.e-40) goto L900;
        if (s->wmu > 1.0) s->wmu = 1.0;
        if (s->wmu < -1.0) s->wmu = -1.0;
        s->n_segments++;
        if (s->pair_switch == 1) pair_enhance = 1.0 + 2.0 * s->f_pair[ZI(s->jph, s->kph)];
        else pair_enhance = 1.0;
        if (scat_flag == -1) {
            if (comac_ar[ZI(s->jph, s->kph)] == 0.0) {
                comac = orc_comtot(s, s->jph, s->kph, s->xnu);
                comac_ar[ZI(s->jph, s->kph)] = comac;
            } else comac = comac_ar[ZI(s->jph, s->kph)];
        } else if (scat_flag == 1) {
            comac = orc_comtot(s, s->jph, s->kph, s->xnu);
        } else comac = 0.;
        sigsc = 1.0 * comac * pair_enhance;                  /* velfact = 1, l.203-205 */
        if (s->kph == 1) xqsqleft = s->rmin * s->rmin; else xqsqleft = RR(s->kph - 1) * RR(s->kph - 1);
        if (scat_flag != 0) dcol = colmfp / sigsc;
        else dcol = 100 * fmax(RR(s->nr), ZZ(s->nz));
        if (s->dcen <= dcol) { trld = s->dcen; ikind = 2; }
        else { trld = dcol; ikind = 3; }
        /* geometry, l.233-365 */
        Eta = cos(s->phi);
        if (s->phi <= PI && s->phi >= 1.e-10) Eta_switch = 1; else Eta_switch = -1;
        disp = Eta * s->rpre;
        psq = s->rpre * s->rpre * (1. - Eta * Eta);
        if (Eta < 0.0 && psq < xqsqleft) {
            s->incounter++;
            kbnd = s->kph - 1; inout = -1;
            rbnd = (s->kph > 1) ? RR(s->kph - 1) : s->rmin;
        } else {
            s->outcounter++;
            kbnd = s->kph; inout = 1;
            rbnd = RR(s->kph);
        }
        dpbsq = rbnd * rbnd - psq;
        if (dpbsq < 1.e-6) dpbsq = 1.e-6;
        disbr = (double)inout * sqrt(dpbsq) - disp;
        trldb = disbr / sqrt(1.0 - (s->wmu * s->wmu));
        f = disbr;
        Zr = s->zpre + s->wmu * trldb;
        {
            double zlow = (s->jph == 1) ? s->zmin : ZZ(s->jph - 1);
            if (Zr > ZZ(s->jph)) {                           /* upward z boundary */
                Zbnd = ZZ(s->jph); knew = s->kph; jnew = s->jph + 1;
                f = (Zbnd - s->zpre) * sqrt(1.0 - s->wmu * s->wmu) / s->wmu;
                Rr = sqrt(s->rpre * s->rpre + f * f + 2.0 * s->rpre * f * Eta);
                rbnd = Rr;
                trldb = sqrt(f * f + (Zbnd - s->zpre) * (Zbnd - s->zpre));
            } else if (Zr < zlow) {                          /* downward z boundary */
                Zbnd = zlow; knew = s->kph; jnew = s->jph - 1;
                f = (Zbnd - s->zpre) * sqrt(1.0 - s->wmu * s->wmu) / s->wmu;
                Rr = sqrt(s->rpre * s->rpre + f * f + 2.0 * s->rpre * f * Eta);
                rbnd = Rr;
                trldb = sqrt(f * f + (Zbnd - s->zpre) * (Zbnd - s->zpre));
            } else {                                         /* r boundary */
                knew = s->kph + inout; jnew = s->jph;
                Rr = (kbnd > 0) ? RR(kbnd) : s->rmin;
                rbnd = Rr; Zbnd = Zr;
            }
        }
        if (trldb < trld) {
            ikind = 1; trld = trldb; rnew = rbnd; znew = Zbnd;
        } else {
            jnew = s->jph; knew = s->kph;
            f = trld * sqrt(1.0 - s->wmu * s->wmu);
            rnew = sqrt(f * f + s->rpre * s->rpre + 2.0 * f * s->rpre * Eta);
            znew = s->zpre + trld * s->wmu;
        }
        for (i = 1; i <= s->n_vol - 1; i++)                   /* l.387-389 */
            if (s->xnu < s->E_ph[i]) break;
        if (s->pair_switch == 1)
            for (i_gg = 1; i_gg <= s->n_gg; i_gg++)
                if (s->xnu < s->E_gg[i_gg - 1]) break;
        sigabs = sigabs + pair_enhance * s->kappa_tot[(size_t)ZI(s->jph, s->kph) * s->n_vol + (i - 1)];
        double kgg_i = 0.0;   /* k_gg(i_gg,jph,kph); i_gg = n_gg+1 reads past the row (zero in every MPI-build run) */
        if (s->pair_switch == 1) {
            const double *kg = s->k_gg + (size_t)ZI(s->jph, s->kph) * s->n_gg;
            if (i_gg <= s->n_gg) kgg_i = kg[i_gg - 1];
            if (s->xnu < s->E_gg[0]) kgg = (s->xnu / s->E_gg[0]) * kg[0];
            else if (i_gg == s->n_gg) kgg = kg[s->n_gg - 1];
            else if (i_gg > s->n_gg) kgg = 0.0;
            else {
                double kgg0 = kg[i_gg - 1], kgg1 = kg[i_gg];
                kgg = kgg0 + (s->xnu - s->E_gg[i_gg - 1]) * (kgg1 - kgg0) / (s->E_gg[i_gg] - s->E_gg[i_gg - 1]);
            }
            sigabs = sigabs + kgg;
        }
        if (sigabs < 1.e-40) sigabs = 1.e-40;
        xabs = sigabs * trld;
        if (xabs < 100.) ewnew = s->ew * exp(-xabs); else ewnew = 0.;
        if (s->xnu > 47.0 && sigabs > 1.e-40 && s->pair_switch == 1)
            deleabs = (s->ew - ewnew) * (sigabs - kgg_i) / sigabs;
        else
            deleabs = s->ew - ewnew;
        if (deleabs < 1.e-50) deleabs = 1.e-50;
        if (xabs <= .00001) {
            wmustar = s->wmu;
        } else {
            do {
                if (philox && scat_flag == -1) mb_ran = orc_abs_uniform(s->rng.key, S_trials, (uint32_t)(iitrk - 1), kabs++);
                else mb_ran = orc_fibran(s);
            } while (!(mb_ran < (s->ew / deleabs)));                                /* l.451-456 */
            sstar = -log(1. - mb_ran * deleabs / s->ew) / sigabs;
            denom = sqrt(s->rpre * s->rpre + 2. * s->wmu * s->rpre * sstar + sstar * sstar);
            wmustar = (s->wmu * s->rpre + sstar) / denom;
        }
        delpr = deleabs * wmustar * C_LIGHT;
        if (scat_flag != 0) {
            s->edep[ZI(s->jph, s->kph)] += deleabs;
            s->prdep[ZI(s->jph, s->kph)] += delpr;
        }
        if (ewnew <= wtmin) { s->n_killed++; goto L900; }
        s->ew = ewnew;
        s->dcen = s->dcen - trld;
        Eta = (f + Eta * s->rpre) / rnew;                    /* l.478 */
        if (Eta > 1.0) Eta = 1.0;
        if (Eta < -1.0) Eta = -1.0;
        s->phi = acos(Eta);
        if (Eta_switch == -1) s->phi = 2.0 * PI - s->phi;
        s->rpre = rnew;
        s->zpre = znew;

        if (ikind == 1) {
            colmfp = colmfp - sigsc * trld;
            if (jnew == s->nz + 1 || jnew == 0 || knew == s->nr + 1 || knew == 0) {
                s->jph = jnew; s->kph = knew;
                if (scat_flag == -1) goto L900;
                orc_imcleak(s);
                s->lkcount++;
                if (s->idead == 1) goto L900;
                goto L110;
            } else {
                s->kph = knew; s->jph = jnew;
                goto L110;
            }
        } else if (ikind == 2 && scat_flag != -1) {          /* census, l.541-586 */
            int zi = ZI(s->jph, s->kph);
            s->npcen[zi]++;
            s->ecens[zi] += s->ew;
            if (s->pair_switch == 1) {
                for (i_gg = 1; i_gg <= s->n_gg; i_gg++)
                    if (s->xnu < s->E_gg[i_gg - 1]) break;
                double Egg_min = (s->E_gg[0] * s->E_gg[0]) / s->E_gg[1];
                if (s->xnu > Egg_min) {
                    double v = (s->ew * 6.25e8) / s->xnu;
                    if (i_gg <= s->n_gg) s->n_ph[(size_t)zi * s->n_gg + (i_gg - 1)] += v;
                    else if (s->jph + 1 <= s->nz)       /* n_ph(n_gg+1,j,k) aliases n_ph(1,j+1,k) in COMMON */
                        s->n_ph[(size_t)ZI(s->jph + 1, s->kph) * s->n_gg] += v;
                }
            }
            for (i_gg = 1; i_gg <= s->nphfield - 1; i_gg++)
                if (s->xnu < s->E_field[i_gg]) break;
            {
                double Egg_min = (s->E_field[0] * s->E_field[0]) / s->E_field[1];
                if (s->xnu > Egg_min)
                    s->n_field[(size_t)zi * s->nphfield + (i_gg - 1)] += 6.25e8 * s->ew / s->xnu;
            }
            if (s->ndxout >= s->cens_cap) { s->fatal = 1; goto L900; }
            {
                double *d = s->dbufout + 6 * s->ndxout;
                int *ib = s->ibufout + 6 * s->ndxout;
                d[0] = s->rpre; d[1] = s->zpre; d[2] = s->wmu; d[3] = s->phi; d[4] = s->ew; d[5] = s->xnu;
                ib[0] = s->jgpsp; ib[1] = s->jgplc; ib[2] = s->jgpmu; ib[3] = s->jph; ib[4] = s->kph;
                s->sidout[s->ndxout] = s->rng.sid;
                ib[5] = (int)(orc_fibran(s) * 1.e5);
                s->ndxout++;
            }
            goto L900;
        } else if (ikind == 3) {                             /* Compton scattering, l.588-689 */
            const uint64_t S_coll = s->rng.sid;
            nscat++;
            ph_save(s, &csv);
            if (scat_flag == 1) s->n_multiple++;
            s->cmcount1++;
            csv.ew = csv.ew / s->split2;
            for (ii = 1; ii <= s->split2; ii++) {
                const uint64_t S_child = orc_mix64(S_coll, (uint64_t)ii);
                if (philox) rng_stream(s, S_child);
                ph_restore(s, &csv);
                ewold = s->ew;
                orc_compb2d(s, &i_gam);
                if (s->ew > ewold * s->split2 * s->split1 * s->spl3_trg) {        /* 3rd split, l.639-669 */
                    ewold = csv.ew / s->split3;
                    for (ii2 = 1; ii2 <= s->split3; ii2++) {
                        /* Philox mode: every pass of the rejection loop owns the stream
                         * mix64(mix64(S_child, ii2), attempt), so attempts can be evaluated in any order */
                        const uint64_t S_sub = orc_mix64(S_child, (uint64_t)ii2);
                        uint64_t attempt = 0;
                        do {
                            if (philox) rng_stream(s, orc_mix64(S_sub, ++attempt));
                            ph_restore(s, &csv);
                            s->ew = csv.ew / s->split3;
                            orc_compb2d(s, &i_gam);
                        } while (s->ew <= ewold * s->split2 * s->split1 * s->spl3_trg);
                        s->edep[ZI(s->jph, s->kph)] += s->ew - ewold;
                        s->E_IC[i_gam - 1] += s->ew - ewold;
                        s->cmcount2++;
                        if (s->phi > 2.0 * PI) s->phi = s->phi - 2.0 * PI;
                        orc_imctrk2d(s, 1);
                        if (s->fatal) goto L_done;
                    }
                } else {
                    s->edep[ZI(s->jph, s->kph)] += s->ew - ewold;
                    s->E_IC[i_gam - 1] += s->ew - ewold;
                    s->cmcount2++;
                    if (s->phi > 2.0 * PI) s->phi = s->phi - 2.0 * PI;
                    orc_imctrk2d(s, 1);
                    if (s->fatal) goto L_done;
                }
            }
        }
    L900:
        if (scat_flag != -1) break;
    }
    /* recombine the unscattered trials, l.698-712 */
    if ((s->split1 - nscat) > 0 && scat_flag == -1) {
        ph_restore(s, &tsv);
        s->ew = (s->split1 - nscat) * tsv.ew;
        if (philox) rng_stream(s, orc_mix64(S_entry, TAG_F0));
        orc_imctrk2d(s, 0);
    }
L_done:
    free(comac_ar);
}

/* ------------------------------------------------------------------ sources */
#define F4(x) ((double)(x##f))

/* imcvol2d_para.f:90-414.  Philox mode: history (zone, ivol) owns stream
 * mix64(mix64('vol', zone), ivol), so results do not depend on how zones are farmed out. */
void orc_vol_calc(orc_state *s, int zone)
{
    int jv = (zone - 1) / s->nr + 1, kv = zone - s->nr * (jv - 1);       /* fp_mpi.f:333-343 get_j_k */
    int zi = ZI(jv, kv), i, ivol, more;
    double ew_save = s->ewsv[zi], f_thermal, delz, f_inn, f_outer, f_upper, f_lower, rnum, rnum0, x1, x2, psi;
    double rlo = (kv == 1) ? s->rmin : RR(kv - 1);
    const double *E_ph = s->E_ph;
    more = s->nsv[zi];
    f_thermal = s->Eloss_th[zi] / s->Eloss_tot[zi];
    delz = (jv == 1) ? ZZ(1) : ZZ(jv) - ZZ(jv - 1);
    f_inn = (4.4e1 / 7.0 * rlo * delz) / s->zsurf[zi];
    f_outer = (4.4e1 / 7.0 * RR(kv) * delz) / s->zsurf[zi];
    f_upper = (2.2e1 / 7.0 * (RR(kv) * RR(kv) - rlo * rlo)) / s->zsurf[zi];
    f_lower = (2.2e1 / 7.0 * (RR(kv) * RR(kv) - rlo * rlo)) / s->zsurf[zi];
    f_outer = f_inn + f_outer;
    f_upper = f_outer + f_upper;
    f_lower = f_upper + f_lower;
    (void)f_lower;
    const uint64_t S_zone = orc_mix64(0x766F6Cull, (uint64_t)zone);
    kn_initialize(s, s->seeds[zi]);                                      /* imcvol2d_para.f:115 initialize_rand(jv, kv) */

    for (ivol = 1; ivol <= more && !s->fatal; ivol++) {
        if (s->rng.mode == 3) rng_stream(s, orc_mix64(S_zone, (uint64_t)ivol));
        s->jph = jv; s->kph = kv;
        s->ew = ew_save;
        s->dcen = C_LIGHT * s->dt * orc_fibran(s);
        rnum = orc_fibran(s);
        if (rnum < f_thermal) {
            const double *eps = s->eps_th + (size_t)zi * s->n_vol;
            i = 0;
            rnum = orc_fibran(s);
            do { i = i + 1; } while (eps[i - 1] < rnum && i < s->n_vol);
            if (i < s->n_vol) s->xnu = E_ph[i - 1] + orc_fibran(s) * (E_ph[i] - E_ph[i - 1]);
            else s->xnu = E_ph[i - 1];
            rnum0 = orc_fibran(s);
            if (rnum0 < f_inn) {
                s->wmu = 2.0 * orc_fibran(s) - 1.0;
                CLAMP8(s->wmu);
                x1 = orc_fibran(s);
                x2 = orc_fibran(s);
                if (x1 < 0.5) {
                    s->phi = 1.1e1 / 7.0 + (1.1e1 / 7.0) * x2;
                    if (s->phi < 1.57079638e0) s->phi = 1.57079638e0;
                } else {
                    s->phi = -1.1e1 / 7.0 - (1.1e1 / 7.0) * x2;
                    if (s->phi > -1.57079638e0) s->phi = -1.57079638e0;
                }
                s->rpre = F4(1.00001) * rlo;
                if (jv == 1) s->zpre = ZZ(1) * orc_fibran(s);
                else s->zpre = ZZ(jv - 1) + orc_fibran(s) * (ZZ(jv) - ZZ(jv - 1));
            } else if (rnum0 < f_outer) {
                s->wmu = 2.0 * orc_fibran(s) - 1.0;
                CLAMP8(s->wmu);
                s->rpre = F4(0.999999) * RR(kv);
                if (jv == 1) s->zpre = ZZ(1) * orc_fibran(s);
                else s->zpre = ZZ(jv - 1) + orc_fibran(s) * (ZZ(jv) - ZZ(jv - 1));
                s->phi = -1.1e1 / 7.0 + 2.2e1 / 7.0 * orc_fibran(s);
                if (s->phi < -1.5707963e0) s->phi = -1.57079063e0;
                if (s->phi > 1.5707963e0) s->phi = 1.5707963e0;
            } else if (rnum0 < f_upper) {
                s->wmu = orc_fibran(s);
                if (s->wmu > 9.9999999e-1) s->wmu = 9.9999999e-1;
                if (s->wmu < 0.0) s->wmu = 0.0;
                s->phi = 4.4e1 / 7.0 * orc_fibran(s);
                if (s->phi > 2.0 * PI) s->phi = 2.0 * PI;
                psi = orc_fibran(s);
                s->zpre = F4(0.999999) * ZZ(jv);
                s->rpre = sqrt((rlo * rlo) + psi * (RR(kv) * RR(kv) - rlo * rlo));
            } else {
                s->wmu = -orc_fibran(s);
                s->phi = 4.4e1 / 7.0 * orc_fibran(s);
                if (s->wmu > 0.0) s->wmu = 0.0;
                if (s->wmu < -9.9999999e-1) s->wmu = -9.9999999e-1;
                if (s->phi > 2.0 * PI) s->phi = 2.0 * PI;
                double zl = (jv == 1) ? s->zmin : ZZ(jv - 1);
                s->zpre = F4(1.000001) * zl;
                if (s->zpre <= zl) s->zpre = zl + 1.e-6;
                psi = orc_fibran(s);
                s->rpre = sqrt((rlo * rlo) + psi * (RR(kv) * RR(kv) - rlo * rlo));
            }
        } else {
            const double *eps = s->eps_tot + (size_t)zi * s->n_vol;
            i = 0;
            rnum = orc_fibran(s);
            do { i = i + 1; } while (eps[i - 1] < rnum && i < s->n_vol);
            if (i < s->n_vol) s->xnu = E_ph[i - 1] + orc_fibran(s) * (E_ph[i] - E_ph[i - 1]);
            else s->xnu = E_ph[i - 1];
            s->wmu = 2.0 * orc_fibran(s) - 1.0;
            s->phi = 4.4e1 / 7.0 * orc_fibran(s);
            CLAMP8(s->wmu);
            if (s->phi > 2.0 * PI) s->phi = 2.0 * PI;
            if (jv == 1) s->zpre = ZZ(1) * orc_fibran(s);
            else s->zpre = ZZ(jv - 1) + orc_fibran(s) * (ZZ(jv) - ZZ(jv - 1));
            psi = orc_fibran(s);
            s->rpre = sqrt((rlo * rlo) + psi * (RR(kv) * RR(kv) - rlo * rlo));
        }
        /* bins, l.334-389: note xnu >= hutop maps to nphtotal here (not 0) */
        s->jgpsp = sp_group(s, s->xnu, F4(1.000001), F4(.999999), s->nphtotal);
        s->jgplc = lc_group(s, s->xnu);
        s->jgpmu = mu_group(s, s->wmu);
        orc_imctrk2d(s, -1);
    }
}

/* ------------------------------------------------------------------ surface sources */
/* planck2d.f:1-141: Planck-spectrum sampler + binning (REAL*4 literals 1.000001, .999999 at l.72-73) */
void orc_planck(orc_state *s, double tpl)
{
    double u4, ap0, ap1, ap2, ap3, rn1;
    do {
        u4 = orc_fibran(s);
        u4 = u4 * orc_fibran(s);
        u4 = u4 * orc_fibran(s);
        u4 = u4 * orc_fibran(s);
    } while (u4 <= 1.e-200);
    ap0 = -log(u4);
    ap1 = 1.0; ap2 = 1.0; ap3 = 1.0;
    rn1 = 1.08232e0 * orc_fibran(s);
    while (!(rn1 <= ap1)) {
        ap2 = ap2 + 1.0;
        ap3 = 1.0 / ap2;
        ap1 = ap1 + (ap3 * ap3) * (ap3 * ap3);       /* ap3**4: ifort expands to two squarings */
    }
    s->xnu = ap0 * ap3 * tpl;
    s->jgpsp = sp_group(s, s->xnu, F4(1.000001), F4(.999999), 0);
    s->jgplc = lc_group(s, s->xnu);
    s->jgpmu = mu_group(s, s->wmu);
}

/* imcsurf2d_para.f:694-788: energy from the piecewise power-law CDF of an external spectrum */
void orc_file_sample(orc_state *s, int spec)
{
    const double *E = s->E_file + (size_t)spec * ORC_NFMAX, *F = s->F_file + (size_t)spec * ORC_NFMAX;
    const double *P = s->P_file + (size_t)spec * ORC_NFMAX, *A = s->a1_file + (size_t)spec * ORC_NFMAX;
    const double *I = s->I_file + (size_t)spec * ORC_NFMAX;
    const int nfile = s->nfile[spec];
    int i;
    double x1 = orc_fibran(s), x2;
    for (i = 1; i <= nfile - 1; i++)
        if (P[i - 1] > x1) break;                    /* DO variable is nfile after a completed loop */
    x2 = orc_fibran(s);
    s->xnu = E[i - 1] * pow(A[i - 1] * I[i - 1] * x2 / (F[i - 1] * E[i - 1]) + 1.0, 1. / A[i - 1]);
    s->jgpsp = sp_group(s, s->xnu, 1.000001e0, .999999e0, 0);
    s->jgplc = lc_group(s, s->xnu);
    s->jgpmu = mu_group(s, s->wmu);
}

#define CLAMP10(v) do { if ((v) > 0.9999999999e0) (v) = 0.9999999999e0; if ((v) < -0.9999999999e0) (v) = -0.9999999999e0; } while (0)

/* imcsurf2d_para.f:228-343: inner (r = rmin) and outer (r = r(nr)) cylinder sources of z-ring js.
 * Philox mode: photon nxs of boundary kind b (0 inner, 1 outer, 2 upper, 3 lower) of ring q owns the stream
 * mix64(mix64('srf' + b, q), nxs). */
void orc_z_surf_calc(orc_state *s, int js)
{
    int nxs;
    double x1, x2;
    kn_initialize(s, s->zseeds[js - 1]);                                 /* imcsurf2d_para.f:242 initialize_zrand(js) */
    for (nxs = 1; nxs <= s->nsurfi[js - 1] && !s->fatal; nxs++) {
        if (s->rng.mode == 3) rng_stream(s, orc_mix64(orc_mix64(0x73726600ull, (uint64_t)js), (uint64_t)nxs));
        s->jph = js;
        s->wmu = 2.0 * orc_fibran(s) - 1.0;
        CLAMP10(s->wmu);
        s->phi = -1.1e1 / 7.0 + 2.2e1 / 7.0 * orc_fibran(s);
        if (s->phi < -1.5707963e0) s->phi = -1.57079063e0;
        if (s->phi > 1.5707963e0) s->phi = 1.5707963e0;
        s->rpre = s->rmin;
        if (js == 1) s->zpre = ZZ(1) * orc_fibran(s);
        else s->zpre = ZZ(js - 1) + orc_fibran(s) * (ZZ(js) - ZZ(js - 1));
        s->ew = s->ewsurfi[js - 1];
        s->dcen = orc_fibran(s) * C_LIGHT * s->dt;
        if (s->tbbi[js - 1] > 0.0) orc_planck(s, s->tbbi[js - 1]);
        else orc_file_sample(s, s->spec_i[js - 1]);
        s->kph = 1;
        orc_imctrk2d(s, -1);
    }
    for (nxs = 1; nxs <= s->nsurfo[js - 1] && !s->fatal; nxs++) {
        if (s->rng.mode == 3) rng_stream(s, orc_mix64(orc_mix64(0x73726601ull, (uint64_t)js), (uint64_t)nxs));
        s->jph = js;
        s->wmu = 2.0 * orc_fibran(s) - 1.0;
        CLAMP10(s->wmu);
        s->rpre = RR(s->nr);
        if (js == 1) s->zpre = ZZ(1) * orc_fibran(s);
        else s->zpre = ZZ(js - 1) + orc_fibran(s) * (ZZ(js) - ZZ(js - 1));
        x1 = orc_fibran(s);
        x2 = orc_fibran(s);
        if (x1 < 0.5) {
            s->phi = 1.1e1 / 7.0 + (1.1e1 / 7.0) * x2;
            if (s->phi < 1.57079638e0) s->phi = 1.57079638e0;
        } else {
            s->phi = -1.1e1 / 7.0 - (1.1e1 / 7.0) * x2;
            if (s->phi > -1.57079638e0) s->phi = -1.57079638e0;
        }
        s->ew = s->ewsurfo[js - 1];
        s->dcen = orc_fibran(s) * C_LIGHT * s->dt;
        if (s->tbbo[js - 1] > 0.0) orc_planck(s, s->tbbo[js - 1]);
        else orc_file_sample(s, s->spec_o[js - 1]);
        s->kph = s->nr;
        orc_imctrk2d(s, -1);
    }
}

/* imcsurf2d_para.f:353-534: upper (z = z(nz)) and lower (z = zmin, beamed: wmu = 0.99999999) disk sources of r-ring ks */
void orc_r_surf_calc(orc_state *s, int ks)
{
    int nxs;
    double psi;
    const double rlo = (ks == 1) ? s->rmin : RR(ks - 1), rhi = RR(ks);
    kn_initialize(s, s->rseeds[ks - 1]);                                 /* imcsurf2d_para.f:372 initialize_rrand(ks) */
    for (nxs = 1; nxs <= s->nsurfu[ks - 1] && !s->fatal; nxs++) {
        if (s->rng.mode == 3) rng_stream(s, orc_mix64(orc_mix64(0x73726602ull, (uint64_t)ks), (uint64_t)nxs));
        s->kph = ks;
        s->wmu = -orc_fibran(s);
        CLAMP10(s->wmu);
        s->phi = 2.0 * PI * orc_fibran(s);
        psi = orc_fibran(s);
        s->zpre = ZZ(s->nz);
        s->rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        s->ew = s->ewsurfu[ks - 1];
        s->dcen = orc_fibran(s) * C_LIGHT * s->dt;
        if (s->tbbu[ks - 1] > 0.0) orc_planck(s, s->tbbu[ks - 1]);
        else orc_file_sample(s, s->spec_u[ks - 1]);
        s->jph = s->nz;
        orc_imctrk2d(s, -1);
    }
    for (nxs = 1; nxs <= s->nsurfl[ks - 1] && !s->fatal; nxs++) {
        if (s->rng.mode == 3) rng_stream(s, orc_mix64(orc_mix64(0x73726603ull, (uint64_t)ks), (uint64_t)nxs));
        s->kph = ks;
        s->wmu = 9.9999999e-1;                        /* l.452-460: star or not, the lower source is beamed upwards */
        s->phi = 2.0 * PI * orc_fibran(s);
        psi = orc_fibran(s);
        s->zpre = s->zmin;
        s->rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        s->ew = s->ewsurfl[ks - 1];
        if (s->tbbl[ks - 1] > 0.e-20) orc_planck(s, s->tbbl[ks - 1]);
        else orc_file_sample(s, s->spec_l[ks - 1]);
        s->dcen = orc_fibran(s) * C_LIGHT * s->dt;      /* drawn after the energy here (l.484) */
        s->jph = 1;
        orc_imctrk2d(s, -1);
    }
}

/* imcsurf2d_para.f:544-685 file_sp: external spectrum (rows E, L_disk, F_blr, F_ir) -> normalised BLR + torus flux
 * (Ghisellini & Madau 1996) and its piecewise power-law CDF.  Returns nfile (< 0: error).  The disk part
 * (Id_file, F_como) is computed by the reference but not used (l.648-649: F_file = F_blr + F_ir). */
int orc_file_sp(int n_rows, const double *E_in, const double *L_disk, const double *F_blr_in, const double *F_ir_in,
                double g_bulk, double R_blr, double fr_blr, double R_ir, double fr_ir, double R_disk, double d_jet,
                double *E_file, double *F_file, double *P_file, double *a1, double *I_file, double *int_file)
{
    double Ld[ORC_NFMAX], Fb[ORC_NFMAX], Fi[ORC_NFMAX];
    int i, nfile;
    double Ltot_disk = 0., Ftot_blr = 0., Ftot_ir = 0., Isum = 0.;
    (void)R_disk; (void)d_jet;
    for (i = 0; i < ORC_NFMAX; i++) { E_file[i] = 0; Ld[i] = 0; Fb[i] = 0; Fi[i] = 0; F_file[i] = 0; P_file[i] = 0; a1[i] = 0; I_file[i] = 0; }
    /* the read loop stops after row nfmax-1 or at the first E <= 0 (l.567-594); end of file is an err= exit */
    i = 0;
    for (;;) {
        i = i + 1;
        if (i > n_rows) break;                         /* err=120 */
        E_file[i - 1] = E_in[i - 1]; Ld[i - 1] = L_disk[i - 1]; Fb[i - 1] = F_blr_in[i - 1]; Fi[i - 1] = F_ir_in[i - 1];
        if (!(E_file[i - 1] > 0. && i <= ORC_NFMAX - 1)) break;
    }
    nfile = i - 1;
    for (i = 1; i <= ORC_NFMAX - 1; i++) {
        Ltot_disk = Ltot_disk + Ld[i - 1] * (E_file[i] - E_file[i - 1]);
        Ftot_blr = Ftot_blr + Fb[i - 1] * (E_file[i] - E_file[i - 1]);
        Ftot_ir = Ftot_ir + Fi[i - 1] * (E_file[i] - E_file[i - 1]);
    }
    Ltot_disk = Ltot_disk / sqrt(E_file[1] / E_file[0]);
    Ftot_blr = Ftot_blr / sqrt(E_file[1] / E_file[0]);
    Ftot_ir = Ftot_ir / sqrt(E_file[1] / E_file[0]);
    {
        const double Ftot_blr_norm = 17. / 48. / PI * (g_bulk * g_bulk) * fr_blr * Ltot_disk / (R_blr * R_blr);
        const double Ftot_ir_norm = 1. / 4. / PI * (g_bulk * g_bulk) * fr_ir * Ltot_disk / (R_ir * R_ir);
        for (i = 0; i < ORC_NFMAX; i++) {
            Fb[i] = Fb[i] / Ftot_blr * Ftot_blr_norm;
            Fi[i] = Fi[i] / Ftot_ir * Ftot_ir_norm;
            F_file[i] = Fb[i] + Fi[i];
        }
    }
    if (nfile < 2) return -1;
    for (i = 1; i <= nfile - 1; i++) {
        const double alpha = log(F_file[i] / F_file[i - 1]) / log(E_file[i] / E_file[i - 1]);
        a1[i - 1] = alpha + 1.;
        if (a1[i - 1] > 2.e1) a1[i - 1] = 2.e1;
        if (a1[i - 1] < -2.e1) a1[i - 1] = -2.e1;
        if (fabs(a1[i - 1]) < 1.e-3)
            I_file[i - 1] = F_file[i - 1] * E_file[i - 1] * log(E_file[i] / E_file[i - 1]);
        else
            I_file[i - 1] = F_file[i - 1] * E_file[i - 1] * (pow(E_file[i] / E_file[i - 1], a1[i - 1]) - 1.) / a1[i - 1];
        Isum = Isum + I_file[i - 1];
        P_file[i - 1] = Isum;
    }
    for (i = 1; i <= nfile - 1; i++) P_file[i - 1] = P_file[i - 1] / Isum;
    *int_file = Isum;
    return nfile;
}

/* imcfield2d.f:57-148.  Philox mode: a census photon's history stream is the id of
 * the flight stream that wrote it (unique per record), used under the new cycle's key. */
void orc_field_calc(orc_state *s)
{
    int64_t nread = s->ndxout, n;
    s->ndxout = 0;
    if (nread == 0) return;
    if (nread >= s->cens_cap) { s->fatal = 2; return; }
    memcpy(s->dbufin, s->dbufout, sizeof(double) * 6 * (size_t)nread);
    memcpy(s->ibufin, s->ibufout, sizeof(int) * 6 * (size_t)nread);
    memcpy(s->sidin, s->sidout, sizeof(uint64_t) * (size_t)nread);
    s->ndxin = nread;
    for (n = 0; n < nread && !s->fatal; n++) {
        const double *d = s->dbufin + 6 * n;
        const int *ib = s->ibufin + 6 * n;
        s->rpre = d[0]; s->zpre = d[1]; s->wmu = d[2]; s->phi = d[3]; s->ew = d[4]; s->xnu = d[5];
        s->jgpsp = ib[0]; s->jgplc = ib[1]; s->jgpmu = ib[2]; s->jph = ib[3]; s->kph = ib[4];
        if (s->rng.mode == 3) rng_stream(s, s->sidin[n]);
        if (s->rng.mode == 1) {                                          /* imcfield2d.f:115-116: the photon carries its seed */
            s->seeds[ZI(s->jph, s->kph)] = ib[5];
            kn_initialize(s, ib[5]);
        }
        s->dcen = C_LIGHT * s->dt;
        if (s->wmu > 1.0) s->wmu = 1.0;
        if (s->wmu < -1.0) s->wmu = -1.0;
        orc_imctrk2d(s, -1);
    }
}

/* imcgen2d.f:559-601 (worker branch): tallies zeroed at the top of every step.
 * ecens / n_field / Ed_* are zeroed only after FP has consumed them (imcgen2d.f:534-546);
 * fout is cumulative on the workers. */
void orc_step_zero_worker(orc_state *s)
{
    memset(s->edep, 0, sizeof(double) * s->nzone);
    memset(s->prdep, 0, sizeof(double) * s->nzone);
    memset(s->npcen, 0, sizeof(int64_t) * s->nzone);
    memset(s->ecens, 0, sizeof(double) * s->nzone);
    memset(s->n_field, 0, sizeof(double) * (size_t)s->nzone * s->nphfield);
    memset(s->n_ph, 0, sizeof(double) * (size_t)s->nzone * s->n_gg);
    memset(s->E_IC, 0, sizeof(double) * s->num_nt);
    memset(s->edout, 0, sizeof(double) * ORC_NMUMAX * ORC_NLCMAX);
    memset(s->erlki, 0, sizeof(double) * s->nz);
    memset(s->erlko, 0, sizeof(double) * s->nz);
    memset(s->erlku, 0, sizeof(double) * s->nr);
    memset(s->erlkl, 0, sizeof(double) * s->nr);
    memset(s->Ed_in, 0, sizeof(double) * s->nr);
    memset(s->Ed_ref, 0, sizeof(double) * s->nr);
    s->incounter = s->outcounter = s->cmcount1 = s->cmcount2 = s->lkcount = 0;
}


/* ------------------------------------------------------------------ event-list reduction */
/* The per-photon transform and binning shared by the reference's two post-processors, postprocessing/plcm.c:384-450
 * (light curves) and postprocessing/pspt.c:249-289 (time-resolved SEDs): Doppler boost into the observer frame (bulk
 * Lorentz factor gam_bulk along +z), light-travel-time correction to the plane through r = rmax, then binning in arrival
 * time, observer-frame direction cosine and photon energy.  The two programs differ only in the bin rules:
 *   plcm: time bins t_lo[k] = k dt by repeated addition (plcm.c:236-240), `time -= t_offset` first; mu bins half open
 *         [mu0, mu1), first match; EVERY band [E0, E1) that holds E is incremented (bands may overlap, l.441-449);
 *   pspt: time bins t_lo[n] = t0 + n dt, t_hi[n] = t0 + n dt + dt (pspt.c:145-148); ONE closed mu window (l.274:
 *         `mu < mu0 || mu > mu1` skips); only the FIRST band that holds E (l.278-280).
 * events: 7 doubles per record (t_bound, xnu, ew, rpre, zpre, wmu, phi), imcleak2d.f:171,181.
 * Accumulators [k][n][l] (time, mu, band), added to (not zeroed). */
void orc_events_reduce(int64_t n_ev, const double *ev, double gam_bulk, double rmax, double t_offset, int n_t,
                       const double *t_lo, const double *t_hi, int n_mu, const double *mu0, const double *mu1, int mu_closed,
                       int n_bins, const double *E0, const double *E1, int first_band_only,
                       double *sum_ew, double *sum_ew2, int64_t *count)
{
    for (int64_t q = 0; q < n_ev; q++) {
        double t_bound = ev[7 * q], E = ev[7 * q + 1], ew = ev[7 * q + 2], r = ev[7 * q + 3], z = ev[7 * q + 4];
        double mu = ev[7 * q + 5], phi = ev[7 * q + 6], betta, doppler, cdt, time;
        int k, n, l;
        mu = -mu;
        betta = sqrt(1. - 1. / (gam_bulk * gam_bulk));
        doppler = gam_bulk * (1. + mu * betta);
        t_bound = (t_bound - betta * z * 3.33333333e-11) / doppler;
        E = E * doppler;
        ew = ew * doppler;
        mu = (mu + betta) / (1. + mu * betta);
        cdt = z * mu / gam_bulk + sqrt(1. - mu * mu) * (rmax - r * cos(phi));
        time = t_bound + 3.33333333e-11 * cdt;
        time -= t_offset;
        for (k = 0; k < n_t; k++)
            if (time >= t_lo[k] && time < t_hi[k]) break;
        for (n = 0; n < n_mu; n++)
            if (mu >= mu0[n] && (mu_closed ? mu <= mu1[n] : mu < mu1[n])) break;
        if (!(k < n_t && n < n_mu)) continue;
        for (l = 0; l < n_bins; l++)
            if (E >= E0[l] && E < E1[l]) {
                const size_t at = ((size_t)k * n_mu + n) * n_bins + l;
                sum_ew[at] += ew;
                sum_ew2[at] += (ew * ew);
                count[at]++;
                if (first_band_only) break;
            }
    }
}
