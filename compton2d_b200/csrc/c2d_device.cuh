// c2d_device.cuh -- device-side physics of the Compton2d hot path for sm_100a.
//
// One thread owns one photon history; every random number is a pure function of
// (rseed, ncycle, stream id, draw index) through Philox-4x32-10, where the stream id encodes the
// history's ancestry (source zone/index -> trial -> scattered child ...), so results do not depend
// on thread scheduling, launch geometry or GPU count.
//
// Reference parity (file:line relative to /root/reference/src_20121026/):
//   fly<MODE>()      imctrk2d.f:158-586   one flight from label 110 to census / leak / collision
//   sigma_kn_bins()  comtot2d.f:219-247   angle-averaged K-N cross section per electron bin
//   dilog_neg()      comtot2d.f:356-433   CERNLIB C332 dilogarithm (argument -x, x > 0)
//   compb2d()        compb_2d.f:1-318     one Compton scattering
//   nth2d()          nontherm2d.f:159-183 electron draw from Pnt
//   imcleak()        imcleak2d.f:2-320    boundary tallies + escape record; specular mirror at the lower surface (cr_sent = 1 / 4)
//   source_volume()  imcvol2d_para.f:157-389
// The arithmetic follows the reference expression by expression (REAL*4 literals included) so that
// integer outcomes replay exactly against the CPU oracle; compile with -fmad=false.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c2d {

// x**2 as the Fortran compiler evaluates it: one multiplication (exactly rounded, unlike the device pow)
__host__ __device__ __forceinline__ double sq(double x) { return x * x; }


constexpr double kPi = 3.1415926536e0;        // general.pa:25
constexpr double kRadCp = 3.333564097e-11;    // general.pa:24
constexpr double kCLight = 2.9979245620e10;   // general.pa:26
#ifndef C2D_MAXSEG
#define C2D_MAXSEG 12
#endif
constexpr int kMaxSeg = C2D_MAXSEG;           // zone crossings held by the fast path of one history
constexpr int kNmuMax = 32;
constexpr int kNlcMax = 16;

constexpr uint64_t kTagTrials = 0xA1ull;
constexpr uint64_t kTagF0 = 0xF0ull;
constexpr uint64_t kTagVol = 0x766F6Cull;
constexpr uint64_t kTagSurf = 0x73726600ull;     // + boundary kind (0 inner, 1 outer, 2 upper, 3 lower)

enum { MODE_PRE = 2, MODE_TRIAL = -1, MODE_TRIAL_SLOW = -2, MODE_F0 = 0, MODE_CHILD = 1 };
enum { END_KILLED = 0, END_CENSUS = 1, END_LEAK = 2, END_DROPPED = 3, END_COLLISION = 4, END_LOWEW = 5,
       END_OVERFLOW = 6, END_FATAL = 7 };
enum { CNT_NDXOUT = 0, CNT_CMCOUNT1, CNT_CMCOUNT2, CNT_LKCOUNT, CNT_NEVENTS, CNT_HISTORIES, CNT_TRIALS,
       CNT_MULTIPLE, CNT_KILLED, CNT_TEAMS, CNT_LITERAL, CNT_QPEAK, CNT_FATAL, CNT_SEGMENTS, CNT_S3PASSES, CNT_S3HARD, CNT_N };   // CNT_N == 16

struct Photon {
    double xnu, wmu, phi, rpre, zpre, dcen, ew;
    int jph, kph, jgpsp, jgplc, jgpmu;
};

struct CollTask {            // a Compton collision waiting for its split2 children (imctrk2d.f:588-689)
    Photon p;
    int pad;
    uint64_t sid;            // stream of the flight that collided
};

struct FlightTask {          // a scattered photon (after compb2d) waiting for its scat_flag = 1 flight
    Photon p;
    int i_gam, zi;
    uint32_t d, pad;         // draw cursor of the stream `sid`
    uint64_t sid;
    double ewold;
};

// an item a block team hands to the grid team: where the team stopped (next attempt / trip) and which item
struct GridItem { unsigned long long first; unsigned idx, pad; };
struct S3Map {               // the all-gathered 3rd-split tasks and this rank's share of their sub-children (split3_sub)
    const CollTask *base;
    long long stride;        // tasks per rank slot
    long long prefix[9];     // prefix[r] = tasks of ranks < r (world <= 8)
    int world, rank;
};

struct Bank {                // census bank, SoA (commonblock.f:8,72 dbuf/ibuf record, imctrk2d.f:566-580)
    double *rpre, *zpre, *wmu, *phi, *ew, *xnu;
    uint32_t *bins;          // jgpsp | jgplc << 16 | jgpmu << 24
    uint32_t *zone;          // jph | kph << 16
    int32_t *seed;
    uint64_t *sid;
};

struct Seg {                 // one zone crossing of the collision-free ray of a history
    int zi;
    double sigsc, sigabs, L, trldb, dcen, ew, dele, rpre;
    double att;              // exp(-sigabs L) (0 when sigabs L >= 100): the factor every weight on this ray is attenuated by
};

struct AbsTask {             // a history whose ray has sigabs L > 1e-5 in some crossing: its per-trial absorption-point draws
    Photon p;                // (imctrk2d.f:451-459) are summed by k_absdraw, one warp per task
    int nlit;                // literal trials (flown one by one, they make their own draws): indices at lit_pool[lit_off ...]
    unsigned long long lit_off;
    uint64_t S_trials;
    double wtmin;
};

struct Params {
    int nz, nr, nzone, n_vol, num_nt, nphfield, n_gg, nphtotal, nph_lc, nmu;
    int pair_switch, spec_switch, cr_sent, split1, split2, split3, spl3_trg, ncycle, sigma_mode;
    int bins_off;                                     // 1: sources are born with jgpsp = jgplc = jgpmu = 0 (c2d_run.quirk_cycle0)
    int n_ref;                                        // Compton-reflection tables of ref_matrix.f (static inputs), 0 = none
    const double *E_ref, *P_ref, *W_abs;              // P_ref / W_abs in COMMON order: (n_out, n_in) at (n_in-1) n_ref + (n_out-1)
    uint32_t key0, key1;
    double rmin, zmin, dt, time;
    // static tables
    const double *z, *r, *E_ph, *E_field, *E_gg, *gnt, *gam0, *opb, *dgam, *kn_pref, *kn_a1, *kn_a2, *kn_cb, *hu, *Elcmin, *Elcmax, *mu;
    // per-step tables
    const double *kappa_tot, *eps_tot, *eps_th, *f_nt, *Pnt, *n_e, *f_pair, *k_gg, *tbbl;
    const double *ewsv, *Eloss_th, *Eloss_tot, *zsurf;
    const int64_t *nsv_prefix;         // [nzone+1]
    // surface sources (c2d_surf_begin): slots in the order (inner, outer) of z ring 1..nz, then (upper, lower) of r ring 1..nr
    const int64_t *surf_prefix;        // [2 nz + 2 nr + 1]
    const double *ewsurf, *tbb_surf;   // [2 nz + 2 nr]
    const int *spec_surf;              // [2 nz + 2 nr] external-spectrum table of the slot (used when tbb <= 0)
    const double *sp_E, *sp_F, *sp_P, *sp_a1, *sp_I;   // [n_spec][sp_nfmax] file_sp tables (imcsurf2d_para.f:544-685)
    const int *sp_n;                   // [n_spec] nfile
    int sp_nfmax, star_switch;
    // sigma_mode 1 (c2d_sigma.cu): per-step piecewise-Chebyshev representation of comtot(zone, E)
    const double *sig_coef;            // [nzone][sig_nint][8] coefficients of the exact-branch sum on each interval
    const double *sig_AB;              // [nzone][num_nt][2] prefix sums of the Thomson-branch bins
    const double *sig_W;               // [nzone][num_nt] f_nt * dgamma
    const double *sig_lo, *sig_iw;     // [sig_nint] interval lower bound in ln(x) and 2 / width
    const double *sig_NB;              // [nzone][num_nt] suffix sums bounding the rounding-noise terms of bins >= i (comtot_table_zone)
    const double *sig_CW;              // [nzone][num_nt] kn_cb[i] * sig_W[z][i]
    const int *sig_if;                 // [nzone] first bin with f_nt > 0
    int sig_nint, sig_nsub; double sig_t0, sig_inv_h;
    // tallies
    double *edep, *prdep, *ecens, *n_field, *n_ph, *E_IC, *fout, *edout, *erlki, *erlko, *erlku, *erlkl,
        *Ed_in, *Ed_ref;
    unsigned long long *npcen, *nelectron, *counters;
    // banks
    Bank in, out;
    unsigned long long census_cap;
    double *events;                    // 7 doubles per record
    unsigned long long event_cap;
    CollTask *q_out;                   // collision tasks produced by this launch
    unsigned long long *q_out_count;
    CollTask *q3;                      // children that triggered the 3rd split (imctrk2d.f:639)
    unsigned long long *q3_count;
    unsigned long long q_cap;
    AbsTask *qabs;                     // histories waiting for k_absdraw
    unsigned long long *qabs_count, qabs_cap;
    const uint32_t *nt_guide;   // [zone][kNtGuide] lo | hi << 16: bracket of nth2d's bisection per 1/kNtGuide of rnum, or null
    int kn_cap;              // trips of the K-N rejection loop a thread makes before it hands its child to a block team (C2D_KN_CAP)
    int *lit_pool;                     // trial indices of the literal trials of queued histories
    unsigned long long *lit_count, lit_cap;
};

// software pipelining depth of the sigma-table noise loops (same arithmetic in the same order; measured)
#ifndef C2D_NOISE_UNROLL
#define C2D_NOISE_UNROLL 2
#endif
constexpr int kNoiseUnroll = C2D_NOISE_UNROLL;
constexpr int kSortE = 24;             // energy classes of the bank's sort key
// energy class of a census photon: one per octave of photon energy from 1 keV down -- the Thomson-branch count m(x) of
// comtot, and with it the length of the sigma table's noise loop, is a function of log x (7 electron bins per octave)
__device__ __forceinline__ int sort_class(double xnu)
{
    int e2;
    (void)frexp(xnu, &e2);
    int c = 1 - e2;
    if (c < 0) c = 0;
    if (c > kSortE - 1) c = kSortE - 1;
    return c;
}

// ------------------------------------------------------------------------------------------ warp
// The kernels are written against a warp of kWarp lanes.  On the GPU kWarp = 32; the g++ debug build of
// tests/emu (TEST INFRASTRUCTURE) runs the same sources with one-lane warps.
#ifdef C2D_EMU
constexpr int kWarp = 1;
__device__ __forceinline__ int lane_id() { return 0; }
__device__ __forceinline__ unsigned warp_ballot(bool pr) { return pr ? 1u : 0u; }
template <class T> __device__ __forceinline__ T warp_shfl(T v, int) { return v; }
template <class T> __device__ __forceinline__ T warp_shfl_xor(T v, int) { return v; }
__device__ __forceinline__ int first_lane(unsigned m) { return __builtin_ffs((int)m) - 1; }
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) { return v; }
#else
constexpr int kWarp = 32;
__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ unsigned warp_ballot(bool pr) { return __ballot_sync(0xffffffffu, pr); }
template <class T> __device__ __forceinline__ T warp_shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class T> __device__ __forceinline__ T warp_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ int first_lane(unsigned m) { return __ffs((int)m) - 1; }
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = kWarp / 2; o > 0; o >>= 1) v += warp_shfl_xor(v, o);
    return v;
}

// ------------------------------------------------------------------------------------------ RNG
__device__ __forceinline__ uint64_t mix64(uint64_t a, uint64_t b)
{
    uint64_t x = a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ void philox4x32_10_inl(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1, uint32_t (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0; c1 = (uint32_t)p1; c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
// The history kernels call Philox from ~100 places; inlined, those copies are 11 % of a 91 000-instruction kernel whose
// warps stall on instruction fetch (no_instructions 23 % of the stall samples).  One out-of-line copy, result in registers.
#if defined(C2D_EMU) || defined(C2D_PHILOX_INLINE)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4])
{
    philox4x32_10_inl(c0, c1, c2, c3, k0, k1, o);
}
#else
static __device__ __noinline__ uint4 philox4x32_10_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    philox4x32_10_inl(c0, c1, c2, c3, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4])
{
    const uint4 v = philox4x32_10_call(c0, c1, c2, c3, k0, k1);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
#endif

__device__ __forceinline__ double u52(uint32_t lo, uint32_t hi)
{   // 52-bit uniform in (0,1): ((w >> 12) + 0.5) * 2^-52, exactly representable
    const uint64_t w = ((uint64_t)hi << 32) | lo;
    return ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

// draw d of stream sid = word pair (d & 1) of philox(ctr = (sid_lo, sid_hi, d >> 1, 0))
__device__ __forceinline__ double rng_uniform(const Params &p, uint64_t sid, uint32_t d)
{
    uint32_t o[4];
    philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), d >> 1, 0u, p.key0, p.key1, o);
    return (d & 1u) ? u52(o[2], o[3]) : u52(o[0], o[1]);
}

// A stream cursor.  One Philox call yields draws 2m and 2m+1; the odd one is kept for the next call.
struct Rng {
    uint64_t sid;
    uint32_t d;
    uint32_t have;          // index of the cached (odd) draw, 0 = none (draw 0 is never odd)
    double cached;
    __device__ __forceinline__ void stream(uint64_t s) { sid = s; d = 0; have = 0; cached = 0.0; }
    __device__ __forceinline__ double next(const Params &p)
    {
        const uint32_t i = d++;
        if (i == have && i != 0u) return cached;
        uint32_t o[4];
        philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), i >> 1, 0u, p.key0, p.key1, o);
        if (i & 1u) return u52(o[2], o[3]);
        cached = u52(o[2], o[3]); have = i + 1u;
        return u52(o[0], o[1]);
    }
};

// The xi of the trial flights of a history with trial stream S: u_t = (K_t + 0.5) 2^-52, K_t = coarse8 << 44 | fine44.
//   Which trials have coarse = 0xFF ("candidates", 1 trial in 256, independently) is decided by the GAPS between them:
//       gap_j = floor(ln U_j * (1 / ln(255/256))),  U_j = word pair (j & 1) of philox(S, j >> 1, 3),
//       candidate_j = candidate_{j-1} + 1 + gap_j  -- a Bernoulli(1/256) process has geometric gaps;
//   every other trial has coarse = (w * 255) >> 32 in [0, 254], w = word (t & 3) of philox(S, t >> 2, 0);
//   fine = word 0 and the low 12 bits of word 1 of philox(S, t, 1).
// A trial collides only if u exceeds a threshold close to 1, so an optically thin ray (tau < 1/256) only has to look at
// its ~split1/256 candidates: history() jumps from candidate to candidate instead of drawing split1 numbers.
constexpr double kInvLnQ = -255.49967384180653;      // 1 / ln(255/256)
__device__ __forceinline__ uint64_t trial_fine44(const Params &p, uint64_t S, uint32_t t)
{
    uint32_t o[4];
    philox4x32_10((uint32_t)S, (uint32_t)(S >> 32), t, 1u, p.key0, p.key1, o);
    return (((uint64_t)(o[1] & 0xFFFu)) << 32) | o[0];
}
// floor(ln U / ln(255/256)) for U in (0, 1).  The double-precision log is 7 % of k_history's instructions; the hardware's
// single-precision logarithm gives the product to 2.3e-3 absolute ((float)U: 6e-8 in ln U; __logf: 3 ulp of |ln U| <= 37, i.e.
// 6.5e-6; times 255.5; the product's own rounding 5.5e-4 at its maximum 9 400), so whenever its fractional part is at least 5e-3
// away from an integer the floor is already decided.  The other 1 % take the double-precision path: same integer either way
// (c2d_selftest_gap counts disagreements over 1e9 uniforms: 0).
__device__ __forceinline__ double trial_gap(double U)
{
#ifndef C2D_EMU
    const float tf = __logf((float)U) * (float)kInvLnQ;
    const float fl = floorf(tf), fr = tf - fl;
    if (fr > 5.e-3f && fr < 1.f - 5.e-3f) return (double)fl;
#endif
    return floor(log(U) * kInvLnQ);
}
struct TrialGen {
    uint64_t S;
    long long next_c;        // position of the next candidate (>= split1: none left)
    uint32_t j;
    __device__ __forceinline__ void advance(const Params &p)
    {
        uint32_t o[4];
        philox4x32_10((uint32_t)S, (uint32_t)(S >> 32), j >> 1, 3u, p.key0, p.key1, o);
        const double U = (j & 1u) ? u52(o[2], o[3]) : u52(o[0], o[1]);
        const double gap = trial_gap(U);
        j++;
        next_c = (gap > 1.e9) ? (1ll << 40) : next_c + 1 + (long long)gap;
    }
    __device__ __forceinline__ void init(const Params &p, uint64_t S_) { S = S_; j = 0; next_c = -1; advance(p); }
    // coarse byte of trial t; t must not decrease between calls
    __device__ __forceinline__ uint32_t coarse(const Params &p, uint32_t t)
    {
        while (next_c < (long long)t) advance(p);
        if (next_c == (long long)t) return 0xFFu;
        uint32_t o[4];
        philox4x32_10((uint32_t)S, (uint32_t)(S >> 32), t >> 2, 0u, p.key0, p.key1, o);
        const uint32_t w = (t & 2u) ? ((t & 1u) ? o[3] : o[2]) : ((t & 1u) ? o[1] : o[0]);
        return (uint32_t)(((uint64_t)w * 255u) >> 32);
    }
    __device__ __forceinline__ double uniform(const Params &p, uint32_t t)
    {
        const uint64_t K = ((uint64_t)coarse(p, t) << 44) | trial_fine44(p, S, t);
        return ((double)K + 0.5) * (1.0 / 4503599627370496.0);
    }
};

// The uniform of the k-th absorption-point draw (imctrk2d.f:451-459) of trial flight t: word pair (t & 1) of
// philox(S, t >> 1, 2 + k) -- two trials share a call.
__device__ __forceinline__ double abs_uniform(const Params &p, uint64_t S, uint32_t t, uint32_t k)
{
    uint32_t o[4];
    philox4x32_10((uint32_t)S, (uint32_t)(S >> 32), t >> 1, 2u + k, p.key0, p.key1, o);
    return (t & 1u) ? u52(o[2], o[3]) : u52(o[0], o[1]);
}

// read-only global load (tables that no kernel of the same launch writes)
#ifdef C2D_EMU
__device__ __forceinline__ double c2d_ldg(const double *a) { return *a; }
__device__ __forceinline__ uint32_t c2d_ldg(const uint32_t *a) { return *a; }
#else
__device__ __forceinline__ double c2d_ldg(const double *a) { return __ldg(a); }
__device__ __forceinline__ uint32_t c2d_ldg(const uint32_t *a) { return __ldg(a); }
#endif

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ void atomic_add_f64(double *a, double v) { atomicAdd(a, v); }
__device__ __forceinline__ void count(const Params &p, int which, unsigned long long n = 1ull)
{
    atomicAdd(p.counters + which, n);
}

#define C2D_CLAMP8(v) do { if ((v) > 9.9999999e-1) (v) = 9.9999999e-1; if ((v) < -9.9999999e-1) (v) = -9.9999999e-1; } while (0)

// first index i in [1, n-1] with x < E[i] (0-based E), else n  == the linear scans of imctrk2d.f:387-389,555-557
__device__ __forceinline__ int scan_lt(const double *E, int n, double x)
{
    int lo = 1, hi = n;               // answer in [1, n]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x < E[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// spectrum group: compb_2d.f:249-274 / imcleak2d.f:345-370 / imcvol2d_para.f:334-359
__device__ __forceinline__ int sp_group(const Params &p, double x, double hubot_f, double hutop_f, int top_value)
{
    int jbot = 1, jtop = p.nphtotal + 1, jmid;
    const double hubot = hubot_f * p.hu[0], hutop = hutop_f * p.hu[jtop - 1];
    if (x >= hutop) return top_value;
    if (x <= hubot) return 0;
    for (;;) {
        jmid = (jbot + jtop) / 2;
        if (jmid == jbot) break;
        const double h = p.hu[jmid - 1];
        if (x == h) break;
        if (x < h) jtop = jmid; else jbot = jmid;
    }
    return jmid;
}
__device__ __forceinline__ int lc_group(const Params &p, double x)
{
    for (int m = 1; m <= p.nph_lc; m++)
        if (x > p.Elcmin[m - 1] && x <= p.Elcmax[m - 1]) return m;
    return 0;
}
__device__ __forceinline__ int mu_group(const Params &p, double w)
{
    for (int n = 1; n <= p.nmu; n++)
        if (w <= p.mu[n - 1]) return n;
    return p.nmu;
}

// ------------------------------------------------------------------------------------------ comtot
// Chebyshev coefficients of CERNLIB C332 (comtot2d.f:371-378)
static __device__ __constant__ double kDilogC[20] = {0.42996693560813697e0, 0.40975987533077105e0,
    -0.01858843665014592e0, 0.00145751084062268e0, -0.00014304184442340e0,
    0.1588415541880e-4, -0.190784959387e-5, 0.024195180854e-5,
    -0.003193341274e-5, 0.000434545063e-5, -0.000060578480e-5,
    0.000008612098e-5, -0.000001244332e-5, 0.000000182256e-5,
    -0.000000027007e-5, 0.000000004042e-5, -0.000000000610e-5,
    0.000000000093e-5, -0.000000000014e-5, 0.000000000002e-5};

// intg_v(xa) - intg_v(xb) of comtot2d.f:238, with both dilogarithms (comtot2d.f:356-433, argument -x, x > 0:
// only the branches T <= 1 and T > 1 are reachable) evaluated side by side: two independent Clenshaw
// recurrences keep the FP64 pipe busy.  The recurrence uses explicit FMAs (the only place the tracker departs
// from the reference's operation-by-operation arithmetic; it moves sigma by <= 1 ulp).
template <bool SMOOTH>
__device__ __forceinline__ double intg_v_diff_t(double xa, double xb)
{
    const double PI2 = kPi * kPi, PI6 = PI2 / 6, PI12 = PI2 / 12;
    double Ya, Sa, Aa, Yb, Sb, Ab;
    if (xa <= 1) { Ya = xa; Sa = 1; Aa = 0; } else { Ya = 1 / xa; Sa = -1; const double l = log(xa); Aa = PI6 + 0.5 * l * l; }
    if (xb <= 1) { Yb = xb; Sb = 1; Ab = 0; } else { Yb = 1 / xb; Sb = -1; const double l = log(xb); Ab = PI6 + 0.5 * l * l; }
    const double Ha = Ya + Ya - 1, ALa = Ha + Ha, Hb = Yb + Yb - 1, ALb = Hb + Hb;
    double B0a = 0, B1a = 0, B2a = 0, B0b = 0, B1b = 0, B2b = 0;
#pragma unroll
    for (int i = 19; i >= 0; i--) {
        const double c = kDilogC[i];
        B0a = fma(ALa, B1a, c) - B2a; B2a = B1a; B1a = B0a;
        B0b = fma(ALb, B1b, c) - B2b; B2b = B1b; B1b = B0b;
    }
    double da = -(Sa * (B0a - Ha * B2a) + Aa), db = -(Sb * (B0b - Hb * B2b) + Ab);
    if (xa == 1.0) da = -PI12;                  // dilog(-1), comtot2d.f:385-386
    if (xb == 1.0) db = -PI12;
    const double la = SMOOTH ? log1p(xa) : log(1.0 + xa), lb = SMOOTH ? log1p(xb) : log(1.0 + xb);
    const double va = (-xa / 2.0 + 0.5 / (1.0 + xa)) + 4.0 * da + (9.0 + xa + 8.0 / xa) * la;
    const double vb = (-xb / 2.0 + 0.5 / (1.0 + xb)) + 4.0 * db + (9.0 + xb + 8.0 / xb) * lb;
    return va - vb;
}
__device__ __forceinline__ double intg_v_diff(double xa, double xb) { return intg_v_diff_t<false>(xa, xb); }

// The reference forms log(1.0 + y) (comtot2d.f:352-353 intg_v): the rounding of 1 + y costs a relative error of up to
// 1.1e-16 / y in the logarithm, which the factor 8 / y turns into noise that exceeds the signal of the high-gamma
// bins near the Thomson switch (y_b = 2 gamma (1 - beta) x ~ 1e-13) and reaches ~2e-4 of comtot for 1e-7 < x < 1e-3.
// That noise is part of the reference's answer.  It is reproduced exactly and cheaply: with s = fl(1 + y) and
// e = (s - 1) - y (exact), log(s) = log1p(y) + e / (1 + y) to first order, so
//   intg_v(y) = intg_v_smooth(y) + kn_eta(y),  kn_eta(y) = (9 + y + 8 / y) e / s.
// intg_v_smooth (log1p instead of log(1 + y)) is what the sigma_mode-1 table interpolates.
__device__ __forceinline__ double kn_eta(double y)
{
    if (!(y < 1.0)) return 0.0;                  // |eta| < 2e-15 there
    const double s = 1.0 + y, e = (s - 1.0) - y, d = y * s;
#ifdef C2D_EMU
    double r = (double)(1.0f / (float)d);
#else
    double r = (double)__frcp_rn((float)d);
#endif
    r = r * (2.0 - d * r);                        // one Newton step: ~1e-14 relative
    return fma(y, 9.0 + y, 8.0) * e * r;
}

// sigma_E of comtot2d.f:234-239 for electron bin i (0-based) and x = xnu/511.  kn_pref = 9.375e-2 sigma_T /
// gamma^2 / beta, kn_a1 = 2 gamma (1 + beta), kn_a2 = 2 gamma (1 - beta) are the reference's own
// left-to-right products, formed once on the host.
__device__ __forceinline__ double sigma_kn_exact(const Params &p, int i, double x)
{
    return p.kn_pref[i] / (x * x) * intg_v_diff(p.kn_a1[i] * x, p.kn_a2[i] * x);
}
__device__ __forceinline__ double sigma_kn_smooth(const Params &p, int i, double x)
{
    return p.kn_pref[i] / (x * x) * intg_v_diff_t<true>(p.kn_a1[i] * x, p.kn_a2[i] * x);
}
__device__ __forceinline__ bool kn_thomson(const Params &p, int i, double x) { return x * p.gam0[i] * p.opb[i] < 1.e-2; }
__device__ __forceinline__ double sigma_kn_bin(const Params &p, int i, double x)
{
    const double sigthom = 6.6516e-25;
    if (kn_thomson(p, i, x)) return sigthom * (1.0 - 2.0 * x * p.gam0[i]);
    return sigma_kn_exact(p, i, x);
}

// comac for up to four zones, one thread walking all electron bins: comtot2d.f:219-247
static __device__ __noinline__ void comtot_zones(const Params &p, double xnu, const int *zi, int nzn, double *comac)
{
    const double x = xnu / 5.11e2;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const double *f0 = p.f_nt + (size_t)zi[0] * p.num_nt;
    const double *f1 = p.f_nt + (size_t)zi[nzn > 1 ? 1 : 0] * p.num_nt;
    const double *f2 = p.f_nt + (size_t)zi[nzn > 2 ? 2 : 0] * p.num_nt;
    const double *f3 = p.f_nt + (size_t)zi[nzn > 3 ? 3 : 0] * p.num_nt;
    for (int i = 0; i < p.num_nt - 1; i++) {
        const double sE = sigma_kn_bin(p, i, x), dg = p.dgam[i];
        c0 = c0 + sE * f0[i] * dg;
        c1 = c1 + sE * f1[i] * dg;
        c2 = c2 + sE * f2[i] * dg;
        c3 = c3 + sE * f3[i] * dg;
    }
    const double c[4] = {c0, c1, c2, c3};
    for (int m = 0; m < nzn; m++) comac[m] = (c[m] < 1.e-40) ? 1.e-40 : p.n_e[zi[m]] * c[m];
}

// The same sum with the warp cooperating on ONE photon energy: electron bins strided over the lanes
// (coalesced f_nt rows, one K-N evaluation per lane and bin, no divergence on the photon energy), then a
// butterfly reduction.  Every lane of the warp must call it; returns comac[0..3] for the zones z0..z3 (nzn used).
__device__ __forceinline__ void comtot_zones_warp(const Params &p, double xnu, int z0, int z1, int z2, int z3, int nzn,
                                                  double (&comac)[4])
{
    const int lane = lane_id();
    const double x = xnu / 5.11e2;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const double *f0 = p.f_nt + (size_t)z0 * p.num_nt, *f1 = p.f_nt + (size_t)z1 * p.num_nt;
    const double *f2 = p.f_nt + (size_t)z2 * p.num_nt, *f3 = p.f_nt + (size_t)z3 * p.num_nt;
    for (int i = lane; i < p.num_nt - 1; i += kWarp) {
        const double sE = sigma_kn_bin(p, i, x), dg = p.dgam[i];
        c0 = c0 + sE * f0[i] * dg;
        if (nzn > 1) c1 = c1 + sE * f1[i] * dg;
        if (nzn > 2) c2 = c2 + sE * f2[i] * dg;
        if (nzn > 3) c3 = c3 + sE * f3[i] * dg;
    }
    c0 = warp_sum(c0);
    if (nzn > 1) c1 = warp_sum(c1);
    if (nzn > 2) c2 = warp_sum(c2);
    if (nzn > 3) c3 = warp_sum(c3);
    comac[0] = (c0 < 1.e-40) ? 1.e-40 : p.n_e[z0] * c0;
    comac[1] = (c1 < 1.e-40) ? 1.e-40 : p.n_e[z1] * c1;
    comac[2] = (c2 < 1.e-40) ? 1.e-40 : p.n_e[z2] * c2;
    comac[3] = (c3 < 1.e-40) ? 1.e-40 : p.n_e[z3] * c3;
}

// one zone, warp-cooperative (every lane must call; all lanes get the result)
__device__ __forceinline__ double comtot_zone_warp(const Params &p, double xnu, int zi)
{
    const int lane = lane_id();
    const double x = xnu / 5.11e2;
    double c0 = 0;
    const double *f0 = p.f_nt + (size_t)zi * p.num_nt;
    for (int i = lane; i < p.num_nt - 1; i += kWarp) c0 = c0 + sigma_kn_bin(p, i, x) * f0[i] * p.dgam[i];
    c0 = warp_sum(c0);
    return (c0 < 1.e-40) ? 1.e-40 : p.n_e[zi] * c0;
}

// sigma_mode 1: comtot(zone, E) from the per-step table of c2d_sigma.cu.  The bins split at m(x) = number of
// bins on the Thomson branch of comtot2d.f:228 (the same floating-point predicate, found by bisection -- it is
// monotone in the bin index):  c = sigma_T (A_z[m] - 2 x B_z[m]) + S_z(m, x),  where A, B are prefix sums over the
// Thomson bins and S_z(m, .) -- the sum of the exact branch over bins >= m, a smooth function of x on the interval
// where m(x) = m -- is a degree-7 Chebyshev series in ln x on that interval (m = 0: sub-intervals of width 1/sig_inv_h).
// The table reproduces the per-history evaluation to ~1e-11 (the cancellation noise of the exact branch itself).
struct SigCtx { int m, j, ia, ib; double x, s; };      // j = -1: all bins on the Thomson branch; j = -2: outside the table;
                                                        // ia / ib: see comtot_table_zone (functions of x only)

__device__ __forceinline__ void sig_begin(const Params &p, double xnu, SigCtx &c)
{
    c.x = xnu / 5.11e2;
    int lo = 0, hi = p.num_nt - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (kn_thomson(p, mid, c.x)) lo = mid + 1; else hi = mid;
    }
    c.m = lo; c.s = 0.0;
    if (c.m == p.num_nt - 1) { c.j = -1; return; }
    const double t = log(c.x);
    if (c.m > 0) c.j = c.m;
    else {
        const double u = (t - p.sig_t0) * p.sig_inv_h;
        const int sub = (u > 0.0) ? (int)u : 0;
        if (!(u < (double)p.sig_nsub)) { c.j = -2; return; }
        c.j = p.num_nt - 1 + sub;
    }
    c.s = (t - p.sig_lo[c.j]) * p.sig_iw[c.j] - 1.0;
    {   // ia = first bin with a1_i x >= 0.1 ; ib = first bin with a2_i x < 1e-3 (a1 grows, a2 falls with the bin index)
        const int N = p.num_nt;
        int lo = c.m, hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.kn_a1[mid] * c.x < 0.1) lo = mid + 1; else hi = mid; }
        c.ia = lo;
        lo = c.m; hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.kn_a2[mid] * c.x < 1.e-3) hi = mid; else lo = mid + 1; }
        c.ib = lo;
    }
}

// Rounding-noise term of one electron bin (see kn_eta): pref_i (eta(a1_i x) - eta(a2_i x)).  The a-side noise is
// < 1e-12 of the bin once y_a = a1_i x > 0.1 (it falls like 1 / y^3 from 1.6e-10 at the branch switch).  On the b side
// y_b = a2_i x is tiny for all but the lowest-gamma bins, where (9 + y + 8 / y) / (1 + y) = (8 / y) (1 + y / 8 + O(y^3)):
// with kn_cb = 8 pref_i / a2_i formed once on the host the bin costs no division.  a1 grows and a2 falls with the bin
// index, so "y_a < 0.1" and "y_b >= 1e-3" are index ranges [.., ia) and [.., ib), found once per photon energy.
//
// Noise sums  sum_{i >= m} noise_i(x) W[z][i]  for up to four zones (rows beyond nzn are not read).
__device__ __forceinline__ void sig_noise_zones(const Params &p, double x, int m, const int *zi, int nzn, double (&ee)[4])
{
    const int N = p.num_nt;
    const double *w0 = p.sig_W + (size_t)zi[0] * N, *w1 = p.sig_W + (size_t)zi[nzn > 1 ? 1 : 0] * N;
    const double *w2 = p.sig_W + (size_t)zi[nzn > 2 ? 2 : 0] * N, *w3 = p.sig_W + (size_t)zi[nzn > 3 ? 3 : 0] * N;
    double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    const double inv_x = 1.0 / x;
    int ia, ib;
    {   // ia = first bin with a1_i x >= 0.1 ; ib = first bin with a2_i x < 1e-3
        int lo = m, hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.kn_a1[mid] * x < 0.1) lo = mid + 1; else hi = mid; }
        ia = lo;
        lo = m; hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.kn_a2[mid] * x < 1.e-3) hi = mid; else lo = mid + 1; }
        ib = lo;
    }
    // bins [m, i_fast): the full expressions (few bins: the lowest-gamma ones and those just above the branch switch)
    const int i_fast = (ia > ib) ? ia : ib;
    int i = m;
    for (; i < i_fast && i < N - 1; i++) {
        const double yb = p.kn_a2[i] * x;
        double q;
        if (i >= ib) {
            const double s = 1.0 + yb, e = (s - 1.0) - yb;
            q = -(p.kn_cb[i] * inv_x) * (e * fma(0.125, yb, 1.0));
        } else q = -(p.kn_pref[i] * kn_eta(yb));
        if (i < ia) q = q + p.kn_pref[i] * kn_eta(p.kn_a1[i] * x);
        e0 = fma(q, w0[i], e0);
        if (nzn > 1) e1 = fma(q, w1[i], e1);
        if (nzn > 2) e2 = fma(q, w2[i], e2);
        if (nzn > 3) e3 = fma(q, w3[i], e3);
    }
    // bins [i_fast, N-1): b side only, division-free, branch-free
    const double *a2 = p.kn_a2, *cb = p.kn_cb;
    for (; i < N - 1; i++) {
        const double yb = c2d_ldg(a2 + i) * x;
        const double s = 1.0 + yb, e = (s - 1.0) - yb;
        const double q = -(c2d_ldg(cb + i) * inv_x) * (e * fma(0.125, yb, 1.0));
        e0 = fma(q, c2d_ldg(w0 + i), e0);
        if (nzn > 1) e1 = fma(q, c2d_ldg(w1 + i), e1);
        if (nzn > 2) e2 = fma(q, c2d_ldg(w2 + i), e2);
        if (nzn > 3) e3 = fma(q, c2d_ldg(w3 + i), e3);
    }
    ee[0] = e0; ee[1] = e1; ee[2] = e2; ee[3] = e3;
}

// Table value for up to four zones of one photon energy, given the noise sums `ee` (zero when c.j < 0)
__device__ __forceinline__ void comtot_table_finish(const Params &p, const SigCtx &c, const int *zi, int nzn, const double (&ee)[4],
                                                    double *comac)
{
    const double sigthom = 6.6516e-25;
    const int N = p.num_nt;
    const double inv_x2 = 1.0 / (c.x * c.x);
    for (int k = 0; k < nzn; k++) {
        const int z = zi[k];
        const double *ab = p.sig_AB + ((size_t)z * N + c.m) * 2;
        double v = sigthom * (ab[0] - 2.0 * c.x * ab[1]);
        if (c.j >= 0) {
            const double *kk = p.sig_coef + ((size_t)z * p.sig_nint + c.j) * 8;
            const double s2 = c.s + c.s;
            double b1 = kk[7], b2 = 0.0, b0;
#pragma unroll
            for (int q = 6; q >= 1; q--) { b0 = fma(s2, b1, kk[q]) - b2; b2 = b1; b1 = b0; }
            v = v + (fma(c.s, b1, kk[0]) - b2) + ee[k] * inv_x2;
        }
        comac[k] = (v < 1.e-40) ? 1.e-40 : p.n_e[z] * v;
    }
}
__device__ __forceinline__ void comtot_table_zones(const Params &p, const SigCtx &c, const int *zi, int nzn, double *comac)
{
    double ee[4] = {0.0, 0.0, 0.0, 0.0};
    if (c.j >= 0) sig_noise_zones(p, c.x, c.m, zi, nzn, ee);
    comtot_table_finish(p, c, zi, nzn, ee, comac);
}

// One zone, with the noise sum cut where it no longer matters.  The rounding-noise term of bin i is bounded by
//   (1.2e-16 / x) (1.125 cb_i + 9 pref_i / a1_i) W[z][i]      (|e| <= 2^-53; b side (1 + y/8) <= 1.125; a side (9 + y + 8/y) <= 9/y, y < 0.1)
// and NB[z][i] holds the suffix sum of those bounds over the bins >= i (k_sigma_prefix).  The electron spectra fall off
// steeply, so the tail of the loop contributes < 1e-13 of the value long before the last bin: checked every 8 bins.
__device__ __forceinline__ double comtot_table_zone(const Params &p, const SigCtx &c, int z)
{
    const double sigthom = 6.6516e-25;
    const int N = p.num_nt;
    const double *ab = p.sig_AB + ((size_t)z * N + c.m) * 2;
    double v = sigthom * (c2d_ldg(ab) - 2.0 * c.x * c2d_ldg(ab + 1));
    if (c.j >= 0) {
        const double *kk = p.sig_coef + ((size_t)z * p.sig_nint + c.j) * 8;
        const double s2 = c.s + c.s;
        double b1 = c2d_ldg(kk + 7), b2 = 0.0, b0;
#pragma unroll
        for (int q = 6; q >= 1; q--) { b0 = fma(s2, b1, c2d_ldg(kk + q)) - b2; b2 = b1; b1 = b0; }
        v = v + (fma(c.s, b1, c2d_ldg(kk)) - b2);
        // noise: sum_{i >= m} noise_i(x) W[z][i], bins [i0, i_end) where the bound still exceeds 1e-13 |v|
        const double x = c.x, inv_x = 1.0 / x;
        const double *w = p.sig_W + (size_t)z * N, *nb = p.sig_NB + (size_t)z * N, *cw = p.sig_CW + (size_t)z * N;
        const double thr = fabs(v) * (x * x * x) * 8.0e2;          // bound * 1.2e-16 / x^3 <= 1e-13 |v|
        int i = c.m;
        const int ifz = p.sig_if[z];
        if (i < ifz) i = ifz;
        double e0 = 0.0;
        if (i < N - 1 && c2d_ldg(nb + i) > thr) {
            int i_end;
            {   // first bin whose suffix bound is below the threshold (NB is non-increasing)
                int lo = i + 1, hi = N - 1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (c2d_ldg(nb + mid) > thr) lo = mid + 1; else hi = mid; }
                i_end = lo;
            }
            const int ia = c.ia, ib = c.ib;
            const int i_full = (ib < i_end) ? ib : i_end;
            for (; i < i_full; i++) {                             // a2_i x >= 1e-3 (hard photons, lowest-gamma bins): both sides in full
                double q = -(p.kn_pref[i] * kn_eta(p.kn_a2[i] * x));
                if (i < ia) q = q + p.kn_pref[i] * kn_eta(p.kn_a1[i] * x);
                e0 = fma(q, w[i], e0);
            }
            // bins just above the branch switch (a1_i x < 0.1, ~17 of them): the a side as well.  eta(y) = (8 + y (9 + y)) e /
            // (y (1 + y)); the noise term needs 1e-4 relative at most (it is <= 2e-4 of sigma, wanted to 1e-10): the reciprocal
            // is taken in single precision, only e -- the rounding error of 1 + y itself -- is exact.
            {
                const int i_a = (ia < i_end) ? ia : i_end;
                double sa = 0.0, sb = 0.0;
#pragma unroll kNoiseUnroll
                for (; i < i_a; i++) {
                    const double ya = c2d_ldg(p.kn_a1 + i) * x, yb = c2d_ldg(p.kn_a2 + i) * x;
                    const double ta = 1.0 + ya, tb = 1.0 + yb;
                    const double ea = (ta - 1.0) - ya, eb = (tb - 1.0) - yb;
#ifdef C2D_EMU
                    const double ra = (double)(1.0f / (float)(ya * ta));
#else
                    const double ra = (double)__frcp_rn((float)(ya * ta));
#endif
                    sa = fma(c2d_ldg(p.kn_pref + i) * c2d_ldg(w + i), fma(ya, 9.0 + ya, 8.0) * (ea * ra), sa);
                    sb = fma(c2d_ldg(cw + i), eb * fma(0.125, yb, 1.0), sb);
                }
                e0 = e0 + sa - sb * inv_x;
            }
            // b side only: -(cb_i W_i / x) e_i (1 + y_i / 8), e_i = (fl(1 + y_i) - 1) - y_i the rounding error of 1 + y_i.
            // cb_i W_i is tabulated per zone (sig_CW); two independent accumulators, -1 / x applied once.
            const double *a2 = p.kn_a2;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll kNoiseUnroll
            for (; i + 1 < i_end; i += 2) {
                const double y0 = c2d_ldg(a2 + i) * x, y1 = c2d_ldg(a2 + i + 1) * x;
                const double t0 = 1.0 + y0, t1 = 1.0 + y1;
                const double r0 = (t0 - 1.0) - y0, r1 = (t1 - 1.0) - y1;
                s0 = fma(c2d_ldg(cw + i), r0 * fma(0.125, y0, 1.0), s0);
                s1 = fma(c2d_ldg(cw + i + 1), r1 * fma(0.125, y1, 1.0), s1);
            }
            if (i < i_end) {
                const double y0 = c2d_ldg(a2 + i) * x, t0 = 1.0 + y0, r0 = (t0 - 1.0) - y0;
                s0 = fma(c2d_ldg(cw + i), r0 * fma(0.125, y0, 1.0), s0);
            }
            e0 = e0 - (s0 + s1) * inv_x;
        }
        v = v + e0 * (inv_x * inv_x);
    }
    return (v < 1.e-40) ? 1.e-40 : p.n_e[z] * v;
}

// comtot for one zone by one thread (scat_flag = 1 flights; `c` was set up by flight_begin when sigma_mode = 1)
__device__ __forceinline__ double comtot_one(const Params &p, int zi, double xnu, const SigCtx &c)
{
    double r;
    if (p.sigma_mode == 1 && c.j != -2) r = comtot_table_zone(p, c, zi);
    else comtot_zones(p, xnu, &zi, 1, &r);
    return r;
}

// ------------------------------------------------------------------------------------------ scattering
// Electron-bin tally of nth2d (nontherm2d.f:180).  A speculative caller (the 3rd-split attempts of k_split3)
// passes a record instead of tallying: up to four bins packed 16 bits each, `n` counts every draw.
struct ElRec { uint64_t bins; int n; };

constexpr int kNtGuide = 256;
// nontherm2d.f:159-183
__device__ __forceinline__ void nth2d(const Params &p, Rng &rng, int zi, double &gamm, double &betb, int &ibin, ElRec *rec)
{
    const double *P = p.Pnt + (size_t)zi * p.num_nt;
    double rnum = rng.next(p);
    rnum = (double)(int)(rnum * 1.e6) / 1.e6 + 1.e-6 * rng.next(p);
    // first i in [2, num_nt] (1-based) with Pnt(i) > rnum; Pnt is a CDF (non-decreasing).  The step's guide table (k_nt_guide)
    // brackets the answer for each 1/256 of rnum: the same bisection, started 0 - 2 steps from its end instead of 8
    int lo = 2, hi = p.num_nt;
    if (p.nt_guide) {
        int b = (int)(rnum * (double)kNtGuide);
        if (b > kNtGuide - 1) b = kNtGuide - 1;
        const uint32_t g = c2d_ldg(p.nt_guide + (size_t)zi * kNtGuide + b);
        lo = (int)(g & 0xFFFFu); hi = (int)(g >> 16);
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P[mid - 1] > rnum) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    gamm = sqrt(p.gnt[i - 1] * p.gnt[i - 2]) + 1.0;
    betb = sqrt(1.0 - 1.0 / (gamm * gamm));
    if (rec) {
        if (rec->n < 4) rec->bins |= (uint64_t)i << (16 * rec->n);
        rec->n++;
    } else atomicAdd(p.nelectron + i, 1ull);
    ibin = i;
}

// compb_2d.f:55-95 (label 100): draw the electron and the collision cosine, K-N rejection
struct ElDraw { double gamm, betb, omeg, znue; };
// one trip through label 100: true = accepted.  Draws: 2 (nth2d) + omeg + tl [+ 1 for the K-N test unless znue < 1e-10].
// sigma_KN / sigma_T at photon energy znue (electron rest frame), compb_2d.f:75-88
__device__ __forceinline__ double kn_xknot(double znue)
{
    if (znue <= 1.e-2) return 1.0 - znue * (2.0 - znue * (5.2 - znue * (13.3 - 1.144e3 * znue / 3.5e1)));
    const double znue3 = znue * znue * znue;
    const double betz = 1.0 + 2.0 * znue;
    const double gamz = znue * (znue - 2.0) - 2.0;
    const double xxx = 4.0 * znue + 2.0 * znue3 * (1.0 + znue) / (betz * betz) + gamz * log(betz);
    return 3.75e-1 * xxx / znue3;
}
// the trip after its nth2d draws: collision cosine (2 draws), K-N test (1 draw unless znue < 1e-10)
__device__ __forceinline__ bool compb2d_trip_rest(const Params &p, Rng &rng, double znu, ElDraw &e)
{
    e.omeg = 2.0 * rng.next(p) - 1.0;
    C2D_CLAMP8(e.omeg);
    const double tl = rng.next(p);
    const double tr = 0.5 * (1.0 - e.betb * e.omeg);
    if (tl > tr) e.omeg = -e.omeg;
    C2D_CLAMP8(e.omeg);
    e.znue = (1. - e.betb * e.omeg) * znu * e.gamm;
    if (e.znue < 1.e-10) return false;
    return !(rng.next(p) > kn_xknot(e.znue));
}
__device__ __forceinline__ bool compb2d_electron_once(const Params &p, Rng &rng, int zi, double znu, ElDraw &e, int &i_gam, ElRec *rec)
{
    nth2d(p, rng, zi, e.gamm, e.betb, i_gam, rec);
    return compb2d_trip_rest(p, rng, znu, e);
}
// For znu >= kKnTeamZnu every trip consumes exactly 5 draws (znue >= (1 - 0.99999999) znu gamma >= 1e-10 always), so trip k of a
// stream starts at draw 5 k and a TEAM of threads can try trips in parallel (k_children_hard): the rejection loop of a hard
// photon deep in the Klein-Nishina regime takes 1 / <sigma_KN / sigma_T> ~ 10^3 .. 10^6 trips.
constexpr double kKnTeamZnu = 2.e-2;
// max_iter > 0: give up after that many rejected trips (returns false; their nth2d tallies are made, the cursor is at 5 max_iter)
__device__ __forceinline__ bool compb2d_electron(const Params &p, Rng &rng, int zi, double znu, ElDraw &e, int &i_gam, ElRec *rec, int max_iter = 0)
{
    for (int it = 0;; it++) {                        // label 100
        if (max_iter > 0 && it >= max_iter) return false;
        if (compb2d_electron_once(p, rng, zi, znu, e, i_gam, rec)) return true;
    }
}

// compb_2d.f:96-318: rest-frame energy ratio, azimuths, Lorentz back-transform, new direction and bins.
// `ew_min` >= 0: a speculative caller (3rd-split passes) only wants the photon if its new weight exceeds ew_min; the
// function returns false as soon as the scattered energy is known to miss it (the rest of the pass's stream is unused).
__device__ __forceinline__ bool compb2d_finish(const Params &p, Rng &rng, Photon &P, const ElDraw &e, double ew_min = -1.0)
{
    const double fuzz = 1.e-10, emasskev = 5.11e2;
    const double znu = P.xnu / emasskev;
    const double gamm = e.gamm, betb = e.betb, omeg = e.omeg, znue = e.znue;
    double sz, games;
    const double betz = 1.0 + 2.0 * znue;
    for (;;) {                                       // labels 200/202
        sz = (1.0 + 2.0 * znue * rng.next(p)) / betz;
        games = 1.0 + (1.0 - 1.0 / sz) / znue;
        if ((1.0 - games * games) < 0.) continue;
        const double tr = games * games - 1.0 + sz + 1.0 / sz;
        const double phat = betz + 1.0 / betz;
        if (rng.next(p) * phat > tr) continue;
        break;
    }
    const double znues = znue * sz;
    double wa, wb, swa;
    do {                                             // label 210
        wa = rng.next(p);
        wb = 2.0 * rng.next(p) - 1.0;
        swa = wa * wa + wb * wb;
    } while (swa >= 1.0 || swa <= 1.e-20);
    const double cazes = (wa * wa - wb * wb) / swa;
    double omege = (omeg - betb) / (1. - betb * omeg);
    C2D_CLAMP8(omege);
    double omeges = games * omege + cazes * sqrt((1. - omege * omege + fuzz) * (1. - games * games));
    C2D_CLAMP8(omeges);
    const double znus = (1.0 + betb * omeges) * gamm * znues;
    if (ew_min >= 0.0 && !(P.ew * (znus * emasskev) / P.xnu > ew_min)) return false;
    double gams = 1.0 - (znue - znues) / (znu * znus);
    C2D_CLAMP8(gams);
    do {                                             // label 220
        wa = rng.next(p);
        wb = 2.0 * rng.next(p) - 1.0;
        swa = wa * wa + wb * wb;
    } while (swa >= 1.0 || swa <= 1.e-20);
    double cazs = (wa * wa - wb * wb) / swa;
    C2D_CLAMP8(cazs);
    double wmus = P.wmu * gams + cazs * sqrt((1.0 - gams * gams) * (1.0 - P.wmu * P.wmu + fuzz));
    C2D_CLAMP8(wmus);
    const double xnus = znus * emasskev;
    double cosdphi = (gams - P.wmu * wmus) / sqrt((1.0 - P.wmu * P.wmu) * (1.0 - wmus * wmus));
    C2D_CLAMP8(cosdphi);
    const double dphi = acos(cosdphi);
    const double phis = P.phi + dphi;
    P.jgpsp = sp_group(p, xnus, 1.000001e0, .999999e0, 0);
    P.jgplc = lc_group(p, xnus);
    P.jgpmu = mu_group(p, wmus);
    P.ew = P.ew * xnus / P.xnu;
    P.xnu = xnus;
    P.wmu = wmus;
    P.phi = phis;
    return true;
}

// compb_2d.f:1-318; P.jph/kph give the zone
static __device__ __noinline__ void compb2d(const Params &p, Rng &rng, Photon &P, int &i_gam)
{
    ElDraw e;
    compb2d_electron(p, rng, (P.jph - 1) * p.nr + (P.kph - 1), P.xnu / 5.11e2, e, i_gam, nullptr);
    compb2d_finish(p, rng, P, e);
}

// ------------------------------------------------------------------------------------------ imcleak
__device__ __forceinline__ void escape_tally(const Params &p, const Photon &P, double t_bound)
{   // imcleak2d.f:171-175
    const unsigned long long n = atomicAdd(p.counters + CNT_NEVENTS, 1ull);
    if (n < p.event_cap) {
        double *e = p.events + 7 * n;
        e[0] = t_bound; e[1] = P.xnu; e[2] = P.ew; e[3] = P.rpre; e[4] = P.zpre; e[5] = P.wmu; e[6] = P.phi;
    }
    if (P.jgplc > 0) atomic_add_f64(p.edout + (P.jgpmu - 1) * kNlcMax + (P.jgplc - 1), P.ew / p.dt);
    if (P.jgpsp > 0 && p.spec_switch == 0) atomic_add_f64(p.fout + (P.jgpmu - 1) * p.nphtotal + (P.jgpsp - 1), P.ew);
}

// imcleak2d.f:329-405
__device__ __forceinline__ void get_bin(const Params &p, Photon &P)
{
    P.jgpsp = sp_group(p, P.xnu, (double)1.000001f, (double).999999f, 0);
    P.jgplc = lc_group(p, P.xnu);
    P.jgpmu = mu_group(p, P.wmu);
}

// imcleak2d.f:126-146 == 224-241: energy of the Compton-reflected photon from the Green's-function tables of ref_matrix.f
// (linear scans as in the reference: a rare event).  Two uniforms.  Returns W_abs(n_out, n_in); P.xnu is replaced.
__device__ __forceinline__ double reflect_energy(const Params &p, Rng &rng, Photon &P)
{
    const int n_ref = p.n_ref;
    int n_in = 1, n_out = 1;
    while (p.E_ref[n_in - 1] < P.xnu && n_in < n_ref) n_in++;
    const double rnum = rng.next(p);
    const double *col = p.P_ref + (size_t)(n_in - 1) * n_ref;
    while (col[n_out - 1] < rnum && n_out < n_ref) n_out++;
    if (n_out > 1) P.xnu = p.E_ref[n_out - 2] + rng.next(p) * (p.E_ref[n_out - 1] - p.E_ref[n_out - 2]);
    else P.xnu = p.E_ref[n_out - 1];
    return p.W_abs[(size_t)(n_in - 1) * n_ref + (n_out - 1)];
}

// The reflection options of imcleak (cr_sent != 0), out of line so that the common cr_sent = 0 path keeps its registers:
//   lower surface, cr_sent = 1 / 3 / 4 (imcleak2d.f:107-165): specular mirror where the boundary has no temperature (or
//     cr_sent = 4), Compton reflection through the tables where it is hot; the flight goes on in row 1 -> returns 0
//   outer cylinder, cr_sent = 2 / 3 and wmu <= 0 (l.216-273): Compton reflection at the outer disk, recorded as escaped -> 1
//   anything else -> -1 (the caller takes the cr_sent = 0 path)
static __device__ __noinline__ int imcleak_reflect(const Params &p, Photon &P, Rng &rng)
{
    const double xnu_sv = P.xnu;
    if (!(P.jph > 0)) {
        if (!(p.cr_sent == 1 || p.cr_sent == 3 || p.cr_sent == 4)) return -1;
        if (p.tbbl[P.kph - 1] <= 0.0 || p.cr_sent == 4) {         // l.113-122: wmu -> -wmu
            P.wmu = -P.wmu;
            if (P.jgpsp > 0 && p.spec_switch == 1) atomic_add_f64(p.fout + (P.jgpmu - 1) * p.nphtotal + (P.jgpsp - 1), P.ew);
        } else {                                                  // l.124-158
            const double w = reflect_energy(p, rng, P);
            const double ewnew = P.ew * w * P.xnu / xnu_sv;
            atomic_add_f64(p.Ed_ref + (P.kph - 1), ewnew);
            P.wmu = fabs(P.wmu);
            P.ew = ewnew;
        }
        get_bin(p, P);                                            // label 60, l.160-165
        P.jph = 1;
        return 0;
    }
    if (P.jph != p.nz + 1 && !(P.wmu > 0. || p.cr_sent == 1 || p.cr_sent == 4)) {
        const double w = reflect_energy(p, rng, P);
        P.ew = P.ew * w * P.xnu / xnu_sv;
        double tb = 1.e20;
        if (fabs(P.wmu) > 1.e-6) {
            tb = p.time + p.dt - kRadCp * (P.dcen + P.zpre / P.wmu);
            P.zpre = 0.;
            const double f = P.zpre * sqrt(1. - P.wmu * P.wmu) / fabs(P.wmu);          // zpre already zeroed (l.249-250): f = 0
            P.rpre = sqrt(P.rpre * P.rpre + f * f + 2. * P.rpre * f * cos(P.phi));
        }
        P.wmu = rng.next(p);
        get_bin(p, P);
        if (p.ncycle > 0) escape_tally(p, P, tb);
        return 1;
    }
    return -1;
}

// imcleak2d.f:2-320.  Returns idead; 0 = the flight goes on (through the axis, or reflected: P.xnu may have changed).
__device__ __forceinline__ int imcleak(const Params &p, Photon &P, Rng &rng)
{
    const double t_bound = p.time + p.dt - kRadCp * P.dcen;     // l.203,302 (lower face: see DESIGN.md)
    if (P.kph == 0) {
        if (p.rmin > 1.e-10) { atomic_add_f64(p.erlki + (P.jph - 1), P.ew); return 1; }
        P.phi = 1.e-6; P.kph = 1;
        return 0;
    }
    if (!(P.jph > 0)) {
        if (p.tbbl[P.kph - 1] > 0.0) {
            atomic_add_f64(p.Ed_in + (P.kph - 1), P.ew);
            atomic_add_f64(p.erlkl + (P.kph - 1), P.ew);
        }
        if (p.cr_sent != 0) { const int r = imcleak_reflect(p, P, rng); if (r >= 0) return r; }
        if (p.ncycle > 0) escape_tally(p, P, t_bound);
        return 1;
    }
    if (P.jph != p.nz + 1) {
        atomic_add_f64(p.erlko + (P.jph - 1), P.ew);
        if (p.cr_sent != 0) { const int r = imcleak_reflect(p, P, rng); if (r >= 0) return r; }
        if (p.ncycle > 0) escape_tally(p, P, t_bound);
        return 1;
    }
    atomic_add_f64(p.erlku + (P.kph - 1), P.ew);
    if (p.ncycle > 0 && P.wmu < (double)0.98f) escape_tally(p, P, t_bound);
    return 1;
}

// census append: imctrk2d.f:541-586, in two halves so that the history kernel can commit the per-zone half (npcen, ecens)
// through its warp / block tally hierarchy
__device__ __forceinline__ void census_zone_tally(const Params &p, int zi, double ew)
{
    atomicAdd(p.npcen + zi, 1ull);
    atomic_add_f64(p.ecens + zi, ew);
}
__device__ __forceinline__ bool census_record(const Params &p, const Photon &P, Rng &rng)
{
    const int zi = (P.jph - 1) * p.nr + (P.kph - 1);
    if (p.pair_switch == 1) {
        int i_gg;
        for (i_gg = 1; i_gg <= p.n_gg; i_gg++)
            if (P.xnu < p.E_gg[i_gg - 1]) break;
        const double Egg_min = (p.E_gg[0] * p.E_gg[0]) / p.E_gg[1];
        if (P.xnu > Egg_min) {
            const double v = (P.ew * 6.25e8) / P.xnu;
            if (i_gg <= p.n_gg) atomic_add_f64(p.n_ph + (size_t)zi * p.n_gg + (i_gg - 1), v);
            else if (P.jph + 1 <= p.nz)   // n_ph(n_gg+1,j,k) aliases n_ph(1,j+1,k) in COMMON /photonfield/
                atomic_add_f64(p.n_ph + (size_t)(zi + p.nr) * p.n_gg, v);
        }
    }
    {
        const int i_f = scan_lt(p.E_field, p.nphfield, P.xnu);
        const double Egg_min = (p.E_field[0] * p.E_field[0]) / p.E_field[1];
        if (P.xnu > Egg_min) atomic_add_f64(p.n_field + (size_t)zi * p.nphfield + (i_f - 1), 6.25e8 * P.ew / P.xnu);
    }
    const unsigned long long n = atomicAdd(p.counters + CNT_NDXOUT, 1ull);
    if (n >= p.census_cap) { count(p, CNT_FATAL); return false; }
    p.out.rpre[n] = P.rpre; p.out.zpre[n] = P.zpre; p.out.wmu[n] = P.wmu; p.out.phi[n] = P.phi;
    p.out.ew[n] = P.ew; p.out.xnu[n] = P.xnu;
    p.out.bins[n] = (uint32_t)P.jgpsp | ((uint32_t)P.jgplc << 16) | ((uint32_t)P.jgpmu << 24);
    p.out.zone[n] = (uint32_t)P.jph | ((uint32_t)P.kph << 16);
    p.out.sid[n] = rng.sid;
    p.out.seed[n] = (int)(rng.next(p) * 1.e5);
    return true;
}
__device__ __forceinline__ bool census_write(const Params &p, const Photon &P, Rng &rng)
{
    census_zone_tally(p, (P.jph - 1) * p.nr + (P.kph - 1), P.ew);
    return census_record(p, P, rng);
}

// ------------------------------------------------------------------------------------------ the flight
// One imctrk2d flight from label 110 (imctrk2d.f:158) until it ends, as a sequence of zone segments.
//   MODE_PRE    collision-free ray of a history: no tallies, no draws; records the zone crossings
//   MODE_TRIAL  scat_flag = -1 trial with a given colmfp: deposits absorption, stops at collision
//   MODE_F0     scat_flag =  0 recombined flight: census / leak tallies, no deposition
//   MODE_CHILD  scat_flag =  1 scattered photon: everything
constexpr int STEP_CONT = -1;

struct FlightCtx {           // energy bins are constant along a flight (xnu only changes in compb2d)
    int i_e, i_gg;
    double dcol_f0;
    SigCtx sig;
    uint64_t abs_S;          // trial flights: absorption-point draws come from abs_uniform(abs_S, abs_t, abs_k++)
    uint32_t abs_t, abs_k;
};
__device__ __forceinline__ void flight_begin(const Params &p, double xnu, FlightCtx &c)
{
    c.i_e = scan_lt(p.E_ph, p.n_vol, xnu);
    c.i_gg = 0;
    if (p.pair_switch == 1)
        for (c.i_gg = 1; c.i_gg <= p.n_gg; c.i_gg++)
            if (xnu < p.E_gg[c.i_gg - 1]) break;
    c.dcol_f0 = 100 * fmax(p.r[p.nr - 1], p.z[p.nz - 1]);
    c.sig.j = -2;
}

// One segment (imctrk2d.f:207-586) of a photon with ew >= 1e-40 and |wmu| <= 1 in zone zi; sigsc is the
// scattering opacity of this segment (comac * pair_enhance; ignored for MODE_PRE / MODE_F0).
// Returns STEP_CONT when the flight goes on in the next zone, or the END_* code.
template <int MODE>
__device__ __forceinline__ int fly_step(const Params &p, Photon &P, Rng &rng, double wtmin, double &colmfp, double sigsc,
                                        FlightCtx &c, Seg *segs, int &s, int &nseg)
{
    const int zi = (P.jph - 1) * p.nr + (P.kph - 1);
    const double pair_enhance = (p.pair_switch == 1) ? 1.0 + 2.0 * p.f_pair[zi] : 1.0;
    const int i_e = c.i_e, i_gg = c.i_gg;
    if (MODE == MODE_PRE || MODE == MODE_F0) sigsc = 0.0;
    const double xqsqleft = (P.kph == 1) ? p.rmin * p.rmin : p.r[P.kph - 2] * p.r[P.kph - 2];
    double dcol;
    if (MODE == MODE_TRIAL || MODE == MODE_TRIAL_SLOW || MODE == MODE_CHILD) dcol = colmfp / sigsc;
    else if (MODE == MODE_F0) dcol = c.dcol_f0;
    else dcol = INFINITY;
    double trld; int ikind;
    if (P.dcen <= dcol) { trld = P.dcen; ikind = 2; } else { trld = dcol; ikind = 3; }
    // geometry, imctrk2d.f:233-365
    double Eta = cos(P.phi);
    const int Eta_switch = (P.phi <= kPi && P.phi >= 1.e-10) ? 1 : -1;
    const double disp = Eta * P.rpre;
    const double psq = P.rpre * P.rpre * (1. - Eta * Eta);
    int kbnd, inout; double rbnd;
    if (Eta < 0.0 && psq < xqsqleft) {
        kbnd = P.kph - 1; inout = -1;
        rbnd = (P.kph > 1) ? p.r[P.kph - 2] : p.rmin;
    } else {
        kbnd = P.kph; inout = 1;
        rbnd = p.r[P.kph - 1];
    }
    double dpbsq = rbnd * rbnd - psq;
    if (dpbsq < 1.e-6) dpbsq = 1.e-6;
    const double disbr = (double)inout * sqrt(dpbsq) - disp;
    double trldb = disbr / sqrt(1.0 - (P.wmu * P.wmu));
    double f = disbr;
    const double Zr = P.zpre + P.wmu * trldb;
    const double zhigh = p.z[P.jph - 1], zlow = (P.jph == 1) ? p.zmin : p.z[P.jph - 2];
    int knew, jnew; double Zbnd;
    if (Zr > zhigh || Zr < zlow) {
        const bool up = (Zr > zhigh);
        Zbnd = up ? zhigh : zlow; knew = P.kph; jnew = up ? P.jph + 1 : P.jph - 1;
        f = (Zbnd - P.zpre) * sqrt(1.0 - P.wmu * P.wmu) / P.wmu;
        rbnd = sqrt(P.rpre * P.rpre + f * f + 2.0 * P.rpre * f * Eta);
        trldb = sqrt(f * f + (Zbnd - P.zpre) * (Zbnd - P.zpre));
    } else {
        knew = P.kph + inout; jnew = P.jph;
        rbnd = (kbnd > 0) ? p.r[kbnd - 1] : p.rmin;
        Zbnd = Zr;
    }
    const double trldb_rec = trldb;
    double rnew, znew;
    if (trldb < trld) {
        ikind = 1; trld = trldb; rnew = rbnd; znew = Zbnd;
    } else {
        jnew = P.jph; knew = P.kph;
        f = trld * sqrt(1.0 - P.wmu * P.wmu);
        rnew = sqrt(f * f + P.rpre * P.rpre + 2.0 * f * P.rpre * Eta);
        znew = P.zpre + trld * P.wmu;
    }
    // absorption, imctrk2d.f:397-460
    double sigabs = 1.e-40;
    sigabs = sigabs + pair_enhance * p.kappa_tot[(size_t)zi * p.n_vol + (i_e - 1)];
    double kgg_i = 0.0;
    if (p.pair_switch == 1) {
        const double *kg = p.k_gg + (size_t)zi * p.n_gg;
        double kgg;
        if (i_gg <= p.n_gg) kgg_i = kg[i_gg - 1];
        if (P.xnu < p.E_gg[0]) kgg = (P.xnu / p.E_gg[0]) * kg[0];
        else if (i_gg == p.n_gg) kgg = kg[p.n_gg - 1];
        else if (i_gg > p.n_gg) kgg = 0.0;
        else {
            const double kgg0 = kg[i_gg - 1], kgg1 = kg[i_gg];
            kgg = kgg0 + (P.xnu - p.E_gg[i_gg - 1]) * (kgg1 - kgg0) / (p.E_gg[i_gg] - p.E_gg[i_gg - 1]);
        }
        sigabs = sigabs + kgg;
    }
    if (sigabs < 1.e-40) sigabs = 1.e-40;
    const double xabs = sigabs * trld;
    const double att = (xabs < 100.) ? exp(-xabs) : 0.;
    const double ewnew = (xabs < 100.) ? P.ew * att : 0.;
    double deleabs;
    if (P.xnu > 47.0 && sigabs > 1.e-40 && p.pair_switch == 1) deleabs = (P.ew - ewnew) * (sigabs - kgg_i) / sigabs;
    else deleabs = P.ew - ewnew;
    if (deleabs < 1.e-50) deleabs = 1.e-50;
    if (MODE == MODE_PRE) {
        if (s >= kMaxSeg) return END_OVERFLOW;
        Seg &g = segs[s];
        g.zi = zi; g.sigabs = sigabs; g.L = trld; g.trldb = trldb_rec; g.dcen = P.dcen; g.ew = P.ew;
        g.dele = deleabs; g.rpre = P.rpre; g.sigsc = pair_enhance;    // sigsc := comac * pair_enhance later
        g.att = att;
        nseg = s + 1;
    } else {
        double wmustar = P.wmu;
        if (!(xabs <= .00001)) {
            if (MODE == MODE_F0) rng.d++;                               // draw consumed, value unused (no deposit)
            else {
                const double mb_ran = (MODE == MODE_TRIAL || MODE == MODE_TRIAL_SLOW) ? abs_uniform(p, c.abs_S, c.abs_t, c.abs_k++)
                                                                                      : rng.next(p);   // always < ew/deleabs (>= 1)
                const double sstar = -log(1. - mb_ran * deleabs / P.ew) / sigabs;
                const double denom = sqrt(P.rpre * P.rpre + 2. * P.wmu * P.rpre * sstar + sstar * sstar);
                wmustar = (P.wmu * P.rpre + sstar) / denom;
            }
        }
        if (MODE != MODE_F0) {
            atomic_add_f64(p.edep + zi, deleabs);
            atomic_add_f64(p.prdep + zi, deleabs * wmustar * kCLight);
        }
    }
    s++;
    if (ewnew <= wtmin) { if (MODE != MODE_PRE) count(p, CNT_KILLED); return END_KILLED; }
    P.ew = ewnew;
    P.dcen = P.dcen - trld;
    Eta = (f + Eta * P.rpre) / rnew;
    if (Eta > 1.0) Eta = 1.0;
    if (Eta < -1.0) Eta = -1.0;
    P.phi = acos(Eta);
    if (Eta_switch == -1) P.phi = 2.0 * kPi - P.phi;
    P.rpre = rnew;
    P.zpre = znew;
    if (ikind == 1) {
        colmfp = colmfp - sigsc * trld;
        if (jnew == p.nz + 1 || jnew == 0 || knew == p.nr + 1 || knew == 0) {
            P.jph = jnew; P.kph = knew;
            if (MODE == MODE_PRE) return END_LEAK;                 // the ray leaves the grid here (nothing tallied)
            if (MODE == MODE_TRIAL || MODE == MODE_TRIAL_SLOW) return END_DROPPED;
            const int idead = imcleak(p, P, rng);
            count(p, CNT_LKCOUNT);
            if (idead == 1) return END_LEAK;
            return STEP_CONT;
        }
        P.kph = knew; P.jph = jnew;
        return STEP_CONT;
    }
    if (ikind == 2) {
        if (MODE == MODE_PRE) return END_CENSUS;                   // the ray ends at census (nothing tallied)
        if (MODE == MODE_TRIAL || MODE == MODE_TRIAL_SLOW) return END_DROPPED;
        return census_write(p, P, rng) ? END_CENSUS : END_FATAL;
    }
    return END_COLLISION;                                       // ikind == 3
}

// a whole flight by one thread
template <int MODE>
__device__ __noinline__ int fly(const Params &p, Photon &P, Rng &rng, double wtmin, double colmfp, Seg *segs, int &nseg,
                                uint64_t abs_S = 0, uint32_t abs_t = 0)
{
    FlightCtx c;
    flight_begin(p, P.xnu, c);
    c.abs_S = abs_S; c.abs_t = abs_t; c.abs_k = 0;
    if ((MODE == MODE_TRIAL_SLOW || MODE == MODE_CHILD) && p.sigma_mode == 1) sig_begin(p, P.xnu, c.sig);
    int s = 0;
    for (;;) {                                                    // label 110
        if (P.ew < 1.e-40) return END_LOWEW;
        if (P.wmu > 1.0) P.wmu = 1.0;
        if (P.wmu < -1.0) P.wmu = -1.0;
        double sigsc = 0.0;
        if (MODE == MODE_TRIAL) sigsc = segs[s].sigsc;
        else if (MODE == MODE_TRIAL_SLOW || MODE == MODE_CHILD) {
            const int zi = (P.jph - 1) * p.nr + (P.kph - 1);
            const double pair_enhance = (p.pair_switch == 1) ? 1.0 + 2.0 * p.f_pair[zi] : 1.0;
            sigsc = 1.0 * comtot_one(p, zi, P.xnu, c.sig) * pair_enhance;
        }
        const double xnu_was = P.xnu;
        const int r = fly_step<MODE>(p, P, rng, wtmin, colmfp, sigsc, c, segs, s, nseg);
        if (r != STEP_CONT) return r;
        if (P.xnu != xnu_was) {                                   // Compton reflection at the lower surface changed the energy
            flight_begin(p, P.xnu, c);
            if ((MODE == MODE_TRIAL_SLOW || MODE == MODE_CHILD) && p.sigma_mode == 1) sig_begin(p, P.xnu, c.sig);
        }
    }
}

// ------------------------------------------------------------------------------------------ sources
#define C2D_F4(x) ((double)(x##f))

// imcvol2d_para.f:157-389: photon number ivol (1-based) of zone `zone` (1-based)
__device__ __forceinline__ void source_volume(const Params &p, Rng &rng, int zone, Photon &P)
{
    const int jv = (zone - 1) / p.nr + 1, kv = zone - p.nr * (jv - 1), zi = zone - 1;
    const double rlo = (kv == 1) ? p.rmin : p.r[kv - 2], rhi = p.r[kv - 1];
    const double zhi = p.z[jv - 1], zlo_tab = (jv == 1) ? 0.0 : p.z[jv - 2];
    const double f_thermal = p.Eloss_th[zi] / p.Eloss_tot[zi];
    const double delz = (jv == 1) ? p.z[0] : zhi - zlo_tab;
    const double zs = p.zsurf[zi];
    const double f_inn = (4.4e1 / 7.0 * rlo * delz) / zs;
    double f_outer = (4.4e1 / 7.0 * rhi * delz) / zs;
    double f_upper = (2.2e1 / 7.0 * (rhi * rhi - rlo * rlo)) / zs;
    f_outer = f_inn + f_outer;
    f_upper = f_outer + f_upper;
    P.jph = jv; P.kph = kv;
    P.ew = p.ewsv[zi];
    P.dcen = kCLight * p.dt * rng.next(p);
    double rnum = rng.next(p);
    const bool thermal = (rnum < f_thermal);
    const double *eps = (thermal ? p.eps_th : p.eps_tot) + (size_t)zi * p.n_vol;
    rnum = rng.next(p);
    {   // first i (1-based) with eps(i) >= rnum, capped at n_vol (the CDF is non-decreasing)
        int lo = 1, hi = p.n_vol;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (eps[mid - 1] < rnum) lo = mid + 1; else hi = mid;
        }
        const int i = lo;
        if (i < p.n_vol) P.xnu = p.E_ph[i - 1] + rng.next(p) * (p.E_ph[i] - p.E_ph[i - 1]);
        else P.xnu = p.E_ph[i - 1];
    }
    if (thermal) {
        const double rnum0 = rng.next(p);
        if (rnum0 < f_inn) {
            P.wmu = 2.0 * rng.next(p) - 1.0;
            C2D_CLAMP8(P.wmu);
            const double x1 = rng.next(p), x2 = rng.next(p);
            if (x1 < 0.5) {
                P.phi = 1.1e1 / 7.0 + (1.1e1 / 7.0) * x2;
                if (P.phi < 1.57079638e0) P.phi = 1.57079638e0;
            } else {
                P.phi = -1.1e1 / 7.0 - (1.1e1 / 7.0) * x2;
                if (P.phi > -1.57079638e0) P.phi = -1.57079638e0;
            }
            P.rpre = C2D_F4(1.00001) * rlo;
            if (jv == 1) P.zpre = p.z[0] * rng.next(p);
            else P.zpre = zlo_tab + rng.next(p) * (zhi - zlo_tab);
        } else if (rnum0 < f_outer) {
            P.wmu = 2.0 * rng.next(p) - 1.0;
            C2D_CLAMP8(P.wmu);
            P.rpre = C2D_F4(0.999999) * rhi;
            if (jv == 1) P.zpre = p.z[0] * rng.next(p);
            else P.zpre = zlo_tab + rng.next(p) * (zhi - zlo_tab);
            P.phi = -1.1e1 / 7.0 + 2.2e1 / 7.0 * rng.next(p);
            if (P.phi < -1.5707963e0) P.phi = -1.57079063e0;
            if (P.phi > 1.5707963e0) P.phi = 1.5707963e0;
        } else if (rnum0 < f_upper) {
            P.wmu = rng.next(p);
            if (P.wmu > 9.9999999e-1) P.wmu = 9.9999999e-1;
            if (P.wmu < 0.0) P.wmu = 0.0;
            P.phi = 4.4e1 / 7.0 * rng.next(p);
            if (P.phi > 2.0 * kPi) P.phi = 2.0 * kPi;
            const double psi = rng.next(p);
            P.zpre = C2D_F4(0.999999) * zhi;
            P.rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        } else {
            P.wmu = -rng.next(p);
            P.phi = 4.4e1 / 7.0 * rng.next(p);
            if (P.wmu > 0.0) P.wmu = 0.0;
            if (P.wmu < -9.9999999e-1) P.wmu = -9.9999999e-1;
            if (P.phi > 2.0 * kPi) P.phi = 2.0 * kPi;
            const double zl = (jv == 1) ? p.zmin : zlo_tab;
            P.zpre = C2D_F4(1.000001) * zl;
            if (P.zpre <= zl) P.zpre = zl + 1.e-6;
            const double psi = rng.next(p);
            P.rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        }
    } else {
        P.wmu = 2.0 * rng.next(p) - 1.0;
        P.phi = 4.4e1 / 7.0 * rng.next(p);
        C2D_CLAMP8(P.wmu);
        if (P.phi > 2.0 * kPi) P.phi = 2.0 * kPi;
        if (jv == 1) P.zpre = p.z[0] * rng.next(p);
        else P.zpre = zlo_tab + rng.next(p) * (zhi - zlo_tab);
        const double psi = rng.next(p);
        P.rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
    }
    P.jgpsp = sp_group(p, P.xnu, C2D_F4(1.000001), C2D_F4(.999999), p.nphtotal);
    P.jgplc = lc_group(p, P.xnu);
    P.jgpmu = mu_group(p, P.wmu);
}

#define C2D_CLAMP10(v) do { if ((v) > 0.9999999999e0) (v) = 0.9999999999e0; if ((v) < -0.9999999999e0) (v) = -0.9999999999e0; } while (0)

// planck2d.f:1-141 (REAL*4 literals at l.72-73); P.wmu must be set
__device__ __forceinline__ void planck(const Params &p, Rng &rng, double tpl, Photon &P)
{
    double u4;
    do {
        u4 = rng.next(p);
        u4 = u4 * rng.next(p);
        u4 = u4 * rng.next(p);
        u4 = u4 * rng.next(p);
    } while (u4 <= 1.e-200);
    const double ap0 = -log(u4);
    double ap1 = 1.0, ap2 = 1.0, ap3 = 1.0;
    const double rn1 = 1.08232e0 * rng.next(p);
    while (!(rn1 <= ap1)) {
        ap2 = ap2 + 1.0;
        ap3 = 1.0 / ap2;
        ap1 = ap1 + (ap3 * ap3) * (ap3 * ap3);
    }
    P.xnu = ap0 * ap3 * tpl;
    P.jgpsp = sp_group(p, P.xnu, C2D_F4(1.000001), C2D_F4(.999999), 0);
    P.jgplc = lc_group(p, P.xnu);
    P.jgpmu = mu_group(p, P.wmu);
}

// imcsurf2d_para.f:694-788; P.wmu must be set.  The CDF scan is a bisection (P_file is non-decreasing): first
// i in [1, nfile-1] with P_file(i) > x1, else nfile.
__device__ __forceinline__ void file_sample(const Params &p, Rng &rng, int spec, Photon &P)
{
    const size_t o = (size_t)spec * p.sp_nfmax;
    const int nfile = p.sp_n[spec];
    const double x1 = rng.next(p);
    int lo = 1, hi = nfile;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p.sp_P[o + mid - 1] > x1) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    const double x2 = rng.next(p);
    const double E = p.sp_E[o + i - 1], A = p.sp_a1[o + i - 1];
    P.xnu = E * pow(A * p.sp_I[o + i - 1] * x2 / (p.sp_F[o + i - 1] * E) + 1.0, 1. / A);
    P.jgpsp = sp_group(p, P.xnu, 1.000001e0, .999999e0, 0);
    P.jgplc = lc_group(p, P.xnu);
    P.jgpmu = mu_group(p, P.wmu);
}

// imcsurf2d_para.f:228-534: photon of surface slot `slot` (see Params::surf_prefix)
__device__ __forceinline__ void source_surface(const Params &p, Rng &rng, int slot, Photon &P)
{
    const double tbb = p.tbb_surf[slot];
    const int spec = p.spec_surf[slot];
    P.ew = p.ewsurf[slot];
    if (slot < 2 * p.nz) {                          // z_surf_calc(js)
        const int js = (slot >> 1) + 1;
        const double zlo = (js == 1) ? 0.0 : p.z[js - 2], zhi = p.z[js - 1];
        P.jph = js;
        P.wmu = 2.0 * rng.next(p) - 1.0;
        C2D_CLAMP10(P.wmu);
        if ((slot & 1) == 0) {                      // inner cylinder, l.255-289
            P.phi = -1.1e1 / 7.0 + 2.2e1 / 7.0 * rng.next(p);
            if (P.phi < -1.5707963e0) P.phi = -1.57079063e0;
            if (P.phi > 1.5707963e0) P.phi = 1.5707963e0;
            P.rpre = p.rmin;
            if (js == 1) P.zpre = p.z[0] * rng.next(p);
            else P.zpre = zlo + rng.next(p) * (zhi - zlo);
            P.kph = 1;
        } else {                                    // outer cylinder, l.296-337
            P.rpre = p.r[p.nr - 1];
            if (js == 1) P.zpre = p.z[0] * rng.next(p);
            else P.zpre = zlo + rng.next(p) * (zhi - zlo);
            const double x1 = rng.next(p), x2 = rng.next(p);
            if (x1 < 0.5) {
                P.phi = 1.1e1 / 7.0 + (1.1e1 / 7.0) * x2;
                if (P.phi < 1.57079638e0) P.phi = 1.57079638e0;
            } else {
                P.phi = -1.1e1 / 7.0 - (1.1e1 / 7.0) * x2;
                if (P.phi > -1.57079638e0) P.phi = -1.57079638e0;
            }
            P.kph = p.nr;
        }
        P.dcen = rng.next(p) * kCLight * p.dt;
        if (tbb > 0.0) planck(p, rng, tbb, P); else file_sample(p, rng, spec, P);
        return;
    }
    const int q = slot - 2 * p.nz, ks = (q >> 1) + 1;   // r_surf_calc(ks)
    const double rlo = (ks == 1) ? p.rmin : p.r[ks - 2], rhi = p.r[ks - 1];
    P.kph = ks;
    if ((q & 1) == 0) {                             // upper disk, l.376-404
        P.wmu = -rng.next(p);
        C2D_CLAMP10(P.wmu);
        P.phi = 2.0 * kPi * rng.next(p);
        const double psi = rng.next(p);
        P.zpre = p.z[p.nz - 1];
        P.rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        P.dcen = rng.next(p) * kCLight * p.dt;
        if (tbb > 0.0) planck(p, rng, tbb, P); else file_sample(p, rng, spec, P);
        P.jph = p.nz;
    } else {                                        // lower disk, beamed upwards, l.444-490
        P.wmu = 9.9999999e-1;
        P.phi = 2.0 * kPi * rng.next(p);
        const double psi = rng.next(p);
        P.zpre = p.zmin;
        P.rpre = sqrt((rlo * rlo) + psi * (rhi * rhi - rlo * rlo));
        if (tbb > 0.e-20) planck(p, rng, tbb, P); else file_sample(p, rng, spec, P);
        P.dcen = rng.next(p) * kCLight * p.dt;      // drawn after the energy on this face (l.484)
        P.jph = 1;
    }
}

}  // namespace c2d
