// c2d_fp.cu -- batched per-zone electron Fokker-Planck update on sm_100a (one warp per zone).
//
// Reference parity (file:line relative to /root/reference/src_20121026/):
//   k_fp_zones      update2d.f:337-1741  FP_calc: implicit Chang-Cooper sub-steps until t_fp >= dt,
//                                        thermal rates -> sub-step length, injection, escape, tridag,
//                                        temperature from <gamma>, power-law re-fit
//   warp_gamma_bar  volume2d.f:572-626   gamma_bar / McDonald quadrature on the t = 1.001^n lattice
//   k_ic_loss       icloss2d.f:1-125     F_IC single-electron inverse-Compton loss table
// Only the terms that reach the solve are evaluated: dgdt = dg_sy + dg_ic + dg_A, disp = disp_A
// (update2d.f:1048-1049).  The Moeller / Coulomb / bremsstrahlung / Landau-damping rates that the
// reference computes, writes to rates/ and then discards are not.
//
// A warp owns a zone: the bins of every O(N) loop are strided over the 32 lanes, the two sequential
// recurrences (Thomas algorithm with the reference's clamp-after-use back-substitution, running CDF)
// are run by lane 0 out of shared memory, the McDonald quadratures and the 200-step p_nth scan are
// lane-parallel reductions.  The 1.005-lattice temperature walk of update2d.f:1446-1463 (up to ~1500
// gamma_bar calls per sub-step in the reference) is replaced by a bracketing search over the same
// lattice, which returns the same lattice point because gamma_bar is monotone on it.
#include "c2d_ctx.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace c2d {

#ifndef C2D_FP_FAT
#define C2D_FP_LEAN 1                  // 4 shared-memory arrays per zone instead of 9 (see k_fp_zones)
#endif
struct FpParams {
    int nz, nr, N, nf, pair_switch;
    int first_zone, n_zones, zone_stride;
    int inj_switch, inj_dis, g2var_switch, pick_sw, cf_sentinel;
    double time, dt, lnL, rmin, zmin, z_nz, r_nr;
    double inj_g1, inj_g2, inj_p, inj_t, inj_gg, inj_sigma, inj_L, inj_v, pick_rate, r_esc, r_acc;
    double r_flare, z_flare, t_flare, sigma_r, sigma_z, sigma_t, flare_amp;
    const double *z, *r, *gnt, *F_IC;
    const double *F_ICt;           // F_IC transposed, [nf][N]: lanes (= electron bins) read consecutive addresses
    const double *qt_x15;          // McDonald quadrature lattice: (ts_n^2 - 1)^(3/2), [kQuadRounds * 32]
    const double *gtab;            // gamma_bar on a uniform grid in ln(Theta): the predictor of the temperature search
    int gtab_n;
    double gtab_x0, gtab_ih;       // first node, 1 / spacing
    unsigned long long *gc_key;    // memo of the quadrature: gamma_bar is a pure function of the bits of Theta, and the zones of
    double *gc_val;                // a run walk the SAME temperature lattice (c2d_fp.cu:gamma_bar_memo); kGcSlots entries
    // per zone, in
    const double *f_nt_in, *Pnt_in, *n_field, *ecens, *n_e_in, *B_field, *tea, *tna, *vol, *f_pair_in, *Eloss_sy,
        *ec_old, *turb_lev;
    // per zone, out
    double *f_nt, *Pnt, *n_e, *Te_new, *gmin, *gmax, *p_nth, *amxwl, *f_pair;
    int *fp_steps;
    double *zE_old, *zE_new, *z_hr, *z_hrst, *z_dT, *z_misc;   // z_misc: [zone][3] = dr, dz, f_t_implicit
    double *scratch;                                           // C2D_FP_LEAN: [zone slot][2 N] dg_ic | dgdt (L1 / L2 resident)
};

struct FpState {
    double *d_FIC = nullptr;      // [N][nf]
    bool have_FIC = false;
    double *d_io = nullptr;       // staging blob for per-zone inputs/outputs
    size_t io_cap = 0;
    double *d_gtab = nullptr;     // [kGtabN] gamma_bar(exp(x0 + i h)), built once by k_gbar_table
    double *d_FICt = nullptr;     // [nf][N]
    double *d_quad = nullptr;     // [kQuadRounds * 32]: (ts^2-1)^(3/2)
    unsigned long long *d_gc_key = nullptr;   // gamma_bar memo (kGcSlots keys / values), kept for the life of the context
    double *d_gc_val = nullptr;
    double *d_scr = nullptr;      // k_fp_zones' global scratch (C2D_FP_LEAN)
    size_t scr_cap = 0;
};
constexpr int kGcSlots = 1 << 16;
constexpr int kQuadRounds = 1024;                                 // 32768 lattice points t = 1.001^n: covers z >= 1e-12
constexpr int kGtabN = 640;
constexpr double kGtabX0 = -1.7, kGtabH = 0.02;                   // Theta = 0.183 ... 6.5e4

#ifndef C2D_EMU      // warp-cooperative device code: not part of the g++ debugging build (tests/emu)

// volume2d.f:649-668
__device__ __forceinline__ double gammln_dev(double xx)
{
    const double cof[6] = {76.18009172947146e0, -86.50532032941677e0, 24.01409824083091e0,
                           -1.231739572450155e0, .1208650973866179e-2, -.5395239384953e-5};
    const double stp = 2.5066282746310005e0;
    double x = xx, y = x, tmp = x + 5.5, ser = 1.000000000190015e0;
    tmp = (x + 0.5) * log(tmp) - tmp;
#pragma unroll
    for (int j = 0; j < 6; j++) { y = y + 1.0; ser = ser + cof[j] / y; }
    return tmp + log(stp * ser / x);
}

// volume2d.f:572-626: gamma_bar(Theta), both McDonald quadratures (nu = 2, 3) in one lane-parallel pass.
// Term n of the reference's loop is t_n = 1.001^n; lane l takes n = l, l+32, ...; each quadrature stops
// after the first term (in n order) with t >= 2 and sd <= 1e-8, exactly like the scalar loop.
// The Theta-independent factor (ts^2-1)^(3/2), ts = t_n (1 + 1.001)/2, comes from a table built once by k_quad_table
// (the first kQuadRounds rounds; later ones are computed), so a term costs one exp: sd = x^(k/2) exp(-z ts)  [the
// reference divides by exp(z ts): same value to 1 ulp].
struct QuadTab { const double *x15; };

__device__ __forceinline__ double quad_x15(double t, double s)
{
    const double ts = t * s, x = ts * ts - 1.0;
    return x * sqrt(x);
}

// force_quad: use the quadrature below Theta = 0.2 too (the predictor table must be smooth across the switch)
__device__ double warp_gamma_bar(double Theta, int lane, const QuadTab tab, bool force_quad = false)
{
    double g;
    if (Theta < (double)0.2f && !force_quad) {
        g = (1. + (double)4.375f * Theta + (double)7.383f * (Theta * Theta) + (double)3.384f * (Theta * Theta * Theta)) /
                (1. + (double)1.875f * Theta + (double).8203f * (Theta * Theta)) - Theta;
    } else {
        const double z = 1.0 / Theta, dt = 1.001, d = dt - 1.0, s = 5.e-1 * (1.0 + dt);
        double sum2 = 0.0, sum3 = 0.0;
        bool done2 = false, done3 = false;
        double t = 1.0, dt32 = 1.0;
        for (int i = 0; i < 32; i++) { if (i < lane) t = t * dt; dt32 = dt32 * dt; }
        // Four rounds (128 terms) per trip: the stop tests are gathered per lane and the warp votes once; only the trip
        // in which a quadrature ends is resolved round by round.  Per-lane summation order is unchanged.
        double pf[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pf[k] = tab.x15[k * 32 + lane];
        for (int round = 0; round < 4096 && !(done2 && done3); round += 4) {
            double tr[4], s2r[4], s3r[4];
            unsigned st2 = 0u, st3 = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double x15 = (round + k < kQuadRounds) ? pf[k] : quad_x15(t, s);
                if (round + 4 + k < kQuadRounds) pf[k] = tab.x15[(round + 4 + k) * 32 + lane];   // next trip's loads in flight
                const double ts = t * s, y = z * ts;
                double sd2 = 0.0, sd3 = 0.0;
                if (y < 2.25e2) {
                    const double e = exp(-y);
                    sd2 = x15 * e;                                // (ts^2 - 1)^(3/2) / exp(y)
                    sd3 = sd2 * (ts * ts - 1.0);                  // (ts^2 - 1)^(5/2) / exp(y)
                }
                tr[k] = t; s2r[k] = sd2; s3r[k] = sd3;
                if (!(t < 2.0 || sd2 > 1.e-8)) st2 |= 1u << k;
                if (!(t < 2.0 || sd3 > 1.e-8)) st3 |= 1u << k;
                t = t * dt32;
            }
            const bool any2 = __any_sync(0xffffffffu, st2 != 0u), any3 = __any_sync(0xffffffffu, st3 != 0u);
            if (!done2) {
                if (!any2) {
#pragma unroll
                    for (int k = 0; k < 4; k++) sum2 += d * tr[k] * s2r[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (done2) break;
                        const unsigned stop2 = __ballot_sync(0xffffffffu, (st2 >> k) & 1u);
                        const int last = stop2 ? (__ffs(stop2) - 1) : 31;
                        if (lane <= last) sum2 += d * tr[k] * s2r[k];
                        done2 = (stop2 != 0u);
                    }
                }
            }
            if (!done3) {
                if (!any3) {
#pragma unroll
                    for (int k = 0; k < 4; k++) sum3 += d * tr[k] * s3r[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (done3) break;
                        const unsigned stop3 = __ballot_sync(0xffffffffu, (st3 >> k) & 1u);
                        const int last = stop3 ? (__ffs(stop3) - 1) : 31;
                        if (lane <= last) sum3 += d * tr[k] * s3r[k];
                        done3 = (stop3 != 0u);
                    }
                }
            }
        }
        sum2 = warp_sum(sum2); sum3 = warp_sum(sum3);
        const double sp = sqrt(3.14159265e0);
        const double K2 = sp * sq(5.e-1 * z) * sum2 / exp(gammln_dev(2.5));
        const double K3 = sp * pow(5.e-1 * z, 3.0) * sum3 / exp(gammln_dev(3.5));
        g = K3 / K2 - Theta;
    }
    if (g < 1.0) g = 1.0;
    return g;
}

// gamma_bar through a memo.  The quadrature is a pure function of the BITS of Theta, the electron temperatures of a run
// live on a few lattices Theta_0 * 1.005^n (update2d.f:1446-1463; every zone that hits the 5 / 1000 keV clamp of
// update2d.f:268-274 restarts from the same Theta_0), and the walk of one sub-step re-evaluates points its neighbours and
// its own previous sub-steps already evaluated: measured, the quadrature was 40 % of the kernel's instructions.  A hit
// returns the value the quadrature itself produced for exactly this argument -- bit-identical, no tolerance involved.
// Open addressing in global memory; a slot whose value has not landed yet reads as a miss.
__device__ double gamma_bar_memo(const FpParams &q, double Theta, int lane, const QuadTab tab)
{
    if (Theta < (double)0.2f || !q.gc_key) return warp_gamma_bar(Theta, lane, tab);
    const unsigned long long key = (unsigned long long)__double_as_longlong(Theta);
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    unsigned slot = (unsigned)h & (kGcSlots - 1);
    double g = 0.0;
    int found = -1;                                               // lane 0 probes, the warp follows
    if (lane == 0) {
        for (int k = 0; k < 8; k++) {
            const unsigned long long kk = *((volatile unsigned long long *)(q.gc_key + slot));
            if (kk == key) { g = *((volatile double *)(q.gc_val + slot)); found = (g != 0.0) ? 1 : 0; break; }
            if (kk == 0ull) { found = -1 - (int)k; break; }
            slot = (slot + 1) & (kGcSlots - 1);
            found = -100;                                         // probe sequence exhausted: compute, do not insert
        }
    }
    found = __shfl_sync(0xffffffffu, found, 0);
    if (found == 1) return __shfl_sync(0xffffffffu, g, 0);
    g = warp_gamma_bar(Theta, lane, tab);
    if (lane == 0 && found < 0 && found > -100) {
        if (atomicCAS(q.gc_key + slot, 0ull, key) == 0ull) *((volatile double *)(q.gc_val + slot)) = g;
    }
    return g;
}

// the quadrature lattice: lane l of round r holds t = 1.001^l (l sequential products) times (1.001^32)^r (r products)
__global__ void k_quad_table(double *x15_out, int rounds)
{
    const int lane = threadIdx.x;
    const double dt = 1.001, s = 5.e-1 * (1.0 + dt);
    double t = 1.0, dt32 = 1.0;
    for (int i = 0; i < 32; i++) { if (i < lane) t = t * dt; dt32 = dt32 * dt; }
    for (int r = 0; r < rounds; r++) {
        x15_out[r * 32 + lane] = quad_x15(t, s);
        t = t * dt32;
    }
}

// F_IC [N][nf] -> [nf][N]
__global__ void k_transpose(const double *in, int rows, int cols, double *out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int r = idx / cols, c = idx - r * cols;
    out[(size_t)c * rows + r] = in[idx];
}

// Theta_0 * c^n (n >= 0) or Theta_0 / c^|n| (n < 0) by the reference's sequential products (update2d.f:1449,1456)
__device__ __forceinline__ double theta_lattice(double Theta0, int n)
{
    const double c = (double)1.005f;
    double t = Theta0;
    if (n >= 0) for (int i = 0; i < n; i++) t = t * c;
    else for (int i = 0; i < -n; i++) t = t / c;
    return t;
}

// The same lattice with check points: lane L keeps Theta_0 * c^(32 L) (the same sequential products), so a point costs a
// shuffle and at most 31 multiplications instead of n.  Built once per search; covers |n| < 1024, beyond it falls back.
struct ThetaLattice {
    double Theta0, ckpt;
    int dir;
    __device__ __forceinline__ void build(double t0, int d, int lane)
    {
        const double c = (double)1.005f;
        Theta0 = t0; dir = d;
        double t = t0;
        ckpt = t0;
        if (d < 0) return;                                        // cooling walks are short (and a division per step): direct
        for (int i = 1; i <= 32 * 31; i++) {
            t = t * c;
            if (i == 32 * lane) ckpt = t;
        }
    }
    __device__ __forceinline__ double at(int n) const          // n >= 0 steps in direction dir; every lane calls
    {
        if (n >= 1024 || dir < 0) return theta_lattice(Theta0, dir * n);
        const double c = (double)1.005f;
        double t = __shfl_sync(0xffffffffu, ckpt, n >> 5);
        for (int i = 0; i < (n & 31); i++) t = t * c;
        return t;
    }
};

// First n >= 1 on the temperature lattice at which the reference's walk stops (update2d.f:1446-1463):
//   dir = +1: Theta_n = Theta_0 * 1.005^n, stop when !(gbar > gamma_bar(Theta_n))
//   dir = -1: Theta_n = Theta_0 / 1.005^n, stop when !(gbar < gamma_bar(Theta_n)) or Theta_n < 1e-2
// gamma_bar is monotone on the lattice, so "stops at n" is monotone in n and a bracketing search started
// at `guess` returns the same n as the step-by-step walk.
__device__ int lattice_search(const FpParams &q, const ThetaLattice &L, int dir, double gbar, int guess, int lane, const QuadTab tab, double &g_out)
{
    auto stops = [&](int n, double &g) {
        const double th = L.at(n);
        g = gamma_bar_memo(q, th, lane, tab);
        return (dir > 0) ? !(gbar > g) : (!(gbar < g) || th < 1.e-2);
    };
    int lo = 0, hi = -1, n = max(1, guess), s = 1;
    double g, g_hi = 0.0;
    if (stops(n, g)) {                     // bracket downwards
        hi = n; g_hi = g;
        while (hi - lo > 1) {
            n = max(lo + 1, hi - s); s *= 2;
            if (stops(n, g)) { hi = n; g_hi = g; } else { lo = n; break; }
        }
    } else {                               // bracket upwards
        lo = n;
        for (;;) {
            n = lo + s; s *= 2;
            if (stops(n, g)) { hi = n; g_hi = g; break; }
            lo = n;
        }
    }
    while (hi - lo > 1) {                  // bisect
        n = lo + (hi - lo) / 2;
        if (stops(n, g)) { hi = n; g_hi = g; } else lo = n;
    }
    g_out = g_hi;
    return hi;
}

// gamma_bar from the table (4-point Lagrange in ln Theta; measured against the quadrature: 3e-9 typical, 5e-8 worst,
// tests/test_gpu_fp.py): a PREDICTOR -- the temperature search confirms its decisions with the exact quadrature unless the
// predicted value clears gbar by kGtabMargin.  Below 0.2 the reference's own rational fit (exact).
// in_table = false beyond the last node (K3/K2 - Theta -> 3 Theta): never trusted without confirmation.
constexpr double kGtabMargin = 1.e-5;
__device__ double gamma_bar_approx(const FpParams &q, double Theta, bool &in_table)
{
    double g;
    in_table = true;
    if (Theta < (double)0.2f) {
        g = (1. + (double)4.375f * Theta + (double)7.383f * (Theta * Theta) + (double)3.384f * (Theta * Theta * Theta)) /
                (1. + (double)1.875f * Theta + (double).8203f * (Theta * Theta)) - Theta;
    } else {
        const double u = (log(Theta) - q.gtab_x0) * q.gtab_ih;
        int i = (int)u;
        if (i < 1) i = 1;
        if (i > q.gtab_n - 3) { in_table = false; return 3.0 * Theta; }
        const double t = u - (double)i;                           // nodes i-1, i, i+1, i+2 at t = -1, 0, 1, 2
        const double *y = q.gtab + i;
        g = -t * (t - 1.0) * (t - 2.0) / 6.0 * y[-1] + (t + 1.0) * (t - 1.0) * (t - 2.0) / 2.0 * y[0]
            - (t + 1.0) * t * (t - 2.0) / 2.0 * y[1] + (t + 1.0) * t * (t - 1.0) / 6.0 * y[2];
    }
    if (g < 1.0) g = 1.0;
    return g;
}

// lattice_search with the predictor in place of the quadrature: the n the exact search will most likely return
__device__ int lattice_predict(const FpParams &q, const ThetaLattice &L, int dir, double gbar, int guess)
{
    auto stops = [&](int n) {
        const double th = L.at(n);
        bool in_table;
        const double g = gamma_bar_approx(q, th, in_table);
        return (dir > 0) ? !(gbar > g) : (!(gbar < g) || th < 1.e-2);
    };
    int lo = 0, hi, n = max(1, guess), s = 1;
    if (stops(n)) {
        hi = n;
        while (hi - lo > 1) {
            n = max(lo + 1, hi - s); s *= 2;
            if (stops(n)) hi = n; else { lo = n; break; }
        }
    } else {
        lo = n;
        for (;;) {
            n = lo + s; s *= 2;
            if (stops(n) || n > (1 << 20)) { hi = n; break; }
            lo = n;
        }
    }
    while (hi - lo > 1) {
        n = lo + (hi - lo) / 2;
        if (stops(n)) hi = n; else lo = n;
    }
    return hi;
}

// update2d.f:1441-1469: the new electron temperature.  Returns The_new; g_av_out = gamma_bar(The_new).
// The predicted lattice point n is confirmed by the exact quadrature: the walk stops at n and does not stop at n - 1 (the
// latter taken from the predictor when it is not a close call); if the confirmation fails the exact bracketing search takes
// over from n.
__device__ double temperature_search(const FpParams &q, double Th_e, double g_av, double gbar, int lane, double &g_av_out)
{
    g_av_out = g_av;
    if (!(gbar > g_av) && !(gbar < g_av)) return Th_e;
    const int dir = (gbar > g_av) ? +1 : -1;
    int guess = 1;
    if (dir > 0 && gbar > 4.0) {
        const double tstar = gbar / 2.985;                          // K3/K2 - Theta -> 3 Theta for Theta >> 1
        if (tstar > Th_e) guess = max(1, (int)(log(tstar / Th_e) / log((double)1.005f)));
    }
    ThetaLattice L;
    L.build(Th_e, dir, lane);
    int n = lattice_predict(q, L, dir, gbar, guess);
    const QuadTab tab{q.qt_x15};
    auto stops = [&](int m, double &g) {
        const double th = L.at(m);
        g = gamma_bar_memo(q, th, lane, tab);
        return (dir > 0) ? !(gbar > g) : (!(gbar < g) || th < 1.e-2);
    };
    // the walk must not stop at n - 1: confirmed by the quadrature unless the predictor clears gbar by kGtabMargin there
    bool sure_m = false;
    if (n > 1) {
        bool in_table;
        const double g_pred = gamma_bar_approx(q, L.at(n - 1), in_table);
        sure_m = in_table && fabs(g_pred - gbar) > kGtabMargin * gbar;
    }
    double g_n, g_m;
    bool ok = stops(n, g_n);                                      // the exact value at n is the next sub-step's g_av
    if (ok && n > 1 && !sure_m) ok = !stops(n - 1, g_m);
    if (ok) g_av_out = g_n;
    else n = lattice_search(q, L, dir, gbar, n, lane, tab, g_av_out);
    return L.at(n);
}

// one warp per node of the predictor table
__global__ void __launch_bounds__(128) k_gbar_table(double *tab, int n, double x0, double h, const QuadTab qt)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const double g = warp_gamma_bar(exp(x0 + (double)i * h), threadIdx.x & 31, qt, true);
    if ((threadIdx.x & 31) == 0) tab[i] = g;
}

// shared-memory layout per warp: 9 arrays of N doubles; C2D_FP_LEAN: 4 (f, a_i, b_i, c_i) -- f_new lives in f_old (the Thomas
// sweeps work in place), the interface quantities bigC / smw are recomputed where they are used (same expressions, same
// values), gam(i) of tridag goes where a_i(i) has just been consumed, dg_ic / dgdt sit in a global scratch row
// Measured on the B200 (4096 zones, ms per call): 9 arrays 1.258 (N = 200) / 5.465 (N = 512); 4 arrays at 128 registers
// 1.146 / 3.738, at 96 registers 1.205 / 3.910, at 72 registers (28 warps / SM: all 4096 zones of a 64 x 64 grid resident
// at once) 1.034 / 4.240 -- so the kernel is instantiated twice and the launcher picks by N.
#ifdef C2D_FP_LEAN
constexpr int kFpArrays = 4;
#else
constexpr int kFpArrays = 9;
#endif
template <int MINB>
__global__ void __launch_bounds__(64, MINB) k_fp_zones(const FpParams q)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = blockIdx.x * (blockDim.x >> 5) + w;
    if (slot >= q.n_zones) return;
    const int zi = q.first_zone + slot * q.zone_stride;
    const int N = q.N, nf = q.nf;
#ifdef C2D_FP_LEAN
    double *f_old = smem + (size_t)w * kFpArrays * N, *f_new = f_old, *a_i = f_old + N, *b_i = a_i + N, *c_i = b_i + N,
           *dg_ic = q.scratch + (size_t)slot * 2 * N, *dgdt = dg_ic + N;
#else
    double *f_old = smem + (size_t)w * kFpArrays * N, *f_new = f_old + N, *a_i = f_new + N, *b_i = a_i + N, *c_i = b_i + N,
           *dg_ic = c_i + N, *dgdt = dg_ic + N, *bigC = dgdt + N, *smw = bigC + N;
#endif
    // disp_A(i) = gamma^2/t_acc/2 is recomputed where needed; bigW = smw/(exp(smw)-1) likewise
    const double *gnt = q.gnt;
    const int j = zi / q.nr + 1, k = zi - q.nr * (j - 1) + 1;
    const double c_light = kCLight, pi = kPi;
    const double t_esc = q.r_esc * q.z_nz / c_light, t_acc = q.r_acc * q.z_nz / c_light;
    const double dt1 = q.dt;
    double t_fp = 0.0, d_t = 0.0;
    int fp_steps = 0;
    double Te_new = q.tea[zi];
    const double volume = q.vol[zi];
    double n_p = q.n_e_in[zi];
    double f_pair = q.f_pair_in ? q.f_pair_in[zi] : 0.0;
    double ne = n_p * (1. + f_pair), n_positron = n_p * f_pair, n_lept = ne + n_positron;
    double E_old = 0.0, E_new = 0.0, hr_tot = 0.0, hr_st_tot = 0.0, f_t_implicit = 0.0;
    const double *f_in = q.f_nt_in + (size_t)zi * N;
    if (n_lept < 1.e-11) {                                        // update2d.f:472: nothing is changed
        for (int i = lane; i < N; i += 32) { q.f_nt[(size_t)zi * N + i] = f_in[i]; q.Pnt[(size_t)zi * N + i] = q.Pnt_in[(size_t)zi * N + i]; }
        if (lane == 0) { q.n_e[zi] = n_p; q.Te_new[zi] = Te_new; q.fp_steps[zi] = 0; q.f_pair[zi] = f_pair;
                         q.zE_old[zi] = 0; q.zE_new[zi] = 0; q.z_hr[zi] = 0; q.z_hrst[zi] = 0; q.z_dT[zi] = 0;
                         q.z_misc[3 * zi] = 0; q.z_misc[3 * zi + 1] = 0; q.z_misc[3 * zi + 2] = 0; }
        return;
    }
    // E_el, sum_p (l.476-509)
    double E_el = 0.0, sum_p = 0.0;
    for (int i = lane; i < N; i += 32) {
        if (i > 0) E_el += (gnt[i] - gnt[i - 1]) * (gnt[i] + 1.0) * f_in[i];
        if (i < N - 1) sum_p += (gnt[i + 1] - gnt[i]) * f_in[i];
    }
    E_el = warp_sum(E_el); sum_p = warp_sum(sum_p);
    E_el = E_el * ne * 8.176e-7 * volume;
    E_old = E_old + E_el + q.ec_old[zi];
    E_new = E_new + q.ecens[zi];
    for (int i = lane; i < N; i += 32) f_old[i] = (i == N - 1) ? 0.0 : f_in[i] / sum_p;
    const double rmid = (k > 1) ? 5.e-1 * (q.r[k - 1] + q.r[k - 2]) : 5.e-1 * (q.r[k - 1] + q.rmin);
    const double zmid = (j > 1) ? 5.e-1 * (q.z[j - 1] + q.z[j - 2]) : 5.e-1 * (q.z[j - 1] + q.zmin);
    double tl_flare = 0.0;
    if (q.cf_sentinel == 1) {
        const double y = 5.e-1 * (sq((rmid - q.r_flare) / q.sigma_r) + sq((zmid - q.z_flare) / q.sigma_z) +
                                  sq((q.time - q.t_flare) / q.sigma_t));
        tl_flare = (y < 1.e2) ? q.flare_amp / exp(y) : 0.0;
    }
    const double tlev = q.turb_lev[zi] + tl_flare;
    const double Tp_flare = q.tna[zi] * (1.0 + tl_flare);
    const double Th_p = Tp_flare / 9.382e5;
    double Th_e = q.tea[zi] / 5.11e2;
    const double f_th = 1.5 * volume * n_lept;
    // dg_ic(i) = -sum_ph n_field(ph) F_IC(i,ph) / volume (l.568-574): one bin per lane, the transposed F_IC streamed from L2.
    // volume ~ 1e45 cm^3 is outside the exponent range of the inline FP64 division (every quotient would take the ~100-
    // instruction slow path), so numerator and denominator are scaled by the same power of two first -- exact, the
    // quotient is bit-identical -- and the empty bins of the photon field (a - 0 = a) are skipped.
    {
        const double *nfld = q.n_field + (size_t)zi * nf;
        const double sc = scalbn(1.0, -ilogb(volume)), vs = volume * sc;
        // scratch a_i | b_i | c_i (3 N doubles): the non-empty bins compacted -- scaled value and bin index
        double *nf_s = a_i;
        int *nf_i = reinterpret_cast<int *>(a_i + nf);
        const bool staged = 2 * nf + nf <= 2 * 3 * N;             // nf doubles + nf ints
        int n_live = 0;
        if (staged) {
            for (int base = 0; base < nf; base += 32) {
                const int ip = base + lane;
                const double a = (ip < nf) ? nfld[ip] * sc : 0.0;
                const unsigned m = __ballot_sync(0xffffffffu, a != 0.0);
                __syncwarp();
                if (a != 0.0) { const int pos = n_live + __popc(m & ((1u << lane) - 1u)); nf_s[pos] = a; nf_i[pos] = ip; }
                n_live += __popc(m);
            }
            __syncwarp();
        }
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            const bool live = i < N - 1;
            const double *F = q.F_ICt + (live ? i : 0);
            double v = 0.0;
            if (staged) {
#pragma unroll 4
                for (int jj = 0; jj < n_live; jj++) v = v - nf_s[jj] * F[(size_t)nf_i[jj] * N] / vs;
            } else {
                for (int ip = 0; ip < nf; ip++) {
                    const double a = nfld[ip] * sc;
                    if (a == 0.0) continue;                       // warp-uniform
                    v = v - a * F[(size_t)ip * N] / vs;
                }
            }
            if (i < N) dg_ic[i] = live ? v : 0.0;
        }
    }
    __syncwarp();
    const double B = q.B_field[zi];
    const double dr = (k == 1) ? q.r[k - 1] - q.rmin : q.r[k - 1] - q.r[k - 2];
    const double dz = (j == 1) ? q.z[j - 1] - q.zmin : q.z[j - 1] - q.z[j - 2];
    double sum_E = 0.0, sum_last = 1.0;
    double g_av = gamma_bar_memo(q, Th_e, lane, QuadTab{q.qt_x15});
    for (;;) {                                                    // label 200 (g_av = gamma_bar(Th_e), l.600)
        double hr_th_c = 0.0;
        for (int i = lane; i < N - 1; i += 32)
            hr_th_c = hr_th_c - 8.176e-7 * dg_ic[i] * f_old[i] * (gnt[i + 1] - gnt[i]) * volume * n_lept;
        hr_th_c = warp_sum(hr_th_c);
        const double gamma_R = 2.1e-3 * sqrt(n_lept) / (B * sqrt(g_av));
        const double h_T = (double).79788f * (2. * sq(Th_e + Th_p) + 2.0 * (Th_e + Th_p) + 1.0) /
                           (pow(Th_e + Th_p, 1.5) * (1.0 + 1.875 * Th_e + .8203 * sq(Th_e)));
        const double hr_th_Coul = f_th * 1.7386e-26 * n_p * q.lnL * h_T * (Tp_flare - Te_new);
        double y = gamma_R / g_av;
        const double hr_th_sy = (y < 100.0) ? -q.Eloss_sy[zi] / (dt1 * exp(y)) : 0.0;
        double hr_th_A = tlev * hr_th_Coul;
        if (hr_th_A < 1.e-20) hr_th_A = 1.e-20;
        const double hr_th_total = hr_th_sy + hr_th_c + hr_th_A;
        const double dT_total = 6.25e8 * dt1 * hr_th_total / f_th;
        f_t_implicit = 1.e-2 * Te_new / fabs(dT_total);               // df_implicit, general.pa:27
        if (f_t_implicit > 2.5e-1) f_t_implicit = 2.5e-1;              // df_T, general.pa:28
        const double f_sy = 1.058e-15 * sq(B) / 8.176e-7;
        const double g_thr = 1.0 + 4.0 * Th_e;
        double hr_st_A = 0.0, hr_nt_A = 0.0;
        for (int i = lane; i < N; i += 32) {                          // loops 300 + 350, live terms
            const double gam = gnt[i] + 1.0;
            y = gamma_R / gam;
            const double dg_sy = (y < 100.0) ? -f_sy * (sq(gam) - 1.0) / exp(y) : -1.e-50;
            const double dg_A = gam / t_acc;
            dgdt[i] = dg_sy + dg_ic[i] + dg_A;
            if (i < N - 1) {
                const double term = dg_A * f_old[i] * ((gnt[i + 1] + 1.0) - gam);
                hr_nt_A += term;
                if (gam > g_thr) hr_st_A += term;
            }
        }
        hr_st_A = warp_sum(hr_st_A) * 8.176e-7 * n_lept * volume;
        hr_nt_A = warp_sum(hr_nt_A) * 8.176e-7 * n_lept * volume;
        const double heat_total = hr_th_Coul + hr_nt_A;
        E_old = E_old + heat_total * f_t_implicit * dt1;
        if (fp_steps == 0) { hr_tot += heat_total; hr_st_tot += hr_st_A; }
        d_t = f_t_implicit * dt1;
        if (d_t > (dt1 - t_fp)) d_t = 1.00001 * (dt1 - t_fp);
        // pairs are inert: only the flush of l.1187 survives when pair_switch = 1
        if (q.pair_switch == 1) {
            for (int i = lane; i < N - 1; i += 32) if (f_old[i] < 1.e-50) f_old[i] = 0.0;
        }
        n_positron = 0.0;
        ne = n_p + n_positron;
        f_pair = 0.0;
        // injection (l.1215-1306)
        double n_inject = 0.0;
        if (q.pick_sw == 1) {
            double inj_sum = 0.0;
            for (int i = lane; i < N - 1; i += 32) {
                const double gam = gnt[i] + 1.0;
                const double v = 1.e2 * exp(-sq(gam - q.inj_gg) / 2 / sq(q.inj_sigma)) / (q.inj_sigma * sqrt(2 * pi));
                a_i[i] = v;
                inj_sum += v * (gnt[i + 1] - gnt[i]);
            }
            inj_sum = warp_sum(inj_sum);
            const double inj_rho = q.pick_rate * d_t;
            double ninj = 0.0;
            for (int i = lane; i < N - 1; i += 32) {
                const double v = inj_rho * a_i[i] / inj_sum;
                f_old[i] = f_old[i] + v / ne;
                ninj += v * (gnt[i + 1] - gnt[i]);
            }
            n_inject += warp_sum(ninj);
        }
        if (q.inj_switch != 0) {
            const double tt = q.time + t_fp - q.inj_t;
            if (tt > dz / q.inj_v * (j - 1) && tt < dz / q.inj_v * j && k <= q.nr) {
                double inj_sum = 0.0, inj_E = 0.0;
                const double inj_g2var = q.inj_g2 * pow(10.0, tt * q.inj_v / q.z_nz);
                for (int i = lane; i < N - 1; i += 32) {
                    const double gam = gnt[i] + 1.0;
                    double v = 0.0;
                    if (q.inj_dis == 1) {
                        v = 1.e2 * exp(-sq(gam - q.inj_gg) / 2 / sq(q.inj_sigma)) / (q.inj_sigma * sqrt(2 * pi));
                    } else if (q.inj_dis == 2) {
                        if (gam > q.inj_g1) {
                            const double inj_y = (q.g2var_switch == 1) ? gam / inj_g2var : gam / q.inj_g2;
                            if (inj_y < 1.e2) v = 1.e2 / (pow(gam, q.inj_p) * exp(inj_y));
                        }
                    }
                    a_i[i] = v;
                    inj_sum += v * (gnt[i + 1] - gnt[i]);
                    inj_E += v * (gnt[i + 1] - gnt[i]) * gam;
                }
                inj_sum = warp_sum(inj_sum); inj_E = warp_sum(inj_E);
                inj_E = inj_E / inj_sum;
                const double inj_rate = q.inj_L / 8.186e-7 / inj_E / (pi * (q.r_nr * q.r_nr) * dz);
                const double inj_rho = inj_rate * d_t;
                double ninj = 0.0;
                for (int i = lane; i < N - 1; i += 32) {
                    const double v = inj_rho * a_i[i] / inj_sum;
                    f_old[i] = f_old[i] + v / ne;
                    ninj += v * (gnt[i + 1] - gnt[i]);
                }
                n_inject += warp_sum(ninj);
            }
        }
        ne = ne + n_inject;
        n_p = n_p + n_inject;
        n_lept = n_lept + n_inject;
        // escape (l.1309-1313)
        ne = ne * t_esc / (t_esc + d_t);
        n_p = n_p * t_esc / (t_esc + d_t);
        n_lept = n_lept * t_esc / (t_esc + d_t);
        __syncwarp();
        // Chang-Cooper coefficients (l.1319-1391): interface quantities bigC, smw first
        // interface m (0-based) between bins m, m + 1: bigC, smw
        auto iface = [&](int m, double &C_m, double &w_m) {
            const double g0 = gnt[m] + 1.0, g1 = gnt[m + 1] + 1.0;
            const double disp0 = g0 * g0 / t_acc / 2.0, disp1 = g1 * g1 / t_acc / 2.0;
            const double bigB = (m == 0) ? -(dgdt[0] + dgdt[1]) : -(dgdt[m] + dgdt[m + 1]) / 2.0;
            C_m = (disp0 + disp1) / 2.0;
            w_m = (gnt[m + 1] - gnt[m]) * bigB / C_m;
        };
#ifndef C2D_FP_LEAN
        for (int i = lane; i < N - 1; i += 32) iface(i, bigC[i], smw[i]);
        __syncwarp();
#endif
        for (int i = lane; i < N; i += 32) {
            if (i == 0 || i == N - 1) { a_i[i] = 0.0; b_i[i] = 1.0; c_i[i] = 0.0; continue; }
            const double D_gminus = gnt[i] - gnt[i - 1], D_gplus = gnt[i + 1] - gnt[i];
            const double Delta_g = sqrt(gnt[i] / gnt[i - 1]) * D_gminus;
#ifdef C2D_FP_LEAN
            double C_i, w_i, C_m, w_m;
            iface(i, C_i, w_i); iface(i - 1, C_m, w_m);
#else
            const double C_i = bigC[i], w_i = smw[i], C_m = bigC[i - 1], w_m = smw[i - 1];
#endif
            const double bigW_i = w_i / (exp(w_i) - 1.0), bigW_m = w_m / (exp(w_m) - 1.0);
            c_i[i] = -d_t * (C_i * w_i / (1.0 - exp(-w_i)) / Delta_g / D_gplus);
            b_i[i] = 1.0 + d_t / Delta_g * (C_i * bigW_i / D_gplus +
                                            C_m * w_m / (1.0 - exp(-w_m)) / D_gminus) + d_t / t_esc;
            a_i[i] = -d_t / Delta_g * C_m * bigW_m / D_gminus;
        }
        __syncwarp();
        // tridag (update2d.f:2485-2528): sequential recurrences, lane 0; gam(i) is kept in dgdt[]
        if (lane == 0) {
#ifdef C2D_FP_LEAN
            double *gam = a_i;                                    // a_i(i) is dead once row i has been eliminated
#else
            double *gam = dgdt;
#endif
            double bet = b_i[0];
            f_new[0] = f_old[0] / bet;
            for (int i = 1; i < N; i++) {
                const double gm = c_i[i - 1] / bet, ai = a_i[i];
                gam[i] = gm;
                bet = b_i[i] - ai * gm;
                f_new[i] = (f_old[i] - ai * f_new[i - 1]) / bet;
            }
            for (int i = N - 2; i >= 0; i--) {
                f_new[i] = f_new[i] - gam[i + 1] * f_new[i + 1];
                if (f_new[i + 1] < 0.0) f_new[i + 1] = 0.0;
            }
            f_new[N - 1] = 0.0;
            f_new[0] = 0.0;
            // running CDF (l.1413-1420) in bin order; c_i[] now holds the un-normalised Pnt
            double sp = 0.0, sE = 0.0;
            for (int i = 0; i < N - 1; i++) {
                sp = sp + (gnt[i + 1] - gnt[i]) * f_new[i];
                sE = sE + (gnt[i + 1] - gnt[i]) * (gnt[i] + 1.0) * f_new[i];
                c_i[i] = sp;
            }
            b_i[0] = sp; b_i[1] = sE;
        }
        __syncwarp();
        sum_last = b_i[0];
        sum_E = b_i[1] / sum_last;
        t_fp = t_fp + d_t;
        fp_steps = fp_steps + 1;
        double gbar = 0.0;
        for (int i = lane; i < N; i += 32) {
            const double v = f_new[i] / sum_last;
            f_new[i] = v; f_old[i] = v;
            if (i < N - 1) gbar += (gnt[i] + 1.0) * v * (gnt[i + 1] - gnt[i]);
        }
        gbar = warp_sum(gbar);
        __syncwarp();
        double g_new;
        const double The_new = temperature_search(q, Th_e, g_av, gbar, lane, g_new);
        Te_new = 5.11e2 * The_new;
        Th_e = The_new;
        g_av = g_new;                                             // == gamma_bar(Th_e) of the next sub-step
        if (!(t_fp < dt1) || fp_steps > 1000000) break;
    }
    // epilogue (l.1480-1500)
    double E_el2 = 0.0;
    const double Pn_in = q.Pnt_in[(size_t)zi * N + (N - 1)];
    for (int i = lane; i < N; i += 32) {
        q.f_nt[(size_t)zi * N + i] = f_new[i];
        q.Pnt[(size_t)zi * N + i] = ((i < N - 1) ? c_i[i] : Pn_in) / sum_last;
        if (i > 0) E_el2 += f_new[i] * (gnt[i] + 1.0) * (gnt[i] - gnt[i - 1]);
    }
    E_el2 = warp_sum(E_el2) * ne * 8.176e-7 * volume;
    E_new = E_new + E_el2;
    const double Delta_T = fabs(Te_new - q.tea[zi]) / Te_new;
    // power-law re-fit (l.1656-1738)
    int i_lo = N - 4, i_hi = 4;                                   // 1-based results of the two scans when nothing is found
    {
        int first = 1 << 30, last = -1;
        for (int i = lane; i < N; i += 32) {
            const int i1 = i + 1;
            if (i1 >= 5 && i1 <= N - 5) {
                if (f_new[i] > 1.e-10) first = min(first, i1);
                if (f_new[i] > 1.e-15) last = max(last, i1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        if (first != (1 << 30)) i_lo = first;
        if (last != -1) i_hi = last;
    }
    const double gmin = gnt[i_lo - 1] + 1.0, gmax = gnt[i_hi - 1] + 1.0;
    double sum_nt = 0.0, sum_th = 0.0;
    for (int i = lane; i < N - 1; i += 32) {
        const double v = ((gnt[i + 1] + 1.0) - (gnt[i] + 1.0)) * f_new[i];
        if (i + 1 < i_lo) sum_th += v; else sum_nt += v;
    }
    sum_th = warp_sum(sum_th); sum_nt = warp_sum(sum_nt);
    double amxwl = sum_th / (sum_nt + sum_th);
    double p_nth = q.p_nth[zi];
    if (amxwl > 9.999e-1) {
        amxwl = 1.0;
    } else {
        // update2d.f:1700-1738: scan p_nth = 0.1, 0.15, ... for the power law whose mean energy matches <gamma>.
        // gamma^-p is carried from step to step by one multiplication with gamma^-0.05 and exp(-gamma/gmax) is hoisted
        // (the reference calls gam**p_nth and exp(y) for every bin and step: same values to ~1e-14)
#ifdef C2D_FP_LEAN
        double *ipw = a_i, *ipstep = b_i, *wq = c_i;              // c_i (the running CDF) has been written out above
#else
        double *ipw = a_i, *ipstep = b_i, *wq = bigC;
#endif
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const int i1 = i + 1;
            const double gam = gnt[i] + 1.0, y = gam / gmax;
            const bool use = !(i1 < i_lo || i1 > N - 2) && y < 100.0;
            ipw[i] = use ? pow(gam, -(double)0.1f) : 0.0;
            ipstep[i] = use ? pow(gam, -0.5e-1) : 0.0;
            wq[i] = use ? (gnt[i + 1] - gnt[i]) * exp(-y) : 0.0;
        }
        __syncwarp();
        double sum_g = 1.e50, sumg_old;
        p_nth = (double)0.1f;
        // gmax^(1-p), gmin^(1-p) of the normalisation N_nt (which cancels in sg / sgg up to rounding) are carried the same way
        double pgx = pow(gmax, 1.0 - p_nth), pgn = pow(gmin, 1.0 - p_nth);
        const double pgx_step = pow(gmax, -0.5e-1), pgn_step = pow(gmin, -0.5e-1);
        for (;;) {
            sumg_old = sum_g;
            const double p_1 = 1.0 - p_nth;
            double N_nt;
            if (fabs(p_1) > 1.e-4) N_nt = (1. - amxwl) * p_1 / (pgx - pgn);
            else N_nt = (1.0 - amxwl) / log(gmax / gmin);
            double sg = 0.0, sgg = 0.0;
            for (int i = lane; i < N; i += 32) {
                const double pw = ipw[i];
                const double f_pl = N_nt * pw * wq[i];            // f_pl (gnt(i+1) - gnt(i))
                sg += f_pl * (gnt[i] + 1.0);
                sgg += f_pl;
                ipw[i] = pw * ipstep[i];
            }
            sg = warp_sum(sg); sgg = warp_sum(sgg);
            sum_g = fabs(sg / sgg - sum_E);
            if (sum_g < sumg_old) {
                if (p_nth < 10.) { p_nth = p_nth + 0.5e-1; pgx = pgx * pgx_step; pgn = pgn * pgn_step; continue; }
            }
            break;
        }
    }
    if (lane == 0) {
        q.n_e[zi] = n_p; q.Te_new[zi] = Te_new; q.fp_steps[zi] = fp_steps; q.f_pair[zi] = f_pair;
        q.gmin[zi] = gmin; q.gmax[zi] = gmax; q.amxwl[zi] = amxwl; q.p_nth[zi] = p_nth;
        q.zE_old[zi] = E_old; q.zE_new[zi] = E_new; q.z_hr[zi] = hr_tot; q.z_hrst[zi] = hr_st_tot; q.z_dT[zi] = Delta_T;
        q.z_misc[3 * zi] = dr; q.z_misc[3 * zi + 1] = dz; q.z_misc[3 * zi + 2] = f_t_implicit;
    }
}

// ------------------------------------------------------------------------------------------ IC_loss
__device__ __forceinline__ double f_Li_dev(double z)
{   // icloss2d.f:103-125
    double sum, sd = 0.0, nn = 1.0;
    const double y = 1.0 + 2.0 * z;
    sum = log(y) * (5.e-1 * log(y) - log(2.0 * z));
    do {
        if (nn < 1.5) sd = 1.0 / y;
        else sd = sd * (((nn - 1.0) / nn) * ((nn - 1.0) / nn)) / y;
        nn = nn + 1.0;
        sum = sum + sd;
    } while (fabs(sd / sum) > 1.e-10);
    return sum;
}
__device__ __forceinline__ double f1_dev(double z)
{   // icloss2d.f:68-82
    const double y = 1.0 + 2.0 * z;
    const double sd1 = (z + 6.0 + 3.0 / z) * log(y);
    const double sd2 = ((2.2e1 / 3.0) * (z * z * z) + 24.0 * (z * z) + 18.0 * z + 4.0) / (y * y);
    return sd1 - sd2 - 2.0 + 2.0 * f_Li_dev(z);
}
__device__ __forceinline__ double f2_dev(double z)
{   // icloss2d.f:85-100
    const double y = 1.0 + 2.0 * z;
    const double sd1 = (z + (3.1e1 / 6.0) + 5.0 / z + 1.5 / (z * z)) * log(y);
    const double sd2 = ((2.2e1 / 3.0) * (z * z * z) + 28.0 * (z * z) + (1.03e2 / 3.0) * z + 1.7e1 + 3.0 / z) / (y * y);
    return sd1 - sd2 - 2.0 + f_Li_dev(z);
}
__global__ void k_ic_loss(const double *gnt, const double *E_field, int N, int nf, double *F_IC)
{   // icloss2d.f:24-45
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * nf) return;
    const int ie = idx / nf, ip = idx - ie * nf;
    const double A = 3.7419e-15;
    const double gamma = 1.0 + gnt[ie], beta = sqrt(1.0 - 1.0 / (gamma * gamma));
    const double epsilon = 1.957e-3 * E_field[ip];
    double v;
    if ((gamma * epsilon) < 1.e-2) v = 2.66e-14 * epsilon * (gamma * gamma - 1.0);
    else {
        const double z1 = epsilon * gamma * (1.0 + beta), z2 = epsilon / (gamma * (1.0 + beta));
        const double F = gamma * (f1_dev(z1) - f1_dev(z2)) - epsilon * (f2_dev(z1) - f2_dev(z2));
        v = A * F / (((epsilon * gamma) * (epsilon * gamma)) * beta);
    }
    F_IC[idx] = v;
}

void launch_gbar_table(double *tab, double *quad, cudaStream_t st)
{
    k_quad_table<<<1, 32, 0, st>>>(quad, kQuadRounds);
    k_gbar_table<<<(kGtabN + 3) / 4, 128, 0, st>>>(tab, kGtabN, kGtabX0, kGtabH, QuadTab{quad});
}
// one warp per query: the quadrature and the predictor table side by side (c2d_gamma_bar)
__global__ void __launch_bounds__(128) k_gbar_query(const FpParams q, int n, const double *theta, double *exact, double *table)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const double g = warp_gamma_bar(theta[i], threadIdx.x & 31, QuadTab{q.qt_x15});
    bool in_table;
    const double ga = gamma_bar_approx(q, theta[i], in_table);
    if ((threadIdx.x & 31) == 0) { exact[i] = g; table[i] = in_table ? ga : -ga; }
}
void launch_gbar_query(const FpParams &q, int n, const double *theta, double *exact, double *table, cudaStream_t st)
{
    k_gbar_query<<<(n + 3) / 4, 128, 0, st>>>(q, n, theta, exact, table);
}
void launch_transpose(const double *in, int rows, int cols, double *out, cudaStream_t st)
{
    k_transpose<<<(rows * cols + 255) / 256, 256, 0, st>>>(in, rows, cols, out);
}
void launch_fp(const FpParams &q, cudaStream_t st)
{
    const int warps = 2;                                          // 6.4 kB of shared memory per warp at N = 200 (4 arrays)
    const size_t smem = (size_t)warps * kFpArrays * q.N * sizeof(double);
    const int blocks = (q.n_zones + warps - 1) / warps;
#ifdef C2D_FP_LEAN
    if (q.N <= 256) {                                             // 72 registers, 14 blocks / SM
        cudaFuncSetAttribute(k_fp_zones<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fp_zones<14><<<blocks, warps * 32, smem, st>>>(q);
        return;
    }
#endif
    cudaFuncSetAttribute(k_fp_zones<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fp_zones<7><<<blocks, warps * 32, smem, st>>>(q);
}
void launch_ic_loss(const double *gnt, const double *E_field, int N, int nf, double *F_IC, cudaStream_t st)
{
    k_ic_loss<<<(N * nf + 127) / 128, 128, 0, st>>>(gnt, E_field, N, nf, F_IC);
}
#else
void launch_fp(const FpParams &, cudaStream_t) {}
void launch_gbar_table(double *, double *, cudaStream_t) {}
void launch_transpose(const double *, int, int, double *, cudaStream_t) {}
void launch_gbar_query(const FpParams &, int, const double *, double *, double *, cudaStream_t) {}
void launch_ic_loss(const double *, const double *, int, int, double *, cudaStream_t) {}
#endif

void Ctx::fp_free()
{
    if (!fp) return;
    cudaFree(fp->d_FIC); cudaFree(fp->d_io); cudaFree(fp->d_gtab); cudaFree(fp->d_FICt); cudaFree(fp->d_quad);
    cudaFree(fp->d_gc_key); cudaFree(fp->d_gc_val); cudaFree(fp->d_scr);
    delete fp;
    fp = nullptr;
}

#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(C2D_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static int fp_state(Ctx *c)
{
    if (c->fp) return 0;
    c->fp = new FpState();
    CUF(cudaMalloc(&c->fp->d_FIC, (size_t)c->dims.num_nt * c->dims.nphfield * 8));
    CUF(cudaMalloc(&c->fp->d_FICt, (size_t)c->dims.num_nt * c->dims.nphfield * 8));
    CUF(cudaMalloc(&c->fp->d_gtab, (size_t)kGtabN * 8));
    CUF(cudaMalloc(&c->fp->d_quad, (size_t)kQuadRounds * 32 * 8));
    CUF(cudaMalloc(&c->fp->d_gc_key, (size_t)kGcSlots * 8));
    CUF(cudaMalloc(&c->fp->d_gc_val, (size_t)kGcSlots * 8));
    CUF(cudaMemsetAsync(c->fp->d_gc_key, 0, (size_t)kGcSlots * 8, c->stream));
    CUF(cudaMemsetAsync(c->fp->d_gc_val, 0, (size_t)kGcSlots * 8, c->stream));
    launch_gbar_table(c->fp->d_gtab, c->fp->d_quad, c->stream);
    CUF(cudaGetLastError());
    return 0;
}


// the broadcast scalars of FP_bcast (fp_mpi.f:579-600) + the static tables
static void fp_fill_scalars(Ctx *c, const c2d_fp_in *in, FpParams &q)
{
    const c2d_dims &d = c->dims;
    memset(&q, 0, sizeof q);
    q.nz = d.nz; q.nr = d.nr; q.N = d.num_nt; q.nf = d.nphfield; q.pair_switch = d.pair_switch;
    q.inj_switch = in->inj_switch; q.inj_dis = in->inj_dis; q.g2var_switch = in->g2var_switch; q.pick_sw = in->pick_sw;
    q.cf_sentinel = in->cf_sentinel;
    q.time = in->time; q.dt = in->dt; q.lnL = in->lnL; q.rmin = d.rmin; q.zmin = d.zmin;
    q.z_nz = c->h_z[d.nz - 1]; q.r_nr = c->h_r[d.nr - 1];
    q.inj_g1 = in->inj_g1; q.inj_g2 = in->inj_g2; q.inj_p = in->inj_p; q.inj_t = in->inj_t; q.inj_gg = in->inj_gg;
    q.inj_sigma = in->inj_sigma; q.inj_L = in->inj_L; q.inj_v = in->inj_v; q.pick_rate = in->pick_rate;
    q.r_esc = in->r_esc; q.r_acc = in->r_acc;
    q.r_flare = in->r_flare; q.z_flare = in->z_flare; q.t_flare = in->t_flare; q.sigma_r = in->sigma_r;
    q.sigma_z = in->sigma_z; q.sigma_t = in->sigma_t; q.flare_amp = in->flare_amp;
    q.z = c->P.z; q.r = c->P.r; q.gnt = c->P.gnt; q.F_IC = c->fp->d_FIC;
    q.F_ICt = c->fp->d_FICt;
    q.qt_x15 = c->fp->d_quad;
    q.gtab = c->fp->d_gtab; q.gtab_n = kGtabN; q.gtab_x0 = kGtabX0; q.gtab_ih = 1.0 / kGtabH;
    q.gc_key = getenv("C2D_FP_NO_MEMO") ? nullptr : c->fp->d_gc_key; q.gc_val = c->fp->d_gc_val;
}

// ---- the Fokker-Planck update on the resident plasma state (c2d_step.cu), zones dealt out rank::world, and the exchange of
// the results: FP_recv_result + FP_end_bcast (fp_mpi.f:926,1017-1079) as one all-gather of packed per-zone records
// [f_nt N | Pnt N | n_e Te_new gmin gmax p_nth amxwl f_pair E_old E_new hr hr_st dT steps dr dz f_t_implicit].
#ifndef C2D_EMU
__global__ void k_fp_pack(const FpParams q, double *send, int rec)
{
    const int slot = blockIdx.x;
    const int zi = q.first_zone + slot * q.zone_stride;
    double *o = send + (size_t)slot * rec;
    const int N = q.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { o[i] = q.f_nt[(size_t)zi * N + i]; o[N + i] = q.Pnt[(size_t)zi * N + i]; }
    if (threadIdx.x == 0) {
        double *e = o + 2 * N;
        e[0] = q.n_e[zi]; e[1] = q.Te_new[zi]; e[2] = q.gmin[zi]; e[3] = q.gmax[zi]; e[4] = q.p_nth[zi]; e[5] = q.amxwl[zi];
        e[6] = q.f_pair[zi]; e[7] = q.zE_old[zi]; e[8] = q.zE_new[zi]; e[9] = q.z_hr[zi]; e[10] = q.z_hrst[zi]; e[11] = q.z_dT[zi];
        e[12] = (double)q.fp_steps[zi]; e[13] = q.z_misc[3 * zi]; e[14] = q.z_misc[3 * zi + 1]; e[15] = q.z_misc[3 * zi + 2];
    }
}
// recv = [world][n_slot][rec]; slot s of rank w holds zone w + s world.  Writes the state every rank carries into the next
// step (update2d.f:262-276: tea = Te_new clamped to [5, 1000] keV where tna > 1) and the per-zone scalars zinfo[zone][16].
__global__ void k_fp_unpack(const double *recv, int world, int n_slot, int rec, int nzone, int N, double *f_nt, double *Pnt,
                            double *n_e, double *f_pair, double *tea, const double *tna, double *gmin, double *gmax, double *p_nth,
                            double *amxwl, double *Te_new, double *zinfo)
{
    const int zi = blockIdx.x;
    if (zi >= nzone) return;
    const int w = zi % world, s = zi / world;
    const double *o = recv + ((size_t)w * n_slot + s) * rec;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { f_nt[(size_t)zi * N + i] = o[i]; Pnt[(size_t)zi * N + i] = o[N + i]; }
    if (threadIdx.x < kFpRecExtra) zinfo[(size_t)zi * kFpRecExtra + threadIdx.x] = o[2 * N + threadIdx.x];
    if (threadIdx.x == 0) {
        const double *e = o + 2 * N;
        n_e[zi] = e[0]; Te_new[zi] = e[1]; gmin[zi] = e[2]; gmax[zi] = e[3]; p_nth[zi] = e[4]; amxwl[zi] = e[5]; f_pair[zi] = e[6];
        if (tna[zi] > 1.) {
            double t = e[1];
            t = fmin(1.e3, t);
            t = fmax(5.0, t);
            tea[zi] = t;
        }
    }
}
#endif

// global scratch rows of k_fp_zones (C2D_FP_LEAN): 2 N doubles per zone of the launch, grown to the largest launch seen
static int fp_scratch(Ctx *c, FpParams &q)
{
    q.scratch = nullptr;
#ifdef C2D_FP_LEAN
    const size_t need = (size_t)std::max(q.n_zones, 1) * 2 * (size_t)q.N;
    if (need > c->fp->scr_cap) {
        cudaFree(c->fp->d_scr); c->fp->d_scr = nullptr; c->fp->scr_cap = 0;
        CUF(cudaMalloc(&c->fp->d_scr, need * 8));
        c->fp->scr_cap = need;
    }
    q.scratch = c->fp->d_scr;
#else
    (void)c;
#endif
    return 0;
}

int fp_resident(Ctx *c, const c2d_fp_in *in, double *zinfo_host)
{
#ifdef C2D_EMU
    (void)c; (void)in; (void)zinfo_host;
    return fail(C2D_EINVAL, "the Fokker-Planck kernel needs the device build");
#else
    int rc = fp_state(c);
    if (rc) return rc;
    const c2d_dims &d = c->dims;
    const int N = d.num_nt, nf = d.nphfield, nzone = c->nzone;
    const int world = comm_world(c), rank = comm_rank(c);
    if (!c->fp->have_FIC) {
        launch_ic_loss(c->P.gnt, c->P.E_field, N, nf, c->fp->d_FIC, c->stream);
        launch_transpose(c->fp->d_FIC, N, nf, c->fp->d_FICt, c->stream);
        c->fp->have_FIC = true;
    }
    const int n_slot = (nzone + world - 1) / world, rec = 2 * N + kFpRecExtra;
    const int n_own = (rank < nzone) ? (nzone - rank + world - 1) / world : 0;
    // scratch: out arrays | send | recv | zinfo
    size_t n = 0;
    auto take = [&](size_t cnt) { size_t o = n; n += (cnt + 1) & ~size_t(1); return o; };
    const size_t p_f = take((size_t)nzone * N), p_P = take((size_t)nzone * N), p_ne = take(nzone), p_Te = take(nzone),
                 p_gmin = take(nzone), p_gmax = take(nzone), p_pn = take(nzone), p_am = take(nzone), p_fp = take(nzone),
                 p_Eo = take(nzone), p_En = take(nzone), p_hr = take(nzone), p_hs = take(nzone), p_dT = take(nzone),
                 p_misc = take((size_t)3 * nzone), p_steps = take(nzone), p_send = take((size_t)n_slot * rec),
                 p_recv = take((size_t)world * n_slot * rec), p_zi = take((size_t)nzone * kFpRecExtra);
    if (n > c->fpx_cap) {
        cudaFree(c->d_fpx); c->d_fpx = nullptr; c->fpx_cap = 0;
        CUF(cudaMalloc(&c->d_fpx, n * 8));
        c->fpx_cap = n;
    }
    double *D = c->d_fpx, *T = c->d_step, *S = c->d_plasma;
    const size_t nzp = ((size_t)nzone + 1) & ~size_t(1);
    double *s_tna = S + PL_TNA * nzp, *s_B = S + PL_B * nzp, *s_gmin = S + PL_GMIN * nzp, *s_gmax = S + PL_GMAX * nzp,
           *s_pn = S + PL_PNTH * nzp, *s_am = S + PL_AMXWL * nzp, *s_tl = S + PL_TURB * nzp, *s_vol = S + PL_VOL * nzp,
           *s_sy = S + PL_ELOSS_SY * nzp, *s_eco = S + PL_EC_OLD * nzp, *s_Te = S + PL_TE_NEW * nzp;
    // gmin / gmax / p_nth / amxwl are in/out of FP_calc (a zone whose amxwl -> 1 keeps its p_nth)
    CUF(cudaMemcpyAsync(D + p_gmin, s_gmin, (size_t)nzone * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUF(cudaMemcpyAsync(D + p_gmax, s_gmax, (size_t)nzone * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUF(cudaMemcpyAsync(D + p_pn, s_pn, (size_t)nzone * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUF(cudaMemcpyAsync(D + p_am, s_am, (size_t)nzone * 8, cudaMemcpyDeviceToDevice, c->stream));
    FpParams q;
    fp_fill_scalars(c, in, q);
    q.first_zone = rank; q.n_zones = n_own; q.zone_stride = world;
    q.f_nt_in = T + c->o_fnt; q.Pnt_in = T + c->o_pnt;
    q.n_field = c->P.n_field; q.ecens = c->P.ecens;                 // this step's tallies, already reduced over the ranks
    q.n_e_in = T + c->o_ne; q.B_field = s_B; q.tea = T + c->o_tea; q.tna = s_tna; q.vol = s_vol;
    q.f_pair_in = T + c->o_fpair; q.Eloss_sy = s_sy; q.ec_old = s_eco; q.turb_lev = s_tl;
    q.f_nt = D + p_f; q.Pnt = D + p_P; q.n_e = D + p_ne; q.Te_new = D + p_Te; q.gmin = D + p_gmin; q.gmax = D + p_gmax;
    q.p_nth = D + p_pn; q.amxwl = D + p_am; q.f_pair = D + p_fp;
    q.zE_old = D + p_Eo; q.zE_new = D + p_En; q.z_hr = D + p_hr; q.z_hrst = D + p_hs; q.z_dT = D + p_dT; q.z_misc = D + p_misc;
    q.fp_steps = reinterpret_cast<int *>(D + p_steps);
    if (int rc = fp_scratch(c, q)) return rc;
    if (n_own > 0) {
        c->timer_begin(3);
        launch_fp(q, c->stream);
        k_fp_pack<<<n_own, 128, 0, c->stream>>>(q, D + p_send, rec);
        c->timer_end();
        CUF(cudaGetLastError());
    }
    c->timer_begin(9);
    rc = comm_allgather(c, D + p_send, D + p_recv, (size_t)n_slot * rec * 8);
    c->timer_end();
    if (rc) return rc;
    k_fp_unpack<<<nzone, 128, 0, c->stream>>>(D + p_recv, world, n_slot, rec, nzone, N, T + c->o_fnt, T + c->o_pnt, T + c->o_ne,
                                              T + c->o_fpair, T + c->o_tea, s_tna, s_gmin, s_gmax, s_pn, s_am, s_Te, D + p_zi);
    CUF(cudaGetLastError());
    CUF(cudaMemcpyAsync(zinfo_host, D + p_zi, (size_t)nzone * kFpRecExtra * 8, cudaMemcpyDeviceToHost, c->stream));
    CUF(cudaStreamSynchronize(c->stream));
    return 0;
#endif
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_ic_loss(c2d_ctx *h, double *F_IC_out)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    CUF(cudaSetDevice(c->device));
    int rc = fp_state(c);
    if (rc) return rc;
    const int N = c->dims.num_nt, nf = c->dims.nphfield;
    c->timer_begin(3);
    launch_ic_loss(c->P.gnt, c->P.E_field, N, nf, c->fp->d_FIC, c->stream);
    launch_transpose(c->fp->d_FIC, N, nf, c->fp->d_FICt, c->stream);
    c->timer_end();
    CUF(cudaGetLastError());
    c->fp->have_FIC = true;
    if (F_IC_out) CUF(cudaMemcpyAsync(F_IC_out, c->fp->d_FIC, (size_t)N * nf * 8, cudaMemcpyDeviceToHost, c->stream));
    CUF(cudaStreamSynchronize(c->stream));
    return 0;
}

int c2d_gamma_bar(c2d_ctx *h, int64_t n, const double *Theta, double *gbar, double *gbar_table)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (n < 0 || (n && (!Theta || !gbar))) return fail(C2D_EINVAL, "NULL argument");
    for (int64_t i = 0; i < n; i++) if (!(Theta[i] > 0.0)) return fail(C2D_EINVAL, "Theta[%lld] = %g", (long long)i, Theta[i]);
    if (n == 0) return 0;
    CUF(cudaSetDevice(c->device));
    int rc = fp_state(c);
    if (rc) return rc;
    double *d = nullptr;
    CUF(cudaMalloc(&d, (size_t)3 * n * 8));
    CUF(cudaMemcpyAsync(d, Theta, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    FpParams q;
    memset(&q, 0, sizeof q);
    q.qt_x15 = c->fp->d_quad; q.gtab = c->fp->d_gtab; q.gtab_n = kGtabN; q.gtab_x0 = kGtabX0; q.gtab_ih = 1.0 / kGtabH;
    launch_gbar_query(q, (int)n, d, d + n, d + 2 * n, c->stream);
    CUF(cudaGetLastError());
    CUF(cudaMemcpyAsync(gbar, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (gbar_table) CUF(cudaMemcpyAsync(gbar_table, d + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUF(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

int c2d_update_fp(c2d_ctx *h, const c2d_fp_in *in, c2d_fp_out *out)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !in || !out) return fail(C2D_ESTATE, "context not initialised / NULL argument");
    CUF(cudaSetDevice(c->device));
    int rc = fp_state(c);
    if (rc) return rc;
    const c2d_dims &d = c->dims;
    const int N = d.num_nt, nf = d.nphfield, nzone = c->nzone;
    if (!in->f_nt || !in->Pnt || !in->n_e || !in->B_field || !in->tea || !in->tna || !in->vol || !in->Eloss_sy || !in->ec_old ||
        !in->turb_lev)
        return fail(C2D_EINVAL, "c2d_update_fp: missing per-zone input");
    if (!out->f_nt || !out->Pnt || !out->n_e || !out->Te_new || !out->gmin || !out->gmax || !out->p_nth || !out->amxwl)
        return fail(C2D_EINVAL, "c2d_update_fp: missing output array");
    if (in->zone_stride < 1 || in->first_zone < 0) return fail(C2D_EINVAL, "bad zone partition");
    if (!(in->dt > 0.0)) return fail(C2D_EINVAL, "dt = %g", in->dt);
    int n_zones = (in->first_zone < nzone) ? (nzone - in->first_zone + in->zone_stride - 1) / in->zone_stride : 0;
    if (in->n_zones >= 0 && in->n_zones < n_zones) n_zones = in->n_zones;
    if (in->F_IC) {
        CUF(cudaMemcpyAsync(c->fp->d_FIC, in->F_IC, (size_t)N * nf * 8, cudaMemcpyHostToDevice, c->stream));
        launch_transpose(c->fp->d_FIC, N, nf, c->fp->d_FICt, c->stream);
        c->fp->have_FIC = true;
    } else if (!c->fp->have_FIC) {
        launch_ic_loss(c->P.gnt, c->P.E_field, N, nf, c->fp->d_FIC, c->stream);
        launch_transpose(c->fp->d_FIC, N, nf, c->fp->d_FICt, c->stream);
        c->fp->have_FIC = true;
    }
    // staging blob: inputs then outputs
    size_t n = 0;
    auto take = [&](size_t cnt) { size_t o = n; n += (cnt + 1) & ~size_t(1); return o; };
    const size_t o_f = take((size_t)nzone * N), o_P = take((size_t)nzone * N), o_nf = take((size_t)nzone * nf), o_ec = take(nzone),
                 o_ne = take(nzone), o_B = take(nzone), o_tea = take(nzone), o_tna = take(nzone), o_vol = take(nzone),
                 o_fp = take(nzone), o_sy = take(nzone), o_eco = take(nzone), o_tl = take(nzone);
    const size_t p_f = take((size_t)nzone * N), p_P = take((size_t)nzone * N), p_ne = take(nzone), p_Te = take(nzone),
                 p_gmin = take(nzone), p_gmax = take(nzone), p_pn = take(nzone), p_am = take(nzone), p_fp = take(nzone),
                 p_Eo = take(nzone), p_En = take(nzone), p_hr = take(nzone), p_hs = take(nzone), p_dT = take(nzone),
                 p_misc = take((size_t)3 * nzone), p_steps = take(nzone);
    if (n > c->fp->io_cap) {
        cudaFree(c->fp->d_io);
        CUF(cudaMalloc(&c->fp->d_io, n * 8));
        c->fp->io_cap = n;
    }
    double *D = c->fp->d_io;
    auto up = [&](size_t off, const double *src, size_t cnt) -> int {
        if (src) CUF(cudaMemcpyAsync(D + off, src, cnt * 8, cudaMemcpyHostToDevice, c->stream));
        else CUF(cudaMemsetAsync(D + off, 0, cnt * 8, c->stream));
        return 0;
    };
    // Pnt: FP_calc rebuilds the running CDF and keeps only its last entry (update2d.f:1485) -- unless a zone is empty
    // (n_lept < 1e-11, l.472), whose Pnt is handed back untouched.  Upload the full table only then.
    bool need_pnt = false;
    for (int z = 0; z < nzone && !need_pnt; z++) {
        const double fpz = in->f_pair ? in->f_pair[z] : 0.0;
        need_pnt = !(in->n_e[z] * (1. + fpz) + in->n_e[z] * fpz >= 1.e-11);        // n_lept as k_fp_zones forms it
    }
    if (need_pnt) rc = up(o_P, in->Pnt, (size_t)nzone * N);
    else {
        CUF(cudaMemcpy2DAsync(D + o_P + (N - 1), (size_t)N * 8, in->Pnt + (N - 1), (size_t)N * 8, 8, (size_t)nzone,
                              cudaMemcpyHostToDevice, c->stream));
        rc = 0;
    }
    if (rc) return rc;
    rc = up(o_f, in->f_nt, (size_t)nzone * N) | up(o_ne, in->n_e, nzone) |
         up(o_B, in->B_field, nzone) | up(o_tea, in->tea, nzone) | up(o_tna, in->tna, nzone) | up(o_vol, in->vol, nzone) |
         up(o_fp, in->f_pair, nzone) | up(o_sy, in->Eloss_sy, nzone) | up(o_eco, in->ec_old, nzone) | up(o_tl, in->turb_lev, nzone);
    if (rc) return rc;
    if (in->n_field) { rc = up(o_nf, in->n_field, (size_t)nzone * nf); if (rc) return rc; }
    if (in->ecens) { rc = up(o_ec, in->ecens, nzone); if (rc) return rc; }
    // gmin / gmax / p_nth / amxwl are in/out (zones whose amxwl -> 1 keep their old p_nth)
    rc = up(p_gmin, out->gmin, nzone) | up(p_gmax, out->gmax, nzone) | up(p_pn, out->p_nth, nzone) | up(p_am, out->amxwl, nzone);
    if (rc) return rc;
    FpParams q;
    fp_fill_scalars(c, in, q);
    q.first_zone = in->first_zone; q.n_zones = n_zones; q.zone_stride = in->zone_stride;
    q.f_nt_in = D + o_f; q.Pnt_in = D + o_P;
    q.n_field = in->n_field ? D + o_nf : c->P.n_field;
    q.ecens = in->ecens ? D + o_ec : c->P.ecens;
    q.n_e_in = D + o_ne; q.B_field = D + o_B; q.tea = D + o_tea; q.tna = D + o_tna; q.vol = D + o_vol;
    q.f_pair_in = D + o_fp; q.Eloss_sy = D + o_sy; q.ec_old = D + o_eco; q.turb_lev = D + o_tl;
    q.f_nt = D + p_f; q.Pnt = D + p_P; q.n_e = D + p_ne; q.Te_new = D + p_Te; q.gmin = D + p_gmin; q.gmax = D + p_gmax;
    q.p_nth = D + p_pn; q.amxwl = D + p_am; q.f_pair = D + p_fp;
    q.zE_old = D + p_Eo; q.zE_new = D + p_En; q.z_hr = D + p_hr; q.z_hrst = D + p_hs; q.z_dT = D + p_dT; q.z_misc = D + p_misc;
    q.fp_steps = reinterpret_cast<int *>(D + p_steps);
    if (int rc2 = fp_scratch(c, q)) return rc2;
    if (n_zones > 0) {
        c->timer_begin(3);
        launch_fp(q, c->stream);
        c->timer_end();
        CUF(cudaGetLastError());
    }
    // copy back only this call's zones
    std::vector<double> hb(n - p_f);
    CUF(cudaMemcpyAsync(hb.data(), D + p_f, (n - p_f) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUF(cudaStreamSynchronize(c->stream));
    auto H = [&](size_t off) { return hb.data() + (off - p_f); };
    const int *steps = reinterpret_cast<const int *>(H(p_steps));
    out->dT_max = 0; out->E_tot_old = 0; out->E_tot_new = 0; out->hr_total = 0; out->hr_st_total = 0;
    for (int sidx = 0; sidx < n_zones; sidx++) {
        const int zi = in->first_zone + sidx * in->zone_stride;
        memcpy(out->f_nt + (size_t)zi * N, H(p_f) + (size_t)zi * N, (size_t)N * 8);
        memcpy(out->Pnt + (size_t)zi * N, H(p_P) + (size_t)zi * N, (size_t)N * 8);
        out->n_e[zi] = H(p_ne)[zi]; out->Te_new[zi] = H(p_Te)[zi]; out->gmin[zi] = H(p_gmin)[zi]; out->gmax[zi] = H(p_gmax)[zi];
        out->p_nth[zi] = H(p_pn)[zi]; out->amxwl[zi] = H(p_am)[zi];
        if (out->f_pair) out->f_pair[zi] = H(p_fp)[zi];
        if (out->fp_steps) out->fp_steps[zi] = steps[zi];
        if (out->zone_E_old) out->zone_E_old[zi] = H(p_Eo)[zi];
        if (out->zone_E_new) out->zone_E_new[zi] = H(p_En)[zi];
        if (out->zone_hr) out->zone_hr[zi] = H(p_hr)[zi];
        if (out->zone_hr_st) out->zone_hr_st[zi] = H(p_hs)[zi];
        if (out->zone_dT) out->zone_dT[zi] = H(p_dT)[zi];
        out->E_tot_old += H(p_Eo)[zi]; out->E_tot_new += H(p_En)[zi];
        out->hr_total += H(p_hr)[zi]; out->hr_st_total += H(p_hs)[zi];
        if (H(p_dT)[zi] > out->dT_max) out->dT_max = H(p_dT)[zi];
        out->dr = H(p_misc)[3 * zi]; out->dz = H(p_misc)[3 * zi + 1]; out->f_t_implicit = H(p_misc)[3 * zi + 2];
    }
    return 0;
}


}  // extern "C"

