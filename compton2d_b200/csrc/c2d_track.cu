// c2d_track.cu -- history kernels of the Compton2d hot path (sm_100a) and their launchers.
//
//   k_history<SRC>   one thread per top-level history = one imctrk2d(-1) call of the reference
//                    (imcvol2d_para.f:392 for volume sources, imcfield2d.f:128 for census photons):
//                    source sampling, the collision-free ray, split1 trial draws, absorption
//                    deposition, the recombined scat_flag = 0 flight with its census / leak tallies.
//   k_children       one thread per (Compton collision, child): the split2 loop of imctrk2d.f:619-689 --
//                    compb2d, then the scat_flag = 1 flight; collisions of those flights go to the next
//                    generation's queue, children that trigger the 3rd split go to the split3 queue.
//   k_split3_attempts one warp per (3rd-split child, ii2): the rejection loop of imctrk2d.f:641-656,
//                    attempts spread over the lanes; k_split3_flights: one thread per accepted sub-child.
#include "c2d_host.h"

#include <algorithm>
#ifndef C2D_EMU
#include <cooperative_groups.h>
#endif

namespace c2d {

__device__ __forceinline__ void push_collision(const Params &p, const Photon &P, uint64_t sid)
{
    const unsigned long long n = atomicAdd(p.q_out_count, 1ull);
    if (n >= p.q_cap) { count(p, CNT_FATAL); return; }
    CollTask &t = p.q_out[n];
    t.p = P; t.pad = 0; t.sid = sid;
}

// ------------------------------------------------------------------------------------------ tally hierarchy
// The per-zone tallies of a history (edep, prdep per crossing; npcen, ecens at census) go warp -> shared memory -> HBM: when
// the lanes of a warp agree on the zone (the census bank is kept sorted by zone, so they usually do) their contributions
// are summed with shuffles and lane 0 adds the sum to the warp's shared-memory slot of that zone; each slot is committed to
// HBM with one atomic per tally when the warp retires.  Lanes that disagree, and zones that find no free slot, fall back to
// global atomics.  The slots are per WARP, not per block: a block-wide cache needs a barrier before its flush, and the
// warps of a block finish their histories at very different times (measured: 10 % of the stall samples at that barrier).
constexpr int kTallySlots = 8;
struct BlockTally {                                      // one per warp
    int key[kTallySlots];                                // zone, -1 = free
    unsigned cen_n[kTallySlots];                         // npcen
    double dep_e[kTallySlots], dep_p[kTallySlots];       // edep, prdep
    double cen_e[kTallySlots];                           // ecens
};
#ifndef C2D_EMU
__device__ __forceinline__ void bt_init(BlockTally &bt)
{
    const int l = lane_id();
    if (l < kTallySlots) { bt.key[l] = -1; bt.dep_e[l] = 0.0; bt.dep_p[l] = 0.0; bt.cen_e[l] = 0.0; bt.cen_n[l] = 0u; }
    __syncwarp();
}
// lane 0 only
__device__ __forceinline__ int bt_slot(BlockTally &bt, int z)
{
    int s = z & (kTallySlots - 1);
#pragma unroll 1
    for (int k = 0; k < kTallySlots; k++) {
        const int old = bt.key[s];
        if (old == z) return s;
        if (old == -1) { bt.key[s] = z; return s; }
        s = (s + 1) & (kTallySlots - 1);
    }
    return -1;
}
__device__ __forceinline__ void bt_flush(const Params &p, BlockTally &bt)
{
    __syncwarp();
    const int s = lane_id();
    if (s < kTallySlots && bt.key[s] >= 0) {
        const int z = bt.key[s];
        if (bt.dep_e[s] != 0.0) atomic_add_f64(p.edep + z, bt.dep_e[s]);
        if (bt.dep_p[s] != 0.0) atomic_add_f64(p.prdep + z, bt.dep_p[s]);
        if (bt.cen_n[s]) { atomicAdd(p.npcen + z, (unsigned long long)bt.cen_n[s]); atomic_add_f64(p.ecens + z, bt.cen_e[s]); }
    }
}
#endif
// every lane of the warp calls; lanes with valid = false contribute nothing
__device__ __forceinline__ void tally_dep(const Params &p, BlockTally *bt, bool valid, int z, double a, double b)
{
#ifndef C2D_EMU
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (!m) return;
    const int z0 = __shfl_sync(0xffffffffu, z, __ffs(m) - 1);
    if (bt && __all_sync(0xffffffffu, !valid || z == z0)) {
        a = warp_sum(valid ? a : 0.0); b = warp_sum(valid ? b : 0.0);
        if (lane_id() == 0) {
            const int s = bt_slot(*bt, z0);
            if (s >= 0) { bt->dep_e[s] += a; bt->dep_p[s] += b; }
            else { atomic_add_f64(p.edep + z0, a); atomic_add_f64(p.prdep + z0, b); }
        }
        return;
    }
#else
    (void)bt;
#endif
    if (valid) { atomic_add_f64(p.edep + z, a); atomic_add_f64(p.prdep + z, b); }
}
__device__ __forceinline__ void tally_census_zone(const Params &p, BlockTally *bt, bool valid, int z, double ew)
{
#ifndef C2D_EMU
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (!m) return;
    const int z0 = __shfl_sync(0xffffffffu, z, __ffs(m) - 1);
    if (bt && __all_sync(0xffffffffu, !valid || z == z0)) {
        ew = warp_sum(valid ? ew : 0.0);
        if (lane_id() == 0) {
            const int s = bt_slot(*bt, z0);
            if (s >= 0) { bt->cen_e[s] += ew; bt->cen_n[s] += (unsigned)__popc(m); }
            else { atomicAdd(p.npcen + z0, (unsigned long long)__popc(m)); atomic_add_f64(p.ecens + z0, ew); }
        }
        return;
    }
#else
    (void)bt;
#endif
    if (valid) census_zone_tally(p, z, ew);
}

// per-block shared-memory copy of nelectron for the team kernels (the passes hammer a handful of low-gamma bins, which global
// atomics would serialise); flushed once per block
constexpr int kHistMax = 1024;
#ifdef C2D_EMU
#define C2D_HIST_BEGIN(p) unsigned *hist = nullptr;
#define C2D_HIST_END(p)
#else
#define C2D_HIST_BEGIN(p)                                                                             \
    __shared__ unsigned s_hist[kHistMax];                                                             \
    unsigned *hist = ((p).num_nt + 2 <= kHistMax) ? s_hist : nullptr;                                 \
    if (hist) for (int i_ = threadIdx.x; i_ < (p).num_nt + 2; i_ += blockDim.x) s_hist[i_] = 0u;      \
    __syncthreads();
#define C2D_HIST_END(p)                                                                               \
    __syncthreads();                                                                                  \
    if (hist) for (int i_ = threadIdx.x; i_ < (p).num_nt + 2; i_ += blockDim.x)                       \
        if (s_hist[i_]) atomicAdd((p).nelectron + i_, (unsigned long long)(long long)(int)s_hist[i_]);   /* signed: the grid teams take tallies back */
#endif


// nelectron tallies of the team kernels: into the block's copy (or global memory when num_nt is too large for it)
__device__ __forceinline__ void tally_bin(const Params &p, unsigned *hist, unsigned bin, int sign)
{
    if (hist) atomicAdd(hist + bin, (unsigned)sign); else atomicAdd(p.nelectron + bin, (unsigned long long)(long long)sign);
}
// ------------------------------------------------------------------------------------------ rare paths of a history
// A trial flight that may collide is flown literally (imctrk2d.f:126-695 body for scat_flag = -1) with comtot evaluated per
// crossing (the same values as the per-history cache comac_ar of l.175-183).  It makes its own deposits, so the share the
// history deposited for it as one of the `split1` identical collision-free trials is taken back first.
static __device__ __noinline__ int literal_trial(const Params &p, const Photon &T0, uint64_t S_trials, int t, double u, double wtmin)
{
    {   // take back one trial's share of the collision-free ray's absorption (what step D below deposited for it)
        Photon Q = T0;
        Rng dummy; dummy.stream(0);
        FlightCtx c;
        flight_begin(p, Q.xnu, c);
        double wmu_c = T0.wmu;
        if (wmu_c > 1.0) wmu_c = 1.0;
        if (wmu_c < -1.0) wmu_c = -1.0;
        for (;;) {
            if (Q.ew < 1.e-40) break;
            if (Q.wmu > 1.0) Q.wmu = 1.0;
            if (Q.wmu < -1.0) Q.wmu = -1.0;
            Seg g; int s = 0, ns = 0; double colmfp = 0.0;
            const int r = fly_step<MODE_PRE>(p, Q, dummy, wtmin, colmfp, 0.0, c, &g, s, ns);
            atomic_add_f64(p.edep + g.zi, -g.dele);
            if (g.sigabs * g.L <= .00001) atomic_add_f64(p.prdep + g.zi, -(g.dele * wmu_c * kCLight));
            if (r != STEP_CONT) break;
        }
    }
    Rng rng; rng.stream(mix64(S_trials, (uint64_t)(t + 1)));
    Photon P = T0;
    int ns = 0;
    count(p, CNT_LITERAL);
    const int end = fly<MODE_TRIAL_SLOW>(p, P, rng, wtmin, -log(u), nullptr, ns, S_trials, (uint32_t)t);
    if (end == END_COLLISION) {
        count(p, CNT_CMCOUNT1);
        push_collision(p, P, rng.sid);
        return 1;
    }
    return 0;
}

// One top-level history = one imctrk2d(-1) call.  Every lane of the warp runs this function (inactive lanes carry
// `active = false`): the crossing loop is warp-uniform so that comtot (sigma_mode 0) and the tallies can cooperate.
//
// The ray of the split1 trial flights is traced ONCE and nothing of it is stored: per crossing the scattering opacity is
// looked up, the ray's optical depth accumulated, the absorption of all split1 trials deposited, and the weight of the
// recombined scat_flag = 0 flight attenuated alongside.  What needs the ray again is rare and re-traces it: a trial that may
// collide (literal_trial), the absorption-point draws of rays with sigabs L > 1e-5 (k_absdraw, queued), a recombined flight
// that outlives the trial ray or follows collisions.
__device__ __forceinline__ void history(const Params &p, bool active, const Photon &P0, uint64_t S_entry, BlockTally *bt)
{
    const double wtmin = 1.e-10 * P0.ew;                          // imctrk2d.f:82,92
    Photon T0 = P0;
    T0.ew = P0.ew / p.split1;                                     // l.110
    const uint64_t S_trials = mix64(S_entry, kTagTrials);
    const bool hist_active = active;
    double wmu_c = T0.wmu;
    if (wmu_c > 1.0) wmu_c = 1.0;
    if (wmu_c < -1.0) wmu_c = -1.0;
    const double n_all = (double)p.split1;

    // A + B + D + E in one pass over the crossings
    Photon Q = T0;                                                // runs along the collision-free ray with a trial's weight
    int endA = END_LOWEW, nseg = 0;
    double tau = 0.0;
    bool has_rand = false;
    // the recombined flight (l.698-712) carries split1 trial weights unless trials collide (then it is flown literally below)
    double F_ew = n_all * T0.ew;
    const double wt0 = 1.e-10 * F_ew;
    uint32_t F_d = 0;
    int F_state = 0;                                              // 0 alive, 1 dropped silently (ew < 1e-40), 2 killed
    SigCtx sc; sc.j = -2;
    FlightCtx fc;
    if (active) {
        flight_begin(p, T0.xnu, fc);
        if (p.sigma_mode == 1) sig_begin(p, T0.xnu, sc);
    }
    int last_z = -1;
    double last_c = 0.0;
    bool go = active;
    while (warp_ballot(go)) {
        Seg g;
        g.zi = 0; g.dele = 0.0; g.sigabs = 0.0; g.L = 0.0; g.sigsc = 1.0; g.att = 1.0;
        int r = STEP_CONT;
        bool have = false;
        if (go) {
            if (Q.ew < 1.e-40) { endA = END_LOWEW; go = false; }   // l.161
            else {
                if (Q.wmu > 1.0) Q.wmu = 1.0;
                if (Q.wmu < -1.0) Q.wmu = -1.0;
                Rng dummy; dummy.stream(0);
                int s = 0, ns = 0; double colmfp = 0.0;
                r = fly_step<MODE_PRE>(p, Q, dummy, wtmin, colmfp, 0.0, fc, &g, s, ns);
                have = true; nseg++;
            }
        }
        // B: comtot of this crossing's zone (comac_ar of l.175-183: one value per zone and history)
        double comac = last_c;
        const bool need = have && g.zi != last_z;
        if (p.sigma_mode == 1) {
            if (need) comac = (sc.j != -2) ? comtot_table_zone(p, sc, g.zi) : comtot_one(p, g.zi, T0.xnu, sc);
        } else {
            unsigned todo = warp_ballot(need);
            while (todo) {                                         // the warp serves one lane's photon at a time
                const int src = first_lane(todo);
                todo &= todo - 1;
                const double xs = warp_shfl(T0.xnu, src);
                const int zs = warp_shfl(g.zi, src);
                const double v = comtot_zone_warp(p, xs, zs);
                if (lane_id() == src) comac = v;
            }
        }
        if (need) { last_z = g.zi; last_c = comac; }
        double d_e = 0.0, d_p = 0.0;
        if (have) {
            const double sigsc = 1.0 * comac * g.sigsc;
            tau += sigsc * g.L;
            const double xabs = g.sigabs * g.L;
            const bool rnd = !(xabs <= .00001);
            has_rand |= rnd;
            // D: absorption of the split1 trials in this crossing (l.465-466); the prdep of a crossing with absorption-point
            // draws comes from k_absdraw
            d_e = n_all * g.dele;
            d_p = rnd ? 0.0 : n_all * (g.dele * wmu_c * kCLight);
            // E: the recombined flight's weight through the same crossing
            if (F_state == 0) {
                if (F_ew < 1.e-40) F_state = 1;
                else {
                    const double ewnew = (xabs < 100.) ? F_ew * g.att : 0.;
                    if (rnd) F_d++;                               // the absorption-point draw is consumed, its value unused
                    if (ewnew <= wt0) F_state = 2; else F_ew = ewnew;
                }
            }
            if (r != STEP_CONT) { endA = r; go = false; }
        }
        tally_dep(p, bt, have, g.zi, d_e, d_p);
    }
    if (active && endA == END_LOWEW && nseg == 0) active = false; // ew < 1e-40 at entry: nothing is tallied (l.161)

    // C: the split1 trial draws (l.126-156).  A trial collides only if -ln(u) is within the ray's scattering depth, i.e.
    // u >= thr; everything below thr provably does not (margins >> rounding).  Whatever reaches thr is flown literally.
    int nscat = 0, nlit = 0;
    unsigned long long lit_off = 0;
    if (active) {
        const double thr = exp(-tau * (1.0 + 1.e-12)) * (1.0 - 1.e-12);
        // u = (K + 0.5) 2^-52 >= thr  <=>  K >= ceil(thr 2^52 - 0.5): the filter runs on integers
        const double tk = thr * 4503599627370496.0 - 0.5;
        const uint64_t Kthr = (tk <= 0.0) ? 0ull : (uint64_t)ceil(tk);
        // K = coarse8 << 44 | fine44 (TrialGen): a trial whose coarse byte is below the threshold's cannot collide.
        // Thin ray (threshold byte 0xFF): only the candidates are visited, ~split1/256 of them; otherwise every trial.
        const uint32_t Cthr = (uint32_t)(Kthr >> 44);
        auto examine = [&](int t, uint32_t coarse) {
            const uint64_t K = ((uint64_t)coarse << 44) | trial_fine44(p, S_trials, (uint32_t)t);
            if (K < Kthr) return;
            const double u = ((double)K + 0.5) * (1.0 / 4503599627370496.0);
            if (has_rand) {                                        // k_absdraw must leave this trial's draws to its literal flight
                if (nlit == 0) lit_off = atomicAdd(p.lit_count, (unsigned long long)p.split1);
                if (lit_off + (unsigned long long)nlit < p.lit_cap) p.lit_pool[lit_off + nlit] = t; else count(p, CNT_FATAL);
            }
            nlit++;
            nscat += literal_trial(p, T0, S_trials, t, u, wtmin);
        };
        if (Cthr <= 0xFFu) {                                      // 0x100: tau = 0, nothing can collide
            TrialGen tg; tg.init(p, S_trials);
            if (Cthr == 0xFFu) {
                while (tg.next_c < (long long)p.split1) { examine((int)tg.next_c, 0xFFu); tg.advance(p); }
            } else {
                for (int t = 0; t < p.split1; t++) {
                    const uint32_t coarse = tg.coarse(p, (uint32_t)t);
                    if (coarse >= Cthr) examine(t, coarse);
                }
            }
        }
        // C': absorption-point draws of l.451-459 for the other trials (they only move prdep): queued for k_absdraw
        if (has_rand) {
            const unsigned long long n = atomicAdd(p.qabs_count, 1ull);
            if (n >= p.qabs_cap) count(p, CNT_FATAL);
            else {
                AbsTask &a = p.qabs[n];
                a.p = T0; a.nlit = nlit; a.lit_off = lit_off; a.S_trials = S_trials; a.wtmin = wtmin;
            }
        }
    }
    {   // counters, aggregated over the warp
        const unsigned long long nh = warp_sum((unsigned long long)(hist_active ? 1 : 0));
        const unsigned long long nt = warp_sum((unsigned long long)(active ? p.split1 : 0));
        const unsigned long long ns = warp_sum((unsigned long long)(active ? nseg : 0));
        if (lane_id() == 0) {
            if (nh) count(p, CNT_HISTORIES, nh);
            if (nt) count(p, CNT_TRIALS, nt);
            if (ns) count(p, CNT_SEGMENTS, ns);
        }
    }
    // E: the recombined flight of the unscattered weight (l.698-712): scat_flag = 0, no collisions, no deposition.  Without
    // collisions it followed the ray above: its census / leak tallies are made at the ray's end point.  Otherwise (trials
    // collided, the trial ray was killed first -- a trial carries 1/split1 of the weight --, the photon passes through the
    // axis) it is flown literally.
    Rng rng; rng.stream(mix64(S_entry, kTagF0));
    Photon F = P0;
    bool do_census = false, literal = false;
    if (active && p.split1 - nscat > 0) {
        if (nscat == 0 && (endA == END_CENSUS || endA == END_LEAK)) {
            if (F_state == 2) count(p, CNT_KILLED);
            if (F_state == 0) {
                F.ew = F_ew; F.rpre = Q.rpre; F.zpre = Q.zpre; F.phi = Q.phi; F.wmu = Q.wmu; F.dcen = Q.dcen; F.jph = Q.jph; F.kph = Q.kph;
                rng.d = F_d;
                if (endA == END_CENSUS) do_census = true;
                else {
                    const int idead = imcleak(p, F, rng);
                    count(p, CNT_LKCOUNT);
                    literal = (idead != 1);                       // through the axis: the flight goes on
                }
            }
        } else {
            F.ew = (p.split1 - nscat) * T0.ew;
            literal = true;
        }
    }
    tally_census_zone(p, bt, do_census, (F.jph - 1) * p.nr + (F.kph - 1), F.ew);
    if (do_census) census_record(p, F, rng);
    if (literal) {
        int dummy = 0;
        const double wtF = 1.e-10 * ((p.split1 - nscat) * T0.ew);
        const int end = fly<MODE_F0>(p, F, rng, wtF, 0.0, nullptr, dummy);
        if (end == END_COLLISION) count(p, CNT_FATAL);
    }
}

// 64 registers, 32 warps / SM (profiles/r01_ncu_v31_k_history.md); measured block sizes at that occupancy:
// 64 x 16: 65.5, 128 x 8: 63.4, 256 x 4: 61.1, 512 x 2: 63.6, 1024 x 1: 70.6 ms / step
#ifndef C2D_HIST_THREADS
#define C2D_HIST_THREADS 256
#endif
#ifndef C2D_HIST_MINB
#define C2D_HIST_MINB 4
#endif
template <int SRC>
__global__ void __launch_bounds__(C2D_HIST_THREADS, C2D_HIST_MINB) k_history(const __grid_constant__ Params p, long long first, long long n, long long stride)
{
#ifndef C2D_EMU
    __shared__ BlockTally s_bt[C2D_HIST_THREADS / 32];
    BlockTally *bt = &s_bt[threadIdx.x >> 5];
    bt_init(*bt);
#else
    BlockTally *bt = nullptr;
#endif
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = idx < n;
    Photon P0 = {1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1, 1, 0, 0, 1};
    uint64_t S_entry = 0;
    if (!active) {
    } else if (SRC == 0) {
        const long long h = first + idx * stride;
        // zone (1-based) with nsv_prefix[zone-1] <= h < nsv_prefix[zone]
        int lo = 0, hi = p.nzone;            // invariant: prefix[lo] <= h < prefix[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.nsv_prefix[mid] <= h) lo = mid; else hi = mid;
        }
        const int zone = lo + 1;
        const long long ivol = h - p.nsv_prefix[lo] + 1;
        S_entry = mix64(mix64(kTagVol, (uint64_t)zone), (uint64_t)ivol);
        Rng rng; rng.stream(S_entry);
        source_volume(p, rng, zone, P0);
    } else if (SRC == 2) {
        const long long h = first + idx * stride;
        const int nslot = 2 * p.nz + 2 * p.nr;
        int lo = 0, hi = nslot;              // invariant: prefix[lo] <= h < prefix[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.surf_prefix[mid] <= h) lo = mid; else hi = mid;
        }
        const long long nxs = h - p.surf_prefix[lo] + 1;
        const int kind = (lo < 2 * p.nz) ? (lo & 1) : 2 + ((lo - 2 * p.nz) & 1);
        const int ring = (lo < 2 * p.nz) ? (lo >> 1) + 1 : ((lo - 2 * p.nz) >> 1) + 1;
        S_entry = mix64(mix64(kTagSurf + (uint64_t)kind, (uint64_t)ring), (uint64_t)nxs);
        Rng rng; rng.stream(S_entry);
        source_surface(p, rng, lo, P0);
    } else {
        const long long h = first + idx * stride;
        P0.rpre = p.in.rpre[h]; P0.zpre = p.in.zpre[h]; P0.wmu = p.in.wmu[h]; P0.phi = p.in.phi[h];
        P0.ew = p.in.ew[h]; P0.xnu = p.in.xnu[h];
        const uint32_t b = p.in.bins[h], zz = p.in.zone[h];
        P0.jgpsp = b & 0xFFFF; P0.jgplc = (b >> 16) & 0xFF; P0.jgpmu = b >> 24;
        P0.jph = zz & 0xFFFF; P0.kph = zz >> 16;
        P0.dcen = kCLight * p.dt;                                 // imcfield2d.f:117
        if (P0.wmu > 1.0) P0.wmu = 1.0;
        if (P0.wmu < -1.0) P0.wmu = -1.0;
        S_entry = p.in.sid[h];
    }
    if (SRC != 1 && p.bins_off) { P0.jgpsp = 0; P0.jgplc = 0; P0.jgpmu = 0; }
    history(p, active, P0, S_entry, bt);
#ifndef C2D_EMU
    bt_flush(p, *bt);
#endif
}

// Absorption-point draws (imctrk2d.f:451-459) of the collision-free trials of one queued history: a warp re-traces the ray
// (every lane the same ray) and, for each crossing with sigabs L > 1e-5, spreads the split1 trials over its lanes -- one
// uniform per (trial, absorbing crossing) from the trial's stream, two trials per Philox call -- leaving out the literal
// trials, which made their own draws.  Only prdep moves.
__device__ __forceinline__ void absdraw_task(const Params &p, const AbsTask &a)
{
    const int lane = lane_id();
    const uint64_t S = a.S_trials;
    const int nlit = a.nlit;
    const int *lit = p.lit_pool + a.lit_off;
    Photon Q = a.p;
    double wm = Q.wmu;
    if (wm > 1.0) wm = 1.0;
    if (wm < -1.0) wm = -1.0;
    FlightCtx fc;
    flight_begin(p, Q.xnu, fc);
    Rng dummy; dummy.stream(0);
    uint32_t k = 0;
    for (;;) {
        if (Q.ew < 1.e-40) break;
        if (Q.wmu > 1.0) Q.wmu = 1.0;
        if (Q.wmu < -1.0) Q.wmu = -1.0;
        Seg g; int s = 0, ns = 0; double colmfp = 0.0;
        const int r = fly_step<MODE_PRE>(p, Q, dummy, a.wtmin, colmfp, 0.0, fc, &g, s, ns);
        if (!(g.sigabs * g.L <= .00001)) {
            double acc = 0.0;
            const double frac = g.dele / g.ew, inv_sa = 1.0 / g.sigabs, rp = g.rpre;   // hoisted out of the trial loop (<= 1 ulp on prdep)
            for (int t2 = 2 * lane; t2 < p.split1; t2 += 2 * kWarp) {                  // trials t2, t2 + 1: one Philox call
                uint32_t o[4];
                philox4x32_10((uint32_t)S, (uint32_t)(S >> 32), (uint32_t)t2 >> 1, 2u + k, p.key0, p.key1, o);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int t = t2 + h;
                    if (t >= p.split1) continue;
                    bool is_lit = false;
                    for (int m = 0; m < nlit; m++) is_lit |= (lit[m] == t);
                    if (is_lit) continue;
                    const double mb_ran = h ? u52(o[2], o[3]) : u52(o[0], o[1]);
                    const double sstar = -log(1. - mb_ran * frac) * inv_sa;
                    const double denom = sqrt(rp * rp + 2. * wm * rp * sstar + sstar * sstar);
                    acc += g.dele * ((wm * rp + sstar) / denom) * kCLight;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) atomic_add_f64(p.prdep + g.zi, acc);
            k++;
        }
        if (r != STEP_CONT) break;
    }
}
// persistent warps over a task count that lives on the device (k_history has just written it: no host round trip)
__global__ void __launch_bounds__(128) k_absdraw(const __grid_constant__ Params p, const AbsTask *q, const unsigned long long *n_dev)
{
    const unsigned long long n = *n_dev;
    const unsigned long long w0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) / kWarp;
    for (unsigned long long w = w0; w < n && w < p.qabs_cap; w += nw) absdraw_task(p, q[w]);
}

// scattered photon after compb2d: tallies of imctrk2d.f:657-661 / 672-677, then the scat_flag = 1 flight
__device__ __forceinline__ void child_flight(const Params &p, Photon &P, Rng &rng, int zi, double ewold, int i_gam)
{
    int dummy = 0;
    atomic_add_f64(p.edep + zi, P.ew - ewold);
    atomic_add_f64(p.E_IC + (i_gam - 1), P.ew - ewold);
    count(p, CNT_CMCOUNT2);
    if (P.phi > 2.0 * kPi) P.phi = P.phi - 2.0 * kPi;
    const double wt = 1.e-10 * P.ew;                               // wtmin of the new imctrk2d(1) frame
    const double colmfp = -log(rng.next(p));
    const int end = fly<MODE_CHILD>(p, P, rng, wt, colmfp, nullptr, dummy);
    if (end == END_COLLISION) {
        count(p, CNT_MULTIPLE); count(p, CNT_CMCOUNT1);
        push_collision(p, P, rng.sid);
    }
}

// one thread per (collision, child ii): the split2 loop of imctrk2d.f:619-689.  Children that trigger
// the 3rd split (l.639) are queued for k_split3 instead of looping split3 times in one thread.
// resident blocks per SM the team kernels are compiled for (registers = 65536 / (blocks x threads)); measured, see DESIGN.md
#ifndef C2D_TEAM_MINB
#define C2D_TEAM_MINB 1
#endif
#ifndef C2D_KN_MINB
#define C2D_KN_MINB 2
#endif
struct ChildCtx { const CollTask *task; int ii, zi; double ewold; uint64_t S_child; };
__device__ __forceinline__ void child_ctx(const Params &p, const S3Map &m, long long i, ChildCtx &c)
{
    const long long idx = (long long)m.rank + (long long)m.world * i;      // global child: this rank takes every world-th one
    const long long tg = idx / p.split2;
    int r = 0;
    while (r + 1 < m.world && tg >= m.prefix[r + 1]) r++;
    c.task = m.base + (size_t)r * m.stride + (size_t)(tg - m.prefix[r]);
    c.ii = (int)(idx % p.split2) + 1;
    c.ewold = c.task->p.ew / p.split2;                            // l.617
    c.zi = (c.task->p.jph - 1) * p.nr + (c.task->p.kph - 1);
    c.S_child = mix64(c.task->sid, (uint64_t)c.ii);
}
// after the electron is drawn: the rest of compb2d, the 3rd-split test (l.639), the scat_flag = 1 flight
__device__ __forceinline__ void child_finish(const Params &p, const ChildCtx &c, Rng &rng, const ElDraw &e, int i_gam)
{
    Photon P = c.task->p;
    P.ew = c.ewold;
    compb2d_finish(p, rng, P, e);
    if (P.ew > c.ewold * p.split2 * p.split1 * p.spl3_trg) {       // 3rd split, l.639-669
        const unsigned long long n = atomicAdd(p.q3_count, 1ull);
        if (n >= p.q_cap) { count(p, CNT_FATAL); return; }
        CollTask &t = p.q3[n];
        t.p = c.task->p; t.pad = 0; t.sid = c.S_child;
        return;
    }
    child_flight(p, P, rng, c.zi, c.ewold, i_gam);
}
__device__ __forceinline__ void children_one(const Params &p, const S3Map &m, long long i, unsigned *hard, unsigned long long hard_cap,
                                             unsigned long long *hard_count)
{
    ChildCtx c;
    child_ctx(p, m, i, c);
    Rng rng; rng.stream(c.S_child);
    const double znu = c.task->p.xnu / 5.11e2;
    ElDraw e; int i_gam = 0;
    // a hard photon's K-N rejection loop can run for 10^3 .. 10^6 trips: after kn_cap of them the child goes to a block team
    const bool team_ok = hard != nullptr && znu >= kKnTeamZnu;
    if (!compb2d_electron(p, rng, c.zi, znu, e, i_gam, nullptr, team_ok ? p.kn_cap : 0)) {
        const unsigned long long slot = atomicAdd(hard_count, 1ull);
        if (slot < hard_cap) {
            hard[slot] = (unsigned)i;
            count(p, CNT_TEAMS);
            return;
        }
        compb2d_electron(p, rng, c.zi, znu, e, i_gam, nullptr, 0);      // list full: carry on from trip kn_cap in this thread
    }
    child_finish(p, c, rng, e, i_gam);
}
// The collisions of all ranks are all-gathered and their children dealt out child by child (S3Map, as for the 3rd split):
// a collision of a hard photon deep in the Klein-Nishina regime costs 10^3-10^4 rejection-loop iterations per child, and
// which rank it happens on is luck.  A fixed grid strides over this rank's n_mine children.
__global__ void __launch_bounds__(128) k_children(const __grid_constant__ Params p, const S3Map m, long long n_mine, unsigned *hard,
                                                  unsigned long long hard_cap, unsigned long long *hard_count)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_mine; i += (long long)gridDim.x * blockDim.x)
        children_one(p, m, i, hard, hard_cap, hard_count);
}
#ifndef C2D_EMU
// One block per listed child: thread t of the team makes trip base + t of the child's electron loop (draws 5 k .. 5 k + 4 of its
// stream), the lowest accepted trip wins, the nth2d tallies of the trips up to it are committed -- exactly the draws the
// sequential loop makes.  Blocks fetch children from a device-side counter (the loops' lengths differ by orders of magnitude).
__global__ void __launch_bounds__(256, C2D_KN_MINB) k_children_hard(const __grid_constant__ Params p, const S3Map m, const unsigned *hard,
                                                       unsigned long long hard_cap, const unsigned long long *n_hard_dev,
                                                       unsigned long long *fetch, int max_rounds, int drain_rounds, GridItem *ultra,
                                                       unsigned long long *ultra_count)
{
    __shared__ int win, s_drain;
    __shared__ unsigned long long s_b;
    C2D_HIST_BEGIN(p)
    const unsigned long long n_hard = min(*n_hard_dev, hard_cap);
    for (;;) {
        if (threadIdx.x == 0) { s_b = atomicAdd(fetch, 1ull); s_drain = 0; }
        __syncthreads();
        const unsigned long long b = s_b;
        if (b >= n_hard) break;
        ChildCtx c;
        child_ctx(p, m, (long long)hard[b], c);
        const double znu = c.task->p.xnu / 5.11e2;
        Rng rng; ElDraw e; int i_gam = 0; ElRec rec;
        bool resolved = false;
        unsigned long long base = (unsigned long long)p.kn_cap;
        for (int round = 0; round < max_rounds; round++, base += blockDim.x) {
            if (threadIdx.x == 0) {
                win = 0x7fffffff;
                // half the blocks have found the list dry and this child is still unresolved after drain_rounds rounds: the GPU is emptying, so the
                // grid team is the faster way to finish it (while other blocks still have children to fetch, carrying on costs nothing)
                if (round >= drain_rounds && (round & 63) == 0) s_drain = *((volatile unsigned long long *)fetch) >= n_hard + gridDim.x / 2;
            }
            __syncthreads();
            if (s_drain) break;
            const unsigned long long k = base + threadIdx.x;
            rng.stream(c.S_child); rng.d = (uint32_t)(5ull * k);
            rec.bins = 0ull; rec.n = 0;
            const bool ok = compb2d_electron_once(p, rng, c.zi, znu, e, i_gam, &rec);
            if (ok) atomicMin(&win, (int)threadIdx.x);
            __syncthreads();
            const int first = win;
            if ((int)threadIdx.x <= first) tally_bin(p, hist, (unsigned)(rec.bins & 0xFFFFull), 1);
            if (first != 0x7fffffff) {
                if ((int)threadIdx.x == first) child_finish(p, c, rng, e, i_gam);
                resolved = true;
                break;
            }
            __syncthreads();                                      // everyone has read `win` before it is reset
        }
        // <sigma_KN / sigma_T> below ~1e-4 (a GeV photon meeting the gamma ~ 1e5 tail): the whole grid takes over at trip `base`
        if (!resolved && threadIdx.x == 0) { GridItem &g = ultra[atomicAdd(ultra_count, 1ull)]; g.first = base; g.idx = hard[b]; g.pad = 0u; }
        __syncthreads();
    }
    C2D_HIST_END(p)
}

// children the block teams gave up on: the grid is the team (see k_split3_ultra for the batching and the winner slots)
constexpr int kGridBatchMax = 64;
__global__ void __launch_bounds__(256, C2D_KN_MINB) k_children_ultra(const __grid_constant__ Params p, const S3Map m, const GridItem *ultra,
                                                        const unsigned long long *n_ultra_dev, unsigned long long *g_win)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    C2D_HIST_BEGIN(p)
    const unsigned long long n = *n_ultra_dev;
    const unsigned long long gsize = (unsigned long long)gridDim.x * blockDim.x, gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned it = 0;
    for (unsigned long long b = 0; b < n; b++) {
        ChildCtx c;
        child_ctx(p, m, (long long)ultra[b].idx, c);
        const double znu = c.task->p.xnu / 5.11e2;
        Rng rng; ElDraw e; int i_gam = 0; ElRec rec;
        int batch = 1;
        for (unsigned long long base = ultra[b].first;; base += gsize * batch, batch = min(2 * batch, kGridBatchMax), it++) {
            unsigned long long mine = ~0ull;
            int n_done = 0;
            for (int i = 0; i < batch; i++) {                    // trips base + gtid + i gsize, tallied as they are made
                const unsigned long long k = base + gtid + (unsigned long long)i * gsize;
                rng.stream(c.S_child); rng.d = (uint32_t)(5ull * k);
                rec.bins = 0ull; rec.n = 0;
                const bool ok = compb2d_electron_once(p, rng, c.zi, znu, e, i_gam, &rec);
                tally_bin(p, hist, (unsigned)(rec.bins & 0xFFFFull), 1);
                n_done++;
                if (ok) { mine = k; break; }
            }
            if (mine != ~0ull) atomicMin(g_win + (it % 3u), mine);
            grid.sync();
            const unsigned long long wn = *((volatile unsigned long long *)(g_win + (it % 3u)));
            if (gtid == 0) g_win[(it + 2u) % 3u] = ~0ull;
            if (wn != ~0ull) {
                for (int i = 0; i < n_done; i++) {                // the sequential loop never made the trips after the winner
                    const unsigned long long k = base + gtid + (unsigned long long)i * gsize;
                    if (k <= wn) continue;
                    Rng r2; ElDraw e2; int ig2; ElRec rec2;
                    r2.stream(c.S_child); r2.d = (uint32_t)(5ull * k);
                    rec2.bins = 0ull; rec2.n = 0;
                    compb2d_electron_once(p, r2, c.zi, znu, e2, ig2, &rec2);
                    tally_bin(p, hist, (unsigned)(rec2.bins & 0xFFFFull), -1);
                }
                if (mine == wn) child_finish(p, c, rng, e, i_gam);
                it++;
                break;
            }
        }
    }
    C2D_HIST_END(p)
}
#endif

// 3rd split, imctrk2d.f:641-669: every sub-child ii2 repeats compb2d until the scattered photon again exceeds
// the trigger.  In Philox replay mode each pass of that rejection loop owns the stream mix64(S_sub, attempt), so
// the passes can be evaluated in any order by a TEAM of threads: thread l of the team tries attempt base + l + 1,
// the lowest accepted attempt wins.  A pass whose electron cannot reach the trigger (gain <= (1+beta)^2 gamma^2
// < 4 gamma^2) is rejected right after the electron draw.  The nth2d tallies of the passes up to and including
// the winner are committed (the sequential loop made exactly those draws); later threads' are discarded.
//   k_split3_attempts  team = one warp per (3rd-split child, ii2), at most max_rounds rounds (0 = no limit); what is
//                      still unresolved (acceptance can be as low as 1e-6 per pass) is listed in `hard`
//   k_split3_hard      team = one block of 512 threads per listed sub-child, continuing at attempt `first_attempt`
struct Split3Sub {
    const CollTask *task;
    int zi;
    double ewcsv, ewold, trigger, gain_min, znu;
    uint64_t S_sub;
};
// Which sub-children a rank works on.  The 3rd-split tasks of all ranks are all-gathered (rank r's n_r tasks at
// base + r * stride) and dealt out again sub-child by sub-child: rank `rank` takes the global sub-children rank, rank + world,
// ... of the concatenated list -- a task's 300 sub-children share an acceptance rate that varies over six decades between
// tasks, so this is what evens out the ranks' tails.  Streams are keyed by ancestry: who computes a sub-child changes nothing.
__device__ __forceinline__ void split3_sub(const Params &p, const S3Map &m, long long i, Split3Sub &u)
{
    const long long w = (long long)m.rank + (long long)m.world * i;          // global sub-child
    const long long tg = w / p.split3;                                       // global task
    int r = 0;
    while (r + 1 < m.world && tg >= m.prefix[r + 1]) r++;
    u.task = m.base + (size_t)r * m.stride + (size_t)(tg - m.prefix[r]);
    const int ii2 = (int)(w % p.split3) + 1;
    u.ewcsv = u.task->p.ew / p.split2;
    u.ewold = u.ewcsv / p.split3;
    u.trigger = u.ewold * p.split2 * p.split1 * p.spl3_trg;
    u.gain_min = (double)p.split2 * p.split1 * p.spl3_trg;
    u.zi = (u.task->p.jph - 1) * p.nr + (u.task->p.kph - 1);
    u.znu = u.task->p.xnu / 5.11e2;
    u.S_sub = mix64(u.task->sid, (uint64_t)ii2);
}
// one pass of the rejection loop (imctrk2d.f:642-656)
__device__ __forceinline__ bool split3_pass(const Params &p, const Split3Sub &u, uint64_t attempt, Rng &rng, Photon &P,
                                            int &i_gam, ElRec &rec)
{
    rng.stream(mix64(u.S_sub, attempt));
    rec.bins = 0ull; rec.n = 0;
    ElDraw e;
    compb2d_electron(p, rng, u.zi, u.znu, e, i_gam, &rec);
    if (!(4.0 * e.gamm * e.gamm * (1.0 + 1.e-9) > u.gain_min)) return false;
    // (measured on the B200: moving the 2 % of the passes that get here out of line, warp-aggregating the tallies, or deciding
    // the K-N test of a hopeless trip from bounds on znue without its two cosine draws all make k_split3_attempts slower --
    // it runs at 32 registers and 64 warps / SM and is bound by latency, not by instruction count)
    P = u.task->p;
    P.ew = u.ewcsv / p.split3;
    if (!compb2d_finish(p, rng, P, e, u.trigger)) return false;
    return P.ew > u.trigger;
}
// nth2d tallies of a committed pass.  `hist` is the block's shared-memory copy of nelectron (flushed once per block:
// the passes hammer a handful of low-gamma bins, which global atomics would serialise), or null.
__device__ __forceinline__ void split3_commit(const Params &p, const Split3Sub &u, uint64_t attempt, const ElRec &rec, unsigned *hist)
{
    if (rec.n <= 4) {
        for (int k = 0; k < rec.n; k++) tally_bin(p, hist, (unsigned)((rec.bins >> (16 * k)) & 0xFFFFull), 1);
    } else {                                                      // > 4 K-N rejections in one pass: replay it, tallying
        Rng r2; r2.stream(mix64(u.S_sub, attempt));
        ElDraw e2; int ig2;
        compb2d_electron(p, r2, u.zi, u.znu, e2, ig2, nullptr);
    }
}
__device__ __forceinline__ void split3_store(FlightTask &f, const Photon &P, int i_gam, const Rng &rng, const Split3Sub &u)
{
    f.p = P; f.i_gam = i_gam; f.sid = rng.sid; f.d = rng.d; f.zi = u.zi; f.ewold = u.ewold;
}
constexpr int kHistWarp = 256;      // bins of the per-warp nelectron copy of k_split3_attempts
#ifndef C2D_S3_MINB
#define C2D_S3_MINB 16
#endif
__global__ void __launch_bounds__(128, C2D_S3_MINB) k_split3_attempts(const __grid_constant__ Params p, const S3Map m, long long i0, long long n_sub, FlightTask *out,
                                                         int max_rounds, unsigned *hard, unsigned long long *hard_count)
{
    // every warp keeps its own shared-memory copy of nelectron and flushes it by itself: no block barrier, so a warp
    // whose sub-child is accepted early retires without waiting for its neighbours
#ifdef C2D_EMU
    unsigned *hist = nullptr;
#else
    __shared__ unsigned s_hist[4][kHistWarp];
    unsigned *hist = (p.num_nt + 2 <= kHistWarp) ? s_hist[threadIdx.x >> 5] : nullptr;
    if (hist) for (int i = lane_id(); i < p.num_nt + 2; i += kWarp) hist[i] = 0u;
    __syncwarp();
#endif
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (w >= n_sub) return;                                       // whole warps together
    const int lane = lane_id();
    Split3Sub u;
    split3_sub(p, m, i0 + w, u);
    Rng rng; Photon P; int i_gam = 0; ElRec rec;
    uint64_t base = 0;
    bool done = false;
    for (int round = 0; max_rounds == 0 || round < max_rounds; round++, base += kWarp) {
        const uint64_t attempt = base + (uint64_t)lane + 1;
        const bool ok = split3_pass(p, u, attempt, rng, P, i_gam, rec);
        const unsigned m = warp_ballot(ok);
        const int first = m ? first_lane(m) : kWarp;
        if (lane <= first) split3_commit(p, u, attempt, rec, hist);
        if (m) {
            if (lane == first) { split3_store(out[w], P, i_gam, rng, u); count(p, CNT_S3PASSES, base + (uint64_t)first + 1); }
            done = true;
            break;
        }
    }
    if (!done && lane == 0) { hard[atomicAdd(hard_count, 1ull)] = (unsigned)w; count(p, CNT_S3HARD); }
#ifndef C2D_EMU
    __syncwarp();
    if (hist) for (int i = lane; i < p.num_nt + 2; i += kWarp)
        if (hist[i]) atomicAdd(p.nelectron + i, (unsigned long long)hist[i]);
#endif
}

#ifndef C2D_EMU
__global__ void __launch_bounds__(512, C2D_TEAM_MINB) k_split3_hard(const __grid_constant__ Params p, const S3Map m, long long i0, const unsigned *hard,
                                                     const unsigned long long *n_hard_dev, FlightTask *out, unsigned long long first_attempt,
                                                     int max_rounds, int drain_rounds, GridItem *ultra, unsigned long long *ultra_count,
                                                     unsigned long long *fetch)
{
    __shared__ int win, s_drain;
    __shared__ unsigned long long s_b;
    C2D_HIST_BEGIN(p)
    for (;;) {
        // blocks fetch sub-children from a device-side counter: their costs differ by orders of magnitude
        if (threadIdx.x == 0) { s_b = atomicAdd(fetch, 1ull); s_drain = 0; }
        __syncthreads();
        const unsigned long long b = s_b;
        const unsigned long long n_hard = *n_hard_dev;            // written by k_split3_attempts just before on this stream
        if (b >= n_hard) break;
        const long long w = hard[b];
        Split3Sub u;
        split3_sub(p, m, i0 + w, u);
        Rng rng; Photon P; int i_gam = 0; ElRec rec;
        bool resolved = false;
        uint64_t base = first_attempt;
        for (int round = 0; round < max_rounds; round++, base += blockDim.x) {
            if (threadIdx.x == 0) {
                win = 0x7fffffff;
                if (round >= drain_rounds && (round & 63) == 0) s_drain = *((volatile unsigned long long *)fetch) >= n_hard + gridDim.x / 2;
            }
            __syncthreads();
            if (s_drain) break;
            const uint64_t attempt = base + threadIdx.x + 1;
            const bool ok = split3_pass(p, u, attempt, rng, P, i_gam, rec);
            if (ok) atomicMin(&win, (int)threadIdx.x);
            __syncthreads();
            const int first = win;
            if ((int)threadIdx.x <= first) split3_commit(p, u, attempt, rec, hist);
            if (first != 0x7fffffff) {
                if ((int)threadIdx.x == first) { split3_store(out[w], P, i_gam, rng, u); count(p, CNT_S3PASSES, attempt); }
                resolved = true;
                break;
            }
            __syncthreads();                                      // everyone has read `win` before it is reset
        }
        // acceptance below ~1e-5 per pass: the whole grid takes over at attempt `base` (k_split3_ultra)
        if (!resolved && threadIdx.x == 0) {
            GridItem &g = ultra[atomicAdd(ultra_count, 1ull)];
            g.first = base; g.idx = (unsigned)w; g.pad = 0u;
            count(p, CNT_TEAMS);
        }
        __syncthreads();
    }
    C2D_HIST_END(p)
}

// The last rung: sub-children still unresolved when the block teams hand over (acceptance 1e-5 .. 1e-9 per pass: a child that
// barely triggered the 3rd split, imctrk2d.f:639, must repeat the feat for each of its split3 sub-children).  The reference
// grinds through them one pass at a time; here the WHOLE GRID is one team.  Super-round j: thread g makes the passes
// base + 1 + g + i G (i < batch_j = 1, 2, 4 .. 64; G = threads of the grid) up to its first accepted one, tallying their nth2d
// draws as it goes; the lowest accepted pass of the grid wins (atomicMin in global memory, one grid-wide barrier per
// super-round); the threads then take back the tallies of the passes they made beyond the winner -- the sequential loop never
// made those.  The doubling batch keeps the barrier (5 - 7 us) off the critical path of the really hard sub-children and
// wastes at most one batch on the easy ones.  Three winner slots rotate so that one barrier per super-round suffices: slot
// (it + 2) % 3 is reset after the barrier of super-round it, i.e. after its last reader and one barrier before its next writer.
__device__ __forceinline__ void split3_tally(const Params &p, const Split3Sub &u, uint64_t attempt, unsigned *hist, int sign)
{   // the nth2d draws of one pass (its K-N rejection loop, label 100), replayed
    Rng r; r.stream(mix64(u.S_sub, attempt));
    for (;;) {
        ElDraw e; int ig; ElRec rec; rec.bins = 0ull; rec.n = 0;
        const bool ok = compb2d_electron_once(p, r, u.zi, u.znu, e, ig, &rec);
        tally_bin(p, hist, (unsigned)(rec.bins & 0xFFFFull), sign);
        if (ok) break;
    }
}
__global__ void __launch_bounds__(512, C2D_TEAM_MINB) k_split3_ultra(const __grid_constant__ Params p, const S3Map m, long long i0, const GridItem *ultra,
                                                      const unsigned long long *n_ultra_dev, FlightTask *out, unsigned long long *g_win)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    C2D_HIST_BEGIN(p)
    const unsigned long long n = *n_ultra_dev;
    const unsigned long long gsize = (unsigned long long)gridDim.x * blockDim.x, gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned it = 0;
    for (unsigned long long b = 0; b < n; b++) {
        const long long w = ultra[b].idx;
        Split3Sub u;
        split3_sub(p, m, i0 + w, u);
        Rng rng; Photon P; int i_gam = 0; ElRec rec;
        int batch = 1;
        for (unsigned long long base = ultra[b].first;; base += gsize * batch, batch = min(2 * batch, kGridBatchMax), it++) {
            unsigned long long mine = ~0ull;
            int n_done = 0;
            for (int i = 0; i < batch; i++) {
                const unsigned long long attempt = base + gtid + 1 + (unsigned long long)i * gsize;
                const bool ok = split3_pass(p, u, attempt, rng, P, i_gam, rec);
                split3_commit(p, u, attempt, rec, hist);
                n_done++;
                if (ok) { mine = attempt; break; }
            }
            if (mine != ~0ull) atomicMin(g_win + (it % 3u), mine);
            grid.sync();
            const unsigned long long wn = *((volatile unsigned long long *)(g_win + (it % 3u)));
            if (gtid == 0) g_win[(it + 2u) % 3u] = ~0ull;
            if (wn != ~0ull) {
                for (int i = 0; i < n_done; i++) {
                    const unsigned long long attempt = base + gtid + 1 + (unsigned long long)i * gsize;
                    if (attempt > wn) split3_tally(p, u, attempt, hist, -1);
                }
                if (mine == wn) { split3_store(out[w], P, i_gam, rng, u); count(p, CNT_S3PASSES, wn); }
                it++;
                break;
            }
        }
    }
    C2D_HIST_END(p)
}
#endif

// one thread per accepted 3rd-split sub-child: tallies of l.657-661, then the scat_flag = 1 flight
__global__ void __launch_bounds__(128, 8) k_split3_flights(const __grid_constant__ Params p, const FlightTask *fl, long long n)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const FlightTask &f = fl[idx];
    Photon P = f.p;
    Rng rng; rng.stream(f.sid); rng.d = f.d;
    child_flight(p, P, rng, f.zi, f.ewold, f.i_gam);
}

// ------------------------------------------------------------------------------------------ launchers
#ifndef C2D_EMU
void launch_history_volume(const Params &p, long long first, long long n, long long stride, cudaStream_t st)
{
    if (n <= 0) return;
    const int threads = C2D_HIST_THREADS;
    const long long blocks = (n + threads - 1) / threads;
    k_history<0><<<(unsigned)blocks, threads, 0, st>>>(p, first, n, stride);
}
void launch_history_surface(const Params &p, long long first, long long n, long long stride, cudaStream_t st)
{
    if (n <= 0) return;
    const int threads = C2D_HIST_THREADS;
    const long long blocks = (n + threads - 1) / threads;
    k_history<2><<<(unsigned)blocks, threads, 0, st>>>(p, first, n, stride);
}
void launch_history_census(const Params &p, long long first, long long n, long long stride, cudaStream_t st)
{
    if (n <= 0) return;
    const int threads = C2D_HIST_THREADS;
    const long long blocks = (n + threads - 1) / threads;
    k_history<1><<<(unsigned)blocks, threads, 0, st>>>(p, first, n, stride);
}
void launch_absdraw(const Params &p, const AbsTask *q, const unsigned long long *n_dev, int n_sm, cudaStream_t st)
{
    k_absdraw<<<n_sm * 8, 128, 0, st>>>(p, q, n_dev);
}
static int coop_blocks_per_sm(const void *kernel, int threads)
{
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, 0) != cudaSuccess || b < 1) b = 1;
    return b;
}
// blocks of a grid team: all that fit (cooperative launch), or C2D_GRID_TEAM_BLOCKS (tests: a small team needs many super-rounds)
static unsigned coop_grid(int n_sm, int per_sm)
{
    if (const char *e = getenv("C2D_GRID_TEAM_BLOCKS")) return (unsigned)std::min(std::max(1, atoi(e)), n_sm * per_sm);
    return (unsigned)(n_sm * per_sm);
}
__global__ void k_selftest_gap(const __grid_constant__ Params p, long long n, unsigned long long seed, unsigned long long *out)
{
    unsigned long long bad = 0, slow = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t o[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)(i >> 32), p.key0, p.key1, o);
        for (int h = 0; h < 2; h++) {
            const double U = h ? u52(o[2], o[3]) : u52(o[0], o[1]);
            const double exact = floor(log(U) * kInvLnQ);
            if (trial_gap(U) != exact) bad++;
            const float tf = __logf((float)U) * (float)kInvLnQ, fr = tf - floorf(tf);
            if (!(fr > 5.e-3f && fr < 1.f - 5.e-3f)) slow++;
        }
    }
    if (bad) atomicAdd(out, bad);
    if (slow) atomicAdd(out + 1, slow);
}
void launch_selftest_gap(const Params &p, long long n, unsigned long long seed, unsigned long long *out, int n_sm, cudaStream_t st)
{
    k_selftest_gap<<<n_sm * 8, 256, 0, st>>>(p, n, seed, out);
}
void launch_children(const Params &p, const S3Map &m, long long n_mine, int n_sm, unsigned *hard, unsigned long long hard_cap,
                     unsigned long long *hard_count, cudaStream_t st)
{
    if (n_mine <= 0) return;
    const long long blocks = std::min<long long>((n_mine + 127) / 128, (long long)n_sm * 8);
    k_children<<<(unsigned)blocks, 128, 0, st>>>(p, m, n_mine, hard, hard_cap, hard_count);
}
// counts: [0] hard list length (k_children), [1] fetch cursor, [2] grid-team list length; all zero before k_children
void launch_children_hard(const Params &p, const S3Map &m, const unsigned *hard, unsigned long long hard_cap, GridItem *ultra,
                          unsigned long long *counts, int max_rounds, int drain_rounds, int n_sm, cudaStream_t st)
{
    k_children_hard<<<n_sm * 2 * C2D_KN_MINB, 256, 0, st>>>(p, m, hard, hard_cap, counts, counts + 1, max_rounds, drain_rounds, ultra, counts + 2);
}
int launch_children_ultra(const Params &p, const S3Map &m, const GridItem *ultra, const unsigned long long *n_ultra, unsigned long long *g_win,
                          int n_sm, cudaStream_t st)
{
    static const int per_sm = coop_blocks_per_sm((const void *)k_children_ultra, 256);
    void *args[] = {(void *)&p, (void *)&m, (void *)&ultra, (void *)&n_ultra, (void *)&g_win};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_children_ultra, dim3(coop_grid(n_sm, per_sm)), dim3(256), args, 0, st);
    return e == cudaSuccess ? 0 : (int)e;
}
void launch_split3_attempts(const Params &p, const S3Map &m, long long i0, long long n_sub, FlightTask *out, int max_rounds,
                            unsigned *hard, unsigned long long *hard_count, cudaStream_t st)
{
    if (n_sub <= 0) return;
    const long long n = n_sub * kWarp;                            // one warp per sub-child
    k_split3_attempts<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, m, i0, n_sub, out, max_rounds, hard, hard_count);
}
// counts: [0] hard list length (written by k_split3_attempts), [1] fetch cursor, [2] grid-team list length
void launch_split3_hard(const Params &p, const S3Map &m, long long i0, const unsigned *hard, GridItem *ultra, unsigned long long *counts,
                        FlightTask *out, unsigned long long first_attempt, int max_rounds, int drain_rounds, int n_sm, cudaStream_t st)
{
    k_split3_hard<<<n_sm * 2 * C2D_TEAM_MINB, 512, 0, st>>>(p, m, i0, hard, counts, out, first_attempt, max_rounds < 0 ? 0x7fffffff : max_rounds, drain_rounds, ultra,
                                            counts + 2, counts + 1);
}
// g_win: 3 slots = ~0
int launch_split3_ultra(const Params &p, const S3Map &m, long long i0, const GridItem *ultra, const unsigned long long *n_ultra, FlightTask *out,
                        unsigned long long *g_win, int n_sm, cudaStream_t st)
{
    static const int per_sm = coop_blocks_per_sm((const void *)k_split3_ultra, 512);
    void *args[] = {(void *)&p, (void *)&m, (void *)&i0, (void *)&ultra, (void *)&n_ultra, (void *)&out, (void *)&g_win};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_split3_ultra, dim3(coop_grid(n_sm, per_sm)), dim3(512), args, 0, st);
    return e == cudaSuccess ? 0 : (int)e;
}
void launch_split3_flights(const Params &p, const FlightTask *fl, long long n, cudaStream_t st)
{
    if (n <= 0) return;
    k_split3_flights<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, fl, n);
}
#endif  // C2D_EMU

}  // namespace c2d
