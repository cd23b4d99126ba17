// c2d_lc.cu -- event-list reduction on the device.
//
// The reference writes every escaping photon as a 105-byte text line (imcleak2d.f:171,181) and turns the files into
// science products afterwards with two C programs that share one per-photon transform and differ in their bin rules:
//   postprocessing/plcm.c:384-450  light curves:  L(t_obs, mu_obs, band), <ew>, <ew^2> per bin
//   postprocessing/pspt.c:249-289  SEDs:          L(E_obs, t_obs) in one viewing-angle window
// Here the escape records never leave HBM: k_lc_reduce streams the step's records (7 doubles each, the layout
// imcleak writes), applies the transform -- Doppler boost with the bulk Lorentz factor, light-travel-time correction to the
// plane through r = rmax -- and bins in arrival time, observer-frame direction cosine and photon energy with the bin
// tables of c2d_lc_config (either program's rules).  The accumulators persist across steps, as the files do.
//
// k_lc_reduce: one warp = 32 consecutive records (1792 contiguous bytes, fetched as 112 16-byte loads, transposed through
// shared memory so every lane gets its own record); the next tile is in flight while the current one is processed.
// A cheap filter (observer-frame mu and E against the bins) runs on every record; the survivors are compacted per warp
// and only they pay for the arrival time (sqrt, cos, division) and the time-bin search.
// Commits are reduced warp-level (__match_any_sync on the cell index, lane-ordered sums) and go to a per-block copy of the
// accumulators in shared memory when the bin table fits (flushed once per block), else straight to HBM.
// Algorithmic bytes: 56 per record (the read; the accumulators are a few kB).
#include "c2d_ctx.h"

#include <algorithm>
#include <vector>

namespace c2d {

struct LcParams {
    double gam_bulk, betta, rmax, t_offset;
    int n_t, n_mu, n_bands, mu_closed, first_band_only;
    const double *t_lo, *t_hi, *mu0, *mu1, *E0, *E1;     // device copies of the bin tables
    int n_edge;                                          // sorted distinct band edges; interval i = [edge[i], edge[i+1])
    const double *edge, *covered;                        // covered[i] != 0: some band holds interval i
    double *sum_ew, *sum_ew2;                            // [n_t][n_mu][n_bands]
    unsigned long long *count;
};

struct LcState {
    LcParams q{};
    double *d_tab = nullptr;       // [t_lo | t_hi | mu0 | mu1 | E0 | E1 | edge | covered]
    double *d_acc = nullptr;       // [sum_ew | sum_ew2 | count (u64)]
    size_t cells = 0;
};

// plcm.c:393-401 == pspt.c:257-264, split in two so that a record outside every mu bin or every band (most records, for
// the narrow viewing-angle windows the programs are run with) skips the expensive half.  Same operations, same order.
// (1) observer-frame E, ew, mu; returns the Doppler factor
__host__ __device__ __forceinline__ double lc_boost(const LcParams &q, double &E, double &ew, double &mu)
{
    mu = -mu;
    const double doppler = q.gam_bulk * (1. + mu * q.betta);
    E = E * doppler;
    ew = ew * doppler;
    mu = (mu + q.betta) / (1. + mu * q.betta);
    return doppler;
}
// (2) observer-frame arrival time (mu already boosted), then `time -= t_offset` (plcm.c:412)
__host__ __device__ __forceinline__ double lc_arrival(const LcParams &q, double doppler, double t_bound, double r, double z,
                                                      double mu, double phi)
{
    t_bound = (t_bound - q.betta * z * 3.33333333e-11) / doppler;
    const double cdt = z * mu / q.gam_bulk + sqrt(1. - mu * mu) * (q.rmax - r * cos(phi));
    double time = t_bound + 3.33333333e-11 * cdt;
    time -= q.t_offset;
    return time;
}

// first k with t_lo[k] <= time < t_hi[k] (plcm.c:416-420, pspt.c:282-284); both tables are nondecreasing (checked by
// c2d_lc_begin), so the first k with t_hi[k] > time is the only candidate
__host__ __device__ __forceinline__ int lc_time_bin(const double *t_lo, const double *t_hi, int n_t, double time)
{
    int lo = 0, hi = n_t;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t_hi[mid] > time) hi = mid; else lo = mid + 1;
    }
    return (lo < n_t && time >= t_lo[lo]) ? lo : -1;
}

__host__ __device__ __forceinline__ int lc_mu_bin(const double *mu0, const double *mu1, int n_mu, int closed, double mu)
{
    for (int n = 0; n < n_mu; n++)
        if (mu >= mu0[n] && (closed ? mu <= mu1[n] : mu < mu1[n])) return n;      // plcm.c:422-426 / pspt.c:274
    return -1;
}

#ifndef C2D_EMU
constexpr int kLcWarps = 8;                     // warps per block
constexpr int kLcTileDoubles = 32 * 7;          // one warp tile: 32 records

// 16-byte chunk c of the record array (n7 = 7 n doubles in total); the odd tail double is fetched alone
__device__ __forceinline__ double2 lc_load_chunk(const double *ev, long long c, long long n7)
{
    const long long d = 2 * c;
    if (d + 1 < n7) return __ldcs(reinterpret_cast<const double2 *>(ev) + c);
    double2 v = make_double2(0., 0.);
    if (d < n7) v.x = __ldcs(ev + d);
    return v;
}

// warp-level pre-reduction: lanes that hit the same cell are summed in lane order (deterministic within the warp) and the
// lowest of them commits.  cell < 0: nothing to commit
template <bool SHARED>
__device__ __forceinline__ void lc_commit(double *s1, double *s2, unsigned long long *cnt, int cell, double ew)
{
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int lane = threadIdx.x & 31;
    const bool alone = peers == (1u << lane);
    double a = ew, b = ew * ew;
    const unsigned crowded = __ballot_sync(0xffffffffu, !alone && cell >= 0);
    if (crowded) {
        double sa = 0., sb = 0.;
        for (unsigned m = crowded; m; m &= m - 1) {                               // only lanes that share a cell are broadcast
            const int src = __ffs(m) - 1;
            const double v = __shfl_sync(0xffffffffu, ew, src);
            if ((peers >> src) & 1u) { sa = sa + v; sb = sb + v * v; }
        }
        if (!alone) { a = sa; b = sb; }
    }
    if (cell >= 0 && lane == __ffs(peers) - 1) {
        atomicAdd(s1 + cell, a);
        atomicAdd(s2 + cell, b);
        if (SHARED) atomicAdd(reinterpret_cast<unsigned *>(cnt) + cell, (unsigned)__popc(peers));   // 32-bit per-block counts
        else atomicAdd(cnt + cell, (unsigned long long)__popc(peers));
    }
}

// does any band hold E?  (the filter stage only needs yes / no; the commit stage walks the bands.)  The band edges cut the
// energy axis into intervals that are either inside some band or inside none: one branch-free bisection over the sorted
// distinct edges (built by c2d_lc_begin) instead of a divergent scan of the bands.  Same answer as the scan for every E.
__device__ __forceinline__ bool lc_any_band(const double *edge, const double *covered, int n_edge, double E)
{
    int lo = 0, hi = n_edge;                                  // first index with edge > E
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edge[mid] > E) hi = mid; else lo = mid + 1;
    }
    return lo > 0 && lo < n_edge && covered[lo - 1] != 0.0;   // E in [edge[lo-1], edge[lo])
}

// The full path for one (compacted) tile of up to 32 records in shared memory: transform, bins, commits.
template <bool SHARED>
__device__ __forceinline__ void lc_process(const LcParams &q, const double *tile, int cnt, const double *s_tlo, const double *s_thi,
                                           const double *s_mu0, const double *s_mu1, const double *s_E0, const double *s_E1,
                                           double *a1, double *a2, unsigned long long *ac)
{
    const int lane = threadIdx.x & 31;
    const double *rec = tile + 7 * lane;
    bool live = lane < cnt;
    int base = -1;
    double E = 0., ew = 0.;
    if (live) {
        double mu = rec[5];
        E = rec[1]; ew = rec[2];
        const double doppler = lc_boost(q, E, ew, mu);
        const int m = lc_mu_bin(s_mu0, s_mu1, q.n_mu, q.mu_closed, mu);
        live = m >= 0;
        if (live) {
            const double time = lc_arrival(q, doppler, rec[0], rec[3], rec[4], mu, rec[6]);
            const int k = lc_time_bin(s_tlo, s_thi, q.n_t, time);
            live = k >= 0;
            base = (k * q.n_mu + m) * q.n_bands;
        }
    }
    // plcm.c:441-449: every band that holds E; pspt.c:278-280: the first one only
    if (!__any_sync(0xffffffffu, live)) return;
    for (int l = 0; l < q.n_bands; l++) {
        const bool hit = live && E >= s_E0[l] && E < s_E1[l];
        if (!__any_sync(0xffffffffu, hit)) continue;
        lc_commit<SHARED>(a1, a2, ac, hit ? base + l : -1, ew);
        if (hit && q.first_band_only) live = false;
        if (!__any_sync(0xffffffffu, live)) break;
    }
}

// Two stages per warp.  FILTER: every record of a streamed tile gets the cheap half of the transform (observer-frame mu
// and E: one division) and is tested against the mu bins and the bands; the programs are run with narrow viewing-angle
// windows and sparse bands, so few records survive (3 % with the reference's mrk421_lc.input).  Survivors are compacted
// into a per-warp queue in shared memory; COMMIT: whenever 32 are queued they take the full path (arrival time: sqrt, cos,
// division; time-bin search; warp-reduced commits) at full lane occupancy.
template <bool SHARED>
__global__ void __launch_bounds__(kLcWarps * 32) k_lc_reduce(const __grid_constant__ LcParams q, const double *__restrict__ ev,
                                                            const unsigned long long *n_ptr, long long n_cap)
{
    extern __shared__ double sh[];
    // shared layout: [tile + 2-tile queue per warp | t_lo | t_hi | mu0 | mu1 | E0 | E1 | edge | covered | (sum_ew | sum_ew2 | count32)]
    double *tile = sh + (threadIdx.x >> 5) * (3 * kLcTileDoubles), *queue = tile + kLcTileDoubles;
    double *s_tlo = sh + kLcWarps * 3 * kLcTileDoubles, *s_thi = s_tlo + q.n_t;
    double *s_mu0 = s_thi + q.n_t, *s_mu1 = s_mu0 + q.n_mu, *s_E0 = s_mu1 + q.n_mu, *s_E1 = s_E0 + q.n_bands;
    double *s_edge = s_E1 + q.n_bands, *s_cov = s_edge + q.n_edge;
    const int cells = q.n_t * q.n_mu * q.n_bands;
    double *a1 = SHARED ? s_cov + q.n_edge : q.sum_ew, *a2 = SHARED ? a1 + cells : q.sum_ew2;
    unsigned long long *ac = SHARED ? reinterpret_cast<unsigned long long *>(a2 + cells) : q.count;
    for (int i = threadIdx.x; i < q.n_t; i += blockDim.x) { s_tlo[i] = q.t_lo[i]; s_thi[i] = q.t_hi[i]; }
    for (int i = threadIdx.x; i < q.n_mu; i += blockDim.x) { s_mu0[i] = q.mu0[i]; s_mu1[i] = q.mu1[i]; }
    for (int i = threadIdx.x; i < q.n_bands; i += blockDim.x) { s_E0[i] = q.E0[i]; s_E1[i] = q.E1[i]; }
    for (int i = threadIdx.x; i < q.n_edge; i += blockDim.x) { s_edge[i] = q.edge[i]; s_cov[i] = q.covered[i]; }
    if (SHARED) {
        for (int i = threadIdx.x; i < 2 * cells; i += blockDim.x) a1[i] = 0.;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) reinterpret_cast<unsigned *>(ac)[i] = 0u;
    }
    __syncthreads();

    long long n = (long long)*n_ptr;
    if (n > n_cap) n = n_cap;
    const long long n7 = 7 * n, n_tiles = (n + 31) / 32, n_full = n / 32;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kLcWarps + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * kLcWarps;
    double2 r0, r1, r2, r3;
    r3 = make_double2(0., 0.);
    auto fetch = [&](long long t) {                                              // 112 chunks of 16 bytes per tile
        const long long c = t * 112 + lane;
        if (t < n_full) {
            const double2 *e2 = reinterpret_cast<const double2 *>(ev) + c;
            r0 = __ldcs(e2); r1 = __ldcs(e2 + 32); r2 = __ldcs(e2 + 64);
            if (lane < 16) r3 = __ldcs(e2 + 96);
        } else {                                                                 // the last, partial tile
            r0 = lc_load_chunk(ev, c, n7); r1 = lc_load_chunk(ev, c + 32, n7); r2 = lc_load_chunk(ev, c + 64, n7);
            if (lane < 16) r3 = lc_load_chunk(ev, c + 96, n7);
        }
    };
    int qn = 0;                                                                  // queued survivors (warp-uniform)
    if (warp0 < n_tiles) fetch(warp0);
    for (long long t = warp0; t < n_tiles; t += n_warps) {
        double2 *t2 = reinterpret_cast<double2 *>(tile);
        t2[lane] = r0; t2[lane + 32] = r1; t2[lane + 64] = r2;
        if (lane < 16) t2[lane + 96] = r3;
        __syncwarp();
        if (t + n_warps < n_tiles) fetch(t + n_warps);                           // next tile in flight during the arithmetic
        const double *rec = tile + 7 * lane;
        bool live = t * 32 + lane < n;
        if (live) {
            double E = rec[1], ew = 0., mu = rec[5];
            lc_boost(q, E, ew, mu);
            live = lc_mu_bin(s_mu0, s_mu1, q.n_mu, q.mu_closed, mu) >= 0 && lc_any_band(s_edge, s_cov, q.n_edge, E);
        }
        const unsigned pass = __ballot_sync(0xffffffffu, live);
        if (pass) {
            if (live) {
                double *dst = queue + 7 * (qn + __popc(pass & ((1u << lane) - 1u)));
#pragma unroll
                for (int f = 0; f < 7; f++) dst[f] = rec[f];
            }
            qn += __popc(pass);
        }
        __syncwarp();
        if (qn >= 32) {
            lc_process<SHARED>(q, queue, 32, s_tlo, s_thi, s_mu0, s_mu1, s_E0, s_E1, a1, a2, ac);
            __syncwarp();
            qn -= 32;
            double keep[7];
            if (lane < qn) {
#pragma unroll
                for (int f = 0; f < 7; f++) keep[f] = queue[7 * (32 + lane) + f];
            }
            __syncwarp();
            if (lane < qn) {
#pragma unroll
                for (int f = 0; f < 7; f++) queue[7 * lane + f] = keep[f];
            }
            __syncwarp();
        }
    }
    if (qn > 0) lc_process<SHARED>(q, queue, qn, s_tlo, s_thi, s_mu0, s_mu1, s_E0, s_E1, a1, a2, ac);
    if (SHARED) {
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += blockDim.x) {
            const unsigned c = reinterpret_cast<unsigned *>(ac)[i];
            if (c) { atomicAdd(q.sum_ew + i, a1[i]); atomicAdd(q.sum_ew2 + i, a2[i]); atomicAdd(q.count + i, (unsigned long long)c); }
        }
    }
}

// synthetic records for c2d_lc_bench: escape records of a 1e16 cm blob over ~30 light-crossing times, from a fixed hash
__global__ void k_lc_gen(long long n, double *ev)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
        auto u = [&]() { h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; return (double)(h >> 11) * (1.0 / 9007199254740992.0); };
        double *e = ev + 7 * i;
        e[0] = 1.0e7 * u();                         // t_bound [s]
        e[1] = exp(log(1e-7) + log(1e15) * u());    // xnu [keV]
        e[2] = 1.0e38 * (0.5 + u());                // ew [erg]
        e[3] = 0.75e16 * sqrt(u());                 // rpre
        e[4] = 1.0e16 * u();                        // zpre
        e[5] = -1.0 + 2.0 * u();                    // wmu
        e[6] = 3.1415926536 * u();                  // phi
    }
}
#endif  // !C2D_EMU

static int lc_launch(Ctx *c, LcState *s, const double *d_ev, const unsigned long long *d_n, long long n_cap, float *ms)
{
#ifdef C2D_EMU
    (void)ms;
    const LcParams &q = s->q;
    long long n = std::min<long long>((long long)*d_n, n_cap);
    for (long long i = 0; i < n; i++) {
        const double *rec = d_ev + 7 * i;
        double E = rec[1], ew = rec[2], mu = rec[5];
        const double doppler = lc_boost(q, E, ew, mu);
        const int m = lc_mu_bin(q.mu0, q.mu1, q.n_mu, q.mu_closed, mu);
        if (m < 0) continue;
        const double time = lc_arrival(q, doppler, rec[0], rec[3], rec[4], mu, rec[6]);
        const int k = lc_time_bin(q.t_lo, q.t_hi, q.n_t, time);
        if (k < 0) continue;
        for (int l = 0; l < q.n_bands; l++)
            if (E >= q.E0[l] && E < q.E1[l]) {
                const size_t at = ((size_t)k * q.n_mu + m) * q.n_bands + l;
                q.sum_ew[at] += ew; q.sum_ew2[at] += ew * ew; q.count[at]++;
                if (q.first_band_only) break;
            }
    }
    return 0;
#else
    const LcParams &q = s->q;
    const size_t tab = (size_t)kLcWarps * 3 * kLcTileDoubles + 2 * (size_t)q.n_t + 2 * (size_t)q.n_mu + 2 * (size_t)q.n_bands +
                       2 * (size_t)q.n_edge;
    const size_t priv = 2 * s->cells + (s->cells + 1) / 2;
    const bool shared = (tab + priv) * 8 <= 100 * 1024;                          // 2 blocks / SM with the private copy
    const size_t smem = (tab + (shared ? priv : 0)) * 8;
    if (smem > 200 * 1024) return fail(C2D_EINVAL, "light-curve bin tables need %zu bytes of shared memory", smem);
    auto kern = shared ? k_lc_reduce<true> : k_lc_reduce<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(C2D_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLcWarps * 32, smem);
    const int grid = c->n_sm * std::max(1, per_sm);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    kern<<<grid, kLcWarps * 32, smem, c->stream>>>(q, d_ev, d_n, n_cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(C2D_ECUDA, "k_lc_reduce launch: %s", cudaGetErrorString(e));
    if (ms) {
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return 0;
#endif
}

void Ctx::lc_free(int slot)
{
    LcState *&s = lc_slot[slot];
    if (!s) return;
    cudaFree(s->d_tab); cudaFree(s->d_acc);
    delete s;
    s = nullptr;
}
void Ctx::lc_free_all()
{
    for (int k = 0; k < kLcSlots; k++) lc_free(k);
    cudaFree(d_lc_stage); d_lc_stage = nullptr; lc_stage_cap = 0;
}
static bool lc_any(const Ctx *c)
{
    for (int k = 0; k < Ctx::kLcSlots; k++) if (c->lc_slot[k]) return true;
    return false;
}

}  // namespace c2d

using namespace c2d;

#define CUL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(C2D_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" int c2d_lc_begin(c2d_ctx *h, const c2d_lc_config *cfg)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (!cfg) return fail(C2D_EINVAL, "NULL config");
    if (cfg->n_t < 1 || cfg->n_mu < 1 || cfg->n_bands < 1 || !cfg->t_lo || !cfg->t_hi || !cfg->mu0 || !cfg->mu1 || !cfg->E0 || !cfg->E1)
        return fail(C2D_EINVAL, "light-curve bin tables: n_t %d n_mu %d n_bands %d or a NULL table", cfg->n_t, cfg->n_mu, cfg->n_bands);
    if (!(cfg->gam_bulk >= 1.0)) return fail(C2D_EINVAL, "gam_bulk = %g", cfg->gam_bulk);
    for (int k = 1; k < cfg->n_t; k++)
        if (!(cfg->t_lo[k] >= cfg->t_lo[k - 1]) || !(cfg->t_hi[k] >= cfg->t_hi[k - 1]))
            return fail(C2D_EINVAL, "time-bin edges must be nondecreasing (bin %d)", k);
    const size_t cells = (size_t)cfg->n_t * cfg->n_mu * cfg->n_bands;
    if (cells > ((size_t)1 << 28)) return fail(C2D_EINVAL, "%zu accumulator cells", cells);
    CUL(cudaSetDevice(c->device));
    c->lc_free(c->lc_active);
    LcState *s = c->lc_cur() = new LcState();
    s->cells = cells;
    // the band edges in ascending order and, per interval between neighbours, whether a band holds it (NaN edges: the
    // band can hold nothing and is left out)
    std::vector<double> edge;
    for (int l = 0; l < cfg->n_bands; l++)
        if (cfg->E0[l] < cfg->E1[l]) { edge.push_back(cfg->E0[l]); edge.push_back(cfg->E1[l]); }
    std::sort(edge.begin(), edge.end());
    edge.erase(std::unique(edge.begin(), edge.end()), edge.end());
    if (edge.empty()) edge.push_back(0.0);
    std::vector<double> covered(edge.size(), 0.0);
    for (size_t i = 0; i + 1 < edge.size(); i++)
        for (int l = 0; l < cfg->n_bands; l++)
            if (cfg->E0[l] <= edge[i] && edge[i] < cfg->E1[l]) { covered[i] = 1.0; break; }
    const size_t nt = cfg->n_t, nm = cfg->n_mu, nb = cfg->n_bands, ne = edge.size(), ntab = 2 * (nt + nm + nb + ne);
    std::vector<double> tab(ntab);
    double *p = tab.data();
    std::copy(cfg->t_lo, cfg->t_lo + nt, p); p += nt; std::copy(cfg->t_hi, cfg->t_hi + nt, p); p += nt;
    std::copy(cfg->mu0, cfg->mu0 + nm, p); p += nm; std::copy(cfg->mu1, cfg->mu1 + nm, p); p += nm;
    std::copy(cfg->E0, cfg->E0 + nb, p); p += nb; std::copy(cfg->E1, cfg->E1 + nb, p); p += nb;
    std::copy(edge.begin(), edge.end(), p); p += ne; std::copy(covered.begin(), covered.end(), p);
    CUL(cudaMalloc(&s->d_tab, ntab * 8));
    CUL(cudaMalloc(&s->d_acc, 3 * cells * 8));
    CUL(cudaMemcpyAsync(s->d_tab, tab.data(), ntab * 8, cudaMemcpyHostToDevice, c->stream));
    CUL(cudaMemsetAsync(s->d_acc, 0, 3 * cells * 8, c->stream));
    CUL(cudaStreamSynchronize(c->stream));
    LcParams &q = s->q;
    q.gam_bulk = cfg->gam_bulk;
    q.betta = sqrt(1. - 1. / (cfg->gam_bulk * cfg->gam_bulk));                  // plcm.c:394 (the same for every photon)
    q.rmax = cfg->rmax; q.t_offset = cfg->t_offset;
    q.n_t = cfg->n_t; q.n_mu = cfg->n_mu; q.n_bands = cfg->n_bands;
    q.mu_closed = cfg->mu_closed ? 1 : 0; q.first_band_only = cfg->first_band_only ? 1 : 0;
    q.t_lo = s->d_tab; q.t_hi = q.t_lo + nt; q.mu0 = q.t_hi + nt; q.mu1 = q.mu0 + nm; q.E0 = q.mu1 + nm; q.E1 = q.E0 + nb;
    q.n_edge = (int)ne; q.edge = q.E1 + nb; q.covered = q.edge + ne;
    q.sum_ew = s->d_acc; q.sum_ew2 = s->d_acc + cells;
    q.count = reinterpret_cast<unsigned long long *>(s->d_acc + 2 * cells);
    return 0;
}

extern "C" int c2d_lc_select(c2d_ctx *h, int slot)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return fail(C2D_EINVAL, "NULL ctx");
    if (slot < 0 || slot >= Ctx::kLcSlots) return fail(C2D_EINVAL, "reducer %d (0 .. %d)", slot, Ctx::kLcSlots - 1);
    c->lc_active = slot;
    return 0;
}

extern "C" int c2d_lc_accumulate(c2d_ctx *h)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (!lc_any(c)) return fail(C2D_ESTATE, "c2d_lc_begin has not been called");
    if (int rc = c2d_sync(h)) return rc;                                        // the scattering queue must have drained
    if (int rc = check_events(c)) return rc;                                    // an overflowed list would silently drop photons
    // the record count lives on the device (counters[C2D_CNT_NEVENTS]); records beyond event_capacity were not stored
    for (int k = 0; k < Ctx::kLcSlots; k++)
        if (c->lc_slot[k])
            if (int rc = lc_launch(c, c->lc_slot[k], c->d_events, c->P.counters + CNT_NEVENTS, c->event_cap, nullptr)) return rc;
    return 0;
}

extern "C" int c2d_lc_accumulate_events(c2d_ctx *h, const c2d_event *ev, int64_t n)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (!lc_any(c)) return fail(C2D_ESTATE, "c2d_lc_begin has not been called");
    if (n < 0 || (n && !ev)) return fail(C2D_EINVAL, "n = %lld", (long long)n);
    if (n == 0) return 0;
    CUL(cudaSetDevice(c->device));
    const size_t need = (size_t)n * 7 + 1;                                      // + the count
    if (need > c->lc_stage_cap) {
        cudaFree(c->d_lc_stage); c->d_lc_stage = nullptr; c->lc_stage_cap = 0;
        CUL(cudaMalloc(&c->d_lc_stage, (need + 1) * 8));
        c->lc_stage_cap = need;
    }
    const unsigned long long cnt = (unsigned long long)n;
    double *d_cnt = c->d_lc_stage + (((size_t)n * 7 + 1) & ~(size_t)1);        // the 8-byte slot after the records
    CUL(cudaMemcpyAsync(c->d_lc_stage, ev, (size_t)n * 56, cudaMemcpyHostToDevice, c->stream));
    CUL(cudaMemcpyAsync(d_cnt, &cnt, 8, cudaMemcpyHostToDevice, c->stream));
    for (int k = 0; k < Ctx::kLcSlots; k++)
        if (c->lc_slot[k])
            if (int rc = lc_launch(c, c->lc_slot[k], c->d_lc_stage, reinterpret_cast<unsigned long long *>(d_cnt), n, nullptr)) return rc;
    CUL(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int c2d_lc_read(c2d_ctx *h, double *sum_ew, double *sum_ew2, int64_t *count)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->lc_cur()) return fail(C2D_ESTATE, "c2d_lc_begin has not been called for reducer %d", c ? c->lc_active : 0);
    CUL(cudaSetDevice(c->device));
    LcState *s = c->lc_cur();
    CUL(cudaStreamSynchronize(c->stream));
    if (sum_ew) CUL(cudaMemcpy(sum_ew, s->q.sum_ew, s->cells * 8, cudaMemcpyDeviceToHost));
    if (sum_ew2) CUL(cudaMemcpy(sum_ew2, s->q.sum_ew2, s->cells * 8, cudaMemcpyDeviceToHost));
    if (count) CUL(cudaMemcpy(count, s->q.count, s->cells * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int c2d_lc_device(c2d_ctx *h, double **sums, int64_t **count, int64_t *cells)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->lc_cur()) return fail(C2D_ESTATE, "c2d_lc_begin has not been called for reducer %d", c ? c->lc_active : 0);
    if (sums) *sums = c->lc_cur()->d_acc;
    if (count) *count = reinterpret_cast<int64_t *>(c->lc_cur()->q.count);
    if (cells) *cells = (int64_t)c->lc_cur()->cells;
    return 0;
}

extern "C" int c2d_lc_bench(c2d_ctx *h, int64_t n_events, float *ms, int64_t *bytes)
{
#ifdef C2D_EMU
    (void)h; (void)n_events; (void)ms; (void)bytes;
    return fail(C2D_EINVAL, "c2d_lc_bench needs the device build");
#else
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->lc_cur()) return fail(C2D_ESTATE, "c2d_lc_begin has not been called for reducer %d", c ? c->lc_active : 0);
    if (n_events < 1) return fail(C2D_EINVAL, "n_events = %lld", (long long)n_events);
    CUL(cudaSetDevice(c->device));
    double *d_ev = nullptr;
    CUL(cudaMalloc(&d_ev, ((size_t)n_events * 7 + 2) * 8));
    k_lc_gen<<<c->n_sm * 8, 256, 0, c->stream>>>(n_events, d_ev);
    CUL(cudaGetLastError());
    const unsigned long long cnt = (unsigned long long)n_events;
    double *d_cnt = d_ev + (((size_t)n_events * 7 + 1) & ~(size_t)1);
    CUL(cudaMemcpyAsync(d_cnt, &cnt, 8, cudaMemcpyHostToDevice, c->stream));
    float t = 0;
    int rc = lc_launch(c, c->lc_cur(), d_ev, reinterpret_cast<unsigned long long *>(d_cnt), n_events, &t);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_ev);
    if (rc) return rc;
    if (ms) *ms = t;
    if (bytes) *bytes = n_events * 56;
    return 0;
#endif
}
