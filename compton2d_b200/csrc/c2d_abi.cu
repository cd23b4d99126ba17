// c2d_abi.cu -- the C ABI of include/compton2d_b200.h: context, uploads, step driver, tallies, census I/O.
// Host code only; the kernels live in c2d_track.cu / c2d_fp.cu.  There is no CPU fallback: every
// entry point needs the CUDA device the context was created on.
#include "c2d_ctx.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace c2d {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(C2D_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static int upload(Ctx *c, double *dst, const double *src, size_t n, double fill = 0.0)
{
    if (!src) {
        std::vector<double> tmp(n, fill);
        CU(cudaMemcpyAsync(dst, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CU(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

static int alloc_bank(Bank &b, size_t cap)
{
    CU(cudaMalloc(&b.rpre, cap * 8)); CU(cudaMalloc(&b.zpre, cap * 8)); CU(cudaMalloc(&b.wmu, cap * 8));
    CU(cudaMalloc(&b.phi, cap * 8)); CU(cudaMalloc(&b.ew, cap * 8)); CU(cudaMalloc(&b.xnu, cap * 8));
    CU(cudaMalloc(&b.bins, cap * 4)); CU(cudaMalloc(&b.zone, cap * 4)); CU(cudaMalloc(&b.seed, cap * 4));
    CU(cudaMalloc(&b.sid, cap * 8));
    return 0;
}
static void free_bank(Bank &b)
{
    cudaFree(b.rpre); cudaFree(b.zpre); cudaFree(b.wmu); cudaFree(b.phi); cudaFree(b.ew); cudaFree(b.xnu);
    cudaFree(b.bins); cudaFree(b.zone); cudaFree(b.seed); cudaFree(b.sid);
    memset(&b, 0, sizeof b);
}

void Ctx::timer_begin(int which)
{
    if (!timing) return;
    if (ev_pool.size() < ev_used + 2) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        ev_pool.push_back(a); ev_pool.push_back(b);
    }
    cudaEventRecord(ev_pool[ev_used], stream);
    ev_which.push_back(which);
}
void Ctx::timer_end()
{
    if (!timing) return;
    cudaEventRecord(ev_pool[ev_used + 1], stream);
    ev_used += 2;
}
void Ctx::timer_collect()
{
    if (!ev_used) return;
    cudaStreamSynchronize(stream);
    for (size_t i = 0; i < ev_used; i += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]);
        const int w = ev_which[i / 2];
        k_ms[w] += ms; k_launches[w]++;
    }
    ev_used = 0; ev_which.clear();
}

// Run the scattering cascade dry, once per step (c2d_sync): the history kernels of the census, volume and surface phases
// have queued their collisions; every generation of collisions spawns split2 children (k_children) whose own collisions
// ('multiple scattering', imctrk2d.f:614) form the next generation, and children that trigger the 3rd split go through
// k_split3_attempts -> k_split3_hard -> k_split3_flights.
//   * device-driven: k_children and k_split3_hard read their work counts from device memory; the host looks at the counts
//     ONCE per generation (it has to size the attempt launches and to know when the cascade has ended);
//   * across ranks: the counts and the 3rd-split tasks of all ranks are all-gathered and every rank takes every world-th
//     sub-child of the concatenated list (S3Map).  The 3rd split's cost per task varies over six decades; without the
//     re-deal the ranks wait for the unluckiest one at the step's all-reduce (SCALE_r01: 57 % efficiency at 8 GPUs).
//     Collective: with world > 1 every rank must call c2d_sync the same number of times per step.
static int drain_queue(Ctx *c)
{
    if (!c->cascade_dirty) return 0;                               // nothing was tracked since the last drain (and no collective is due)
    c->cascade_dirty = false;
    const int world = comm_world(c), rank = comm_rank(c);
    static const bool trace = getenv("C2D_TRACE") != nullptr;     // host clock at every look at the device (each follows a stream sync)
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what, int gen, unsigned long long n) {
        if (trace) fprintf(stderr, "[c2d drain r%d] gen %d %-10s n=%llu t=%.3f ms\n", rank, gen, what, n,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    auto lap_count = [&](const char *what, int gen, const unsigned long long *dev) -> int {
        if (!trace) return 0;
        unsigned long long v = 0;
        if (dev) CU(cudaMemcpyAsync(&v, dev, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        lap(what, gen, v);
        return 0;
    };
    if (!c->d_cnt3) {
        CU(cudaMalloc(&c->d_cnt3, (size_t)(3 + 3 * 8) * 8));
        CU(cudaMalloc(&c->d_casc, 16 * 8));
    }
    auto gather_counts = [&](const unsigned long long *dev_count, unsigned long long *h, int fam) -> int {
        unsigned long long *mine = c->d_cnt3, *all = c->d_cnt3 + 3;
        CU(cudaMemcpyAsync(mine, dev_count, 8, cudaMemcpyDeviceToDevice, c->stream));
        if (world > 1) { c->timer_begin(fam); int rc = comm_allgather(c, mine, all, 8); c->timer_end(); if (rc) return rc; }
        else CU(cudaMemcpyAsync(all, mine, 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(h, all, (size_t)world * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    };
    // the tasks of all ranks, rank r's at slot r (n_max each); one rank: the queue itself
    auto gather_tasks = [&](const CollTask *mine_q, const unsigned long long *h, S3Map &m) -> int {
        memset(&m, 0, sizeof m);
        m.world = world; m.rank = rank;
        unsigned long long n_max = 0;
        for (int r = 0; r < world; r++) { m.prefix[r + 1] = m.prefix[r] + (long long)h[r]; n_max = std::max(n_max, h[r]); }
        if (world == 1) { m.base = mine_q; m.stride = (long long)h[0]; return 0; }
        if (!c->d_q3_all) {                                        // once, at the queues' capacity: an allocation (or a free) inside a step
            const size_t need = (size_t)world * (size_t)c->q_cap;  // stalls the stream for 10 .. 900 ms (measured on the B200 box)
            CU(cudaMalloc(&c->d_q3_all, need * sizeof(CollTask)));
            c->q3all_cap = need;
        }
        c->timer_begin(10);
        int rc = comm_allgather(c, mine_q, c->d_q3_all, (size_t)n_max * sizeof(CollTask));
        c->timer_end();
        if (rc) return rc;
        m.base = c->d_q3_all; m.stride = (long long)n_max;
        return 0;
    };
    for (int gen = 0; gen < 1000; gen++) {
        const int other = c->q_cur ^ 1;
        Params p = c->P;
        p.q_out = c->d_q[other]; p.q_out_count = c->d_qcount + other;
        // this generation's collisions on every rank (the first look of a step is where a rank waits for the others' history
        // kernels: that wait is part of family 10's time)
        unsigned long long h[8];
        if (int rc = gather_counts(c->d_qcount + c->q_cur, h, 10)) return rc;
        unsigned long long n_tot = 0, n3_tot = 0;
        for (int r = 0; r < world; r++) {
            if (h[r] > (unsigned long long)c->q_cap)
                return fail(C2D_EINVAL, "collision queue overflow: %llu tasks > collision_capacity %lld", h[r], (long long)c->q_cap);
            n_tot += h[r];
        }
        lap("collisions", gen, n_tot);
        if (n_tot == 0) return 0;                                  // no rank has a collision left: the cascade has ended
        c->q_peak = std::max<long long>(c->q_peak, (long long)h[rank]);
        {
            S3Map mq;
            if (int rc = gather_tasks(c->d_q[c->q_cur], h, mq)) return rc;
            const long long G = (long long)n_tot * p.split2;
            const long long n_mine = (G > rank) ? (G - rank + world - 1) / world : 0;
            if (!c->d_chard) {                                     // fixed size: a child that finds the list full finishes its loop itself
                c->chard_cap = (size_t)1 << 20;
                CU(cudaMalloc(&c->d_chard, c->chard_cap * sizeof(unsigned)));
                CU(cudaMalloc(&c->d_cultra, c->chard_cap * sizeof(GridItem)));
            }
            // d_casc: [0] hard children [1] fetch cursor [2] grid-team children | [3] hard sub-children [4] cursor [5] grid-team sub-children
            CU(cudaMemsetAsync(c->d_casc, 0, 24, c->stream));
            CU(cudaMemsetAsync(c->d_casc + 8, 0xFF, 24, c->stream));
            c->timer_begin(1);
            launch_children(p, mq, n_mine, c->n_sm, c->d_chard, (unsigned long long)c->chard_cap, c->d_casc, c->stream);
            c->timer_end();
            CU(cudaGetLastError());
            if (int rc = lap_count("children", gen, c->d_casc)) return rc;
#ifndef C2D_EMU
            if (n_mine > 0) {
                // the ladder of the electron / Klein-Nishina rejection loop: thread -> block team -> the whole grid
                c->timer_begin(5);
                launch_children_hard(p, mq, c->d_chard, (unsigned long long)c->chard_cap, c->d_cultra, c->d_casc, c->kn_hard_rounds, c->kn_drain_rounds,
                                     c->n_sm, c->stream);
                if (int rc = lap_count("c_hard", gen, c->d_casc + 2)) return rc;
                const int e = launch_children_ultra(p, mq, c->d_cultra, c->d_casc + 2, c->d_casc + 8, c->n_sm, c->stream);
                c->timer_end();
                if (e) return fail(C2D_ECUDA, "cooperative launch of k_children_ultra failed: %s", cudaGetErrorString((cudaError_t)e));
                CU(cudaGetLastError());
                if (int rc = lap_count("c_ultra", gen, nullptr)) return rc;
            }
#endif
        }
        if (int rc = gather_counts(c->d_qcount + 2, h, 10)) return rc;
        unsigned long long n3_max = 0;
        for (int r = 0; r < world; r++) { n3_tot += h[r]; n3_max = std::max(n3_max, h[r]); }
        lap("split3", gen, n3_tot);
        if (n3_max > (unsigned long long)c->q_cap)
            return fail(C2D_EINVAL, "split3 queue overflow: %llu tasks > collision_capacity %lld", n3_max, (long long)c->q_cap);
        if (n3_tot) {
            S3Map m;
            if (int rc = gather_tasks(c->d_q3, h, m)) return rc;
            // this rank's sub-children: global sub-child rank, rank + world, ...
            const long long G = (long long)n3_tot * p.split3;
            const long long n_mine = (G > rank) ? (G - rank + world - 1) / world : 0;
            if (!c->d_flights) {
                // one launch sequence per chunk of this many sub-children: every chunk ends with the tail of its hardest sub-children,
                // so few large chunks (0.9 GB of flight tasks) beat many small ones
                int64_t want = (int64_t)1 << 23;
                if (const char *e = getenv("C2D_FLIGHT_CAP")) want = std::max<int64_t>(1024, atoll(e));
                c->flight_cap = std::max<int64_t>((int64_t)p.split3, want);
                CU(cudaMalloc(&c->d_flights, (size_t)c->flight_cap * sizeof(FlightTask)));
                CU(cudaMalloc(&c->d_hard, (size_t)c->flight_cap * sizeof(unsigned)));
                CU(cudaMalloc(&c->d_ultra, (size_t)c->flight_cap * sizeof(GridItem)));
            }
            for (long long i0 = 0; i0 < n_mine; i0 += c->flight_cap) {
                const long long ns = std::min<long long>(c->flight_cap, n_mine - i0);
                CU(cudaMemsetAsync(c->d_casc + 3, 0, 24, c->stream));
                CU(cudaMemsetAsync(c->d_casc + 8, 0xFF, 24, c->stream));
                c->timer_begin(2);
                launch_split3_attempts(p, m, i0, ns, c->d_flights, c->split3_rounds, c->d_hard, c->d_casc + 3, c->stream);
                c->timer_end();
                CU(cudaGetLastError());
                if (int rc = lap_count("attempts", gen, c->d_casc + 3)) return rc;
                if (c->split3_rounds > 0) {
                    // the ladder: warp team -> block team (k_split3_hard) -> the whole grid (k_split3_ultra, cooperative launch)
                    const unsigned long long first_hard = (unsigned long long)c->split3_rounds * kWarp;
                    c->timer_begin(12);
                    launch_split3_hard(p, m, i0, c->d_hard, c->d_ultra, c->d_casc + 3, c->d_flights, first_hard, c->split3_hard_rounds,
                                       c->split3_drain_rounds, c->n_sm, c->stream);
                    c->timer_end();
                    CU(cudaGetLastError());
                    if (int rc = lap_count("s3_hard", gen, c->d_casc + 5)) return rc;
                    if (c->split3_hard_rounds >= 0) {
                        c->timer_begin(13);
                        const int e = launch_split3_ultra(p, m, i0, c->d_ultra, c->d_casc + 5, c->d_flights, c->d_casc + 8, c->n_sm, c->stream);
                        c->timer_end();
                        if (e) return fail(C2D_ECUDA, "cooperative launch of k_split3_ultra failed: %s", cudaGetErrorString((cudaError_t)e));
                        if (int rc = lap_count("s3_ultra", gen, nullptr)) return rc;
                    }
                }
                c->timer_begin(6);
                launch_split3_flights(p, c->d_flights, ns, c->stream);
                c->timer_end();
                CU(cudaGetLastError());
                if (trace) {
                    unsigned long long hc[8];
                    CU(cudaMemcpyAsync(hc, c->d_casc, sizeof hc, cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaStreamSynchronize(c->stream));
                    fprintf(stderr, "[c2d drain r%d]   hard children %llu, hard sub-children %llu, grid-team sub-children %llu\n", rank, hc[0], hc[3], hc[5]);
                    lap("flights", gen, (unsigned long long)ns);
                }
            }
        }
        CU(cudaMemsetAsync(c->d_qcount + c->q_cur, 0, 8, c->stream));
        CU(cudaMemsetAsync(c->d_qcount + 2, 0, 8, c->stream));
        c->q_cur = other;
        c->P.q_out = c->d_q[c->q_cur]; c->P.q_out_count = c->d_qcount + c->q_cur;
    }
    return fail(C2D_EINVAL, "scattering cascade did not terminate after 1000 generations");
}

// cr_sent = 1 / 3 reflect specularly only where the lower boundary has no temperature (imcleak2d.f:113); a hot boundary
// reflects through the P_ref / W_abs tables of ref_matrix.f (l.124-158), which the caller must have supplied
static int check_mirror(Ctx *c, const double *tbbl)
{
    if ((c->dims.cr_sent != 1 && c->dims.cr_sent != 3) || !tbbl || c->P.n_ref > 0) return 0;
    for (int k = 0; k < c->dims.nr; k++)
        if (tbbl[k] > 0.0)
            return fail(C2D_EINVAL, "cr_sent = %d with tbbl(%d) = %g > 0: Compton reflection off a hot lower boundary (imcleak2d.f:124-158) "
                        "needs the tables E_ref, P_ref, W_abs in c2d_dims", c->dims.cr_sent, k + 1, tbbl[k]);
    return 0;
}

static int check_fatal(Ctx *c)
{
    unsigned long long cnt[CNT_N];
    CU(cudaMemcpyAsync(cnt, c->P.counters, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (cnt[CNT_NDXOUT] > (unsigned long long)c->census_cap)
        return fail(C2D_ECENSUS, "too many photons: census needs %llu records, capacity %lld (imctrk2d.f:581-585)",
                    cnt[CNT_NDXOUT], (long long)c->census_cap);
    if (cnt[CNT_FATAL]) return fail(C2D_EINVAL, "device reported %llu fatal conditions (queue overflow or scat_flag=0 collision)", cnt[CNT_FATAL]);
    return 0;
}

// records beyond event_capacity are counted, not stored -- whoever consumes the stored list must fail loudly
int check_events(Ctx *c)
{
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->P.counters + CNT_NEVENTS, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (v > (unsigned long long)c->event_cap)
        return fail(C2D_EEVENTS, "%llu escape records produced, event_capacity %lld: the list is incomplete", v, (long long)c->event_cap);
    return 0;
}

// ---- sigma_mode 1: see c2d_sigma.cu
int Ctx::sigma_setup()
{
    const int N = dims.num_nt, nb = N - 1, D = 8;
    std::vector<double> th(nb);
    for (int i = 0; i < nb; i++) th[i] = std::log(1.e-2 / (h_gam0[i] * h_opb[i]));
    for (int i = 1; i < nb; i++)
        if (!(th[i] < th[i - 1])) return fail(C2D_EINVAL, "sigma table: the electron grid gnt must increase strictly (bin %d)", i);
    const double t0 = th[0], h = 0.2, tmax = std::log(1.e14 / 5.11e2);
    const int nsub = std::max(1, (int)std::ceil((tmax - t0) / h));
    const int n_int = nb + nsub;
    sig_nodes = n_int * D;
    std::vector<double> lo(n_int, 0.0), iw(n_int, 0.0), xs(sig_nodes, 1.0), M(D * D);
    std::vector<int> ms(sig_nodes, nb);
    auto interval = [&](int j, double a, double b, int m) {
        lo[j] = a; iw[j] = 2.0 / (b - a);
        for (int n = 0; n < D; n++) {
            const double sn = std::cos(kPi * (n + 0.5) / D);
            xs[j * D + n] = std::exp(0.5 * (a + b) + 0.5 * (b - a) * sn);
            ms[j * D + n] = m;
        }
    };
    for (int m = 1; m < nb; m++) interval(m, th[m], th[m - 1], m);       // m(x) = m on [theta_m, theta_{m-1})
    for (int q = 0; q < nsub; q++) interval(nb + q, t0 + q * h, t0 + (q + 1) * h, 0);
    for (int k = 0; k < D; k++)
        for (int n = 0; n < D; n++) M[k * D + n] = (k == 0 ? 1.0 : 2.0) / D * std::cos(kPi * k * (n + 0.5) / D);
    const size_t n_static = (size_t)2 * n_int + sig_nodes + D * D + (size_t)nb * sig_nodes;
    CU(cudaMalloc(&d_sig_static, n_static * 8));
    double *S = d_sig_static;
    double *d_lo = S, *d_iw = S + n_int, *d_xs = S + 2 * n_int;
    d_sig_M = d_xs + sig_nodes; d_sig_Kn = d_sig_M + D * D;
    CU(cudaMemcpy(d_lo, lo.data(), n_int * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_iw, iw.data(), n_int * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_xs, xs.data(), (size_t)sig_nodes * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_sig_M, M.data(), D * D * 8, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_sig_ms, (size_t)sig_nodes * 4));
    CU(cudaMemcpy(d_sig_ms, ms.data(), (size_t)sig_nodes * 4, cudaMemcpyHostToDevice));
    const size_t n_step = (size_t)nzone * N * 5 + (size_t)nzone * sig_nodes * 2 + (((size_t)nzone + 1) & ~size_t(1));
    CU(cudaMalloc(&d_sig_step, n_step * 8));
    d_sig_W = d_sig_step; d_sig_AB = d_sig_W + (size_t)nzone * N; d_sig_T = d_sig_AB + (size_t)nzone * N * 2;
    d_sig_coef = d_sig_T + (size_t)nzone * sig_nodes;
    double *d_NB = d_sig_coef + (size_t)nzone * sig_nodes;
    P.sig_NB = d_NB; P.sig_CW = d_NB + (size_t)nzone * N; P.sig_if = reinterpret_cast<const int *>(d_NB + (size_t)2 * nzone * N);
    P.sig_coef = d_sig_coef; P.sig_AB = d_sig_AB; P.sig_W = d_sig_W; P.sig_lo = d_lo; P.sig_iw = d_iw;
    P.sig_nint = n_int; P.sig_nsub = nsub; P.sig_t0 = t0; P.sig_inv_h = 1.0 / h;
    launch_sigma_kn(P, d_xs, d_sig_ms, sig_nodes, d_sig_Kn, stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(stream));
    return 0;
}

int Ctx::sigma_ensure()
{
    if (sigma_ready) return 0;                                    // once per step: the tables that follow Pnt / f_nt / n_e
    sigma_ready = true;
#ifndef C2D_EMU
    if (dims.num_nt < 65536 && !getenv("C2D_NO_NT_GUIDE")) {
        if (!d_nt_guide) CU(cudaMalloc(&d_nt_guide, (size_t)dims.nz * dims.nr * kNtGuide * sizeof(uint32_t)));
        P.nt_guide = d_nt_guide;
        timer_begin(4);
        launch_nt_guide(P, d_nt_guide, stream);
        timer_end();
        CU(cudaGetLastError());
    }
#endif
    if (dims.sigma_mode != 1) return 0;
    return sigma_build();
}

int Ctx::sigma_build()
{
    timer_begin(4);
    launch_sigma_build(P, d_sig_W, d_sig_AB, d_sig_Kn, sig_nodes, d_sig_T, d_sig_M, d_sig_coef, stream);
    timer_end();
    CU(cudaGetLastError());
    return 0;
}

}  // namespace c2d

// k_history in chunks of at most qabs_cap histories (each queues at most one absorption-draw task), every chunk followed by
// k_absdraw on the device-side count
namespace c2d {
int launch_history_chunks(Ctx *c, int src, int64_t first, int64_t count, int64_t stride)
{
    const int64_t chunk = std::min<int64_t>((int64_t)1 << 30, c->qabs_cap);
    for (int64_t done = 0; done < count; done += chunk) {
        const int64_t n = std::min(chunk, count - done);
        c->timer_begin(0);
        if (src == 0) launch_history_volume(c->P, first + done * stride, n, stride, c->stream);
        else if (src == 1) launch_history_census(c->P, first + done * stride, n, stride, c->stream);
        else launch_history_surface(c->P, first + done * stride, n, stride, c->stream);
        launch_absdraw(c->P, c->d_qabs, c->d_qcount + 4, c->n_sm, c->stream);
        c->timer_end();
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(c->d_qcount + 4, 0, 16, c->stream));        // absorption-draw queue and literal pool
    }
    return 0;
}
}  // namespace c2d

// ---- pieces of the step prologue shared by c2d_step_begin_gen (host arrays) and c2d_step (resident plasma state)
namespace c2d {

int latch_census(Ctx *c)
{   // the photon count in flight: take over what the previous step wrote
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, c->P.counters + CNT_NDXOUT, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->n_census = (int64_t)std::min<unsigned long long>(n, (unsigned long long)c->census_cap);
    return 0;
}

// volume_em for every zone from the f_nt / n_e / tea already in the step blob and the B_field at d_B (device), then the
// per-zone energy budget in the reference's zone order (imcgen2d.f:280-283,291,330) and the volume photon budget (l.405-459)
int gen_core(Ctx *c, double dt, int64_t nst, const double *d_B, const double *h_ne, const double *h_B, const double *h_vol,
             const double *h_zsurf, GenBudget &g)
{
    const c2d_dims &d = c->dims;
    const int nzone = c->nzone, nv = d.n_vol;
    // volume_em regenerates E_ph on every call (volume2d.f:53,96,112-113); the grid given to c2d_init must be that one
    {
        const double dE = std::exp(std::log(1.e20) / (double)nv);
        double E = 1.e-10 / dE;
        for (int i = 0; i < nv; i++) {
            E = E * dE;
            if (std::fabs(c->h_Eph[i] - E) > 1.e-13 * E)
                return fail(C2D_EINVAL, "E_ph(%d) = %.17g differs from the grid volume_em generates (%.17g, volume2d.f:53,96)", i + 1, c->h_Eph[i], E);
        }
    }
    double *T = c->d_step;
    // scratch: j_sy | sum_th | sum_g1
    const size_t n_scr = (size_t)nzone * (nv + 3);
    if (n_scr > c->gen_cap) {
        cudaFree(c->d_gen); c->d_gen = nullptr; c->gen_cap = 0;
        CU(cudaMalloc(&c->d_gen, n_scr * 8));
        c->gen_cap = n_scr;
    }
    double *G = c->d_gen;
    GenParams q;
    q.nzone = nzone; q.nz = d.nz; q.nr = d.nr; q.N = d.num_nt; q.nv = nv; q.rmin = d.rmin;
    q.z = c->P.z; q.r = c->P.r; q.gnt = c->P.gnt; q.E_ph = c->P.E_ph;
    q.f_nt = T + c->o_fnt; q.n_e = T + c->o_ne; q.tea = T + c->o_tea; q.B_field = d_B;
    q.j_sy = G + nzone; q.sum_th = G + nzone + (size_t)nzone * nv; q.sum_g1 = q.sum_th + nzone;
    q.kappa_tot = T + c->o_kappa; q.eps_tot = T + c->o_epstot; q.eps_th = T + c->o_epsth;
    c->timer_begin(4);
    launch_volume_em(q, c->stream);
    c->timer_end();
    CU(cudaGetLastError());
    std::vector<double> hs((size_t)2 * nzone);
    CU(cudaMemcpyAsync(hs.data(), q.sum_th, (size_t)2 * nzone * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    g.e_sy.assign(nzone, 0.0); g.e_th.assign(nzone, 0.0); g.ew.assign(nzone, 0.0); g.nsv.assign(nzone, 0);
    g.Emiss_tot = 0.0;
    for (int zi = 0; zi < nzone; zi++) {
        g.e_sy[zi] = 1.058e-15 * h_ne[zi] * dt * std::pow(h_B[zi], 2.0) * hs[nzone + zi] * h_vol[zi];
        g.e_th[zi] = dt * h_zsurf[zi] * hs[zi];
        g.Emiss_tot = g.Emiss_tot + g.e_sy[zi];
    }
    for (int zi = 0; zi < nzone; zi++) {
        g.nsv[zi] = (g.Emiss_tot > 0.0) ? (int32_t)(0.5 * (double)nst * g.e_sy[zi] / g.Emiss_tot) : 0;
        g.ew[zi] = (g.nsv[zi] > 0) ? g.e_sy[zi] / (double)g.nsv[zi] : 0.0;
    }
    return 0;
}

int budget_upload(Ctx *c, const GenBudget &g)
{
    const int nzone = c->nzone;
    double *T = c->d_step;
    c->h_prefix.assign(nzone + 1, 0);
    for (int zi = 0; zi < nzone; zi++) c->h_prefix[zi + 1] = c->h_prefix[zi] + g.nsv[zi];
    int rc = 0;
    rc |= upload(c, T + c->o_ewsv, g.ew.data(), nzone);
    rc |= upload(c, T + c->o_elth, g.e_th.data(), nzone);
    rc |= upload(c, T + c->o_eltot, g.e_sy.data(), nzone);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_prefix, c->h_prefix.data(), (size_t)(nzone + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// zero the per-step tallies (imcgen2d.f:87-104,534-551,559-601), swap the census banks, open the step
int begin_common(Ctx *c, int ncycle, double time, double dt)
{
    CU(cudaMemsetAsync(c->d_f64, 0, c->n_f64 * 8, c->stream));
    CU(cudaMemsetAsync(c->d_i64, 0, c->n_i64 * 8, c->stream));
    // census: last step's output bank becomes this step's input (no 360 MB copy as in imcfield2d.f:96-97) -- sorted by
    // (zone, energy class) into the other bank on the way (c2d_bank.cu), in which case the roles stay as they are
    // (c2d_tally_reduce has usually done it already at the end of the previous step)
    const int sorted = c->bank_sorted ? 1 : bank_sort(c, c->n_census);
    c->bank_sorted = false;
    if (sorted < 0) return C2D_ECUDA;
    if (!sorted) c->cur_out ^= 1;
    c->P.out = c->bank[c->cur_out]; c->P.in = c->bank[c->cur_out ^ 1];
    c->P.ncycle = ncycle; c->P.key1 = (uint32_t)ncycle; c->P.dt = dt; c->P.time = time;
    c->field_done = false; c->vol_next = 0; c->surf_open = false; c->surf_total = 0;
    c->sigma_ready = false; c->step_reduced = false; c->step_tracked = false;
    CU(cudaStreamSynchronize(c->stream));      // the caller may reuse its host arrays
    c->step_open = true;
    return 0;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

const char *c2d_last_error(void) { return g_err; }
int c2d_version(void) { return 100; }

int c2d_create(c2d_ctx **out, int device)
{
    if (!out) return fail(C2D_EINVAL, "ctx pointer is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(C2D_ENODEV, "no CUDA device: compton2d_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(C2D_ENODEV, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(C2D_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200)", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    Ctx *c = new Ctx();
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
#ifdef C2D_EMU
    c->split3_rounds = 0;                                         // the serial debug build has no block-team kernel
#else
    if (const char *e = getenv("C2D_SPLIT3_ROUNDS")) c->split3_rounds = std::max(1, atoi(e));
    if (const char *e = getenv("C2D_SPLIT3_HARD_ROUNDS")) c->split3_hard_rounds = atoi(e);      // < 0: block teams without limit, no grid team
    if (const char *e = getenv("C2D_KN_HARD_ROUNDS")) c->kn_hard_rounds = std::max(0, atoi(e));
    if (const char *e = getenv("C2D_KN_DRAIN_ROUNDS")) c->kn_drain_rounds = std::max(0, atoi(e));
    if (const char *e = getenv("C2D_SPLIT3_DRAIN_ROUNDS")) c->split3_drain_rounds = std::max(0, atoi(e));
    if (const char *e = getenv("C2D_KN_CAP")) c->kn_cap = std::max(1, atoi(e));
#endif
    *out = reinterpret_cast<c2d_ctx *>(c);
    return 0;
}

int c2d_destroy(c2d_ctx *h)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_ref); cudaFree(c->d_static); cudaFree(c->d_step); cudaFree(c->d_prefix); cudaFree(c->d_surf); cudaFree(c->d_gen); cudaFree(c->d_f64); cudaFree(c->d_i64);
    cudaFree(c->d_events); cudaFree(c->d_q[0]); cudaFree(c->d_q[1]); cudaFree(c->d_q3); cudaFree(c->d_flights); cudaFree(c->d_hard); cudaFree(c->d_qcount);
    cudaFree(c->d_sig_static); cudaFree(c->d_sig_step); cudaFree(c->d_sig_ms);
    cudaFree(c->d_sortkeys); cudaFree(c->d_qabs); cudaFree(c->d_litpool); cudaFree(c->d_cnt3); cudaFree(c->d_q3_all); cudaFree(c->d_nt_guide); cudaFree(c->d_casc); cudaFree(c->d_chard); cudaFree(c->d_cultra); cudaFree(c->d_ultra);
    cudaFree(c->d_fout_total); cudaFree(c->d_cnt_global); cudaFree(c->d_plasma); cudaFree(c->d_fpx);
    comm_free(c);
    free_bank(c->bank[0]); free_bank(c->bank[1]);
    c->fp_free();
    c->lc_free_all();
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int c2d_init(c2d_ctx *h, const c2d_dims *d)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !d) return fail(C2D_EINVAL, "NULL argument");
    if (c->inited) return fail(C2D_ESTATE, "context already initialised");
    if (d->cr_sent < 0 || d->cr_sent > 4) return fail(C2D_EINVAL, "cr_sent = %d", d->cr_sent);
    const bool have_ref = d->n_ref > 1 && d->E_ref && d->P_ref && d->W_abs;
    if ((d->cr_sent == 2 || d->cr_sent == 3) && !have_ref)
        return fail(C2D_EINVAL, "cr_sent = %d: Compton reflection (imcleak2d.f:124-165,216-273) needs the tables E_ref, P_ref, W_abs of "
                    "ref_matrix.f in c2d_dims", d->cr_sent);
    if (d->nz < 1 || d->nr < 1 || d->nz > 65535 || d->nr > 65535) return fail(C2D_EINVAL, "bad grid %d x %d", d->nz, d->nr);
    if (d->nmu < 1 || d->nmu > C2D_NMUMAX || d->nph_lc < 0 || d->nph_lc > C2D_NLCMAX || d->nphtotal < 1 || d->nphtotal > 65535)
        return fail(C2D_EINVAL, "bad bin counts nmu=%d nph_lc=%d nphtotal=%d", d->nmu, d->nph_lc, d->nphtotal);
    if (d->n_vol < 2 || d->num_nt < 3 || d->nphfield < 2 || d->n_gg < 2) return fail(C2D_EINVAL, "bad table sizes");
    if (d->split1 < 1 || d->split2 < 1 || d->split3 < 1) return fail(C2D_EINVAL, "split factors must be >= 1");
    if (!d->z || !d->r || !d->E_ph || !d->E_field || !d->gnt || !d->hu || !d->mu) return fail(C2D_EINVAL, "missing static table");
    CU(cudaSetDevice(c->device));
    c->dims = *d;
    const int nzone = d->nz * d->nr;
    c->nzone = nzone;
    c->census_cap = std::max<int64_t>(d->census_capacity, 1);
    c->event_cap = std::max<int64_t>(d->event_capacity, 1);
    c->q_cap = std::max<int64_t>(d->collision_capacity, 1024);

    // ---- static tables, one blob
    std::vector<double> hs;
    auto put = [&](const double *src, int n, double fill = 0.0) {
        size_t off = hs.size();
        for (int i = 0; i < n; i++) hs.push_back(src ? src[i] : fill);
        while (hs.size() % 2) hs.push_back(0.0);
        return off;
    };
    const size_t o_z = put(d->z, d->nz), o_r = put(d->r, d->nr), o_Eph = put(d->E_ph, d->n_vol);
    const size_t o_Ef = put(d->E_field, d->nphfield), o_Egg = put(d->E_gg, d->n_gg), o_gnt = put(d->gnt, d->num_nt);
    std::vector<double> gam0(d->num_nt), opb(d->num_nt), dgam(d->num_nt, 0.0), pref(d->num_nt), a1(d->num_nt), a2(d->num_nt), cb(d->num_nt);
    for (int i = 0; i < d->num_nt; i++) {
        // comtot2d.f:223-238, the bin-only factors in the reference's own left-to-right order
        const double sigthom = 6.6516e-25;
        gam0[i] = d->gnt[i] + 1.0;
        const volatile double betta = std::sqrt(1.0 - 1.0 / (gam0[i] * gam0[i]));
        opb[i] = 1 + betta;
        volatile double t = 9.375e-2 * sigthom;
        t = t / (gam0[i] * gam0[i]);
        t = t / betta;
        pref[i] = t;
        volatile double u = 2 * gam0[i];
        a1[i] = u * (1 + betta);
        a2[i] = u * (1 - betta);
        cb[i] = 8.0 * pref[i] / a2[i];                         // sigma_mode 1 noise term, see kn_noise_bin
        if (i + 1 < d->num_nt) dgam[i] = d->gnt[i + 1] - d->gnt[i];
    }
    const size_t o_g0 = put(gam0.data(), d->num_nt), o_be = put(opb.data(), d->num_nt), o_dg = put(dgam.data(), d->num_nt);
    const size_t o_pref = put(pref.data(), d->num_nt), o_a1 = put(a1.data(), d->num_nt), o_a2 = put(a2.data(), d->num_nt), o_cb = put(cb.data(), d->num_nt);
    const size_t o_hu = put(d->hu, d->nphtotal + 1);
    const size_t o_lmin = put(d->Elcmin, d->nph_lc), o_lmax = put(d->Elcmax, d->nph_lc), o_mu = put(d->mu, d->nmu);
    CU(cudaMalloc(&c->d_static, hs.size() * 8));
    CU(cudaMemcpy(c->d_static, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice));
    c->h_gnt.assign(d->gnt, d->gnt + d->num_nt);
    c->h_Eph.assign(d->E_ph, d->E_ph + d->n_vol);
    c->h_gam0 = gam0; c->h_opb = opb;
    c->h_Efield.assign(d->E_field, d->E_field + d->nphfield);
    c->h_z.assign(d->z, d->z + d->nz); c->h_r.assign(d->r, d->r + d->nr);

    // ---- per-step tables, one blob
    size_t n = 0;
    auto take = [&](size_t cnt) { size_t o = n; n += (cnt + 1) & ~size_t(1); return o; };
    c->o_kappa = take((size_t)nzone * d->n_vol); c->o_epstot = take((size_t)nzone * d->n_vol);
    c->o_epsth = take((size_t)nzone * d->n_vol); c->o_fnt = take((size_t)nzone * d->num_nt);
    c->o_pnt = take((size_t)nzone * d->num_nt); c->o_ne = take(nzone); c->o_fpair = take(nzone);
    c->o_kgg = take((size_t)nzone * d->n_gg); c->o_tbbl = take(d->nr); c->o_ewsv = take(nzone);
    c->o_elth = take(nzone); c->o_eltot = take(nzone); c->o_zsurf = take(nzone); c->o_tea = take(nzone);
    c->n_step = n;
    CU(cudaMalloc(&c->d_step, n * 8));
    CU(cudaMemset(c->d_step, 0, n * 8));
    CU(cudaMalloc(&c->d_prefix, (size_t)(nzone + 1) * 8));

    // ---- tallies: one packed FP64 buffer + one packed int64 buffer (one all-reduce each)
    n = 0;
    c->t_edep = take(nzone); c->t_prdep = take(nzone); c->t_ecens = take(nzone);
    c->t_nfield = take((size_t)nzone * d->nphfield); c->t_nph = take((size_t)nzone * d->n_gg);
    c->t_EIC = take(d->num_nt); c->t_fout = take((size_t)C2D_NMUMAX * d->nphtotal);
    c->t_edout = take((size_t)C2D_NMUMAX * C2D_NLCMAX);
    c->t_erlki = take(d->nz); c->t_erlko = take(d->nz); c->t_erlku = take(d->nr); c->t_erlkl = take(d->nr);
    c->t_Edin = take(d->nr); c->t_Edref = take(d->nr);
    c->n_f64 = n;
    CU(cudaMalloc(&c->d_f64, n * 8));
    CU(cudaMemset(c->d_f64, 0, n * 8));
    c->i_npcen = 0; c->i_nel = nzone; c->i_cnt = nzone + d->num_nt + 2;
    c->n_i64 = c->i_cnt + CNT_N;
    CU(cudaMalloc(&c->d_i64, c->n_i64 * 8));
    CU(cudaMemset(c->d_i64, 0, c->n_i64 * 8));
    CU(cudaMalloc(&c->d_fout_total, (size_t)C2D_NMUMAX * d->nphtotal * 8));
    CU(cudaMemset(c->d_fout_total, 0, (size_t)C2D_NMUMAX * d->nphtotal * 8));
    CU(cudaMalloc(&c->d_cnt_global, CNT_N * 8));
    CU(cudaMemset(c->d_cnt_global, 0, CNT_N * 8));

    // ---- banks, events, collision queues
    for (int b = 0; b < 2; b++) { int rc = alloc_bank(c->bank[b], c->census_cap); if (rc) return rc; }
    CU(cudaMalloc(&c->d_events, (size_t)c->event_cap * 7 * 8));
    for (int b = 0; b < 2; b++) CU(cudaMalloc(&c->d_q[b], (size_t)c->q_cap * sizeof(CollTask)));
    CU(cudaMalloc(&c->d_q3, (size_t)c->q_cap * sizeof(CollTask)));
    CU(cudaMalloc(&c->d_qcount, 64));      // [0],[1] collision queues, [2] split3 queue, [3] hard list, [4] absorption-draw queue, [5] literal pool
    CU(cudaMemset(c->d_qcount, 0, 64));
    // absorption-draw queue: k_history is launched in chunks of at most qabs_cap histories, each of which queues at most one task
    c->qabs_cap = std::min<int64_t>((int64_t)1 << 24, std::max<int64_t>((int64_t)1 << 16, c->census_cap));
    if (const char *e = getenv("C2D_QABS_CAP")) c->qabs_cap = std::max<int64_t>(1024, atoll(e));
    c->lit_cap = std::max<int64_t>((int64_t)1 << 22, (int64_t)64 * d->split1);
    CU(cudaMalloc(&c->d_qabs, (size_t)c->qabs_cap * sizeof(AbsTask)));
    CU(cudaMalloc(&c->d_litpool, (size_t)c->lit_cap * sizeof(int)));
    if (const char *e = getenv("C2D_SORT_BANK")) c->sort_bank = atoi(e) != 0;

    Params &P = c->P;
    memset(&P, 0, sizeof P);
    P.kn_cap = c->kn_cap;
    P.nz = d->nz; P.nr = d->nr; P.nzone = nzone; P.n_vol = d->n_vol; P.num_nt = d->num_nt; P.nphfield = d->nphfield;
    P.n_gg = d->n_gg; P.nphtotal = d->nphtotal; P.nph_lc = d->nph_lc; P.nmu = d->nmu;
    P.pair_switch = d->pair_switch; P.spec_switch = d->spec_switch; P.cr_sent = d->cr_sent;
    P.n_ref = 0; P.E_ref = P.P_ref = P.W_abs = nullptr;
    if (have_ref) {
        const size_t nr_ = (size_t)d->n_ref;
        CU(cudaMalloc(&c->d_ref, (nr_ + 2 * nr_ * nr_) * 8));
        CU(cudaMemcpy(c->d_ref, d->E_ref, nr_ * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_ref + nr_, d->P_ref, nr_ * nr_ * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_ref + nr_ + nr_ * nr_, d->W_abs, nr_ * nr_ * 8, cudaMemcpyHostToDevice));
        P.n_ref = d->n_ref; P.E_ref = c->d_ref; P.P_ref = c->d_ref + nr_; P.W_abs = c->d_ref + nr_ + nr_ * nr_;
    }
    P.split1 = d->split1; P.split2 = d->split2; P.split3 = d->split3; P.spl3_trg = d->spl3_trg;
    P.sigma_mode = d->sigma_mode;
    P.key0 = d->rseed; P.rmin = d->rmin; P.zmin = d->zmin;
    const double *S = c->d_static;
    P.z = S + o_z; P.r = S + o_r; P.E_ph = S + o_Eph; P.E_field = S + o_Ef; P.E_gg = S + o_Egg; P.gnt = S + o_gnt;
    P.gam0 = S + o_g0; P.opb = S + o_be; P.dgam = S + o_dg; P.kn_pref = S + o_pref; P.kn_a1 = S + o_a1; P.kn_a2 = S + o_a2; P.kn_cb = S + o_cb; P.hu = S + o_hu; P.Elcmin = S + o_lmin; P.Elcmax = S + o_lmax;
    P.mu = S + o_mu;
    const double *T = c->d_step;
    P.kappa_tot = T + c->o_kappa; P.eps_tot = T + c->o_epstot; P.eps_th = T + c->o_epsth; P.f_nt = T + c->o_fnt;
    P.Pnt = T + c->o_pnt; P.n_e = T + c->o_ne; P.f_pair = T + c->o_fpair; P.k_gg = T + c->o_kgg; P.tbbl = T + c->o_tbbl;
    P.ewsv = T + c->o_ewsv; P.Eloss_th = T + c->o_elth; P.Eloss_tot = T + c->o_eltot; P.zsurf = T + c->o_zsurf;
    P.nsv_prefix = c->d_prefix;
    double *F = c->d_f64;
    P.edep = F + c->t_edep; P.prdep = F + c->t_prdep; P.ecens = F + c->t_ecens; P.n_field = F + c->t_nfield;
    P.n_ph = F + c->t_nph; P.E_IC = F + c->t_EIC; P.fout = F + c->t_fout; P.edout = F + c->t_edout;
    P.erlki = F + c->t_erlki; P.erlko = F + c->t_erlko; P.erlku = F + c->t_erlku; P.erlkl = F + c->t_erlkl;
    P.Ed_in = F + c->t_Edin; P.Ed_ref = F + c->t_Edref;
    P.npcen = c->d_i64 + c->i_npcen; P.nelectron = c->d_i64 + c->i_nel; P.counters = c->d_i64 + c->i_cnt;
    c->cur_out = 0;
    P.out = c->bank[0]; P.in = c->bank[1];
    P.census_cap = c->census_cap; P.events = c->d_events; P.event_cap = c->event_cap;
    c->q_cur = 0;
    P.q_out = c->d_q[0]; P.q_out_count = c->d_qcount; P.q_cap = c->q_cap;
    P.q3 = c->d_q3; P.q3_count = c->d_qcount + 2;
    P.qabs = c->d_qabs; P.qabs_count = c->d_qcount + 4; P.qabs_cap = (unsigned long long)c->qabs_cap;
    P.lit_pool = c->d_litpool; P.lit_count = c->d_qcount + 5; P.lit_cap = (unsigned long long)c->lit_cap;
    c->n_census = 0;

    if (d->sigma_mode == 1) { int rc = c->sigma_setup(); if (rc) return rc; }
    c->inited = true;
    return 0;
}

int c2d_step_begin(c2d_ctx *h, const c2d_step_in *in)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !in) return fail(C2D_EINVAL, "NULL argument");
    if (!c->inited) return fail(C2D_ESTATE, "c2d_init has not been called");
    CU(cudaSetDevice(c->device));
    const c2d_dims &d = c->dims;
    const int nzone = c->nzone;
    if (!in->kappa_tot || !in->f_nt || !in->Pnt || !in->n_e) return fail(C2D_EINVAL, "missing per-step table");
    if (!(in->dt >= 0.0)) return fail(C2D_EINVAL, "dt = %g", in->dt);
    if (int rc_l = latch_census(c)) return rc_l;
    double *T = c->d_step;
    int rc = 0;
    rc |= upload(c, T + c->o_kappa, in->kappa_tot, (size_t)nzone * d.n_vol);
    rc |= upload(c, T + c->o_epstot, in->eps_tot, (size_t)nzone * d.n_vol);
    rc |= upload(c, T + c->o_epsth, in->eps_th, (size_t)nzone * d.n_vol);
    rc |= upload(c, T + c->o_fnt, in->f_nt, (size_t)nzone * d.num_nt);
    rc |= upload(c, T + c->o_pnt, in->Pnt, (size_t)nzone * d.num_nt);
    rc |= upload(c, T + c->o_ne, in->n_e, nzone);
    rc |= upload(c, T + c->o_fpair, in->f_pair, nzone);
    rc |= upload(c, T + c->o_kgg, in->k_gg, (size_t)nzone * d.n_gg);
    rc |= upload(c, T + c->o_tbbl, in->tbbl, d.nr);
    if (int rc_m = check_mirror(c, in->tbbl)) return rc_m;
    rc |= upload(c, T + c->o_ewsv, in->ewsv, nzone);
    rc |= upload(c, T + c->o_elth, in->Eloss_th, nzone);
    rc |= upload(c, T + c->o_eltot, in->Eloss_tot, nzone, 1.0);
    rc |= upload(c, T + c->o_zsurf, in->zsurf, nzone, 1.0);
    rc |= upload(c, T + c->o_tea, in->tea, nzone);
    if (rc) return rc;
    c->h_prefix.assign(nzone + 1, 0);
    for (int zi = 0; zi < nzone; zi++) {
        const int nsv = in->nsv ? in->nsv[zi] : 0;
        if (nsv < 0) return fail(C2D_EINVAL, "nsv[%d] = %d", zi, nsv);
        c->h_prefix[zi + 1] = c->h_prefix[zi] + nsv;
    }
    CU(cudaMemcpyAsync(c->d_prefix, c->h_prefix.data(), (size_t)(nzone + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    return begin_common(c, in->ncycle, in->time, in->dt);
}

// == imcgen2d for the volume sources, on the device: volume_em for every zone from the live electron
// spectrum (csrc/c2d_gen.cu), the per-zone energy budget (imcgen2d.f:280-283,291,330) and the photon budget
// (imcgen2d.f:405-459), then everything c2d_step_begin does.  The opacity / emissivity tables never visit the host.
int c2d_step_begin_gen(c2d_ctx *h, const c2d_gen_in *in, c2d_gen_out *out)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !in) return fail(C2D_EINVAL, "NULL argument");
    if (!c->inited) return fail(C2D_ESTATE, "c2d_init has not been called");
    CU(cudaSetDevice(c->device));
    const c2d_dims &d = c->dims;
    const int nzone = c->nzone;
    if (!in->f_nt || !in->Pnt || !in->n_e || !in->tea || !in->B_field || !in->vol || !in->zsurf)
        return fail(C2D_EINVAL, "c2d_step_begin_gen: missing per-zone input");
    if (!(in->dt >= 0.0) || in->nst < 0) return fail(C2D_EINVAL, "dt = %g, nst = %lld", in->dt, (long long)in->nst);
    if (int rc = latch_census(c)) return rc;
    double *T = c->d_step;
    int rc = 0;
    rc |= upload(c, T + c->o_fnt, in->f_nt, (size_t)nzone * d.num_nt);
    rc |= upload(c, T + c->o_pnt, in->Pnt, (size_t)nzone * d.num_nt);
    rc |= upload(c, T + c->o_ne, in->n_e, nzone);
    rc |= upload(c, T + c->o_fpair, in->f_pair, nzone);
    rc |= upload(c, T + c->o_kgg, in->k_gg, (size_t)nzone * d.n_gg);
    rc |= upload(c, T + c->o_tbbl, in->tbbl, d.nr);
    if (int rc_m = check_mirror(c, in->tbbl)) return rc_m;
    rc |= upload(c, T + c->o_zsurf, in->zsurf, nzone, 1.0);
    rc |= upload(c, T + c->o_tea, in->tea, nzone);
    if (rc) return rc;
    // B_field rides in the first nzone doubles of the generator's scratch
    const size_t n_scr = (size_t)nzone * (d.n_vol + 3);
    if (n_scr > c->gen_cap) {
        cudaFree(c->d_gen); c->d_gen = nullptr; c->gen_cap = 0;
        CU(cudaMalloc(&c->d_gen, n_scr * 8));
        c->gen_cap = n_scr;
    }
    CU(cudaMemcpyAsync(c->d_gen, in->B_field, (size_t)nzone * 8, cudaMemcpyHostToDevice, c->stream));
    GenBudget g;
    if (int rc2 = gen_core(c, in->dt, in->nst, c->d_gen, in->n_e, in->B_field, in->vol, in->zsurf, g)) return rc2;
    if (int rc2 = budget_upload(c, g)) return rc2;
    if (int rc2 = begin_common(c, in->ncycle, in->time, in->dt)) return rc2;
    if (out) {
        if (out->nsv) memcpy(out->nsv, g.nsv.data(), (size_t)nzone * 4);
        if (out->ewsv) memcpy(out->ewsv, g.ew.data(), (size_t)nzone * 8);
        if (out->Eloss_sy) memcpy(out->Eloss_sy, g.e_sy.data(), (size_t)nzone * 8);
        if (out->Eloss_th) memcpy(out->Eloss_th, g.e_th.data(), (size_t)nzone * 8);
        if (out->Eloss_tot) memcpy(out->Eloss_tot, g.e_sy.data(), (size_t)nzone * 8);
        out->Emiss_tot = g.Emiss_tot;
        out->n_volume = c->h_prefix.back();
    }
    return 0;
}

// D2H copy of the step tables the device holds (kappa_tot, eps_tot, eps_th: [zone][n_vol]); any pointer may be NULL
int c2d_step_tables_read(c2d_ctx *h, double *kappa_tot, double *eps_tot, double *eps_th)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->nzone * c->dims.n_vol * 8;
    if (kappa_tot) CU(cudaMemcpyAsync(kappa_tot, c->d_step + c->o_kappa, n, cudaMemcpyDeviceToHost, c->stream));
    if (eps_tot) CU(cudaMemcpyAsync(eps_tot, c->d_step + c->o_epstot, n, cudaMemcpyDeviceToHost, c->stream));
    if (eps_th) CU(cudaMemcpyAsync(eps_th, c->d_step + c->o_epsth, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int c2d_imcfield(c2d_ctx *h)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    if (c->field_done) return fail(C2D_ESTATE, "c2d_imcfield called twice in one step");
    CU(cudaSetDevice(c->device));
    c->field_done = true; c->step_tracked = true; c->cascade_dirty = true;
    if (int rc = c->sigma_ensure()) return rc;
    if (int rc = launch_history_chunks(c, 1, 0, c->n_census, 1)) return rc;
    c->histories += c->n_census;
    c->n_census = 0;
    return 0;                                                     // collisions stay queued until c2d_sync
}

int c2d_imcvol(c2d_ctx *h, int64_t first, int64_t count, int64_t stride)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    CU(cudaSetDevice(c->device));
    const int64_t total = c->h_prefix.back();
    if (stride < 1 || first < 0) return fail(C2D_EINVAL, "bad shard first=%lld stride=%lld", (long long)first, (long long)stride);
    int64_t avail = (first < total) ? (total - first + stride - 1) / stride : 0;
    if (count < 0 || count > avail) count = avail;
    c->step_tracked = true; c->cascade_dirty = true;
    if (int rc = c->sigma_ensure()) return rc;
    if (int rc = launch_history_chunks(c, 0, first, count, stride)) return rc;
    c->histories += count;
    return 0;
}

int c2d_surf_begin(c2d_ctx *h, const c2d_surf_in *in)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !in) return fail(C2D_EINVAL, "NULL argument");
    if (!c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    CU(cudaSetDevice(c->device));
    const int nz = c->dims.nz, nr = c->dims.nr, nslot = 2 * nz + 2 * nr;
    if (in->n_spec < 0 || in->nfmax < 0) return fail(C2D_EINVAL, "bad spectrum table sizes");
    std::vector<int64_t> prefix(nslot + 1, 0);
    std::vector<double> ew(nslot, 0.0), tbb(nslot, 0.0);
    std::vector<int> spec(nslot, 0);
    auto slot = [&](int q, const int32_t *n, const double *w, const double *t, const int32_t *sp, int ring) -> int {
        const int cnt = n ? n[ring] : 0;
        if (cnt < 0) return fail(C2D_EINVAL, "negative surface photon count (slot %d)", q);
        prefix[q + 1] = cnt;
        ew[q] = w ? w[ring] : 0.0; tbb[q] = t ? t[ring] : 0.0; spec[q] = sp ? sp[ring] : 0;
        if (cnt > 0 && !(tbb[q] > 0.0)) {
            if (spec[q] < 0 || spec[q] >= in->n_spec || !in->nfile || in->nfile[spec[q]] < 2 || in->nfile[spec[q]] > in->nfmax)
                return fail(C2D_EINVAL, "surface slot %d samples spectrum table %d, which is missing or has < 2 rows "
                            "(imcsurf2d_para.f:652-657 'Less than 2 lines of input read')", q, spec[q]);
        }
        return 0;
    };
    for (int j = 0; j < nz; j++) {
        if (int rc = slot(2 * j, in->nsurfi, in->ewsurfi, in->tbbi, in->spec_i, j)) return rc;
        if (int rc = slot(2 * j + 1, in->nsurfo, in->ewsurfo, in->tbbo, in->spec_o, j)) return rc;
    }
    for (int k = 0; k < nr; k++) {
        if (int rc = slot(2 * nz + 2 * k, in->nsurfu, in->ewsurfu, in->tbbu, in->spec_u, k)) return rc;
        if (int rc = slot(2 * nz + 2 * k + 1, in->nsurfl, in->ewsurfl, in->tbbl, in->spec_l, k)) return rc;
    }
    for (int q = 0; q < nslot; q++) prefix[q + 1] += prefix[q];
    c->surf_total = prefix[nslot];
    // tbbl(k,ti) is one COMMON array in the reference: the lower-face leak test of imcleak (imcleak2d.f:95) reads the
    // same values the source uses, so a tbbl given here replaces the one of c2d_step_begin
    if (in->tbbl) { if (int rc = upload(c, c->d_step + c->o_tbbl, in->tbbl, nr)) return rc; }
    if (int rc_m = check_mirror(c, in->tbbl)) return rc_m;
    // one blob: prefix | ew | tbb | E F P a1 I | spec | nfile
    const size_t nt = (size_t)in->n_spec * in->nfmax;
    const size_t bytes = (size_t)(nslot + 1) * 8 + (size_t)nslot * 16 + nt * 40 + (size_t)nslot * 4 + (size_t)std::max(in->n_spec, 1) * 4 + 64;
    if (bytes > c->surf_bytes) {
        cudaFree(c->d_surf);
        CU(cudaMalloc(&c->d_surf, bytes));
        c->surf_bytes = bytes;
    }
    std::vector<char> hb(bytes, 0);
    char *base = hb.data();
    size_t off = 0;
    auto put = [&](const void *src, size_t n) { size_t o = off; if (src && n) memcpy(base + off, src, n); off += (n + 7) & ~size_t(7); return o; };
    const size_t o_pre = put(prefix.data(), (size_t)(nslot + 1) * 8), o_ew = put(ew.data(), (size_t)nslot * 8), o_tbb = put(tbb.data(), (size_t)nslot * 8);
    const size_t o_E = put(in->E_file, nt * 8), o_F = put(in->F_file, nt * 8), o_P = put(in->P_file, nt * 8), o_a = put(in->a1, nt * 8), o_I = put(in->I_file, nt * 8);
    const size_t o_sp = put(spec.data(), (size_t)nslot * 4), o_nf = put(in->nfile, (size_t)in->n_spec * 4);
    CU(cudaMemcpyAsync(c->d_surf, base, off, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    char *D = static_cast<char *>(c->d_surf);
    Params &P = c->P;
    P.surf_prefix = reinterpret_cast<const int64_t *>(D + o_pre);
    P.ewsurf = reinterpret_cast<const double *>(D + o_ew); P.tbb_surf = reinterpret_cast<const double *>(D + o_tbb);
    P.sp_E = reinterpret_cast<const double *>(D + o_E); P.sp_F = reinterpret_cast<const double *>(D + o_F);
    P.sp_P = reinterpret_cast<const double *>(D + o_P); P.sp_a1 = reinterpret_cast<const double *>(D + o_a);
    P.sp_I = reinterpret_cast<const double *>(D + o_I);
    P.spec_surf = reinterpret_cast<const int *>(D + o_sp); P.sp_n = reinterpret_cast<const int *>(D + o_nf);
    P.sp_nfmax = in->nfmax; P.star_switch = in->star_switch;
    c->surf_open = true;
    return 0;
}

int c2d_imcsurf(c2d_ctx *h, int64_t first, int64_t count, int64_t stride)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    if (!c->surf_open) return fail(C2D_ESTATE, "c2d_surf_begin has not been called in this step");
    CU(cudaSetDevice(c->device));
    if (stride < 1 || first < 0) return fail(C2D_EINVAL, "bad shard first=%lld stride=%lld", (long long)first, (long long)stride);
    const int64_t total = c->surf_total;
    int64_t avail = (first < total) ? (total - first + stride - 1) / stride : 0;
    if (count < 0 || count > avail) count = avail;
    c->step_tracked = true; c->cascade_dirty = true;
    if (int rc = c->sigma_ensure()) return rc;
    if (int rc = launch_history_chunks(c, 2, first, count, stride)) return rc;
    c->histories += count;
    return 0;
}

int c2d_sync(c2d_ctx *h)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    CU(cudaSetDevice(c->device));
    int rc = drain_queue(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return check_fatal(c);
}

int c2d_tally_device(c2d_ctx *h, double **f, int64_t *nf, int64_t **i, int64_t *ni)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (f) *f = c->d_f64;
    if (nf) *nf = (int64_t)c->n_f64;
    if (i) *i = reinterpret_cast<int64_t *>(c->d_i64);
    if (ni) *ni = (int64_t)c->n_i64;
    return 0;
}

int c2d_tally_read(c2d_ctx *h, c2d_tallies *t)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !t) return fail(C2D_ESTATE, "context not initialised");
    CU(cudaSetDevice(c->device));
    c->h_f64.resize(c->n_f64); c->h_i64.resize(c->n_i64);
    unsigned long long cg[CNT_N];
    CU(cudaMemcpyAsync(c->h_f64.data(), c->d_f64, c->n_f64 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_i64.data(), c->d_i64, c->n_i64 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(cg, c->d_cnt_global, sizeof cg, cudaMemcpyDeviceToHost, c->stream));
    const c2d_dims &d = c->dims;
    if (t->fout_total)
        CU(cudaMemcpyAsync(t->fout_total, c->d_fout_total, (size_t)C2D_NMUMAX * d.nphtotal * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int nzone = c->nzone;
    auto cp = [&](double *dst, size_t off, size_t n) { if (dst) memcpy(dst, c->h_f64.data() + off, n * 8); };
    cp(t->edep, c->t_edep, nzone); cp(t->prdep, c->t_prdep, nzone); cp(t->ecens, c->t_ecens, nzone);
    cp(t->n_field, c->t_nfield, (size_t)nzone * d.nphfield); cp(t->n_ph, c->t_nph, (size_t)nzone * d.n_gg);
    cp(t->E_IC, c->t_EIC, d.num_nt); cp(t->fout, c->t_fout, (size_t)C2D_NMUMAX * d.nphtotal);
    cp(t->edout, c->t_edout, (size_t)C2D_NMUMAX * C2D_NLCMAX);
    cp(t->erlki, c->t_erlki, d.nz); cp(t->erlko, c->t_erlko, d.nz); cp(t->erlku, c->t_erlku, d.nr);
    cp(t->erlkl, c->t_erlkl, d.nr); cp(t->Ed_in, c->t_Edin, d.nr); cp(t->Ed_ref, c->t_Edref, d.nr);
    if (t->npcen) memcpy(t->npcen, c->h_i64.data() + c->i_npcen, (size_t)nzone * 8);
    if (t->nelectron) memcpy(t->nelectron, c->h_i64.data() + c->i_nel, (size_t)(d.num_nt + 2) * 8);
    for (int k = 0; k < 16; k++) t->counters[k] = (k < CNT_N) ? (int64_t)c->h_i64[c->i_cnt + k] : 0;
    t->counters[C2D_CNT_QPEAK] = c->q_peak;
    // before c2d_tally_reduce ran in this step the global counters are this rank's
    for (int k = 0; k < 16; k++) t->counters_global[k] = c->step_reduced && k < CNT_N ? (int64_t)cg[k] : t->counters[k];
    return 0;
}

#ifndef C2D_EMU
__global__ void k_add_into(double *dst, const double *src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = dst[i] + src[i];
}
#endif
static int add_into(Ctx *c, double *dst, const double *src, size_t n)
{
#ifndef C2D_EMU
    k_add_into<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dst, src, n);
    CU(cudaGetLastError());
#else
    (void)c;
    for (size_t i = 0; i < n; i++) dst[i] = dst[i] + src[i];
#endif
    return 0;
}

// The end-of-step reduction: cens_add_up + E_add_up + xec_add + graphics_collect (update2d.f:1931-2083, xec2d.f:325-394).
// In place on the device: every rank ends up with the sums (the reference reduces to the master and re-broadcasts what the
// workers need with FP_bcast).  counters[NDXOUT] and [NEVENTS] index this rank's census bank / event list, so the counter
// block is reduced out of place (c2d_tallies.counters_global).  The run-long escaping spectrum (the master's fout,
// xec2d.f:389-394: sum of the workers' never-zeroed copies) is accumulated here from the reduced per-step one.
int c2d_tally_reduce(c2d_ctx *h, c2d_tallies *t)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (int rc = c2d_sync(h)) return rc;
    if (!c->step_reduced) {
        if (comm_world(c) > 1) {
            c->timer_begin(8);
            int rc = comm_allreduce(c, c->d_f64, c->d_f64, c->n_f64, 0);
            if (!rc) rc = comm_allreduce(c, c->d_i64, c->d_i64, c->i_cnt, 1);
            if (!rc) rc = comm_allreduce(c, c->d_i64 + c->i_cnt, c->d_cnt_global, CNT_N, 1);
            c->timer_end();
            if (rc) return rc;
        } else {
            CU(cudaMemcpyAsync(c->d_cnt_global, c->d_i64 + c->i_cnt, CNT_N * 8, cudaMemcpyDeviceToDevice, c->stream));
        }
        if (int rc = add_into(c, c->d_fout_total, c->d_f64 + c->t_fout, (size_t)C2D_NMUMAX * c->dims.nphtotal)) return rc;
        c->step_reduced = true;
        // the step's census is final: sort it by (zone, energy class) for the next step's re-tracking (c2d_bank.cu) while
        // it is still this step's cost
        if (c->step_open && !c->bank_sorted) {
            if (int rc = latch_census(c)) return rc;
            const int sorted = bank_sort(c, c->n_census);
            if (sorted < 0) return C2D_ECUDA;
            c->bank_sorted = sorted == 1;
        }
        CU(cudaStreamSynchronize(c->stream));
    }
    return t ? c2d_tally_read(h, t) : 0;
}

int c2d_finalize(c2d_ctx *h) { return c2d_destroy(h); }

int c2d_census_count(c2d_ctx *h, int64_t *n)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited || !n) return fail(C2D_ESTATE, "context not initialised");
    CU(cudaSetDevice(c->device));
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->P.counters + CNT_NDXOUT, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n = (int64_t)v;
    return 0;
}

int c2d_census_export(c2d_ctx *h, double *dbuf, int32_t *ibuf, uint64_t *sid, int64_t cap, int64_t *n_out)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    int64_t n = 0;
    int rc = c2d_census_count(h, &n);
    if (rc) return rc;
    if (n > c->census_cap) return fail(C2D_ECENSUS, "too many photons: %lld > %lld", (long long)n, (long long)c->census_cap);
    if (n_out) *n_out = n;
    if (n > cap) return fail(C2D_EINVAL, "export buffer holds %lld records, census has %lld", (long long)cap, (long long)n);
    if (n == 0) return 0;
    const Bank &b = c->bank[c->bank_sorted ? c->cur_out ^ 1 : c->cur_out];
    std::vector<double> col(n);
    const double *cols[6] = {b.rpre, b.zpre, b.wmu, b.phi, b.ew, b.xnu};
    if (dbuf)
        for (int k = 0; k < 6; k++) {
            CU(cudaMemcpy(col.data(), cols[k], n * 8, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < n; i++) dbuf[6 * i + k] = col[i];
        }
    if (ibuf) {
        std::vector<uint32_t> bins(n), zone(n);
        std::vector<int32_t> seed(n);
        CU(cudaMemcpy(bins.data(), b.bins, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(zone.data(), b.zone, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(seed.data(), b.seed, n * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; i++) {
            ibuf[6 * i + 0] = bins[i] & 0xFFFF; ibuf[6 * i + 1] = (bins[i] >> 16) & 0xFF; ibuf[6 * i + 2] = bins[i] >> 24;
            ibuf[6 * i + 3] = zone[i] & 0xFFFF; ibuf[6 * i + 4] = zone[i] >> 16; ibuf[6 * i + 5] = seed[i];
        }
    }
    if (sid) CU(cudaMemcpy(sid, b.sid, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int c2d_census_import(c2d_ctx *h, const double *dbuf, const int32_t *ibuf, const uint64_t *sid, int64_t n)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (n < 0 || n > c->census_cap) return fail(C2D_ECENSUS, "too many photons needed to read: %lld > %lld (imcfield2d.f:84-86)", (long long)n, (long long)c->census_cap);
    if (n && (!dbuf || !ibuf)) return fail(C2D_EINVAL, "NULL census buffers");
    // a bad restart file must not become out-of-bounds device indices (zone in census_write, (jgpmu-1) nphtotal in escape_tally)
    for (int64_t i = 0; i < n; i++) {
        const int32_t *r = ibuf + 6 * i;
        const bool ok = r[0] >= 0 && r[0] <= c->dims.nphtotal && r[1] >= 0 && r[1] <= c->dims.nph_lc && r[2] >= 0 && r[2] <= c->dims.nmu &&
                        r[3] >= 1 && r[3] <= c->dims.nz && r[4] >= 1 && r[4] <= c->dims.nr && (r[2] >= 1 || (r[0] == 0 && r[1] == 0));
        if (!ok)
            return fail(C2D_EINVAL, "census record %lld out of range: jgpsp %d jgplc %d jgpmu %d jph %d kph %d (grid %d x %d, %d groups, %d bands, %d mu bins)",
                        (long long)i + 1, r[0], r[1], r[2], r[3], r[4], c->dims.nz, c->dims.nr, c->dims.nphtotal, c->dims.nph_lc, c->dims.nmu);
    }
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    // between steps the photons go where the last step left its census; right after c2d_step_begin (banks swapped, count
    // latched, nothing tracked yet) they are this step's census input
    const bool into_input = c->step_open && !c->step_tracked;
    c->bank_sorted = c->bank_sorted && into_input;                // a census imported between steps replaces the sorted copy
    Bank &b = into_input ? c->bank[c->cur_out ^ 1] : c->bank[c->cur_out];
    if (n) {
        std::vector<double> col(n);
        double *cols[6] = {b.rpre, b.zpre, b.wmu, b.phi, b.ew, b.xnu};
        for (int k = 0; k < 6; k++) {
            for (int64_t i = 0; i < n; i++) col[i] = dbuf[6 * i + k];
            CU(cudaMemcpy(cols[k], col.data(), n * 8, cudaMemcpyHostToDevice));
        }
        std::vector<uint32_t> bins(n), zone(n);
        std::vector<int32_t> seed(n);
        std::vector<uint64_t> sd(n);
        for (int64_t i = 0; i < n; i++) {
            bins[i] = (uint32_t)ibuf[6 * i] | ((uint32_t)ibuf[6 * i + 1] << 16) | ((uint32_t)ibuf[6 * i + 2] << 24);
            zone[i] = (uint32_t)ibuf[6 * i + 3] | ((uint32_t)ibuf[6 * i + 4] << 16);
            seed[i] = ibuf[6 * i + 5];
            // without a replay id (a census written by the reference), derive one from the record's seed and slot
            sd[i] = sid ? sid[i] : (0x63656E73ull << 32 | (uint64_t)i) * 0x9E3779B97F4A7C15ull + (uint64_t)(uint32_t)ibuf[6 * i + 5];
        }
        CU(cudaMemcpy(b.bins, bins.data(), n * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(b.zone, zone.data(), n * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(b.seed, seed.data(), n * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(b.sid, sd.data(), n * 8, cudaMemcpyHostToDevice));
    }
    unsigned long long v = (unsigned long long)n;
    if (into_input) { c->n_census = n; return 0; }
    CU(cudaMemcpy(c->P.counters + CNT_NDXOUT, &v, 8, cudaMemcpyHostToDevice));
    return 0;
}

int c2d_events_drain(c2d_ctx *h, c2d_event *buf, int64_t cap, int64_t *n_out)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    CU(cudaSetDevice(c->device));
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->P.counters + CNT_NEVENTS, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_out) *n_out = (int64_t)v;
    const int64_t stored = std::min<int64_t>((int64_t)v, c->event_cap);
    const int64_t n = std::min<int64_t>(stored, cap);
    if (n && buf) CU(cudaMemcpy(buf, c->d_events, (size_t)n * 7 * 8, cudaMemcpyDeviceToHost));
    if ((int64_t)v > c->event_cap) return fail(C2D_EEVENTS, "%llu escape records produced, event_capacity %lld", v, (long long)c->event_cap);
    return 0;
}

int c2d_comtot(c2d_ctx *h, int64_t n, const int32_t *zone, const double *xnu, int mode, double *comac)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->step_open) return fail(C2D_ESTATE, "c2d_step_begin has not been called");
    if (n < 0 || (n && (!zone || !xnu || !comac))) return fail(C2D_EINVAL, "NULL argument");
    if (mode != 0 && mode != 1) return fail(C2D_EINVAL, "mode %d", mode);
    if (mode == 1 && c->dims.sigma_mode != 1) return fail(C2D_EINVAL, "the context was created with sigma_mode 0: no table");
    for (int64_t i = 0; i < n; i++)
        if (zone[i] < 0 || zone[i] >= c->nzone || !(xnu[i] > 0.0)) return fail(C2D_EINVAL, "query %lld: zone %d xnu %g", (long long)i, zone[i], xnu[i]);
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    if (int rc = c->sigma_ensure()) return rc;
    int *d_zone = nullptr; double *d_x = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_zone, n * 4)); CU(cudaMalloc(&d_x, n * 8)); CU(cudaMalloc(&d_out, n * 8));
    CU(cudaMemcpyAsync(d_zone, zone, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_x, xnu, n * 8, cudaMemcpyHostToDevice, c->stream));
    Params p = c->P;
    p.sigma_mode = mode;
    launch_comtot_query(p, n, d_zone, d_x, d_out, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(comac, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_zone); cudaFree(d_x); cudaFree(d_out);
    return 0;
}

int c2d_pref_calc(c2d_ctx *h, int n_ref, double *E_ref, double *P_ref)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return fail(C2D_EINVAL, "NULL ctx");
    if (n_ref < 2 || !E_ref || !P_ref) return fail(C2D_EINVAL, "c2d_pref_calc: n_ref = %d, E_ref / P_ref must not be NULL", n_ref);
    CU(cudaSetDevice(c->device));
    double *d = nullptr;
    const size_t n = (size_t)n_ref * n_ref + n_ref;
    CU(cudaMalloc(&d, n * 8));
    launch_pref_calc(n_ref, d, d + n_ref, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(E_ref, d, (size_t)n_ref * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(P_ref, d + n_ref, (size_t)n_ref * n_ref * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

int c2d_selftest_gap(c2d_ctx *h, int64_t n, uint64_t seed, int64_t *mismatches, int64_t *slow_path)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    if (n < 0 || !mismatches) return fail(C2D_EINVAL, "NULL argument");
    CU(cudaSetDevice(c->device));
    unsigned long long *d = nullptr, host[2] = {0, 0};
    CU(cudaMalloc(&d, 16));
    CU(cudaMemsetAsync(d, 0, 16, c->stream));
    if (n > 0) launch_selftest_gap(c->P, n, seed, d, c->n_sm, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, d, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    *mismatches = (int64_t)host[0];
    if (slow_path) *slow_path = (int64_t)host[1];
    return 0;
}

int c2d_kernel_time_ms(c2d_ctx *h, int which, float *ms, int64_t *launches)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return fail(C2D_EINVAL, "NULL ctx");
    if (which == -1) { c->timing = true; return 0; }          // enable
    if (which == -5) { c->timer_collect(); c->timing = false; return 0; }   // disable
    if (which == -3 || which == -4) {                         // bracket a region on the library's stream -> family 7
        CU(cudaSetDevice(c->device));
        if (!c->ev_mark[0]) { CU(cudaEventCreate(&c->ev_mark[0])); CU(cudaEventCreate(&c->ev_mark[1])); }
        CU(cudaEventRecord(c->ev_mark[which == -3 ? 0 : 1], c->stream));
        if (which == -4) {
            float t = 0;
            CU(cudaEventSynchronize(c->ev_mark[1]));
            CU(cudaEventElapsedTime(&t, c->ev_mark[0], c->ev_mark[1]));
            c->k_ms[7] += t; c->k_launches[7]++;
        }
        return 0;
    }
    if (which == -2) { c->timer_collect(); for (int k = 0; k < Ctx::kFamilies; k++) { c->k_ms[k] = 0; c->k_launches[k] = 0; } return 0; }   // reset
    if (which < 0 || which >= Ctx::kFamilies) return fail(C2D_EINVAL, "kernel family %d", which);
    c->timer_collect();
    if (ms) *ms = (float)c->k_ms[which];
    if (launches) *launches = c->k_launches[which];
    return 0;
}

void c2d_pack_bin_first(const double *common, int nbin_pad, int jmax, int kmax, int nbin, int nz, int nr, double *dense)
{   // COMMON a(nbin_pad, jmax, kmax) column-major -> dense [zone][bin]
    (void)kmax;
    for (int j = 0; j < nz; j++)
        for (int k = 0; k < nr; k++)
            for (int i = 0; i < nbin; i++)
                dense[((size_t)j * nr + k) * nbin + i] = common[((size_t)k * jmax + j) * nbin_pad + i];
}
void c2d_pack_bin_last(const double *common, int jmax, int kmax, int nbin, int nz, int nr, double *dense)
{   // COMMON a(jmax, kmax, nbin) column-major -> dense [zone][bin]
    for (int j = 0; j < nz; j++)
        for (int k = 0; k < nr; k++)
            for (int i = 0; i < nbin; i++)
                dense[((size_t)j * nr + k) * nbin + i] = common[((size_t)i * kmax + k) * jmax + j];
}
void c2d_unpack_bin_first(const double *dense, int nbin_pad, int jmax, int kmax, int nbin, int nz, int nr, double *common)
{
    (void)kmax;
    for (int j = 0; j < nz; j++)
        for (int k = 0; k < nr; k++)
            for (int i = 0; i < nbin; i++)
                common[((size_t)k * jmax + j) * nbin_pad + i] = dense[((size_t)j * nr + k) * nbin + i];
}
void c2d_unpack_bin_last(const double *dense, int jmax, int kmax, int nbin, int nz, int nr, double *common)
{
    for (int j = 0; j < nz; j++)
        for (int k = 0; k < nr; k++)
            for (int i = 0; i < nbin; i++)
                common[((size_t)i * kmax + k) * jmax + j] = dense[((size_t)j * nr + k) * nbin + i];
}

}  // extern "C"
