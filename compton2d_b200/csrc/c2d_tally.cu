// c2d_tally.cu -- stand-alone tally-commit kernels fed a pre-generated event list (c2d_tally_bench).
//
// The history kernels commit their tallies inline (one FP64 atomic per zone crossing / census event); these kernels
// are the same tallies in their bandwidth-bound form -- an event list streamed from HBM, reduced
// warp-level with __match_any_sync / __shfl, accumulated in shared-memory bins and committed to HBM once per block --
// and exist to measure that form against the HBM roofline (BASELINE.json: ">= 60 % of HBM roofline on the tally kernel"):
//   kind 0  deposition      edep(zone) += de, prdep(zone) += dp                  (imctrk2d.f:465-466)
//                           algorithmic bytes / event: 8 (index) + 16 (values) + 32 (two FP64 RMW) = 56
//   kind 1  census          record copied into the bank (imctrk2d.f:566-580) + npcen, ecens, n_field (l.541-564)
//                           algorithmic bytes / event: 72 (record) + 40 (tallies) = 112
// kind | 0x10: no warp-level pre-reduction; kind | 0x20: events sorted by zone (default: zones hashed).
// Events are generated on the device from their index by a fixed integer hash, so tests can rebuild them in numpy.
#include "c2d_ctx.h"

#include <algorithm>

namespace c2d {

__host__ __device__ inline uint32_t tally_hash(uint64_t i) { return (uint32_t)((i * 0x9E3779B97F4A7C15ull) >> 33); }

__global__ void k_tally_gen(long long n, int nzone, int nphfield, int sorted, long long *zone, double *de, double *dp, int *bin)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t h = tally_hash((uint64_t)i);
        zone[i] = sorted ? (long long)(((unsigned long long)i * (unsigned long long)nzone) / (unsigned long long)n)
                         : (long long)(h % (uint32_t)nzone);
        de[i] = 1.0 + 1.e-3 * (double)(h & 1023u);
        dp[i] = 0.5 * de[i];
        if (bin) bin[i] = (int)((h >> 10) % (uint32_t)nphfield);
    }
}

#ifndef C2D_EMU
// warp-level pre-reduction: when all 32 lanes hit the same bin (events sorted by zone, as a zone-sorted bank produces
// them) one butterfly sum and one shared-memory atomic pair replace 64 atomics; mixed warps go lane by lane
__device__ __forceinline__ void warp_commit2(double *a, double *b, int key, double va, double vb)
{
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (peers == 0xffffffffu) {
        va = warp_sum(va); vb = warp_sum(vb);
        if ((threadIdx.x & 31) == 0) { atomicAdd(a + key, va); atomicAdd(b + key, vb); }
    } else {
        atomicAdd(a + key, va); atomicAdd(b + key, vb);
    }
}

// kind 0: shared-memory bins [2][nzone] per block, flushed once
__global__ void __launch_bounds__(256) k_tally_deposit(long long n, int nzone, const long long *zone, const double *de,
                                                       const double *dp, double *edep, double *prdep, int aggregate)
{
    extern __shared__ double sh[];
    double *s_e = sh, *s_p = sh + nzone;
    for (int i = threadIdx.x; i < 2 * nzone; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (n + 31) & ~31ll;                   // whole warps stay in the loop (match_any needs all lanes)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool live = i < n;                                  // a dead lane adds 0 to bin 0
        const int z = live ? (int)zone[i] : 0;
        const double a = live ? de[i] : 0.0, b = live ? dp[i] : 0.0;
        if (aggregate) warp_commit2(s_e, s_p, z, a, b);
        else if (live) { atomicAdd(s_e + z, a); atomicAdd(s_p + z, b); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nzone; i += blockDim.x) {
        if (s_e[i] != 0.0) atomicAdd(edep + i, s_e[i]);
        if (s_p[i] != 0.0) atomicAdd(prdep + i, s_p[i]);
    }
}

// kind 1: census commit.  Records are synthesised from the event (the bank write is the point), tallies go through
// shared-memory bins for npcen / ecens and straight to HBM for the diffuse n_field (400 bins per zone)
__global__ void __launch_bounds__(256) k_tally_census(long long n, int nzone, int nphfield, const long long *zone, const double *de,
                                                      const double *dp, const int *bin, Bank out, double *ecens,
                                                      unsigned long long *npcen, double *n_field)
{
    extern __shared__ double sh[];
    double *s_e = sh;
    unsigned *s_n = reinterpret_cast<unsigned *>(sh + nzone);
    for (int i = threadIdx.x; i < nzone; i += blockDim.x) { s_e[i] = 0.0; s_n[i] = 0u; }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int z = (int)zone[i];
        const double ew = de[i], x = dp[i];
        out.rpre[i] = ew; out.zpre[i] = x; out.wmu[i] = 0.5; out.phi[i] = 1.0; out.ew[i] = ew; out.xnu[i] = x;
        out.bins[i] = 1u; out.zone[i] = (uint32_t)z; out.seed[i] = (int)(i & 0xFFFF); out.sid[i] = (uint64_t)i;
        atomicAdd(s_e + z, ew);
        atomicAdd(s_n + z, 1u);
        atomicAdd(n_field + (size_t)z * nphfield + bin[i], 6.25e8 * ew / x);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nzone; i += blockDim.x) {
        if (s_n[i]) { atomicAdd(ecens + i, s_e[i]); atomicAdd(npcen + i, (unsigned long long)s_n[i]); }
    }
}
#endif

}  // namespace c2d

using namespace c2d;

#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(C2D_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" int c2d_tally_bench(c2d_ctx *h, int64_t n_events, int kind, float *ms, int64_t *bytes)
{
#ifdef C2D_EMU
    (void)h; (void)n_events; (void)kind; (void)ms; (void)bytes;
    return fail(C2D_EINVAL, "c2d_tally_bench needs the device build");
#else
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !c->inited) return fail(C2D_ESTATE, "context not initialised");
    const int mode = kind & 0xF, aggregate = (kind & 0x10) ? 0 : 1;          // kind | 0x10: no warp-level pre-reduction
    if (n_events < 1 || (mode != 0 && mode != 1)) return fail(C2D_EINVAL, "n_events = %lld, kind = %d", (long long)n_events, kind);
    if (mode == 1 && n_events > c->census_cap) return fail(C2D_ECENSUS, "census bench needs census_capacity >= %lld", (long long)n_events);
    CUT(cudaSetDevice(c->device));
    const int nzone = c->nzone, nf = c->dims.nphfield;
    long long *d_zone = nullptr; double *d_de = nullptr, *d_dp = nullptr; int *d_bin = nullptr;
    CUT(cudaMalloc(&d_zone, n_events * 8)); CUT(cudaMalloc(&d_de, n_events * 8)); CUT(cudaMalloc(&d_dp, n_events * 8));
    if (mode == 1) CUT(cudaMalloc(&d_bin, n_events * 4));
    const int grid = c->n_sm * 8;
    k_tally_gen<<<grid, 256, 0, c->stream>>>(n_events, nzone, nf, (kind & 0x20) ? 1 : 0, d_zone, d_de, d_dp, d_bin);
    CUT(cudaGetLastError());
    const size_t smem = (size_t)2 * nzone * sizeof(double);
    if (smem > 200 * 1024) return fail(C2D_EINVAL, "tally bench: %d zones do not fit the shared-memory bins", nzone);
    CUT(cudaFuncSetAttribute(k_tally_deposit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUT(cudaFuncSetAttribute(k_tally_census, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool was_timing = c->timing;
    c->timing = true;
    c->timer_collect();
    const double ms0 = c->k_ms[5];
    c->timer_begin(5);
    if (mode == 0)
        k_tally_deposit<<<c->n_sm * 4, 256, smem, c->stream>>>(n_events, nzone, d_zone, d_de, d_dp, c->P.edep, c->P.prdep, aggregate);
    else
        k_tally_census<<<c->n_sm * 4, 256, smem, c->stream>>>(n_events, nzone, nf, d_zone, d_de, d_dp, d_bin, c->P.out, c->P.ecens,
                                                               c->P.npcen, c->P.n_field);
    c->timer_end();
    CUT(cudaGetLastError());
    c->timer_collect();
    c->timing = was_timing;
    if (ms) *ms = (float)(c->k_ms[5] - ms0);
    if (bytes) *bytes = n_events * (mode == 0 ? 56 : 112);
    cudaFree(d_zone); cudaFree(d_de); cudaFree(d_dp); cudaFree(d_bin);
    return 0;
#endif
}
