// c2d_bank.cu -- the census bank is kept sorted by (zone, energy class) between time steps.
//
// The reference's census is an append-only buffer (dbufout / ibufout, imctrk2d.f:566-580) that field_calc walks front to
// back (imcfield2d.f:100-143): order means nothing there.  On the GPU it decides what a warp of k_history<census> looks
// like: 32 photons of one zone and one energy decade read the same rows of the step tables (kappa_tot, the sigma table's
// coefficients and noise weights), cross into the same few neighbours and walk noise loops of the same length, and their
// per-zone tallies collapse into one shuffle sum.  So before a step re-tracks the bank it is counting-sorted:
//   k_bank_hist     key = zone * kSortE + sort_class(xnu) per photon, histogram (shared-memory copy per block)
//   k_bank_scan     exclusive prefix sum of the histogram (one block)
//   k_bank_scatter  every photon moves to the next free slot of its key's range in the other bank (warp-aggregated cursor
//                   atomics; the unsorted bank is already clustered by zone, so the writes land in a few dense streams)
// 68 B read + 68 B written per photon.  Random-number streams are keyed by ancestry, not by slot: results do not change.
#include "c2d_ctx.h"

#include <algorithm>

namespace c2d {

#ifndef C2D_EMU
constexpr int kHistSmemKeys = 8192;

__device__ __forceinline__ int bank_key(const Bank &b, long long i, int nr)
{
    const uint32_t zz = b.zone[i];
    const int jph = (int)(zz & 0xFFFF), kph = (int)(zz >> 16);
    return ((jph - 1) * nr + (kph - 1)) * kSortE + sort_class(b.xnu[i]);
}

__global__ void __launch_bounds__(256) k_bank_hist(Bank b, long long n, int nr, int nkeys, unsigned long long *hist)
{
    extern __shared__ unsigned s_hist[];
    const bool use_s = nkeys <= kHistSmemKeys;
    if (use_s) {
        for (int k = threadIdx.x; k < nkeys; k += blockDim.x) s_hist[k] = 0u;
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int key = bank_key(b, i, nr);
        if (use_s) atomicAdd(&s_hist[key], 1u); else atomicAdd(hist + key, 1ull);
    }
    if (use_s) {
        __syncthreads();
        for (int k = threadIdx.x; k < nkeys; k += blockDim.x)
            if (s_hist[k]) atomicAdd(hist + k, (unsigned long long)s_hist[k]);
    }
}

// hist -> exclusive offsets (in place), one block of 1024 threads
__global__ void __launch_bounds__(1024) k_bank_scan(unsigned long long *hist, int nkeys)
{
    __shared__ unsigned long long s_part[1024];
    const int per = (nkeys + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, nkeys);
    unsigned long long sum = 0;
    for (int k = lo; k < hi; k++) sum += hist[k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                         // Hillis-Steele inclusive scan of the 1024 partial sums
        const unsigned long long v = (threadIdx.x >= (unsigned)o) ? s_part[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;
    for (int k = lo; k < hi; k++) { const unsigned long long c = hist[k]; hist[k] = run; run += c; }
}

__global__ void __launch_bounds__(256) k_bank_scatter(Bank src, Bank dst, long long n, int nr, unsigned long long *cursor)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int key = valid ? bank_key(src, i, nr) : -1;
    // one cursor atomic per (warp, key)
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (valid && lane == leader) base = atomicAdd(cursor + key, (unsigned long long)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!valid) return;
    const unsigned long long d = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
    dst.rpre[d] = src.rpre[i]; dst.zpre[d] = src.zpre[i]; dst.wmu[d] = src.wmu[i]; dst.phi[d] = src.phi[i];
    dst.ew[d] = src.ew[i]; dst.xnu[d] = src.xnu[i];
    dst.bins[d] = src.bins[i]; dst.zone[d] = src.zone[i]; dst.seed[d] = src.seed[i]; dst.sid[d] = src.sid[i];
}
#endif

// Sort bank[cur_out] (n photons) into bank[cur_out ^ 1].  Returns 1 when the sorted copy exists, 0 when the sort was skipped.
int bank_sort(Ctx *c, long long n)
{
#ifdef C2D_EMU
    (void)c; (void)n;
    return 0;
#else
    if (n <= 0 || !c->sort_bank) return 0;
    const int nkeys = c->nzone * kSortE;
    if (!c->d_sortkeys) {
        if (cudaMalloc(&c->d_sortkeys, (size_t)nkeys * 8) != cudaSuccess) { fail(C2D_ECUDA, "cudaMalloc(sort keys) failed"); return -1; }
    }
    cudaMemsetAsync(c->d_sortkeys, 0, (size_t)nkeys * 8, c->stream);
    c->timer_begin(11);
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)c->n_sm * 8);
    const size_t smem = (nkeys <= kHistSmemKeys) ? (size_t)nkeys * 4 : 0;
    k_bank_hist<<<blocks, 256, smem, c->stream>>>(c->bank[c->cur_out], n, c->dims.nr, nkeys, c->d_sortkeys);
    k_bank_scan<<<1, 1024, 0, c->stream>>>(c->d_sortkeys, nkeys);
    k_bank_scatter<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->bank[c->cur_out], c->bank[c->cur_out ^ 1], n, c->dims.nr, c->d_sortkeys);
    c->timer_end();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fail(C2D_ECUDA, "bank sort launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 1;
#endif
}

}  // namespace c2d
