// c2d_sigma.cu -- sigma_mode 1: the per-step table of the macroscopic Compton opacity comtot(zone, E)
// (comtot2d.f:1-334, live branch icoms = 6, l.219-247) and the c2d_comtot entry point.
//
// comtot sums, over the electron bins i, sigma_i(x) f_nt(zone, i) dgamma_i, where sigma_i is the Thomson series
// when x gamma_i (1 + beta_i) < 0.01 and the angle-averaged Klein-Nishina expression (two dilogarithms) otherwise.
// sigma_i depends on the bin and the photon energy only -- not on the zone and not on the time step -- so:
//   once, at c2d_init:   Kn[i][node] = exact-branch sigma_i at the Chebyshev nodes of every interval of ln x on
//                        which the set of exact-branch bins {i >= m} is constant (k_sigma_kn);
//   every step:          W[z][i] = f_nt dgamma, prefix sums A_z[m], B_z[m] of the Thomson-branch bins
//                        (k_sigma_prefix); node values T[z][node] = sum_i Kn[i][node] W[z][i] (k_sigma_nodes, an
//                        FP64 product nzone x (num_nt-1) x n_nodes); Chebyshev coefficients per interval (k_sigma_coef).
// A history then needs one bisection, one log and an 8-term Clenshaw sum per zone instead of 199 bins x 2
// dilogarithms (comtot_table in c2d_device.cuh).  The bin-by-bin evaluation stays available as sigma_mode 0 and
// through c2d_comtot(mode 0); tests/test_gpu_sigma.py holds the two together.
#include "c2d_ctx.h"

#include <algorithm>
#include <cmath>

namespace c2d {

constexpr int kSigDeg = 8;          // nodes (= coefficients) per interval

__global__ void k_sigma_kn(const __grid_constant__ Params p, const double *xs, const int *ms, int n_nodes, double *Kn)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_nodes * (p.num_nt - 1)) return;
    const int i = (int)(idx / n_nodes), node = (int)(idx % n_nodes);
    Kn[idx] = (i >= ms[node]) ? sigma_kn_smooth(p, i, xs[node]) : 0.0;
}

__global__ void k_sigma_prefix(const __grid_constant__ Params p, double *W, double *AB, double *NB, int *IF, double *CW)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= p.nzone) return;
    const int N = p.num_nt;
    const double *f = p.f_nt + (size_t)z * N;
    double A = 0.0, B = 0.0;
    for (int i = 0; i < N - 1; i++) {
        double *ab = AB + ((size_t)z * N + i) * 2;
        ab[0] = A; ab[1] = B;
        const double w = f[i] * p.dgam[i];
        W[(size_t)z * N + i] = w;
        A = A + w;
        B = B + p.gam0[i] * w;
    }
    double *ab = AB + ((size_t)z * N + (N - 1)) * 2;
    ab[0] = A; ab[1] = B;
    W[(size_t)z * N + (N - 1)] = 0.0;
    // bounds of the rounding-noise terms (comtot_table_zone): suffix sums of (1.125 cb_i + 9 pref_i / a1_i) |W_i|, and the
    // first bin that holds electrons
    double nb = 0.0;
    int first = N - 1;
    NB[(size_t)z * N + (N - 1)] = 0.0;
    for (int i = N - 2; i >= 0; i--) {
        const double w = fabs(W[(size_t)z * N + i]);
        nb = nb + (1.125 * p.kn_cb[i] + 9.0 * p.kn_pref[i] / p.kn_a1[i]) * w;
        NB[(size_t)z * N + i] = nb;
        CW[(size_t)z * N + i] = p.kn_cb[i] * W[(size_t)z * N + i];
        if (w > 0.0) first = i;
    }
    CW[(size_t)z * N + (N - 1)] = 0.0;
    IF[z] = first;
}

constexpr int kSigZG = 4;           // zones per thread of k_sigma_nodes
__global__ void k_sigma_nodes(const __grid_constant__ Params p, const double *W, const double *Kn, int n_nodes, double *T)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ngrp = (p.nzone + kSigZG - 1) / kSigZG;
    if (idx >= (long long)n_nodes * ngrp) return;
    const int node = (int)(idx % n_nodes), z0 = (int)(idx / n_nodes) * kSigZG;
    const int N = p.num_nt;
    double acc[kSigZG];
    const double *w[kSigZG];
#pragma unroll
    for (int q = 0; q < kSigZG; q++) { acc[q] = 0.0; w[q] = W + (size_t)((z0 + q < p.nzone) ? z0 + q : z0) * N; }
    for (int i = 0; i < N - 1; i++) {
        const double k = Kn[(size_t)i * n_nodes + node];
#pragma unroll
        for (int q = 0; q < kSigZG; q++) acc[q] = acc[q] + k * w[q][i];
    }
#pragma unroll
    for (int q = 0; q < kSigZG; q++)
        if (z0 + q < p.nzone) T[(size_t)(z0 + q) * n_nodes + node] = acc[q];
}

// node values -> Chebyshev coefficients, M[k][n] = (2 - [k == 0]) / 8 cos(pi k (n + 1/2) / 8)
__global__ void k_sigma_coef(const double *T, const double *M, long long n_intervals_total, double *coef)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_intervals_total) return;
    double v[kSigDeg];
#pragma unroll
    for (int n = 0; n < kSigDeg; n++) v[n] = T[idx * kSigDeg + n];
#pragma unroll
    for (int k = 0; k < kSigDeg; k++) {
        double c = 0.0;
#pragma unroll
        for (int n = 0; n < kSigDeg; n++) c = fma(M[k * kSigDeg + n], v[n], c);
        coef[idx * kSigDeg + k] = c;
    }
}

// c2d_comtot: comac for n (zone, xnu) pairs; mode 0 = bin by bin, 1 = table
__global__ void k_comtot_query(const __grid_constant__ Params p, long long n, const int *zone, const double *xnu, double *out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    SigCtx c; c.j = -2;
    if (p.sigma_mode == 1) sig_begin(p, xnu[idx], c);
    out[idx] = comtot_one(p, zone[idx], xnu[idx], c);
}

#ifndef C2D_EMU
// nth2d's guide (c2d_device.cuh): for every zone and every 1/kNtGuide of rnum the bracket [lo, hi] of the bisection
//   lo = first i with Pnt(i) > b / kNtGuide, hi = first i with Pnt(i) > (b + 1) / kNtGuide   (the answer is monotone in rnum).
// A zone whose Pnt(1 .. num_nt - 1) is not non-decreasing gets the full range [2, num_nt]: the plain bisection.
__device__ __forceinline__ int nt_first_above(const double *P, int num_nt, double x)
{
    int lo = 2, hi = num_nt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P[mid - 1] > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__global__ void __launch_bounds__(kNtGuide) k_nt_guide(const __grid_constant__ Params p, uint32_t *guide)
{
    __shared__ int s_bad;
    const int zi = blockIdx.x, b = threadIdx.x;
    const double *P = p.Pnt + (size_t)zi * p.num_nt;
    if (b == 0) s_bad = 0;
    __syncthreads();
    for (int i = 1 + b; i < p.num_nt - 1; i += kNtGuide)
        if (P[i] < P[i - 1]) s_bad = 1;
    __syncthreads();
    int lo = 2, hi = p.num_nt;
    if (!s_bad) {
        lo = nt_first_above(P, p.num_nt, (double)b / kNtGuide);
        hi = nt_first_above(P, p.num_nt, (double)(b + 1) / kNtGuide);
    }
    guide[(size_t)zi * kNtGuide + b] = (uint32_t)lo | ((uint32_t)hi << 16);
}
void launch_nt_guide(const Params &p, uint32_t *guide, cudaStream_t st)
{
    k_nt_guide<<<p.nzone, kNtGuide, 0, st>>>(p, guide);
}
void launch_sigma_kn(const Params &p, const double *xs, const int *ms, int n_nodes, double *Kn, cudaStream_t st)
{
    const long long n = (long long)n_nodes * (p.num_nt - 1);
    k_sigma_kn<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, xs, ms, n_nodes, Kn);
}
void launch_sigma_build(const Params &p, double *W, double *AB, const double *Kn, int n_nodes, double *T, const double *M,
                        double *coef, cudaStream_t st)
{
    k_sigma_prefix<<<(p.nzone + 63) / 64, 64, 0, st>>>(p, W, AB, const_cast<double *>(p.sig_NB), const_cast<int *>(p.sig_if), const_cast<double *>(p.sig_CW));
    const long long n = (long long)n_nodes * ((p.nzone + kSigZG - 1) / kSigZG);
    k_sigma_nodes<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, W, Kn, n_nodes, T);
    const long long ni = (long long)p.nzone * (n_nodes / kSigDeg);
    k_sigma_coef<<<(unsigned)((ni + 127) / 128), 128, 0, st>>>(T, M, ni, coef);
}
void launch_comtot_query(const Params &p, long long n, const int *zone, const double *xnu, double *out, cudaStream_t st)
{
    if (n > 0) k_comtot_query<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, n, zone, xnu, out);
}
#endif

}  // namespace c2d
