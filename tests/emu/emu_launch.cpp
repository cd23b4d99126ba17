// tests/emu/emu_launch.cpp -- TEST INFRASTRUCTURE ONLY: serial "launches" of the kernels for the g++ build.
#include "cuda_runtime.h"
EmuDim blockIdx, threadIdx, blockDim, gridDim;
#include "../../compton2d_b200/csrc/c2d_track.cu"
#include "../../compton2d_b200/csrc/c2d_sigma.cu"
#include "../../compton2d_b200/csrc/c2d_gen.cu"

namespace c2d {
template <int SRC> static void run_hist(const Params &p, long long first, long long n, long long stride)
{
    blockDim.x = 128;
    const long long blocks = (n + 127) / 128;
    for (long long b = 0; b < blocks; b++)
        for (unsigned t = 0; t < 128; t++) {
            blockIdx.x = (unsigned)b; threadIdx.x = t;
            k_history<SRC>(p, first, n, stride);
        }
}
void launch_history_volume(const Params &p, long long first, long long n, long long stride, cudaStream_t) { if (n > 0) run_hist<0>(p, first, n, stride); }
void launch_history_surface(const Params &p, long long first, long long n, long long stride, cudaStream_t) { if (n > 0) run_hist<2>(p, first, n, stride); }
void launch_history_census(const Params &p, long long first, long long n, long long stride, cudaStream_t) { if (n > 0) run_hist<1>(p, first, n, stride); }
void launch_absdraw(const Params &p, const AbsTask *q, const unsigned long long *n_dev, int, cudaStream_t)
{
    blockDim.x = 1; gridDim.x = 1; blockIdx.x = 0; threadIdx.x = 0;       // one serial "warp" walks the whole queue
    k_absdraw(p, q, n_dev);
}
template <class F> static void run_grid(long long n, F body)
{
    blockDim.x = 128;
    const long long blocks = (n + 127) / 128;
    for (long long b = 0; b < blocks; b++)
        for (unsigned t = 0; t < 128; t++) {
            blockIdx.x = (unsigned)b; threadIdx.x = t;
            body();
        }
}
void launch_selftest_gap(const Params &, long long, unsigned long long, unsigned long long *, int, cudaStream_t) {}   // one path only in this build
void launch_children(const Params &p, const S3Map &m, long long n_mine, int, unsigned *, unsigned long long, unsigned long long *, cudaStream_t)
{
    for (long long i = 0; i < n_mine; i++) children_one(p, m, i, nullptr, 0ull, nullptr);     // no team kernels in the serial build
}
void launch_split3_attempts(const Params &p, const S3Map &m, long long i0, long long n_sub, FlightTask *out, int max_rounds,
                            unsigned *hard, unsigned long long *hard_count, cudaStream_t)
{   // kWarp == 1 in this build: one serial "warp" per sub-child, no round limit
    (void)max_rounds;
    if (n_sub > 0) run_grid(n_sub, [&] { k_split3_attempts(p, m, i0, n_sub, out, 0, hard, hard_count); });
}
void launch_children_hard(const Params &, const S3Map &, const unsigned *, unsigned long long, GridItem *, unsigned long long *, int, int, int, cudaStream_t) {}
int launch_children_ultra(const Params &, const S3Map &, const GridItem *, const unsigned long long *, unsigned long long *, int, cudaStream_t) { return 0; }
void launch_split3_hard(const Params &, const S3Map &, long long, const unsigned *, GridItem *, unsigned long long *, FlightTask *, unsigned long long,
                        int, int, int, cudaStream_t) {}
int launch_split3_ultra(const Params &, const S3Map &, long long, const GridItem *, const unsigned long long *, FlightTask *, unsigned long long *, int,
                        cudaStream_t) { return 0; }
void launch_split3_flights(const Params &p, const FlightTask *fl, long long n, cudaStream_t)
{
    if (n > 0) run_grid(n, [&] { k_split3_flights(p, fl, n); });
}
void launch_sigma_kn(const Params &p, const double *xs, const int *ms, int n_nodes, double *Kn, cudaStream_t)
{
    run_grid((long long)n_nodes * (p.num_nt - 1), [&] { k_sigma_kn(p, xs, ms, n_nodes, Kn); });
}
void launch_sigma_build(const Params &p, double *W, double *AB, const double *Kn, int n_nodes, double *T, const double *M,
                        double *coef, cudaStream_t)
{
    run_grid(p.nzone, [&] { k_sigma_prefix(p, W, AB, const_cast<double *>(p.sig_NB), const_cast<int *>(p.sig_if), const_cast<double *>(p.sig_CW)); });
    run_grid((long long)n_nodes * ((p.nzone + kSigZG - 1) / kSigZG), [&] { k_sigma_nodes(p, W, Kn, n_nodes, T); });
    const long long ni = (long long)p.nzone * (n_nodes / kSigDeg);
    run_grid(ni, [&] { k_sigma_coef(T, M, ni, coef); });
}
void launch_volume_em(const GenParams &q, cudaStream_t)
{
    run_grid((long long)q.nzone * q.nv, [&] { k_volume_sy(q); });
    blockDim.x = 64;
    for (int b = 0; b < (q.nzone + 63) / 64; b++)
        for (unsigned t = 0; t < 64; t++) { blockIdx.x = (unsigned)b; threadIdx.x = t; k_volume_cdf(q); }
}
void launch_pref_calc(int n_ref, double *E_ref, double *P_ref, cudaStream_t)
{
    run_grid(n_ref, [&] { k_pref_calc(n_ref, E_ref, P_ref); });
}
void launch_comtot_query(const Params &p, long long n, const int *zone, const double *xnu, double *out, cudaStream_t)
{
    if (n > 0) run_grid(n, [&] { k_comtot_query(p, n, zone, xnu, out); });
}
}  // namespace c2d
