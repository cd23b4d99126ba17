// c2d_host.h -- host-side declarations shared by the translation units of libcompton2d_b200.so
#pragma once
#include "c2d_device.cuh"

namespace c2d {

void launch_history_volume(const Params &p, long long first, long long n, long long stride, cudaStream_t st);
void launch_history_surface(const Params &p, long long first, long long n, long long stride, cudaStream_t st);
void launch_history_census(const Params &p, long long first, long long n, long long stride, cudaStream_t st);
void launch_absdraw(const Params &p, const AbsTask *q, const unsigned long long *n_dev, int n_sm, cudaStream_t st);
void launch_selftest_gap(const Params &p, long long n, unsigned long long seed, unsigned long long *out, int n_sm, cudaStream_t st);
void launch_children(const Params &p, const S3Map &m, long long n_mine, int n_sm, unsigned *hard, unsigned long long hard_cap,
                     unsigned long long *hard_count, cudaStream_t st);
void launch_children_hard(const Params &p, const S3Map &m, const unsigned *hard, unsigned long long hard_cap, GridItem *ultra,
                          unsigned long long *counts, int max_rounds, int drain_rounds, int n_sm, cudaStream_t st);
int launch_children_ultra(const Params &p, const S3Map &m, const GridItem *ultra, const unsigned long long *n_ultra, unsigned long long *g_win,
                          int n_sm, cudaStream_t st);
void launch_split3_attempts(const Params &p, const S3Map &m, long long i0, long long n_sub, FlightTask *out, int max_rounds,
                            unsigned *hard, unsigned long long *hard_count, cudaStream_t st);
void launch_split3_hard(const Params &p, const S3Map &m, long long i0, const unsigned *hard, GridItem *ultra, unsigned long long *counts,
                        FlightTask *out, unsigned long long first_attempt, int max_rounds, int drain_rounds, int n_sm, cudaStream_t st);
int launch_split3_ultra(const Params &p, const S3Map &m, long long i0, const GridItem *ultra, const unsigned long long *n_ultra, FlightTask *out,
                        unsigned long long *g_win, int n_sm, cudaStream_t st);
void launch_split3_flights(const Params &p, const FlightTask *fl, long long n, cudaStream_t st);
void launch_nt_guide(const Params &p, uint32_t *guide, cudaStream_t st);
void launch_sigma_kn(const Params &p, const double *xs, const int *ms, int n_nodes, double *Kn, cudaStream_t st);
void launch_sigma_build(const Params &p, double *W, double *AB, const double *Kn, int n_nodes, double *T, const double *M,
                        double *coef, cudaStream_t st);
void launch_comtot_query(const Params &p, long long n, const int *zone, const double *xnu, double *out, cudaStream_t st);

// step prologue (c2d_gen.cu)
struct GenParams {
    int nzone, nz, nr, N, nv;
    double rmin;
    const double *z, *r, *gnt, *E_ph;
    const double *f_nt, *n_e, *tea, *B_field;      // [zone][N], [zone] ...
    double *j_sy;                                   // scratch [zone][nv]
    double *kappa_tot, *eps_tot, *eps_th;           // step tables [zone][nv]
    double *sum_th, *sum_g1;                        // [zone] raw thermalised emissivity sum, sum_g_1
};
void launch_volume_em(const GenParams &q, cudaStream_t st);
void launch_pref_calc(int n_ref, double *E_ref, double *P_ref, cudaStream_t st);

// Fokker-Planck (c2d_fp.cu)
struct FpParams;
}  // namespace c2d
