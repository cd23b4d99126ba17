// c2d_gen.cu -- the step prologue on the device: volume_em (volume2d.f:10-394) for every zone, from the
// live electron spectrum, written straight into the step tables the tracker reads.
//
// Only the terms that reach an output are evaluated.  Since the reference's 2009-2011 edits kappa_tot = kappa_sy
// (volume2d.f:330-331) and the emission CDF eps_tot holds synchrotron only (l.338-340; imcgen2d.f:330 Eloss_tot =
// Eloss_sy); bremsstrahlung, cyclotron and pair-annihilation emissivities are computed by the reference and discarded.
//   k_volume_sy   one thread per (zone, photon energy E_ph(i)): j_sy and kappa_sy as sums over the electron bins of the
//                 exact single-electron synchrotron emissivity (expk43 / expk13 fits of exp(t) K_{4/3}, K_{1/3}), l.206-239
//   k_volume_cdf  one thread per zone: the running sums P, P_th over the energy grid with the optically-thick switch
//                 (l.335-357), their normalisation into eps_tot / eps_th (l.364-375), and sum_g_1 of imcgen2d.f:280-283
// The photon budget (Emiss_tot, nsv, ewsv: imcgen2d.f:405-459) is a sequential sum over a few hundred zones and is
// finished on the host by c2d_step_begin_gen.
#include "c2d_ctx.h"

namespace c2d {

// volume2d.f:676-713: exp(t) K_{1/3}(t)
__device__ __forceinline__ double expk13(double t)
{
    const double c1 = (double)0.35502805f, c2 = (double)0.25881940f;   // DATA with REAL*4 literals (l.686)
    if (t <= 1.0) {
        double z3 = 1.5 * t;
        const double zs = pow(z3, 0.3333333333333333e0);
        const double z = zs * zs;
        z3 = z3 * z3;
        const double f1 = 1.0 + z3 / 6.0 * (1.0 + z3 / 30.0 * (1.0 + z3 / 56.0));
        const double f2 = z * (1.0 + z3 / 12.0 * (1.0 + z3 / 42.0 * (1.0 + z3 / 90.0)));
        return exp(t) * kPi * 1.7320508e0 / zs * (c1 * f1 - c2 * f2);
    }
    const double z = 1.0 / (72.0 * t);
    const double poly = 1.0 - 5.0 * z * (1.0 - 38.5 * z);
    return sqrt(0.5 * kPi / t) * poly / (1.0 + 1.0 / (1.0 + 58.0 * t * t));
}
// volume2d.f:717-746: exp(t) K_{4/3}(t)
__device__ __forceinline__ double expk43(double t)
{
    if (t <= 1.0) {
        const double poly = 1.0 + t * (0.9757317e0 - 7.6790616e-2 * t);
        return 0.44648975e0 * pow(2.0 / t, 1.333333333e0) * poly;
    }
    const double z = 1.0 / (72.0 * t);
    const double poly = 1.0 + 55.0 * z * (1.0 - 8.5 * z);
    return sqrt(0.5 * kPi / t) * poly * (1.0 + 1.0 / (1.0 + 50.0 * t * t));
}

__global__ void __launch_bounds__(128) k_volume_sy(const __grid_constant__ GenParams q)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)q.nzone * q.nv) return;
    const int zi = (int)(idx / q.nv), i = (int)(idx % q.nv);
    const double ee = 4.803e-10, em = 9.109e-28, sigmaT = 6.6524616e-25;
    const double B = q.B_field[zi], ne_local = q.n_e[zi];
    const double nu_b = ee * B / (2 * kPi * em * kCLight);
    const double Ub = sq(B) / (8.0 * kPi);
    const double face = pow(3.0, 1.5) * sigmaT * kCLight * Ub / (kPi * nu_b);
    const double nu_p = 9.e3 * sqrt(ne_local);
    const double nu = 2.41487e17 * q.E_ph[i];
    double j_sy = 0., kappa_sy = 0.;
    if (!(nu <= nu_p)) {
        const double *f = q.f_nt + (size_t)zi * q.N;
        double sum = 0., sum_k = 0.;
        double g_cur = q.gnt[0] + 1.0;
        double gamp_cur = g_cur * sqrt(sq(g_cur) - 1.0);
        for (int i2 = 0; i2 < q.N - 1; i2++) {
            const double g_next = q.gnt[i2 + 1] + 1.0;
            const double gamp_next = g_next * sqrt(sq(g_next) - 1.0);
            const double facg = 3.0 * sq(g_cur) * nu_b;
            const double tt = nu / facg;
            double es = 0.0;
            if (tt < 1.0e4) {
                const double eq43 = expk43(tt), eq13 = expk13(tt);
                const double ff = tt * tt * (eq43 * eq13 - (double)0.6f * tt * (eq43 - eq13) * (eq43 + eq13));
                es = face * ff * exp(-2.0 * tt);
            }
            const double sd = f[i2] * es, sd_k = gamp_cur * es;
            sum = sum + (q.gnt[i2 + 1] - q.gnt[i2]) * sd;
            sum_k = sum_k + (f[i2] / gamp_cur - f[i2 + 1] / gamp_next) * sd_k;
            g_cur = g_next; gamp_cur = gamp_next;
        }
        j_sy = sum * ne_local / (4.0 * kPi);
        kappa_sy = sum_k * ne_local / (8.0 * kPi * em * (nu * nu));
    }
    if (kappa_sy < 0.) kappa_sy = -1.0 * kappa_sy;
    q.j_sy[idx] = j_sy;
    q.kappa_tot[idx] = kappa_sy;
}

__global__ void __launch_bounds__(64) k_volume_cdf(const __grid_constant__ GenParams q)
{
    const int zi = blockIdx.x * blockDim.x + threadIdx.x;
    if (zi >= q.nzone) return;
    const int j = zi / q.nr + 1, k = zi - q.nr * (j - 1) + 1;
    const double dzz = (j == 1) ? q.z[0] : q.z[j - 1] - q.z[j - 2], drr = (k == 1) ? q.r[0] - q.rmin : q.r[k - 1] - q.r[k - 2];
    const double l_min = fmin(dzz, drr);                          // imcgen2d.f:238-246
    const double T_keV = q.tea[zi], ne_local = q.n_e[zi];
    const double kappa_C = 6.65e-25 * ne_local;
    const double dE = exp(log(1.e20) / (double)q.nv);
    const double thick = fmax(1.0 / l_min, 1e1 * kappa_C);
    double *et = q.eps_tot + (size_t)zi * q.nv, *eh = q.eps_th + (size_t)zi * q.nv;
    const double *kap = q.kappa_tot + (size_t)zi * q.nv, *js = q.j_sy + (size_t)zi * q.nv;
    double P_sum = 0., sum_th = 0.;
    for (int i = 0; i < q.nv; i++) {
        const double E = q.E_ph[i], nu = 2.41487e17 * E;
        if (kap[i] < thick) P_sum = P_sum + js[i] * E * (dE - 1.0);
        else {
            const double x = E / T_keV, tau_tot = kap[i] * l_min;
            double j_th = (x < 1.e2) ? 1.47e-47 * (nu * nu * nu) / (exp(x) - 1.0) : 1.e-50;
            if (tau_tot < 5.e1) j_th = j_th * (1.0 - exp(-tau_tot));
            sum_th = sum_th + j_th * E * (dE - 1.0);
        }
        et[i] = P_sum; eh[i] = sum_th;
    }
    for (int i = 0; i < q.nv; i++) {
        et[i] = (P_sum > 1.e-50) ? et[i] / P_sum : 0.;
        eh[i] = (sum_th > 1.e-50) ? eh[i] / sum_th : 0.;
    }
    const double *f = q.f_nt + (size_t)zi * q.N;
    double sum_g_1 = 0.;
    for (int i = 0; i < q.N - 1; i++) sum_g_1 = sum_g_1 + (sq(q.gnt[i] + 1.0) - 1.0) * f[i] * (q.gnt[i + 1] - q.gnt[i]);
    q.sum_th[zi] = sum_th;
    q.sum_g1[zi] = sum_g_1;
}

// == Pref_calc (ref_matrix.f:7-89): the Green's function of Compton reflection off cold matter (White, Lightman & Zdziarski
// 1988) on the grid E_ref(n) = 10^(3 (n-1) / n_ref) keV, as the cumulative distribution over the outgoing energy that
// imcleak samples (imcleak2d.f:126-146): P_ref(n_out, n_in), column n_in contiguous.  Below E_trans = 20 keV a photon keeps
// its energy (a step at n_out = n_in).  One thread per incoming energy: the column is a running sum.  The REAL*4 literals
// of the source (1.12, the exponents 0.785 1.04 .51 .824 .85) are kept as the single-precision values Fortran promotes.
__global__ void k_pref_calc(int n_ref, double *E_ref, double *P_ref)
{
    const int n_in = (int)(blockIdx.x * blockDim.x + threadIdx.x);           // 0-based
    if (n_in >= n_ref) return;
    const double dE = exp(log(1.e3) / (double)n_ref), e_trans = 20.0;
    double e_in = 1.0;
    for (int n = 1; n <= n_in; n++) e_in = e_in * dE;                         // E_ref(n) = E_ref(n-1) dE, as the reference rounds it
    E_ref[n_in] = e_in;                                                       // (every thread writes its own entry)
    double *col = P_ref + (size_t)n_in * n_ref;
    for (int n_out = n_in + 1; n_out < n_ref; n_out++) col[n_out] = 1.0;
    if (e_in <= e_trans) {
        for (int n_out = 0; n_out < n_in; n_out++) col[n_out] = 0.0;
        col[n_in] = 1.0;
        return;
    }
    const double x0 = 1.957e-3 * e_in, y0 = 1.0 / x0, dyc = 1.e3 - y0;
    const double A = 5.6e-1 + (double)1.12f / pow(y0, (double)0.785f) - 3.4e-1 / pow(y0, (double)1.04f);
    const double al = -3.e-1 / pow(y0, (double).51f) + 6.e-2 / pow(y0, (double).824f);
    const double beta = 3.7e-1 - pow(y0, (double).85f);
    const double tail = (fabs(al + .5) < 1.e-4) ? (2.0 + log(.5 * dyc)) : (2.0 + (pow(.5 * dyc, al + .5) - 1.0) / (al + .5));
    const double B = (1.0 - A * tail / sqrt(dyc))
                     / (pow(y0, 1.0 - beta) * pow(y0 + 2.0, beta) * (pow(1.0 + 2.0 / y0, 1.0 - beta) - 1.0)) * (1.0 - beta);
    double sum = 0.0, e_out = 1.0;
    for (int n_out = 0; n_out <= n_in; n_out++) {
        if (n_out > 0) e_out = e_out * dE;
        const double x1 = 1.957e-3 * e_out, dy = 1.0 / x1 - y0;
        double Gy;
        if (dy < 2.0) Gy = B * pow((y0 + 2.0) / (y0 + dy), beta);
        else if (dy < dyc) Gy = A * pow(dyc / dy, al) / pow(dy, 1.5);
        else Gy = A / pow(dy, 1.5);
        sum = sum + (Gy / (x1 * x1)) * (dE * x1);
        col[n_out] = sum;
    }
    for (int n_out = 0; n_out <= n_in; n_out++) col[n_out] = col[n_out] / sum;
}

#ifndef C2D_EMU
void launch_pref_calc(int n_ref, double *E_ref, double *P_ref, cudaStream_t st)
{
    k_pref_calc<<<(n_ref + 63) / 64, 64, 0, st>>>(n_ref, E_ref, P_ref);
}
void launch_volume_em(const GenParams &q, cudaStream_t st)
{
    const long long n = (long long)q.nzone * q.nv;
    k_volume_sy<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(q);
    k_volume_cdf<<<(q.nzone + 63) / 64, 64, 0, st>>>(q);
}
#endif

}  // namespace c2d
