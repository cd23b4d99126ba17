/*
 * c2d_oracle_fp.c -- CPU restatement, part 2 (TEST INFRASTRUCTURE ONLY, see c2d_oracle.h):
 * the per-zone electron Fokker-Planck update FP_calc (update2d.f:337-1741) with its helpers
 * tridag (update2d.f:2485-2528), gamma_bar / McDonald / GammaF / gammln (volume2d.f:572-668) and
 * IC_loss (icloss2d.f:1-125).  Paths relative to /root/reference/src_20121026/.
 *
 * Only the terms that reach the solve are restated: dgdt = dg_sy + dg_ic + dg_A and disp = disp_A
 * (update2d.f:1048-1049).  The Moeller/Coulomb/bremsstrahlung/Landau-damping rates of
 * update2d.f:880-1002 and the te_mo iteration (l.1086-1103) are computed by the reference, written to
 * rates/ files and then discarded; they influence neither f_new, Te_new, n_e, nor the E_tot / hr totals
 *  and are not restated.
 */
#include "c2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.1415926536e0
#define C_LIGHT 2.9979245620e10
#define ZI(j, k) (((j) - 1) * s->nr + ((k) - 1))
#define ZZ(j) (s->z[(j) - 1])
#define RR(k) (s->r[(k) - 1])

/* volume2d.f:649-668 (Numerical Recipes gammln) */
double orc_gammln(double xx)
{
    static const double cof[6] = {76.18009172947146e0, -86.50532032941677e0, 24.01409824083091e0,
                                  -1.231739572450155e0, .1208650973866179e-2, -.5395239384953e-5};
    const double stp = 2.5066282746310005e0;
    double x = xx, y = x, tmp = x + 5.5, ser = 1.000000000190015e0;
    tmp = (x + 0.5) * log(tmp) - tmp;
    for (int j = 0; j < 6; j++) { y = y + 1.0; ser = ser + cof[j] / y; }
    return tmp + log(stp * ser / x);
}

/* volume2d.f:598-626 */
double orc_McDonald(double nu, double z)
{
    double sum = 0.0, t = 1.0, dt = 1.001, d = dt - 1.0, s = 5.e-1 * (1.0 + dt), a = nu - 5.e-1, ts, y, sd;
    do {
        ts = t * s;
        y = z * ts;
        if (y < 2.25e2) sd = pow(ts * ts - 1.0, a) / exp(y); else sd = 0.0;
        sum = sum + d * t * sd;
        t = t * dt;
    } while (t < 2.0 || sd > 1.e-8);
    return sqrt(3.14159265e0) * pow(5.e-1 * z, nu) * sum / exp(orc_gammln(5.e-1 + nu));
}

/* volume2d.f:572-594.  The polynomial's literals are REAL*4. */
double orc_gamma_bar(double Theta)
{
    double g;
    if (Theta < (double)0.2f) {
        g = (1. + (double)4.375f * Theta + (double)7.383f * (Theta * Theta) + (double)3.384f * (Theta * Theta * Theta)) /
                (1. + (double)1.875f * Theta + (double).8203f * (Theta * Theta)) - Theta;
    } else {
        double K2 = orc_McDonald(2.0, 1.0 / Theta), K3 = orc_McDonald(3.0, 1.0 / Theta);
        g = K3 / K2 - Theta;
    }
    if (g < 1.0) g = 1.0;
    return g;
}

/* icloss2d.f:103-125 */
static double f_Li(double z)
{
    double sum, sd = 0.0, nn = 1.0, y = 1.0 + 2.0 * z;
    sum = log(y) * (5.e-1 * log(y) - log(2.0 * z));
    do {
        if (nn < 1.5) sd = 1.0 / y;
        else sd = sd * (((nn - 1.0) / nn) * ((nn - 1.0) / nn)) / y;
        nn = nn + 1.0;
        sum = sum + sd;
    } while (fabs(sd / sum) > 1.e-10);
    return sum;
}
/* icloss2d.f:68-82 */
static double f1(double z)
{
    double y = 1.0 + 2.0 * z;
    double sd1 = (z + 6.0 + 3.0 / z) * log(y);
    double sd2 = ((2.2e1 / 3.0) * (z * z * z) + 24.0 * (z * z) + 18.0 * z + 4.0) / (y * y);
    return sd1 - sd2 - 2.0 + 2.0 * f_Li(z);
}
/* icloss2d.f:85-100 */
static double f2(double z)
{
    double y = 1.0 + 2.0 * z;
    double sd1 = (z + (3.1e1 / 6.0) + 5.0 / z + 1.5 / (z * z)) * log(y);
    double sd2 = ((2.2e1 / 3.0) * (z * z * z) + 28.0 * (z * z) + (1.03e2 / 3.0) * z + 1.7e1 + 3.0 / z) / (y * y);
    return sd1 - sd2 - 2.0 + f_Li(z);
}

/* icloss2d.f:1-50: F_IC[i_el][i_ph] */
void orc_IC_loss(orc_state *s)
{
    const double A = 3.7419e-15;
    for (int ie = 0; ie < s->num_nt; ie++) {
        double gamma = 1.0 + s->gnt[ie];
        double beta = sqrt(1.0 - 1.0 / (gamma * gamma));
        for (int ip = 0; ip < s->nphfield; ip++) {
            double epsilon = 1.957e-3 * s->E_field[ip], v;
            if ((gamma * epsilon) < 1.e-2) v = 2.66e-14 * epsilon * (gamma * gamma - 1.0);
            else {
                double z1 = epsilon * gamma * (1.0 + beta), z2 = epsilon / (gamma * (1.0 + beta));
                double F = gamma * (f1(z1) - f1(z2)) - epsilon * (f2(z1) - f2(z2));
                v = A * F / (((epsilon * gamma) * (epsilon * gamma)) * beta);
            }
            s->F_IC[(size_t)ie * s->nphfield + ip] = v;
        }
    }
}

/* update2d.f:2485-2528 */
void orc_tridag(int n, const double *a, const double *b, const double *c, const double *f_old, double *f_new)
{
    double *gam = (double *)calloc((size_t)n + 1, sizeof(double));
    double bet;
    if (fabs(b[0]) <= 1.e-100) { free(gam); return; }
    bet = b[0];
    f_new[0] = f_old[0] / bet;
    for (int i = 1; i < n; i++) {
        gam[i] = c[i - 1] / bet;
        bet = b[i] - a[i] * gam[i];
        if (fabs(bet) <= 1.e-100) { memset(f_new, 0, sizeof(double) * n); free(gam); return; }
        f_new[i] = (f_old[i] - a[i] * f_new[i - 1]) / bet;
    }
    for (int i = n - 2; i >= 0; i--) {
        f_new[i] = f_new[i] - gam[i + 1] * f_new[i + 1];
        if (f_new[i + 1] < 0.0) f_new[i + 1] = 0.0;
    }
    free(gam);
}

/* update2d.f:337-1741 */
void orc_fp_calc(orc_state *s, int zone)
{
    const int N = s->num_nt, nf = s->nphfield;
    const int j = (zone - 1) / s->nr + 1, k = zone - s->nr * (j - 1), zi = ZI(j, k);
    const double temp_min = 5.0, temp_max = 1.e3;
    const double df_implicit = 1.e-2, df_T = 2.5e-1;      /* general.pa:27-28 */
    const double *gnt = s->gnt;
    double *f_nt = s->f_nt + (size_t)zi * N, *Pnt = s->Pnt + (size_t)zi * N, *n_pos = s->n_pos + (size_t)zi * N;
    const double *dn_pp = s->dn_pp + (size_t)zi * N, *dne_pa = s->dne_pa + (size_t)zi * N, *dnp_pa = s->dnp_pa + (size_t)zi * N;
    const double *n_field = s->n_field + (size_t)zi * nf;
    const double dt1 = s->dt;
    double *gamma = (double *)calloc(12 * (size_t)N + 16, sizeof(double));
    double *f_old = gamma + N, *f_new = f_old + N, *a_i = f_new + N, *b_i = a_i + N, *c_i = b_i + N,
           *dg_ic = c_i + N, *dg_sy = dg_ic + N, *dgdt = dg_sy + N, *disp = dgdt + N, *smw = disp + N, *bigC = smw + N;
    double *bigW = (double *)calloc((size_t)N + 1, sizeof(double)), *npos_new = (double *)calloc((size_t)N + 1, sizeof(double));
    double *inject_ne = (double *)calloc((size_t)N + 1, sizeof(double));
    (void)temp_min; (void)temp_max;

    const double t_esc = s->r_esc * s->z[s->nz - 1] / C_LIGHT, t_acc = s->r_acc * s->z[s->nz - 1] / C_LIGHT;
    double t_fp = 0.0, d_t = 0.0;
    int fp_steps = 0, i;
    s->Te_new[zi] = s->tea[zi];
    const double volume = s->vol[zi];
    double E_el = 0.0, E_pos = 0.0, n_p = s->n_e[zi];
    double ne = n_p * (1. + s->f_pair[zi]), ne_new = ne, n_positron = n_p * s->f_pair[zi], n_lept = ne + n_positron;
    double sum_p = 0.0, sum_E = 0.0;
    if (n_lept < 1.e-11) goto done;
    for (i = 1; i <= N; i++) {
        gamma[i - 1] = gnt[i - 1] + 1.0;
        if (i > 1) {
            double Delta_g = gnt[i - 1] - gnt[i - 2];
            if (s->pair_switch == 1) {
                n_positron = n_positron + Delta_g * n_pos[i - 1];
                E_pos = E_pos + Delta_g * gamma[i - 1] * n_pos[i - 1];
            }
            E_el = E_el + Delta_g * gamma[i - 1] * f_nt[i - 1];
        }
    }
    E_el = E_el * ne * 8.176e-7 * volume;
    if (s->pair_switch == 1) E_pos = E_pos * 8.176e-7 * volume;
    s->E_tot_old = s->E_tot_old + E_el + E_pos + s->ec_old[zi];
    s->E_tot_new = s->E_tot_new + s->ecens[zi];
    sum_p = 0.;
    for (i = 1; i <= N - 1; i++) sum_p = sum_p + (gnt[i] - gnt[i - 1]) * f_nt[i - 1];
    for (i = 1; i <= N; i++) {
        if (s->pair_switch == 1) npos_new[i - 1] = n_pos[i - 1];     /* npos_old; positrons are inert */
        f_nt[i - 1] = f_nt[i - 1] / sum_p;
        f_old[i - 1] = f_nt[i - 1];
    }
    f_old[N - 1] = 0.0;

    double rmid = (k > 1) ? 5.e-1 * (s->r[k - 1] + s->r[k - 2]) : 5.e-1 * (s->r[k - 1] + s->rmin);
    double zmid = (j > 1) ? 5.e-1 * (s->z[j - 1] + s->z[j - 2]) : 5.e-1 * (s->z[j - 1] + s->zmin);
    double tl_flare = 0.0, y;
    if (s->cf_sentinel == 1) {
        y = 5.e-1 * (pow((rmid - s->r_flare) / s->sigma_r, 2.0) + pow((zmid - s->z_flare) / s->sigma_z, 2.0) +
                     pow((s->time - s->t_flare) / s->sigma_t, 2.0));
        tl_flare = (y < 1.e2) ? s->flare_amp / exp(y) : 0.0;
    }
    const double tlev = s->turb_lev[zi] + tl_flare;
    const double Tp_flare = s->tna[zi] * (1.0 + tl_flare);
    const double Th_p = Tp_flare / 9.382e5;
    double Th_e = s->tea[zi] / 5.11e2;
    const double f_th = 1.5 * volume * n_lept;
    for (i = 1; i <= N - 1; i++) {                                    /* l.568-574 */
        double v = 0.0;
        for (int ip = 0; ip < nf; ip++) v = v - n_field[ip] * s->F_IC[(size_t)(i - 1) * nf + ip] / volume;
        dg_ic[i - 1] = v;
    }
    dg_ic[N - 1] = 0.0;          /* never assigned in the FORTRAN (static storage, stays 0) */
    const double B = s->B_field[zi];
    double g_av;
    for (;;) {                                                        /* label 200 */
        g_av = orc_gamma_bar(Th_e);
        double hr_th_c = 0.0;
        for (i = 1; i <= N - 1; i++)
            hr_th_c = hr_th_c - 8.176e-7 * dg_ic[i - 1] * f_old[i - 1] * (gnt[i] - gnt[i - 1]) * volume * n_lept;
        const double gamma_R = 2.1e-3 * sqrt(n_lept) / (B * sqrt(g_av));
        const double h_T = (double).79788f * (2. * pow(Th_e + Th_p, 2.0) + 2.0 * (Th_e + Th_p) + 1.0) /
                           (pow(Th_e + Th_p, 1.5) * (1.0 + 1.875 * Th_e + .8203 * pow(Th_e, 2.0)));
        const double hr_th_Coul = f_th * 1.7386e-26 * n_p * s->lnL * h_T * (Tp_flare - s->Te_new[zi]);
        y = gamma_R / g_av;
        const double hr_th_sy = (y < 100.0) ? -s->Eloss_sy[zi] / (dt1 * exp(y)) : 0.0;
        s->dr = (k == 1) ? s->r[k - 1] - s->rmin : s->r[k - 1] - s->r[k - 2];
        s->dz = (j == 1) ? s->z[j - 1] - s->zmin : s->z[j - 1] - s->z[j - 2];
        double hr_th_A = tlev * hr_th_Coul;
        if (hr_th_A < 1.e-20) hr_th_A = 1.e-20;
        const double hr_th_total = hr_th_sy + hr_th_c + hr_th_A;
        const double dT_total = 6.25e8 * dt1 * hr_th_total / f_th;
        s->f_t_implicit = df_implicit * s->Te_new[zi] / fabs(dT_total);
        if (s->f_t_implicit > df_T) s->f_t_implicit = df_T;
        const double f_sy = 1.058e-15 * pow(B, 2) / 8.176e-7;
        const double g_thr = 1.0 + 4.0 * Th_e;
        double hr_st_A = 0.0, hr_nt_A = 0.0;
        for (i = 1; i <= N; i++) {                                    /* loop 300, live part */
            y = gamma_R / gamma[i - 1];
            if (y < 100.0) dg_sy[i - 1] = -f_sy * (pow(gamma[i - 1], 2.0) - 1.0) / exp(y);
            else dg_sy[i - 1] = -1.e-50;
        }
        for (i = 1; i <= N; i++) {                                    /* loop 350 */
            const double dg_A = gamma[i - 1] / t_acc;
            const double disp_A = gamma[i - 1] * gamma[i - 1] / t_acc / 2.0;
            disp[i - 1] = disp_A;
            dgdt[i - 1] = dg_sy[i - 1] + dg_ic[i - 1] + dg_A;
            if (i < N) {
                hr_nt_A = hr_nt_A + dg_A * f_old[i - 1] * (gamma[i] - gamma[i - 1]);
                if (gamma[i - 1] > g_thr) hr_st_A = hr_st_A + dg_A * f_old[i - 1] * (gamma[i] - gamma[i - 1]);
            }
        }
        hr_st_A = hr_st_A * 8.176e-7 * n_lept * volume;
        hr_nt_A = hr_nt_A * 8.176e-7 * n_lept * volume;
        const double heat_total = hr_th_Coul + hr_nt_A;               /* hr_nt_Coul = hr_th_Coul, l.1032 */
        s->E_tot_old = s->E_tot_old + heat_total * s->f_t_implicit * dt1;      /* label 360 */
        if (fp_steps == 0) { s->hr_total = s->hr_total + heat_total; s->hr_st_total = s->hr_st_total + hr_st_A; }
        d_t = s->f_t_implicit * dt1;
        if (d_t > (dt1 - t_fp)) d_t = 1.00001 * (dt1 - t_fp);
        /* pairs, l.1162-1213 */
        if (s->pair_switch == 0) {
            s->f_pair[zi] = 0.0;
            n_positron = 0.0;
        } else {
            n_positron = s->f_pair[zi] * n_p;
            if (n_positron > 1.e-20) {
                double sp = 0.0;
                for (i = 1; i <= N - 1; i++) sp = sp + npos_new[i - 1] * (gamma[i] - gamma[i - 1]);
                sp = sp / n_positron;
                if (sp > 1.e-20) for (i = 1; i <= N; i++) npos_new[i - 1] = npos_new[i - 1] / sp;
            }
            ne_new = ne;
            n_positron = 0.0;
            for (i = 1; i <= N - 1; i++) {
                const double Delta_ne = (dn_pp[i - 1] + dne_pa[i - 1]) * d_t;
                const double Delta_np = (dn_pp[i - 1] + dnp_pa[i - 1]) * d_t;
                f_old[i - 1] = f_old[i - 1] + Delta_ne / ne;
                npos_new[i - 1] = npos_new[i - 1] + Delta_np;
                if (f_old[i - 1] < 1.e-50) f_old[i - 1] = 0.0;
                if (npos_new[i - 1] < 1.e-50) npos_new[i - 1] = 0.0;
                n_positron = n_positron + npos_new[i - 1] * (gamma[i] - gamma[i - 1]);
            }
            if (n_positron < 1.e-50) n_positron = 0.0;
        }
        ne_new = n_p + n_positron;                                    /* label 465 */
        ne = ne_new;
        s->f_pair[zi] = fmax(0.0, (n_positron / n_p));
        /* injection, l.1215-1306 */
        double n_inject = 0.0, inj_sum, inj_E;
        if (s->pick_sw == 1) {
            inj_sum = 0.0; inj_E = 0.0;
            for (i = 1; i <= N - 1; i++) {
                inject_ne[i - 1] = 1.e2 * exp(-pow(gamma[i - 1] - s->inj_gg, 2) / 2 / pow(s->inj_sigma, 2)) /
                                   (s->inj_sigma * sqrt(2 * PI));
                inj_sum = inj_sum + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]);
                inj_E = inj_E + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]) * gamma[i - 1];
            }
            const double inj_rho = s->pick_rate * d_t;
            for (i = 1; i <= N - 1; i++) {
                inject_ne[i - 1] = inj_rho * inject_ne[i - 1] / inj_sum;
                f_old[i - 1] = f_old[i - 1] + inject_ne[i - 1] / ne;
                n_inject = n_inject + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]);
            }
        }
        if (s->inj_switch != 0) {
            inj_sum = 0.0; inj_E = 0.0;
            const double tt = s->time + t_fp - s->inj_t;
            if (tt > s->dz / s->inj_v * (j - 1) && tt < s->dz / s->inj_v * j && k <= s->nr) {
                for (i = 1; i <= N - 1; i++) {
                    if (s->inj_dis == 1) {
                        inject_ne[i - 1] = 1.e2 * exp(-pow(gamma[i - 1] - s->inj_gg, 2) / 2 / pow(s->inj_sigma, 2)) /
                                           (s->inj_sigma * sqrt(2 * PI));
                    } else if (s->inj_dis == 2) {
                        const double inj_g2var = s->inj_g2 * pow(10.0, tt * s->inj_v / s->z[s->nz - 1]);
                        if (gamma[i - 1] > s->inj_g1) {
                            const double inj_y = (s->g2var_switch == 1) ? gamma[i - 1] / inj_g2var : gamma[i - 1] / s->inj_g2;
                            if (inj_y < 1.e2) inject_ne[i - 1] = 1.e2 / (pow(gamma[i - 1], s->inj_p) * exp(inj_y));
                            else inject_ne[i - 1] = 0.0;
                        } else inject_ne[i - 1] = 0.0;
                    }
                    inj_sum = inj_sum + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]);
                    inj_E = inj_E + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]) * gamma[i - 1];
                }
                inj_E = inj_E / inj_sum;
                const double rn = s->r[s->nr - 1];
                const double inj_rate = s->inj_L / 8.186e-7 / inj_E / (PI * (rn * rn) * s->dz);
                const double inj_rho = inj_rate * d_t;
                for (i = 1; i <= N - 1; i++) {
                    inject_ne[i - 1] = inj_rho * inject_ne[i - 1] / inj_sum;
                    f_old[i - 1] = f_old[i - 1] + inject_ne[i - 1] / ne;
                    n_inject = n_inject + inject_ne[i - 1] * (gnt[i] - gnt[i - 1]);
                }
            }
        }
        ne_new = ne + n_inject;
        ne = ne_new;
        n_p = n_p + n_inject;
        s->n_e[zi] = n_p;
        n_lept = n_lept + n_inject;
        /* escape, l.1309-1313 */
        ne_new = ne_new * t_esc / (t_esc + d_t);
        ne = ne_new;
        n_p = n_p * t_esc / (t_esc + d_t);
        s->n_e[zi] = n_p;
        n_lept = n_lept * t_esc / (t_esc + d_t);
        /* Chang-Cooper coefficients, l.1319-1391 */
        a_i[0] = 0.0; b_i[0] = 1.0; c_i[0] = 0.0;
        a_i[N - 1] = 0.0; b_i[N - 1] = 1.0; c_i[N - 1] = 0.0;
        for (i = 2; i <= N - 1; i++) {
            const double D_gminus = gnt[i - 1] - gnt[i - 2], D_gplus = gnt[i] - gnt[i - 1];
            const double Delta_g = sqrt(gnt[i - 1] / gnt[i - 2]) * D_gminus;
            double bigB;
            if (i == 2) {
                bigB = -(dgdt[0] + dgdt[1]);
                bigC[0] = (disp[0] + disp[1]) / 2.0;
                smw[0] = D_gminus * bigB / bigC[0];
                bigW[0] = smw[0] / (exp(smw[0]) - 1.0);
            }
            bigB = -(dgdt[i - 1] + dgdt[i]) / 2.0;
            bigC[i - 1] = (disp[i - 1] + disp[i]) / 2.0;
            smw[i - 1] = D_gplus * bigB / bigC[i - 1];
            bigW[i - 1] = smw[i - 1] / (exp(smw[i - 1]) - 1.0);
            c_i[i - 1] = -d_t * (bigC[i - 1] * smw[i - 1] / (1.0 - exp(-smw[i - 1])) / Delta_g / D_gplus);
            b_i[i - 1] = 1.0 + d_t / Delta_g * (bigC[i - 1] * bigW[i - 1] / D_gplus +
                                                 bigC[i - 2] * smw[i - 2] / (1.0 - exp(-smw[i - 2])) / D_gminus) + d_t / t_esc;
            a_i[i - 1] = -d_t / Delta_g * bigC[i - 2] * bigW[i - 2] / D_gminus;
        }
        orc_tridag(N, a_i, b_i, c_i, f_old, f_new);
        f_new[N - 1] = 0.0;
        f_new[0] = 0.0;
        sum_p = 0.; sum_E = 0.;
        for (i = 1; i <= N - 1; i++) {
            sum_p = sum_p + (gnt[i] - gnt[i - 1]) * f_new[i - 1];
            sum_E = sum_E + (gnt[i] - gnt[i - 1]) * gamma[i - 1] * f_new[i - 1];
            Pnt[i - 1] = sum_p;
        }
        sum_E = sum_E / sum_p;
        t_fp = t_fp + d_t;
        fp_steps = fp_steps + 1;
        for (i = 1; i <= N; i++) { f_new[i - 1] = f_new[i - 1] / sum_p; f_old[i - 1] = f_new[i - 1]; }
        /* new electron temperature, l.1441-1469 */
        double gbar = 0.0;
        for (i = 1; i <= N - 1; i++) gbar = gbar + gamma[i - 1] * f_new[i - 1] * (gnt[i] - gnt[i - 1]);
        double The_new = Th_e;
        if (gbar > g_av) {
            while (gbar > g_av) { The_new = The_new * (double)1.005f; g_av = orc_gamma_bar(The_new); }   /* REAL*4 literal */
        } else {
            while (gbar < g_av) {
                The_new = The_new / (double)1.005f;
                g_av = orc_gamma_bar(The_new);
                if (The_new < 1.e-2) break;
            }
        }
        s->Te_new[zi] = 5.11e2 * The_new;
        Th_e = The_new;
        if (!(t_fp < dt1)) break;
        if (fp_steps > 1000000) { s->fatal = 4; break; }
    }
    s->fp_steps[zi] = fp_steps;
    /* epilogue, l.1480-1741 */
    E_el = 0.0; E_pos = 0.0;
    for (i = 1; i <= N; i++) {
        f_nt[i - 1] = f_new[i - 1];
        if (s->pair_switch == 1) n_pos[i - 1] = npos_new[i - 1];
        Pnt[i - 1] = Pnt[i - 1] / sum_p;
        if (i > 1) {
            E_el = E_el + f_nt[i - 1] * gamma[i - 1] * (gnt[i - 1] - gnt[i - 2]);
            if (s->pair_switch == 1) E_pos = E_pos + n_pos[i - 1] * gamma[i - 1] * (gnt[i - 1] - gnt[i - 2]);
        }
    }
    E_el = E_el * ne * 8.176e-7 * volume;
    if (s->pair_switch == 1) E_pos = E_pos * 8.176e-7 * volume;
    s->E_tot_new = s->E_tot_new + E_el + E_pos;
    {
        const double Delta_T = fabs(s->Te_new[zi] - s->tea[zi]) / s->Te_new[zi];
        if (Delta_T > s->dT_max) s->dT_max = Delta_T;
    }
    /* power-law re-fit, l.1656-1738 */
    {
        int i_nt;
        for (i = 5; i <= N - 5; i++) if (f_new[i - 1] > 1.e-10) break;
        s->gmin[zi] = gamma[i - 1];
        i_nt = i;
        for (i = N - 5; i >= 5; i--) if (f_new[i - 1] > 1.e-15) break;
        s->gmax[zi] = gamma[i - 1];
        double sum_nt = 0.0, sum_th = 0.0;
        for (i = 1; i <= N - 1; i++) {
            if (i < i_nt) sum_th = sum_th + (gamma[i] - gamma[i - 1]) * f_new[i - 1];
            else sum_nt = sum_nt + (gamma[i] - gamma[i - 1]) * f_new[i - 1];
        }
        s->amxwl[zi] = sum_th / (sum_nt + sum_th);
        if (s->amxwl[zi] > 9.999e-1) {
            s->amxwl[zi] = 1.0;
        } else {
            double sum_g = 1.e50, sumg_old;
            s->p_nth[zi] = (double)0.1f;
            for (;;) {
                sumg_old = sum_g;
                sum_g = 0.0;
                double sum_gg = 0.0, N_nt;
                const double p_1 = 1.0 - s->p_nth[zi];
                if (fabs(p_1) > 1.e-4) N_nt = (1. - s->amxwl[zi]) * p_1 / (pow(s->gmax[zi], p_1) - pow(s->gmin[zi], p_1));
                else N_nt = (1.0 - s->amxwl[zi]) / log(s->gmax[zi] / s->gmin[zi]);
                for (i = i_nt; i <= N - 2; i++) {
                    y = gamma[i - 1] / s->gmax[zi];
                    if (y < 100.0) {
                        const double f_pl = N_nt / (pow(gamma[i - 1], s->p_nth[zi]) * exp(y));
                        sum_g = sum_g + f_pl * gamma[i - 1] * (gnt[i] - gnt[i - 1]);
                        sum_gg = sum_gg + f_pl * (gnt[i] - gnt[i - 1]);
                    } else break;
                }
                sum_g = sum_g / sum_gg;
                sum_g = fabs(sum_g - sum_E);
                if (sum_g < sumg_old) {
                    if (s->p_nth[zi] < 10.) s->p_nth[zi] = s->p_nth[zi] + 0.5e-1;
                    else break;
                    continue;
                }
                break;
            }
        }
    }
done:
    free(gamma); free(bigW); free(npos_new); free(inject_ne);
}


/* ------------------------------------------------------------------ step prologue */
/* volume2d.f:676-713: exp(t) K_{1/3}(t) */
static double expk13(double t)
{
    const double c1 = (double)0.35502805f, c2 = (double)0.25881940f;   /* DATA with REAL*4 literals (l.686) */
    if (t <= 1.0) {
        double z3 = 1.5 * t;
        const double zs = pow(z3, 0.3333333333333333e0);
        const double z = zs * zs;
        z3 = z3 * z3;
        const double f1 = 1.0 + z3 / 6.0 * (1.0 + z3 / 30.0 * (1.0 + z3 / 56.0));
        const double f2 = z * (1.0 + z3 / 12.0 * (1.0 + z3 / 42.0 * (1.0 + z3 / 90.0)));
        return exp(t) * PI * 1.7320508e0 / zs * (c1 * f1 - c2 * f2);
    }
    const double z = 1.0 / (72.0 * t);
    const double poly = 1.0 - 5.0 * z * (1.0 - 38.5 * z);
    return sqrt(0.5 * PI / t) * poly / (1.0 + 1.0 / (1.0 + 58.0 * t * t));
}
/* volume2d.f:717-746: exp(t) K_{4/3}(t) */
static double expk43(double t)
{
    if (t <= 1.0) {
        const double poly = 1.0 + t * (0.9757317e0 - 7.6790616e-2 * t);
        return 0.44648975e0 * pow(2.0 / t, 1.333333333e0) * poly;
    }
    const double z = 1.0 / (72.0 * t);
    const double poly = 1.0 + 55.0 * z * (1.0 - 8.5 * z);
    return sqrt(0.5 * PI / t) * poly * (1.0 + 1.0 / (1.0 + 50.0 * t * t));
}

/* volume2d.f:10-394 volume_em(j,k,T_keV,ne_local,B,l_min): only the terms that reach an output.  Since the 2009-2011
 * edits (l.330-341) kappa_tot = kappa_sy and the emission CDF holds synchrotron only; bremsstrahlung, cyclotron and
 * pair annihilation are computed by the reference and discarded (Eloss_cy only goes to the log).  Outputs: E_ph,
 * kappa_tot, eps_tot, eps_th of the zone and the raw thermalised emissivity sum in Eloss_th (imcgen2d scales it). */
void orc_volume_em(orc_state *s, int zone, double l_min)
{
    const int N = s->num_nt, nv = s->n_vol;
    const int j = (zone - 1) / s->nr + 1, k = zone - s->nr * (j - 1), zi = ZI(j, k);
    const double T_keV = s->tea[zi], ne_local = s->n_e[zi], B = s->B_field[zi];
    const double *f = s->f_nt + (size_t)zi * N, *gnt = s->gnt;
    const double ee = 4.803e-10, em = 9.109e-28, sigmaT = 6.6524616e-25;
    const double nu_b = ee * B / (2 * PI * em * C_LIGHT);
    const double Ub = pow(B, 2.0) / (8.0 * PI);
    const double face = pow(3.0, 1.5) * sigmaT * C_LIGHT * Ub / (PI * nu_b);
    const double dE = exp(log(1.e20) / (double)nv);
    const double kappa_C = 6.65e-25 * ne_local;
    const double nu_p = 9.e3 * sqrt(ne_local);
    double *gamma0 = (double *)malloc(sizeof(double) * 2 * (size_t)N), *gamp = gamma0 + N;
    double *P = (double *)malloc(sizeof(double) * 2 * (size_t)nv), *P_th = P + nv;
    double P_sum = 0., sum_th = 0., E = 1.e-10 / dE;
    int i, i2;
    for (i = 0; i < N; i++) { gamma0[i] = gnt[i] + 1.0; gamp[i] = gamma0[i] * sqrt(pow(gamma0[i], 2.0) - 1.0); }
    s->Eloss_th[zi] = 0.;
    for (i = 1; i <= nv; i++) {
        double j_sy, kappa_sy, nu;
        E = E * dE;
        s->E_ph[i - 1] = E;
        nu = 2.41487e17 * E;
        if (nu <= nu_p) { j_sy = 0.; kappa_sy = 0.; }
        else {
            double sum = 0., sum_k = 0.;
            for (i2 = 1; i2 <= N - 1; i2++) {
                const double facg = 3.0 * pow(gamma0[i2 - 1], 2.0) * nu_b;
                const double tt = nu / facg;
                double es;
                if (tt < 1.0e4) {
                    const double eq43 = expk43(tt), eq13 = expk13(tt);
                    const double ff = tt * tt * (eq43 * eq13 - (double)0.6f * tt * (eq43 - eq13) * (eq43 + eq13));
                    es = face * ff * exp(-2.0 * tt);
                } else es = 0.0;
                const double sd = f[i2 - 1] * es, sd_k = gamp[i2 - 1] * es;
                sum = sum + (gnt[i2] - gnt[i2 - 1]) * sd;
                sum_k = sum_k + (f[i2 - 1] / gamp[i2 - 1] - f[i2] / gamp[i2]) * sd_k;
            }
            j_sy = sum * ne_local / (4.0 * PI);
            kappa_sy = sum_k * ne_local / (8.0 * PI * em * (nu * nu));
        }
        if (kappa_sy < 0.) kappa_sy = -1.0 * kappa_sy;
        s->kappa_tot[(size_t)zi * nv + (i - 1)] = kappa_sy;
        if (kappa_sy < fmax(1.0 / l_min, 1e1 * kappa_C)) {
            P_sum = P_sum + j_sy * E * (dE - 1.0);
        } else {
            const double x = E / T_keV, tau_tot = kappa_sy * l_min;
            double j_th = (x < 1.e2) ? 1.47e-47 * (nu * nu * nu) / (exp(x) - 1.0) : 1.e-50;
            if (tau_tot < 5.e1) j_th = j_th * (1.0 - exp(-tau_tot));
            sum_th = sum_th + j_th * E * (dE - 1.0);
            s->Eloss_th[zi] = s->Eloss_th[zi] + j_th * E * (dE - 1.0);
        }
        P[i - 1] = P_sum;
        P_th[i - 1] = sum_th;
    }
    for (i = 0; i < nv; i++) {
        s->eps_tot[(size_t)zi * nv + i] = (P_sum > 1.e-50) ? P[i] / P_sum : 0.;
        s->eps_th[(size_t)zi * nv + i] = (sum_th > 1.e-50) ? P_th[i] / sum_th : 0.;
    }
    free(gamma0); free(P);
}

/* imcgen2d.f:203-336 (per-zone energy budget; the emission is synchrotron only, l.330) and l.405-459 (photon budget).
 * ep_switch = 0 (B_field as given).  Fills kappa_tot, eps_*, E_ph, Eloss_sy, Eloss_th, Eloss_tot, nsv, ewsv. */
void orc_imcgen_volume(orc_state *s, int nst)
{
    const int N = s->num_nt;
    double Emiss_tot = 0.;
    for (int j = 1; j <= s->nz; j++)
        for (int k = 1; k <= s->nr; k++) {
            const int zi = ZI(j, k), zone = k + s->nr * (j - 1);
            const double dzz = (j == 1) ? ZZ(1) : ZZ(j) - ZZ(j - 1), drr = (k == 1) ? RR(1) - s->rmin : RR(k) - RR(k - 1);
            const double l_min = fmin(dzz, drr);
            const double *f = s->f_nt + (size_t)zi * N;
            orc_volume_em(s, zone, l_min);
            double sum_g_1 = 0.;
            for (int i = 1; i <= N - 1; i++)
                sum_g_1 = sum_g_1 + (pow(s->gnt[i - 1] + 1.0, 2.0) - 1.0) * f[i - 1] * (s->gnt[i] - s->gnt[i - 1]);
            s->Eloss_sy[zi] = 1.058e-15 * s->n_e[zi] * s->dt * pow(s->B_field[zi], 2.0) * sum_g_1 * s->vol[zi];
            s->Eloss_th[zi] = s->dt * s->zsurf[zi] * s->Eloss_th[zi];
            s->Eloss_tot[zi] = s->Eloss_sy[zi];
            Emiss_tot = Emiss_tot + s->Eloss_tot[zi];
        }
    for (int zi = 0; zi < s->nzone; zi++) {
        s->nsv[zi] = (int)(0.5 * nst * s->Eloss_tot[zi] / Emiss_tot);
        s->ewsv[zi] = (s->nsv[zi] > 0) ? s->Eloss_tot[zi] / (double)s->nsv[zi] : 0.;
    }
}
