/*
 * c2d_oracle.h -- CPU restatement of Compton2d's per-time-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement of
 * the reference's FORTRAN (every function cites the reference file:line it
 * follows, paths relative to /root/reference/src_20121026/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (compton2d_b200/) never links, imports or calls it.
 *
 * Parity pin: checked against COMMON-block snapshots of the UNMODIFIED reference
 * binary (tests/golden/refbin_*.npz, made by oracle/refbin/make_golden.py) in
 * rand_switch=2 (ran1) mode: integer tallies, census integers and the RNG end
 * state bit-exact, FP64 tallies to ~1e-12 (the binary uses Intel's libm).
 * See tests/test_oracle_vs_refbin.py.
 *
 * Two RNG modes:
 *   rand_switch = 2  Numerical-Recipes ran1, one serial stream (rand.f:322-359);
 *                    reproduces the reference binary with one worker.
 *   rand_switch = 3  (extension) counter-based Philox-4x32-10 streams keyed by
 *                    history ancestry, so that results do not depend on execution
 *                    order -- the replay mode the CUDA path is compared in.
 *
 * Array layout (dense, 0-based C): zone index zi = (j-1)*nr + (k-1) with j the
 * z index and k the r index (fp_mpi.f:333-352); per-zone tables are [zi][bin].
 */
#ifndef C2D_ORACLE_H
#define C2D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int mode;                 /* 1 = Knuth lagged Fibonacci (rand_switch = 1), 2 = ran1 (rand_switch = 2), 3 = philox */
    /* Knuth ranf_array state: COMMON /RSTATE/ RANX(100), /rlist/ randlist(randmax = 10000), /rcount/ randcounter (rand.f:230-316) */
    double ranx[100], randlist[10000];
    int randcounter;
    /* ran1 state (rand.f:330-343) */
    int idum, iv[32], iy;
    /* philox */
    uint32_t key[2];
    uint64_t sid;             /* current stream id */
    uint32_t d;               /* next draw index in the stream */
    int has_pending;          /* first draw of a trial stream is supplied by the parent */
    double pending;
    uint64_t ndraw;           /* statistics */
} orc_rng;

#define ORC_NSPEC 4        /* external-spectrum tables held at once (file_sp results, imcsurf2d_para.f:544-685) */
#define ORC_NFMAX 500      /* general.pa:17 nfmax */
#define ORC_NMUMAX 32
#define ORC_NLCMAX 16

/* X-macro: name, C type, element-count expression (in terms of dims) */
#define ORC_NREF 500        /* general.pa:19 n_ref */
#define ORC_FIELDS(X) \
    X(z, double, nz) X(r, double, nr) \
    X(E_ph, double, n_vol) X(kappa_tot, double, nzone * n_vol) \
    X(eps_tot, double, nzone * n_vol) X(eps_th, double, nzone * n_vol) \
    X(E_field, double, nphfield) X(E_gg, double, n_gg) X(k_gg, double, nzone * n_gg) \
    X(f_pair, double, nzone) X(gnt, double, num_nt) X(f_nt, double, nzone * num_nt) \
    X(Pnt, double, nzone * num_nt) X(n_e, double, nzone) X(tea, double, nzone) \
    X(hu, double, nphtotal + 1) X(Elcmin, double, ORC_NLCMAX) X(Elcmax, double, ORC_NLCMAX) \
    X(mu, double, ORC_NMUMAX) \
    X(tbbl, double, nr) X(tbbu, double, nr) X(tbbi, double, nz) X(tbbo, double, nz) \
    X(nsurfi, int, nz) X(nsurfo, int, nz) X(nsurfu, int, nr) X(nsurfl, int, nr) \
    X(ewsurfi, double, nz) X(ewsurfo, double, nz) X(ewsurfu, double, nr) X(ewsurfl, double, nr) \
    X(spec_i, int, nz) X(spec_o, int, nz) X(spec_u, int, nr) X(spec_l, int, nr) \
    X(nfile, int, ORC_NSPEC) X(E_file, double, ORC_NSPEC * ORC_NFMAX) X(F_file, double, ORC_NSPEC * ORC_NFMAX) \
    X(P_file, double, ORC_NSPEC * ORC_NFMAX) X(a1_file, double, ORC_NSPEC * ORC_NFMAX) \
    X(I_file, double, ORC_NSPEC * ORC_NFMAX) \
    /* COMMON /random/ seeds(j,k), zseeds(j), rseeds(k) of the lagged-Fibonacci generator (commonblock.f:9-10,130) */ \
    X(seeds, int, nzone) X(zseeds, int, nz) X(rseeds, int, nr) \
    X(nsv, int, nzone) X(ewsv, double, nzone) X(Eloss_th, double, nzone) X(Eloss_tot, double, nzone) \
    X(zsurf, double, nzone) \
    X(edep, double, nzone) X(prdep, double, nzone) X(ecens, double, nzone) X(npcen, int64_t, nzone) \
    X(n_field, double, nzone * nphfield) X(n_ph, double, nzone * n_gg) X(E_IC, double, num_nt) \
    X(nelectron, int64_t, num_nt + 2) X(fout, double, ORC_NMUMAX * nphtotal) \
    X(edout, double, ORC_NMUMAX * ORC_NLCMAX) \
    X(erlki, double, nz) X(erlko, double, nz) X(erlku, double, nr) X(erlkl, double, nr) \
    X(Ed_in, double, nr) X(Ed_ref, double, nr) \
    /* Compton-reflection tables of ref_matrix.f (inputs: general.pa n_ref = 500); P_ref / W_abs in COMMON order, \
       element (n_out, n_in) at (n_in - 1) n_ref + (n_out - 1) */ \
    X(E_ref, double, ORC_NREF) X(P_ref, double, ORC_NREF * ORC_NREF) X(W_abs, double, ORC_NREF * ORC_NREF) \
    /* FP inputs / outputs */ \
    X(tna, double, nzone) X(B_field, double, nzone) X(vol, double, nzone) \
    X(Eloss_sy, double, nzone) X(Eloss_br, double, nzone) X(ec_old, double, nzone) \
    X(q_turb, double, nzone) X(turb_lev, double, nzone) X(amxwl, double, nzone) \
    X(gmin, double, nzone) X(gmax, double, nzone) X(p_nth, double, nzone) X(Te_new, double, nzone) \
    X(n_pos, double, nzone * num_nt) X(dn_pp, double, nzone * num_nt) \
    X(dne_pa, double, nzone * num_nt) X(dnp_pa, double, nzone * num_nt) \
    X(F_IC, double, num_nt * nphfield) X(fp_steps, int, nzone)

typedef struct orc_state {
    /* dims */
    int nz, nr, nzone, n_vol, num_nt, nphfield, n_gg, nphtotal, nph_lc, nmu;
    /* switches / scalars (commonblock.f) */
    int pair_switch, cr_sent, spec_switch, rand_switch, ncycle, star_switch;
    int split1, split2, split3, spl3_trg;
    int inj_switch, inj_dis, g2var_switch, pick_sw, cf_sentinel;
    int fp_dead_terms;        /* 1 = also evaluate the numerically dead rate terms (log parity only) */
    double rmin, zmin, time, dt;
    double inj_g1, inj_g2, inj_p, inj_t, inj_gg, inj_sigma, inj_L, inj_v, pick_rate;
    double r_esc, r_acc, lnL;
    double r_flare, z_flare, t_flare, sigma_r, sigma_z, sigma_t, flare_amp;
    /* FP accumulators (to_fp_calc, d_update) */
    double E_tot_old, E_tot_new, hr_total, hr_st_total, f_t_implicit, dT_max, dr, dz;
    /* event-related */
    double t_bound;
    /* counters (commonblock.f:134) */
    int64_t incounter, outcounter, cmcount1, cmcount2, lkcount, n_multiple, n_killed;
    int64_t n_segments, n_trials, n_comtot;     /* work statistics (not in the reference) */
    /* current photon (commonblock.f:126,181) */
    double xnu, wmu, phi, rpre, zpre, dcen, ew;
    int jph, kph, jgpsp, jgplc, jgpmu, idead;
    /* census banks: 6 doubles + 6 ints per photon (imctrk2d.f:566-580) (+ philox stream id) */
    double *dbufout, *dbufin;
    int *ibufout, *ibufin;
    uint64_t *sidout, *sidin;
    int64_t ndxout, ndxin, cens_cap;
    /* event list: 7 doubles per record (imcleak2d.f:181) */
    double *events;
    int64_t nev, ev_cap;
    int fatal;                /* set instead of Fortran STOP */
    orc_rng rng;
#define X(name, type, count) type *name;
    ORC_FIELDS(X)
#undef X
} orc_state;

/* life cycle */
orc_state *orc_create(int nz, int nr, int n_vol, int num_nt, int nphfield, int n_gg, int nphtotal,
                      int nph_lc, int nmu, int64_t cens_cap, int64_t ev_cap);
void orc_destroy(orc_state *s);
/* generic accessors used by the ctypes harness */
void *orc_array(orc_state *s, const char *name, int64_t *count, int *elem_size);
int orc_set_scalar(orc_state *s, const char *name, double v);
double orc_get_scalar(orc_state *s, const char *name);

/* RNG */
void orc_ran1_seed(orc_state *s, int idum, const int *iv, int iy);   /* restore rand.f SAVE state */
void orc_ran1_get(orc_state *s, int *idum, int *iv, int *iy);
/* rand_switch = 1 (o["rand_switch"] = 1 selects it): seed_zone (rand.f:9-88; *seed_io <- seeds(1,1)), RANX / randlist / randcounter */
void orc_seed_zone(orc_state *s, int *seed_io);
void orc_knuth_get(orc_state *s, double *ranx, double *randlist, int *randcounter);
void orc_knuth_set(orc_state *s, const double *ranx, const double *randlist, int randcounter);
void orc_philox_key(orc_state *s, uint32_t seed, uint32_t cycle);
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t orc_mix64(uint64_t a, uint64_t b);
double orc_philox_uniform(const uint32_t key[2], uint64_t sid, uint32_t d);
void orc_trial_uniforms(const uint32_t key[2], uint64_t S, int n, double *u);
double orc_abs_uniform(const uint32_t key[2], uint64_t S, uint32_t t, uint32_t k);
double orc_fibran(orc_state *s);

/* step prologue: emissivity / opacity tables from the live electron spectrum and the photon budget */
void orc_volume_em(orc_state *s, int zone, double l_min);            /* volume2d.f:10-394, live (synchrotron) terms */
void orc_imcgen_volume(orc_state *s, int nst);                         /* imcgen2d.f:203-336,405-459: Eloss_*, nsv, ewsv */

/* event-list reduction: the per-photon transform and binning of postprocessing/plcm.c:384-450 and
 * postprocessing/pspt.c:249-289 */
void orc_events_reduce(int64_t n, const double *events, double gam_bulk, double rmax, double t_offset, int n_t,
                       const double *t_lo, const double *t_hi, int n_mu, const double *mu0, const double *mu1, int mu_closed,
                       int n_bins, const double *E0, const double *E1, int first_band_only,
                       double *sum_ew, double *sum_ew2, int64_t *count);

/* surface sources (imcsurf2d_para.f) */
void orc_planck(orc_state *s, double tpl);                             /* planck2d.f:1-141 */
void orc_file_sample(orc_state *s, int spec);                          /* imcsurf2d_para.f:694-788 */
void orc_z_surf_calc(orc_state *s, int js);                            /* imcsurf2d_para.f:228-343 */
void orc_r_surf_calc(orc_state *s, int ks);                            /* imcsurf2d_para.f:353-534 */
int orc_file_sp(int n_rows, const double *E_in, const double *L_disk, const double *F_blr_in, const double *F_ir_in,
                double g_bulk, double R_blr, double fr_blr, double R_ir, double fr_ir, double R_disk, double d_jet,
                double *E_file, double *F_file, double *P_file, double *a1, double *I_file, double *int_file);
                                                                       /* imcsurf2d_para.f:544-685 */

/* leaves */
double orc_dilog(double x);                                            /* comtot2d.f:356-433 */
double orc_intg_v(double x);                                           /* comtot2d.f:337-352 */
double orc_comtot(orc_state *s, int j, int k, double xnuc);            /* comtot2d.f:219-247 */
void orc_nth2d(orc_state *s, int j, int k, double *gamm, double *betb, int *i); /* nontherm2d.f:159 */
void orc_compb2d(orc_state *s, int *i_gam);                            /* compb_2d.f */
void orc_get_bin(orc_state *s);                                        /* imcleak2d.f:329-405 */
void orc_imcleak(orc_state *s);                                        /* imcleak2d.f:2-320 */
void orc_imctrk2d(orc_state *s, int scat_flag);                        /* imctrk2d.f:8-716 */
void orc_vol_calc(orc_state *s, int zone);                             /* imcvol2d_para.f:90-414 */
void orc_field_calc(orc_state *s);                                     /* imcfield2d.f:57-148 */
void orc_step_zero_worker(orc_state *s);                               /* imcgen2d.f:559-601 */

/* Fokker-Planck */
double orc_gammln(double xx);                                          /* volume2d.f:649-668 */
double orc_McDonald(double nu, double z);                              /* volume2d.f:598-626 */
double orc_gamma_bar(double Theta);                                    /* volume2d.f:572-594 */
void orc_IC_loss(orc_state *s);                                        /* icloss2d.f:1-125 */
void orc_tridag(int n, const double *a, const double *b, const double *c, const double *f_old,
                double *f_new);                                        /* update2d.f:2485-2528 */
void orc_fp_calc(orc_state *s, int zone);                              /* update2d.f:337-1741 */

#ifdef __cplusplus
}
#endif
#endif
