// c2d_step.cu -- one closed IMC time step on a device-resident plasma state (host code; kernels live in the other units).
//
// The reference's step is the body of xec's loop (xec2d.f:100-110 master / 141-193 worker):
//     xec_add -> imcgen2d -> imcfield2d -> imcvol2d -> imcsurf2d -> imcredist -> update -> graphics_collect -> graphics
// with the plasma (f_nt, Pnt, n_e, tea, ...) living in the master's COMMON blocks, shipped to the workers every step
// (vol_bcast, z_surf_bcast, FP_bcast: ~30 array broadcasts) and shipped back zone by zone after the Fokker-Planck update
// (FP_recv_result, FP_end_bcast).  Here the plasma stays in HBM on every rank: c2d_plasma_set once, then c2d_step per
// cycle does
//   1. ec_old <- last step's reduced ecens (imcgen2d.f:90-96), bingo (l.405-421)
//   2. volume_em on the device from the live f_nt (csrc/c2d_gen.cu) + the volume budget (imcgen2d.f:280-283,330,444-452)
//   3. the surface budget and the > 10 nst cap (imcgen2d.f:120-190,424-443,457-517: host arithmetic over nz + nr rings)
//   4. zero the tallies in the reference's order (ecens / n_field only after they fed step 1; imcgen2d.f:87-104,534-551)
//   5. field_calc, vol_calc, surface sources (this rank's share of the histories) and the scattering cascade
//   6. the tally all-reduce (c2d_tally_reduce)
//   7. if T_const = 0 and ncycle > 0: FP_calc for this rank's zones straight from the reduced n_field / ecens in HBM, the
//      all-gather of the new electron state, tea <- clamp(Te_new) (update2d.f:262-276), the sums of E_add_up
//      (update2d.f:2002-2083) and the time-step controller (update2d.f:233-243: dt_new is reported, not applied -- as in
//      the reference, l.246-261)
// and returns a report holding the scalars the reference prints to log.txt.  The host advances time and ncycle
// (xec2d.f:94-103).
//
// What "as the reference" means here: the MPI build's workers see several inputs one broadcast late.  The
// step follows the evident intent -- tables of the current cycle everywhere -- with ONE optional quirk, c2d_run.quirk_cycle0:
// in cycle 0 the workers still have dt = 0 and no energy groups, so every source photon goes to census unflown, at its
// birth point, with bins 0 (never tallied in fout / edout until it scatters).  Runs that are compared with the binary's
// golden values need it, because the census that cycle 1 re-tracks is exactly that.
#include "c2d_ctx.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace c2d {

#define CUS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(C2D_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static inline size_t nz_pad(const Ctx *c) { return ((size_t)c->nzone + 1) & ~size_t(1); }
static inline double *pl_row(Ctx *c, int row) { return c->d_plasma + (size_t)row * nz_pad(c); }

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_plasma_set(c2d_ctx *h, const c2d_plasma *p)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !p) return fail(C2D_EINVAL, "NULL argument");
    if (!c->inited) return fail(C2D_ESTATE, "c2d_init has not been called");
    CUS(cudaSetDevice(c->device));
    const int nzone = c->nzone, N = c->dims.num_nt;
    const bool first = !c->plasma_set;
    if (first) {
        if (!p->f_nt || !p->Pnt || !p->n_e || !p->tea || !p->tna || !p->B_field || !p->vol || !p->zsurf || !p->gmin || !p->gmax ||
            !p->p_nth || !p->amxwl)
            return fail(C2D_EINVAL, "c2d_plasma_set: the first call needs f_nt, Pnt, n_e, tea, tna, B_field, vol, zsurf, gmin, gmax, p_nth, amxwl");
        CUS(cudaMalloc(&c->d_plasma, (size_t)PL_ROWS * nz_pad(c) * 8));
        CUS(cudaMemset(c->d_plasma, 0, (size_t)PL_ROWS * nz_pad(c) * 8));
        c->h_ne.assign(nzone, 0.0); c->h_B.assign(nzone, 0.0); c->h_vol.assign(nzone, 0.0); c->h_zsurf.assign(nzone, 0.0);
        c->h_tna.assign(nzone, 0.0);
    }
    double *T = c->d_step;
    auto up = [&](double *dst, const double *src, size_t n) -> int {
        if (src) CUS(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyHostToDevice, c->stream));
        return 0;
    };
    int rc = 0;
    rc |= up(T + c->o_fnt, p->f_nt, (size_t)nzone * N);
    rc |= up(T + c->o_pnt, p->Pnt, (size_t)nzone * N);
    rc |= up(T + c->o_ne, p->n_e, nzone);
    rc |= up(T + c->o_tea, p->tea, nzone);
    rc |= up(T + c->o_fpair, p->f_pair, nzone);
    rc |= up(T + c->o_zsurf, p->zsurf, nzone);
    rc |= up(T + c->o_kgg, p->k_gg, (size_t)nzone * c->dims.n_gg);
    rc |= up(pl_row(c, PL_TNA), p->tna, nzone);
    rc |= up(pl_row(c, PL_B), p->B_field, nzone);
    rc |= up(pl_row(c, PL_GMIN), p->gmin, nzone);
    rc |= up(pl_row(c, PL_GMAX), p->gmax, nzone);
    rc |= up(pl_row(c, PL_PNTH), p->p_nth, nzone);
    rc |= up(pl_row(c, PL_AMXWL), p->amxwl, nzone);
    rc |= up(pl_row(c, PL_TURB), p->turb_lev, nzone);
    rc |= up(pl_row(c, PL_VOL), p->vol, nzone);
    if (rc) return rc;
    CUS(cudaStreamSynchronize(c->stream));
    if (p->n_e) c->h_ne.assign(p->n_e, p->n_e + nzone);
    if (p->B_field) c->h_B.assign(p->B_field, p->B_field + nzone);
    if (p->vol) c->h_vol.assign(p->vol, p->vol + nzone);
    if (p->zsurf) c->h_zsurf.assign(p->zsurf, p->zsurf + nzone);
    if (p->tna) c->h_tna.assign(p->tna, p->tna + nzone);
    c->plasma_set = true;
    return 0;
}

int c2d_plasma_get(c2d_ctx *h, c2d_plasma_out *p)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !p) return fail(C2D_EINVAL, "NULL argument");
    if (!c->plasma_set) return fail(C2D_ESTATE, "c2d_plasma_set has not been called");
    CUS(cudaSetDevice(c->device));
    const int nzone = c->nzone, N = c->dims.num_nt;
    double *T = c->d_step;
    auto down = [&](double *dst, const double *src, size_t n) -> int {
        if (dst) CUS(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, c->stream));
        return 0;
    };
    int rc = 0;
    rc |= down(p->f_nt, T + c->o_fnt, (size_t)nzone * N);
    rc |= down(p->Pnt, T + c->o_pnt, (size_t)nzone * N);
    rc |= down(p->n_e, T + c->o_ne, nzone);
    rc |= down(p->tea, T + c->o_tea, nzone);
    rc |= down(p->f_pair, T + c->o_fpair, nzone);
    rc |= down(p->B_field, pl_row(c, PL_B), nzone);
    rc |= down(p->gmin, pl_row(c, PL_GMIN), nzone);
    rc |= down(p->gmax, pl_row(c, PL_GMAX), nzone);
    rc |= down(p->p_nth, pl_row(c, PL_PNTH), nzone);
    rc |= down(p->amxwl, pl_row(c, PL_AMXWL), nzone);
    rc |= down(p->Te_new, pl_row(c, PL_TE_NEW), nzone);
    rc |= down(p->Eloss_sy, pl_row(c, PL_ELOSS_SY), nzone);
    rc |= down(p->ec_old, pl_row(c, PL_EC_OLD), nzone);
    if (rc) return rc;
    CUS(cudaStreamSynchronize(c->stream));
    return 0;
}

// == the surface half of imcgen2d (imcgen2d.f:120-190 energies entering through each boundary ring, 424-443 photon numbers,
// 457-476 weights): host arithmetic over nz + nr rings.  Fills the count / weight arrays of `s` (caller-owned storage).
static void surface_budget(const Ctx *c, const c2d_run *run, double time, double dt, std::vector<int32_t> n[4], std::vector<double> ew[4],
                           std::vector<double> erin[4])
{
    const c2d_surf_run *sr = run->surf;
    const int nz = c->dims.nz, nr = c->dims.nr;
    const double sigma = 1.0267e24, pi = kPi;                     // imcgen2d.f:14, general.pa:25
    const std::vector<double> &z = c->h_z, &r = c->h_r;
    const double rmin = c->dims.rmin, zmin = c->dims.zmin;
    const double r_nr = r[nr - 1];
    for (int f = 0; f < 4; f++) {                                  // 0 inner, 1 outer, 2 upper, 3 lower
        const int nring = (f < 2) ? nz : nr;
        n[f].assign(nring, 0); ew[f].assign(nring, 0.0); erin[f].assign(nring, 0.0);
        if (!sr) continue;
        const double *tbb = (f == 0) ? sr->tbbi : (f == 1) ? sr->tbbo : (f == 2) ? sr->tbbu : sr->tbbl;
        const int32_t *spec = (f == 0) ? sr->spec_i : (f == 1) ? sr->spec_o : (f == 2) ? sr->spec_u : sr->spec_l;
        if (!tbb) continue;
        for (int q = 0; q < nring; q++) {
            double A;                                              // setup2d.f:80-114 Asurf*
            if (f < 2) {
                const double zlo = (q == 0) ? zmin : z[q - 1];
                A = 2.0 * pi * ((f == 0) ? rmin : r_nr) * (z[q] - zlo);
            } else {
                const double rlo = (q == 0) ? rmin : r[q - 1];
                A = pi * (r[q] * r[q] - rlo * rlo);
            }
            const double t = tbb[q];
            const bool from_file = (t < 0.0) && (time + 0.5 * dt >= sr->t0);
            if (from_file) {
                const int sp = spec ? spec[q] : 0;
                const double intf = (sr->int_file && sp >= 0 && sp < sr->n_spec) ? sr->int_file[sp] : 0.0;
                erin[f][q] = dt * A * intf;
            } else {
                erin[f][q] = dt * A * sigma * std::pow(t, 4.0);
                if (f == 2 && sr->star_switch == 1) erin[f][q] = erin[f][q] * std::pow(sr->Rstar / sr->dist_star, 2.0);
            }
            if (t < 0.0) {
                if (f < 2) n[f][q] = (int32_t)(run->nst / nz);                                  // integer division, l.427-428
                else {
                    const double rlo = (q == 0) ? rmin : r[q - 1];                               // r(0): rmin of COMMON /zones/
                    n[f][q] = (int32_t)((double)run->nst * (r[q] * r[q] - rlo * rlo) / (r_nr * r_nr));   // truncation, l.435-436
                }
            }
        }
    }
    for (int f = 0; f < 4; f++)
        for (size_t q = 0; q < n[f].size(); q++) ew[f][q] = (n[f][q] > 0) ? erin[f][q] / (double)n[f][q] : 0.0;
}

int c2d_step(c2d_ctx *h, const c2d_run *run, int ncycle, double time, double dt, c2d_step_report *rep)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !run || !rep) return fail(C2D_EINVAL, "NULL argument");
    if (!c->inited) return fail(C2D_ESTATE, "c2d_init has not been called");
    if (!c->plasma_set) return fail(C2D_ESTATE, "c2d_plasma_set has not been called");
    if (!(dt > 0.0) || run->nst < 0) return fail(C2D_EINVAL, "dt = %g, nst = %lld", dt, (long long)run->nst);
    CUS(cudaSetDevice(c->device));
    const c2d_dims &d = c->dims;
    const int nzone = c->nzone, nz = d.nz, nr = d.nr;
    const int world = comm_world(c), rank = comm_rank(c);
    memset(rep, 0, sizeof *rep);
    rep->ncycle = ncycle; rep->time = time; rep->dt = dt;
    double *T = c->d_step;

    // 1. ec_old <- the census energy the previous step left (reduced over the ranks); bingo needs its sum
    if (int rc = latch_census(c)) return rc;
    rep->n_census_in = c->n_census;
    std::vector<double> ecens_prev(nzone, 0.0);
    if (ncycle > 0) {
        CUS(cudaMemcpyAsync(pl_row(c, PL_EC_OLD), c->d_f64 + c->t_ecens, (size_t)nzone * 8, cudaMemcpyDeviceToDevice, c->stream));
        CUS(cudaMemcpyAsync(ecens_prev.data(), c->d_f64 + c->t_ecens, (size_t)nzone * 8, cudaMemcpyDeviceToHost, c->stream));
    } else CUS(cudaMemsetAsync(pl_row(c, PL_EC_OLD), 0, (size_t)nzone * 8, c->stream));

    // 2. volume_em + volume budget from the resident state
    GenBudget g;
    if (int rc = gen_core(c, dt, run->nst, pl_row(c, PL_B), c->h_ne.data(), c->h_B.data(), c->h_vol.data(), c->h_zsurf.data(), g)) return rc;
    CUS(cudaMemcpyAsync(pl_row(c, PL_ELOSS_SY), g.e_sy.data(), (size_t)nzone * 8, cudaMemcpyHostToDevice, c->stream));

    // 3. surface budget, bingo, bias cap
    std::vector<int32_t> ns[4];
    std::vector<double> ews[4], erin[4];
    surface_budget(c, run, time, dt, ns, ews, erin);
    double bingo = 0.0;
    for (int zi = 0; zi < nzone; zi++) bingo = bingo + 1.0 * (ecens_prev[zi] + g.e_sy[zi]);      // l.413-417 (bias = 1)
    for (int j = 0; j < nz; j++) bingo = bingo + erin[0][j] + erin[1][j];
    for (int k = 0; k < nr; k++) bingo = bingo + erin[2][k] + erin[3][k];
    long long n_new = 0;
    for (int f = 0; f < 4; f++) for (int32_t v : ns[f]) n_new += v;
    for (int32_t v : g.nsv) n_new += v;
    double fbias = 1.0;
    if (n_new > 10 * (long long)run->nst) {                        // l.491-517
        fbias = (double)(10 * run->nst) / (double)n_new;
        for (int f = 0; f < 4; f++)
            for (size_t q = 0; q < ns[f].size(); q++) { ns[f][q] = (int32_t)((double)ns[f][q] * fbias); ews[f][q] = ews[f][q] / fbias; }
        for (int zi = 0; zi < nzone; zi++) { g.nsv[zi] = (int32_t)(g.nsv[zi] * fbias); g.ew[zi] = g.ew[zi] / fbias; }
    }
    if (int rc = budget_upload(c, g)) return rc;
    rep->bingo = bingo; rep->Emiss_tot = g.Emiss_tot; rep->fbias = fbias;
    rep->n_volume = c->h_prefix.back();
    long long n_surf = 0;
    for (int f = 0; f < 4; f++) for (int32_t v : ns[f]) n_surf += v;
    rep->n_surface = n_surf;
    if (run->surf && run->surf->tbbl)      // the lower-face leak test of imcleak reads tbbl even when the ring emits nothing
        CUS(cudaMemcpyAsync(T + c->o_tbbl, run->surf->tbbl, (size_t)nr * 8, cudaMemcpyHostToDevice, c->stream));

    // 4./5. open the step (zeroes every per-step tally) and track
    const bool quirk0 = run->quirk_cycle0 && ncycle == 0;
    if (int rc = begin_common(c, ncycle, time, quirk0 ? 0.0 : dt)) return rc;
    c->P.bins_off = quirk0 ? 1 : 0;
    int rc = 0;
    if (c->n_census > 0) rc = c2d_imcfield(h);
    if (!rc) rc = c2d_imcvol(h, rank, -1, world);
    if (!rc && n_surf > 0) {
        const c2d_surf_run *sr = run->surf;
        c2d_surf_in si;
        memset(&si, 0, sizeof si);
        si.star_switch = sr->star_switch;
        si.nsurfi = ns[0].data(); si.nsurfo = ns[1].data(); si.nsurfu = ns[2].data(); si.nsurfl = ns[3].data();
        si.ewsurfi = ews[0].data(); si.ewsurfo = ews[1].data(); si.ewsurfu = ews[2].data(); si.ewsurfl = ews[3].data();
        si.tbbi = sr->tbbi; si.tbbo = sr->tbbo; si.tbbu = sr->tbbu; si.tbbl = sr->tbbl;
        si.spec_i = sr->spec_i; si.spec_o = sr->spec_o; si.spec_u = sr->spec_u; si.spec_l = sr->spec_l;
        si.n_spec = sr->n_spec; si.nfmax = sr->nfmax; si.nfile = sr->nfile;
        si.E_file = sr->E_file; si.F_file = sr->F_file; si.P_file = sr->P_file; si.a1 = sr->a1; si.I_file = sr->I_file;
        rc = c2d_surf_begin(h, &si);
        if (!rc) rc = c2d_imcsurf(h, rank, -1, world);
    }
    c->P.bins_off = 0;
    if (rc) return rc;

    // 6. cens_add_up + E_add_up + xec_add + graphics_collect
    if ((rc = c2d_tally_reduce(h, nullptr)) != 0) return rc;
    {
        unsigned long long cg[CNT_N];
        std::vector<double> lk((size_t)2 * nz + 2 * nr);
        CUS(cudaMemcpyAsync(cg, c->d_cnt_global, sizeof cg, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaMemcpyAsync(lk.data(), c->d_f64 + c->t_erlki, (size_t)nz * 8, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaMemcpyAsync(lk.data() + nz, c->d_f64 + c->t_erlko, (size_t)nz * 8, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaMemcpyAsync(lk.data() + 2 * nz, c->d_f64 + c->t_erlku, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaMemcpyAsync(lk.data() + 2 * nz + nr, c->d_f64 + c->t_erlkl, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaStreamSynchronize(c->stream));
        rep->ndxout = (int64_t)cg[CNT_NDXOUT]; rep->nevents = (int64_t)cg[CNT_NEVENTS];
        rep->cmcount1 = (int64_t)cg[CNT_CMCOUNT1]; rep->cmcount2 = (int64_t)cg[CNT_CMCOUNT2];
        rep->histories = (int64_t)cg[CNT_HISTORIES];
        // E_add_up's boundary terms (update2d.f:2038-2049), summed in its order
        double E_in = 0.0, E_out = 0.0;
        for (int j = 0; j < nz; j++) { E_in = E_in + erin[0][j] + erin[1][j]; E_out = E_out + lk[j] + lk[nz + j]; }
        for (int k = 0; k < nr; k++) { E_in = E_in + erin[2][k] + erin[3][k]; E_out = E_out + lk[2 * nz + k] + lk[2 * nz + nr + k]; }
        rep->E_tot_old = E_in; rep->E_tot_new = E_out; rep->luminosity = E_out / dt;
    }

    // 7. update (xec2d.f:87-88: only when T_const = 0 and ncycle > 0)
    if (run->T_const == 0 && ncycle > 0) {
        c2d_fp_in fi;
        memset(&fi, 0, sizeof fi);
        fi.inj_switch = run->inj_switch; fi.inj_dis = run->inj_dis; fi.g2var_switch = run->g2var_switch; fi.pick_sw = run->pick_sw;
        fi.cf_sentinel = run->cf_sentinel; fi.ncycle = ncycle; fi.time = time; fi.dt = dt; fi.lnL = 20.0;   // update2d.f:139
        fi.inj_g1 = run->inj_g1; fi.inj_g2 = run->inj_g2; fi.inj_p = run->inj_p; fi.inj_t = run->inj_t; fi.inj_gg = run->inj_gg;
        fi.inj_sigma = run->inj_sigma; fi.inj_L = run->inj_L; fi.inj_v = run->inj_v; fi.pick_rate = run->pick_rate;
        fi.r_esc = run->r_esc; fi.r_acc = run->r_acc;
        fi.r_flare = run->r_flare; fi.z_flare = run->z_flare; fi.t_flare = run->t_flare; fi.sigma_r = run->sigma_r;
        fi.sigma_z = run->sigma_z; fi.sigma_t = run->sigma_t; fi.flare_amp = run->flare_amp;
        std::vector<double> zi_((size_t)nzone * kFpRecExtra);
        if ((rc = fp_resident(c, &fi, zi_.data())) != 0) return rc;
        double dT_max = 0.0, hr = 0.0, hrst = 0.0, Eo = rep->E_tot_old, En = rep->E_tot_new;
        int smin = 1 << 30, smax = 0;
        for (int z = 0; z < nzone; z++) {
            const double *e = zi_.data() + (size_t)z * kFpRecExtra;
            c->h_ne[z] = e[0];
            Eo = Eo + e[7]; En = En + e[8]; hr = hr + e[9]; hrst = hrst + e[10];
            if (e[11] > dT_max) dT_max = e[11];
            smin = std::min(smin, (int)e[12]); smax = std::max(smax, (int)e[12]);
        }
        const double df_T = 0.25;                                  // general.pa:28
        if (ncycle <= 1 && dT_max < df_T) dT_max = df_T;           // photon_fill leaves dT_max = df_T (update2d.f:1917), E_add_up takes the MAX
        rep->fp_ran = 1; rep->fp_steps_min = smin; rep->fp_steps_max = smax;
        rep->E_tot_old = Eo; rep->E_tot_new = En; rep->hr_total = hr; rep->hr_st_total = hrst; rep->dT_max = dT_max;
        rep->dE_fraction = (En - Eo) / Eo;
        // update2d.f:233-243 (logged; the reference does not apply it: l.246-261)
        double dt_new;
        if (dT_max < (0.2 * df_T)) dt_new = 3.0 * dt;
        else if (dT_max < (.75 * df_T)) dt_new = 1.1 * dt;
        else if (dT_max > (5. * df_T)) dt_new = 0.33 * dt;
        else if (dT_max > (1.25 * df_T)) dt_new = 7.5e-1 * dt;
        else dt_new = dt;
        rep->dt_new = dt_new;
    } else {
        rep->dt_new = dt;
    }
    for (int k = 0; k < Ctx::kFamilies && k < 16; k++) { c->timer_collect(); rep->ms[k] = c->k_ms[k]; }
    return 0;
}

}  // extern "C"
