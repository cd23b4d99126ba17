// c2d_common.cu -- the COMMON-block side of the boundary (host code).
//
// The reference's subroutines take no arguments: everything travels through COMMON blocks whose arrays are column-major
// at their padded compile-time extents (general.pa:7-31: jmax = kmax = 99, ...; commonblock.f:1-236).  A Fortran shim that
// keeps the reference's subroutine names (integration/c2d_shim.f90) therefore has exactly one job: hand the addresses of
// those arrays to the library.  The packing into the dense [zone][bin] layout of the C ABI, the call, and the scatter of
// the results back into COMMON happen here, in one place, where they can be tested without a Fortran compiler
// (tests/test_common_bridge.py lays the blocks out with numpy in Fortran order at the padded extents).
//
//   element (j, k)     of a(jmax, kmax)          at (k-1) jmax + (j-1)                    dense zone = (k-1) + nr (j-1)
//   element (i, j, k)  of a(nbin, jmax, kmax)    at ((k-1) jmax + (j-1)) nbin + (i-1)     "bin first" (kappa_tot, n_field, k_gg)
//   element (j, k, i)  of a(jmax, kmax, nbin)    at ((i-1) kmax + (k-1)) jmax + (j-1)     "bin last"  (f_nt, Pnt)
#include "c2d_ctx.h"

#include <cstring>
#include <vector>

namespace c2d {

struct ZoneMap {
    int nz, nr, jmax, kmax;
    size_t at(int zi) const { const int j = zi / nr, k = zi - nr * j; return (size_t)k * jmax + j; }
};
static void gather_zone(const ZoneMap &m, const double *common, std::vector<double> &dense)
{
    dense.assign((size_t)m.nz * m.nr, 0.0);
    if (common) for (int zi = 0; zi < m.nz * m.nr; zi++) dense[zi] = common[m.at(zi)];
}
static void scatter_zone(const ZoneMap &m, const double *dense, double *common)
{
    if (common) for (int zi = 0; zi < m.nz * m.nr; zi++) common[m.at(zi)] = dense[zi];
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_common_step_begin(c2d_ctx *h, const c2d_common *cm, int ncycle, double time, double dt)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !cm) return fail(C2D_EINVAL, "NULL argument");
    if (!c->inited) return fail(C2D_ESTATE, "c2d_init has not been called");
    const c2d_dims &d = c->dims;
    if (cm->jmax < d.nz || cm->kmax < d.nr) return fail(C2D_EINVAL, "COMMON extents %d x %d smaller than the grid %d x %d", cm->jmax, cm->kmax, d.nz, d.nr);
    if (!cm->kappa_tot || !cm->eps_tot || !cm->eps_th || !cm->f_nt || !cm->Pnt || !cm->n_e || !cm->nsv || !cm->ewsv)
        return fail(C2D_EINVAL, "c2d_common_step_begin: missing COMMON member");
    const ZoneMap m{d.nz, d.nr, cm->jmax, cm->kmax};
    const int nzone = c->nzone;
    std::vector<double> kappa((size_t)nzone * d.n_vol), etot((size_t)nzone * d.n_vol), eth((size_t)nzone * d.n_vol),
        fnt((size_t)nzone * d.num_nt), pnt((size_t)nzone * d.num_nt), kgg((size_t)nzone * d.n_gg);
    c2d_pack_bin_first(cm->kappa_tot, d.n_vol, cm->jmax, cm->kmax, d.n_vol, d.nz, d.nr, kappa.data());
    c2d_pack_bin_first(cm->eps_tot, d.n_vol, cm->jmax, cm->kmax, d.n_vol, d.nz, d.nr, etot.data());
    c2d_pack_bin_first(cm->eps_th, d.n_vol, cm->jmax, cm->kmax, d.n_vol, d.nz, d.nr, eth.data());
    c2d_pack_bin_last(cm->f_nt, cm->jmax, cm->kmax, d.num_nt, d.nz, d.nr, fnt.data());
    c2d_pack_bin_last(cm->Pnt, cm->jmax, cm->kmax, d.num_nt, d.nz, d.nr, pnt.data());
    if (cm->k_gg) c2d_pack_bin_first(cm->k_gg, d.n_gg, cm->jmax, cm->kmax, d.n_gg, d.nz, d.nr, kgg.data());
    std::vector<double> ne, tea, fpair, ewsv, elth, eltot, zsurf;
    gather_zone(m, cm->n_e, ne); gather_zone(m, cm->tea, tea); gather_zone(m, cm->f_pair, fpair); gather_zone(m, cm->ewsv, ewsv);
    gather_zone(m, cm->Eloss_th, elth); gather_zone(m, cm->Eloss_tot, eltot); gather_zone(m, cm->zsurf, zsurf);
    std::vector<int32_t> nsv(nzone);
    for (int zi = 0; zi < nzone; zi++) nsv[zi] = cm->nsv[m.at(zi)];
    std::vector<double> tbbl(d.nr, 0.0);
    if (cm->tbbl) for (int k = 0; k < d.nr; k++) tbbl[k] = cm->tbbl[(size_t)(cm->ti > 0 ? cm->ti - 1 : 0) * cm->kmax + k];   // tbbl(k, ti)
    c2d_step_in s;
    memset(&s, 0, sizeof s);
    s.ncycle = ncycle; s.time = time; s.dt = dt;
    s.kappa_tot = kappa.data(); s.eps_tot = etot.data(); s.eps_th = eth.data(); s.f_nt = fnt.data(); s.Pnt = pnt.data();
    s.n_e = ne.data(); s.tea = tea.data(); s.f_pair = cm->f_pair ? fpair.data() : nullptr; s.k_gg = cm->k_gg ? kgg.data() : nullptr;
    s.nsv = nsv.data(); s.ewsv = ewsv.data(); s.Eloss_th = cm->Eloss_th ? elth.data() : nullptr;
    s.Eloss_tot = cm->Eloss_tot ? eltot.data() : nullptr; s.zsurf = cm->zsurf ? zsurf.data() : nullptr; s.tbbl = tbbl.data();
    return c2d_step_begin(h, &s);
}

// == cens_add_up + E_add_up + xec_add + graphics_collect: reduce over the ranks, then put every tally where the reference's
// master has it after those calls (fout: the run-long sum, xec2d.f:389-394)
int c2d_common_collect(c2d_ctx *h, c2d_common *cm)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !cm) return fail(C2D_EINVAL, "NULL argument");
    const c2d_dims &d = c->dims;
    if (cm->jmax < d.nz || cm->kmax < d.nr || cm->nmumax < d.nmu || cm->nphomax < d.nphtotal || cm->nphlcmax < d.nph_lc)
        return fail(C2D_EINVAL, "COMMON extents smaller than the problem");
    const ZoneMap m{d.nz, d.nr, cm->jmax, cm->kmax};
    const int nzone = c->nzone;
    std::vector<double> edep(nzone), prdep(nzone), ecens(nzone), nfield((size_t)nzone * d.nphfield), nph((size_t)nzone * d.n_gg), eic(d.num_nt),
        fout((size_t)C2D_NMUMAX * d.nphtotal), ftot((size_t)C2D_NMUMAX * d.nphtotal), edout((size_t)C2D_NMUMAX * C2D_NLCMAX), lki(d.nz), lko(d.nz), lku(d.nr),
        lkl(d.nr), edin(d.nr), edref(d.nr);
    std::vector<int64_t> npcen(nzone), nel(d.num_nt + 2);
    c2d_tallies t;
    memset(&t, 0, sizeof t);
    t.edep = edep.data(); t.prdep = prdep.data(); t.ecens = ecens.data(); t.npcen = npcen.data(); t.n_field = nfield.data();
    t.n_ph = nph.data(); t.E_IC = eic.data(); t.nelectron = nel.data(); t.fout = fout.data(); t.fout_total = ftot.data(); t.edout = edout.data();
    t.erlki = lki.data(); t.erlko = lko.data(); t.erlku = lku.data(); t.erlkl = lkl.data(); t.Ed_in = edin.data(); t.Ed_ref = edref.data();
    if (int rc = c2d_tally_reduce(h, &t)) return rc;
    scatter_zone(m, edep.data(), cm->edep); scatter_zone(m, prdep.data(), cm->prdep); scatter_zone(m, ecens.data(), cm->ecens);
    if (cm->npcen) for (int zi = 0; zi < nzone; zi++) cm->npcen[m.at(zi)] = (int32_t)npcen[zi];
    if (cm->n_field) c2d_unpack_bin_first(nfield.data(), d.nphfield, cm->jmax, cm->kmax, d.nphfield, d.nz, d.nr, cm->n_field);
    if (cm->n_ph) c2d_unpack_bin_first(nph.data(), d.n_gg, cm->jmax, cm->kmax, d.n_gg, d.nz, d.nr, cm->n_ph);
    if (cm->E_IC) memcpy(cm->E_IC, eic.data(), (size_t)d.num_nt * 8);
    if (cm->nelectron) for (int i = 0; i < d.num_nt; i++) cm->nelectron[i] = (int32_t)nel[i + 1];     // nelectron(i), i = 1 .. num_nt
    // fout(nmumax, nphomax), edout(nmumax, nphlcmax): column-major, first index = angle bin
    if (cm->fout) for (int n = 0; n < d.nmu; n++) for (int g = 0; g < d.nphtotal; g++) cm->fout[(size_t)g * cm->nmumax + n] = ftot[(size_t)n * d.nphtotal + g];
    if (cm->edout) for (int n = 0; n < d.nmu; n++) for (int l = 0; l < d.nph_lc; l++) cm->edout[(size_t)l * cm->nmumax + n] = edout[(size_t)n * C2D_NLCMAX + l];
    if (cm->erlki) memcpy(cm->erlki, lki.data(), (size_t)d.nz * 8);
    if (cm->erlko) memcpy(cm->erlko, lko.data(), (size_t)d.nz * 8);
    if (cm->erlku) memcpy(cm->erlku, lku.data(), (size_t)d.nr * 8);
    if (cm->erlkl) memcpy(cm->erlkl, lkl.data(), (size_t)d.nr * 8);
    if (cm->Ed_in) memcpy(cm->Ed_in, edin.data(), (size_t)d.nr * 8);
    if (cm->Ed_ref) memcpy(cm->Ed_ref, edref.data(), (size_t)d.nr * 8);
    if (cm->ndxout) *cm->ndxout = (int32_t)t.counters[C2D_CNT_NDXOUT];                // this rank's census, as in the reference
    if (cm->cmcount1) *cm->cmcount1 = (int32_t)t.counters[C2D_CNT_CMCOUNT1];
    if (cm->cmcount2) *cm->cmcount2 = (int32_t)t.counters[C2D_CNT_CMCOUNT2];
    if (cm->lkcount) *cm->lkcount = (int32_t)t.counters[C2D_CNT_LKCOUNT];
    return 0;
}

// == update minus its MPI (update2d.f:7-334): every rank solves every zone (the FP kernel takes ~2 ms for 4096 zones, far
// less than shipping 1830 + 830 doubles per zone through the host would), results land in COMMON where FP_end_bcast leaves
// them on every rank; tea is clamped as update2d.f:262-276 does.
int c2d_common_update(c2d_ctx *h, c2d_common *cm, const c2d_fp_in *scalars)
{
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c || !cm || !scalars) return fail(C2D_EINVAL, "NULL argument");
    const c2d_dims &d = c->dims;
    if (!cm->f_nt || !cm->Pnt || !cm->n_e || !cm->tea || !cm->tna || !cm->B_field || !cm->vol || !cm->Eloss_sy || !cm->ec_old ||
        !cm->gmin || !cm->gmax || !cm->p_nth || !cm->amxwl)
        return fail(C2D_EINVAL, "c2d_common_update: missing COMMON member");
    const ZoneMap m{d.nz, d.nr, cm->jmax, cm->kmax};
    const int nzone = c->nzone, N = d.num_nt;
    std::vector<double> fnt((size_t)nzone * N), pnt((size_t)nzone * N), ne, tea, tna, B, vol, fpair, sy, eco, tl, gmin, gmax, pn, am;
    c2d_pack_bin_last(cm->f_nt, cm->jmax, cm->kmax, N, d.nz, d.nr, fnt.data());
    c2d_pack_bin_last(cm->Pnt, cm->jmax, cm->kmax, N, d.nz, d.nr, pnt.data());
    gather_zone(m, cm->n_e, ne); gather_zone(m, cm->tea, tea); gather_zone(m, cm->tna, tna); gather_zone(m, cm->B_field, B);
    gather_zone(m, cm->vol, vol); gather_zone(m, cm->f_pair, fpair); gather_zone(m, cm->Eloss_sy, sy); gather_zone(m, cm->ec_old, eco);
    gather_zone(m, cm->turb_lev, tl); gather_zone(m, cm->gmin, gmin); gather_zone(m, cm->gmax, gmax); gather_zone(m, cm->p_nth, pn);
    gather_zone(m, cm->amxwl, am);
    c2d_fp_in fi = *scalars;
    fi.first_zone = 0; fi.n_zones = -1; fi.zone_stride = 1;
    fi.f_nt = fnt.data(); fi.Pnt = pnt.data(); fi.n_field = nullptr; fi.ecens = nullptr;      // the reduced tallies in HBM
    fi.n_e = ne.data(); fi.B_field = B.data(); fi.tea = tea.data(); fi.tna = tna.data(); fi.vol = vol.data();
    fi.f_pair = cm->f_pair ? fpair.data() : nullptr; fi.Eloss_sy = sy.data(); fi.ec_old = eco.data(); fi.turb_lev = tl.data(); fi.F_IC = nullptr;
    std::vector<double> ofnt((size_t)nzone * N), opnt((size_t)nzone * N), one(nzone), ote(nzone), ofp(nzone);
    c2d_fp_out fo;
    memset(&fo, 0, sizeof fo);
    fo.f_nt = ofnt.data(); fo.Pnt = opnt.data(); fo.n_e = one.data(); fo.Te_new = ote.data(); fo.gmin = gmin.data(); fo.gmax = gmax.data();
    fo.p_nth = pn.data(); fo.amxwl = am.data(); fo.f_pair = ofp.data();
    if (int rc = c2d_update_fp(h, &fi, &fo)) return rc;
    c2d_unpack_bin_last(ofnt.data(), cm->jmax, cm->kmax, N, d.nz, d.nr, cm->f_nt);
    c2d_unpack_bin_last(opnt.data(), cm->jmax, cm->kmax, N, d.nz, d.nr, cm->Pnt);
    scatter_zone(m, one.data(), cm->n_e); scatter_zone(m, ote.data(), cm->Te_new); scatter_zone(m, gmin.data(), cm->gmin);
    scatter_zone(m, gmax.data(), cm->gmax); scatter_zone(m, pn.data(), cm->p_nth); scatter_zone(m, am.data(), cm->amxwl);
    if (cm->f_pair) scatter_zone(m, ofp.data(), cm->f_pair);
    for (int zi = 0; zi < nzone; zi++)                                            // update2d.f:262-276
        if (tna[zi] > 1.) {
            double t = ote[zi];
            t = t < 1.e3 ? t : 1.e3;
            t = t > 5.0 ? t : 5.0;
            cm->tea[m.at(zi)] = t;
        }
    if (cm->dT_max) *cm->dT_max = fo.dT_max;
    if (cm->E_tot_old) *cm->E_tot_old = fo.E_tot_old;
    if (cm->E_tot_new) *cm->E_tot_new = fo.E_tot_new;
    if (cm->hr_total) *cm->hr_total = fo.hr_total;
    if (cm->hr_st_total) *cm->hr_st_total = fo.hr_st_total;
    if (cm->dr) *cm->dr = fo.dr;
    if (cm->dz) *cm->dz = fo.dz;
    if (cm->f_t_implicit) *cm->f_t_implicit = fo.f_t_implicit;
    return 0;
}

}  // extern "C"
