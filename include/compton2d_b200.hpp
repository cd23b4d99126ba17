// compton2d_b200.hpp -- C++ host-side mirror of the reference's per-step interface, on top of the C ABI
// (include/compton2d_b200.h).  Header only; link with -lcompton2d_b200.
//
// The reference's "API" is the sequence of argument-less subroutine calls in `xec` (src_20121026/xec2d.f:141-193)
// operating on COMMON blocks:  imcgen2d -> imcfield2d -> imcvol2d -> imcsurf2d -> imcredist -> update -> graphics_collect.
// c2d::Compton2d keeps those names and that order, with the COMMON members passed as dense arrays ([zone][bin],
// zone = (k-1) + nr (j-1), fp_mpi.f:333-352).  Nothing here computes physics: every method is one or two calls into the
// library.  Errors: the reference prints a line and STOPs (imctrk2d.f:583-584, imcfield2d.f:85-86); here every non-zero
// return of the C ABI becomes a c2d::Error carrying the code and c2d_last_error() -- a host that wants the reference's
// behaviour catches it at top level, prints what() and exits.
// The Fortran binding of the same entry points is INTEGRATION.md / integration/c2d_shim.f90; the Python one is
// compton2d_b200/host.py (what the tests and bench.py drive).
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "compton2d_b200.h"

namespace c2d {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &msg) : std::runtime_error("compton2d_b200 error " + std::to_string(c) + ": " + msg), code(c) {}
};
inline void check(int rc) { if (rc != 0) throw Error(rc, c2d_last_error()); }

// static grids: what setup (setup2d.f:48-222) leaves in COMMON and setup_bcast ships (fp_mpi.f:12-319)
struct Grid {
    int nz = 0, nr = 0;
    double rmin = 0.0, zmin = 0.0;
    std::vector<double> z, r, E_ph, E_field, E_gg, gnt, hu, Elcmin, Elcmax, mu;
    std::vector<double> E_ref, P_ref, W_abs;      // Compton-reflection tables of ref_matrix.f (COMMON order), empty = none
};
struct Options {
    int pair_switch = 0, cr_sent = 0, spec_switch = 0;
    int split1 = 1000, split2 = 1000, split3 = 300, spl3_trg = 10;       // reader.f:587-590
    uint32_t rseed = 9857;
    int64_t census_capacity = 5000000, event_capacity = 1000000, collision_capacity = 65536;
    int sigma_mode = 0, device = 0;
    int rank = 0, world = 1;                                              // myid / numprocs of the reference
};

// dense copies of the step's tallies
struct Tallies {
    std::vector<double> edep, prdep, ecens, n_field, n_ph, E_IC, fout, edout, erlki, erlko, erlku, erlkl, Ed_in, Ed_ref;
    std::vector<double> fout_total;                 // escaping spectrum summed over the run and the ranks (xec2d.f:389-394)
    std::vector<int64_t> npcen, nelectron;
    int64_t counters[16] = {0}, counters_global[16] = {0};
};

// == Pref_calc (ref_matrix.f:7-89) on device `device`: E_ref [n_ref], P_ref(n_out, n_in) with column n_in contiguous
inline void pref_calc(int device, int n_ref, std::vector<double> &E_ref, std::vector<double> &P_ref)
{
    c2d_ctx *h = nullptr;
    check(c2d_create(&h, device));
    E_ref.assign((size_t)n_ref, 0.0); P_ref.assign((size_t)n_ref * n_ref, 0.0);
    const int rc = c2d_pref_calc(h, n_ref, E_ref.data(), P_ref.data());
    c2d_destroy(h);
    check(rc);
}

class Compton2d {
  public:
    // Ranks: one Compton2d per GPU.  comm_init joins them: rank 0 makes the NCCL id (unique_id()), the host ships its 128 bytes
    // with whatever it has (MPI_Bcast in the reference's setting), every rank passes it here.  All later collectives -- the
    // tally all-reduce, the FP all-gather, the re-dealing of the scattering cascade -- run inside the library.
    static std::vector<char> unique_id() { std::vector<char> id(C2D_COMM_ID_BYTES); check(c2d_comm_unique_id(id.data())); return id; }
    void comm_init(const std::vector<char> &id) { check(c2d_comm_init(h_, id.data(), opt_.rank, opt_.world)); }

    Compton2d(const Grid &g, const Options &o) : grid_(g), opt_(o), nzone_(g.nz * g.nr)
    {
        c2d_dims d{};
        d.nz = g.nz; d.nr = g.nr;
        d.n_vol = (int)g.E_ph.size(); d.num_nt = (int)g.gnt.size(); d.nphfield = (int)g.E_field.size(); d.n_gg = (int)g.E_gg.size();
        d.nphtotal = (int)g.hu.size() - 1; d.nph_lc = (int)g.Elcmin.size(); d.nmu = (int)g.mu.size();
        d.pair_switch = o.pair_switch; d.cr_sent = o.cr_sent; d.spec_switch = o.spec_switch;
        d.split1 = o.split1; d.split2 = o.split2; d.split3 = o.split3; d.spl3_trg = o.spl3_trg;
        d.census_capacity = o.census_capacity; d.event_capacity = o.event_capacity; d.collision_capacity = o.collision_capacity;
        d.rseed = o.rseed; d.sigma_mode = o.sigma_mode; d.rmin = g.rmin; d.zmin = g.zmin;
        d.z = ptr(grid_.z); d.r = ptr(grid_.r); d.E_ph = ptr(grid_.E_ph); d.E_field = ptr(grid_.E_field); d.E_gg = ptr(grid_.E_gg);
        d.gnt = ptr(grid_.gnt); d.hu = ptr(grid_.hu); d.Elcmin = ptr(grid_.Elcmin); d.Elcmax = ptr(grid_.Elcmax); d.mu = ptr(grid_.mu);
        if (!grid_.E_ref.empty()) { d.n_ref = (int)grid_.E_ref.size(); d.E_ref = ptr(grid_.E_ref); d.P_ref = ptr(grid_.P_ref); d.W_abs = ptr(grid_.W_abs); }
        check(c2d_create(&h_, o.device));
        const int rc = c2d_init(h_, &d);
        if (rc != 0) { const std::string msg = c2d_last_error(); c2d_destroy(h_); h_ = nullptr; throw Error(rc, msg); }
    }
    ~Compton2d() { if (h_) c2d_destroy(h_); }
    Compton2d(const Compton2d &) = delete;
    Compton2d &operator=(const Compton2d &) = delete;

    int nzone() const { return nzone_; }
    c2d_ctx *handle() const { return h_; }

    // ---- one time step, in the order of xec2d.f:141-193 ------------------------------------------------------------
    // what imcgen2d leaves in COMMON for the workers (vol_bcast, the volume job messages, z_surf_bcast, xec_bcast)
    void imcgen2d(const c2d_step_in &s)
    {
        check(c2d_step_begin(h_, &s));
        n_volume_ = 0;
        if (s.nsv) for (int i = 0; i < nzone_; i++) n_volume_ += s.nsv[i];
    }
    // ... or imcgen2d itself for the volume sources, volume_em on the device (volume2d.f:10-394)
    c2d_gen_out imcgen2d_device(const c2d_gen_in &g, std::vector<int32_t> *nsv = nullptr, std::vector<double> *ewsv = nullptr)
    {
        c2d_gen_out out{};
        if (nsv) { nsv->assign(nzone_, 0); out.nsv = nsv->data(); }
        if (ewsv) { ewsv->assign(nzone_, 0.0); out.ewsv = ewsv->data(); }
        check(c2d_step_begin_gen(h_, &g, &out));
        n_volume_ = out.n_volume;
        return out;
    }
    void imcfield2d() { check(c2d_imcfield(h_)); }                                       // field_calc, imcfield2d.f:57-148
    void imcvol2d() { check(c2d_imcvol(h_, opt_.rank, -1, opt_.world)); }               // sum of vol_calc, this rank's share
    void imcsurf2d(const c2d_surf_in *s = nullptr)                                       // z_surf_calc + r_surf_calc
    {
        if (!s) return;                                                                  // no surface sources (shipped input)
        check(c2d_surf_begin(h_, s));
        check(c2d_imcsurf(h_, opt_.rank, -1, opt_.world));
    }
    void imcredist() {}                       // imcredist.f: histories are sharded statically, census photons never move
    void sync() { check(c2d_sync(h_)); }
    // cens_add_up + E_add_up + xec_add + graphics_collect (update2d.f:1931-2083, xec2d.f:325-394): c2d_tally_reduce -- the
    // cascade is drained, the two packed tally buffers are all-reduced in place over the ranks of comm_init
    void graphics_collect() { check(c2d_tally_reduce(h_, nullptr)); }

    // ---- the closed step on a device-resident plasma (c2d_step): what the whole body of xec's loop does ----------
    void plasma_set(const c2d_plasma &p) { check(c2d_plasma_set(h_, &p)); }
    void plasma_get(c2d_plasma_out &p) { check(c2d_plasma_get(h_, &p)); }
    c2d_step_report step(const c2d_run &run, double dt)
    {
        c2d_step_report rep{};
        check(c2d_step(h_, &run, ncycle_, time_, dt, &rep));
        if (ncycle_ > 0) time_ += dt;                                                    // xec2d.f:94-103
        ncycle_++;
        return rep;
    }
    int ncycle() const { return ncycle_; }
    double time() const { return time_; }
    // update minus MPI (update2d.f:7-334): FP_calc for the zones rank, rank + world, ...; the caller all-gathers the outputs
    void update(c2d_fp_in in, c2d_fp_out &out)
    {
        in.first_zone = opt_.rank; in.n_zones = -1; in.zone_stride = opt_.world;
        check(c2d_update_fp(h_, &in, &out));
    }
    std::vector<double> IC_loss()
    {
        std::vector<double> F(grid_.gnt.size() * grid_.E_field.size());
        check(c2d_ic_loss(h_, F.data()));
        return F;
    }

    // ---- results ------------------------------------------------------------------------------------------------
    Tallies tallies()
    {
        const size_t nzn = (size_t)nzone_, nf = grid_.E_field.size(), ng = grid_.E_gg.size(), N = grid_.gnt.size(), nph = grid_.hu.size() - 1;
        Tallies t;
        t.edep.assign(nzn, 0); t.prdep.assign(nzn, 0); t.ecens.assign(nzn, 0); t.npcen.assign(nzn, 0);
        t.n_field.assign(nzn * nf, 0); t.n_ph.assign(nzn * ng, 0); t.E_IC.assign(N, 0); t.nelectron.assign(N + 2, 0);
        t.fout.assign((size_t)C2D_NMUMAX * nph, 0); t.edout.assign((size_t)C2D_NMUMAX * C2D_NLCMAX, 0);
        t.erlki.assign(grid_.nz, 0); t.erlko.assign(grid_.nz, 0); t.erlku.assign(grid_.nr, 0); t.erlkl.assign(grid_.nr, 0);
        t.Ed_in.assign(grid_.nr, 0); t.Ed_ref.assign(grid_.nr, 0); t.fout_total.assign((size_t)C2D_NMUMAX * nph, 0);
        c2d_tallies c{};
        c.fout_total = t.fout_total.data();
        c.edep = t.edep.data(); c.prdep = t.prdep.data(); c.ecens = t.ecens.data(); c.npcen = t.npcen.data();
        c.n_field = t.n_field.data(); c.n_ph = t.n_ph.data(); c.E_IC = t.E_IC.data(); c.nelectron = t.nelectron.data();
        c.fout = t.fout.data(); c.edout = t.edout.data(); c.erlki = t.erlki.data(); c.erlko = t.erlko.data();
        c.erlku = t.erlku.data(); c.erlkl = t.erlkl.data(); c.Ed_in = t.Ed_in.data(); c.Ed_ref = t.Ed_ref.data();
        check(c2d_tally_read(h_, &c));
        for (int k = 0; k < 16; k++) { t.counters[k] = c.counters[k]; t.counters_global[k] = c.counters_global[k]; }
        return t;
    }
    int64_t census_count() { int64_t n = 0; check(c2d_census_count(h_, &n)); return n; }
    // record layout of imctrk2d.f:566-580: dbuf[6n] = rpre zpre wmu phi ew xnu, ibuf[6n] = jgpsp jgplc jgpmu jph kph seed
    void census(std::vector<double> &dbuf, std::vector<int32_t> &ibuf, std::vector<uint64_t> *sid = nullptr)
    {
        const int64_t n = census_count();
        dbuf.assign((size_t)(6 * n + 6), 0.0); ibuf.assign((size_t)(6 * n + 6), 0);
        if (sid) sid->assign((size_t)n + 1, 0);
        int64_t got = 0;
        check(c2d_census_export(h_, dbuf.data(), ibuf.data(), sid ? sid->data() : nullptr, n, &got));
        dbuf.resize((size_t)(6 * got)); ibuf.resize((size_t)(6 * got));
        if (sid) sid->resize((size_t)got);
    }
    void census_import(const std::vector<double> &dbuf, const std::vector<int32_t> &ibuf, const std::vector<uint64_t> *sid = nullptr)
    {
        check(c2d_census_import(h_, dbuf.data(), ibuf.data(), sid ? sid->data() : nullptr, (int64_t)(dbuf.size() / 6)));
    }
    std::vector<c2d_event> events()                                // the 7 doubles of imcleak2d.f:171,181 per escaping photon
    {
        std::vector<c2d_event> ev((size_t)opt_.event_capacity);
        int64_t n = 0;
        check(c2d_events_drain(h_, ev.data(), opt_.event_capacity, &n));
        ev.resize((size_t)(n < opt_.event_capacity ? n : opt_.event_capacity));
        return ev;
    }
    // reference file formats (census2d.f:1-76, imcleak2d.f:181)
    int64_t write_cens(const std::string &path) { int64_t n = 0; check(c2d_census_write(h_, path.c_str(), &n)); return n; }
    void read_cens(const std::string &path, int64_t entries) { check(c2d_census_read(h_, path.c_str(), entries)); }
    int64_t write_events(const std::string &path, bool append = true)
    {
        int64_t n = 0;
        check(c2d_events_write(h_, path.c_str(), append ? 1 : 0, &n));
        return n;
    }
    // event-list reduction with plcm's / pspt's bin rules (postprocessing/plcm.c:384-450, pspt.c:249-289)
    void lc_select(int slot) { check(c2d_lc_select(h_, slot)); }           // up to 4 reducers per context, all fed by lc_accumulate
    void lc_begin(const c2d_lc_config &cfg) { check(c2d_lc_begin(h_, &cfg)); lc_cells_ = (size_t)cfg.n_t * cfg.n_mu * cfg.n_bands; }
    void lc_accumulate() { check(c2d_lc_accumulate(h_)); }
    void lc_read(std::vector<double> &sum_ew, std::vector<double> &sum_ew2, std::vector<int64_t> &count)
    {
        sum_ew.assign(lc_cells_, 0); sum_ew2.assign(lc_cells_, 0); count.assign(lc_cells_, 0);
        check(c2d_lc_read(h_, sum_ew.data(), sum_ew2.data(), count.data()));
    }

  private:
    static const double *ptr(const std::vector<double> &v) { return v.empty() ? nullptr : v.data(); }
    Grid grid_;
    Options opt_;
    int nzone_;
    int64_t n_volume_ = 0;
    size_t lc_cells_ = 0;
    int ncycle_ = 0;
    double time_ = 0.0;
    c2d_ctx *h_ = nullptr;
};

}  // namespace c2d
