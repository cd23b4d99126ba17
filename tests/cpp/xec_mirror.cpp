// tests/cpp/xec_mirror.cpp -- drives two time steps of the hot path through the C++ host mirror (include/compton2d_b200.hpp)
// in the order of the reference's xec (xec2d.f:141-193).  Inputs: a directory of raw arrays (<name>.f64 / <name>.i32) and
// scalars.txt written by tests/test_cpp_host.py; outputs: raw arrays into the same directory (out_<name>).
// usage: xec_mirror <dir> [cr_sent]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "compton2d_b200.hpp"

static std::string g_dir;
template <class T> static std::vector<T> load(const std::string &name, const char *ext)
{
    std::ifstream f(g_dir + "/" + name + ext, std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "missing " << name << ext << "\n"; std::exit(2); }
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<T> v((size_t)n / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), n);
    return v;
}
static std::vector<double> F(const std::string &n) { return load<double>(n, ".f64"); }
template <class T> static void save(const std::string &name, const std::vector<T> &v)
{
    std::ofstream f(g_dir + "/out_" + name, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    g_dir = argv[1];
    std::map<std::string, double> sc;
    { std::ifstream f(g_dir + "/scalars.txt"); std::string k; double v; while (f >> k >> v) sc[k] = v; }
    c2d::Grid g;
    g.nz = (int)sc["nz"]; g.nr = (int)sc["nr"]; g.rmin = sc["rmin"]; g.zmin = sc["zmin"];
    g.z = F("z"); g.r = F("r"); g.E_ph = F("E_ph"); g.E_field = F("E_field"); g.E_gg = F("E_gg"); g.gnt = F("gnt"); g.hu = F("hu");
    g.Elcmin = F("Elcmin"); g.Elcmax = F("Elcmax"); g.mu = F("mu");
    c2d::Options o;
    o.split1 = (int)sc["split1"]; o.split2 = (int)sc["split2"]; o.split3 = (int)sc["split3"]; o.spl3_trg = (int)sc["spl3_trg"];
    o.pair_switch = (int)sc["pair_switch"]; o.spec_switch = (int)sc["spec_switch"]; o.rseed = (uint32_t)sc["rseed"];
    o.census_capacity = 400000; o.event_capacity = 400000;
    const bool closed = argc > 2 && std::string(argv[2]) == "closed";
    o.cr_sent = (argc > 2 && !closed) ? std::atoi(argv[2]) : 0;
    try {
        c2d::Compton2d x(g, o);
        if (closed) {
            // the closed time step on the resident plasma: xec's loop body as ONE library call per cycle (c2d_step)
            const auto f_nt = F("f_nt"), Pnt = F("Pnt"), n_e = F("n_e"), tea = F("tea"), tna = F("tna"), B = F("B_field"), gmin = F("gmin"),
                       gmax = F("gmax"), p_nth = F("p_nth"), amxwl = F("amxwl"), turb = F("turb_lev"), vol = F("vol"), zsurf = F("zsurf");
            c2d_plasma pl{};
            pl.f_nt = f_nt.data(); pl.Pnt = Pnt.data(); pl.n_e = n_e.data(); pl.tea = tea.data(); pl.tna = tna.data(); pl.B_field = B.data();
            pl.gmin = gmin.data(); pl.gmax = gmax.data(); pl.p_nth = p_nth.data(); pl.amxwl = amxwl.data(); pl.turb_lev = turb.data();
            pl.vol = vol.data(); pl.zsurf = zsurf.data();
            x.plasma_set(pl);
            c2d_run run{};
            run.nst = (int64_t)sc["nst"]; run.T_const = (int)sc["T_const"]; run.quirk_cycle0 = 1; run.r_esc = 0.3; run.r_acc = 1.0;
            std::vector<int64_t> ndx; std::vector<double> bingo;
            for (int c = 0; c < 3; c++) {
                const c2d_step_report rep = x.step(run, sc["dt"]);
                ndx.push_back(rep.ndxout); ndx.push_back(rep.nevents); ndx.push_back(rep.n_volume); ndx.push_back(rep.histories);
                bingo.push_back(rep.bingo); bingo.push_back(rep.luminosity);
            }
            const c2d::Tallies t = x.tallies();
            save("closed_counts", ndx); save("closed_bingo", bingo); save("closed_ecens", t.ecens); save("closed_fout_total", t.fout_total);
            std::printf("xec_mirror closed: ncycle %d, time %.17g\n", x.ncycle(), x.time());
            return 0;
        }
        const auto kappa = F("kappa_tot"), et = F("eps_tot"), eh = F("eps_th"), f_nt = F("f_nt"), Pnt = F("Pnt"), n_e = F("n_e"), tea = F("tea"),
                   f_pair = F("f_pair"), k_gg = F("k_gg"), ewsv = F("ewsv"), elth = F("Eloss_th"), eltot = F("Eloss_tot"), zsurf = F("zsurf"),
                   tbbl = F("tbbl");
        const auto nsv = load<int32_t>("nsv", ".i32");
        c2d::Tallies t;
        for (int ncycle = 1; ncycle <= 2; ncycle++) {                     // xec2d.f:141-193
            c2d_step_in s{};
            s.ncycle = ncycle; s.time = sc["time"]; s.dt = sc["dt"];
            s.kappa_tot = kappa.data(); s.eps_tot = et.data(); s.eps_th = eh.data(); s.f_nt = f_nt.data(); s.Pnt = Pnt.data();
            s.n_e = n_e.data(); s.tea = tea.data(); s.f_pair = f_pair.data(); s.k_gg = k_gg.data(); s.nsv = nsv.data(); s.ewsv = ewsv.data();
            s.Eloss_th = elth.data(); s.Eloss_tot = eltot.data(); s.zsurf = zsurf.data(); s.tbbl = tbbl.data();
            x.imcgen2d(s);
            if (ncycle > 1) x.imcfield2d();
            x.imcvol2d();
            x.imcsurf2d();
            x.imcredist();
            x.graphics_collect();
            t = x.tallies();
        }
        save("edep", t.edep); save("prdep", t.prdep); save("ecens", t.ecens); save("npcen", t.npcen); save("n_field", t.n_field);
        save("E_IC", t.E_IC); save("nelectron", t.nelectron); save("fout", t.fout); save("erlku", t.erlku);
        save("counters", std::vector<int64_t>(t.counters, t.counters + 16));
        std::vector<double> d; std::vector<int32_t> iv; std::vector<uint64_t> sid;
        x.census(d, iv, &sid);
        save("dbuf", d); save("ibuf", iv); save("sid", sid);
        const auto ev = x.events();
        std::vector<double> evd(ev.size() * 7);
        for (size_t i = 0; i < ev.size(); i++) { const double *r = &ev[i].t_bound; for (int k = 0; k < 7; k++) evd[7 * i + k] = r[k]; }
        save("events", evd);
        const int64_t n = x.write_cens(g_dir + "/p001_census.dat");
        std::printf("xec_mirror: 2 cycles, census %lld, events %zu, histories %lld\n", (long long)n, ev.size(), (long long)t.counters[C2D_CNT_HISTORIES]);
    } catch (const c2d::Error &e) {                                       // the reference: write(*,*) ...; stop
        std::fprintf(stderr, "%s\n", e.what());
        return 10 + e.code;
    }
    return 0;
}
