"""Generate tests/golden/refbin_*.npz from the UNMODIFIED reference binary.

Runs ONLY in the authoring container (needs /root/reference).  Two passes of
src_20121026/compton2d under oracle/refbin/fakempi.c with one worker rank:
  pass 1  records the worker's MPI call trace,
  pass 2  re-runs (the run is deterministic with one worker) and
          dumps the worker's COMMON blocks -- full binary precision -- at the
          phase boundaries of every cycle:
            A  before field_calc          (imcfield2d.f:38   field_recv_synch)
            B  after  field_calc          (imcfield2d.f:42   field_send)
            C  after  all vol_calc jobs   (surf_mpi.f:18     first z_surf_bcast)
            D  after  surface sources     (imcredist.f       barrier)
            E  before the FP farm         (fp_mpi.f:596      FP_bcast)
            F  after  the FP farm         (update2d.f:2023   first E_add_up reduce)
The snapshots hold, for the used nz x nr extent, every input the worker's hot
path reads and every tally it writes, plus ran1's SAVE state, so the CPU
restatement (oracle/c2d_oracle.c) can be replayed phase by phase against them.

usage: python oracle/refbin/make_golden.py <name>      (see CONFIGS)
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refrun  # noqa: E402

REPO = Path(__file__).resolve().parent.parent.parent
GOLD = REPO / "tests" / "golden"
NCENS = 16000     # census photons dumped per snapshot (runs must stay below)

CONFIGS = {
    # shipped physics on a 4x3 grid, ~3000 sources/step, 3 cycles (0,1,2)
    "small": dict(over=dict(nz=4, nr=3, nst=6000, tstop="1.e+3"), zone={}),
    # dense blob: Compton scattering actually happens, small split factors so the
    # 2nd and 3rd splitting and 'multiple scattering' are all exercised
    "dense": dict(over=dict(nz=3, nr=3, nst=1200, tstop="1.e+3", split1=60, split2=40,
                            split3=12), zone=dict(n_e="0.300e06")),
    # rmin > 0 (inner boundary absorbs), nmu = 1, different seed
    "annulus": dict(over=dict(nz=3, nr=4, nst=5000, tstop="1.e+3", rmin="0.2d16", rseed=4242),
                    zone={}),
    # specular reflection at the lower surface: cr_sent = 1 with the shipped cold boundary (tbbl = -0.d0),
    # imcleak2d.f:107-122; denser plasma and small splits so that scattered photons bounce too
    "mirror": dict(over=dict(nz=3, nr=3, nst=1500, tstop="1.e+3", cr_sent=1, split1=60, split2=40, split3=12, rseed=777),
                   zone=dict(n_e="0.300e06")),
    # Compton reflection off a hot lower boundary and off the outer cylinder (cr_sent = 3, tbbl = 0.5 keV: imcleak2d.f:124-165,
    # 216-273) with the Green's-function tables of ref_matrix.f, which are dumped too (extra=)
    # (t1(1) = 1e20 keeps the time-interval index ti at 1: with the shipped t1(1) = 1 s the workers read tbbl(k, 2) = 0 from
    # cycle 2 on and fall back to the mirror, imcsurf2d_para.f:55-63)
    "reflect": dict(over=dict(nz=3, nr=3, nst=1500, tstop="1.e+3", cr_sent=3, split1=60, split2=40, split3=12, rseed=4711,
                              **{"t0(1)": "0.d0", "t1(1)": "1.d20"}),
                    zone=dict(n_e="0.300e06"), wi=dict(tbbl="0.5d0"),
                    extra=[("p_ref", (500, 500)), ("w_abs", (500, 500)), ("e_ref", (500,))]),
    # BASELINE configs[0]: src_20121026/input.dat + inputm.dat AS SHIPPED (30 x 9 zones, nst = 8e5, split 1000/1000/300,
    # pair_switch = 1, T_const = 0), cycles 0-2 (tstop = 1e3 as in SURVEY.md 8 "golden values").  ~20 min per pass on one
    # core; only the first 2000 census records of each snapshot are kept, and the first 4000 event records.
    "c0": dict(over=dict(tstop="1.e+3"), zone={}, ncens=2000, census_head=True, events_head=4000, keep_logs=True),
    # rand_switch = 1: Knuth's lagged-Fibonacci generator, re-seeded per zone job and per census photon (rand.f:9-316);
    # the generator state (COMMON /RSTATE/ RANX, /rlist/, /rcount/) and the seeds are dumped too
    "knuth": dict(over=dict(nz=4, nr=3, nst=6000, tstop="1.e+3", rand_switch=1), zone={}, knuth=True,
                  extra=[("seeds", (4, 3)), ("zseeds", (4,)), ("rseeds", (3,)), ("randlist", (10000,)), ("randcounter", None)]),
    # external-Compton seed photons (BASELINE configs[2]): upper and lower disk sources sampled from a spectrum
    # file (the only live surface-source path of the reference: nsurf* > 0 requires tbb* < 0, imcgen2d.f:441-442)
    "surface": dict(over=dict(nz=3, nr=4, nst=1200, tstop="1.e+3", **{"t0(1)": "0.d0", "t1(1)": "1.d20"}), zone={},
                    wi=dict(tbbl="-1.d0", l_fname="blackbody_G15_4spectra.in", tbbu="-1.d0",
                            u_fname="blackbody_G15_4spectra.in"),
                    files=["/root/reference/disk/blackbody_G15_4spectra.in"]),
}

# (name, dims-to-keep as function of (nz, nr, nmu)) -- None = scalar / full
def varlist(nz, nr, ncens=NCENS):
    Z = lambda *lead: (lambda: tuple(lead) + (nz, nr))
    v = [
        ("ndxout", None), ("nz", None), ("nr", None), ("ncycle", None), ("pair_switch", None),
        ("npcen", (nz, nr)), ("nsv", (nz, nr)), ("rseed", None), ("rand_switch", None),
        ("incounter", None), ("outcounter", None), ("cmcount1", None), ("cmcount2", None), ("lkcount", None),
        ("nmu", None), ("nelectron", (200,)), ("nphtotal", None), ("nph_lc", None),
        ("cr_sent", None), ("upper_sent", None), ("spec_switch", None), ("ti", None),
        ("split1", None), ("split2", None), ("split3", None), ("spl3_trg", None),
        ("t_const", None), ("inj_switch", None), ("inj_dis", None), ("g2var_switch", None), ("pick_sw", None),
        ("cf_sentinel", None), ("nst", None),
        ("ibufout", (6 * ncens,)), ("dbufout", (6 * ncens,)),
        ("amxwl", (nz, nr)), ("gmin", (nz, nr)), ("gmax", (nz, nr)), ("p_nth", (nz, nr)),
        ("q_turb", (nz, nr)), ("turb_lev", (nz, nr)), ("ec_old", (nz, nr)),
        ("n_pos", (nz, nr, 200)), ("dn_pp", (nz, nr, 200)), ("dne_pa", (nz, nr, 200)), ("dnp_pa", (nz, nr, 200)),
        ("erlki", (nz,)), ("erlko", (nz,)), ("erlku", (nr,)), ("erlkl", (nr,)),
        ("ecens", (nz, nr)), ("edep", (nz, nr)), ("prdep", (nz, nr)),
        ("kappa_tot", (400, nz, nr)), ("eps_tot", (400, nz, nr)), ("eps_th", (400, nz, nr)), ("e_ph", (400,)),
        ("eloss_tot", (nz, nr)), ("eloss_br", (nz, nr)), ("eloss_cy", (nz, nr)), ("eloss_sy", (nz, nr)),
        ("eloss_th", (nz, nr)),
        ("tea", (nz, nr)), ("tna", (nz, nr)), ("b_field", (nz, nr)), ("n_e", (nz, nr)),
        ("z", (nz,)), ("r", (nr,)), ("rmin", None), ("zmin", None), ("vol", (nz, nr)), ("zsurf", (nz, nr)),
        ("time", None), ("dt", (2,)), ("tstop", None), ("mcdt", None), ("t0", (1,)), ("t1", (1,)),
        ("f_pair", (nz, nr)), ("pnt", (nz, nr, 200)), ("gnt", (200,)), ("f_nt", (nz, nr, 200)),
        ("e_ic", (200,)), ("n_field", (400, nz, nr)), ("e_field", (400,)),
        ("n_ph", (100, nz, nr)), ("k_gg", (100, nz, nr)), ("e_gg", (100,)),
        ("mu", (32,)), ("ed_abs", (nr,)), ("ed_ref", (nr,)), ("ed_in", (nr,)),
        ("hu", (129,)), ("elcmin", (10,)), ("elcmax", (10,)),
        ("tbbu", (nr, 1)), ("tbbl", (nr, 1)), ("tbbi", (nz, 1)), ("tbbo", (nz, 1)),
        ("nsurfu", (nr,)), ("nsurfl", (nr,)), ("nsurfi", (nz,)), ("nsurfo", (nz,)),
        ("ewsurfu", (nr,)), ("ewsurfl", (nr,)), ("ewsurfi", (nz,)), ("ewsurfo", (nz,)), ("star_switch", None),
        ("nfile", None), ("e_file", (500,)), ("p_file", (500,)), ("a1", (500,)), ("i_file", (500,)), ("f_file", (500,)),
        ("int_file", None), ("r_blr", None), ("fr_blr", None), ("r_ir", None), ("fr_ir", None), ("r_disk", None),
        ("d_jet", None),
        ("inj_g1", None), ("inj_g2", None), ("inj_p", None), ("inj_t", None), ("inj_gg", None),
        ("inj_sigma", None), ("inj_l", None), ("inj_v", None), ("g_bulk", None), ("pick_rate", None),
        ("r_esc", None), ("r_acc", None), ("f_ic", (200, 400)),
        ("te_new", (nz, nr)), ("dt_new", None), ("lnl", None), ("dt_max", None),
        ("e_tot_old", None), ("e_tot_new", None), ("hr_total", None), ("hr_st_total", None),
        ("f_t_implicit", None), ("dr", None), ("dz", None),
        ("ewsv", (nz, nr)), ("edout", (1, 10)), ("fout", (1, 128)), ("t_bound", None),
        ("r_flare", None), ("z_flare", None), ("t_flare", None), ("sigma_r", None), ("sigma_z", None),
        ("sigma_t", None), ("flare_amp", None),
    ]
    return v


# nothing is treated as static: several "constants" reach the worker only with the first
# z_surf_bcast / FP_bcast, so their cycle-0 values differ; de-duplication below
# makes unchanged arrays free.
STATIC = set()


def snapshot_points(trace):
    """Worker call indices of the phase boundaries, per cycle."""
    pts = []
    cyc = -1
    i = 0
    n = len(trace)
    state = None
    while i < n:
        t = trace[i]
        sig = (t["op"], t["count"], t["dtype"])
        # A: the 1-int receive that opens imcfield2d (followed by a 1-double send)
        if sig == ("recv", 1, 7) and i + 1 < n and (trace[i + 1]["op"], trace[i + 1]["count"], trace[i + 1]["dtype"]) == ("send", 1, 17):
            cyc += 1
            pts.append((t["idx"], f"c{cyc}_A"))
            pts.append((trace[i + 1]["idx"], f"c{cyc}_B"))
            state = "vol"
        # C: first z_surf_bcast after the vol end signal (recv of derived type with tag 9802)
        elif state == "vol" and t["op"] == "recv" and t["dtype"] >= 1000 and t.get("rtag") == 9802:
            pts.append((trace[i + 1]["idx"], f"c{cyc}_C"))
            state = "surf"
        elif state == "surf" and t["op"] == "barrier":
            pts.append((t["idx"], f"c{cyc}_D"))
            state = "redist"
        elif state == "redist" and t["op"] == "bcast" and t["dtype"] >= 1000:
            pts.append((t["idx"], f"c{cyc}_E"))
            state = "fp"
        elif state == "fp" and t["op"] == "reduce" and t["dtype"] == 17:      # E_add_up's first reduce: erlki, count nz (update2d.f:2023)
            pts.append((t["idx"], f"c{cyc}_F"))
            state = None
        i += 1
    return pts


def main(name):
    cfg = CONFIGS[name]
    L = refrun.CommonLayout()
    rd = Path("/tmp") / f"c2d_gold_{name}"
    shutil.rmtree(rd, ignore_errors=True)
    rd.mkdir()
    vals = refrun.write_inputs(rd, cfg["over"], cfg["zone"], **cfg.get("wi", {}))
    for f in cfg.get("files", []):
        shutil.copy(f, rd / Path(f).name)
    nz, nr = int(vals["nz"]), int(vals["nr"])
    r = refrun.run(rd)
    assert r.returncode == 0, (rd / "stdout.txt").read_text()[-2000:]
    trace = refrun.read_trace(rd, 1)
    pts = snapshot_points(trace)
    print("snapshot points:", pts)

    ncens = int(cfg.get("ncens", NCENS))
    vl = varlist(nz, nr, ncens) + list(cfg.get("extra", []))
    spec = []
    layout = []
    for vname, keep in vl:
        if vname in ("incounter", "outcounter", "cmcount1", "cmcount2", "lkcount", "split1", "split2",
                     "split3", "spl3_trg", "inj_switch", "inj_dis", "g2var_switch", "pick_sw", "nz", "nr",
                     "nphtotal", "nph_lc", "cr_sent", "upper_sent", "rmin", "zmin", "time", "tstop", "mcdt",
                     "dt_new", "lnl", "dt_max", "e_tot_old", "e_tot_new", "hr_total", "hr_st_total",
                     "f_t_implicit", "dr", "dz", "inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma",
                     "inj_l", "inj_v", "g_bulk", "pick_rate", "r_esc", "r_acc", "r_flare", "z_flare",
                     "t_flare", "sigma_r", "sigma_z", "sigma_t", "flare_amp", "rseed", "rand_switch"):
            pass
        rngs, used = L.ranges(vname, list(keep) if keep else None)
        es = L.where[vname][2]
        nbytes = sum(x[2] for x in rngs)
        layout.append((vname, es, used, nbytes))
        spec += [f"R {s} {o} {n}" for s, o, n in rngs]
    # ran1 SAVE state (rand.f:339-343): iv(32), iy
    spec += ["R ran1_$IV.0.1 0 128", "R ran1_$IY.0.1 0 4"]
    layout += [("ran1_iv", 4, (32,), 128), ("ran1_iy", 4, (), 4)]
    if cfg.get("knuth"):
        spec += ["R rstate_ 0 800"]
        layout += [("knuth_ranx", 8, (100,), 800)]
    specf = rd / "snapspec.txt"
    specf.write_text("".join(f"AT 1 {i} {lab}\n" for i, lab in pts) + "\n".join(spec) + "\n")

    # pass 2 (fresh directory state: rates/ cache and event files must not carry over)
    evt1 = (rd / "p001_evb.dat").read_text()
    for d in ("output", "temp", "rates"):
        shutil.rmtree(rd / d); (rd / d).mkdir()
    for f in rd.glob("*.dat"):
        f.unlink()
    r = refrun.run(rd, snapspec=specf)
    assert r.returncode == 0
    assert (rd / "p001_evb.dat").read_text() == evt1, "reference run is not reproducible?!"

    out = {}
    for idx, lab in pts:
        raw = (rd / f"{lab}.bin").read_bytes()
        off = 0
        for vname, es, used, nbytes in layout:
            is_int = vname in L.where and L.vars[vname][0] == 4 and vname not in ()
            dt = np.int32 if (vname.startswith("ran1") or (vname in L.where and es == 4)) else np.float64
            a = np.frombuffer(raw[off: off + nbytes], dtype=dt).copy()
            off += nbytes
            if used:
                a = a.reshape(tuple(reversed(used))).transpose()   # Fortran order -> a[j-1,k-1,...]
            else:
                a = a.reshape(())
            if vname in STATIC:
                if lab.endswith("_C") and lab.startswith("c0"):
                    pass
                key = f"static/{vname}"
                # statics are taken from the last snapshot where everything has arrived
                out[key] = a
            else:
                out[f"{lab}/{vname}"] = a
        assert off == len(raw)
    # trim census dumps to the live count
    for idx, lab in pts:
        n = int(out[f"{lab}/ndxout"])
        if cfg.get("census_head"):
            n = min(n, ncens)                # big runs: only the head of the census is kept (the count is in ndxout)
        assert n <= ncens, f"census {n} exceeds dump cap"
        out[f"{lab}/dbufout"] = out[f"{lab}/dbufout"][: 6 * n]
        out[f"{lab}/ibufout"] = out[f"{lab}/ibufout"][: 6 * n]
    # de-duplicate: a variable is stored under a label only if it differs from the value at
    # the previous label (tests/golden_util.py walks back to the last stored value)
    labs = [lab for _, lab in pts]
    last = {}
    for lab in labs:
        for k in [k for k in list(out) if k.startswith(lab + "/")]:
            v = k.split("/", 1)[1]
            if v in last and last[v].shape == out[k].shape and np.array_equal(last[v], out[k]):
                del out[k]
            else:
                last[v] = out[k]
    # event file of the worker (text, 7 doubles per record; imcleak2d.f:181)
    ev = np.loadtxt(rd / "p001_evb.dat", ndmin=2) if evt1.strip() else np.zeros((0, 7))
    if cfg.get("events_head"):
        # big runs: head of the list + the sums the parity gates use (SURVEY.md 8 notes)
        out["events_count"] = np.array(len(ev)); out["events_sum_ew"] = np.array(ev[:, 2].sum())
        out["events_t_hist"] = np.histogram(ev[:, 0], bins=64)[0]; out["events_t_edges"] = np.histogram(ev[:, 0], bins=64)[1]
        out["events_lnE_hist"], out["events_lnE_edges"] = np.histogram(np.log10(ev[:, 1]), bins=np.linspace(-10, 12, 89))
        out["events_lnE_ew"] = np.histogram(np.log10(ev[:, 1]), bins=np.linspace(-10, 12, 89), weights=ev[:, 2])[0]
        # unscattered (synchrotron) escapes: source photons carry the zone's ewsv and soft energies, Compton children are far
        # lighter or far harder (tests/test_statistical_parity.py:_is_unscattered); per cycle by t_bound
        w_med = float(np.median(ev[:, 2])); dt1 = float(out[[k for k in out if k.endswith("/dt")][-1]][0])
        heavy = (ev[:, 2] > 0.3 * w_med) & (ev[:, 2] < 3.0 * w_med) & (ev[:, 1] < 10.0)
        in_cyc = lambda c: (ev[:, 0] >= (c - 1) * dt1) & (ev[:, 0] < c * dt1)
        out["events_w_median"] = np.array(w_med); out["events_unscattered_count"] = np.array(int(heavy.sum()))
        out["events_unscattered_sum_ew"] = np.array(float(ev[heavy, 2].sum()))
        out["events_unscattered_per_cycle"] = np.array([int((heavy & in_cyc(c)).sum()) for c in (1, 2)])
        out["events_per_cycle"] = np.array([int(in_cyc(c).sum()) for c in (1, 2)])
        ev = ev[: cfg["events_head"]]
    out["events"] = ev
    if cfg.get("keep_logs"):
        # what the master wrote: log.txt (bingo, Luminosity, (Delta E)/E, dT_max), the light curve and temperature files
        for fn in ("log.txt", "output/lcb_01.dat", "output/temp_b.dat", "output/spb.dat", "output/phb.dat", "stdout.txt"):
            f = rd / fn
            if f.exists():
                txt = f.read_text()
                if len(txt) > 200000:          # log.txt: keep the per-cycle summary lines, not the per-zone chatter
                    keys = ("time step", "bingo", "Luminosity", "Delta E", "dT_max", "Total energy", "dt_new", "Total heat", "Acceleration rate")
                    txt = "".join(l + "\n" for l in txt.splitlines() if any(k in l for k in keys))
                out["file/" + fn] = np.array(txt)
    # the worker's event file as the binary formatted it (format 105 of imcleak2d.f:181), first 600 records: pins the
    # e14.7 formatter of the library's reference-format writers (c2d_events_write / c2d_census_write)
    if name == "dense":
        (GOLD / f"evb_{name}.txt").write_text("".join(evt1.splitlines(keepends=True)[:600]))
    out["labels"] = np.array([lab for _, lab in pts])
    # scalars the log prints (FP energy check etc.) stay in the log; keep stdout tail for provenance
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / f"refbin_{name}.npz", **out)
    print("wrote", GOLD / f"refbin_{name}.npz", (GOLD / f"refbin_{name}.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
