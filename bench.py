#!/usr/bin/env python
"""bench.py -- photon histories/s tracked + tallied (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through its C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU build (prebuilt binary
                                                           under oracle/refbin/fakempi.c) on the host cores

Workload (config.workload): BASELINE.json configs[1] -- Mrk 421-style SSC blob, 16x16 zones, 100 photon
energy groups, 1e7 volume sources per step and GPU (nst = 2e7), split1/2/3 = 1000/1000/300, plus the census
photons the previous steps left (weak scaling: every GPU sources 1e7 histories per step).
A step = one IMC time step of the hot path: imcfield2d (census re-track) + imcvol2d (volume sources) +
scattering queue + the packed tally all-reduce.  value = histories of all ranks / max-over-ranks device time of that
region, bracketed by CUDA events on the library's stream (host gaps between the launches of the queue included).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle", ROOT / "oracle" / "refbin"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

METRIC = "photon_histories_per_s"
UNIT = "histories/s"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference (CPU) arm
def run_reference_sample(nst, ncores=1, warm=0, timed=2, workdir=None, fp=False, nz=16, nr=16):
    """Time the UNMODIFIED reference binary (oracle/_ref/compton2d, staged by __graft_entry__.build()) on a
    bounded sample of the workload: one launch of a 16x16-zone, 100-group run with `nst` particles per cycle,
    cycles 0 .. warm+timed; `ncores` independent single-worker instances in parallel (distinct rseed,
    SURVEY.md 8c item 1).  A "step" of the reference is one cycle; the timed window of a cycle is the worker's
    tracking phase (imcfield2d + imcvol2d + imcsurf2d) taken from the MPI trace.  Cycle 0 never counts: the
    worker's dt is still zero there.  -> dict(value, seconds per cycle, ...)"""
    import math
    import refrun
    from make_golden import snapshot_points
    if not (refrun.REFDIR / "compton2d").exists():
        return None
    refrun.build_shim()
    base = Path(workdir or tempfile.mkdtemp(prefix="c2d_ref_"))
    ncyc = warm + timed                                   # cycles 1..ncyc are tracked with dt > 0
    inj_v = 2.9979245620e10 * math.sqrt(1.0 - 1.0 / 33.0 ** 2)
    dt = min(0.75e16 / nr, 1.0e16 / nz) / inj_v           # setup2d.f:50-52 with the shipped z(nz), r(nr), mcdt, g_bulk
    # master continues after cycle c while (c-1)*dt < tstop (xec2d.f:109); so does the worker (xec2d.f:189)
    over = dict(nz=nz, nr=nr, nst=int(nst), tstop=f"{(ncyc - 0.5) * dt:.6e}".replace("e", "d"), pair_switch=0,
                T_const=0 if fp else 1)
    over["nphbins(1)"] = 30
    over["nphbins(2)"] = 40          # 30 + 40 + 30 = 100 groups (the shipped file reuses the key nphbins(1) for region 3)
    dirs = []
    for c in range(ncores):
        rd = base / f"inst{c}"
        shutil.rmtree(rd, ignore_errors=True)
        rd.mkdir(parents=True)
        refrun.write_inputs(rd, dict(over, rseed=9857 + 101 * c), {}, staged=refrun.REFDIR)
        dirs.append(rd)
    t0 = time.time()
    res = [None] * ncores

    def work(c):
        res[c] = refrun.run(dirs[c], trace=True, staged=refrun.REFDIR, cpu=None, timeout=3000)

    th = [threading.Thread(target=work, args=(c,)) for c in range(ncores)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.time() - t0
    hist, cyc_secs, fp_steps, fp_secs = 0, [0.0] * timed, 0, []
    for c, rd in enumerate(dirs):
        if res[c] is None or res[c].returncode != 0:
            return {"error": (rd / "stdout.txt").read_text()[-400:]}
        tr = refrun.read_trace(rd, 1)
        pts = dict((lab, idx) for idx, lab in snapshot_points(tr))
        tt = {t["idx"]: t["t"] for t in tr}
        out = (rd / "stdout.txt").read_text()
        nct = [int(l.split("=")[-1]) for l in out.splitlines() if "nctot=" in l]
        if len(nct) < ncyc + 1 or f"c{ncyc}_D" not in pts:
            return {"error": f"expected {ncyc + 1} cycles, saw {len(nct)}: " + out[-300:]}
        # cycle c tracks the census of cycle c-1 (nctot printed by imcredist) + the volume sources, whose
        # number per cycle equals cycle 0's census (cycle-0 photons are banked unflown)
        for k, cyc in enumerate(range(warm + 1, ncyc + 1)):
            hist += nct[cyc - 1] + nct[0]
            cyc_secs[k] = max(cyc_secs[k], tt[pts[f"c{cyc}_D"]] - tt[pts[f"c{cyc}_A"]])
        if fp and "c1_F" in pts:
            # the FP farm of cycle 1 (fp_mpi.f:596 FP_bcast .. first E_add_up reduce); sub-steps as the worker prints them
            fp_secs.append(tt[pts["c1_F"]] - tt[pts["c1_E"]])
            steps = [int(l.split("=")[-1]) for l in out.splitlines() if "fp_steps=" in l]
            fp_steps += sum(steps[: nz * nr]) if len(steps) >= nz * nr else 5 * nz * nr
    shutil.rmtree(base, ignore_errors=True)
    secs = sum(cyc_secs)
    r = {"histories": hist, "seconds": secs, "value": hist / secs, "wall_s": wall, "cores": ncores,
         "ms_per_step": 1e3 * secs / timed}
    if fp_secs:
        r["fp_zone_steps_per_s"] = fp_steps / max(fp_secs)
        r["fp_zone_steps"], r["fp_seconds"] = fp_steps, max(fp_secs)
    return r


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  One
    launch per core of the prebuilt binary; its cycles are the steps (warm-up cycles 1..W, timed W+1..W+K)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = min(max(1, (os.cpu_count() or 2) - 1), 64)
    nst = args.ref_nst
    r = run_reference_sample(nst, ncores=ncores, warm=args.warmup, timed=args.steps)
    if r is None or "error" in r:
        print(json.dumps({"impl": "reference", "unavailable": "reference binary not staged in oracle/_ref or run failed: "
                          + str((r or {}).get("error", ""))[:200].replace("\n", " ")}))
        return
    sample = (f"cycles {args.warmup + 1}-{args.warmup + args.steps} (census + volume sources each) of {ncores} independent "
              f"single-worker instances of the prebuilt ifort binary, 16x16 zones, 100 groups, nst={nst} per instance, "
              f"T_const=1 (FP off), split 1000/1000/300; {r['histories']} histories in {r['seconds']:.1f} s")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": dict(workload_config(args, ref=True), nst_per_gpu=nst, nst_per_instance=nst, instances=ncores,
                                                          note="bounded sample of the same workload shape (grid, groups, splits): nst is what "
                                                               "each single-worker instance of the CPU binary ran, not the GPU arm's 2e7"),
                      "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": ncores, "kind": "reference", "sample": sample},
                      "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def workload_config(args, ref=False):
    if getattr(args, "workload", "c1") == "c3":
        name = ("BASELINE configs[3]: synthetic grid 64x64 zones, 200 photon energy groups, 1.25e8 volume sources/step/GPU "
                "(nst=2.5e8; 1e9 histories/step on 8 GPUs) + census, split1/2/3=1000/1000/300, pair_switch=0")
    else:
        name = ("BASELINE configs[1]: Mrk 421-style SSC blob, 16x16 zones, 100 photon energy groups, "
                "1e7 volume sources/step/GPU (nst=2e7) + census, split1/2/3=1000/1000/300, pair_switch=0")
    return {"workload": name,
            "nz": args.nz, "nr": args.nr, "groups": args.groups, "nst_per_gpu": args.nst, "split": [1000, 1000, 300, 10],
            "sigma_mode": args.sigma_mode, "l2": "inputs larger than L2 (census bank >= 0.6 GB/step after warm-up)",
            "step": "imcfield2d + imcvol2d + scattering queue + tally all-reduce"}


# ------------------------------------------------------------------------------------------ Fokker-Planck leg
def fp_bench(local, hbm, reps=3, nz=64, nr=64, bins=200):
    """FP zone-steps/s (second half of BASELINE.json's metric) on the BASELINE configs[4] grid: 64 x 64 zones, the
    reference's 200 momentum bins (its compile-time num_nt; 512 bins exceed what the reference can run),
    every zone carrying the reference binary's own pre-FP state of the shipped plasma (tests/golden/refbin_small.npz,
    snapshot c2_E, zone (1,1): electron spectrum, tallied photon field, shipped injection block with the pick-up
    Gaussian on), photon field jittered +-10 % per zone.  One call = FP_calc for all zones (c2d_update_fp); a zone-step =
    one implicit Chang-Cooper sub-step of one zone.  bins = 512: the configs[4] momentum grid gamma - 1 = 0.2 q^(i-2),
    q = (1.1^198)^(1/510), the binary's spectrum resampled onto it (as tests/test_gpu_configs.py does)."""
    import numpy as np
    from golden_util import Golden
    from compton2d_b200.synth import context_for as make_gpu, make_problem
    E = Golden("small").snap("c2_E")
    P = make_problem(nz=nz, nr=nr, nst=1000, ngroups=100)
    nzone = nz * nr
    f_nt, pnt = np.asarray(E.f_nt[0, 0], float), np.asarray(E.pnt[0, 0], float)
    if bins != 200:
        gnt = 0.2 * ((1.1 ** 198) ** (1.0 / (bins - 2))) ** (np.arange(1, bins + 1) - 2.0)
        f = np.exp(np.interp(np.log(gnt), np.log(np.asarray(E.gnt)), np.log(np.maximum(f_nt, 1e-300))))
        f[f < 1e-250] = 0.0; f[-1] = 0.0
        f_nt = f / (f[:-1] * np.diff(gnt)).sum()
        pnt = np.concatenate(([0.0], np.cumsum(f_nt[:-1] * np.diff(gnt))))
        P["gnt"] = gnt
    g = make_gpu(P, census_capacity=1000, event_capacity=1000, device=local)
    try:
        g.timing(True)
        F_IC = g.IC_loss()
        rng = np.random.default_rng(7)
        vr = P["vol"] / float(E.vol[0, 0])
        tr = P["dt"] / float(E.dt[0])
        til = lambda a: np.ascontiguousarray(np.tile(np.asarray(a, float), (nzone, 1)))
        n_field = til(E.n_field[:, 0, 0]) * vr[:, None] * (1.0 + 0.2 * (rng.random((nzone, 1)) - 0.5))
        inj = {k: float(E[k.lower()]) for k in ("inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma", "inj_L", "inj_v", "pick_rate")}
        inj.update({k: int(E[k]) for k in ("inj_switch", "inj_dis", "g2var_switch", "pick_sw")})
        kw = dict(time=2.0 * P["dt"], dt=P["dt"], f_nt=til(f_nt), Pnt=til(pnt), n_e=np.full(nzone, float(E.n_e[0, 0])),
                  B_field=np.full(nzone, float(E.b_field[0, 0])), tea=np.full(nzone, float(E.tea[0, 0])),
                  tna=np.full(nzone, float(E.tna[0, 0])), vol=P["vol"], Eloss_sy=float(E.eloss_sy[0, 0]) * vr * tr,
                  ec_old=float(E.ec_old[0, 0]) * vr, turb_lev=np.full(nzone, float(E.turb_lev[0, 0])),
                  gmin=np.full(nzone, float(E.gmin[0, 0])), gmax=np.full(nzone, float(E.gmax[0, 0])),
                  p_nth=np.full(nzone, float(E.p_nth[0, 0])), amxwl=np.full(nzone, float(E.amxwl[0, 0])),
                  n_field=n_field, ecens=float(E.ecens[0, 0]) * vr, F_IC=F_IC, lnL=float(E.lnl), r_esc=float(E.r_esc),
                  r_acc=float(E.r_acc), inj=inj)
        out = g.update(**kw)                       # warm-up
        g.timing_reset()
        t0 = time.perf_counter()
        steps = 0
        for _ in range(reps):
            out = g.update(**kw)
            steps += int(out["fp_steps"].sum())
        wall = time.perf_counter() - t0
        ms, nl = g.kernel_time(3)
        N, nf = P["gnt"].size, P["E_field"].size
        bytes_alg = steps * 5.0 * 8.0 * N + reps * nzone * (8.0 * nf + 200.0)      # 5*8*N B per zone-step
        ach = bytes_alg / (ms * 1e-3) / 1e9
        return {"metric": "fp_zone_steps_per_s", "value": steps / (ms * 1e-3), "unit": "zone-steps/s",
                "e2e_value": steps / wall, "zones": nzone, "bins": N, "sub_steps_per_zone": steps / (reps * nzone),
                "kernel_ms_per_call": ms / reps, "launches": int(nl),
                "roofline": {"bound": "hbm", "kernel": "k_fp_zones", "achieved": ach, "peak": hbm, "unit": "GB/s",
                             "frac": ach / hbm, "traffic": 21.0e6 if bins == 200 else None,
                             "note": "latency-bound (McDonald quadratures, Thomas recurrence): HBM is not the limiter, see DESIGN.md 4"},
                "workload": f"BASELINE configs[4] grid 64x64 zones x {bins} bins" + (" (reference's num_nt)" if bins == 200 else
                            " (configs[4] momentum grid)") + ", shipped plasma + injection block"}
    finally:
        g.close()


MRK421_LC = "p001_evb.dat\n33\n1e16\nlc07_ev0.dat\n1\nlc07_ev0.dat\n0.99944\n0.99964\n7e2\n0\n7e4\n7\n" + \
    "".join(f"{a}\n{b}\n1\n0\n" for a, b in ((1e-3, 3e-3), (2, 4), (9, 15), (15, 20), (20, 60), (5e5, 5e7), (1e9, 1e10)))


def lc_bench(g, hbm, n_events=100_000_000):
    """Event-list reduction (csrc/c2d_lc.cu) on a generated list of escape records with the bin tables of
    the reference's own postprocessing/mrk421_lc.input (content reproduced above): 56 algorithmic B/record."""
    from compton2d_b200.lightcurve import lc_reduction, parse_lc_input
    cfg = parse_lc_input(MRK421_LC)
    per_rec = None
    try:
        tf = ROOT / "profiles" / "r02_traffic.json"
        per_rec = json.loads((tf if tf.exists() else ROOT / "profiles" / "r01_traffic.json").read_text())["k_lc_reduce_dram_bytes_per_record"]
    except Exception:
        pass
    out = {"kernel": "k_lc_reduce", "events": n_events, "bins": "postprocessing/mrk421_lc.input (1 mu window, 7 bands)"}
    for label, st_max in (("shared_acc_128_time_bins", 128), ("hbm_acc_1024_time_bins", 1024)):
        g.lc_begin(lc_reduction(cfg, st_max=st_max))
        g.lc_bench(n_events // 10)                  # warm-up
        best = min(g.lc_bench(n_events)[0] for _ in range(3))
        ach = 56.0 * n_events / (best * 1e-3) / 1e9
        out[label] = {"ms": best, "events_per_s": n_events / (best * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                   "traffic": per_rec * n_events if per_rec else None}}
    return out


# ------------------------------------------------------------------------------------------ our arm
FAM = {"k_history": 0, "k_children": 1, "k_children_teams": 5, "k_split3_attempts": 2, "k_split3_hard": 12, "k_split3_ultra": 13,
       "k_split3_flights": 6, "k_sigma_table": 4, "bank_sort": 11,
       "tally_allreduce": 8, "cascade_exchange": 10}      # c2d_kernel_time_ms families (tally_allreduce includes the wait for the slowest rank)


def run_tracking(args, dist, rank, world, local, nz, nr, groups, nst, steps, warmup, sampler=None):
    """K timed IMC steps of the tracking hot path on one workload.  Every step goes through the public host API with HOST
    buffers: c2d_step_begin uploads the step's tables (H2D), imcfield2d + imcvol2d track, graphics_collect drains the cascade
    and all-reduces the tallies inside the library (NCCL), tallies() reads them back (D2H).
      value = histories / device time of [imcfield2d .. graphics_collect] (CUDA events on the library's stream)
      e2e   = the same histories / host wall clock of the whole loop, copies included -- the same K cycles."""
    import numpy as np
    import torch
    from compton2d_b200.synth import context_for, make_problem
    P = make_problem(nz=nz, nr=nr, nst=nst * world, ngroups=groups)
    per_step = int(P["nsv"].sum()) // world + 1
    cap = int(per_step * (steps + warmup + 2) * 1.1) + 1000
    g = context_for(P, sigma_mode=args.sigma_mode, census_capacity=cap, event_capacity=max(per_step, 1000),
                    device=local, rank=rank, world=world, collision_capacity=1 << 18)
    if world > 1:
        g.comm_init()
    g.timing(True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(cyc):
        g.imcgen2d(ncycle=cyc, time=cyc * P["dt"], dt=P["dt"], kappa_tot=P["kappa_tot"], eps_tot=P["eps_tot"],
                   eps_th=P["eps_th"], f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], tea=P["tea"], f_pair=P["f_pair"],
                   k_gg=P["k_gg"], nsv=P["nsv"], ewsv=P["ewsv"], Eloss_th=P["Eloss_th"], Eloss_tot=P["Eloss_tot"],
                   zsurf=P["zsurf"], tbbl=P["tbbl"])
        g.timing_reset()
        g.mark()                                   # CUDA events on the library's stream bracket the tracked region
        g.imcfield2d()
        g.imcvol2d()
        g.graphics_collect()
        g.mark(stop=True)
        fam = {k: g.kernel_time(i) for k, i in FAM.items()}
        ms = g.kernel_time(7)[0]
        t = g.tallies()
        return ms, fam, t

    cyc = 1
    for _ in range(warmup):
        step(cyc); cyc += 1
    if sampler is not None:
        sampler.start()
    barrier()
    dev_ms, hist, launches = 0.0, 0, 0
    fam_ms = {k: 0.0 for k in FAM}
    seg = ndx = nev = 0
    wall0 = time.perf_counter()
    for _ in range(steps):
        h0 = g.census_count()
        ms, fam, t = step(cyc); cyc += 1
        dev_ms += ms
        if os.environ.get("BENCH_TRACE"):
            print(f"[bench r{rank}] cycle {cyc - 1} device {ms:.2f} ms, families " +
                  " ".join(f"{k}={m:.2f}/{n}" for k, (m, n) in fam.items()), file=sys.stderr)
        for k, (m, n) in fam.items():
            fam_ms[k] += m
            if k not in ("tally_allreduce", "cascade_exchange"):
                launches += n
        hist += h0 + (int(P["nsv"].sum()) - rank + world - 1) // world
        c = t["counters"]
        seg += c["segments"]; ndx += c["ndxout"]; nev += c["nevents"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler is not None else None
    red = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([hist, launches], dtype=torch.float64, device="cuda")
    per_rank = [dict(fam_ms, device_ms=dev_ms, histories=hist)]
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
    dev_ms_max, wall_max = red.tolist()
    hist_all, launches_all = cnt.tolist()
    nfl = P["E_field"].size
    h2d = 8 * sum(int(np.asarray(P[k]).size) for k in ("kappa_tot", "eps_tot", "eps_th", "f_nt", "Pnt", "n_e", "tea",
                                                       "f_pair", "k_gg", "ewsv", "Eloss_th", "Eloss_tot", "zsurf",
                                                       "tbbl")) + 4 * P["nsv"].size
    d2h = 8 * (P["nzone"] * (4 + nfl + P["E_gg"].size) + P["gnt"].size * 2 + 2 * 32 * (P["hu"].size - 1) + 32 * 16 + 64)
    # algorithmic bytes of the history kernel on this rank (DESIGN.md "bytes per history"): census record read 68 B, census
    # record written 68 B, deposition RMW 2 x 16 B per zone crossing, census tallies (npcen, ecens, n_field) 3 x 16 B,
    # escape (erlk, fout, edout RMW + 56 B record)
    n_in = hist - per_step * steps
    bytes_all = 68.0 * n_in + 68.0 * ndx + 32.0 * seg + 48.0 * ndx + (48.0 + 56.0) * nev
    stats = {k: {"min": min(r[k] for r in per_rank) / steps, "mean": sum(r[k] for r in per_rank) / len(per_rank) / steps,
                 "max": max(r[k] for r in per_rank) / steps} for k in list(FAM) + ["device_ms"]}
    out = dict(value=hist_all / (dev_ms_max * 1e-3), ms_per_step=dev_ms_max / steps, e2e=hist_all / wall_max,
               wall_ms_per_step=1e3 * wall_max / steps, histories_per_step=hist_all / steps, launches=int(launches_all),
               h2d=h2d, d2h=d2h, kernel_ms_per_step={k: fam_ms[k] / steps for k in FAM}, per_rank_ms_per_step=stats,
               bytes_per_step=bytes_all / steps, hist_kernel_ms=fam_ms["k_history"] / steps, census_in_per_step=n_in / steps,
               sources_per_step=per_step, counters=t["counters"], clocks=clocks)
    return g, P, out


def our_arm(args):
    import torch
    from compton2d_b200.host import Compton2d  # noqa: F401  (fails loudly without the CUDA library)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None

    # weak scaling: nst grows with the GPU count, every rank tracks h = rank (mod world)
    sampler = ClockSampler(local) if rank == 0 else None
    g, P, r = run_tracking(args, dist, rank, world, local, args.nz, args.nr, args.groups, args.nst, args.steps, args.warmup, sampler)

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    k_ms = r["hist_kernel_ms"]
    achieved = r["bytes_per_step"] / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # DRAM traffic of the same launches: per-history figures measured with ncu (profiles/r02_traffic.json)
    traffic, tsrc = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.loads((ROOT / "profiles" / name).read_text())
            traffic = tj["k_history_census_dram_bytes_per_history"] * r["census_in_per_step"] + \
                tj["k_history_volume_dram_bytes_per_history"] * r["sources_per_step"]
            tsrc = f"profiles/{name} (ncu dram bytes per history x histories of the step)"
            break
        except Exception:
            pass
    out = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args),
           "histories_per_step": r["histories_per_step"], "trial_flights_per_s": r["value"] * 1000,
           "wall_ms_per_step": r["wall_ms_per_step"],
           "e2e": {"value": r["e2e"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                   "note": "the same K cycles as `value`: host tables uploaded and tallies read back every step, host wall clock"},
           "gpu_launches": r["launches"],
           "kernel_ms_per_step": r["kernel_ms_per_step"], "per_rank_ms_per_step": r["per_rank_ms_per_step"],
           "roofline": {"bound": "hbm", "kernel": "k_history", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                        "frac": achieved / hbm, "traffic": traffic, "traffic_source": tsrc,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                        "bytes_per_step": r["bytes_per_step"],
                        "note": "latency-bound kernel (dependent L1/L2 loads at 28 warps / SM, issue slots 41 % busy: DESIGN.md 4, profiles/r02_ncu_summary.md); the HBM fraction is reported because the contract asks for it"},
           "counters": r["counters"], "clocks": r["clocks"]}
    if rank == 0 and not args.no_fp and world == 1:
        for leg, fn in (("lc", lambda: lc_bench(g, hbm)),):
            try:
                out[leg] = fn()
            except Exception as e:
                out[leg] = {"error": repr(e)[:300]}
    g.close()
    del g
    torch.cuda.empty_cache()
    # BASELINE configs[3] shape (64 x 64 zones, 200 groups, 1.25e8 sources per step and GPU: 1e9 histories per step on 8 GPUs),
    # a short leg sized so that its census fits HBM next to nothing else
    if not args.no_c3 and args.workload == "c1":
        try:
            a3 = argparse.Namespace(**vars(args))
            g3, P3, r3 = run_tracking(a3, dist, rank, world, local, 64, 64, 200, args.c3_nst, args.c3_steps, 2)
            g3.close()
            del g3
            torch.cuda.empty_cache()
            out["c3"] = {"workload": "BASELINE configs[3] shape: 64x64 zones, 200 groups, %d volume sources/step/GPU + census, split 1000/1000/300"
                                     % (args.c3_nst // 2), "steps": args.c3_steps, "warmup": 2,
                         "value": r3["value"], "unit": UNIT, "ms_per_step": r3["ms_per_step"], "e2e": r3["e2e"],
                         "histories_per_step": r3["histories_per_step"], "kernel_ms_per_step": r3["kernel_ms_per_step"],
                         "per_rank_ms_per_step": r3["per_rank_ms_per_step"]}
        except Exception as e:
            out["c3"] = {"error": repr(e)[:300]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if not args.no_fp and world == 1:
        try:
            out["fp"] = fp_bench(local, hbm)
            out["fp512"] = fp_bench(local, hbm, bins=512)
        except Exception as e:                     # the FP leg must never take the headline down with it
            out["fp"] = {"error": repr(e)[:300]}
    if not args.no_cpu_baseline and not args.no_fp and world == 1 and isinstance(out.get("fp"), dict) and "value" in out["fp"]:
        r = run_reference_sample(3000, ncores=1, warm=0, timed=1, fp=True, nz=8, nr=8)
        if r and "fp_zone_steps_per_s" in r:
            out["fp"]["cpu_baseline"] = {"value": r["fp_zone_steps_per_s"], "unit": "zone-steps/s", "cores": 1, "kind": "reference",
                                         "sample": f"FP farm of cycle 1 of the prebuilt reference binary, 8x8 zones x 200 bins: "
                                                   f"{r['fp_zone_steps']} zone-steps in {r['fp_seconds']:.1f} s on 1 worker core"}
    if not args.no_cpu_baseline and world == 1:
        r = run_reference_sample(args.cpu_nst, ncores=1, warm=0, timed=2)
        if r and "error" not in r:
            out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                                   "sample": f"cycles 1-2 of the prebuilt reference binary, 16x16 zones, 100 groups, nst={args.cpu_nst}, "
                                             f"{r['histories']} histories in {r['seconds']:.1f} s on 1 worker core"}
        else:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "unavailable: " + str(r)[:200]}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def legs_only(args):
    """--only fp,lc,tally: the stand-alone kernel legs without the tracking headline (profiling runs)."""
    import torch
    from compton2d_b200.synth import context_for as make_gpu, make_problem
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = {}
    legs = args.only.split(",")
    if "lc" in legs:
        P = make_problem(nz=16, nr=16, nst=1000, ngroups=100)
        g = make_gpu(P, census_capacity=1000, event_capacity=1000)
        if "lc" in legs:
            out["lc"] = lc_bench(g, hbm)
        g.close()
    if "fp" in legs:
        out["fp"] = fp_bench(0, hbm)
    if "fp512" in legs:
        out["fp512"] = fp_bench(0, hbm, bins=512)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nst", type=int, default=20_000_000, help="particles per photon cycle per GPU (nst; sources = nst/2)")
    ap.add_argument("--sigma-mode", type=int, default=1, help="0: comtot bin by bin per history; 1: per-step table (same values to 5e-11)")
    ap.add_argument("--ref-nst", type=int, default=6_000, help="--impl reference: nst per instance (census grows every cycle)")
    ap.add_argument("--cpu-nst", type=int, default=30_000, help="nst of the bounded cpu_baseline sample (cycles 1-2, one core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp", action="store_true", help="skip the Fokker-Planck zone-steps/s leg")
    ap.add_argument("--workload", default="c1", choices=["c1", "c3"],
                    help="c1 = BASELINE configs[1] (the bench line); c3 = configs[3] shape, 64x64 zones x 200 groups, 1.25e8 sources/GPU")
    ap.add_argument("--no-c3", action="store_true", help="skip the short BASELINE configs[3]-shape leg")
    ap.add_argument("--c3-nst", type=int, default=250_000_000, help="nst per GPU of the configs[3] leg (sources = nst/2)")
    ap.add_argument("--c3-steps", type=int, default=2)
    ap.add_argument("--only", default="", help="comma list of stand-alone legs (fp, fp512, lc) to run instead of the bench line")
    args = ap.parse_args()
    args.nz, args.nr, args.groups = (64, 64, 200) if args.workload == "c3" else (16, 16, 100)
    if args.workload == "c3" and args.nst == 20_000_000:
        args.nst = 250_000_000
    if args.only:
        legs_only(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
