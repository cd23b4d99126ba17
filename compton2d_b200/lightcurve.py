"""Host side of the event-list reduction: the reference's post-processors postprocessing/plcm.c (light
curves) and postprocessing/pspt.c (time-resolved SEDs), whose per-photon transform and binning run on the device
(csrc/c2d_lc.cu) straight from the escape records instead of from a 105-bytes-per-photon text file.

`parse_lc_input` / `parse_sed_input` read the same answer files the two programs are driven with (`plcm < mrk421_lc.input`,
prompts of plcm.c:104-300; `pspt < mrk421_sed.input`, pspt.c:98-196); `lc_reduction` / `sed_reduction` turn them into the
bin tables of `c2d_lc_config` (include/compton2d_b200.h); `finalize` / `finalize_sed` turn the raw accumulators into the
quantities the programs print (plcm.c:470-496, pspt.c:318-341).  No binning happens here."""
from __future__ import annotations

import math

import numpy as np

ST_MAX, MU_MAX, CH_MAX = 1024, 10, 20          # plcm.c:6-8
SED_CH_MAX, SED_T_MAX = 200, 90                # pspt.c:8-10



def _energy_grid(answers, num, defaults, ch_max, start=(1e-3, 3e-3, 1), carry_upper=False):
    """The energy-region prompts shared by plcm.c:243-300 and pspt.c:159-193."""
    regions = num(len(defaults) if defaults else 1, int)
    E0, E1 = [], []
    lo, hi, n_r = start
    for reg in range(regions):
        if defaults:
            lo, hi = defaults[reg] if reg < len(defaults) else (hi, hi)
        lo = num(lo); hi = num(hi); n_r = num(n_r, int)
        if len(E0) + n_r > ch_max:
            raise ValueError(f"Not more than {ch_max} energy channels (plcm.c:279, pspt.c:172)")
        linear = answers()[:1] == "1"
        e = lo
        step = (hi - lo) / n_r if linear else math.exp(math.log(hi / lo) / n_r)     # libm, as the C programs
        for _ in range(n_r):
            E0.append(e)
            e = e + step if linear else e * step
            E1.append(e)
        if carry_upper:
            lo = hi                                       # pspt.c:192 E_lower = E_upper
    return np.array(E0), np.array(E1)


def _answers(text):
    it = iter([l.split("!")[0].strip() for l in text.splitlines()])

    def ans():
        return next(it, "")

    def num(default, conv=float):
        a = ans()
        return conv(float(a.split()[0])) if a else default
    return ans, num


def parse_lc_input(text):
    """-> dict(infile, gam_bulk, rmax, outfile, n_mu, mu0, mu1, dt, t_offset, t_max, E0, E1).  An empty answer keeps the
    default of the prompt, as in plcm.c's dinput / iinput."""
    ans, num = _answers(text)
    infile = ans() or "p001_evb.dat"
    gam_bulk = num(33.0)
    rmax = num(1.0e16)
    outfile = ans() or "lc07_ev0.dat"
    n_mu = min(num(1, int), MU_MAX)
    mu0, mu1 = [0.99944] + [0.0] * MU_MAX, [0.99964] + [0.0] * MU_MAX
    for k in range(n_mu):
        ans()                                            # output file of this bin
        mu0[k] = num(mu0[k]); mu1[k] = num(mu1[k])
        if k < MU_MAX - 1:
            mu0[k + 1] = mu1[k + 1] = mu1[k]
    dt = num(7.0e2); t_offset = num(0.0); t_max = num(7.0e4)
    defaults = [(1e-3, 3e-3), (2, 4), (9, 15), (15, 20), (20, 60), (5e5, 5e7), (1e9, 1e10)]
    E0, E1 = _energy_grid(ans, num, defaults, CH_MAX)
    return dict(infile=infile, gam_bulk=gam_bulk, rmax=rmax, outfile=outfile, n_mu=n_mu, mu0=np.array(mu0[:n_mu]),
                mu1=np.array(mu1[:n_mu]), dt=dt, t_offset=t_offset, t_max=t_max, E0=E0, E1=E1)


def parse_sed_input(text):
    """Answers of pspt.c:98-196 -> dict(infile, gam_bulk, rmax, outfile, n_t, t_start, t_end, mu0, mu1, E0, E1)."""
    ans, num = _answers(text)
    infile = ans() or "p001_evb.dat"
    gam_bulk = num(33.0); rmax = num(1.0e16)
    outfile = ans() or "seds_30.dat"
    n_t = min(num(30, int), SED_T_MAX)
    t_start = num(1.6e4); t_end = num(6e4)
    mu0 = num(0.99944); mu1 = num(0.99964)
    E0, E1 = _energy_grid(ans, num, None, SED_CH_MAX, start=(1e-7, 1e10, 100), carry_upper=True)
    return dict(infile=infile, gam_bulk=gam_bulk, rmax=rmax, outfile=outfile, n_t=n_t, t_start=t_start, t_end=t_end,
                mu0=mu0, mu1=mu1, E0=E0, E1=E1)


def lc_reduction(cfg, st_max=ST_MAX):
    """plcm's bin rules as the tables of c2d_lc_config: time edges by repeated addition (plcm.c:236-240), half-open mu bins,
    every matching band."""
    t_lo = np.zeros(st_max)
    for k in range(st_max - 1):
        t_lo[k + 1] = t_lo[k] + cfg["dt"]
    return dict(gam_bulk=cfg["gam_bulk"], rmax=cfg["rmax"], t_offset=cfg["t_offset"], t_lo=t_lo, t_hi=t_lo + cfg["dt"],
                mu0=np.asarray(cfg["mu0"], float), mu1=np.asarray(cfg["mu1"], float), mu_closed=0,
                E0=np.asarray(cfg["E0"], float), E1=np.asarray(cfg["E1"], float), first_band_only=0)


def sed_reduction(cfg):
    """pspt's bin rules (pspt.c:143-148, 274-280): t0[n] = t_start + n dt, t1[n] = t_start + n dt + dt, one closed mu
    window, first matching band only."""
    dt = (cfg["t_end"] - cfg["t_start"]) / cfg["n_t"]
    n = np.arange(cfg["n_t"], dtype=np.float64)
    t_lo = cfg["t_start"] + n * dt
    return dict(gam_bulk=cfg["gam_bulk"], rmax=cfg["rmax"], t_offset=0.0, t_lo=t_lo, t_hi=t_lo + dt,
                mu0=np.array([cfg["mu0"]]), mu1=np.array([cfg["mu1"]]), mu_closed=1,
                E0=np.asarray(cfg["E0"], float), E1=np.asarray(cfg["E1"], float), first_band_only=1)


def finalize_sed(sum_ew, count, cfg):
    """pspt.c:318-341: -> dict(E = sqrt(E0 E1) [band], F [band][time] = luminosity per keV, time and solid-angle window
    (floored at 1e-20 as printed), particle_sum [band] of the LAST time bin (the only one pspt prints))."""
    dt = (cfg["t_end"] - cfg["t_start"]) / cfg["n_t"]
    s = np.asarray(sum_ew, float)[:, 0, :]                       # [time][band]
    F = s / (dt * (cfg["E1"] - cfg["E0"])[None, :] * (cfg["mu1"] - cfg["mu0"]) / 2.0)
    return dict(E=np.maximum(1e-20, np.sqrt(cfg["E0"] * cfg["E1"])), F=np.maximum(1e-20, F.T),
                particle_sum=np.asarray(count)[-1, 0, :])


def finalize(sum_ew, sum_ew2, count, dt, mu0, mu1):
    """plcm.c:470-496 from the raw accumulators [time bin][mu bin][band] -> dict(F, F1, F2, sF, particle_sum):
    F = luminosity per time and solid-angle bin, F1 = <ew>, F2 = <ew^2>, sF = sqrt(var(ew)) (0 for a single photon,
    1e-20 for an empty bin)."""
    sum_ew = np.asarray(sum_ew, float); sum_ew2 = np.asarray(sum_ew2, float); count = np.asarray(count)
    F = sum_ew / (dt * (np.asarray(mu1) - np.asarray(mu0))[None, :, None] / 2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        F1 = np.where(count > 0, sum_ew / np.maximum(count, 1), 1e-20)
        F2 = np.where(count > 0, sum_ew2 / np.maximum(count, 1), 1e-20)
        sF = np.where(count > 0, np.sqrt(np.maximum(F2 - F1 * F1, 0.0)), 1e-20)
    sF = np.where(count == 1, 0.0, sF)
    return dict(F=F, F1=F1, F2=F2, sF=sF, particle_sum=count)
