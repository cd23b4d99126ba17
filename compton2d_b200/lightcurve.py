"""Host side of the event-list reduction: the reference's light-curve post-processor
postprocessing/plcm.c, whose binning runs on the device (csrc/c2d_lc.cu) straight from the escape records instead of from
a 105-bytes-per-photon text file.

`parse_lc_input` reads the same answer files plcm is driven with (`plcm < mrk421_lc.input`, prompts of plcm.c:104-300);
`finalize` turns the raw accumulators into the quantities plcm prints (plcm.c:470-496)."""
from __future__ import annotations

import numpy as np

ST_MAX, MU_MAX, CH_MAX = 1024, 10, 20          # plcm.c:6-8


def parse_lc_input(text):
    """-> dict(infile, gam_bulk, rmax, outfile, n_mu, mu0, mu1, dt, t_offset, t_max, E0, E1).  An empty answer keeps the
    default of the prompt, as in plcm.c's dinput / iinput."""
    lines = [l.split("!")[0].strip() for l in text.splitlines()]
    it = iter(lines)

    def ans():
        return next(it, "")

    def num(default, conv=float):
        a = ans()
        return conv(float(a.split()[0])) if a else default
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
    regions = num(7, int)
    defaults = [(1e-3, 3e-3), (2, 4), (9, 15), (15, 20), (20, 60), (5e5, 5e7), (1e9, 1e10)]
    E0, E1 = [], []
    lo, hi, n_r = 1e-3, 3e-3, 1
    for reg in range(regions):
        lo, hi = defaults[reg] if reg < len(defaults) else (hi, hi)
        lo = num(lo); hi = num(hi); n_r = num(n_r, int)
        if len(E0) + n_r > CH_MAX:
            raise ValueError("Not more than 20 energy channels (plcm.c:279)")
        linear = ans()[:1] == "1"
        e = lo
        step = (hi - lo) / n_r if linear else np.exp(np.log(hi / lo) / n_r)
        for _ in range(n_r):
            E0.append(e)
            e = e + step if linear else e * step
            E1.append(e)
    return dict(infile=infile, gam_bulk=gam_bulk, rmax=rmax, outfile=outfile, n_mu=n_mu, mu0=np.array(mu0[:n_mu]),
                mu1=np.array(mu1[:n_mu]), dt=dt, t_offset=t_offset, t_max=t_max, E0=np.array(E0), E1=np.array(E1))


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
