"""Host-side preparation of external seed-photon spectra for the surface sources.

`file_sp` mirrors the reference subroutine of the same name (src_20121026/imcsurf2d_para.f:544-685): it reads
a table `E [keV], L_disk, F_blr, F_ir`, normalises the BLR and torus fluxes after Ghisellini & Madau (1996), and
builds the piecewise power-law cumulative distribution the device samples from (`file_sample`,
imcsurf2d_para.f:694-788).  The reference re-reads and re-parses the file once per boundary ring per step on the
master (imcgen2d.f:157-158,175-176) and again on the worker (imcsurf2d_para.f:374-375,441-442); here it is
done once and the tables are handed to `Compton2d.imcsurf2d`.  Pure host logic (file parsing): nothing on the
hot path runs here."""
from __future__ import annotations

import numpy as np

NFMAX = 500                 # general.pa:17
PI = 3.1415926536e0         # general.pa:25


def read_spectrum(path):
    """The list-directed READ of file_sp: rows of at least four numbers."""
    return np.loadtxt(path, ndmin=2)[:, :4]


def file_sp(rows, *, g_bulk, R_blr, fr_blr, R_ir, fr_ir, R_disk=None, d_jet=None):
    """-> dict(E_file, F_file, P_file, a1, I_file [NFMAX each], int_file, nfile).

    rows: [n, >=4] array or a file name.  The read loop stops after row NFMAX-1 or at the first E <= 0
    (imcsurf2d_para.f:567-594); the disk component (Id_file, F_como) is computed by the reference but does not enter
    F_file (l.648-649), so R_disk and d_jet are accepted and unused."""
    if isinstance(rows, (str, bytes)) or hasattr(rows, "__fspath__"):
        rows = read_spectrum(rows)
    rows = np.asarray(rows, float)
    E = np.zeros(NFMAX); Ld = np.zeros(NFMAX); Fb = np.zeros(NFMAX); Fi = np.zeros(NFMAX)
    i = 0
    while True:
        i += 1
        if i > rows.shape[0]:
            break
        E[i - 1], Ld[i - 1], Fb[i - 1], Fi[i - 1] = rows[i - 1, :4]
        if not (E[i - 1] > 0.0 and i <= NFMAX - 1):
            break
    nfile = i - 1
    if nfile < 2:
        raise ValueError("Error in input spectrum file: less than 2 lines of input read (imcsurf2d_para.f:652-657)")
    dE = E[1:] - E[:-1]
    Ltot_disk = Ftot_blr = Ftot_ir = 0.0
    for k in range(NFMAX - 1):                      # sequential sums, as the reference accumulates them
        Ltot_disk += Ld[k] * dE[k]; Ftot_blr += Fb[k] * dE[k]; Ftot_ir += Fi[k] * dE[k]
    sq = np.sqrt(E[1] / E[0])
    Ltot_disk /= sq; Ftot_blr /= sq; Ftot_ir /= sq
    blr_norm = 17.0 / 48.0 / PI * g_bulk ** 2 * fr_blr * Ltot_disk / R_blr ** 2
    ir_norm = 1.0 / 4.0 / PI * g_bulk ** 2 * fr_ir * Ltot_disk / R_ir ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        F = Fb / Ftot_blr * blr_norm + Fi / Ftot_ir * ir_norm
        a1 = np.zeros(NFMAX); I = np.zeros(NFMAX); P = np.zeros(NFMAX)
        Isum = 0.0
        for k in range(nfile - 1):
            alpha = np.log(F[k + 1] / F[k]) / np.log(E[k + 1] / E[k])
            a = min(max(alpha + 1.0, -2.0e1), 2.0e1)
            a1[k] = a
            if abs(a) < 1.0e-3:
                I[k] = F[k] * E[k] * np.log(E[k + 1] / E[k])
            else:
                I[k] = F[k] * E[k] * ((E[k + 1] / E[k]) ** a - 1.0) / a
            Isum += I[k]
            P[k] = Isum
    P[: nfile - 1] /= Isum
    return dict(E_file=E, F_file=F, P_file=P, a1=a1, I_file=I, int_file=float(Isum), nfile=int(nfile))


SIGMA_SB = 1.0267e24        # imcgen2d.f:14 (erg cm^-2 s^-1 keV^-4)


def surface_budget(*, nst, time, dt, z, r, rmin=0.0, tbbi=None, tbbo=None, tbbu=None, tbbl=None, t0=0.0, int_file=None,
                   star_switch=0, Rstar=1.0, dist_star=1.0, n_volume=0):
    """The surface-source half of imcgen2d (imcgen2d.f:120-190 energies entering through each boundary ring, 424-443 photon
    numbers, 457-476 weights, 491-517 the > 10 nst cap) -- host arithmetic over nz + nr rings, the remaining piece of
    SURVEY.md 8f N1 next to `Compton2d.imcgen2d_device` (volume sources).

    tbb*: boundary temperatures of the current time interval [keV]; a ring with tbb < 0 (and time + dt/2 >= t0) is fed by a
    spectrum file whose integral `int_file` (file_sp) may be a scalar or a dict(i=, o=, u=, l=) of per-ring arrays.
    Returns dict(nsurf{i,o,u,l}, ewsurf{i,o,u,l}, erin{i,o,u,l}, fbias) ready for `Compton2d.imcsurf2d`; `n_volume` (sum of
    nsv) only enters the cap, whose factor the caller applies to nsv / ewsv too when fbias != 1."""
    z = np.asarray(z, float); r = np.asarray(r, float)
    nz, nr = z.size, r.size
    zlo = np.concatenate(([0.0], z[:-1])); rlo = np.concatenate(([rmin], r[:-1]))
    A = dict(i=2.0 * PI * rmin * (z - zlo), o=2.0 * PI * r[-1] * (z - zlo), u=PI * (r ** 2 - rlo ** 2), l=PI * (r ** 2 - rlo ** 2))
    tb = dict(i=tbbi, o=tbbo, u=tbbu, l=tbbl)
    n_ring = dict(i=nz, o=nz, u=nr, l=nr)
    out, n_new = {}, int(n_volume)
    for f in "ioul":
        t = np.zeros(n_ring[f]) if tb[f] is None else np.asarray(tb[f], float)
        intf = int_file.get(f, 0.0) if isinstance(int_file, dict) else (0.0 if int_file is None else int_file)
        intf = np.broadcast_to(np.asarray(intf, float), t.shape)
        from_file = (t < 0.0) & (time + 0.5 * dt >= t0)
        erin = np.where(from_file, dt * A[f] * intf, dt * A[f] * SIGMA_SB * t ** 4.0)
        if f == "u" and star_switch == 1:
            erin = np.where(from_file, erin, erin * (Rstar / dist_star) ** 2.0)
        if f in "io":
            n = np.where(t < 0.0, nst // nz, 0).astype(np.int64)                       # integer division, l.427-428
        else:
            n = np.where(t < 0.0, (nst * (r ** 2 - rlo ** 2) / r[-1] ** 2).astype(np.int64), 0)     # truncation, l.435-436
        out["erin" + f] = erin
        out["nsurf" + f] = n
        n_new += int(n.sum())
    for f in "ioul":
        n = out["nsurf" + f]
        out["ewsurf" + f] = np.where(n > 0, out["erin" + f] / np.maximum(n, 1), 0.0)
    out["fbias"] = 1.0
    if n_new > 10 * nst:                                                             # l.491-517
        fb = float(10 * nst) / float(n_new)
        for f in "ioul":
            out["nsurf" + f] = (out["nsurf" + f].astype(float) * fb).astype(np.int64)
            out["ewsurf" + f] = out["ewsurf" + f] / fb
        out["fbias"] = fb
    for f in "ioul":
        out["nsurf" + f] = out["nsurf" + f].astype(np.int32)
    return out
