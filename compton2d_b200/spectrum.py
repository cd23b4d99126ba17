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
