"""ctypes binding of libcompton2d_b200.so (the C ABI in include/compton2d_b200.h).

The library is the product; this module only loads it.  It fails loudly when the
shared object is missing -- there is no Python/CPU fallback of any kernel.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIBPATH = Path(os.environ.get("C2D_LIBPATH", HERE / "libcompton2d_b200.so"))   # override: kernel-tuning experiments

NMUMAX = 32
NLCMAX = 16
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nz", "nr", "n_vol", "num_nt", "nphfield", "n_gg", "nphtotal", "nph_lc",
                                       "nmu", "pair_switch", "cr_sent", "spec_switch", "split1", "split2",
                                       "split3", "spl3_trg")] + \
        [("census_capacity", C.c_int64), ("event_capacity", C.c_int64), ("collision_capacity", C.c_int64),
         ("rseed", C.c_uint32), ("sigma_mode", C.c_int), ("rmin", C.c_double), ("zmin", C.c_double)] + \
        [(n, dp) for n in ("z", "r", "E_ph", "E_field", "E_gg", "gnt", "hu", "Elcmin", "Elcmax", "mu")] + \
        [("n_ref", C.c_int), ("E_ref", dp), ("P_ref", dp), ("W_abs", dp)]


class StepIn(C.Structure):
    _fields_ = [("ncycle", C.c_int), ("time", C.c_double), ("dt", C.c_double)] + \
        [(n, dp) for n in ("kappa_tot", "eps_tot", "eps_th", "f_nt", "Pnt", "n_e", "tea", "f_pair", "k_gg")] + \
        [("nsv", ip)] + [(n, dp) for n in ("ewsv", "Eloss_th", "Eloss_tot", "zsurf", "tbbl")]


class GenIn(C.Structure):
    _fields_ = [("ncycle", C.c_int), ("time", C.c_double), ("dt", C.c_double), ("nst", C.c_int64)] + \
        [(n, dp) for n in ("f_nt", "Pnt", "n_e", "tea", "B_field", "vol", "zsurf", "f_pair", "k_gg", "tbbl")]


class GenOut(C.Structure):
    _fields_ = [("nsv", ip)] + [(n, dp) for n in ("ewsv", "Eloss_sy", "Eloss_th", "Eloss_tot")] + \
        [("Emiss_tot", C.c_double), ("n_volume", C.c_int64)]


class SurfIn(C.Structure):
    _fields_ = [("star_switch", C.c_int)] + [(n, ip) for n in ("nsurfi", "nsurfo", "nsurfu", "nsurfl")] + \
        [(n, dp) for n in ("ewsurfi", "ewsurfo", "ewsurfu", "ewsurfl", "tbbi", "tbbo", "tbbu", "tbbl")] + \
        [(n, ip) for n in ("spec_i", "spec_o", "spec_u", "spec_l")] + [("n_spec", C.c_int), ("nfmax", C.c_int), ("nfile", ip)] + \
        [(n, dp) for n in ("E_file", "F_file", "P_file", "a1", "I_file")]


class Tallies(C.Structure):
    _fields_ = [("edep", dp), ("prdep", dp), ("ecens", dp), ("npcen", lp), ("n_field", dp), ("n_ph", dp),
                ("E_IC", dp), ("nelectron", lp), ("fout", dp), ("edout", dp), ("erlki", dp), ("erlko", dp),
                ("erlku", dp), ("erlkl", dp), ("Ed_in", dp), ("Ed_ref", dp), ("counters", C.c_int64 * 16),
                ("fout_total", dp), ("counters_global", C.c_int64 * 16)]


class FpIn(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("first_zone", "n_zones", "zone_stride", "inj_switch", "inj_dis",
                                       "g2var_switch", "pick_sw", "cf_sentinel", "ncycle")] + \
        [(n, C.c_double) for n in ("time", "dt", "lnL", "inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma",
                                   "inj_L", "inj_v", "pick_rate", "r_esc", "r_acc", "r_flare", "z_flare",
                                   "t_flare", "sigma_r", "sigma_z", "sigma_t", "flare_amp")] + \
        [(n, dp) for n in ("f_nt", "Pnt", "n_field", "ecens", "n_e", "B_field", "tea", "tna", "vol", "f_pair",
                           "Eloss_sy", "ec_old", "turb_lev", "F_IC")]


class FpOut(C.Structure):
    _fields_ = [(n, dp) for n in ("f_nt", "Pnt", "n_e", "Te_new", "gmin", "gmax", "p_nth", "amxwl", "f_pair")] + \
        [("fp_steps", ip)] + \
        [(n, dp) for n in ("zone_E_old", "zone_E_new", "zone_hr", "zone_hr_st", "zone_dT")] + \
        [(n, C.c_double) for n in ("dT_max", "E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dr", "dz",
                                   "f_t_implicit")]


class Plasma(C.Structure):
    _fields_ = [(n, dp) for n in ("f_nt", "Pnt", "n_e", "tea", "tna", "B_field", "f_pair", "gmin", "gmax", "p_nth", "amxwl",
                                  "turb_lev", "vol", "zsurf", "k_gg")]


class PlasmaOut(C.Structure):
    _fields_ = [(n, dp) for n in ("f_nt", "Pnt", "n_e", "tea", "f_pair", "B_field", "gmin", "gmax", "p_nth", "amxwl", "Te_new",
                                  "Eloss_sy", "ec_old")]


class SurfRun(C.Structure):
    _fields_ = [("star_switch", C.c_int), ("Rstar", C.c_double), ("dist_star", C.c_double), ("t0", C.c_double)] + \
        [(n, dp) for n in ("tbbi", "tbbo", "tbbu", "tbbl")] + [(n, ip) for n in ("spec_i", "spec_o", "spec_u", "spec_l")] + \
        [("n_spec", C.c_int), ("nfmax", C.c_int), ("nfile", ip), ("int_file", dp)] + \
        [(n, dp) for n in ("E_file", "F_file", "P_file", "a1", "I_file")]


class Run(C.Structure):
    _fields_ = [("nst", C.c_int64), ("T_const", C.c_int), ("quirk_cycle0", C.c_int)] + \
        [(n, C.c_int) for n in ("inj_switch", "inj_dis", "g2var_switch", "pick_sw", "cf_sentinel")] + \
        [(n, C.c_double) for n in ("inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma", "inj_L", "inj_v", "pick_rate",
                                   "r_esc", "r_acc", "r_flare", "z_flare", "t_flare", "sigma_r", "sigma_z", "sigma_t",
                                   "flare_amp")] + [("surf", C.POINTER(SurfRun))]


class StepReport(C.Structure):
    _fields_ = [("ncycle", C.c_int), ("time", C.c_double), ("dt", C.c_double), ("n_census_in", C.c_int64),
                ("n_volume", C.c_int64), ("n_surface", C.c_int64)] + \
        [(n, C.c_int64) for n in ("histories", "ndxout", "nevents", "cmcount1", "cmcount2")] + \
        [(n, C.c_double) for n in ("bingo", "Emiss_tot", "fbias", "luminosity")] + \
        [(n, C.c_int) for n in ("fp_ran", "fp_steps_min", "fp_steps_max")] + \
        [(n, C.c_double) for n in ("E_tot_old", "E_tot_new", "dE_fraction", "dT_max", "dt_new", "hr_total", "hr_st_total")] + \
        [("ms", C.c_double * 16)]


class Common(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("jmax", "kmax", "ntmax", "nmumax", "nphomax", "nphlcmax", "ti")] + \
        [(n, dp) for n in ("kappa_tot", "eps_tot", "eps_th", "f_nt", "Pnt", "n_e", "tea", "tna", "B_field", "f_pair", "k_gg")] + \
        [("nsv", ip)] + [(n, dp) for n in ("ewsv", "Eloss_th", "Eloss_tot", "zsurf", "tbbl", "edep", "prdep", "ecens")] + \
        [("npcen", ip)] + [(n, dp) for n in ("n_field", "n_ph", "E_IC")] + [("nelectron", ip)] + \
        [(n, dp) for n in ("fout", "edout", "erlki", "erlko", "erlku", "erlkl", "Ed_in", "Ed_ref")] + \
        [(n, ip) for n in ("ndxout", "cmcount1", "cmcount2", "lkcount")] + \
        [(n, dp) for n in ("vol", "Eloss_sy", "ec_old", "turb_lev", "gmin", "gmax", "p_nth", "amxwl", "Te_new", "dT_max", "E_tot_old",
                           "E_tot_new", "hr_total", "hr_st_total", "dr", "dz", "f_t_implicit")]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class CommExt(C.Structure):
    _fields_ = [("user", C.c_void_p), ("allreduce", ALLREDUCE_FN), ("allgather", ALLGATHER_FN)]


COMM_ID_BYTES = 128
FAMILIES = ["k_history", "k_children", "k_split3_attempts", "k_fp_zones", "k_sigma_table", "k_children_teams", "k_split3_flights",
            "bracket", "tally_allreduce", "fp_allgather", "cascade_exchange", "bank_sort", "k_split3_hard", "k_split3_ultra",
            "(unused)", "(unused)"]


class LcConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gam_bulk", "rmax", "t_offset")] + \
        [(n, C.c_int) for n in ("n_t", "n_mu", "n_bands", "mu_closed", "first_band_only")] + \
        [(n, dp) for n in ("t_lo", "t_hi", "mu0", "mu1", "E0", "E1")]


COUNTERS = ["ndxout", "cmcount1", "cmcount2", "lkcount", "nevents", "histories", "trials", "multiple", "killed",
            "teams", "literal", "qpeak", "fatal", "segments", "s3passes", "s3hard"]

EXPORTS = ["c2d_create", "c2d_init", "c2d_destroy", "c2d_last_error", "c2d_version", "c2d_step_begin",
           "c2d_imcfield", "c2d_imcvol", "c2d_sync", "c2d_tally_read", "c2d_tally_device", "c2d_census_count",
           "c2d_census_export", "c2d_census_import", "c2d_events_drain", "c2d_ic_loss", "c2d_update_fp",
           "c2d_pack_bin_first", "c2d_pack_bin_last", "c2d_unpack_bin_first", "c2d_unpack_bin_last",
           "c2d_kernel_time_ms", "c2d_comtot", "c2d_surf_begin", "c2d_imcsurf", "c2d_step_begin_gen", "c2d_step_tables_read",
           "c2d_lc_select", "c2d_lc_begin", "c2d_lc_accumulate", "c2d_lc_accumulate_events", "c2d_lc_read", "c2d_lc_device", "c2d_lc_bench", "c2d_gamma_bar", "c2d_selftest_gap", "c2d_pref_calc",
           "c2d_census_write", "c2d_census_read", "c2d_events_write", "c2d_format_e14_7", "c2d_tally_reduce", "c2d_finalize",
           "c2d_comm_unique_id", "c2d_comm_init", "c2d_comm_init_ext", "c2d_comm_info", "c2d_comm_destroy",
           "c2d_plasma_set", "c2d_plasma_get", "c2d_step", "c2d_common_step_begin", "c2d_common_collect", "c2d_common_update"]

_lib = None


class C2dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"compton2d_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIBPATH.exists():
        raise ImportError(f"{LIBPATH} is missing: build it with `python compton2d_b200/build.py` "
                          "(nvcc, sm_100a). compton2d_b200 has no CPU fallback.")
    L = C.CDLL(str(LIBPATH))
    vp = C.c_void_p
    L.c2d_last_error.restype = C.c_char_p
    L.c2d_create.argtypes = [C.POINTER(vp), C.c_int]
    L.c2d_init.argtypes = [vp, C.POINTER(Dims)]
    L.c2d_destroy.argtypes = [vp]
    L.c2d_step_begin.argtypes = [vp, C.POINTER(StepIn)]
    L.c2d_imcfield.argtypes = [vp]
    L.c2d_step_begin_gen.argtypes = [vp, C.POINTER(GenIn), C.POINTER(GenOut)]
    L.c2d_step_tables_read.argtypes = [vp, vp, vp, vp]
    L.c2d_imcvol.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.c2d_sync.argtypes = [vp]
    L.c2d_surf_begin.argtypes = [vp, C.POINTER(SurfIn)]
    L.c2d_imcsurf.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.c2d_tally_read.argtypes = [vp, C.POINTER(Tallies)]
    L.c2d_tally_reduce.argtypes = [vp, C.POINTER(Tallies)]
    L.c2d_finalize.argtypes = [vp]
    L.c2d_tally_device.argtypes = [vp, C.POINTER(vp), lp, C.POINTER(vp), lp]
    L.c2d_census_count.argtypes = [vp, lp]
    L.c2d_census_export.argtypes = [vp, vp, vp, vp, C.c_int64, lp]
    L.c2d_census_import.argtypes = [vp, vp, vp, vp, C.c_int64]
    L.c2d_events_drain.argtypes = [vp, vp, C.c_int64, lp]
    L.c2d_ic_loss.argtypes = [vp, vp]
    L.c2d_update_fp.argtypes = [vp, C.POINTER(FpIn), C.POINTER(FpOut)]
    L.c2d_comtot.argtypes = [vp, C.c_int64, vp, vp, C.c_int, vp]
    L.c2d_kernel_time_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float), lp]
    L.c2d_census_write.argtypes = [vp, C.c_char_p, lp]
    L.c2d_census_read.argtypes = [vp, C.c_char_p, C.c_int64]
    L.c2d_events_write.argtypes = [vp, C.c_char_p, C.c_int, lp]
    L.c2d_format_e14_7.argtypes = [C.c_double, C.c_char_p]
    L.c2d_format_e14_7.restype = None
    L.c2d_gamma_bar.argtypes = [vp, C.c_int64, vp, vp, vp]
    L.c2d_selftest_gap.argtypes = [vp, C.c_int64, C.c_uint64, vp, vp]
    L.c2d_pref_calc.argtypes = [vp, C.c_int, vp, vp]
    L.c2d_lc_select.argtypes = [vp, C.c_int]
    L.c2d_lc_begin.argtypes = [vp, C.POINTER(LcConfig)]
    L.c2d_lc_accumulate.argtypes = [vp]
    L.c2d_lc_accumulate_events.argtypes = [vp, vp, C.c_int64]
    L.c2d_lc_read.argtypes = [vp, vp, vp, vp]
    L.c2d_lc_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), lp]
    L.c2d_lc_bench.argtypes = [vp, C.c_int64, C.POINTER(C.c_float), lp]
    L.c2d_comm_unique_id.argtypes = [vp]
    L.c2d_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.c2d_comm_init_ext.argtypes = [vp, C.c_int, C.c_int, C.POINTER(CommExt)]
    L.c2d_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.c2d_comm_destroy.argtypes = [vp]
    L.c2d_plasma_set.argtypes = [vp, C.POINTER(Plasma)]
    L.c2d_plasma_get.argtypes = [vp, C.POINTER(PlasmaOut)]
    L.c2d_step.argtypes = [vp, C.POINTER(Run), C.c_int, C.c_double, C.c_double, C.POINTER(StepReport)]
    L.c2d_common_step_begin.argtypes = [vp, C.POINTER(Common), C.c_int, C.c_double, C.c_double]
    L.c2d_common_collect.argtypes = [vp, C.POINTER(Common)]
    L.c2d_common_update.argtypes = [vp, C.POINTER(Common), C.POINTER(FpIn)]
    for fn in ("c2d_pack_bin_first", "c2d_unpack_bin_first"):
        getattr(L, fn).argtypes = [vp] + [C.c_int] * 6 + [vp]
        getattr(L, fn).restype = None
    for fn in ("c2d_pack_bin_last", "c2d_unpack_bin_last"):
        getattr(L, fn).argtypes = [vp] + [C.c_int] * 5 + [vp]
        getattr(L, fn).restype = None
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise C2dError(rc, load().c2d_last_error().decode())
