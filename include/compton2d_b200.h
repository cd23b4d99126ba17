/*
 * compton2d_b200.h -- C ABI of the B200-native hot path of Compton2d.
 *
 * This is the drop-in boundary.  The reference has no FFI: its
 * "interface" is `call name` with no arguments plus ~90 COMMON blocks
 * (src_20121026/commonblock.f:118-219).  A thin ISO_C_BINDING Fortran shim that
 * keeps the reference's subroutine names (imcfield2d, imcvol2d, imcsurf2d, update, ...)
 * packs the COMMON members listed below into dense arrays and calls these entry
 * points (INTEGRATION.md shows the shim).  Plain C types only; every function
 * returns 0 on success and a non-zero C2D_E* code otherwise -- the shim prints
 * c2d_last_error() and STOPs, which is the reference's own error behaviour
 * (e.g. imctrk2d.f:583-584 `write(*,*) 'too many photons'; stop`).
 *
 * Layout convention of every per-zone table: dense, row-major [zone][bin], with
 *   zone = (k-1) + nr*(j-1),  j = z index 1..nz,  k = r index 1..nr
 * (fp_mpi.f:333-352 get_j_k/get_zone, 0-based).  c2d_pack_bin_first/_bin_last convert from
 * the padded column-major COMMON arrays (kappa_tot(n_vol,jmax,kmax) / f_nt(jmax,kmax,num_nt)).
 *
 * There is NO CPU fallback: without a CUDA device c2d_create fails with C2D_ENODEV.
 */
#ifndef COMPTON2D_B200_H
#define COMPTON2D_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2D_NMUMAX 32      /* general.pa:20 nmumax */
#define C2D_NLCMAX 16      /* >= general.pa:22 nphlcmax = 10 */

enum {
    C2D_OK = 0,
    C2D_ENODEV = 1,        /* no CUDA device / wrong architecture */
    C2D_EINVAL = 2,        /* bad argument or unsupported combination (Compton reflection without its tables, ...) */
    C2D_ECUDA = 3,         /* CUDA runtime error (message in c2d_last_error) */
    C2D_ECENSUS = 4,       /* census bank full: the reference's 'too many photons' STOP (imctrk2d.f:581-585) */
    C2D_EEVENTS = 5,       /* event buffer full (records beyond capacity were counted, not stored) */
    C2D_ESTATE = 6,        /* call order violated (step_begin before init, ...) */
    C2D_ECOMM = 7          /* collective failed / no transport (NCCL missing, rank mismatch) */
};

typedef struct c2d_ctx c2d_ctx;      /* opaque; one per GPU / per reference worker rank */

/* Static problem description: replaces setup_bcast (fp_mpi.f:12-319) + the grid part of
 * z_surf_bcast (surf_mpi.f:18-98).  Arrays are copied; the caller keeps ownership. */
typedef struct {
    int nz, nr;                       /* commonblock.f /zones/ */
    int n_vol, num_nt, nphfield, n_gg; /* general.pa:13-18 (400, 200, 400, 100 in the reference build) */
    int nphtotal, nph_lc, nmu;        /* spectrum groups, light-curve bands, angle bins (setup2d.f:148-173) */
    int pair_switch, cr_sent, spec_switch;   /* reader.f sentinels.  cr_sent (imcleak2d.f:107-165,216-273): 0 none, 1 lower surface
                                                (mirror where tbbl <= 0, Compton reflection where hot), 2 outer cylinder, 3 both,
                                                4 mirror at the lower surface; Compton reflection needs the tables below */
    int split1, split2, split3, spl3_trg;    /* reader.f:587-590 */
    int64_t census_capacity;          /* ucens (general.pa:7), photons */
    int64_t event_capacity;           /* escape records kept per step */
    int64_t collision_capacity;       /* pending Compton-collision tasks (scattering work queue) */
    uint32_t rseed;                   /* reader.f:466; Philox key word 0 (key word 1 = ncycle) */
    int sigma_mode;                   /* 0: comtot evaluated bin by bin for every history and zone (comtot2d.f:219-247)
                                         1: per-step table of comtot(zone, E): Thomson-branch bins as prefix sums,
                                            exact-branch bins as piecewise degree-7 Chebyshev series in ln E between
                                            the branch switch points (same values to ~1e-11, see c2d_comtot) */
    double rmin, zmin;
    const double *z;                  /* [nz]  upper faces */
    const double *r;                  /* [nr]  outer faces */
    const double *E_ph;               /* [n_vol]    volume2d.f:112-113 */
    const double *E_field;            /* [nphfield] setup2d.f:217-222 */
    const double *E_gg;               /* [n_gg]     setup2d.f:199-202 */
    const double *gnt;                /* [num_nt]   gamma-1 grid, nontherm2d.f:52-54 */
    const double *hu;                 /* [nphtotal+1] */
    const double *Elcmin, *Elcmax;    /* [nph_lc] */
    const double *mu;                 /* [nmu] */
    /* Compton-reflection tables (COMMON /reflection/, commonblock.f:76-77,185; computed once by Pref_calc / Wref_calc,
       ref_matrix.f, setup2d.f:40-41; shipped every step by xec_bcast, xec2d.f:313-318).  Needed for cr_sent = 2, 3 and for
       cr_sent = 1 with a hot lower boundary; n_ref = 0 / NULL otherwise. */
    int n_ref;                        /* general.pa:19 (500) */
    const double *E_ref;              /* [n_ref] keV */
    const double *P_ref, *W_abs;      /* [n_ref][n_ref] as they lie in COMMON: element (n_out, n_in) at (n_in-1) n_ref + (n_out-1) */
} c2d_dims;

/* Per-step inputs: what imcgen2d (imcgen2d.f) leaves in COMMON for the workers, i.e. the union of
 * vol_bcast (vol_mpi.f:39-45), the per-zone volume job (vol_mpi.f:100-119), z_surf_bcast
 * (surf_mpi.f:18-98) and xec_bcast (xec2d.f:293-318). */
typedef struct {
    int ncycle;
    double time, dt;                  /* dt(1) */
    const double *kappa_tot;          /* [zone][n_vol] */
    const double *eps_tot, *eps_th;   /* [zone][n_vol] emission CDFs */
    const double *f_nt, *Pnt;         /* [zone][num_nt] */
    const double *n_e, *tea, *f_pair; /* [zone] */
    const double *k_gg;               /* [zone][n_gg] (may be NULL when pair_switch = 0) */
    const int32_t *nsv;               /* [zone] volume photons to source */
    const double *ewsv;               /* [zone] their energy weight */
    const double *Eloss_th, *Eloss_tot, *zsurf;   /* [zone] */
    const double *tbbl;               /* [nr] lower-boundary temperature of the current time interval tbbl(k,ti) */
} c2d_step_in;

/* The step prologue on the device: instead of the finished tables of c2d_step_in, hand over the plasma
 * state and let the device run volume_em (volume2d.f:10-394, the terms that reach an output: synchrotron emissivity and
 * self-absorption from the live electron spectrum) plus the volume-source budget of imcgen2d (imcgen2d.f:280-283,291,
 * 330,405-459).  ep_switch = 0 (B_field as given); the > 10 nst bias cap (imcgen2d.f:491-517) cannot trigger for volume
 * sources alone.  c2d_dims.E_ph must be the grid volume_em generates (1e-10 (1e20)^(i/n_vol) keV). */
typedef struct {
    int ncycle;
    double time, dt;
    int64_t nst;                      /* reader.f nst: particles per photon cycle; nsv = 0.5 nst Eloss_tot / sum(Eloss_tot) */
    const double *f_nt, *Pnt;         /* [zone][num_nt] */
    const double *n_e, *tea, *B_field, *vol, *zsurf;   /* [zone] */
    const double *f_pair, *k_gg, *tbbl;               /* as in c2d_step_in; may be NULL */
} c2d_gen_in;

typedef struct {
    int32_t *nsv;                     /* [zone] (any pointer may be NULL) */
    double *ewsv, *Eloss_sy, *Eloss_th, *Eloss_tot;   /* [zone] */
    double Emiss_tot;
    int64_t n_volume;                 /* sum(nsv) */
} c2d_gen_out;

/* Surface sources of one step (optional; BASELINE configs[2]): what r_surf_bcast / z_surf_bcast and the ring job
 * messages carry (surf_mpi.f:18-98,136,337) plus the tables file_sp builds from an external spectrum file
 * (imcsurf2d_para.f:544-685; parsing and normalisation stay on the host, compton2d_b200.spectrum.file_sp).
 * A boundary ring emits nsurf* photons of weight ewsurf*; its energies are Planckian at tbb* when tbb* > 0
 * (planck2d.f) and drawn from spectrum table spec_* otherwise (file_sample, imcsurf2d_para.f:694-788). */
typedef struct {
    int star_switch;                  /* reader.f; both settings beam the lower source (imcsurf2d_para.f:452-460) */
    const int32_t *nsurfi, *nsurfo;   /* [nz] inner / outer cylinder */
    const int32_t *nsurfu, *nsurfl;   /* [nr] upper / lower disk */
    const double *ewsurfi, *ewsurfo, *ewsurfu, *ewsurfl;
    const double *tbbi, *tbbo;        /* [nz] tbbi(j,ti), tbbo(j,ti) */
    const double *tbbu, *tbbl;        /* [nr] tbbu(k,ti), tbbl(k,ti); tbbl also replaces c2d_step_in.tbbl (same COMMON array) */
    const int32_t *spec_i, *spec_o, *spec_u, *spec_l;   /* spectrum-table index per ring (may be NULL = table 0) */
    int n_spec, nfmax;                /* tables, entries per table (general.pa:17 nfmax = 500) */
    const int32_t *nfile;             /* [n_spec] */
    const double *E_file, *F_file, *P_file, *a1, *I_file;   /* [n_spec][nfmax] COMMON /insp/ */
} c2d_surf_in;

/* All tallies of one step, dense.  Filled by c2d_tally_read. */
typedef struct {
    double *edep, *prdep, *ecens;     /* [zone] */
    int64_t *npcen;                   /* [zone] */
    double *n_field;                  /* [zone][nphfield] */
    double *n_ph;                     /* [zone][n_gg] */
    double *E_IC;                     /* [num_nt] */
    int64_t *nelectron;               /* [num_nt+2], index = electron bin i (1-based as in nth2d) */
    double *fout;                     /* [C2D_NMUMAX][nphtotal]  escaping spectrum of THIS step (zeroed by c2d_step_begin) */
    double *edout;                    /* [C2D_NMUMAX][C2D_NLCMAX] */
    double *erlki, *erlko;            /* [nz] */
    double *erlku, *erlkl;            /* [nr] */
    double *Ed_in, *Ed_ref;           /* [nr] */
    int64_t counters[16];             /* see C2D_CNT_*; this rank's */
    double *fout_total;               /* [C2D_NMUMAX][nphtotal]  escaping spectrum summed over the run and over the ranks: the
                                         master's fout after graphics_collect (xec2d.f:389-394 sums the workers' never-zeroed
                                         copies; graphics2d.f:147 divides it by time + dt).  Advanced by c2d_tally_reduce. */
    int64_t counters_global[16];      /* counters summed over the ranks by c2d_tally_reduce (== counters on one rank) */
} c2d_tallies;

enum {
    C2D_CNT_NDXOUT = 0,      /* census photons written this step (ndxout) */
    C2D_CNT_CMCOUNT1 = 1,    /* Compton collisions (imctrk2d.f:615) */
    C2D_CNT_CMCOUNT2 = 2,    /* scattered children tracked (imctrk2d.f:661,677) */
    C2D_CNT_LKCOUNT = 3,     /* imcleak calls (imctrk2d.f:517) */
    C2D_CNT_NEVENTS = 4,     /* escape records produced */
    C2D_CNT_HISTORIES = 5,   /* top-level imctrk2d(-1) calls */
    C2D_CNT_TRIALS = 6,      /* split1 trial flights evaluated */
    C2D_CNT_MULTIPLE = 7,    /* 'multiple scattering' events (imctrk2d.f:614) */
    C2D_CNT_KILLED = 8,      /* flights ended by ew <= wtmin (imctrk2d.f:470) */
    C2D_CNT_TEAMS = 9,       /* children finished by k_children_hard + 3rd-split sub-children finished by k_split3_ultra */
    C2D_CNT_LITERAL = 10,    /* trial flights executed literally (collision candidates) */
    C2D_CNT_QPEAK = 11,      /* peak number of queued collision tasks */
    C2D_CNT_FATAL = 12,      /* device-side fatal conditions (0 on success) */
    C2D_CNT_SEGMENTS = 13,   /* zone crossings of the collision-free rays */
    C2D_CNT_S3PASSES = 14,   /* passes of the 3rd-split rejection loop (imctrk2d.f:642-656) up to each accepted one */
    C2D_CNT_S3HARD = 15      /* 3rd-split sub-children finished by the block-team kernel */
};

/* Escape record: the 7 doubles of imcleak2d.f:171,181 (t_bound,xnu,ew,rpre,zpre,wmu,phi). */
typedef struct { double t_bound, xnu, ew, rpre, zpre, wmu, phi; } c2d_event;

/* ---- life cycle --------------------------------------------------------------------------- */
int c2d_create(c2d_ctx **ctx, int device);                 /* binds the context to a CUDA device */
int c2d_init(c2d_ctx *ctx, const c2d_dims *dims);          /* after setup / read_record */
int c2d_destroy(c2d_ctx *ctx);
const char *c2d_last_error(void);
int c2d_version(void);

/* ---- one IMC time step: same order as xec (xec2d.f:141-193) --------------------------------- */
int c2d_step_begin(c2d_ctx *ctx, const c2d_step_in *in);   /* replaces imcgen2d's worker branch + the broadcasts */
int c2d_step_begin_gen(c2d_ctx *ctx, const c2d_gen_in *in, c2d_gen_out *out);
        /* == imcgen2d for the volume sources, computed on the device; replaces c2d_step_begin for the step */
int c2d_step_tables_read(c2d_ctx *ctx, double *kappa_tot, double *eps_tot, double *eps_th);
        /* D2H copy of the step's [zone][n_vol] tables (for COMMON copy-back / checks); pointers may be NULL */
int c2d_imcfield(c2d_ctx *ctx);                            /* == field_calc (imcfield2d.f:57-148) on the resident bank */
int c2d_imcvol(c2d_ctx *ctx, int64_t first, int64_t count, int64_t stride);
        /* == sum over zones of vol_calc (imcvol2d_para.f:90-414); this context tracks the volume
           histories with global index first, first+stride, ... (count of them; count<0 = all that remain),
           which is how histories shard over GPUs */
int c2d_surf_begin(c2d_ctx *ctx, const c2d_surf_in *in);   /* after c2d_step_begin, when the step has surface sources */
int c2d_imcsurf(c2d_ctx *ctx, int64_t first, int64_t count, int64_t stride);
        /* == imcsurf2d: sum over rings of z_surf_calc + r_surf_calc (imcsurf2d_para.f:228-534); sharded like c2d_imcvol */
int c2d_sync(c2d_ctx *ctx);                                /* drains the scattering queue, waits for the device */
int c2d_tally_read(c2d_ctx *ctx, c2d_tallies *out);        /* D2H copy of this context's (or the reduced) tallies */
int c2d_tally_reduce(c2d_ctx *ctx, c2d_tallies *out);
        /* the end-of-step reduction: c2d_sync, then ONE all-reduce(sum) of the packed FP64 tally buffer and one of the packed
           int64 buffer over the ranks of c2d_comm_init, in place in HBM -- cens_add_up + E_add_up + xec_add + graphics_collect
           (update2d.f:1931-2083, xec2d.f:325-394) -- then fout_total += fout, then c2d_tally_read (out may be NULL).  Once per
           step (a second call in the same step only reads).  Collective: every rank must call it. */
int c2d_finalize(c2d_ctx *ctx);                             /* == c2d_destroy */
int c2d_tally_device(c2d_ctx *ctx, double **fp64_buf, int64_t *n_fp64, int64_t **i64_buf, int64_t *n_i64);
        /* device pointers of the packed tally buffers, for the per-step all-reduce
           (cens_add_up + E_add_up + xec_add + graphics_collect: update2d.f:1931-2083, xec2d.f:325-394) */

int c2d_comtot(c2d_ctx *ctx, int64_t n, const int32_t *zone, const double *xnu, int mode, double *comac);
        /* == comtot(j,k,xnuc,1,comac,enexc) (comtot2d.f:1-334) for n (0-based zone, xnu [keV]) pairs with the tables
           of the current step; mode 0 = bin by bin, 1 = the sigma_mode-1 table (context must have sigma_mode 1) */

/* ---- ranks: one context per GPU, NCCL over NVLink inside the library -----------------------------------
 * The reference's MPI calls on this path become: c2d_tally_reduce (all-reduce), the all-gather of the electron
 * state after the Fokker-Planck update (c2d_step / c2d_update_fp_resident; FP_recv_result + FP_end_bcast, fp_mpi.f:926,
 * 1017-1079) and the re-dealing of 3rd-split work between the ranks (c2d_sync).  NCCL is bound at run time (the libnccl.so.2
 * the process already holds, else the system's).  Rank 0 makes an id, the host ships its 128 bytes to the other ranks with
 * whatever it has (MPI_BCAST from the Fortran host, torch.distributed from Python), every rank calls c2d_comm_init. */
#define C2D_COMM_ID_BYTES 128
int c2d_comm_unique_id(void *id128);                              /* ncclGetUniqueId */
int c2d_comm_init(c2d_ctx *ctx, const void *id128, int rank, int world);   /* collective; world = 1 needs no id */
/* A host that brings its own transport (the reference's MPI on CUDA-aware buffers; gloo in the CPU tests of the host logic).
 * Pointers are DEVICE pointers, the call must be ordered after the work already queued on `cuda_stream` and complete (or be
 * stream-ordered) before it returns.  kind: 0 = FP64 sum, 1 = int64 sum, 2 = FP64 max; send may equal recv. */
typedef struct {
    void *user;
    int (*allreduce)(void *user, const void *send, void *recv, int64_t n, int kind, void *cuda_stream);
    int (*allgather)(void *user, const void *send, void *recv, int64_t bytes_per_rank, void *cuda_stream);
} c2d_comm_ext;
int c2d_comm_init_ext(c2d_ctx *ctx, int rank, int world, const c2d_comm_ext *transport);
int c2d_comm_info(c2d_ctx *ctx, int *rank, int *world);
int c2d_comm_destroy(c2d_ctx *ctx);

/* ---- census bank: write_record / read_record contract (census2d.f:1-76) ---------------------- */
int c2d_census_count(c2d_ctx *ctx, int64_t *n);
int c2d_census_export(c2d_ctx *ctx, double *dbuf, int32_t *ibuf, uint64_t *sid, int64_t cap, int64_t *n);
        /* dbuf[6n] = rpre,zpre,wmu,phi,ew,xnu ; ibuf[6n] = jgpsp,jgplc,jgpmu,jph,kph,seed (imctrk2d.f:566-580);
           sid[n] = replay stream id (may be NULL) */
int c2d_census_import(c2d_ctx *ctx, const double *dbuf, const int32_t *ibuf, const uint64_t *sid, int64_t n);
int c2d_events_drain(c2d_ctx *ctx, c2d_event *buf, int64_t cap, int64_t *n);

/* ---- the device-owned state in the reference's own file formats ---------------------
 * What write_record / read_record keep of a worker between runs is its census (write_record.f:432-437 ->
 * p###_census.dat); what a worker writes while it runs is its event file (unit 1000+myid, imcleak2d.f:171,181).
 * Both are text with Fortran E14.7 / I5 fields; these entry points produce and read exactly those files, so a run can be
 * handed over between the reference and this library in either direction, and plcm / pspt keep working on the events. */
int c2d_census_write(c2d_ctx *ctx, const char *path, int64_t *n_written);
        /* == write_cens(ndxout, unit) (census2d.f:1-36): per photon `6e14.7` then `6i5` */
int c2d_census_read(c2d_ctx *ctx, const char *path, int64_t entries);
        /* == read_cens(entries, unit) (census2d.f:41-73) into the resident bank; entries = ndxout of p###_misc.dat */
int c2d_events_write(c2d_ctx *ctx, const char *path, int append, int64_t *n_written);
        /* the step's escape records as imcleak writes them: format 105 = 6(e14.7,1x),e14.7 (imcleak2d.f:181) */
void c2d_format_e14_7(double v, char *out14);
        /* the E14.7 edit descriptor as ifort prints it (14 characters, no terminator) */

/* ---- event-list reduction -------------------------------------------------------
 * The reference turns its p###_evb.dat event files (one 105-byte text line per escaping photon, imcleak2d.f:181) into
 * light curves and SEDs with two stand-alone C programs that share one per-photon transform (Doppler boost with the bulk
 * Lorentz factor, light-travel-time correction to the plane through r = rmax) and differ in their bin rules:
 *   postprocessing/plcm.c:384-450   L(t_obs, mu_obs, band) with <ew>, <ew^2>;  mu bins [mu0, mu1), EVERY band holding E
 *   postprocessing/pspt.c:249-289   L(E_obs, t_obs);  one CLOSED mu window [mu0, mu1], only the FIRST band holding E
 * These entry points do the same reduction on the escape records while they are still in HBM.  The accumulators
 * [time bin][mu bin][band] persist across steps (as the event files do) until the next c2d_lc_begin.
 * compton2d_b200/lightcurve.py builds the tables from the programs' own answer files (mrk421_lc.input, ...). */
typedef struct {
    double gam_bulk, rmax, t_offset;  /* plcm.c:104-130 prompts; t_offset is subtracted from the arrival time (plcm.c:412) */
    int n_t, n_mu, n_bands;
    int mu_closed;                    /* 0: mu0 <= mu < mu1 (plcm.c:423), 1: mu0 <= mu <= mu1 (pspt.c:274) */
    int first_band_only;              /* 0: plcm.c:441-449, 1: pspt.c:278-280 */
    const double *t_lo, *t_hi;        /* [n_t]   t_lo <= t < t_hi, first match; both nondecreasing (plcm.c:236-240, pspt.c:145-148) */
    const double *mu0, *mu1;          /* [n_mu]  first match */
    const double *E0, *E1;            /* [n_bands]  E0 <= E < E1 */
} c2d_lc_config;

int c2d_lc_select(c2d_ctx *ctx, int slot);
        /* a context holds up to 4 independent reducers (e.g. plcm's light curves and pspt's SEDs of the same run); c2d_lc_begin,
           _read, _device and _bench act on the selected one (default 0), the _accumulate calls feed every reducer that was begun */
int c2d_lc_begin(c2d_ctx *ctx, const c2d_lc_config *cfg);  /* sets the bin tables, zeroes the accumulators */
int c2d_lc_accumulate(c2d_ctx *ctx);
        /* bins the escape records of the current step (those c2d_events_drain would return); call once per step after
           the tracking calls -- it drains the scattering queue like c2d_sync */
int c2d_lc_accumulate_events(c2d_ctx *ctx, const c2d_event *ev, int64_t n);
        /* the same for n records in host memory (event files of an earlier run / of the reference) */
int c2d_lc_read(c2d_ctx *ctx, double *sum_ew, double *sum_ew2, int64_t *count);
        /* [n_t][n_mu][n_bands] each (any pointer may be NULL): sum of ew, of ew^2, photon count -- plcm's F, F2,
           particle_sum before its normalisation (plcm.c:470-496); pspt's F, particle_sum (pspt.c:286-287) */
int c2d_lc_device(c2d_ctx *ctx, double **sums, int64_t **count, int64_t *cells);
        /* device pointers of the accumulators ([2][cells] FP64, [cells] int64) for the reduction over GPUs */
int c2d_lc_bench(c2d_ctx *ctx, int64_t n_events, float *ms, int64_t *bytes);
        /* k_lc_reduce on n generated records with the current tables; CUDA-event time and algorithmic bytes (56 / record) */

/* ---- Fokker-Planck update (update2d.f:337-1741 FP_calc, batched over zones) -------------------
 * Inputs are what the per-zone job message carries (fp_mpi.f:612-700) plus the broadcast scalars
 * (fp_mpi.f:579-600); outputs what FP_send_result returns (fp_mpi.f:860-930) and FP_end_bcast
 * re-broadcasts (fp_mpi.f:1017-1079).  Only the zones first_zone, first_zone+zone_stride, ... (n_zones of
 * them; n_zones < 0 = all that remain) are updated -- the multi-GPU partition; the other zones' output
 * entries are left untouched.  Pair terms (n_pos, dn_pp, dne_pa, dnp_pa) are taken as zero: they are inert
 * in the reference's MPI build. */
typedef struct {
    int first_zone, n_zones, zone_stride;       /* 0-based zone ids */
    int inj_switch, inj_dis, g2var_switch, pick_sw, cf_sentinel, ncycle;
    double time, dt, lnL;
    double inj_g1, inj_g2, inj_p, inj_t, inj_gg, inj_sigma, inj_L, inj_v, pick_rate;
    double r_esc, r_acc;
    double r_flare, z_flare, t_flare, sigma_r, sigma_z, sigma_t, flare_amp;
    const double *f_nt, *Pnt;                   /* [zone][num_nt] */
    const double *n_field;                      /* [zone][nphfield]; NULL = this step's (reduced) device tally */
    const double *ecens;                        /* [zone]; NULL = device tally */
    const double *n_e, *B_field, *tea, *tna, *vol, *f_pair;   /* [zone] */
    const double *Eloss_sy, *ec_old, *turb_lev; /* [zone] */
    const double *F_IC;                         /* [num_nt][nphfield] (icloss2d.f); NULL = table built on device by c2d_ic_loss */
} c2d_fp_in;

typedef struct {
    double *f_nt, *Pnt;                         /* [zone][num_nt] */
    double *n_e, *Te_new, *gmin, *gmax, *p_nth, *amxwl, *f_pair;   /* [zone] (p_nth, gmin, gmax: in/out) */
    int32_t *fp_steps;                          /* [zone] sub-steps taken */
    double *zone_E_old, *zone_E_new, *zone_hr, *zone_hr_st, *zone_dT;   /* [zone] per-zone terms of the sums below (may be NULL) */
    double dT_max, E_tot_old, E_tot_new, hr_total, hr_st_total;   /* over this call's zones, summed in zone order */
    double dr, dz, f_t_implicit;                /* COMMON /d_update/ leftovers of the last zone (update2d.f:233-243) */
} c2d_fp_out;

int c2d_ic_loss(c2d_ctx *ctx, double *F_IC_out);           /* == IC_loss (icloss2d.f:1-125); F_IC_out may be NULL */
int c2d_update_fp(c2d_ctx *ctx, const c2d_fp_in *in, c2d_fp_out *out);
int c2d_gamma_bar(c2d_ctx *ctx, int64_t n, const double *Theta, double *gbar, double *gbar_table);
        /* == gamma_bar(Theta) (volume2d.f:572-594: <gamma> of a Maxwellian, McDonald quadratures above Theta = 0.2) for n
           temperatures; gbar_table (may be NULL) = the interpolation table that predicts the lattice point of FP_calc's
           temperature walk (update2d.f:1446-1463), negated where Theta lies beyond the table */

/* ---- the closed time step on a device-resident plasma state ---------------------------------------------
 * One iteration of xec's loop (xec2d.f:100-110 / 141-193): imcgen2d -> imcfield2d -> imcvol2d -> imcsurf2d -> update ->
 * graphics_collect, with the plasma the master keeps in COMMON (and re-broadcasts every step: vol_bcast, z_surf_bcast,
 * FP_bcast, FP_end_bcast) resident in HBM on every rank instead.  See csrc/c2d_step.cu for the exact order. */
typedef struct {                      /* per-zone plasma: inputm.dat (reader.f:608-657) + what FP_end_bcast carries (fp_mpi.f:1017-1079) */
    const double *f_nt, *Pnt;         /* [zone][num_nt] (P_nontherm, nontherm2d.f:2-150) */
    const double *n_e, *tea, *tna, *B_field;          /* [zone]; B_field as given (ep_switch = 0, imcgen2d.f:213-238) */
    const double *f_pair;             /* [zone] or NULL */
    const double *gmin, *gmax, *p_nth, *amxwl;        /* [zone] power-law fit carried between steps (update2d.f:1656-1738) */
    const double *turb_lev;           /* [zone] or NULL */
    const double *vol, *zsurf;        /* [zone] setup2d.f:80-114 */
    const double *k_gg;               /* [zone][n_gg] or NULL */
} c2d_plasma;                         /* c2d_plasma_set: NULL members keep their value (the first call needs all but f_pair, turb_lev, k_gg) */

typedef struct {                      /* c2d_plasma_get: any pointer may be NULL */
    double *f_nt, *Pnt;
    double *n_e, *tea, *f_pair, *B_field, *gmin, *gmax, *p_nth, *amxwl;
    double *Te_new;                   /* unclamped temperature of the last FP update (update2d.f:1441-1469) */
    double *Eloss_sy, *ec_old;        /* this step's synchrotron loss (imcgen2d.f:280-283) and the census energy carried in (l.90-96) */
} c2d_plasma_out;

typedef struct {                      /* surface sources of a run: the boundary block of input.dat for the current time interval */
    int star_switch;
    double Rstar, dist_star;          /* imcgen2d.f:163-164 (star_switch = 1) */
    double t0;                        /* t0(ti): file-fed rings switch on at time + dt/2 >= t0 (imcgen2d.f:127) */
    const double *tbbi, *tbbo;        /* [nz] */
    const double *tbbu, *tbbl;        /* [nr] */
    const int32_t *spec_i, *spec_o, *spec_u, *spec_l;   /* spectrum table per ring (NULL = table 0) */
    int n_spec, nfmax;
    const int32_t *nfile;             /* [n_spec] */
    const double *int_file;           /* [n_spec] integral of each table (file_sp, imcsurf2d_para.f:640-650) */
    const double *E_file, *F_file, *P_file, *a1, *I_file;   /* [n_spec][nfmax] */
} c2d_surf_run;

typedef struct {                      /* run constants of input.dat the step needs (reader.f:157-597) */
    int64_t nst;                      /* particles per photon cycle */
    int T_const;                      /* 1: electron spectrum frozen, update is skipped (xec2d.f:87) */
    int quirk_cycle0;                 /* 1: cycle 0 as the reference's MPI workers run it: dt = 0 and no energy groups,
                                         every source photon is banked unflown at its birth point with bins 0 */
    int inj_switch, inj_dis, g2var_switch, pick_sw, cf_sentinel;
    double inj_g1, inj_g2, inj_p, inj_t, inj_gg, inj_sigma, inj_L, inj_v, pick_rate;
    double r_esc, r_acc;
    double r_flare, z_flare, t_flare, sigma_r, sigma_z, sigma_t, flare_amp;
    const c2d_surf_run *surf;         /* NULL: no surface sources (the shipped input: every tbb = -0.d0) */
} c2d_run;

typedef struct {                      /* what the reference prints to log.txt / stdout for the cycle, summed over the ranks */
    int ncycle;
    double time, dt;
    int64_t n_census_in;              /* census photons this RANK re-tracked */
    int64_t n_volume, n_surface;      /* source photons of the cycle (all ranks) */
    int64_t histories, ndxout, nevents, cmcount1, cmcount2;   /* all ranks */
    double bingo, Emiss_tot, fbias;   /* imcgen2d.f:405-421 'Total ingoing energy', l.491-517 */
    double luminosity;                /* E_add_up: sum(erlk*) / dt (update2d.f:2068) */
    int fp_ran, fp_steps_min, fp_steps_max;
    double E_tot_old, E_tot_new, dE_fraction, dT_max, dt_new, hr_total, hr_st_total;   /* update2d.f:206-243 */
    double ms[16];                    /* CUDA-event ms per kernel family since timing was enabled (c2d_kernel_time_ms) */
} c2d_step_report;

int c2d_plasma_set(c2d_ctx *ctx, const c2d_plasma *p);
int c2d_plasma_get(c2d_ctx *ctx, c2d_plasma_out *p);
int c2d_step(c2d_ctx *ctx, const c2d_run *run, int ncycle, double time, double dt, c2d_step_report *report);
        /* collective when c2d_comm_init was called with world > 1.  The host advances the clock as xec does
           (xec2d.f:94-103): if ncycle > 0 then time += dt; ncycle += 1. */

/* ---- the COMMON-block side: what a Fortran shim hands over --------------------------------------------
 * Addresses of the reference's COMMON members as they lie (column-major, padded compile-time extents of general.pa:7-31;
 * a(jmax,kmax): element (j,k) at (k-1) jmax + (j-1); a(nbin,jmax,kmax) and a(jmax,kmax,nbin) likewise) -- `c_loc(edep)` etc.
 * from integration/c2d_shim.f90.  Members that are NULL are skipped.  The three calls pack into the dense layout above, call
 * the entry points above and scatter the results back where the reference's own subroutines leave them. */
typedef struct {
    int jmax, kmax, ntmax;                    /* general.pa:10-12 (99, 99, 100) */
    int nmumax, nphomax, nphlcmax;            /* general.pa:20-22 (32, 128, 10) */
    int ti;                                   /* current time interval of the boundary temperatures (1-based; imcgen2d.f:113-119) */
    /* step inputs: what imcgen2d leaves for the tracker */
    const double *kappa_tot, *eps_tot, *eps_th;   /* (n_vol, jmax, kmax)   /vol_em/ */
    double *f_nt, *Pnt;                       /* (jmax, kmax, num_nt)  /Pnth/           (in; out of c2d_common_update) */
    double *n_e, *tea;                        /* (jmax, kmax)          /zone_quantities/ (in; out of c2d_common_update) */
    const double *tna, *B_field;              /* (jmax, kmax) */
    double *f_pair;                           /* (jmax, kmax) */
    const double *k_gg;                       /* (n_gg, jmax, kmax) */
    const int32_t *nsv;                       /* (jmax, kmax) */
    const double *ewsv, *Eloss_th, *Eloss_tot, *zsurf;   /* (jmax, kmax) */
    const double *tbbl;                       /* (kmax, ntmax): column ti is used */
    /* tallies: where cens_add_up / E_add_up / xec_add / graphics_collect leave them on the master */
    double *edep, *prdep, *ecens;             /* (jmax, kmax) */
    int32_t *npcen;                           /* (jmax, kmax) */
    double *n_field;                          /* (nphfield, jmax, kmax) */
    double *n_ph;                             /* (n_gg, jmax, kmax) */
    double *E_IC;                             /* (num_nt) */
    int32_t *nelectron;                       /* (num_nt) */
    double *fout;                             /* (nmumax, nphomax): the run-long sum (xec2d.f:389-394) */
    double *edout;                            /* (nmumax, nphlcmax) */
    double *erlki, *erlko;                    /* (jmax) */
    double *erlku, *erlkl, *Ed_in, *Ed_ref;   /* (kmax) */
    int32_t *ndxout, *cmcount1, *cmcount2, *lkcount;   /* scalars of /counters/ etc. */
    /* Fokker-Planck */
    const double *vol, *Eloss_sy, *ec_old, *turb_lev;    /* (jmax, kmax) */
    double *gmin, *gmax, *p_nth, *amxwl, *Te_new;        /* (jmax, kmax)  /nontherm/, /d_update/ */
    double *dT_max, *E_tot_old, *E_tot_new, *hr_total, *hr_st_total, *dr, *dz, *f_t_implicit;   /* scalars */
} c2d_common;

int c2d_common_step_begin(c2d_ctx *ctx, const c2d_common *cm, int ncycle, double time, double dt);
        /* pack + c2d_step_begin: called by the shim's imcfield2d (the first tracking subroutine xec calls, xec2d.f:71,167) */
int c2d_common_collect(c2d_ctx *ctx, c2d_common *cm);
        /* c2d_tally_reduce + scatter into COMMON: the shim's graphics_collect (xec2d.f:91,183; also stands in for xec_add,
           cens_add_up and E_add_up, whose sums it delivers) */
int c2d_common_update(c2d_ctx *ctx, c2d_common *cm, const c2d_fp_in *scalars);
        /* the shim's update (update2d.f:7-334): FP_calc for every zone from the reduced tallies in HBM; scalars = the
           non-pointer members of c2d_fp_in, i.e. what FP_bcast ships (fp_mpi.f:579-600); f_nt, Pnt, n_e, tea, gmin, gmax, p_nth, amxwl, Te_new
           and the sums of E_add_up land in COMMON */

/* ---- COMMON <-> dense packing helpers (host code) ------------------------------------------- */
void c2d_pack_bin_first(const double *common, int nbin_pad, int jmax, int kmax, int nbin, int nz, int nr, double *dense);
void c2d_pack_bin_last(const double *common, int jmax, int kmax, int nbin, int nz, int nr, double *dense);
void c2d_unpack_bin_first(const double *dense, int nbin_pad, int jmax, int kmax, int nbin, int nz, int nr, double *common);
void c2d_unpack_bin_last(const double *dense, int jmax, int kmax, int nbin, int nz, int nr, double *common);

int c2d_pref_calc(c2d_ctx *ctx, int n_ref, double *E_ref, double *P_ref);
        /* == Pref_calc (ref_matrix.f:7-89): E_ref[n_ref] and the reflection Green's function P_ref(n_out, n_in) as COMMON
           /reflection/ holds it (column n_in contiguous), built on the device -> host arrays the caller hands to c2d_init in
           c2d_dims when cr_sent = 2 / 4 or a hot lower boundary reflects (imcleak2d.f:124-158).  The ctx only needs
           c2d_create.  W_abs (Wref_calc, ref_matrix.f:101-499: photo-absorption of a partially ionised disk, 400 lines of
           atomic DATA) is not rebuilt: it stays an input. */

/* ---- self-tests of the device arithmetic --------------------------------------------------- */
int c2d_selftest_gap(c2d_ctx *ctx, int64_t n, uint64_t seed, int64_t *mismatches, int64_t *slow_path);
        /* candidate gaps of the trial uniforms, floor(ln U / ln(255/256)): the single-precision fast path of the tracker
           against the double-precision expression it stands for, over 2 n uniforms of Philox stream `seed`;
           mismatches must come back 0; slow_path = how many uniforms fell back to the double-precision log */

/* ---- measurement hooks (bench.py) ------------------------------------------------------------ */
int c2d_kernel_time_ms(c2d_ctx *ctx, int which, float *ms, int64_t *launches);
        /* CUDA-event time per kernel family on the library's stream.  which >= 0: read family (0 k_history, 1 k_children,
           2 k_split3_attempts/_hard, 3 FP, 4 sigma table, 6 k_split3_flights, 7 bracket, 8 tally all-reduce, 9 FP all-gather,
           10 cascade exchange (incl. the wait for the other ranks), 11 census-bank sort, 5 block / grid teams of k_children,
           12 k_split3_hard, 13 k_split3_ultra);
           -1 enable, -5 disable, -2 reset, -3 / -4 record the start / stop event of a bracket (accumulated in family 7) */

#ifdef __cplusplus
}
#endif
#endif
