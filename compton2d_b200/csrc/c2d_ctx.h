// c2d_ctx.h -- the context behind the opaque c2d_ctx handle (host side)
#pragma once
#include <vector>

#include "../../include/compton2d_b200.h"
#include "c2d_host.h"

namespace c2d {

int fail(int code, const char *fmt, ...);
struct Ctx;
int check_events(Ctx *c);      // C2D_EEVENTS when the step produced more escape records than event_capacity

struct FpState;
struct LcState;
struct Comm;
struct FpParams;

// collectives (c2d_comm.cu).  kind: 0 = FP64 sum, 1 = int64 sum, 2 = FP64 max; device pointers, library stream
struct Ctx;
int comm_world(const Ctx *c);
int comm_rank(const Ctx *c);
void comm_free(Ctx *c);
int comm_allreduce(Ctx *c, const void *send, void *recv, size_t n, int kind);
int comm_allgather(Ctx *c, const void *send, void *recv, size_t bytes_per_rank);

// rows of the resident plasma blob Ctx::d_plasma (each padded to an even number of zones)
enum { PL_TNA = 0, PL_B, PL_GMIN, PL_GMAX, PL_PNTH, PL_AMXWL, PL_TURB, PL_VOL, PL_ELOSS_SY, PL_EC_OLD, PL_TE_NEW, PL_ROWS };
constexpr int kFpRecExtra = 16;        // per-zone scalars of the FP exchange record (c2d_fp.cu)
int fp_resident(Ctx *c, const c2d_fp_in *in, double *zinfo_host);   // FP on the resident state + all-gather; zinfo [zone][16]

struct GenBudget { std::vector<int32_t> nsv; std::vector<double> ew, e_sy, e_th; double Emiss_tot = 0.0; };
int latch_census(Ctx *c);
int gen_core(Ctx *c, double dt, int64_t nst, const double *d_B, const double *h_ne, const double *h_B, const double *h_vol,
             const double *h_zsurf, GenBudget &g);
int budget_upload(Ctx *c, const GenBudget &g);
int begin_common(Ctx *c, int ncycle, double time, double dt);
int bank_sort(Ctx *c, long long n);    // -> 1 sorted into the other bank, 0 skipped, -1 error (c2d_bank.cu)
int launch_history_chunks(Ctx *c, int src, int64_t first, int64_t count, int64_t stride);

struct Ctx {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    c2d_dims dims{};
    int nzone = 0;
    bool inited = false, step_open = false, field_done = false, step_tracked = false, cascade_dirty = false;
    int64_t census_cap = 0, event_cap = 0, q_cap = 0, n_census = 0, vol_next = 0, histories = 0;
    long long q_peak = 0;
    // device memory
    double *d_ref = nullptr;               // [E_ref | P_ref | W_abs] Compton-reflection tables (static)
    double *d_static = nullptr, *d_step = nullptr, *d_f64 = nullptr, *d_events = nullptr;
    int64_t *d_prefix = nullptr;
    void *d_surf = nullptr;                // surface-source blob of the current step
    size_t surf_bytes = 0;
    int64_t surf_total = 0;
    bool surf_open = false;
    unsigned long long *d_i64 = nullptr, *d_qcount = nullptr;
    CollTask *d_q[2] = {nullptr, nullptr}, *d_q3 = nullptr;
    FlightTask *d_flights = nullptr;       // accepted 3rd-split sub-children of one chunk of split3 tasks
    int64_t flight_cap = 0;
    unsigned *d_hard = nullptr;            // sub-children the warp phase of the 3rd split left unresolved
    GridItem *d_ultra = nullptr;           // ... and those the block teams of k_split3_hard gave up on (k_split3_ultra: the whole grid)
    GridItem *d_cultra = nullptr;          // ... and those the block teams gave up on (k_children_ultra)
    unsigned *d_chard = nullptr;           // children whose K-N rejection loop outlasted kn_cap trips (k_children_hard)
    size_t chard_cap = 0;
    unsigned long long *d_casc = nullptr;  // [0] hard children [1] their fetch cursor [2] hard sub-children [3] fetch cursor [4] ultra list
                                           // [8..10] k_split3_ultra's winner slots
    int kn_cap = 64;                       // trips of the K-N rejection loop a thread makes before a block team takes the child (C2D_KN_CAP)
    // block teams -> grid team: after *_hard_rounds rounds whatever happens (tests force the path with 0), or after *_drain_rounds
    // rounds once half the blocks have run out of work (C2D_KN_HARD_ROUNDS, C2D_KN_DRAIN_ROUNDS, C2D_SPLIT3_HARD_ROUNDS, C2D_SPLIT3_DRAIN_ROUNDS)
    int kn_hard_rounds = 1 << 22, kn_drain_rounds = 1024;          // x 256 trips
    int split3_drain_rounds = 128;                                 // x 512 passes
    int split3_hard_rounds = 1 << 21;          // block-team rounds before a sub-child goes to the grid team (C2D_SPLIT3_HARD_ROUNDS)
    int split3_rounds = 64;                // warp rounds before a sub-child goes to k_split3_hard (C2D_SPLIT3_ROUNDS; measured 8..128)
    Bank bank[2]{};
    int cur_out = 0, q_cur = 0;
    // census bank sort (c2d_bank.cu) and the absorption-draw queue of k_history (c2d_track.cu)
    bool sort_bank = true;                 // C2D_SORT_BANK=0 keeps the bank in arrival order (A/B measurements)
    bool bank_sorted = false;              // the census lives sorted in bank[cur_out ^ 1] (c2d_tally_reduce did it)
    unsigned long long *d_sortkeys = nullptr;
    unsigned long long *d_cnt3 = nullptr;  // cascade counts: [mine x3 | all ranks]
    CollTask *d_q3_all = nullptr;          // all-gathered 3rd-split tasks
    size_t q3all_cap = 0;
    AbsTask *d_qabs = nullptr;
    int *d_litpool = nullptr;
    int64_t qabs_cap = 0, lit_cap = 0;
    // offsets (in doubles) into d_step / d_f64, (in int64) into d_i64
    size_t o_kappa = 0, o_epstot = 0, o_epsth = 0, o_fnt = 0, o_pnt = 0, o_ne = 0, o_fpair = 0, o_kgg = 0, o_tbbl = 0,
           o_ewsv = 0, o_elth = 0, o_eltot = 0, o_zsurf = 0, o_tea = 0, n_step = 0;
    size_t t_edep = 0, t_prdep = 0, t_ecens = 0, t_nfield = 0, t_nph = 0, t_EIC = 0, t_fout = 0, t_edout = 0, t_erlki = 0,
           t_erlko = 0, t_erlku = 0, t_erlkl = 0, t_Edin = 0, t_Edref = 0, n_f64 = 0;
    size_t i_npcen = 0, i_nel = 0, i_cnt = 0, n_i64 = 0;
    Params P{};
    // collectives + what only exists once per rank
    Comm *comm = nullptr;
    double *d_fout_total = nullptr;        // [C2D_NMUMAX][nphtotal] escaping spectrum summed over the run (xec2d.f:389-394)
    unsigned long long *d_cnt_global = nullptr;   // [CNT_N] counters summed over the ranks by c2d_tally_reduce
    bool step_reduced = false;             // c2d_tally_reduce already ran in this step
    // resident plasma state (c2d_step.cu): f_nt, Pnt, n_e, f_pair, tea, zsurf live in d_step (the tracker reads them there)
    double *d_plasma = nullptr;            // [tna | B_field | gmin | gmax | p_nth | amxwl | turb_lev | vol | Eloss_sy | ec_old | Te_new] x nzone
    bool plasma_set = false;
    std::vector<double> h_ne, h_B, h_vol, h_zsurf, h_tna;   // host mirrors of the per-zone scalars the budget arithmetic needs
    double *d_fpx = nullptr;               // FP exchange buffers: [out arrays | send | recv]
    size_t fpx_cap = 0;
    std::vector<int64_t> h_prefix;
    std::vector<double> h_f64, h_gnt, h_Efield, h_z, h_r, h_Eph;
    double *d_gen = nullptr;               // scratch of c2d_step_begin_gen
    size_t gen_cap = 0;
    std::vector<unsigned long long> h_i64;
    // sigma(zone, E) table (sigma_mode 1, c2d_sigma.cu)
    int sig_nodes = 0;
    double *d_sig_static = nullptr, *d_sig_step = nullptr;   // [lo | iw | xs | M | Kn] and [W | AB | T | coef]
    int *d_sig_ms = nullptr;
    double *d_sig_W = nullptr, *d_sig_AB = nullptr, *d_sig_T = nullptr, *d_sig_coef = nullptr, *d_sig_Kn = nullptr, *d_sig_M = nullptr;
    std::vector<double> h_gam0, h_opb;
    bool sigma_ready = false;
    uint32_t *d_nt_guide = nullptr;        // nth2d's bisection brackets, rebuilt with the step tables (k_nt_guide)
    int sigma_setup();
    int sigma_build();
    int sigma_ensure();                    // builds the table on the step's first use
    // Fokker-Planck state (c2d_fp.cu)
    FpState *fp = nullptr;
    void fp_free();
    // event-list reduction (c2d_lc.cu)
    static constexpr int kLcSlots = 4;     // independent reducers (light curves and SEDs of one run at once)
    LcState *lc_slot[kLcSlots] = {nullptr, nullptr, nullptr, nullptr};
    int lc_active = 0;
    LcState *&lc_cur() { return lc_slot[lc_active]; }
    double *d_lc_stage = nullptr;          // host-provided records, shared by the slots
    size_t lc_stage_cap = 0;
    void lc_free(int slot);
    void lc_free_all();
    // CUDA-event timing per kernel family: 0 k_history, 1 k_children, 2 k_split3_attempts, 3 FP, 4 sigma table, 5 k_children_hard / _ultra,
    // 12 k_split3_hard, 13 k_split3_ultra,
    // 6 k_split3_flights, 7 bracket, 8 tally all-reduce (includes the wait for the slowest rank), 9 FP all-gather,
    // 10 cascade exchange (all-gather of the 3rd-split tasks), 11 bank sort
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_which;
    size_t ev_used = 0;
    cudaEvent_t ev_mark[2] = {nullptr, nullptr};
    static constexpr int kFamilies = 16;
    double k_ms[kFamilies] = {0};
    int64_t k_launches[kFamilies] = {0};
    void timer_begin(int which);
    void timer_end();
    void timer_collect();
};

}  // namespace c2d
