/*
 * fakempi.c -- a 2-rank stand-in for the OpenMPI-1.x Fortran bindings that the
 * reference's prebuilt binary (src_20121026/compton2d, ifort 11) imports.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): it lets the UNMODIFIED reference binary
 * run in a container without MPI so that its results can pin the CPU
 * restatement (oracle/c2d_oracle.c) and serve as the timed CPU baseline.
 * Nothing in the product path (compton2d_b200/) links or loads this file.
 *
 * The binary imports exactly 12 entry points:
 *   mpi_init_ mpi_finalize_ mpi_comm_rank_ mpi_comm_size_ mpi_send_ mpi_recv_
 *   mpi_bcast_ mpi_reduce_ mpi_barrier_ mpi_address_ mpi_type_struct_
 *   mpi_type_commit_   (+ 13 dummy data symbols mpi_*_fn_)
 * Constants follow OpenMPI 1.x mpif.h: MPI_INTEGER=7, MPI_DOUBLE_PRECISION=17,
 * MPI_CHARACTER=5, MPI_SUM=3, MPI_MAX=1, MPI_ANY_SOURCE=MPI_ANY_TAG=-1,
 * status(1)=MPI_SOURCE, status(2)=MPI_TAG.
 *
 * Rank 0 (dispatcher, reference xec2d.f:29-125) and rank 1 (the single worker,
 * xec2d.f:127-193) are joined by one socketpair created in mpi_init_ (legal
 * because MPI_INIT is the program's first statement, compton2d.f:11). With one
 * worker every receive matches the next message in FIFO order.
 *
 * Extras (all driven by environment variables, all optional):
 *   C2D_TRACE=<prefix>    write <prefix>.<rank> : one line per MPI call
 *                         "idx op count dtype peer tag bytes t_monotonic"
 *   C2D_SNAPSPEC=<file>   snapshot requests, one per line:
 *                           AT <rank> <callidx> <label>
 *                           R <symbol> <offset> <nbytes>
 *                         "AT" lines select the MPI call index (on entry to the
 *                         call) at which a snapshot is taken; "R" lines list the
 *                         byte ranges (relative to ELF symbols of the binary,
 *                         e.g. COMMON block zones_) that every snapshot dumps,
 *                         in order, to <C2D_SNAPDIR>/<label>.bin.
 *   C2D_SNAPDIR=<dir>     where snapshots go (default ".")
 * Symbol addresses come from the static symbol table of /proc/self/exe (the
 * binary is a non-PIE EXEC and is not stripped).
 */
#define _GNU_SOURCE
#include <elf.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

static int g_rank = 0, g_fd = -1;
static long g_idx = 0;
static FILE *g_trace = NULL;

/* ---------------------------------------------------------------- helpers */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static void die(const char *m) {
    fprintf(stderr, "[fakempi rank %d] fatal: %s (errno %d)\n", g_rank, m, errno);
    abort();
}
static void wr(const void *p, size_t n) {
    const char *c = p;
    while (n) {
        ssize_t k = write(g_fd, c, n);
        if (k <= 0) { if (errno == EINTR) continue; die("write"); }
        c += k; n -= (size_t)k;
    }
}
static void rd(void *p, size_t n) {
    char *c = p;
    while (n) {
        ssize_t k = read(g_fd, c, n);
        if (k <= 0) { if (k < 0 && errno == EINTR) continue; die("read (peer gone?)"); }
        c += k; n -= (size_t)k;
    }
}

/* ------------------------------------------------------- derived datatypes */
#define MAXT 8192   /* vol_create_job_type makes a new type every cycle (vol_mpi.f:54-81) */
#define MAXF 64
typedef struct { int n, len[MAXF], disp[MAXF], typ[MAXF]; } dtype_t;
static dtype_t g_types[MAXT];
static int g_ntypes = 0;

static int base_size(int t) {
    switch (t) {
    case 5: return 1;   /* MPI_CHARACTER */
    case 7: return 4;   /* MPI_INTEGER */
    case 13: return 4;  /* MPI_REAL */
    case 17: return 8;  /* MPI_DOUBLE_PRECISION */
    default: fprintf(stderr, "[fakempi] unknown datatype %d\n", t); abort();
    }
}
static size_t type_bytes(int t) {
    if (t >= 1000) {
        dtype_t *d = &g_types[t - 1000];
        size_t s = 0;
        for (int i = 0; i < d->n; i++) s += (size_t)d->len[i] * base_size(d->typ[i]);
        return s;
    }
    return (size_t)base_size(t);
}
/* pack/unpack `count` items of type t at buf to/from a contiguous stream */
static void send_payload(char *buf, int count, int t) {
    if (t >= 1000) {
        dtype_t *d = &g_types[t - 1000];
        for (int c = 0; c < count; c++)   /* count is always 1 in the reference */
            for (int i = 0; i < d->n; i++)
                wr(buf + d->disp[i], (size_t)d->len[i] * base_size(d->typ[i]));
    } else wr(buf, (size_t)count * base_size(t));
}
static void recv_payload(char *buf, int count, int t) {
    if (t >= 1000) {
        dtype_t *d = &g_types[t - 1000];
        for (int c = 0; c < count; c++)
            for (int i = 0; i < d->n; i++)
                rd(buf + d->disp[i], (size_t)d->len[i] * base_size(d->typ[i]));
    } else rd(buf, (size_t)count * base_size(t));
}

/* ------------------------------------------------------------- ELF symbols */
typedef struct { char name[48]; uint64_t addr, size; } sym_t;
static sym_t *g_syms = NULL;
static int g_nsyms = 0;

static void load_symbols(void) {
    int fd = open("/proc/self/exe", O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    fstat(fd, &st);
    unsigned char *m = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return;
    Elf64_Ehdr *eh = (Elf64_Ehdr *)m;
    Elf64_Shdr *sh = (Elf64_Shdr *)(m + eh->e_shoff);
    for (int i = 0; i < eh->e_shnum; i++) {
        if (sh[i].sh_type != SHT_SYMTAB) continue;
        Elf64_Sym *sy = (Elf64_Sym *)(m + sh[i].sh_offset);
        int n = sh[i].sh_size / sizeof(Elf64_Sym);
        const char *str = (const char *)(m + sh[sh[i].sh_link].sh_offset);
        g_syms = calloc(n, sizeof(sym_t));
        for (int k = 0; k < n; k++) {
            if (ELF64_ST_TYPE(sy[k].st_info) != STT_OBJECT) continue;
            const char *nm = str + sy[k].st_name;
            if (strlen(nm) >= sizeof(g_syms[0].name)) continue;
            strcpy(g_syms[g_nsyms].name, nm);
            g_syms[g_nsyms].addr = sy[k].st_value;
            g_syms[g_nsyms].size = sy[k].st_size;
            g_nsyms++;
        }
    }
    munmap(m, st.st_size);
}
static sym_t *find_sym(const char *nm) {
    for (int i = 0; i < g_nsyms; i++)
        if (!strcmp(g_syms[i].name, nm)) return &g_syms[i];
    return NULL;
}

/* --------------------------------------------------------------- snapshots */
typedef struct { long idx; char label[64]; } snap_at_t;
typedef struct { uint64_t addr, n; } snap_rng_t;
static snap_at_t *g_at = NULL; static int g_nat = 0;
static snap_rng_t *g_rng = NULL; static int g_nrng = 0;
static char g_snapdir[512] = ".";

static void load_snapspec(void) {
    const char *f = getenv("C2D_SNAPSPEC");
    if (!f) return;
    if (getenv("C2D_SNAPDIR")) snprintf(g_snapdir, sizeof g_snapdir, "%s", getenv("C2D_SNAPDIR"));
    FILE *fp = fopen(f, "r");
    if (!fp) die("cannot open C2D_SNAPSPEC");
    load_symbols();
    char line[256];
    int cap_at = 0, cap_r = 0;
    while (fgets(line, sizeof line, fp)) {
        if (line[0] == 'A') {
            int rk; long idx; char lab[64];
            if (sscanf(line, "AT %d %ld %63s", &rk, &idx, lab) != 3) die("bad AT line");
            if (rk != g_rank) continue;
            if (g_nat == cap_at) { cap_at = cap_at ? 2 * cap_at : 16; g_at = realloc(g_at, cap_at * sizeof *g_at); }
            g_at[g_nat].idx = idx; strcpy(g_at[g_nat].label, lab); g_nat++;
        } else if (line[0] == 'R') {
            char s[64]; unsigned long off, n;
            if (sscanf(line, "R %63s %lu %lu", s, &off, &n) != 3) die("bad R line");
            sym_t *sy = find_sym(s);
            if (!sy) { fprintf(stderr, "[fakempi] no symbol %s\n", s); abort(); }
            if (off + n > sy->size) { fprintf(stderr, "[fakempi] range beyond %s\n", s); abort(); }
            if (g_nrng == cap_r) { cap_r = cap_r ? 2 * cap_r : 64; g_rng = realloc(g_rng, cap_r * sizeof *g_rng); }
            g_rng[g_nrng].addr = sy->addr + off; g_rng[g_nrng].n = n; g_nrng++;
        }
    }
    fclose(fp);
}
static void maybe_snapshot(void) {
    for (int i = 0; i < g_nat; i++) {
        if (g_at[i].idx != g_idx) continue;
        char path[700];
        snprintf(path, sizeof path, "%s/%s.bin", g_snapdir, g_at[i].label);
        FILE *fp = fopen(path, "wb");
        if (!fp) die("snapshot open");
        for (int r = 0; r < g_nrng; r++)
            fwrite((void *)(uintptr_t)g_rng[r].addr, 1, g_rng[r].n, fp);
        fclose(fp);
    }
}

/* ------------------------------------------------------------------- trace */
static void enter(const char *op, int count, int t, int peer, int tag) {
    maybe_snapshot();
    if (g_trace) {
        size_t bytes = (count > 0 && t) ? (size_t)count * type_bytes(t) : 0;
        fprintf(g_trace, "%ld %s %d %d %d %d %zu %.6f\n", g_idx, op, count, t, peer, tag, bytes, now_s());
    }
    g_idx++;
}

/* --------------------------------------------------------------- MPI calls */
void mpi_init_(int *ierr) {
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv)) die("socketpair");
    fflush(NULL);
    pid_t p = fork();
    if (p < 0) die("fork");
    if (p == 0) { g_rank = 1; g_fd = sv[1]; close(sv[0]); }
    else        { g_rank = 0; g_fd = sv[0]; close(sv[1]); }
    const char *tp = getenv("C2D_TRACE");
    if (tp) {
        char path[600];
        snprintf(path, sizeof path, "%s.%d", tp, g_rank);
        g_trace = fopen(path, "w");
    }
    load_snapspec();
    *ierr = 0;
}
void mpi_finalize_(int *ierr) {
    enter("finalize", 0, 0, -1, -1);
    if (g_trace) fclose(g_trace);
    if (g_rank == 0) { int st; wait(&st); }
    *ierr = 0;
}
void mpi_comm_rank_(int *c, int *r, int *e) { (void)c; *r = g_rank; *e = 0; }
void mpi_comm_size_(int *c, int *s, int *e) { (void)c; *s = 2; *e = 0; }
void mpi_address_(void *p, int *a, int *e) { *a = (int)(intptr_t)p; *e = 0; }
void mpi_type_struct_(int *n, int *len, int *disp, int *typ, int *newt, int *e) {
    if (g_ntypes >= MAXT || *n > MAXF) die("too many derived types/fields");
    dtype_t *d = &g_types[g_ntypes];
    d->n = *n;
    for (int i = 0; i < *n; i++) { d->len[i] = len[i]; d->disp[i] = disp[i]; d->typ[i] = typ[i]; }
    *newt = 1000 + g_ntypes++;
    *e = 0;
}
void mpi_type_commit_(int *t, int *e) { (void)t; *e = 0; }

typedef struct { int tag; int pad; uint64_t bytes; } hdr_t;

void mpi_send_(void *buf, int *count, int *t, int *dst, int *tag, int *c, int *e) {
    (void)c;
    enter("send", *count, *t, *dst, *tag);
    hdr_t h = { *tag, 0, (uint64_t)*count * type_bytes(*t) };
    wr(&h, sizeof h);
    if (h.bytes) send_payload(buf, *count, *t);
    *e = 0;
}
void mpi_recv_(void *buf, int *count, int *t, int *src, int *tag, int *c, int *st, int *e) {
    (void)c; (void)src;
    enter("recv", *count, *t, *src, *tag);
    hdr_t h;
    rd(&h, sizeof h);
    uint64_t want = (uint64_t)*count * type_bytes(*t);
    if (h.bytes > want) die("recv: message longer than buffer");
    if (h.bytes == want) { if (want) recv_payload(buf, *count, *t); }
    else if (h.bytes) {             /* shorter message of a basic type */
        if (*t >= 1000) die("recv: short derived message");
        rd(buf, h.bytes);
    }
    st[0] = 1 - g_rank;   /* MPI_SOURCE */
    st[1] = h.tag;        /* MPI_TAG */
    st[2] = 0;
    if (g_trace) fprintf(g_trace, "# recv tag %d bytes %lu\n", h.tag, (unsigned long)h.bytes);
    *e = 0;
}
void mpi_bcast_(void *buf, int *count, int *t, int *root, int *c, int *e) {
    (void)c;
    enter("bcast", *count, *t, *root, -1);
    if (g_rank == *root) send_payload(buf, *count, *t);
    else recv_payload(buf, *count, *t);
    *e = 0;
}
void mpi_reduce_(void *sbuf, void *rbuf, int *count, int *t, int *op, int *root, int *c, int *e) {
    (void)c;
    enter("reduce", *count, *t, *root, *op);
    size_t n = (size_t)*count, bytes = n * base_size(*t);
    if (g_rank != *root) { wr(sbuf, bytes); *e = 0; return; }
    void *tmp = malloc(bytes ? bytes : 1);
    rd(tmp, bytes);
    if (*t == 17) {
        double *s = sbuf, *r = rbuf, *o = tmp;
        for (size_t i = 0; i < n; i++)
            r[i] = (*op == 3) ? s[i] + o[i] : (s[i] > o[i] ? s[i] : o[i]);
    } else if (*t == 7) {
        int *s = sbuf, *r = rbuf, *o = tmp;
        for (size_t i = 0; i < n; i++)
            r[i] = (*op == 3) ? s[i] + o[i] : (s[i] > o[i] ? s[i] : o[i]);
    } else die("reduce: unsupported type");
    free(tmp);
    *e = 0;
}
void mpi_barrier_(int *c, int *e) {
    (void)c;
    enter("barrier", 0, 0, -1, -1);
    char b = 1;
    if (g_rank == 0) { wr(&b, 1); rd(&b, 1); } else { rd(&b, 1); wr(&b, 1); }
    *e = 0;
}

/* data symbols referenced from mpif.h-derived COMMON/externals */
int mpi_comm_dup_fn_, mpi_comm_null_copy_fn_, mpi_comm_null_delete_fn_,
    mpi_dup_fn_, mpi_null_copy_fn_, mpi_null_delete_fn_,
    mpi_type_dup_fn_, mpi_type_null_copy_fn_, mpi_type_null_delete_fn_,
    mpi_win_dup_fn_, mpi_win_null_copy_fn_, mpi_win_null_delete_fn_,
    mpi_conversion_fn_null_;
