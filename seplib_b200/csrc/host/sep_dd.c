/* sep_dd.c -- SEP_NGPU: an unchanged seplib program on N GPUs.
 *
 * The reference is one address space ("offers nothing here").  The device layer decomposes the
 * box into slabs along z, one process per GPU (include/sepgpu.h, "slab domain decomposition").  This file puts that
 * behind the sep_* API without asking the program to change:
 *
 *   SEP_NGPU=N ./prg1
 *
 * At the first hot call -- before this process has touched CUDA -- the library forks N-1 copies of the program.
 * Every copy holds the whole atoms[] array in the state the program had prepared, carries on executing the
 * program's own loop, and drives one GPU with the atoms of its slab.  Copies 1..N-1 are silent: stdout and every file
 * already open for writing are redirected to /dev/null, files the library writes are not written.  Whenever atoms[]
 * has to be brought up to date (sep_gpu_sync, a host reader of the library, a page fault in SEP_SYNC=auto) each
 * process fetches its own atoms by global index and the processes exchange them through a shared mapping, so that
 * every copy sees all atoms and keeps taking the same decisions.  Sums (sepret, the thermostat, the rebuild trigger)
 * are global on the device already.
 *
 * What runs decomposed: sep_force_pairs / sep_force_lj with the built-in pair functions, sep_nosehoover,
 * sep_leapfrog, sep_reset_*, and every host-side reader.  Anything else stops with an error under SEP_NGPU > 1.
 * The program must be deterministic up to its first hot call (an initial state seeded from the clock differs between
 * the copies), and files it opens itself for writing after that point are written by every copy.
 */
#define _DEFAULT_SOURCE
#include "sep_host.h"

#include <dirent.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#undef fopen

typedef struct {
    double x[3], v[3], f[3], a[3], xn[3], pv[3], pa[3];
    int cn[3], cr[3];
} sepdd_rec;

typedef struct {
    unsigned char id[128];
    int id_ready;
    int bar_count, bar_sense;
    int failed;
    size_t npart;
} sepdd_shared;

static int g_world = 0, g_rank = 0;
static sepdd_shared *g_sh = NULL;
static sepdd_rec *g_rec = NULL;
static int g_local_sense = 0;
static pid_t g_children[64];
static int g_nchildren = 0;
static int g_allow = 0;
static int *g_rows = NULL; static size_t g_rows_cap = 0;

int sepdd_world(void)
{
    if (g_world == 0) {
        const char *e = getenv("SEP_NGPU");
        g_world = e ? atoi(e) : 1;
        if (g_world < 1) g_world = 1;
        if (g_world > 64) sep_error("SEP_NGPU: at most 64");
    }
    return g_world;
}

int sepdd_rank(void) { return g_rank; }
void sepdd_allow(void) { g_allow = 1; }

void sepdd_guard(const sep_binding *b)
{
    const int ok = g_allow;
    g_allow = 0;
    if (b->dd && !ok)
        sep_error("this call is not available with SEP_NGPU > 1 (decomposed runs shard sep_force_pairs / sep_force_lj with the "
                  "built-in pair functions, sep_nosehoover and sep_leapfrog)");
}

void sepdd_mark_failed(void)
{
    if (g_sh) __atomic_store_n(&g_sh->failed, 1, __ATOMIC_RELEASE);
}

FILE *sepdd_fopen(const char *path, const char *mode)
{
    if (g_rank > 0 && mode && (mode[0] == 'w' || mode[0] == 'a')) return fopen("/dev/null", mode);
    return fopen(path, mode);
}

static void sepdd_barrier(void)
{
    const int sense = !g_local_sense;
    g_local_sense = sense;
    if (__atomic_add_fetch(&g_sh->bar_count, 1, __ATOMIC_ACQ_REL) == g_world) {
        __atomic_store_n(&g_sh->bar_count, 0, __ATOMIC_RELAXED);
        __atomic_store_n(&g_sh->bar_sense, sense, __ATOMIC_RELEASE);
        return;
    }
    const time_t t0 = time(NULL);
    unsigned spins = 0;
    while (__atomic_load_n(&g_sh->bar_sense, __ATOMIC_ACQUIRE) != sense) {
        if ((++spins & 1023u) == 0) {
            if (__atomic_load_n(&g_sh->failed, __ATOMIC_ACQUIRE)) { fflush(NULL); _exit(EXIT_FAILURE); }
            if (time(NULL) - t0 > 300) sep_error("SEP_NGPU: the other processes never reached the same library call (did the copies diverge?)");
            sched_yield();
        }
    }
}

static void sepdd_reap(void)
{
    fflush(NULL);
    for (int k = 0; k < g_nchildren; k++) {
        if (g_sh && __atomic_load_n(&g_sh->failed, __ATOMIC_ACQUIRE)) kill(g_children[k], SIGTERM);
        int st;
        waitpid(g_children[k], &st, 0);
    }
    g_nchildren = 0;
}

/* in a forked copy: nothing this process prints or has open for writing reaches the outside */
static void sepdd_silence(void)
{
    const int nul = open("/dev/null", O_WRONLY);
    if (nul < 0) return;
    if (!getenv("SEP_DD_STDOUT_ALL")) dup2(nul, 1);
    DIR *d = opendir("/proc/self/fd");
    if (d) {
        struct dirent *e;
        int fds[256], nf = 0;
        while ((e = readdir(d)) && nf < 256) {
            const int fd = atoi(e->d_name);
            if (fd > 2 && fd != nul && fd != dirfd(d)) fds[nf++] = fd;
        }
        closedir(d);
        for (int k = 0; k < nf; k++) {
            struct stat st;
            const int fl = fcntl(fds[k], F_GETFL);
            if (fl < 0 || (fl & O_ACCMODE) == O_RDONLY) continue;
            if (fstat(fds[k], &st) == 0 && S_ISREG(st.st_mode)) dup2(nul, fds[k]);
        }
    }
    close(nul);
}

static const int *sepdd_rows(sep_binding *b, int *n_own_out)
{
    int z0, z1, n_own, n_halo;
    sepb_check(sepgpu_dd_layers(b->gpu, &z0, &z1, &n_own, &n_halo), "SEP_NGPU: layers");
    if ((size_t)n_own > g_rows_cap) {
        free(g_rows);
        g_rows_cap = (size_t)n_own + (size_t)n_own / 4 + 1024;
        g_rows = (int *)malloc(sizeof(int) * g_rows_cap);
        if (!g_rows) sep_error("SEP_NGPU: out of memory");
    }
    sepb_check(sepgpu_set_host_rows(b->gpu, NULL), "SEP_NGPU: rows");
    sepb_check(sepgpu_get(b->gpu, SEPGPU_F_GID, g_rows, 0), "SEP_NGPU: global ids");
    *n_own_out = n_own;
    return g_rows;
}

/* first hot call on this array: fork, then every process creates its context and takes the atoms of its slab */
void sepdd_start(sep_binding *b, sepsys *sys)
{
    const int world = sepdd_world();
    if (g_sh) sep_error("SEP_NGPU: one atom array per program");
    if (sys->neighb_update == SEP_BRUTE) sep_error("SEP_NGPU: decomposed runs need a neighbour list (SEP_LLIST_NEIGHBLIST)");
    const size_t n = (size_t)sys->npart;
    const size_t bytes = sizeof(sepdd_shared) + 64 + sizeof(sepdd_rec) * n;
    void *m = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) sep_error("SEP_NGPU: cannot map the exchange buffer");
    g_sh = (sepdd_shared *)m;
    g_rec = (sepdd_rec *)((char *)m + ((sizeof(sepdd_shared) + 63) & ~(size_t)63));
    memset(g_sh, 0, sizeof *g_sh);
    g_sh->npart = n;
    fflush(NULL);
    g_rank = 0;
    for (int r = 1; r < world; r++) {
        const pid_t pid = fork();
        if (pid < 0) sep_error("SEP_NGPU: fork failed");
        if (pid == 0) { g_rank = r; g_nchildren = 0; sepdd_silence(); break; }
        g_children[g_nchildren++] = pid;
    }
    if (g_rank == 0) atexit(sepdd_reap);

    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    const int nz = sys->nsubbox[2];
    const size_t ncap = (size_t)(1.6 * (double)n / world) + (size_t)(3.0 * (double)n / (nz > 0 ? nz : 1)) + 1024;
    sepb_check(sepgpu_create(&b->gpu, ncap, g_rank), "sepgpu_create");
    if (g_rank == 0) {
        sepb_check(sepgpu_dd_unique_id(g_sh->id), "SEP_NGPU: unique id");
        __atomic_store_n(&g_sh->id_ready, 1, __ATOMIC_RELEASE);
    } else {
        const time_t t0 = time(NULL);
        while (!__atomic_load_n(&g_sh->id_ready, __ATOMIC_ACQUIRE)) {
            if (time(NULL) - t0 > 120 || __atomic_load_n(&g_sh->failed, __ATOMIC_ACQUIRE)) { fflush(NULL); _exit(EXIT_FAILURE); }
            sched_yield();
        }
    }
    sepb_check(sepgpu_dd_init(b->gpu, g_rank, world, g_sh->id, &gs, (long long)n), "SEP_NGPU: dd_init");
    int z0, z1, no, nh;
    sepb_check(sepgpu_dd_layers(b->gpu, &z0, &z1, &no, &nh), "SEP_NGPU: layers");
    /* my atoms: wrapped z in the cell layers [z0, z1) -- the binning the device uses (reference source/sepprfrc.c:404-409) */
    g_rows_cap = ncap;
    g_rows = (int *)malloc(sizeof(int) * g_rows_cap);
    if (!g_rows) sep_error("SEP_NGPU: out of memory");
    int mine = 0;
    for (size_t i = 0; i < n; i++) {
        int cz = (int)(b->atoms[i].x[2] / sys->lsubbox[2]);
        if (cz < 0) cz = 0;
        if (cz >= nz) cz = nz - 1;
        if (cz >= z0 && cz < z1) {
            if ((size_t)mine >= g_rows_cap) sep_error("SEP_NGPU: slab %d holds more atoms than its capacity", g_rank);
            g_rows[mine++] = (int)i;
        }
    }
    sepb_check(sepgpu_dd_set_owned(b->gpu, mine), "SEP_NGPU: set_owned");
    sepb_check(sepgpu_put(b->gpu, SEPGPU_F_GID, g_rows, 0), "SEP_NGPU: global ids");
    sepb_check(sepgpu_set_host_rows(b->gpu, g_rows), "SEP_NGPU: rows");
    b->dd = 1;
    b->host_dirty = ~0u;
    b->dev_dirty = 0;
    b->dd_rows_valid = 1;
}

/* rows of atoms[] that belong to the atoms this process owns NOW (ownership moves with migration) */
void sepdd_before_upload(sep_binding *b)
{
    if (b->dd_rows_valid) { b->dd_rows_valid = 0; return; }     /* initial upload: rows were just set */
    int n_own;
    const int *rows = sepdd_rows(b, &n_own);
    sepb_check(sepgpu_set_host_rows(b->gpu, rows), "SEP_NGPU: rows");
}

/* device -> every process's atoms[]: own atoms into the shared records by global index, barrier, everyone copies all */
void sepdd_download(sep_binding *b, int nf, const int *fl, const size_t *offs)
{
    static const struct { int fid; size_t roff; size_t bytes; } map[] = {
        {SEPGPU_F_X, offsetof(sepdd_rec, x), 24}, {SEPGPU_F_V, offsetof(sepdd_rec, v), 24}, {SEPGPU_F_F, offsetof(sepdd_rec, f), 24},
        {SEPGPU_F_A, offsetof(sepdd_rec, a), 24}, {SEPGPU_F_XN, offsetof(sepdd_rec, xn), 24}, {SEPGPU_F_PV, offsetof(sepdd_rec, pv), 24},
        {SEPGPU_F_PA, offsetof(sepdd_rec, pa), 24}, {SEPGPU_F_CROSS_NEIGHB, offsetof(sepdd_rec, cn), 12},
        {SEPGPU_F_CROSSINGS, offsetof(sepdd_rec, cr), 12}};
    size_t roffs[16], bytes[16];
    for (int f = 0; f < nf; f++) {
        int hit = -1;
        for (size_t q = 0; q < sizeof map / sizeof map[0]; q++) if (map[q].fid == fl[f]) hit = (int)q;
        if (hit < 0) sep_error("SEP_NGPU: field %d cannot be exchanged", fl[f]);
        roffs[f] = map[hit].roff; bytes[f] = map[hit].bytes;
    }
    int n_own;
    const int *rows = sepdd_rows(b, &n_own);
    sepb_check(sepgpu_set_host_rows(b->gpu, rows), "SEP_NGPU: rows");
    sepb_check(sepgpu_get_fields(b->gpu, g_rec, sizeof(sepdd_rec), nf, fl, roffs), "SEP_NGPU: download");
    sepdd_barrier();
    const size_t n = g_sh->npart;
    for (size_t i = 0; i < n; i++) {
        char *dst = (char *)&b->atoms[i];
        const char *src = (const char *)&g_rec[i];
        for (int f = 0; f < nf; f++) memcpy(dst + offs[f], src + roffs[f], bytes[f]);
    }
    sepdd_barrier();
}
