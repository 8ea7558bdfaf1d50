/* sep_core.c -- host side of the seplib API: allocation, system setup, host<->device binding,
 * coherence, error reporting and the small scalar routines (pressure tensor, momentum, ...).
 *
 * The functions here mirror the behaviour of the reference's source/sepinit.c, source/sepret.c and
 * the parts of source/sepmisc.c that the example programs use; each cites its counterpart.  They
 * never compute forces or integrate: that is the device layer's job (include/sepgpu.h).
 */
#define _DEFAULT_SOURCE
#include "sep_host.h"
#include <float.h>

#include <ctype.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

/* ---------------------------------------------------------------------------------------------------
 * error reporting (source/sepmisc.c:16-113): mini-printf to stdout, then exit for errors
 * ------------------------------------------------------------------------------------------------- */
static void emit(const char *prefix, const char *fmt, va_list ap)
{
    fputs(prefix, stdout);
    for (const char *p = fmt; *p; p++) {
        if (*p != '%') { putchar(*p); continue; }
        p++;
        if (*p == 'd' || *p == 'i') printf("%d", va_arg(ap, int));
        else if (*p == 'f') printf("%f", va_arg(ap, double));
        else if (*p == 'c') printf("%c", va_arg(ap, int));
        else if (*p == 's') { const char *s = va_arg(ap, char *); fputs(s ? s : "(null)", stdout); }
        else if (*p == '\0') break;
        else putchar(*p);
    }
    putchar('\n');
}

void sep_error(char *str, ...)
{
    va_list ap;
    va_start(ap, str);
    emit("sep-error -> ", str, ap);
    va_end(ap);
    printf("BAILING OUT\n");
    fflush(stdout);
    if (sepdd_rank() > 0) {                      /* copies 1..N-1 have no stdout: say it where somebody can see it */
        fprintf(stderr, "sep-error in copy %d of a SEP_NGPU run: ", sepdd_rank());
        va_start(ap, str);
        vfprintf(stderr, str, ap);
        va_end(ap);
        fputc('\n', stderr);
    }
    sepdd_mark_failed();
    exit(EXIT_FAILURE);
}

void sep_warning(char *str, ...)
{
    va_list ap;
    va_start(ap, str);
    emit("sep-warning -> ", str, ap);
    va_end(ap);
    fflush(stdout);
}

void sepb_check(int rc, const char *where)
{
    if (rc == 0) return;
    if (rc == SEPGPU_ENEIGHB) sep_error("%s: Too many neighbours", (char *)where);
    if (rc == SEPGPU_ECELL) sep_error("%s: Index larger than array length", (char *)where);
    if (rc == SEPGPU_ETABLE) sep_error("%s: a pair came closer than the table of the pair function reaches (SEP_TABLE_RMIN)", (char *)where);
    if (rc == SEPGPU_ENODEV)
        sep_error("%s: no CUDA device -- seplib-b200 has no CPU path (%s)", (char *)where, (char *)sepgpu_last_error());
    sep_error("%s: device layer failed (%d): %s", (char *)where, rc, (char *)sepgpu_last_error());
}

/* ---------------------------------------------------------------------------------------------------
 * registry of atom arrays
 * ------------------------------------------------------------------------------------------------- */
static sep_binding *g_bindings = NULL;
static int g_sync_mode = -1;
static unsigned long long g_dpd_seed = 0x5EB11B200ULL;

/* ---------------------------------------------------------------------------------------------------
 * SEP_SYNC=auto: page protection of atoms[].  The reference exposes seppart as a public
 * struct (include/sepstrct.h:23-61) that callers read and write between library calls; with the state resident on the
 * device the library has to notice.  The array is an anonymous mapping of its own:
 *   PROT_NONE   the device copy is newer (after any hot call)
 *   PROT_READ   host and device agree; a write faults once more and marks the host copy as changed
 *   READ|WRITE  the host copy may have changed: everything is uploaded before the next hot call
 * A fault elsewhere goes to whoever handled SIGSEGV before.
 * ------------------------------------------------------------------------------------------------- */
static struct sigaction g_old_segv;
static int g_segv_installed = 0;

static void sepb_set_prot(sep_binding *b, int prot)
{
    if (!b->managed || b->prot == prot) return;
    if (mprotect(b->map_base, b->map_bytes, prot) != 0) sep_error("%s: mprotect failed", (char *)__func__);
    b->prot = prot;
}

void sepb_unprotect_all(sep_binding *b)
{
    if (!b->managed || b->prot == (PROT_READ | PROT_WRITE)) return;
    const int was = b->prot;
    sepb_set_prot(b, PROT_READ | PROT_WRITE);
    if (was == PROT_NONE && b->gpu) sepb_download(b, ~0u);
}

static void sep_segv_handler(int sig, siginfo_t *si, void *uctx)
{
    const char *addr = (const char *)si->si_addr;
    for (sep_binding *b = sepb_first(); b; b = b->next) {
        if (!b->managed || addr < (const char *)b->map_base || addr >= (const char *)b->map_base + b->map_bytes) continue;
        if (b->prot == PROT_NONE) {                      /* first touch since the device moved on: bring atoms[] up to date */
            sepb_set_prot(b, PROT_READ | PROT_WRITE);
            if (b->gpu) sepb_download(b, ~0u);
            sepb_set_prot(b, PROT_READ);
            return;
        }
        if (b->prot == PROT_READ) {                      /* a write: the host copy is the newer one from here on */
            sepb_set_prot(b, PROT_READ | PROT_WRITE);
            b->host_dirty |= SEPB_ALL_STATE | SEPB_EXCL;
            if (b->dpd_state_on_device) b->host_dirty |= SEPB_PV | SEPB_PA;
            if (b->x0_on_device) b->host_dirty |= SEPB_X0;
            b->dev_dirty = 0;
            return;
        }
        break;                                           /* already writable: a genuine fault */
    }
    /* not ours: previous handler, or the default action */
    if (g_old_segv.sa_flags & SA_SIGINFO) {
        if (g_old_segv.sa_sigaction) { g_old_segv.sa_sigaction(sig, si, uctx); return; }
    } else if (g_old_segv.sa_handler != SIG_DFL && g_old_segv.sa_handler != SIG_IGN && g_old_segv.sa_handler) {
        g_old_segv.sa_handler(sig);
        return;
    }
    signal(SIGSEGV, SIG_DFL);
    raise(SIGSEGV);
}

static void sep_install_segv(void)
{
    if (g_segv_installed) return;
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = sep_segv_handler;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    if (sigaction(SIGSEGV, &sa, &g_old_segv) == 0) g_segv_installed = 1;
}

void sepb_dev_newer(sep_binding *b, unsigned bits)
{
    b->dev_dirty |= bits;
    if (b->managed && sep_sync_mode() == SEP_SYNC_AUTO) {
        /* what the host changed and has not uploaded yet must not be lost: those fields stay the host's */
        if (b->host_dirty == 0) sepb_set_prot(b, PROT_NONE);
    }
}

int sep_sync_mode(void)
{
    if (g_sync_mode < 0) {
        const char *e = getenv("SEP_SYNC");
        g_sync_mode = SEP_SYNC_AUTO;
        if (e) {
            if (!strcmp(e, "lazy")) g_sync_mode = SEP_SYNC_LAZY;
            else if (!strcmp(e, "full")) g_sync_mode = SEP_SYNC_FULL;
            else if (!strcmp(e, "step")) g_sync_mode = SEP_SYNC_STEP;
            else if (!strcmp(e, "auto")) g_sync_mode = SEP_SYNC_AUTO;
            else sep_warning("SEP_SYNC=%s not understood (auto|lazy|step|full); using auto", (char *)e);
        }
    }
    return g_sync_mode;
}

void sepb_unprotect_all(sep_binding *b);

void sep_gpu_set_sync(int mode)
{
    if (mode == SEP_SYNC_LAZY || mode == SEP_SYNC_STEP || mode == SEP_SYNC_FULL || mode == SEP_SYNC_AUTO) g_sync_mode = mode;
    if (mode != SEP_SYNC_AUTO)           /* leaving auto mode: nothing stays protected */
        for (sep_binding *b = sepb_first(); b; b = b->next) sepb_unprotect_all(b);
}

int sepb_eager(const sep_binding *b)
{
    const int mode = sep_sync_mode();
    return mode == SEP_SYNC_STEP || mode == SEP_SYNC_FULL || (mode == SEP_SYNC_AUTO && !(b && b->managed));
}

void sep_gpu_set_dpd_seed(unsigned long long seed) { g_dpd_seed = seed; }
unsigned long long sep_dpd_seed(void) { return g_dpd_seed; }

sep_binding *sepb_find(const seppart *atoms)
{
    for (sep_binding *b = g_bindings; b; b = b->next)
        if (b->atoms == atoms) return b;
    return NULL;
}

sep_binding *sepb_find_mol(const sepmolinfo *molptr)
{
    if (!molptr) return NULL;
    for (sep_binding *b = g_bindings; b; b = b->next)
        if (b->molptr == molptr) return b;
    return NULL;
}

sep_binding *sepb_first(void) { return g_bindings; }

sep_binding *sepb_register(seppart *atoms, size_t npart)
{
    sep_binding *b = calloc(1, sizeof *b);
    if (!b) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    b->atoms = atoms;
    b->npart = npart;
    b->host_dirty = ~0u;
    b->next = g_bindings;
    g_bindings = b;
    return b;
}

void sepb_unregister(seppart *atoms)
{
    sep_binding **pp = &g_bindings;
    while (*pp) {
        if ((*pp)->atoms == atoms) {
            sep_binding *b = *pp;
            *pp = b->next;
            if (b->gpu) sepgpu_destroy(b->gpu);
            free(b->blengths_host); free(b->angles_host); free(b->dihedrals_host); free(b->noise);
            for (int q = 0; q < 4; q++) free(b->pairtab[q].fu);
            free(b);
            return;
        }
        pp = &(*pp)->next;
    }
}

void sepb_mark_host_dirty(seppart *atoms, unsigned fields)
{
    sep_binding *b = sepb_find(atoms);
    if (b) { b->host_dirty |= fields; b->dev_dirty &= ~fields; }
}

void sepb_fill_sys(const sepsys *sys, sepgpu_sys *out)
{
    for (int k = 0; k < 3; k++) {
        out->length[k] = sys->length[k];
        out->lsubbox[k] = sys->lsubbox[k];
        out->nsubbox[k] = sys->nsubbox[k];
    }
    out->cf = sys->cf;
    out->skin = sys->skin;
    out->dt = sys->dt;
    out->neighb_update = (int)sys->neighb_update;
}

#define FIELD_PTR(b, member) ((void *)&(b)->atoms[0].member)

sep_binding *sepb_prepare(seppart *atoms, sepsys *sys)
{
    sep_binding *b = sepb_find(atoms);
    if (!b) b = sepb_register(atoms, (size_t)sys->npart);    /* array not from sep_init (e.g. user malloc) */
    if ((long)b->npart != sys->npart) {
        if (b->gpu) sep_error("%s: sys.npart changed under a live device context", (char *)__func__);
        b->npart = (size_t)sys->npart;
    }
    if (!b->gpu) {
        if (sepdd_world() > 1) sepdd_start(b, sys);
        else sepb_check(sepgpu_create(&b->gpu, b->npart, -1), "sepgpu_create");
        b->host_dirty = ~0u;
        b->dev_dirty = 0;
    }
    sepdd_guard(b);
    if (sys->molptr) b->molptr = sys->molptr;
    if (sys->molptr && sys->molptr->flag_Fij == 1 && !b->fij_on && sys->molptr->num_mols > 0) {
        sepb_check(sepgpu_fij_enable(b->gpu, (int)sys->molptr->num_mols), "molecular force table");
        b->fij_on = 1;
    }
    if (!b->host_dirty) return b;

    /* every dirty field in ONE pass over the 568-byte records and one PCIe transfer */
    {
        int fields[16]; size_t offs[16]; int nf = 0;
#define WANT(bit, fid, member) if (b->host_dirty & (bit)) { fields[nf] = fid; offs[nf] = offsetof(seppart, member); nf++; }
        WANT(SEPB_X, SEPGPU_F_X, x) WANT(SEPB_V, SEPGPU_F_V, v) WANT(SEPB_M, SEPGPU_F_M, m) WANT(SEPB_Z, SEPGPU_F_Z, z)
        WANT(SEPB_TYPE, SEPGPU_F_TYPE, type) WANT(SEPB_MOL, SEPGPU_F_MOLINDEX, molindex) WANT(SEPB_XN, SEPGPU_F_XN, xn)
        WANT(SEPB_CN, SEPGPU_F_CROSS_NEIGHB, cross_neighb) WANT(SEPB_CR, SEPGPU_F_CROSSINGS, crossings)
        if (b->uploaded_once) WANT(SEPB_F, SEPGPU_F_F, f)
        if (b->dpd_state_on_device) { WANT(SEPB_PV, SEPGPU_F_PV, pv) WANT(SEPB_PA, SEPGPU_F_PA, pa) }
        if (b->x0_on_device) WANT(SEPB_X0, SEPGPU_F_X0, x0)
#undef WANT
        if (nf && b->dd) sepdd_before_upload(b);
        if (nf) sepb_check(sepgpu_put_fields(b->gpu, b->atoms, sizeof(seppart), nf, fields, offs), "upload");
    }
    if ((b->host_dirty & SEPB_EXCL) && sys->molptr &&
        (sys->molptr->flag_bonds || sys->molptr->flag_angles || sys->molptr->flag_dihedrals)) {
        sepb_check(sepgpu_put(b->gpu, SEPGPU_F_BOND, FIELD_PTR(b, bond), sizeof(seppart)), "upload bond table");
        sepb_check(sepgpu_put(b->gpu, SEPGPU_F_ANGLE, FIELD_PTR(b, angle), sizeof(seppart)), "upload angle table");
        sepb_check(sepgpu_put(b->gpu, SEPGPU_F_DIHED, FIELD_PTR(b, dihed), sizeof(seppart)), "upload dihed table");
    }
    if ((b->host_dirty & SEPB_TOPO) && sys->molptr && sys->molptr->flag_bonds) {
        const sepmolinfo *mp = sys->molptr;
        sepb_check(sepgpu_set_topology(b->gpu, mp->blist, mp->num_bonds,
                                       mp->flag_angles ? mp->alist : NULL, mp->flag_angles ? mp->num_angles : 0,
                                       mp->flag_dihedrals ? mp->dlist : NULL, mp->flag_dihedrals ? mp->num_dihedrals : 0),
                   "upload topology");
    }
    b->host_dirty = 0;
    b->uploaded_once = 1;
    return b;
}

void sepb_download(sep_binding *b, unsigned fields)
{
    if (!b || !b->gpu) return;
    fields &= b->dev_dirty;
    if (!fields) return;
    {
        int fl[16]; size_t offs[16]; int nf = 0;
#define PULL(bit, fid, member) if (fields & (bit)) { fl[nf] = fid; offs[nf] = offsetof(seppart, member); nf++; }
        PULL(SEPB_X, SEPGPU_F_X, x) PULL(SEPB_V, SEPGPU_F_V, v) PULL(SEPB_F, SEPGPU_F_F, f) PULL(SEPB_A, SEPGPU_F_A, a)
        PULL(SEPB_XN, SEPGPU_F_XN, xn) PULL(SEPB_CN, SEPGPU_F_CROSS_NEIGHB, cross_neighb) PULL(SEPB_CR, SEPGPU_F_CROSSINGS, crossings)
        PULL(SEPB_PV, SEPGPU_F_PV, pv) PULL(SEPB_PA, SEPGPU_F_PA, pa)
#undef PULL
        if (nf && b->dd) sepdd_download(b, nf, fl, offs);
        else if (nf) sepb_check(sepgpu_get_fields(b->gpu, b->atoms, sizeof(seppart), nf, fl, offs), "download");
    }
    b->dev_dirty &= ~fields;
}

void sepb_pull_scalars(sep_binding *b, sepsys *sys, sepret *ret, sepgpu_scalars *out)
{
    sepgpu_scalars s;
    sepb_check(sepgpu_read_scalars(b->gpu, &s), "read scalars");
    if (ret) {
        ret->epot = s.epot; ret->ecoul = s.ecoul; ret->ekin = s.ekin;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) {
                ret->pot_P[k][kk] = s.pot_P[3 * k + kk];
                ret->kin_P[k][kk] = s.kin_P[3 * k + kk];
                ret->pot_P_bond[k][kk] = s.pot_P_bond[3 * k + kk];
            }
    }
    if (sys) sys->max_dist2 = s.max_dist2;
    for (int k = 0; k < 4; k++)
        if (b->alpha_ptr[k]) { *b->alpha_ptr[k] = s.alpha[k]; b->alpha_seen[k] = s.alpha[k]; }
    if (out) *out = s;
}

void sepb_after_force(sep_binding *b, sepsys *sys, sepret *ret)
{
    sepb_dev_newer(b, SEPB_F | SEPB_A);
    b->last_ret = ret;
    const int mode = sep_sync_mode();
    if (mode != SEP_SYNC_LAZY) sepb_pull_scalars(b, sys, ret, NULL);
    if (mode == SEP_SYNC_FULL) sepb_download(b, SEPB_F);
}

void sep_gpu_sync(seppart *ptr)
{
    sep_binding *b = sepb_find(ptr);
    if (!b || !b->gpu) return;
    if (b->managed && b->prot == PROT_NONE) {
        sepb_set_prot(b, PROT_READ | PROT_WRITE);
        sepb_download(b, ~0u);
        sepb_set_prot(b, PROT_READ);
    } else {
        sepb_download(b, ~0u);
    }
}

void sep_gpu_invalidate(seppart *ptr)
{
    sep_binding *b = sepb_find(ptr);
    if (!b) return;
    b->host_dirty |= SEPB_ALL_STATE | SEPB_EXCL;
    if (b->dpd_state_on_device) b->host_dirty |= SEPB_PV | SEPB_PA;
    b->dev_dirty = 0;
    sepb_set_prot(b, PROT_READ | PROT_WRITE);
}

void sep_gpu_sync_scalars(seppart *ptr, sepsys *sys, sepret *ret)
{
    sep_binding *b = sepb_find(ptr);
    if (b && b->gpu) sepb_pull_scalars(b, sys, ret, NULL);
}

void *sep_gpu_handle(seppart *ptr)
{
    sep_binding *b = sepb_find(ptr);
    return b ? (void *)b->gpu : NULL;
}

long sep_gpu_export_neighb(seppart *ptr, sepsys *sys, int *pairs, long max_pairs)
{
    (void)sys;
    sep_binding *b = sepb_find(ptr);
    if (!b || !b->gpu) return -1;
    return (long)sepgpu_get_pairs(b->gpu, pairs, max_pairs);
}

/* ---------------------------------------------------------------------------------------------------
 * allocation (source/sepinit.c:15-66)
 * ------------------------------------------------------------------------------------------------- */
seppart *sep_init(size_t npart, size_t nneighb)
{
    /* zero-filled so that xn starts at 0 (the reference leaves it uninitialised; fresh heap pages
     * make it 0 in practice and the first leapfrog then requests a rebuild) */
    seppart *p = NULL;
    size_t map_bytes = 0;
    const size_t want = (npart ? npart : 1) * sizeof(seppart);
    if (sep_sync_mode() == SEP_SYNC_AUTO) {
        /* a mapping of its own (whole pages, zero-filled) whose protection the library can switch */
        const size_t pg = (size_t)sysconf(_SC_PAGESIZE);
        map_bytes = (want + pg - 1) / pg * pg;
        void *m = mmap(NULL, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) map_bytes = 0; else p = m;
    }
    if (!p) p = calloc(npart ? npart : 1, sizeof(seppart));
    if (!p) sep_error("%s at line %d: Couldn't allocate memory\n", (char *)__func__, __LINE__);
    /* The reference gives every atom an int[nneighb] row (12 kB/atom at SEP_NEIGHB).  The device owns
     * the list here; rows are not materialised (1 M atoms would need 12 GB of host memory). */
    (void)nneighb;
    for (size_t n = 0; n < npart; n++) {
        p[n].type = 'A';
        p[n].m = 1.0;
        p[n].molindex = -1;
        p[n].ldiff = 1.0;
        p[n].neighb = NULL;
        for (int k = 0; k < SEP_BOND; k++) p[n].bond[k] = -1;
        for (int k = 0; k < SEP_ANGLE; k++) p[n].angle[k] = -1;
        for (int k = 0; k < SEP_DIHED; k++) p[n].dihed[k] = -1;
    }
    sep_binding *b = sepb_register(p, npart);
    if (map_bytes) {
        b->managed = 1; b->map_base = p; b->map_bytes = map_bytes; b->prot = PROT_READ | PROT_WRITE;
        sep_install_segv();
    }
    return p;
}

void sep_close(seppart *ptr, size_t npart)
{
    (void)npart;
    sep_binding *b = sepb_find(ptr);
    void *map_base = NULL; size_t map_bytes = 0;
    if (b) {
        if (b->managed) { sepb_set_prot(b, PROT_READ | PROT_WRITE); map_base = b->map_base; map_bytes = b->map_bytes; }
        if (b->gpu) sepb_download(b, ~0u);      /* final state visible to anyone still holding a copy */
    }
    sepb_unregister(ptr);
    if (map_base) munmap(map_base, map_bytes); else free(ptr);
}

/* source/sepinit.c:69-109 */
seppart *sep_init_xyz(double *lbox, int *npart, const char *file, char verbose)
{
    if (verbose == 'v') fprintf(stdout, "Opening %s\n", file);
    FILE *fin = fopen(file, "r");
    if (!fin) sep_error("%s at line %d: Couldn't open file\n", (char *)__func__, __LINE__);
    if (fscanf(fin, "%d", npart) != 1)
        sep_error("%s at line %d: Error reading xyz file\n", (char *)__func__, __LINE__);
    seppart *p = sep_init((size_t)*npart, SEP_NUM_NEIGHB);
    if (fscanf(fin, "%lf%lf%lf\n", &lbox[0], &lbox[1], &lbox[2]) != 3)
        sep_error("%s at line %d: Error reading xyz file\n", (char *)__func__, __LINE__);
    for (int n = 0; n < *npart; n++) {
        seppart *a = &p[n];
        if (fscanf(fin, "%c%lf%lf%lf%lf%lf%lf%lf%lf\n", &a->type, &a->x[0], &a->x[1], &a->x[2],
                   &a->v[0], &a->v[1], &a->v[2], &a->m, &a->z) != 9)
            sep_error("%s at line %d: Error reading xyz file\n", (char *)__func__, __LINE__);
    }
    if (verbose == 'v') {
        fprintf(stdout, "Closing file\n");
        fprintf(stdout, "Number of particles: %d\n", *npart);
        SEP_FLUSH;
    }
    fclose(fin);
    return p;
}

/* source/sepmisc.c:454-463 */
int sep_nsubbox(double cf, double delta, double lbox)
{
    const double cut = cf + delta;
    return (int)(lbox / cut);
}

double sep_box_length(double dens, int npart, int ndim)
{
    return pow((double)npart / dens, (double)1.0 / ndim);
}

/* source/sepinit.c:246-305 */
sepsys sep_sys_setup(double lengthx, double lengthy, double lengthz,
                     double cf, double dt, size_t npart, size_t update)
{
    sepsys sys;
    memset(&sys, 0, sizeof sys);
    const double skinparam = 0.25;
    const double len[3] = {lengthx, lengthy, lengthz};
    static const char *dirname[3] = {"x", "y", "z"};
    sys.volume = lengthx * lengthy * lengthz;
    for (int k = 0; k < 3; k++) {
        sys.length[k] = len[k];
        sys.nsubbox[k] = sep_nsubbox(cf, skinparam, len[k]);
        if (update > 1 && sys.nsubbox[k] < 3) {
            char msg[128];
            snprintf(msg, sizeof msg, "sep_sys_setup: Number of subboxes in %s direction are less than three", dirname[k]);
            sep_warning(msg);
        }
        sys.lsubbox[k] = len[k] / sys.nsubbox[k];
    }
    sys.npart = (long)npart;
    sys.cf = cf;
    sys.dt = dt;
    sys.ndof = (unsigned)(3 * npart - 3);
    sys.tnow = 0.0;
    sys.nupdate_neighb = 0;
    sys.neighb_update = (unsigned)update;
    sys.neighb_flag = 1;
    sys.skin = skinparam;
    sys.molptr = calloc(1, sizeof(sepmolinfo));
    if (!sys.molptr) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    sys.omp_flag = false;
    sys.fun_cstate = 0;
    return sys;
}

void sep_free_sys(sepsys *ptr)
{
    sep_free_bonds(ptr->molptr);
    sep_free_angles(ptr->molptr);
    sep_free_dihedrals(ptr->molptr);
}

/* source/sepinit.c:317-357 : simple cubic lattice starting at (1,1,1) */
void sep_set_lattice(seppart *ptr, sepsys sys)
{
    int numb[3];
    double gap[3];
    const double dens = sys.npart / (sys.length[0] * sys.length[1] * sys.length[2]);
    for (int k = 0; k < 3; k++) {
        numb[k] = (int)ceil(pow(dens, 1. / 3.) * sys.length[k]);
        gap[k] = sys.length[k] / numb[k];
    }
    long n = 0;
    for (int iz = 0; iz < numb[2] && n < sys.npart; iz++)
        for (int iy = 0; iy < numb[1] && n < sys.npart; iy++)
            for (int ix = 0; ix < numb[0] && n < sys.npart; ix++) {
                ptr[n].x[0] = ix * gap[0] + 1.0;
                ptr[n].x[1] = iy * gap[1] + 1.0;
                ptr[n].x[2] = iz * gap[2] + 1.0;
                n++;
            }
    sepb_mark_host_dirty(ptr, SEPB_X);
}

/* shared body of sep_set_vel / sep_set_vel_seed / sep_set_vel_type (source/sepinit.c:114-243).
 * The glibc rand() stream is consumed in the reference's order so seeded runs start identically. */
static void draw_velocities(seppart *ptr, long npart, int use_type, char type, double temp)
{
    double smom[3] = {0.0, 0.0, 0.0}, sekin = 0.0;
    long ntype = 0;
    for (long n = 0; n < npart; n++) {
        if (use_type && ptr[n].type != type) continue;
        ntype++;
        for (int k = 0; k < 3; k++) {
            ptr[n].v[k] = sep_rand() - 0.5;
            smom[k] += ptr[n].v[k] * ptr[n].m;
            sekin += ptr[n].v[k] * ptr[n].v[k] * ptr[n].m;
        }
    }
    const int cnt = (int)ntype, ndim = 3;
    const double scale = sqrt(cnt * ndim * temp / sekin);
    for (long n = 0; n < npart; n++) {
        if (use_type && ptr[n].type != type) continue;
        for (int k = 0; k < 3; k++)
            ptr[n].v[k] = (ptr[n].v[k] - smom[k] / (ptr[n].m * cnt)) * scale;
    }
}

void sep_set_vel(seppart *ptr, double temp, sepsys sys)
{
    srand((unsigned)time(NULL));
    draw_velocities(ptr, sys.npart, 0, 0, temp);
    for (long n = 0; n < sys.npart; n++)
        for (int k = 0; k < 3; k++) {
            ptr[n].px[k] = ptr[n].x[k] + (sep_rand() - 0.5) * 0.01;
            ptr[n].pv[k] = ptr[n].v[k] + (sep_rand() - 0.5) * 0.01;
        }
    sepb_mark_host_dirty(ptr, SEPB_V | SEPB_PV);
}

void sep_set_vel_seed(seppart *ptr, double temp, unsigned int seed, sepsys sys)
{
    srand(seed);
    draw_velocities(ptr, sys.npart, 0, 0, temp);
    for (long n = 0; n < sys.npart; n++)
        for (int k = 0; k < 3; k++) {
            (void)sep_rand(); (void)sep_rand();          /* the reference draws and discards two numbers */
            ptr[n].px[k] = ptr[n].x[k];
            ptr[n].pv[k] = ptr[n].v[k];
            ptr[n].pa[k] = 0.0;
        }
    sepb_mark_host_dirty(ptr, SEPB_V | SEPB_PV | SEPB_PA);
}

void sep_set_vel_type(seppart *ptr, char type, double temp, unsigned int seed, sepsys sys)
{
    if (sep_count_type(ptr, type, (int)sys.npart) == 0) return;
    srand(seed);
    draw_velocities(ptr, sys.npart, 1, type, temp);
    for (long n = 0; n < sys.npart; n++) {
        if (ptr[n].type != type) continue;
        for (int k = 0; k < 3; k++) {
            ptr[n].px[k] = ptr[n].x[k] + (sep_rand() - 0.5) * 0.01;
            ptr[n].pv[k] = ptr[n].v[k] + (sep_rand() - 0.5) * 0.01;
        }
    }
    sepb_mark_host_dirty(ptr, SEPB_V | SEPB_PV);
}

/* ---------------------------------------------------------------------------------------------------
 * pair functions (source/sepmisc.c:115-164).  On the device these are recognised by address and
 * mapped to the Lennard-Jones kernel; the host versions exist for user code that calls them.
 * ------------------------------------------------------------------------------------------------- */
static double lj_family(double r2, char opt, double ushift)
{
    const double rri = 1.0 / r2, rri3 = rri * rri * rri;
    if (opt == 'f') return 48.0 * rri3 * (rri3 - 0.5) * rri;
    if (opt == 'u') return 4.0 * rri3 * (rri3 - 1.0) + ushift;
    return 0.0;
}
double sep_lj(double r2, char opt) { return lj_family(r2, opt, 0.0); }
double sep_lj_shift(double r2, char opt) { return lj_family(r2, opt, SEP_LJCF2); }
double sep_wca(double r2, char opt) { return lj_family(r2, opt, 1.0); }

/* ---------------------------------------------------------------------------------------------------
 * per-step resets (source/sepret.c:19-47, source/sepmisc.c:393-430)
 * ------------------------------------------------------------------------------------------------- */
void sep_reset_retval(sepret *r)
{
    r->epot = 0; r->ecoul = 0; r->ekin = 0; r->sumv2 = 0;
    double (*tens[])[3] = {r->pot_P, r->kin_P, r->P, r->pot_P_conservative, r->pot_P_random,
                           r->pot_P_dissipative, r->pot_P_bond, r->pot_P_mol, r->kin_P_mol, r->P_mol,
                           r->pot_T_mol, r->kin_T_mol, r->T_mol};
    for (size_t t = 0; t < sizeof tens / sizeof tens[0]; t++)
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) tens[t][k][kk] = 0.0;
    /* the device accumulators follow: every live context whose results go to this struct (or whose
     * target is not known yet) is reset in stream order */
    for (sep_binding *b = sepb_first(); b; b = b->next)
        if (b->gpu && (b->last_ret == r || b->last_ret == NULL))
            sepb_check(sepgpu_reset_ret(b->gpu), "sep_reset_retval");
}

void sep_reset_force(seppart *ptr, sepsys *sys)
{
    sepdd_allow();
    sep_binding *b = sepb_prepare(ptr, sys);
    sepb_check(sepgpu_reset_force(b->gpu), "sep_reset_force");
    sys->max_dist2 = 0.0;
    b->host_dirty &= ~SEPB_F;
    if (sepb_eager(b)) {
        for (long n = 0; n < sys->npart; n++) ptr[n].f[0] = ptr[n].f[1] = ptr[n].f[2] = 0.0;
        b->dev_dirty &= ~(SEPB_F | SEPB_A);
    } else {
        sepb_dev_newer(b, SEPB_F | SEPB_A);
    }
}

/* ---------------------------------------------------------------------------------------------------
 * scalar results (source/sepret.c:50-82)
 * ------------------------------------------------------------------------------------------------- */
static void refresh_ret(sepret *ret, sepsys *sys)
{
    /* in lazy mode the force calls have not copied their sums out yet */
    for (sep_binding *b = sepb_first(); b; b = b->next)
        if (b->gpu && b->last_ret == ret) sepb_pull_scalars(b, sys, ret, NULL);
}

void sep_pressure_tensor(sepret *ret, sepsys *sys)
{
    if (sep_sync_mode() == SEP_SYNC_LAZY) refresh_ret(ret, sys);
    const double ivol = 1.0 / sys->volume;
    ret->p = 0.0;
    for (int k = 0; k < 3; k++)
        for (int kk = 0; kk < 3; kk++) ret->P[k][kk] = (ret->kin_P[k][kk] + ret->pot_P[k][kk]) * ivol;
    for (int k = 0; k < 3; k++) ret->p += ret->P[k][k];
    ret->p /= 3.0;
}

double sep_get_pressure(sepret *ret, sepsys *sys)
{
    sep_pressure_tensor(ret, sys);
    return ret->p;
}

double sep_get_temperature(sepret *ret, sepsys *sys)
{
    if (sep_sync_mode() == SEP_SYNC_LAZY) refresh_ret(ret, sys);
    return 2.0 / (3.0 * sys->ndof) * ret->ekin;
}

/* ---------------------------------------------------------------------------------------------------
 * readers of the host array: bring it up to date first
 * ------------------------------------------------------------------------------------------------- */
double sep_eval_mom(seppart *ptr, int npart)
{
    sep_gpu_sync(ptr);
    double mom = 0.0;
    for (int n = 0; n < npart; n++)
        for (int k = 0; k < 3; k++) mom += ptr[n].v[k] * ptr[n].m;
    return mom / (npart * 3);
}

double sep_eval_mom_type(seppart *ptr, char type, int dir, int npart)
{
    sep_gpu_sync(ptr);
    double mom = 0.0;
    int cnt = 0;
    for (int n = 0; n < npart; n++)
        if (ptr[n].type == type) { mom += ptr[n].v[dir] * ptr[n].m; cnt++; }
    return cnt ? mom / cnt : 0.0;
}

int sep_count_type(seppart *ptr, char spec, int npart)
{
    int c = 0;
    for (int n = 0; n < npart; n++) c += ptr[n].type == spec;
    return c;
}

/* source/sepmisc.c:536-573 */
void sep_save_xyz(seppart *ptr, const char *partnames, const char *file, char *mode, sepsys *sys)
{
    sep_gpu_sync(ptr);
    const long ntype = (long)strlen(partnames);
    FILE *fout = fopen(file, mode);
    if (!fout) sep_error("%s at line %d: I couldn't open file\n", (char *)__func__, __LINE__);
    long ntotal = 0;
    for (long k = 0; k < ntype; k++) ntotal += sep_count_type(ptr, partnames[k], (int)sys->npart);
    fprintf(fout, "%lu\n%f %f %f\n", (unsigned long)ntotal, sys->length[0], sys->length[1], sys->length[2]);
    for (long n = 0; n < sys->npart; n++)
        for (long k = 0; k < ntype; k++)
            if (ptr[n].type == partnames[k])
                fprintf(fout, "%c %.15f %.15f %.15f %.15f %.15f %.15f %.15f %.15f\n", ptr[n].type,
                        ptr[n].x[0], ptr[n].x[1], ptr[n].x[2], ptr[n].v[0], ptr[n].v[1], ptr[n].v[2],
                        ptr[n].m, ptr[n].z);
    fclose(fout);
}

void sep_set_x0(seppart *ptr, int npart)
{
    sep_gpu_sync(ptr);
    for (int n = 0; n < npart; n++)
        for (int k = 0; k < 3; k++) ptr[n].x0[k] = ptr[n].x[k];
    sepb_mark_host_dirty(ptr, SEPB_X0);
}

void sep_set_xn(seppart *ptr, int npart)
{
    sep_gpu_sync(ptr);
    for (int n = 0; n < npart; n++)
        for (int k = 0; k < 3; k++) ptr[n].xn[k] = ptr[n].x[k];
    sepb_mark_host_dirty(ptr, SEPB_XN);
}

double sep_dist_ij(double *r, seppart *ptr, int i, int j, sepsys *sys)
{
    sep_gpu_sync(ptr);
    double r2 = 0.0;
    for (int k = 0; k < 3; k++) {
        r[k] = ptr[i].x[k] - ptr[j].x[k];
        sep_Wrap(r[k], sys->length[k]);
        r2 += r[k] * r[k];
    }
    return sqrt(r2);
}

void sep_eval_xtrue(seppart *ptr, sepsys *sys)
{
    sep_gpu_sync(ptr);
    for (long n = 0; n < sys->npart; n++)
        for (int k = 0; k < 3; k++) ptr[n].xtrue[k] = ptr[n].x[k] + ptr[n].crossings[k] * sys->length[k];
}

/* ---------------------------------------------------------------------------------------------------
 * setters (source/sepmisc.c:484-513, 1087-1129, 1196-1200)
 * ------------------------------------------------------------------------------------------------- */
void sep_set_charge(seppart *ptr, char type, double z, sepsys sys)
{
    for (long n = 0; n < sys.npart; n++) if (ptr[n].type == type) ptr[n].z = z;
    sepb_mark_host_dirty(ptr, SEPB_Z);
}

void sep_set_mass(seppart *ptr, char type, double m, sepsys sys)
{
    for (long n = 0; n < sys.npart; n++) if (ptr[n].type == type) ptr[n].m = m;
    sepb_mark_host_dirty(ptr, SEPB_M);
}

/* relabel `numb` randomly chosen 'A' atoms (source/sepmisc.c:484-513) */
void sep_set_type(seppart *ptr, char spec, int numb, sepsys *sys)
{
    int done = 0;
    long guard = 0;
    while (done < numb && guard++ < 1000L * (sys->npart + 1)) {
        long i = (long)(sep_rand() * sys->npart);
        if (ptr[i].type == 'A') { ptr[i].type = spec; done++; }
    }
    sepb_mark_host_dirty(ptr, SEPB_TYPE);
}

void sep_set_omp(unsigned nthreads, sepsys *sys)
{
    /* accepted for source compatibility: the device path does not use host threads */
    sys->omp_flag = true;
    sys->nthreads = nthreads;
}

void sep_set_skin(sepsys *sys, double value) { sys->skin = value; }
void sep_set_ndof(size_t ndof, sepsys *sys) { sys->ndof = (unsigned)ndof; }

/* source/sepmisc.c:1173-1192 */
void sep_reset_momentum(seppart *ptr, const char type, sepsys *sys)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepb_check(sepgpu_reset_momentum(b->gpu, type), "sep_reset_momentum");
    sepb_dev_newer(b, SEPB_V);
    if (sepb_eager(b)) sepb_download(b, SEPB_V);
}

/* source/sepmisc.c:994-1026 */
/* ---- box-changing callers (reference source/sepmisc.c:892-1083) ------------------------------------------
 * The host keeps the reference's bookkeeping of sys->length / nsubbox / lsubbox / volume line by line; the
 * positions are scaled on the device (sepgpu_scale_box), which also re-derives its sorted copy for the new box.
 * As in the reference the neighbour list is NOT invalidated by a box change. */
static void scale_box_on_device(sepatom *ptr, sepsys *sys, double sx, double sy, double sz, const char *who)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    const double sc[3] = {sx, sy, sz};
    sepb_check(sepgpu_scale_box(b->gpu, sc, sys->length), who);
    sepb_dev_newer(b, SEPB_X);
    if (sepb_eager(b)) sepb_download(b, SEPB_X);
}

static void resize_subbox(sepsys *sys, int k, double skin, const char *who)
{
    sys->nsubbox[k] = sep_nsubbox(sys->cf, skin, sys->length[k]);
    if (sys->nsubbox[k] < 3)
        sep_warning("%s: Number of subboxes in x direction are less than three", (char *)who);
    sys->lsubbox[k] = sys->length[k] / sys->nsubbox[k];
}

void sep_compress_box(sepatom *ptr, double rhoD, double xi, sepsys *sys)
{
    const double density = sys->npart / sys->volume;
    if (fabs(density - rhoD) < 1e-6) return;
    if (density > rhoD) xi = 1.0 / xi;
    for (int k = 0; k < 3; k++) {
        sys->length[k] *= xi;
        if (sys->length[k] < sys->cf * 2.0)
            sep_warning("sep_compress_box: Box length too small compared to the maximum cut-off");
    }
    scale_box_on_device(ptr, sys, xi, xi, xi, "sep_compress_box");
    if (sys->neighb_update != 0)
        for (int k = 0; k < 3; k++) resize_subbox(sys, k, sys->skin, "sep_compress_box");      /* :1014-1021, with skin */
    sys->volume = sys->length[0] * sys->length[1] * sys->length[2];
}

/* one direction; positions move by xi^(1/3) while the length moves by xi (:1037-1038) and the sub-boxes are
 * resized WITHOUT the skin (:1042) -- both reference quirks kept */
void sep_compress_box_dir(sepatom *ptr, double rhoD, double xi, int dir, sepsys *sys)
{
    if (sys->npart / sys->volume < rhoD) {
        sys->length[dir] *= xi;
        if (sys->length[dir] < sys->cf * 2.0)
            sep_warning("sep_compress_box_dir: Box length too small compared to the maximum cut-off");
        const double scale = pow(xi, 1.0 / 3.0);
        scale_box_on_device(ptr, sys, dir == 0 ? scale : 1.0, dir == 1 ? scale : 1.0, dir == 2 ? scale : 1.0, "sep_compress_box_dir");
        if (sys->neighb_update != 0) resize_subbox(sys, dir, 0.0, "sep_compress_box_dir");
        sys->volume = sys->length[0] * sys->length[1] * sys->length[2];
    }
}

void sep_compress_box_dir_length(sepatom *ptr, double length, double xi, int dir, sepsys *sys)
{
    if (sys->length[dir] > length) {
        sys->length[dir] *= xi;
        if (sys->length[dir] < sys->cf * 2.0)
            sep_warning("sep_compress_box_dir_length: Box length too small compared to the maximum cut-off");
        const double scale = pow(xi, 1.0 / 3.0);
        scale_box_on_device(ptr, sys, dir == 0 ? scale : 1.0, dir == 1 ? scale : 1.0, dir == 2 ? scale : 1.0, "sep_compress_box_dir_length");
        if (sys->neighb_update != 0) resize_subbox(sys, dir, 0.0, "sep_compress_box_dir_length");
        sys->volume = sys->length[0] * sys->length[1] * sys->length[2];
    }
}

/* Berendsen barostat along z (:892-914) and isotropic (:918-944) */
void sep_berendsen(sepatom *ptr, double Pd, double beta, sepret *ret, sepsys *sys)
{
    sep_pressure_tensor(ret, sys);
    const double xi = 1 - beta * sys->dt * (Pd - ret->p);
    sys->length[2] *= xi;
    const double scale = pow(xi, 1.0 / 3.0);
    scale_box_on_device(ptr, sys, 1.0, 1.0, scale, "sep_berendsen");
    sys->volume = sys->length[0] * sys->length[1] * sys->length[2];
    if (sys->neighb_update != 0) resize_subbox(sys, 2, 0.0, "sep_berendsen");
}

void sep_berendsen_iso(sepatom *ptr, double Pd, double beta, sepret *ret, sepsys *sys)
{
    sep_pressure_tensor(ret, sys);
    const double xi = 1 - beta * sys->dt * (Pd - ret->p);
    for (int k = 0; k < 3; k++) sys->length[k] *= xi;
    const double scale = pow(xi, 1.0 / 3.0);
    scale_box_on_device(ptr, sys, scale, scale, scale, "sep_berendsen_iso");
    sys->volume = sys->length[0] * sys->length[1] * sys->length[2];
    if (sys->neighb_update != 0)
        for (int k = 0; k < 3; k++) resize_subbox(sys, k, 0.0, "sep_berendsen_iso");
}

/* ---- per-type temperature relaxation and tethering springs (:357-390, :167-181, :645-670) ------------------ */
void sep_relax_temp(seppart *ptr, char type, double Td, double tau, sepsys *sys)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    double ekin = 0.0;
    sepb_check(sepgpu_relax_temp(b->gpu, &gs, type, Td, tau, &ekin), "sep_relax_temp");
    if (ekin < DBL_EPSILON)
        sep_warning("sep_relax_temp: Zero kinetic energy - check your the types.");
    sepb_dev_newer(b, SEPB_V);
    if (sepb_eager(b)) sepb_download(b, SEPB_V);
}

double sep_spring_x0(double r2, char opt)
{
    const double k = 500.0;
    double ret = 0.0;
    switch (opt) {
    case 'f': ret = -k; break;
    case 'u': ret = 0.5 * r2 * k; break;
    }
    return ret;
}

void sep_force_x0(seppart *ptr, char type, double (*fun)(double, char), sepsys *sys)
{
    if (fun != sep_spring_x0)
        sep_error("sep_force_x0: only sep_spring_x0 can run on the device (no CPU path)");
    sep_binding *b = sepb_find(ptr);
    if (!b) b = sepb_register(ptr, (size_t)sys->npart);
    if (!b->x0_on_device) { b->x0_on_device = 1; b->host_dirty |= SEPB_X0; }
    b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_force_x0(b->gpu, &gs, type, -fun(0.0, 'f')), "sep_force_x0");
    sepb_dev_newer(b, SEPB_F | SEPB_A);
    if (sep_sync_mode() == SEP_SYNC_FULL) sepb_download(b, SEPB_F);
}
