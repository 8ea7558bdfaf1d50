/* sep_mol.c -- molecular topology (host) and the bonded-force entry points (device).
 *
 * Topology reader and molecule table follow the behaviour of the reference's source/sepmol.c:22-369
 * and :590-685 (.top format: "[ bonds ]" / "[ angles ]" / "[ dihedrals ]" sections, one comment
 * line after each header, then "mol a b [c [d]] type" rows).  The force routines
 * sep_stretch_harmonic / sep_angle_harmonic / sep_angle_cossq / sep_torsion_Ryckaert
 * (source/sepmol.c:372-587) are device kernels (sepgpu_bonded.cu).
 */
#include <pthread.h>
#include "sep_host.h"
#include <float.h>

/* ---- .top reader ----------------------------------------------------------------------------------- */
typedef struct { unsigned *v; size_t n, cap; } uvec;

static void uvec_push(uvec *u, unsigned x)
{
    if (u->n == u->cap) {
        u->cap = u->cap ? 2 * u->cap : 1024;
        u->v = realloc(u->v, u->cap * sizeof(unsigned));
        if (!u->v) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    }
    u->v[u->n++] = x;
}

static int blank(const char *s)
{
    for (; *s; s++) if (*s != ' ' && *s != '\t' && *s != '\n' && *s != '\r') return 0;
    return 1;
}

/* Reads one section into rows of `ncol` unsigned numbers.  Returns the number of rows; 0 when the
 * section does not exist (the reference leaves empty lists and the flag set in that case). */
static size_t read_section(const char *file, const char *header, int ncol, uvec *out)
{
    FILE *fp = fopen(file, "r");
    if (!fp) sep_error("%s at line %d: Couldn't open file", (char *)__func__, __LINE__);
    char line[256];
    int found = 0;
    while (fgets(line, sizeof line, fp))
        if (strcmp(line, header) == 0) { found = 1; break; }
    size_t rows = 0;
    if (found && fgets(line, sizeof line, fp)) {           /* the line after the header is a comment */
        while (fgets(line, sizeof line, fp)) {
            if (line[0] == '[') break;
            if (blank(line)) continue;
            unsigned v[6];
            int got = sscanf(line, "%u%u%u%u%u%u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
            if (got != ncol)
                sep_error("%s at line %d: Format in top file not correct", (char *)__func__, __LINE__);
            for (int k = 0; k < ncol; k++) uvec_push(out, v[k]);
            rows++;
        }
    }
    fclose(fp);
    return rows;
}

static void add_partner(int *row, int width, int partner, const char *what)
{
    for (int k = 0; k < width; k++)
        if (row[k] == -1) { row[k] = partner; return; }
    sep_error("sep_read_topology_file: Index exceeds the allowed number of %s", (char *)what);
}

static void check_atom(unsigned a, long npart)
{
    if ((long)a >= npart) sep_error("sep_read_topology_file: atom index %d outside the system", (int)a);
}

void sep_read_topology_file(sepatom *aptr, const char *file, sepsys *sysptr, char opt)
{
    sepmolinfo *mp = sysptr->molptr;
    const long npart = sysptr->npart;
    uvec raw = {0};

    /* bonds: mol a b type (source/sepmol.c:22-135) */
    size_t nb = read_section(file, "[ bonds ]\n", 4, &raw);
    mp->flag_bonds = 1;
    mp->num_bonds = (unsigned)nb;
    mp->num_btypes = 0;
    mp->num_mols = 0;
    mp->blist = malloc(sizeof(unsigned) * 3 * (nb ? nb : 1));
    mp->blengths = calloc(nb ? nb : 1, sizeof(double));
    for (size_t n = 0; n < nb; n++) {
        const unsigned mol = raw.v[4 * n], a = raw.v[4 * n + 1], b = raw.v[4 * n + 2], t = raw.v[4 * n + 3];
        check_atom(a, npart); check_atom(b, npart);
        mp->blist[3 * n] = a; mp->blist[3 * n + 1] = b; mp->blist[3 * n + 2] = t;
        aptr[a].molindex = (int)mol;
        aptr[b].molindex = (int)mol;
        add_partner(aptr[a].bond, SEP_BOND, (int)b, "bonds");
        add_partner(aptr[b].bond, SEP_BOND, (int)a, "bonds");
        if (t > mp->num_btypes) mp->num_btypes = t;
        if (mol > mp->num_mols) mp->num_mols = mol;
    }
    if (nb) { mp->num_btypes++; mp->num_mols++; }
    if (opt == 'v' && nb) {
        printf("Succesfully read 'bond' section in file %s -> ", file);
        printf("Found %d molecules, %d bond(s) and %d bond type(s).\n", mp->num_mols, mp->num_bonds, mp->num_btypes);
    }

    /* angles: mol a b c type (source/sepmol.c:147-245) */
    raw.n = 0;
    size_t na = read_section(file, "[ angles ]\n", 5, &raw);
    mp->flag_angles = 1;
    mp->num_angles = (unsigned)na;
    mp->num_atypes = 0;
    mp->alist = malloc(sizeof(unsigned) * 4 * (na ? na : 1));
    mp->angles = calloc(na ? na : 1, sizeof(double));
    for (size_t n = 0; n < na; n++) {
        const unsigned a = raw.v[5 * n + 1], b = raw.v[5 * n + 2], c = raw.v[5 * n + 3], t = raw.v[5 * n + 4];
        check_atom(a, npart); check_atom(b, npart); check_atom(c, npart);
        mp->alist[4 * n] = a; mp->alist[4 * n + 1] = b; mp->alist[4 * n + 2] = c; mp->alist[4 * n + 3] = t;
        add_partner(aptr[a].angle, SEP_ANGLE, (int)b, "angles"); add_partner(aptr[a].angle, SEP_ANGLE, (int)c, "angles");
        add_partner(aptr[b].angle, SEP_ANGLE, (int)a, "angles"); add_partner(aptr[b].angle, SEP_ANGLE, (int)c, "angles");
        add_partner(aptr[c].angle, SEP_ANGLE, (int)a, "angles"); add_partner(aptr[c].angle, SEP_ANGLE, (int)b, "angles");
        if (t > mp->num_atypes) mp->num_atypes = t;
    }
    if (na) mp->num_atypes++;
    if (opt == 'v' && na) {
        printf("Succesfully read 'angles' section in file %s -> ", file);
        printf("Found %d angles(s) and %d bond angles(s).\n", mp->num_angles, mp->num_atypes);
    }

    /* dihedrals: mol a b c d type (source/sepmol.c:256-358) */
    raw.n = 0;
    size_t nd = read_section(file, "[ dihedrals ]\n", 6, &raw);
    mp->flag_dihedrals = 1;
    mp->num_dihedrals = (unsigned)nd;
    mp->num_dtypes = 0;
    mp->dlist = malloc(sizeof(unsigned) * 5 * (nd ? nd : 1));
    mp->dihedrals = calloc(nd ? nd : 1, sizeof(double));
    for (size_t n = 0; n < nd; n++) {
        unsigned q[4];
        for (int r = 0; r < 4; r++) { q[r] = raw.v[6 * n + 1 + r]; check_atom(q[r], npart); mp->dlist[5 * n + r] = q[r]; }
        const unsigned t = raw.v[6 * n + 5];
        mp->dlist[5 * n + 4] = t;
        for (int r = 0; r < 4; r++)
            for (int s = 0; s < 4; s++)
                if (s != r) add_partner(aptr[q[r]].dihed, SEP_DIHED, (int)q[s], "dihedrals");
        if (t > mp->num_dtypes) mp->num_dtypes = t;
    }
    if (nd) mp->num_dtypes++;
    if (opt == 'v' && nd) {
        printf("Succesfully read 'dihedrals' section in file %s -> ", file);
        printf("Found %d dihedrals(s) and %d dihedral types(s).\n", mp->num_dihedrals, mp->num_dtypes);
    }
    free(raw.v);
    sepb_mark_host_dirty(aptr, SEPB_MOL | SEPB_EXCL | SEPB_TOPO);
    sep_binding *bd = sepb_find(aptr);
    if (bd) bd->molptr = mp;
}

void sep_free_bonds(sepmolinfo *p) { if (p && p->flag_bonds == 1) { free(p->blist); free(p->blengths); p->flag_bonds = 0; } }
void sep_free_angles(sepmolinfo *p) { if (p && p->flag_angles == 1) { free(p->alist); free(p->angles); p->flag_angles = 0; } }
void sep_free_dihedrals(sepmolinfo *p) { if (p && p->flag_dihedrals == 1) { free(p->dlist); free(p->dihedrals); p->flag_dihedrals = 0; } }

/* ---- molecule table (source/sepmol.c:590-685) -------------------------------------------------------- */
sepmol *sep_init_mol(sepatom *atom, sepsys *sys)
{
    const unsigned nmol = sys->molptr->num_mols;
    sepmol *mols = calloc(nmol ? nmol : 1, sizeof(sepmol));
    int *count = sep_vector_int(nmol ? nmol : 1);
    if (!mols) sep_error("%s at %d: Couldn't allocate memory", (char *)__func__, __LINE__);
    for (long n = 0; n < sys->npart; n++)
        if (atom[n].molindex > -1) count[atom[n].molindex]++;
    sys->molptr->max_nuau = 0;
    for (unsigned i = 0; i < nmol; i++) {
        mols[i].nuau = (unsigned)count[i];
        mols[i].type = 'A';
        mols[i].index = malloc(sizeof(int) * (count[i] ? count[i] : 1));
        if (!mols[i].index) sep_error("%s at line %d: Couldn't allocate memory", (char *)__func__, __LINE__);
        if (mols[i].nuau > sys->molptr->max_nuau) sys->molptr->max_nuau = mols[i].nuau;
        count[i] = 0;
    }
    for (long n = 0; n < sys->npart; n++) {
        const int a = atom[n].molindex;
        if (a > -1) { mols[a].index[count[a]++] = (int)n; mols[a].m += atom[n].m; }
    }
    free(count);
    if (nmol <= SEP_MAX_NUM_MOL) {
        sys->molptr->flag_Fij = 1;
        sys->molptr->Fij = sep_tensor_float(nmol, nmol, 3);
    } else {
        sys->molptr->flag_Fij = 0;
        sep_warning("Molecular force/torque evaluation disabled");
    }
    return mols;
}

void sep_free_mol(sepmol *ptr, sepsys *sys)
{
    const unsigned nmol = sys->molptr->num_mols;
    for (unsigned n = 0; n < nmol; n++) {
        free(ptr[n].index);
        if (ptr[n].shake_flag == 1) free(ptr[n].blength);
    }
    free(ptr);
    if (sys->molptr->flag_Fij == 1) sep_free_tensor_float(sys->molptr->Fij, nmol, nmol);
}

/* source/sepmol.c:688-744: centre of mass by unfolding consecutive atoms of a molecule */
void sep_mol_cm(seppart *ptr, sepmol *mol, sepsys *sys)
{
    sep_gpu_sync(ptr);
    for (unsigned n = 0; n < sys->molptr->num_mols; n++) {
        if (mol[n].nuau == 0) continue;
        const unsigned first = (unsigned)mol[n].index[0];
        double u[3] = {ptr[first].x[0], ptr[first].x[1], ptr[first].x[2]};
        double cm[3];
        for (int k = 0; k < 3; k++) cm[k] = ptr[first].m * u[k];
        for (unsigned m = 1; m < mol[n].nuau; m++) {
            for (int k = 0; k < 3; k++) {
                double r = ptr[first + m].x[k] - ptr[first + m - 1].x[k];
                sep_Wrap(r, sys->length[k]);
                u[k] = u[k] + r;
            }
            for (int k = 0; k < 3; k++) cm[k] += ptr[first + m].m * u[k];
        }
        for (int k = 0; k < 3; k++) {
            cm[k] = cm[k] / mol[n].m;
            sep_Periodic(cm[k], sys->length[k]);
            mol[n].x[k] = cm[k];
        }
    }
}

void sep_mol_velcm(seppart *atom, sepmol *mol, sepsys *sys)
{
    sep_gpu_sync(atom);
    for (unsigned i = 0; i < sys->molptr->num_mols; i++) {
        for (int k = 0; k < 3; k++) mol[i].v[k] = 0.0;
        for (unsigned n = 0; n < mol[i].nuau; n++) {
            const int a = mol[i].index[n];
            for (int k = 0; k < 3; k++) mol[i].v[k] += atom[a].v[k] * atom[a].m;
        }
        for (int k = 0; k < 3; k++) mol[i].v[k] /= mol[i].m;
    }
}

/* ---- per-molecule derived quantities used by the molecular samplers (host, on synchronised atoms[]) ---------- */
/* unwrapped centre of mass, source/sepmol.c:1127-1148 */
void sep_mol_eval_xtrue(seppart *ptr, sepmol *mol, sepsys sys)
{
    sep_eval_xtrue(ptr, &sys);
    for (unsigned n = 0; n < sys.molptr->num_mols; n++) {
        double acc[3] = {0.0, 0.0, 0.0};
        for (unsigned m = 0; m < mol[n].nuau; m++) {
            const int i = mol[n].index[m];
            for (int k = 0; k < 3; k++) acc[k] += ptr[i].m * ptr[i].xtrue[k];
        }
        for (int k = 0; k < 3; k++) mol[n].xtrue[k] = acc[k] / mol[n].m;
    }
}

static double det3(double a[3][3])
{
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

/* angular momentum s about the centre of mass, inertia tensor and (unless safe) angular velocity w = I^-1 s;
 * a singular tensor (linear molecule) falls back to the mean principal moment, source/sepmol.c:750-817 */
void sep_mol_spin(sepatom *atom, sepmol *mol, sepsys *sys, bool safe)
{
    sep_mol_cm(atom, mol, sys);
    for (unsigned i = 0; i < sys->molptr->num_mols; i++) {
        double in[3][3] = {{0}}, s[3] = {0.0, 0.0, 0.0}, w[3] = {0.0, 0.0, 0.0};
        for (unsigned n = 0; n < mol[i].nuau; n++) {
            const int ia = mol[i].index[n];
            if (ia == -1) break;
            const double mia = atom[ia].m;
            double d[3], p[3];
            for (int k = 0; k < 3; k++) {
                d[k] = atom[ia].x[k] - mol[i].x[k];
                sep_Wrap(d[k], sys->length[k]);
                p[k] = atom[ia].v[k] * mia;
            }
            s[0] += d[1] * p[2] - d[2] * p[1];
            s[1] += d[2] * p[0] - d[0] * p[2];
            s[2] += d[0] * p[1] - d[1] * p[0];
            in[0][0] += mia * (sep_Sq(d[1]) + sep_Sq(d[2]));
            in[1][1] += mia * (sep_Sq(d[0]) + sep_Sq(d[2]));
            in[2][2] += mia * (sep_Sq(d[0]) + sep_Sq(d[1]));
            in[0][1] -= mia * d[0] * d[1];
            in[0][2] -= mia * d[0] * d[2];
            in[1][2] -= mia * d[1] * d[2];
        }
        in[1][0] = in[0][1]; in[2][0] = in[0][2]; in[2][1] = in[1][2];
        if (!safe) {
            const double det = det3(in);
            if (fabs(det) < DBL_EPSILON) {
                const double Ip = (in[0][0] + in[1][1] + in[2][2]) / 3.0;       /* mean eigenvalue = trace / 3 */
                for (int k = 0; k < 3; k++) w[k] = s[k] / Ip;
                mol[i].method_w = 0;
            } else {                                                           /* Cramer's rule on the 3x3 system */
                for (int c = 0; c < 3; c++) {
                    double t[3][3];
                    for (int r = 0; r < 3; r++)
                        for (int q = 0; q < 3; q++) t[r][q] = q == c ? s[r] : in[r][q];
                    w[c] = det3(t) / det;
                }
                mol[i].method_w = 1;
            }
        }
        for (int k = 0; k < 3; k++) {
            if (!safe) mol[i].w[k] = w[k];
            mol[i].s[k] = s[k];
            for (int kk = 0; kk < 3; kk++) mol[i].inertia[k][kk] = in[k][kk];
        }
    }
}

/* electric dipole from the mean positions of the positive and of the negative sites, source/sepmol.c:1036-1083 */
void sep_mol_dipoles(seppart *atom, sepmol *mol, sepsys *sys)
{
    sep_mol_cm(atom, mol, sys);
    for (unsigned i = 0; i < sys->molptr->num_mols; i++) {
        double sumz = 0.0, rpos[3] = {0.0, 0.0, 0.0}, rneg[3] = {0.0, 0.0, 0.0};
        int npos = 0, nneg = 0;
        for (unsigned n = 0; n < mol[i].nuau; n++) {
            const int ia = mol[i].index[n];
            double off[3] = {0.0, 0.0, 0.0};
            for (int k = 0; k < 3; k++) {
                const double d = atom[ia].x[k] - mol[i].x[k];
                if (fabs(d) > 0.5 * sys->length[k]) off[k] = d > 0.0 ? -sys->length[k] : sys->length[k];
            }
            if (atom[ia].z > 0.0) {
                for (int k = 0; k < 3; k++) rpos[k] += atom[ia].x[k] + off[k];
                sumz += atom[ia].z;
                npos++;
            } else if (atom[ia].z < 0.0) {
                for (int k = 0; k < 3; k++) rneg[k] += atom[ia].x[k] + off[k];
                nneg++;
            }
        }
        if (nneg > 0 && npos > 0)
            for (int k = 0; k < 3; k++) mol[i].pel[k] = sumz * (rpos[k] / npos - rneg[k] / nneg);
    }
}

double sep_average_bondlengths(int type, sepsys *sys)
{
    const sepmolinfo *mp = sys->molptr;
    sep_binding *b = sepb_find_mol(mp);
    if (b && b->gpu)
        sepb_check(sepgpu_get_bonded_values(b->gpu, mp->blengths, NULL, NULL), "sep_average_bondlengths");
    double sum = 0.0; int cnt = 0;
    for (unsigned n = 0; n < mp->num_bonds; n++)
        if ((int)mp->blist[3 * n + 2] == type) { sum += mp->blengths[n]; cnt++; }
    return sum / cnt;
}

/* ---- bonded forces: device kernels --------------------------------------------------------------------- */
static sep_binding *bonded_prepare(sepatom *ptr, sepsys *sys, sepgpu_sys *gs, const char *who)
{
    if (!sys->molptr || !sys->molptr->flag_bonds)
        sep_error("%s: no topology loaded (sep_read_topology_file)", (char *)who);
    sep_binding *b = sepb_prepare(ptr, sys);
    sepb_fill_sys(sys, gs);
    return b;
}

void sep_stretch_harmonic(sepatom *aptr, int type, const double lbond, const double ks, sepsys *sys, sepret *ret)
{
    sepgpu_sys gs;
    sep_binding *b = bonded_prepare(aptr, sys, &gs, "sep_stretch_harmonic");
    sepb_check(sepgpu_stretch_harmonic(b->gpu, &gs, type, lbond, ks), "sep_stretch_harmonic");
    sepb_after_force(b, sys, ret);
}

void sep_angle_harmonic(sepatom *ptr, int type, const double angle0, const double k, sepsys *sys, sepret *ret)
{
    sepgpu_sys gs;
    sep_binding *b = bonded_prepare(ptr, sys, &gs, "sep_angle_harmonic");
    sepb_check(sepgpu_angle_harmonic(b->gpu, &gs, type, angle0, k), "sep_angle_harmonic");
    sepb_after_force(b, sys, ret);
}

void sep_angle_cossq(sepatom *ptr, int type, const double angle0, const double k, sepsys *sys, sepret *ret)
{
    sepgpu_sys gs;
    sep_binding *b = bonded_prepare(ptr, sys, &gs, "sep_angle_cossq");
    sepb_check(sepgpu_angle_cossq(b->gpu, &gs, type, angle0, k), "sep_angle_cossq");
    sepb_after_force(b, sys, ret);
}

void sep_torsion_Ryckaert(sepatom *ptr, int type, const double g[6], sepsys *sys, sepret *ret)
{
    sepgpu_sys gs;
    sep_binding *b = bonded_prepare(ptr, sys, &gs, "sep_torsion_Ryckaert");
    sepb_check(sepgpu_torsion_ryckaert(b->gpu, &gs, type, g), "sep_torsion_Ryckaert");
    sepb_after_force(b, sys, ret);
}

/* ---- molecular pressure tensor (source/sepmol.c:913-963, source/sepret.c:85-102) --------------------------
 * The molecule-molecule force table Fij is accumulated on the device by the pair kernels (FP64 there,
 * float in the reference); it is copied into the caller-visible float table when the tensor is evaluated. */
/* ---- prg5's helpers: one term kind into the caller's matrix (reference source/sepomp.c:179-329) ---------------------- */
static pthread_mutex_t g_omp_mu = PTHREAD_MUTEX_INITIALIZER;      /* prg5 calls them from two OpenMP sections at once */

static void omp_side(double **ftot, sepatom *ptr, sepsys *sys, int kind, int type, const double *par, const char *who)
{
    pthread_mutex_lock(&g_omp_mu);
    sepgpu_sys gs;
    sep_binding *b = bonded_prepare(ptr, sys, &gs, who);
    if (b->dd) sep_error("%s: not available with SEP_NGPU", (char *)who);
    const size_t n = (size_t)sys->npart;
    double *flat = malloc(sizeof(double) * 3 * (n ? n : 1));
    if (!flat) sep_error("%s: out of memory", (char *)who);
    sepb_check(sepgpu_bonded_side(b->gpu, &gs, kind, type, par, flat), who);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) ftot[i][k] += flat[3 * i + k];
    free(flat);
    pthread_mutex_unlock(&g_omp_mu);
}

void sep_omp_bond(double **ftot, sepatom *aptr, int type, const double lbond, const double ks, sepsys *sys)
{
    const double par[2] = {lbond, ks};
    omp_side(ftot, aptr, sys, 0, type, par, "sep_omp_bond");
}

void sep_omp_angle(double **ftot, sepatom *ptr, int type, const double angle0, const double k, sepsys *sys)
{
    const double par[2] = {angle0, k};
    omp_side(ftot, ptr, sys, 1, type, par, "sep_omp_angle");
}

void sep_omp_torsion(double **ftot, sepatom *ptr, int type, const double g[6], sepsys *sys)
{
    omp_side(ftot, ptr, sys, 2, type, g, "sep_omp_torsion");
}

void sep_reset_force_mol(sepsys *sys)
{
    sys->fun_cstate = 0;
    if (sys->molptr->flag_Fij == 0)
        sep_error("%s: Tried to reset mol force, but flag is zero", (char *)__func__);
    /* the host copy too (source/sepmisc.c:418-422): it is what callers read before any device call has bound the table */
    for (unsigned i = 0; i < sys->molptr->num_mols; i++)
        for (unsigned j = 0; j < sys->molptr->num_mols; j++)
            for (int k = 0; k < 3; k++) sys->molptr->Fij[i][j][k] = 0.0f;
    sep_binding *b = sepb_find_mol(sys->molptr);
    if (b && b->gpu) {
        sepb_check(sepgpu_fij_enable(b->gpu, (int)sys->molptr->num_mols), "sep_reset_force_mol");
        sepb_check(sepgpu_fij_reset(b->gpu), "sep_reset_force_mol");
    }
}

void sep_eval_mol_pressure_tensor(sepatom *atoms, sepmol *mols, sepret *ret, sepsys *sys)
{
    if (sys->molptr->flag_Fij == 0) return;
    const int nmol = (int)sys->molptr->num_mols;
    float ***Fij = sys->molptr->Fij;
    sep_binding *b = sepb_find(atoms);
    if (b && b->gpu) {
        float *flat = malloc(sizeof(float) * 3 * (size_t)nmol * nmol);
        if (!flat) sep_error("%s at %d: Couldn't allocate memory", (char *)__func__, __LINE__);
        int rc = sepgpu_fij_get(b->gpu, flat);
        if (rc == 0)
            for (int i = 0; i < nmol; i++)
                for (int j = 0; j < nmol; j++)
                    for (int k = 0; k < 3; k++) Fij[i][j][k] = flat[((size_t)i * nmol + j) * 3 + k];
        free(flat);
    }
    sep_mol_cm(atoms, mols, sys);
    sep_mol_velcm(atoms, mols, sys);
    for (int k = 0; k < 3; k++)
        for (int kk = 0; kk < 3; kk++) ret->kin_P_mol[k][kk] = ret->pot_P_mol[k][kk] = 0.0;
    for (int i = 0; i < nmol; i++)
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->kin_P_mol[k][kk] += mols[i].m * mols[i].v[k] * mols[i].v[kk];
    for (int i = 0; i < nmol - 1; i++)
        for (int j = i + 1; j < nmol; j++) {
            double rij[3];
            for (int k = 0; k < 3; k++) {
                rij[k] = mols[i].x[k] - mols[j].x[k];
                sep_Wrap(rij[k], sys->length[k]);
            }
            for (int k = 0; k < 3; k++)
                for (int kk = 0; kk < 3; kk++) ret->pot_P_mol[k][kk] += Fij[i][j][k] * rij[kk];
        }
    /* P_mol and p_mol are filled here as well (source/sepmol.c:950-961) */
    const double ivol = 1.0 / sys->volume;
    ret->p_mol = 0.0;
    for (int k = 0; k < 3; k++) {
        for (int kk = 0; kk < 3; kk++) ret->P_mol[k][kk] = (ret->kin_P_mol[k][kk] + ret->pot_P_mol[k][kk]) * ivol;
        ret->p_mol += ret->P_mol[k][k];
    }
    ret->p_mol /= 3.0;
}

void sep_mol_pressure_tensor(sepatom *atoms, sepmol *mols, sepret *ret, sepsys *sys)
{
    /* source/sepret.c:85-102: evaluated again here, also when the table is switched off (then from whatever the
     * kinetic / potential parts hold) */
    const double ivol = 1.0 / sys->volume;
    sep_eval_mol_pressure_tensor(atoms, mols, ret, sys);
    ret->p_mol = 0.0;
    for (int k = 0; k < 3; k++) {
        for (int kk = 0; kk < 3; kk++) ret->P_mol[k][kk] = (ret->kin_P_mol[k][kk] + ret->pot_P_mol[k][kk]) * ivol;
        ret->p_mol += ret->P_mol[k][k];
    }
    ret->p_mol /= 3.0;
}
