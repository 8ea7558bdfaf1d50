/* sep_sampler.c -- run-time samplers (host post-processing of the synchronised atoms[] / sepret data).
 *
 * SURVEY.md section 8b keeps samplers as host C.  Implemented here, with the reference's file names and
 * output formats so that existing analysis scripts keep working (reference source/sepsampler.c):
 *   "sacf"   stress autocorrelation            -> sacf.dat                      (:280-358)
 *   "vacf"   velocity autocorrelation          -> vacf.dat                      (:658-722)
 *   "msd"    mean square displacement, non-Gaussian parameter, self-intermediate scattering
 *                                              -> msd-k.dat msd.dat msd-gaussparam.dat msd-incoherent.dat  (:471-655)
 *   "profs"  density / momentum / temperature profile along z -> profs.dat     (:1416-1525)
 *   "radial" radial distribution functions     -> radial_info.dat radial.dat   (:361-468)
 *   "msacf"  molecular stress autocorrelation  -> msacf.dat                    (:725-821)
 *   "mvacf"  molecular velocity autocorrelation-> mvacf.dat                    (:1757-1826)
 *   "gh"     generalised hydrodynamics: density, momentum, energy correlations at wave vectors (0, k, 0)
 *                                              -> gh-wavevector.dat gh-*-acf.dat gh-*-ccf.dat   (:823-1107)
 *   "mgh"    the molecular counterpart (centre-of-mass fields, angular momentum, dipole)
 *                                              -> mgh-wavevector.dat mgh-*-acf.dat mgh-*-ccf.dat (:1110-1415)
 * "mprofs", "mcacf", "mavacf", "mmsd" and "scatt" are accepted and record nothing (one warning each).
 *
 * Sampler feeds (SEP_SAMPLER_FEEDS=1, single-GPU runs): "vacf", "msd", "profs", "gh" and "radial" take their per-sample sums
 * from the device (include/sepgpu.h, sepgpu_feed_*) instead of reading atoms[] -- no download of the array per sample, and
 * none per step for "msd", whose boundary tracking is replaced by the device's own crossing counters.  Same files, same
 * numbers to printed precision ("radial" exactly: integer counts).  Off by default this round: the feed kernels were
 * written after the round's GPU budget was spent (tests/test_gpu_zzzz_feeds.py compares both paths on hardware).
 * Not taken over in feed mode: the side effect of "gh" on atoms[].xtrue (sep_eval_xtrue on the host copy).
 *
 * All correlation samplers share one block accumulator: lvec rows of ncol channels are collected, then every
 * channel's products x[t0] x[t0+t] are added to acf[t] and the file is rewritten.
 */
#include "sep_host.h"

#include <complex.h>
#include <math.h>

static int feeds_enabled(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SEP_SAMPLER_FEEDS");
        v = e && atoi(e) > 0 ? 1 : 0;
    }
    return v;
}

/* ---- block correlation accumulator ------------------------------------------------------------------- */
typedef struct {
    unsigned lvec, fill, nblocks, isample;
    size_t ncol;
    double dtsample;
    double *rows;          /* [lvec][ncol] */
    double *acf;           /* [nacf][lvec] */
    int nacf;              /* 1, or 2 when the channels are split into two groups of ncol/2 */
} sep_corr;

static sep_corr *corr_new(const char *who, int lvec, double tsample, double dt, size_t ncol, int nacf)
{
    sep_corr *c = calloc(1, sizeof *c);
    if (!c || lvec <= 0) sep_error("%s: Couldn't allocate memory", (char *)who);
    c->lvec = (unsigned)lvec;
    c->dtsample = tsample / lvec;
    c->isample = (unsigned)(int)(c->dtsample / dt);
    if ((int)(c->dtsample / dt) < 1) sep_error("%s: isample is too small - CHECK lvec argument", (char *)who);
    c->ncol = ncol;
    c->nacf = nacf;
    c->rows = calloc((size_t)lvec * ncol, sizeof(double));
    c->acf = calloc((size_t)lvec * nacf, sizeof(double));
    if (!c->rows || !c->acf) sep_error("%s: Couldn't allocate memory", (char *)who);
    return c;
}

static void corr_free(sep_corr *c)
{
    if (!c) return;
    free(c->rows); free(c->acf); free(c);
}

static double *corr_row(sep_corr *c) { return c->rows + (size_t)c->fill * c->ncol; }

/* returns 1 when a block was completed (acf updated, time origin restarted) */
static int corr_push(sep_corr *c)
{
    if (++c->fill < c->lvec) return 0;
    const size_t per = c->ncol / c->nacf;
    for (int a = 0; a < c->nacf; a++)
        for (size_t col = a * per; col < (a + 1) * per; col++)
            for (unsigned t = 0; t < c->lvec; t++) {
                double s = 0.0;
                for (unsigned t0 = 0; t0 + t < c->lvec; t0++)
                    s += c->rows[(size_t)t0 * c->ncol + col] * c->rows[(size_t)(t0 + t) * c->ncol + col];
                c->acf[(size_t)a * c->lvec + t] += s;
            }
    c->nblocks++;
    c->fill = 0;
    return 1;
}

/* ---- the sampler set --------------------------------------------------------------------------------- */
typedef struct {
    unsigned lvec, nsample, isample;
    char type;
    double *momc, *dens, *temp, *svel;
} sep_profile;

typedef struct {
    int lvec, isample, ntypes, ncomb, nsample;
    long *hist;            /* [lvec][ncomb] */
    char types[256];
} sep_rdf;

typedef struct {
    int lvec, fill, nsample, isample, nk, npart;
    char type;
    int logmode, logcounter;
    double *time, *msd, *msdsq, *k;
    double *fs;            /* [lvec][nk], real part of the self-intermediate scattering sum */
    double *prev, *pos0;   /* [npart][3] */
    int *cross;            /* [npart][3] */
} sep_msdacc;

/* ---- wave-vector dependent correlators ("gh", "mgh") ------------------------------------------------------
 * A set of Fourier fields A_f(+k, t) and A_f(-k, t) at wave vectors (0, k_n, 0) is recorded for lvec sample
 * times; when the block is full every listed product  A(t0) B(t0 + t)  (or A(t0 + t) B(t0)) is added to its
 * correlation function and all files are rewritten (real and imaginary part per wave vector). */
#define FK_MAXFIELD 10
typedef struct { const char *file; int fa, sa, la, fb, sb, lb; int unsafe_only; } fk_corr_spec;   /* field, sign (0:+k 1:-k), lagged? */

typedef struct {
    unsigned lvec, fill, nsample, isample, nwave, ncalls;
    int nfield, ncorr, safe;
    const fk_corr_spec *spec;
    double dtsample, avekin;
    double *k;
    double complex *field;      /* [nfield][2][lvec][nwave] */
    double complex *corr;       /* [ncorr][lvec][nwave]     */
} sep_fkacc;

static double complex *fk_field(sep_fkacc *g, int f, int sign, unsigned t)
{
    return g->field + (((size_t)f * 2 + sign) * g->lvec + t) * g->nwave;
}

static sep_fkacc *fk_new(const char *who, const char *kfile, int lvec, double tsample, double dt, int nwave, double Ldir,
                         int nfield, const fk_corr_spec *spec, int ncorr)
{
    sep_fkacc *g = calloc(1, sizeof *g);
    if (!g || lvec <= 0 || nwave <= 0) sep_error("%s: Couldn't allocate memory", (char *)who);
    g->lvec = (unsigned)lvec; g->nwave = (unsigned)nwave; g->nfield = nfield; g->spec = spec; g->ncorr = ncorr; g->safe = 1;
    g->dtsample = tsample / lvec;
    g->isample = (unsigned)(int)(g->dtsample / dt);
    if ((int)(g->dtsample / dt) < 1) sep_error("%s: isample is too small - CHECK lvec argument", (char *)who);
    g->k = sep_vector((size_t)nwave);
    g->field = calloc((size_t)nfield * 2 * lvec * nwave, sizeof(double complex));
    g->corr = calloc((size_t)ncorr * lvec * nwave, sizeof(double complex));
    FILE *fout = fopen(kfile, "w");
    if (!fout || !g->field || !g->corr) sep_error("%s: Couldn't open file k.dat", (char *)who);
    for (int n = 1; n <= nwave; n++) {
        g->k[n - 1] = 2 * SEP_PI * n / Ldir;
        fprintf(fout, "%f\n", g->k[n - 1]);
    }
    fclose(fout);
    return g;
}

static void fk_free(sep_fkacc *g)
{
    if (!g) return;
    free(g->k); free(g->field); free(g->corr); free(g);
}

/* called after the fields of sample time g->fill were stored */
static void fk_push(sep_fkacc *g, const char *who, double volume)
{
    if (++g->fill < g->lvec) return;
    for (int c = 0; c < g->ncorr; c++) {
        const fk_corr_spec *sp = &g->spec[c];
        if (sp->unsafe_only && g->safe) continue;
        for (unsigned k = 0; k < g->nwave; k++)
            for (unsigned n = 0; n < g->lvec; n++) {
                double complex sum = 0.0;
                for (unsigned nn = 0; nn + n < g->lvec; nn++)
                    sum += fk_field(g, sp->fa, sp->sa, sp->la ? nn + n : nn)[k] * fk_field(g, sp->fb, sp->sb, sp->lb ? nn + n : nn)[k];
                g->corr[((size_t)c * g->lvec + n) * g->nwave + k] += sum;
            }
    }
    g->nsample++;
    for (int c = 0; c < g->ncorr; c++) {
        FILE *fout = fopen(g->spec[c].file, "w");
        if (!fout) sep_error("%s: Error opening files", (char *)who);
        if (!(g->spec[c].unsafe_only && g->safe))
            for (unsigned n = 0; n < g->lvec; n++) {
                const double fac = 1.0 / (g->nsample * volume * (g->lvec - n));
                fprintf(fout, "%f ", n * g->dtsample);
                for (unsigned k = 0; k < g->nwave; k++) {
                    const double complex v = g->corr[((size_t)c * g->lvec + n) * g->nwave + k] * fac;
                    fprintf(fout, "%f %f ", creal(v), cimag(v));
                }
                fprintf(fout, "\n");
            }
        fclose(fout);
    }
    g->fill = 0;
}

/* atomic: density, transverse / longitudinal momentum, energy, auxiliary X (source/sepsampler.c:926-1107) */
enum { GH_RHO, GH_TV, GH_LV, GH_E, GH_X, GH_NFIELD };
static const fk_corr_spec GH_CORR[] = {
    {"gh-trans-momentum-acf.dat", GH_TV, 1, 0, GH_TV, 0, 1, 0},
    {"gh-long-momentum-acf.dat", GH_LV, 1, 0, GH_LV, 0, 1, 0},
    {"gh-rho-acf.dat", GH_RHO, 0, 0, GH_RHO, 1, 1, 0},
    {"gh-energy-acf.dat", GH_E, 0, 0, GH_E, 1, 1, 0},
    {"gh-rho-energy-ccf.dat", GH_RHO, 0, 0, GH_E, 1, 1, 0},
    {"gh-rho-long-momentum-ccf.dat", GH_RHO, 0, 0, GH_LV, 1, 1, 0},
    {"gh-energy-long-momentum-ccf.dat", GH_E, 0, 0, GH_LV, 1, 1, 0},
    {"gh-energy-rho-ccf.dat", GH_RHO, 0, 1, GH_E, 1, 0, 0},
    {"gh-long-momentum-rho-ccf.dat", GH_RHO, 0, 1, GH_LV, 1, 0, 0},
    {"gh-long-momentum-energy-ccf.dat", GH_E, 0, 1, GH_LV, 1, 0, 0},
    {"gh-X-acf.dat", GH_X, 0, 0, GH_LV, 1, 1, 0},
};

/* the same fields from the device's Fourier sums (sepgpu_feed_fourier): S1 = sum e, Sm = sum m e, ... with e = exp(+i k y);
 * the -k fields are the conjugates, except that the auxiliary field keeps -i k on both sides (:975-977) */
static void sample_gh_feed(sep_fkacc *g, sep_binding *b, sepsys *sys)
{
    const unsigned t = g->fill;
    double *o = sep_vector((size_t)16 * g->nwave);
    sepb_check(sepgpu_feed_fourier(b->gpu, sys->length[1], (int)g->nwave, g->k, o), "gh sampler feed");
    g->ncalls++;
    g->avekin += o[14] / (sys->npart * g->ncalls);
    for (unsigned n = 0; n < g->nwave; n++) {
        const double *q = o + 16 * (size_t)n;
        const double complex S1 = q[0] + I * q[1], Sm = q[2] + I * q[3], Stv = q[4] + I * q[5], Slv = q[6] + I * q[7],
                             Se = q[8] + I * q[9], Sa = q[10] + I * q[11], Svv = q[12] + I * q[13];
        const double complex E = Se - g->avekin * S1;
        fk_field(g, GH_RHO, 0, t)[n] = Sm;  fk_field(g, GH_RHO, 1, t)[n] = conj(Sm);
        fk_field(g, GH_TV, 0, t)[n] = Stv;  fk_field(g, GH_TV, 1, t)[n] = conj(Stv);
        fk_field(g, GH_LV, 0, t)[n] = Slv;  fk_field(g, GH_LV, 1, t)[n] = conj(Slv);
        fk_field(g, GH_E, 0, t)[n] = E;     fk_field(g, GH_E, 1, t)[n] = conj(E);
        fk_field(g, GH_X, 0, t)[n] = Sa - I * g->k[n] * Svv;
        fk_field(g, GH_X, 1, t)[n] = conj(Sa) - I * g->k[n] * conj(Svv);
    }
    free(o);
    fk_push(g, "sep_gh_sampler", sys->volume);
}

static void sample_gh(sep_fkacc *g, seppart *atoms, sepsys *sys)
{
    const int kdir = 1, tdir = 0;                 /* wave vector (0, k, 0), transverse direction x (:848-849) */
    const unsigned t = g->fill;
    sep_eval_xtrue(atoms, sys);
    double sumv2 = 0.0;
    for (long m = 0; m < sys->npart; m++)
        for (int kk = 0; kk < 3; kk++) sumv2 += 0.5 * atoms[m].m * sep_Sq(atoms[m].v[kk]);
    g->ncalls++;
    g->avekin += sumv2 / (sys->npart * g->ncalls);             /* the reference's running value, :940 */
    for (unsigned n = 0; n < g->nwave; n++) {
        double complex acc[GH_NFIELD][2];
        for (int f = 0; f < GH_NFIELD; f++) acc[f][0] = acc[f][1] = 0.0;
        for (long m = 0; m < sys->npart; m++) {
            const double mass = atoms[m].m;
            const double complex kf[2] = {cexp(I * g->k[n] * atoms[m].xtrue[kdir]), cexp(-I * g->k[n] * atoms[m].xtrue[kdir])};
            double ekin = 0.0;
            for (int kk = 0; kk < 3; kk++) ekin += 0.5 * mass * sep_Sq(atoms[m].v[kk]);
            const double complex aux = mass * (atoms[m].a[1] - I * g->k[n] * sep_Sq(atoms[m].v[1]));
            for (int sgn = 0; sgn < 2; sgn++) {
                acc[GH_RHO][sgn] += mass * kf[sgn];
                acc[GH_TV][sgn] += mass * atoms[m].v[tdir] * kf[sgn];
                acc[GH_LV][sgn] += mass * atoms[m].v[kdir] * kf[sgn];
                acc[GH_E][sgn] += (ekin - g->avekin) * kf[sgn];
                acc[GH_X][sgn] += aux * kf[sgn];
            }
        }
        for (int f = 0; f < GH_NFIELD; f++)
            for (int sgn = 0; sgn < 2; sgn++) fk_field(g, f, sgn, t)[n] = acc[f][sgn];
    }
    fk_push(g, "sep_gh_sampler", sys->volume);
}

/* molecular: centre-of-mass density / momentum / energy, angular momentum, dipole (source/sepsampler.c:1231-1415).
 * The wave factor uses the WRAPPED centre of mass (mols[m].x[1], :1270) although xtrue is evaluated, and the
 * momentum / angular-momentum cross function pairs s_z(+k) with m v_x(-k) (:1300-1301): both kept. */
enum { MGH_RHO, MGH_TV, MGH_LV, MGH_E, MGH_TAV, MGH_LAV, MGH_VAV, MGH_DIP, MGH_X, MGH_NFIELD };
static const fk_corr_spec MGH_CORR[] = {
    {"mgh-trans-momentum-acf.dat", MGH_TV, 1, 0, MGH_TV, 0, 1, 0},
    {"mgh-long-momentum-acf.dat", MGH_LV, 1, 0, MGH_LV, 0, 1, 0},
    {"mgh-rho-acf.dat", MGH_RHO, 0, 0, MGH_RHO, 1, 1, 0},
    {"mgh-energy-acf.dat", MGH_E, 0, 0, MGH_E, 1, 1, 0},
    {"mgh-trans-angmomentum-acf.dat", MGH_TAV, 1, 0, MGH_TAV, 0, 1, 0},
    {"mgh-long-angmomentum-acf.dat", MGH_LAV, 1, 0, MGH_LAV, 0, 1, 0},
    {"mgh-momentum-angmomentum-ccf.dat", MGH_VAV, 1, 0, MGH_VAV, 0, 1, 0},
    {"mgh-dipole-acf.dat", MGH_DIP, 1, 0, MGH_DIP, 0, 1, 0},
    {"mgh-X-cf.dat", MGH_X, 1, 0, MGH_X, 0, 1, 1},
};

static void sample_mgh(sep_fkacc *g, seppart *atoms, sepmol *mols, sepsys *sys)
{
    const unsigned nmol = sys->molptr->num_mols;
    const unsigned t = g->fill;
    sep_mol_cm(atoms, mols, sys);
    sep_mol_velcm(atoms, mols, sys);
    sep_mol_eval_xtrue(atoms, mols, *sys);
    sep_mol_spin(atoms, mols, sys, g->safe ? true : false);
    sep_mol_dipoles(atoms, mols, sys);
    double sumv2 = 0.0;
    for (unsigned m = 0; m < nmol; m++)
        for (int kk = 0; kk < 3; kk++) sumv2 += 0.5 * mols[m].m * sep_Sq(mols[m].v[kk]);
    g->ncalls++;
    g->avekin += sumv2 / ((double)nmol * g->ncalls);
    for (unsigned n = 0; n < g->nwave; n++) {
        double complex acc[MGH_NFIELD][2];
        for (int f = 0; f < MGH_NFIELD; f++) acc[f][0] = acc[f][1] = 0.0;
        for (unsigned m = 0; m < nmol; m++) {
            const double mass = mols[m].m;
            const double complex kf[2] = {cexp(I * g->k[n] * mols[m].x[1]), cexp(-I * g->k[n] * mols[m].x[1])};
            double ekin = 0.0;
            for (int kk = 0; kk < 3; kk++) ekin += 0.5 * mass * sep_Sq(mols[m].v[kk]);
            for (int sgn = 0; sgn < 2; sgn++) {
                acc[MGH_RHO][sgn] += mass * kf[sgn];
                acc[MGH_TV][sgn] += mass * mols[m].v[0] * kf[sgn];
                acc[MGH_LV][sgn] += mass * mols[m].v[1] * kf[sgn];
                acc[MGH_E][sgn] += (ekin - g->avekin) * kf[sgn];
                acc[MGH_TAV][sgn] += mols[m].s[2] * kf[sgn];
                acc[MGH_LAV][sgn] += mols[m].s[1] * kf[sgn];
                acc[MGH_VAV][sgn] += (sgn == 0 ? mols[m].s[2] : mass * mols[m].v[0]) * kf[sgn];
                acc[MGH_DIP][sgn] += mols[m].pel[1] * kf[sgn];
                if (!g->safe) acc[MGH_X][sgn] += mols[m].w[0] * kf[sgn];
            }
        }
        for (int f = 0; f < MGH_NFIELD; f++)
            for (int sgn = 0; sgn < 2; sgn++) fk_field(g, f, sgn, t)[n] = acc[f][sgn];
    }
    fk_push(g, "sep_mgh_sampler", sys->volume);
}

struct sep_sampler_set {
    sep_corr *sacf, *vacf, *msacf, *mvacf;
    sep_fkacc *gh, *mgh;
    sep_profile *profs;
    sep_rdf *radial;
    sep_msdacc *msd;
    unsigned warned;       /* bit per unimplemented sampler name */
};

static struct sep_sampler_set *set_of(sepsampler *s)
{
    if (!s->impl) {
        s->impl = calloc(1, sizeof(struct sep_sampler_set));
        if (!s->impl) sep_error("sep_add_sampler: Couldn't allocate memory");
    }
    return (struct sep_sampler_set *)s->impl;
}

sepsampler sep_init_sampler(void)
{
    sepsampler s;
    s.molptr = NULL;
    s.impl = NULL;
    s.msd_counter = 1;                                        /* source/sepsampler.c:268 */
    return s;
}

void sep_add_mol_sampler(sepsampler *sptr, sepmol *mols) { sptr->molptr = mols; }

static sep_msdacc *msd_new(int lvec, double tsample, int nk, char type, const sepsys *sys)
{
    sep_msdacc *m = calloc(1, sizeof *m);
    if (!m) sep_error("sep_msd_init: Couldn't allocate memory");
    m->nk = nk; m->npart = (int)sys->npart; m->type = type;
    if (lvec > 0) {                                           /* equidistant samples */
        m->lvec = lvec;
        const double dts = tsample / lvec;
        m->isample = (int)(dts / sys->dt);
        m->time = sep_vector((size_t)lvec);
        for (int n = 0; n < lvec; n++) m->time[n] = dts * (n + 1);
    } else {                                                  /* lvec == 0: samples at dt, 2dt, 4dt, ... (:495-509) */
        double t = sys->dt;
        lvec = 1;
        while (t < tsample) { lvec++; t = 2 * t; }
        m->lvec = lvec; m->logmode = 1; m->logcounter = 1; m->isample = 1;
        m->time = sep_vector((size_t)lvec);
        int mult = 1;
        for (int c = 0; c < lvec; c++) { m->time[c] = sys->dt * mult; mult *= 2; }
    }
    m->msd = sep_vector((size_t)m->lvec); m->msdsq = sep_vector((size_t)m->lvec);
    m->fs = calloc((size_t)m->lvec * (nk > 0 ? nk : 1), sizeof(double));
    m->k = sep_vector((size_t)(nk > 0 ? nk : 1));
    m->prev = calloc((size_t)m->npart * 3, sizeof(double));
    m->pos0 = calloc((size_t)m->npart * 3, sizeof(double));
    m->cross = calloc((size_t)m->npart * 3, sizeof(int));
    if (!m->fs || !m->prev || !m->pos0 || !m->cross) sep_error("sep_msd_init: Couldn't allocate memory");
    FILE *fout = fopen("msd-k.dat", "w");
    if (!fout) sep_error("sep_msd_init: Couldn't open file");
    for (int n = 1; n <= nk; n++) {
        m->k[n - 1] = 2 * SEP_PI / sys->length[0] * n;
        fprintf(fout, "%f\n", m->k[n - 1]);
    }
    fclose(fout);
    return m;
}

void sep_add_sampler(sepsampler *sptr, const char *sampler, sepsys sys, int lvec, ...)
{
    struct sep_sampler_set *S = set_of(sptr);
    va_list args;
    va_start(args, lvec);
    if (!strcmp(sampler, "sacf")) {
        if (!S->sacf) S->sacf = corr_new("sep_sacf_init", lvec, va_arg(args, double), sys.dt, 3, 1);
    } else if (!strcmp(sampler, "vacf")) {
        if (!S->vacf) S->vacf = corr_new("sep_vacf_init", lvec, va_arg(args, double), sys.dt, (size_t)sys.npart, 1);
    } else if (!strcmp(sampler, "msacf")) {
        if (!sptr->molptr) sep_error("sep_add_sampler: molpointer not initialized");
        if (!S->msacf) S->msacf = corr_new("sep_msacf_init", lvec, va_arg(args, double), sys.dt, 6, 2);
    } else if (!strcmp(sampler, "mvacf")) {
        if (!S->mvacf) S->mvacf = corr_new("sep_mvacf_init", lvec, va_arg(args, double), sys.dt, (size_t)sys.molptr->num_mols, 1);
    } else if (!strcmp(sampler, "profs")) {
        if (!S->profs) {
            sep_profile *p = calloc(1, sizeof *p);
            if (!p) sep_error("sep_profs_init: Couldn't allocate memory");
            p->type = (char)va_arg(args, int);
            p->isample = (unsigned)va_arg(args, int);
            p->lvec = (unsigned)lvec;
            p->momc = sep_vector((size_t)lvec); p->dens = sep_vector((size_t)lvec);
            p->temp = sep_vector((size_t)lvec); p->svel = sep_vector((size_t)lvec);
            S->profs = p;
        }
    } else if (!strcmp(sampler, "radial")) {
        if (!S->radial) {
            sep_rdf *r = calloc(1, sizeof *r);
            if (!r) sep_error("sep_radial_init: Couldn't allocate memory");
            r->lvec = lvec;
            r->isample = va_arg(args, int);
            const char *t = va_arg(args, char *);
            r->ntypes = (int)strlen(t);
            if (r->ntypes > 255) r->ntypes = 255;
            memcpy(r->types, t, (size_t)r->ntypes);
            for (int n = 1; n <= r->ntypes; n++) r->ncomb += n;
            r->hist = calloc((size_t)lvec * (r->ncomb ? r->ncomb : 1), sizeof(long));
            FILE *fout = fopen("radial_info.dat", "w");
            if (!fout || !r->hist) sep_error("sep_radial_init: Couldn't open file");
            fprintf(fout, "Pairs in radial.dat columns are\n");
            for (int a = 0; a < r->ntypes; a++)
                for (int b = a; b < r->ntypes; b++) fprintf(fout, "%c%c  ", r->types[a], r->types[b]);
            fclose(fout);
            S->radial = r;
        }
    } else if (!strcmp(sampler, "gh")) {
        if (!S->gh) {
            const double tsample = va_arg(args, double);
            const int nwave = va_arg(args, int);
            S->gh = fk_new("sep_gh_init", "gh-wavevector.dat", lvec, tsample, sys.dt, nwave, sys.length[1], GH_NFIELD, GH_CORR,
                           (int)(sizeof GH_CORR / sizeof GH_CORR[0]));
        }
    } else if (!strcmp(sampler, "mgh")) {
        if (!sptr->molptr) sep_error("sep_add_sampler: molpointer not initialized");
        if (!S->mgh) {
            const double tsample = va_arg(args, double);
            const int nwave = va_arg(args, int);
            const int safe = va_arg(args, int);
            S->mgh = fk_new("sep_mgh_init", "mgh-wavevector.dat", lvec, tsample, sys.dt, nwave, sys.length[1], MGH_NFIELD, MGH_CORR,
                            (int)(sizeof MGH_CORR / sizeof MGH_CORR[0]));
            S->mgh->safe = safe ? 1 : 0;
            if (!safe) sep_warning("ACHTUNG - unsafe mode for mgh sampler. Assuming uniaxial single component system");
        }
    } else if (!strcmp(sampler, "msd")) {
        if (!S->msd) {
            const double tsample = va_arg(args, double);
            const int nk = va_arg(args, int);
            const char type = (char)va_arg(args, int);
            S->msd = msd_new(lvec, tsample, nk, type, &sys);
        }
    } else {
        static const char *later[] = {"mprofs", "mcacf", "mavacf", "mmsd", "scatt"};
        int known = -1;
        for (int k = 0; k < 5; k++) if (!strcmp(sampler, later[k])) known = k;
        if (known < 0) sep_error("sep_add_sampler: Sampler %s is not recognized", (char *)sampler);
        if (!(S->warned & (1u << known))) {
            sep_warning("sampler '%s' is not implemented in seplib-b200; it records nothing", (char *)sampler);
            S->warned |= 1u << known;
        }
    }
    va_end(args);
}

/* ---- individual samplers ------------------------------------------------------------------------------ */
static void write_acf(const char *file, const sep_corr *c, double prefactor, size_t per_channel_norm)
{
    FILE *fout = fopen(file, "w");
    if (!fout) sep_error("sep_sample: Couldn't open file %s", (char *)file);
    for (unsigned t = 0; t < c->lvec; t++) {
        const double fac = prefactor / ((double)(c->lvec - t) * per_channel_norm * c->nblocks);
        fprintf(fout, "%f", t * c->dtsample);
        for (int a = 0; a < c->nacf; a++) fprintf(fout, " %f", c->acf[(size_t)a * c->lvec + t] * fac);
        fprintf(fout, "\n");
    }
    fclose(fout);
}

static void sample_sacf(sep_corr *c, sepret *ret, sepsys *sys)
{
    sep_pressure_tensor(ret, sys);
    double *row = corr_row(c);
    row[0] = -ret->P[0][1]; row[1] = -ret->P[0][2]; row[2] = -ret->P[1][2];
    if (corr_push(c)) write_acf("sacf.dat", c, sys->volume, 3);             /* V / (3 (lvec-t) nsample), :338 */
}

static void sample_vacf(sep_corr *c, const seppart *atoms, const sepsys *sys)
{
    double *row = corr_row(c);
    for (long i = 0; i < sys->npart; i++) row[i] = atoms[i].v[0];
    if (corr_push(c)) write_acf("vacf.dat", c, 1.0, (size_t)sys->npart);
}

/* the device keeps the block's rows; a completed block comes back already summed over atoms and time origins */
static void sample_vacf_feed(sep_corr *c, sep_binding *b, const sepsys *sys)
{
    int done = 0;
    sepb_check(sepgpu_feed_vacf(b->gpu, (int)c->lvec, c->rows, &done), "vacf sampler feed");
    if (!done) return;
    for (unsigned t = 0; t < c->lvec; t++) c->acf[t] += c->rows[t];
    c->nblocks++;
    write_acf("vacf.dat", c, 1.0, (size_t)sys->npart);
}

static void sample_msacf(sep_corr *c, seppart *atoms, sepmol *mols, sepret *ret, sepsys *sys)
{
    sep_mol_pressure_tensor(atoms, mols, ret, sys);
    double *row = corr_row(c);
    const int a[3] = {0, 0, 1}, b[3] = {1, 2, 2};
    for (int k = 0; k < 3; k++) {
        row[k] = 0.5 * (ret->P_mol[a[k]][b[k]] + ret->P_mol[b[k]][a[k]]);     /* symmetric part  */
        row[3 + k] = 0.5 * (ret->P_mol[a[k]][b[k]] - ret->P_mol[b[k]][a[k]]); /* antisymmetric   */
    }
    if (corr_push(c)) write_acf("msacf.dat", c, sys->volume, 3);
}

static void sample_mvacf(sep_corr *c, seppart *atoms, sepmol *mols, sepsys *sys)
{
    sep_mol_velcm(atoms, mols, sys);
    double *row = corr_row(c);
    const size_t nmol = sys->molptr->num_mols;
    for (size_t i = 0; i < nmol; i++) row[i] = mols[i].v[0];
    if (corr_push(c)) write_acf("mvacf.dat", c, 1.0, nmol);
}

static void sample_profs(sep_profile *p, const seppart *atoms, const sepsys *sys, sep_binding *feed)
{
    const int dir = 2, dirvel = 0;                                             /* fixed in the reference, :1430-1431 */
    const double dl = sys->length[dir] / p->lvec;
    const double dV = sys->length[0] * sys->length[1] * dl;
    double *j = sep_vector(p->lvec), *rho = sep_vector(p->lvec), *sumv2 = sep_vector(p->lvec);
    int *numb = sep_vector_int(p->lvec);
    if (feed) {                                                                /* the slab sums, formed on the device */
        double *o = sep_vector((size_t)4 * p->lvec);
        sepb_check(sepgpu_feed_profile(feed->gpu, p->type, sys->length[dir], (int)p->lvec, o), "profs sampler feed");
        for (unsigned n = 0; n < p->lvec; n++) {
            j[n] = o[n]; rho[n] = o[p->lvec + n]; sumv2[n] = o[2 * p->lvec + n]; numb[n] = (int)o[3 * p->lvec + n];
        }
        free(o);
    }
    else for (long n = 0; n < sys->npart; n++) {
        if (atoms[n].type != p->type) continue;
        int i = (int)(atoms[n].x[dir] / dl);
        if (i < 0) i = 0;
        if (i >= (int)p->lvec) i = (int)p->lvec - 1;
        j[i] += atoms[n].m * atoms[n].v[dirvel];
        rho[i] += atoms[n].m;
        for (int k = 0; k < 3; k++)
            if (k != dirvel) sumv2[i] += atoms[n].m * atoms[n].v[k] * atoms[n].v[k];
        numb[i]++;
    }
    p->nsample++;
    const double idV = 1.0 / dV;
    for (unsigned n = 0; n < p->lvec; n++) {
        p->momc[n] += j[n] * idV;
        p->dens[n] += rho[n] * idV;
        if (numb[n] > 0) p->temp[n] += sumv2[n];
    }
    free(j); free(rho); free(sumv2); free(numb);
    if (p->nsample % 100 == 0) {
        FILE *fout = fopen("profs.dat", "w");
        if (!fout) sep_error("sep_profs_sampler: Couldn't open file");
        const double insample = 1.0 / p->nsample;
        for (unsigned n = 0; n < p->lvec; n++) {
            double temp_fac = 0;
            if (p->dens[n] > 0.0) {
                p->svel[n] = p->momc[n] / p->dens[n];
                temp_fac = 1.0 / (2.0 * dV * p->dens[n]);
            }
            fprintf(fout, "%f %f %f %f %f\n", (n + 0.5) * dl, p->momc[n] * insample, p->dens[n] * insample,
                    p->temp[n] * temp_fac, p->svel[n]);
        }
        fclose(fout);
    }
}

static void sample_radial(sep_rdf *r, const seppart *atoms, const sepsys *sys, sep_binding *feed)
{
    const long npart = sys->npart;
    const double lbox = sys->length[0];
    const double dg = 0.5 * lbox / r->lvec;
    if (feed) {                                                                /* this configuration's counts from the device */
        const size_t nb = (size_t)r->lvec * r->ncomb;
        long long *cnt = malloc(nb * sizeof *cnt);
        if (!cnt) sep_error("sep_radial_sample: Couldn't allocate memory");
        sepb_check(sepgpu_feed_radial(feed->gpu, lbox, r->lvec, r->ntypes, r->types, cnt), "radial sampler feed");
        for (size_t q = 0; q < nb; q++) r->hist[q] += (long)cnt[q];
        free(cnt);
    }
    else for (long i = 0; i < npart - 1; i++)
        for (long jx = i + 1; jx < npart; jx++) {
            double r2 = 0.0;
            for (int k = 0; k < 3; k++) {
                double d = atoms[i].x[k] - atoms[jx].x[k];
                sep_Wrap(d, lbox);
                r2 += d * d;
            }
            const int index = (int)(sqrt(r2) / dg);
            if (index >= r->lvec) continue;
            int counter = 0;
            for (int a = 0; a < r->ntypes; a++)
                for (int b = a; b < r->ntypes; b++) {
                    if ((atoms[i].type == r->types[a] && atoms[jx].type == r->types[b]) ||
                        (atoms[i].type == r->types[b] && atoms[jx].type == r->types[a]))
                        r->hist[(size_t)index * r->ncomb + counter] += 1;
                    counter++;
                }
        }
    r->nsample++;
    FILE *fout = fopen("radial.dat", "w");
    if (!fout) sep_error("sep_radial_sample: Couldn't open file");
    for (int i = 0; i < r->lvec; i++) {
        const double vi = pow(i * dg, 3.0), vii = pow((i + 1) * dg, 3.0);
        fprintf(fout, "%f ", (i + 0.5) * dg);
        for (int n = 0; n < r->ncomb; n++)
            fprintf(fout, "%f ", (double)r->hist[(size_t)i * r->ncomb + n] / ((vii - vi) * r->nsample));
        fprintf(fout, "\n");
    }
    fclose(fout);
}

/* follows the atoms across the periodic boundaries on its own (prev position per step), :537-552 */
static void msd_track(sep_msdacc *m, const seppart *atoms, const sepsys *sys)
{
    for (long n = 0; n < sys->npart; n++)
        for (int k = 0; k < 3; k++) {
            const double d = m->prev[3 * n + k] - atoms[n].x[k];
            if (d > 0.5 * sys->length[k]) m->cross[3 * n + k]++;
            else if (d < -0.5 * sys->length[k]) m->cross[3 * n + k]--;
            m->prev[3 * n + k] = atoms[n].x[k];
        }
}

static void msd_take(sep_msdacc *m, const seppart *atoms, const sepsys *sys, sep_binding *feed)
{
    int index = m->fill;
    double sd = 0.0, qd = 0.0;
    int ntype_feed = 0;
    if (feed) {                 /* sums from the device; its crossing counters stand in for msd_track's (origin reset at index 0) */
        double sums[3], *fsv = sep_vector((size_t)(m->nk > 0 ? m->nk : 1));
        sepb_check(sepgpu_feed_msd(feed->gpu, index == 0, m->type, sys->length, m->nk, m->k, sums, fsv), "msd sampler feed");
        sd = sums[0]; qd = sums[1]; ntype_feed = (int)sums[2];
        for (int i = 0; i < m->nk; i++) m->fs[(size_t)index * m->nk + i] += fsv[i];
        free(fsv);
    }
    else {
    if (index == 0)
        for (long n = 0; n < sys->npart; n++)
            for (int k = 0; k < 3; k++) {
                m->prev[3 * n + k] = m->pos0[3 * n + k] = atoms[n].x[k];
                m->cross[3 * n + k] = 0;
            }
    for (long n = 0; n < sys->npart; n++) {
        if (atoms[n].type != m->type) continue;
        double a = 0.0, dx0 = 0.0;
        for (int k = 0; k < 3; k++) {
            const double dr = atoms[n].x[k] + m->cross[3 * n + k] * sys->length[k] - m->pos0[3 * n + k];
            if (k == 0) dx0 = dr;
            a += dr * dr;
        }
        sd += a;
        qd += a * a;
        for (int i = 0; i < m->nk; i++) m->fs[(size_t)index * m->nk + i] += cos(m->k[i] * dx0);    /* Re exp(i k dx) */
    }
    }
    m->msd[index] += sd;
    m->msdsq[index] += qd;
    index++;
    if (index == m->lvec) {
        m->nsample++;
        const int ntype = feed ? ntype_feed : sep_count_type((seppart *)atoms, m->type, m->npart);
        const double norm = (double)ntype * m->nsample;
        FILE *fout = fopen("msd.dat", "w");
        if (!fout) sep_error("sep_msd_sample: Couldn't open file");
        for (int n = 0; n < m->lvec; n++) fprintf(fout, "%f %f \n", m->time[n], m->msd[n] / norm);
        fclose(fout);
        fout = fopen("msd-gaussparam.dat", "w");
        if (!fout) sep_error("sep_msd_sample: Couldn't open file");
        for (int n = 0; n < m->lvec; n++) {
            const double a = m->msdsq[n] / norm, b = sep_Sq(m->msd[n] / norm);
            fprintf(fout, "%f %f \n", m->time[n], 3.0 * a / (5.0 * b) - 1.);
        }
        fclose(fout);
        fout = fopen("msd-incoherent.dat", "w");
        if (!fout) sep_error("sep_msd_sample: Couldn't open file");
        for (int n = 0; n < m->lvec; n++) {
            fprintf(fout, "%f ", m->time[n]);
            for (int i = 0; i < m->nk; i++) fprintf(fout, "%f ", m->fs[(size_t)n * m->nk + i] / norm);
            fprintf(fout, "\n");
        }
        fclose(fout);
        index = 0;
    }
    m->fill = index;
}

void sep_sample(seppart *pptr, sepsampler *sptr, sepret *ret, sepsys sys, unsigned n)
{
    struct sep_sampler_set *S = (struct sep_sampler_set *)sptr->impl;
    if (!S) return;
    /* sampler feeds: the device forms the sums, atoms[] stays where it is (pending host writes are uploaded first) */
    sep_binding *feed = NULL;
    if (feeds_enabled() && sepdd_world() <= 1 && (S->vacf || S->profs || S->gh || S->radial || S->msd)) {
        feed = sepb_prepare(pptr, &sys);
        if (!feed->gpu || feed->dd) feed = NULL;
    }
    if (!feed && ((S->vacf && n % S->vacf->isample == 0) || (S->profs && n % S->profs->isample == 0) || (S->gh && n % S->gh->isample == 0) ||
        (S->radial && n % (unsigned)S->radial->isample == 0) || S->msd))
        sep_gpu_sync(pptr);                                    /* the samplers below read atoms[] on the host */
    if (S->sacf && n % S->sacf->isample == 0) sample_sacf(S->sacf, ret, &sys);
    if (S->vacf && n % S->vacf->isample == 0) { if (feed) sample_vacf_feed(S->vacf, feed, &sys); else sample_vacf(S->vacf, pptr, &sys); }
    if (S->msacf && n % S->msacf->isample == 0) sample_msacf(S->msacf, pptr, sptr->molptr, ret, &sys);
    if (S->gh && n % S->gh->isample == 0) { if (feed) sample_gh_feed(S->gh, feed, &sys); else sample_gh(S->gh, pptr, &sys); }
    if (S->mgh && n % S->mgh->isample == 0) sample_mgh(S->mgh, pptr, sptr->molptr, &sys);
    if (S->profs && n % S->profs->isample == 0) sample_profs(S->profs, pptr, &sys, feed);
    if (S->mvacf && n % S->mvacf->isample == 0) sample_mvacf(S->mvacf, pptr, sptr->molptr, &sys);
    if (S->radial && n % (unsigned)S->radial->isample == 0) sample_radial(S->radial, pptr, &sys, feed);
    if (S->msd) {                                              /* source/sepsampler.c:213-231 */
        sep_msdacc *m = S->msd;
        if (!feed) msd_track(m, pptr, &sys);
        if (!m->logmode && n % (unsigned)m->isample == 0) msd_take(m, pptr, &sys, feed);
        else if (m->logmode && sptr->msd_counter % (unsigned long)m->logcounter == 0) {
            msd_take(m, pptr, &sys, feed);
            m->logcounter = m->fill == 0 ? 1 : 2 * m->logcounter;
        }
        sptr->msd_counter++;
    }
}

void sep_close_sampler(sepsampler *ptr)
{
    struct sep_sampler_set *S = (struct sep_sampler_set *)ptr->impl;
    if (!S) return;
    corr_free(S->sacf); corr_free(S->vacf); corr_free(S->msacf); corr_free(S->mvacf);
    if (S->profs) { free(S->profs->momc); free(S->profs->dens); free(S->profs->temp); free(S->profs->svel); free(S->profs); }
    if (S->radial) { free(S->radial->hist); free(S->radial); }
    fk_free(S->gh); fk_free(S->mgh);
    if (S->msd) {
        sep_msdacc *m = S->msd;
        free(m->time); free(m->msd); free(m->msdsq); free(m->k); free(m->fs); free(m->prev); free(m->pos0); free(m->cross);
        free(m);
    }
    free(S);
    ptr->impl = NULL;
}
