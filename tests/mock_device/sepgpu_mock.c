/* sepgpu_mock.c -- TEST INFRASTRUCTURE ONLY.  A stand-in for the device layer (include/sepgpu.h) that keeps the state
 * in host memory and does the arithmetic with the CPU oracle (oracle/sep_oracle.c).
 *
 * Purpose: exercise the HOST layer of seplib-b200 (seplib_b200/csrc/host/sep_*.c: dispatch, brute/list control flow,
 * rebuild flag, epot assign/accumulate rules, dirty-bit coherence in the step / lazy / full modes, the noise stream of
 * the stochastic integrators, ...) in the CPU test run, where no GPU exists.  tests/test_cpu_hostlayer.py links the
 * unchanged host sources with this file into tests/_build/libsep_hostmock.so and drives the same sep_* loops the GPU
 * tests drive.
 *
 * It is NOT a fallback: nothing outside tests/ builds, links or loads it, and seplib_b200/libsep.so contains no CPU
 * path (tests/test_cpu_host.py::test_no_cpu_fallback_without_device).  Only the entry points the host layer calls
 * are provided.  Everything is eager: there is no deferred thermostat term and no lazy reset here -- the host layer
 * must not depend on those device-internal optimisations, which is part of what this checks.
 */
#include "sepgpu.h"
#include "sep_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct sepgpu_ctx {
    int n;
    double *x, *v, *f, *a, *m, *z, *xn, *pv, *pa, *x0, *prevf, *randn;
    char *type;
    int *mol, *cn, *cr, *bond, *angle, *dihed;
    unsigned *blist, *alist, *dlist, nb, na, nd;
    double *blengths, *angles, *dihedrals;
    int *pairs; long npairs; int list_valid;
    orc_ret ret;
    double max_dist2, alpha[4];
    int neighb_flag, nbuild, have_excl;
};

static char g_err[512] = "";
static void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *sepgpu_last_error(void) { return g_err; }
int sepgpu_device_count(void) { return 1; }

static void *zalloc(size_t count, size_t size) { return calloc(count ? count : 1, size); }

int sepgpu_create(sepgpu_ctx **out, size_t npart, int device)
{
    (void)device;
    sepgpu_ctx *c = calloc(1, sizeof *c);
    const size_t n = npart;
    c->n = (int)n;
    c->x = zalloc(3 * n, 8); c->v = zalloc(3 * n, 8); c->f = zalloc(3 * n, 8); c->a = zalloc(3 * n, 8);
    c->m = zalloc(n, 8); c->z = zalloc(n, 8); c->xn = zalloc(3 * n, 8); c->pv = zalloc(3 * n, 8); c->pa = zalloc(3 * n, 8);
    c->x0 = zalloc(3 * n, 8); c->prevf = zalloc(3 * n, 8); c->randn = zalloc(3 * n, 8);
    c->type = zalloc(n, 1); c->mol = zalloc(n, 4); c->cn = zalloc(3 * n, 4); c->cr = zalloc(3 * n, 4);
    c->bond = zalloc(10 * n, 4); c->angle = zalloc(10 * n, 4); c->dihed = zalloc(20 * n, 4);
    for (size_t i = 0; i < n; i++) { c->m[i] = 1.0; c->type[i] = 'A'; c->mol[i] = -1; }
    for (size_t i = 0; i < 10 * n; i++) c->bond[i] = c->angle[i] = -1;
    for (size_t i = 0; i < 20 * n; i++) c->dihed[i] = -1;
    *out = c;
    return 0;
}

void sepgpu_destroy(sepgpu_ctx *c)
{
    if (!c) return;
    void *p[] = {c->x, c->v, c->f, c->a, c->m, c->z, c->xn, c->pv, c->pa, c->x0, c->prevf, c->randn, c->type, c->mol, c->cn,
                 c->cr, c->bond, c->angle, c->dihed, c->blist, c->alist, c->dlist, c->blengths, c->angles, c->dihedrals, c->pairs};
    for (size_t k = 0; k < sizeof p / sizeof p[0]; k++) free(p[k]);
    free(c);
}

/* ---- fields ------------------------------------------------------------------------------------------- */
typedef struct { void *base; size_t elem; int width; } field_t;

static field_t field_of(sepgpu_ctx *c, int field)
{
    switch (field) {
    case SEPGPU_F_X: return (field_t){c->x, 8, 3};
    case SEPGPU_F_V: return (field_t){c->v, 8, 3};
    case SEPGPU_F_F: return (field_t){c->f, 8, 3};
    case SEPGPU_F_A: return (field_t){c->a, 8, 3};
    case SEPGPU_F_XN: return (field_t){c->xn, 8, 3};
    case SEPGPU_F_PV: return (field_t){c->pv, 8, 3};
    case SEPGPU_F_PA: return (field_t){c->pa, 8, 3};
    case SEPGPU_F_X0: return (field_t){c->x0, 8, 3};
    case SEPGPU_F_M: return (field_t){c->m, 8, 1};
    case SEPGPU_F_Z: return (field_t){c->z, 8, 1};
    case SEPGPU_F_TYPE: return (field_t){c->type, 1, 1};
    case SEPGPU_F_MOLINDEX: return (field_t){c->mol, 4, 1};
    case SEPGPU_F_CROSS_NEIGHB: return (field_t){c->cn, 4, 3};
    case SEPGPU_F_CROSSINGS: return (field_t){c->cr, 4, 3};
    case SEPGPU_F_BOND: return (field_t){c->bond, 4, 10};
    case SEPGPU_F_ANGLE: return (field_t){c->angle, 4, 10};
    case SEPGPU_F_DIHED: return (field_t){c->dihed, 4, 20};
    default: return (field_t){NULL, 0, 0};
    }
}

static void refresh_accel(sepgpu_ctx *c)
{
    for (int i = 0; i < c->n; i++)
        for (int k = 0; k < 3; k++) c->a[3 * i + k] = c->f[3 * i + k] / c->m[i];
}

static int move_field(sepgpu_ctx *c, int field, void *host, size_t stride, int to_device)
{
    field_t fd = field_of(c, field);
    if (!fd.base) { set_error("mock: bad field %d", field); return SEPGPU_EINVAL; }
    const size_t row = fd.elem * fd.width;
    if (!stride) stride = row;
    if (!to_device && field == SEPGPU_F_A) refresh_accel(c);
    for (int i = 0; i < c->n; i++) {
        char *h = (char *)host + (size_t)i * stride, *d = (char *)fd.base + (size_t)i * row;
        if (to_device) memcpy(d, h, row); else memcpy(h, d, row);
    }
    if (to_device && (field == SEPGPU_F_X || field == SEPGPU_F_TYPE || field == SEPGPU_F_MOLINDEX || field >= SEPGPU_F_BOND))
        c->list_valid = 0;
    if (to_device && field >= SEPGPU_F_BOND && field <= SEPGPU_F_DIHED) c->have_excl = 1;
    return 0;
}

int sepgpu_put(sepgpu_ctx *c, int field, const void *host, size_t stride) { return move_field(c, field, (void *)host, stride, 1); }
int sepgpu_get(sepgpu_ctx *c, int field, void *host, size_t stride) { return move_field(c, field, host, stride, 0); }

int sepgpu_put_fields(sepgpu_ctx *c, const void *base, size_t stride, int nfields, const int *fields, const size_t *offsets)
{
    for (int k = 0; k < nfields; k++) {
        int rc = move_field(c, fields[k], (char *)base + offsets[k], stride, 1);
        if (rc) return rc;
    }
    return 0;
}

int sepgpu_get_fields(sepgpu_ctx *c, void *base, size_t stride, int nfields, const int *fields, const size_t *offsets)
{
    for (int k = 0; k < nfields; k++) {
        int rc = move_field(c, fields[k], (char *)base + offsets[k], stride, 0);
        if (rc) return rc;
    }
    return 0;
}

static unsigned *dup_u(const unsigned *src, size_t count)
{
    unsigned *p = zalloc(count, sizeof(unsigned));
    if (src && count) memcpy(p, src, count * sizeof(unsigned));
    return p;
}

int sepgpu_set_topology(sepgpu_ctx *c, const unsigned *blist, unsigned nb, const unsigned *alist, unsigned na,
                        const unsigned *dlist, unsigned nd)
{
    free(c->blist); free(c->alist); free(c->dlist); free(c->blengths); free(c->angles); free(c->dihedrals);
    c->blist = dup_u(blist, 3 * (size_t)nb); c->alist = dup_u(alist, 4 * (size_t)na); c->dlist = dup_u(dlist, 5 * (size_t)nd);
    c->nb = nb; c->na = na; c->nd = nd;
    c->blengths = zalloc(nb, 8); c->angles = zalloc(na, 8); c->dihedrals = zalloc(nd, 8);
    return 0;
}

int sepgpu_get_bonded_values(sepgpu_ctx *c, double *bl, double *an, double *di)
{
    if (bl && c->nb) memcpy(bl, c->blengths, 8 * (size_t)c->nb);
    if (an && c->na) memcpy(an, c->angles, 8 * (size_t)c->na);
    if (di && c->nd) memcpy(di, c->dihedrals, 8 * (size_t)c->nd);
    return 0;
}

/* ---- per-step path ----------------------------------------------------------------------------------------- */
int sepgpu_reset_ret(sepgpu_ctx *c) { memset(&c->ret, 0, sizeof c->ret); return 0; }

int sepgpu_reset_force(sepgpu_ctx *c)
{
    memset(c->f, 0, 24 * (size_t)c->n);
    c->max_dist2 = 0.0;
    return 0;
}

static orc_topo topo_of(sepgpu_ctx *c)
{
    orc_topo t = {c->mol, c->bond, c->angle, c->dihed};
    return t;
}

int sepgpu_neighb_build(sepgpu_ctx *c, const sepgpu_sys *sys, unsigned opt)
{
    const double vol = sys->length[0] * sys->length[1] * sys->length[2];
    const double cut = sys->cf + sys->skin;
    long cap = (long)(1.5 * c->n * (2.1 * cut * cut * cut * c->n / vol) + 65536);
    free(c->pairs);
    c->pairs = malloc(sizeof(int) * 2 * (size_t)cap);
    orc_topo t = topo_of(c);
    long k;
    if (sys->neighb_update == 1)
        k = orc_neighb_pairs_n2(c->n, c->x, sys->length, cut, opt, &t, c->pairs, cap);
    else
        k = orc_neighb_pairs(c->n, c->x, sys->length, sys->nsubbox, sys->lsubbox, cut, opt, &t, c->pairs, cap);
    if (k == -2) { set_error("mock: too many neighbours"); return SEPGPU_ENEIGHB; }
    if (k < 0) { set_error("mock: pair capacity exceeded"); return SEPGPU_EINVAL; }
    c->npairs = k;
    c->list_valid = 1;
    c->nbuild++;
    return 0;
}

static void add_ret(sepgpu_ctx *c, const orc_ret *t, int epot_assign, int bond_virial)
{
    if (epot_assign) c->ret.epot = t->epot; else c->ret.epot += t->epot;
    c->ret.ecoul += t->ecoul;
    for (int k = 0; k < 9; k++) {
        c->ret.pot_P[k] += t->pot_P[k];
        if (bond_virial) c->ret.pot_P_bond[k] += t->pot_P_bond[k];
    }
}

/* sep_force_pairs with a caller-made pair function: the same four-point cubic over the r^2 table the device kernels use
 * (sepgpu_pair.cuh table_eval), written as a plain loop over the pair list */
static int table_pair(const double *tab, int n, double lo, double inv, double r2, double *ft, double *u)
{
    double s = (r2 - lo) * inv;
    if (s < 0.0) return -1;
    int k = (int)s;
    if (k < 1) k = 1;
    if (k > n - 3) k = n - 3;
    const double t = s - (double)k;
    const double w0 = -t * (t - 1.0) * (t - 2.0) / 6.0, w1 = (t + 1.0) * (t - 1.0) * (t - 2.0) / 2.0;
    const double w2 = -(t + 1.0) * t * (t - 2.0) / 2.0, w3 = (t + 1.0) * t * (t - 1.0) / 6.0;
    *ft = w0 * tab[2 * (k - 1)] + w1 * tab[2 * k] + w2 * tab[2 * (k + 1)] + w3 * tab[2 * (k + 2)];
    *u = w0 * tab[2 * (k - 1) + 1] + w1 * tab[2 * k + 1] + w2 * tab[2 * (k + 1) + 1] + w3 * tab[2 * (k + 2) + 1];
    return 0;
}

int sepgpu_force_table(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], double cf, const double *tab_fu, int n, double r2_lo,
                       unsigned opt, int epot_assign)
{
    if (!c || !tab_fu || n < 8 || !(cf * cf > r2_lo)) return SEPGPU_EINVAL;
    if (sys->neighb_update == 0) { set_error("mock device: tabulated pair functions need a neighbour list"); return SEPGPU_EINVAL; }
    if (!c->list_valid) { int rc = sepgpu_neighb_build(c, sys, opt); if (rc) return rc; }
    const double inv = (double)(n - 1) / (cf * cf - r2_lo);
    orc_ret t;
    memset(&t, 0, sizeof t);
    for (long p = 0; p < c->npairs; p++) {
        const int i = c->pairs[2 * p], j = c->pairs[2 * p + 1];
        const char a = c->type[i], b = c->type[j];
        if (!((a == types[0] && b == types[1]) || (a == types[1] && b == types[0]))) continue;
        double r[3], r2 = 0.0, ft, u;
        for (int k = 0; k < 3; k++) {
            r[k] = c->x[3 * i + k] - c->x[3 * j + k];
            if (r[k] > 0.5 * sys->length[k]) r[k] -= sys->length[k];
            else if (r[k] < -0.5 * sys->length[k]) r[k] += sys->length[k];
            r2 += r[k] * r[k];
        }
        if (!(r2 < cf * cf)) continue;
        if (table_pair(tab_fu, n, r2_lo, inv, r2, &ft, &u)) { set_error("pair below the table"); return SEPGPU_ETABLE; }
        for (int k = 0; k < 3; k++) {
            c->f[3 * i + k] += ft * r[k];
            c->f[3 * j + k] -= ft * r[k];
            for (int kk = 0; kk < 3; kk++) t.pot_P[3 * k + kk] += ft * r[k] * r[kk];
        }
        t.epot += u;
    }
    add_ret(c, &t, epot_assign, 0);
    return 0;
}

int sepgpu_force_lj(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], const sepgpu_ljparam *p, unsigned opt, int epot_assign)
{
    int pot;
    double param[4] = {p->cf, p->eps, p->sigma, p->aw};
    if (p->eps == 1.0 && p->sigma == 1.0 && p->aw == 1.0 && p->shift == 0.0) pot = ORC_POT_LJ;
    else if (p->eps == 1.0 && p->sigma == 1.0 && p->aw == 1.0 && p->shift == -0.016316891136) pot = ORC_POT_LJ_SHIFT;
    else if (p->eps == 1.0 && p->sigma == 1.0 && p->aw == 1.0 && p->shift == -1.0) pot = ORC_POT_WCA;
    else pot = ORC_POT_LJ_PARAM;
    orc_ret t;
    memset(&t, 0, sizeof t);
    orc_topo tp = topo_of(c);
    if (sys->neighb_update == 0) {
        orc_force_pairs_brute(c->n, c->x, c->type, sys->length, types, p->cf, pot, pot == ORC_POT_LJ_PARAM ? param : NULL, opt, &tp, c->f, &t);
    } else {
        if (!c->list_valid) { int rc = sepgpu_neighb_build(c, sys, opt); if (rc) return rc; }
        orc_force_pairs_list(c->n, c->x, c->type, sys->length, c->pairs, c->npairs, types, p->cf, pot,
                             pot == ORC_POT_LJ_PARAM ? param : NULL, c->f, &t);
    }
    add_ret(c, &t, epot_assign, 0);
    return 0;
}

int sepgpu_coulomb_sf(sepgpu_ctx *c, const sepgpu_sys *sys, double cf, unsigned opt)
{
    orc_ret t;
    memset(&t, 0, sizeof t);
    orc_topo tp = topo_of(c);
    if (sys->neighb_update == 0) orc_coulomb_sf_brute(c->n, c->x, c->z, sys->length, cf, opt, &tp, c->f, &t);
    else {
        if (!c->list_valid) { set_error("mock: coulomb_sf without a list"); return SEPGPU_ESTATE; }
        orc_coulomb_sf_list(c->n, c->x, c->z, sys->length, c->pairs, c->npairs, cf, c->f, &t);
    }
    c->ret.epot += t.epot;                      /* the oracle adds ecoul to both (source/sepcoulomb.c:150-153) */
    c->ret.ecoul += t.ecoul;
    for (int k = 0; k < 9; k++) c->ret.pot_P[k] += t.pot_P[k];
    return 0;
}

int sepgpu_force_dpd(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], double cf, double aij, double temp,
                     double sigma, unsigned opt, unsigned long long seed, unsigned long long step)
{
    orc_ret t;
    memset(&t, 0, sizeof t);
    if (sys->neighb_update == 0) {
        orc_topo tp = topo_of(c);
        long cap = (long)c->n * (c->n - 1) / 2 + 16;
        int *all = malloc(sizeof(int) * 2 * (size_t)cap);
        long k = orc_neighb_pairs_n2(c->n, c->x, sys->length, cf, opt, &tp, all, cap);
        if (k < 0) { free(all); set_error("mock: dpd brute pair list failed"); return SEPGPU_EINVAL; }
        orc_dpd_force_list(c->n, c->x, c->pv, c->type, sys->length, all, k, types, cf, aij, temp, sigma, sys->dt, seed, step, c->f, &t);
        free(all);
    } else {
        if (!c->list_valid) { int rc = sepgpu_neighb_build(c, sys, opt); if (rc) return rc; }
        orc_dpd_force_list(c->n, c->x, c->pv, c->type, sys->length, c->pairs, c->npairs, types, cf, aij, temp, sigma, sys->dt,
                           seed, step, c->f, &t);
    }
    add_ret(c, &t, 1, 0);
    return 0;
}

int sepgpu_stretch_harmonic(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double lbond, double ks)
{
    orc_stretch_harmonic(c->x, sys->length, c->blist, c->nb, type, lbond, ks, c->f, &c->ret, c->blengths);
    return 0;
}
int sepgpu_angle_harmonic(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double angle0, double k)
{
    orc_angle_harmonic(c->x, sys->length, c->alist, c->na, type, angle0, k, c->f, &c->ret, c->angles);
    return 0;
}
int sepgpu_angle_cossq(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double angle0, double k)
{
    orc_angle_cossq(c->x, sys->length, c->alist, c->na, type, angle0, k, c->f, &c->ret, c->angles);
    return 0;
}
int sepgpu_torsion_ryckaert(sepgpu_ctx *c, const sepgpu_sys *sys, int type, const double g[6])
{
    orc_torsion_ryckaert(c->x, sys->length, c->dlist, c->nd, type, g, c->f, &c->ret, c->dihedrals);
    return 0;
}

int sepgpu_bonded_side(sepgpu_ctx *c, const sepgpu_sys *sys, int kind, int type, const double *par, double *out3)
{
    orc_ret t;
    memset(&t, 0, sizeof t);
    memset(out3, 0, sizeof(double) * 3 * (size_t)c->n);
    if (kind == 0) orc_stretch_harmonic(c->x, sys->length, c->blist, c->nb, type, par[0], par[1], out3, &t, c->blengths);
    else if (kind == 1) orc_angle_cossq(c->x, sys->length, c->alist, c->na, type, par[0], par[1], out3, &t, c->angles);
    else if (kind == 2) orc_torsion_ryckaert(c->x, sys->length, c->dlist, c->nd, type, par, out3, &t, c->dihedrals);
    else return SEPGPU_EINVAL;
    return 0;
}

int sepgpu_nosehoover(sepgpu_ctx *c, const sepgpu_sys *sys, double temp0, int slot, double tau)
{
    c->alpha[slot] = orc_nosehoover(c->n, c->v, c->m, c->f, temp0, c->alpha[slot], tau, sys->dt);
    return 0;
}
int sepgpu_nosehoover_type(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double Td, double alpha3[3], double Q)
{
    orc_nosehoover_type(c->n, c->v, c->m, c->type, type, c->f, Td, alpha3, Q, sys->dt);
    return 0;
}
int sepgpu_set_alpha(sepgpu_ctx *c, int slot, double alpha) { c->alpha[slot] = alpha; return 0; }

static void after_integrator(sepgpu_ctx *c, int flag)
{
    c->neighb_flag = flag;
    if (flag) c->list_valid = 0;
}

int sepgpu_leapfrog(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    after_integrator(c, orc_leapfrog(c->n, c->x, c->v, c->f, c->m, c->a, c->xn, c->cn, c->cr, sys->length, sys->dt, sys->skin,
                                     &c->max_dist2, &c->ret));
    return 0;
}
int sepgpu_verlet_dpd(sepgpu_ctx *c, const sepgpu_sys *sys, double lambda, int stepnow)
{
    after_integrator(c, orc_verlet_dpd(c->n, c->x, c->v, c->f, c->m, c->a, c->pv, c->pa, c->xn, c->cn, c->cr, sys->length, sys->dt,
                                       lambda, stepnow, sys->skin, &c->max_dist2, &c->ret));
    return 0;
}

/* the oracle's stochastic integrators draw their own noise; the device ABI receives it from the host layer, so the
 * two routines are restated here around the given numbers (same statements as oracle/sep_oracle.c: orc_fp,
 * orc_langevin_gjf) */
static int wrap_and_trigger(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    if (sqrt(c->max_dist2) > sys->skin * 0.5) {
        for (int q = 0; q < 3 * c->n; q++) { c->xn[q] = c->x[q]; c->cn[q] = 0; }
        return 1;
    }
    return 0;
}

int sepgpu_fp(sepgpu_ctx *c, const sepgpu_sys *sys, double temp, const double *noise4)
{
    const double dt = sys->dt, fac = sqrt(1.0 / 12.0);
    double sumekin = 0.0;
    for (int i = 0; i < c->n; i++) {
        double d2 = 0.0;
        const double im = 1.0 / c->m[i], fric = temp / noise4[4 * i + 3], gaussfac = sqrt(24 * temp * fric / dt);
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            const double a = noise4[4 * i + k] * fac * gaussfac;
            c->x[q] += dt * c->v[q];
            c->v[q] += im * dt * (c->f[q] - fric * c->v[q] + a);
            if (c->x[q] > sys->length[k]) { c->x[q] -= sys->length[k]; c->cn[q]++; c->cr[q]++; }
            else if (c->x[q] < 0.0) { c->x[q] += sys->length[k]; c->cn[q]--; c->cr[q]--; }
            sumekin += c->v[q] * c->v[q] * c->m[i];
            const double ri = (c->x[q] + c->cn[q] * sys->length[k]) - c->xn[q];
            d2 += ri * ri;
        }
        if (d2 > c->max_dist2) c->max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) c->ret.kin_P[3 * k + kk] += c->v[3 * i + k] * c->v[3 * i + kk] * c->m[i];
    }
    c->ret.ekin += 0.5 * sumekin;
    after_integrator(c, wrap_and_trigger(c, sys));
    return 0;
}

int sepgpu_langevin_gjf(sepgpu_ctx *c, const sepgpu_sys *sys, double temp, double alpha, const double *noise4)
{
    const double dt = sys->dt, cc = exp(-alpha * dt);
    double sumekin = 0.0;
    for (int i = 0; i < c->n; i++) {
        const double mass = c->m[i], imass = 1.0 / mass, imass2 = 0.5 * imass;
        const double fac = sqrt(temp * (1.0 - cc * cc));
        const double cq = alpha * dt * imass2, ca = (1.0 - cq) / (1.0 + cq), cb = 1.0 / (1.0 + cq);
        double d2 = 0.0;
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            c->v[q] = ca * c->v[q] + dt * imass2 * (ca * c->prevf[q] + c->f[q]) + cb * imass * c->randn[q];
            c->prevf[q] = c->f[q];
            sumekin += c->v[q] * c->v[q] * mass;
            c->randn[q] = fac * noise4[4 * i + k];
            c->x[q] += cb * dt * c->v[q] + cb * dt * dt * imass2 * c->f[q] + cb * dt * imass2 * c->randn[q];
        }
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            if (c->x[q] > sys->length[k]) { c->x[q] -= sys->length[k]; c->cn[q]++; c->cr[q]++; }
            else if (c->x[q] < 0.0) { c->x[q] += sys->length[k]; c->cn[q]--; c->cr[q]--; }
            const double ri = (c->x[q] + c->cn[q] * sys->length[k]) - c->xn[q];
            d2 += ri * ri;
        }
        if (d2 > c->max_dist2) c->max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) c->ret.kin_P[3 * k + kk] += c->v[3 * i + k] * c->v[3 * i + kk] * mass;
    }
    c->ret.ekin += 0.5 * sumekin;
    after_integrator(c, wrap_and_trigger(c, sys));
    return 0;
}

int sepgpu_reset_momentum(sepgpu_ctx *c, char type)
{
    double mom[3] = {0.0, 0.0, 0.0}, mass = 0.0;
    for (int i = 0; i < c->n; i++)
        if (c->type[i] == type) {
            for (int k = 0; k < 3; k++) mom[k] += c->v[3 * i + k] * c->m[i];
            mass += c->m[i];
        }
    for (int i = 0; i < c->n; i++)
        if (c->type[i] == type)
            for (int k = 0; k < 3; k++) c->v[3 * i + k] -= mom[k] / mass;
    return 0;
}

int sepgpu_scale_box(sepgpu_ctx *c, const double scale[3], const double new_length[3])
{
    (void)new_length;
    for (int i = 0; i < c->n; i++)
        for (int k = 0; k < 3; k++) c->x[3 * i + k] *= scale[k];
    return 0;
}

int sepgpu_relax_temp(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double Td, double tau, double *ekin_type)
{
    const double e = orc_relax_temp(c->n, c->v, c->m, c->type, type, Td, tau, sys->dt);
    if (ekin_type) *ekin_type = e;
    return 0;
}

int sepgpu_force_x0(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double kspring)
{
    if (kspring != 500.0) { set_error("mock: only the reference's spring constant"); return SEPGPU_EINVAL; }
    orc_force_x0(c->n, c->x, c->x0, c->type, type, sys->length, c->f);
    return 0;
}

/* the molecule-pair force table needs the pair loops themselves: not modelled */
int sepgpu_fij_enable(sepgpu_ctx *c, int nmol) { (void)c; (void)nmol; return 0; }
int sepgpu_fij_reset(sepgpu_ctx *c) { (void)c; return 0; }
int sepgpu_fij_get(sepgpu_ctx *c, float *out) { (void)c; (void)out; set_error("mock: no Fij table"); return SEPGPU_ESTATE; }

int sepgpu_read_scalars(sepgpu_ctx *c, sepgpu_scalars *out)
{
    memset(out, 0, sizeof *out);
    out->epot = c->ret.epot; out->ecoul = c->ret.ecoul; out->ekin = c->ret.ekin;
    memcpy(out->pot_P, c->ret.pot_P, sizeof out->pot_P);
    memcpy(out->kin_P, c->ret.kin_P, sizeof out->kin_P);
    memcpy(out->pot_P_bond, c->ret.pot_P_bond, sizeof out->pot_P_bond);
    out->max_dist2 = c->max_dist2;
    for (int i = 0; i < c->n; i++)
        for (int k = 0; k < 3; k++) out->sum_mv2 += c->m[i] * c->v[3 * i + k] * c->v[3 * i + k];
    memcpy(out->alpha, c->alpha, sizeof out->alpha);
    out->neighb_flag = c->neighb_flag;
    out->nbuild = c->nbuild;
    out->npairs_listed = 2 * (long long)c->npairs;
    return 0;
}

long long sepgpu_get_pairs(sepgpu_ctx *c, int *pairs, long long max_pairs)
{
    if (!c->list_valid) return SEPGPU_ESTATE;
    if (c->npairs > max_pairs) return SEPGPU_EINVAL;
    for (long k = 0; k < c->npairs; k++) {
        const int a = c->pairs[2 * k], b = c->pairs[2 * k + 1];
        pairs[2 * k] = a < b ? a : b;
        pairs[2 * k + 1] = a < b ? b : a;
    }
    return c->npairs;
}

/* SEP_NGPU needs real devices: the mock only satisfies the linker */
/* sampler feeds: the mock keeps no device-side sampler state (the host samplers run on the synchronised atoms[] here) */
static int no_feed(void) { set_error("mock device: no sampler feeds"); return SEPGPU_ESTATE; }
int sepgpu_feed_vacf(sepgpu_ctx *c, int lvec, double *acf_block, int *completed) { (void)c; (void)lvec; (void)acf_block; (void)completed; return no_feed(); }
int sepgpu_feed_msd(sepgpu_ctx *c, int new_origin, char type, const double length[3], int nk, const double *k, double *sums, double *fs)
{ (void)c; (void)new_origin; (void)type; (void)length; (void)nk; (void)k; (void)sums; (void)fs; return no_feed(); }
int sepgpu_feed_profile(sepgpu_ctx *c, char type, double lz, int nbins, double *out4) { (void)c; (void)type; (void)lz; (void)nbins; (void)out4; return no_feed(); }
int sepgpu_feed_fourier(sepgpu_ctx *c, double ly, int nwave, const double *k, double *out16) { (void)c; (void)ly; (void)nwave; (void)k; (void)out16; return no_feed(); }
int sepgpu_feed_radial(sepgpu_ctx *c, double lbox, int lvec, int ntypes, const char *types, long long *counts)
{ (void)c; (void)lbox; (void)lvec; (void)ntypes; (void)types; (void)counts; return no_feed(); }

int sepgpu_set_host_rows(sepgpu_ctx *c, const int *rows) { (void)c; return rows ? SEPGPU_ESTATE : 0; }
int sepgpu_dd_unique_id(void *out128) { (void)out128; set_error("mock device: no decomposition"); return SEPGPU_ESTATE; }
int sepgpu_dd_init(sepgpu_ctx *c, int rank, int nranks, const void *id128, const sepgpu_sys *sys, long long n_global)
{
    (void)c; (void)rank; (void)nranks; (void)id128; (void)sys; (void)n_global;
    set_error("mock device: no decomposition");
    return SEPGPU_ESTATE;
}
int sepgpu_dd_set_owned(sepgpu_ctx *c, int n_own) { (void)c; (void)n_own; return SEPGPU_ESTATE; }
int sepgpu_dd_layers(sepgpu_ctx *c, int *z0, int *z1, int *n_own, int *n_halo)
{
    (void)c; (void)z0; (void)z1; (void)n_own; (void)n_halo;
    return SEPGPU_ESTATE;
}
