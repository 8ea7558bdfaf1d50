/* sep_force.c -- dispatchers for the pair-force entry points.
 *
 * Mirrors the control flow of sep_force_pairs / sep_force_lj / sep_force_dpd (reference
 * source/sepprfrc.c:226-301, 743-780) and sep_coulomb_sf (source/sepcoulomb.c:5-18): cutoff check,
 * brute vs. list, rebuild when sys->neighb_flag is set, then the force evaluation -- which here is a
 * device kernel.  The caller's function pointer is recognised by address and turned into kernel
 * parameters; any other pair function is sampled once on a fine r^2 grid and evaluated on the device
 * by cubic interpolation (pair_table below).
 */
#include "sep_host.h"
#include <stdlib.h>
#include <math.h>

/* A pair function of the caller's own -- double fun(double r2, char opt), 'f' = force factor F/r, 'u' = energy, the
 * contract of source/sepmisc.c:113-158 -- cannot be called from a kernel.  It is sampled at SEP_TABLE_N (default 65536)
 * points uniform in r^2 over [rmin^2, cf^2] (rmin = SEP_TABLE_RMIN, default 0.15 cf) and interpolated with a four-point
 * Lagrange cubic: for r^-12-like functions the relative error is below 1e-9 down to r = 0.15 cf and below 1e-12 where
 * pairs of a liquid actually sit.  A pair closer than rmin stops the run with an error instead of
 * extrapolating.  The table is made once per (function, cutoff); sep_pairs_retabulate() drops them when the function's
 * own parameters have changed. */
static const struct sep_pairtab *pair_table(sep_binding *b, double (*fun)(double, char), double cf)
{
    for (int q = 0; q < 4; q++)
        if (b->pairtab[q].fu && b->pairtab[q].fun == fun && b->pairtab[q].cf == cf) return &b->pairtab[q];
    struct sep_pairtab *t = &b->pairtab[b->pairtab_next++ % 4];
    free(t->fu);
    int n = 65536;
    double rmin = 0.15 * cf;
    const char *e;
    if ((e = getenv("SEP_TABLE_N")) && atoi(e) >= 64) n = atoi(e);
    if ((e = getenv("SEP_TABLE_RMIN")) && atof(e) > 0.0 && atof(e) < cf) rmin = atof(e);
    t->fu = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    if (!t->fu) sep_error("sep_force_pairs: out of memory for the pair-function table");
    t->fun = fun; t->cf = cf; t->n = n; t->r2_lo = rmin * rmin;
    const double h = (cf * cf - t->r2_lo) / (double)(n - 1);
    for (int k = 0; k < n; k++) {
        const double r2 = t->r2_lo + h * (double)k;
        t->fu[2 * k] = fun(r2, 'f');
        t->fu[2 * k + 1] = fun(r2, 'u');
        if (!isfinite(t->fu[2 * k]) || !isfinite(t->fu[2 * k + 1]))
            sep_error("sep_force_pairs: the pair function is not finite on [rmin, cf] (set SEP_TABLE_RMIN)");
    }
    return t;
}

void sep_pairs_retabulate(void)
{
    for (sep_binding *b = sepb_first(); b; b = b->next)
        for (int q = 0; q < 4; q++) { free(b->pairtab[q].fu); b->pairtab[q].fu = NULL; }
}

static void rebuild_if_flagged(sep_binding *b, sepsys *sys, const sepgpu_sys *gs, unsigned opt)
{
    if (sys->neighb_flag != 1) return;
    /* SEP_NEIGHBLIST (all-pairs builder, source/sepprfrc.c:306-343) and SEP_LLIST_NEIGHBLIST produce
     * the same pair set whenever the cell grid is valid; both map to the cell-binned device build. */
    sepb_check(sepgpu_neighb_build(b->gpu, gs, opt), "sep_neighb");
    sys->neighb_flag = 0;
}

static void brute_omp_warning(sepsys *sys)
{
    if (sys->omp_flag) {                                   /* source/sepprfrc.c:242-246 */
        sep_warning("omp flag set, SEP_BRUTE does not support threads.");
        sep_warning("Resetting omp flag");
        sys->omp_flag = false;
    }
}

int sep_force_pairs(seppart *ptr, const char *types, double cf, double (*fun)(double, char),
                    sepsys *sys, sepret *retval, const unsigned opt)
{
    if (cf > sys->cf)
        sep_error("cutoff for an interaction cannot be larger than maximum cutoff");

    sepgpu_ljparam p;
    int builtin = 1;
    p.cf = cf; p.eps = 1.0; p.sigma = 1.0; p.aw = 1.0; p.shift = 0.0;
    if (fun == sep_lj) p.shift = 0.0;
    else if (fun == sep_lj_shift) p.shift = -SEP_LJCF2;    /* u + SEP_LJCF2, whatever cf is (source/sepmisc.c:142) */
    else if (fun == sep_wca) p.shift = -1.0;
    else if (fun) builtin = 0;
    else {
        sep_error("sep_force_pairs: no pair function given");
        return SEP_FAILURE;
    }

    if (builtin) sepdd_allow();
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    const struct sep_pairtab *t = builtin ? NULL : pair_table(b, fun, cf);
    /* brute path accumulates epot (source/sepprfrc.c:64); list path ASSIGNS it (source/sepprfrc.c:222) */
    const int assign = sys->neighb_update != SEP_BRUTE;
    if (!assign) brute_omp_warning(sys);
    else rebuild_if_flagged(b, sys, &gs, opt);
    if (builtin) sepb_check(sepgpu_force_lj(b->gpu, &gs, types, &p, opt, assign), "sep_force_pairs");
    else sepb_check(sepgpu_force_table(b->gpu, &gs, types, cf, t->fu, t->n, t->r2_lo, opt, assign), "sep_force_pairs");
    sepb_after_force(b, sys, retval);
    return SEP_SUCCESS;
}

void sep_force_lj(seppart *ptr, const char *types, const double *param, sepsys *sys,
                  sepret *retval, const unsigned opt)
{
    /* param = {cf, eps, sigma, aw} in the order the code reads them (source/sepprfrc.c:785) */
    const double cf = param[0], eps = param[1], sigma = param[2], aw = param[3];
    if (cf > sys->cf)
        sep_error("cutoff for an interaction cannot be larger than maximum cutoff");
    sepgpu_ljparam p;
    p.cf = cf; p.eps = eps; p.sigma = sigma; p.aw = aw;
    p.shift = 4.0 * eps * (pow(sigma / cf, 12.) - aw * pow(sigma / cf, 6.));

    sepdd_allow();
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    if (sys->neighb_update == SEP_BRUTE) brute_omp_warning(sys);
    else rebuild_if_flagged(b, sys, &gs, opt);
    /* both sep_lj_pair_brute and sep_lj_pair_neighb accumulate epot (source/sepprfrc.c:922, 964) */
    sepb_check(sepgpu_force_lj(b->gpu, &gs, types, &p, opt, 0), "sep_force_lj");
    sepb_after_force(b, sys, retval);
}

void sep_coulomb_sf(seppart *ptr, double cf, sepsys *sys, sepret *retval, const unsigned opt)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_coulomb_sf(b->gpu, &gs, cf, opt), "sep_coulomb_sf");
    sepb_after_force(b, sys, retval);
}

void sep_force_dpd(seppart *ptr, const char *types, const double cf, const double aij,
                   const double temp_desired, const double sigma, sepsys *sys, sepret *retval,
                   const unsigned opt)
{
    sep_binding *b = sepb_find(ptr);
    if (b && !b->dpd_state_on_device) {
        /* pv (predicted velocity) feeds the dissipative force: needs to be on the device from now on */
        b->dpd_state_on_device = 1;
        b->host_dirty |= SEPB_PV | SEPB_PA;
    }
    b = sepb_prepare(ptr, sys);
    if (!b->dpd_state_on_device) { b->dpd_state_on_device = 1; b->host_dirty |= SEPB_PV | SEPB_PA; b = sepb_prepare(ptr, sys); }
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    if (sys->neighb_update == SEP_BRUTE) brute_omp_warning(sys);
    else rebuild_if_flagged(b, sys, &gs, opt);            /* the DPD list routine builds on demand (:1021-1031) */
    sepb_check(sepgpu_force_dpd(b->gpu, &gs, types, cf, aij, temp_desired, sigma, opt,
                                sep_dpd_seed(), b->dpd_calls++), "sep_force_dpd");
    sepb_after_force(b, sys, retval);
}

/* explicit list builders (source/sepprfrc.c:347-378) */
static void build_now(seppart *ptr, sepsys *sys, unsigned opt)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_neighb_build(b->gpu, &gs, opt), "sep_neighb");
}
void sep_neighb(seppart *ptr, sepsys *sys) { build_now(ptr, sys, SEP_ALL); }
void sep_neighb_nonbonded(seppart *ptr, sepsys *sys) { build_now(ptr, sys, SEP_EXCL_BONDED); }
void sep_neighb_excl_same_mol(seppart *ptr, sepsys *sys) { build_now(ptr, sys, SEP_EXCL_SAME_MOL); }

/* bonded-partner predicates on the host tables (source/sepprfrc.c:703-740) */
static unsigned share(const int *ra, const int *rb, int width, int a, int b)
{
    for (int k = 0; k < width; k++) {
        if (ra[k] == -1 || rb[k] == -1) break;
        if (ra[k] == b || rb[k] == a) return 1;
    }
    return 0;
}
unsigned int sep_bond_share(seppart *p, int a, int b) { return share(p[a].bond, p[b].bond, SEP_BOND, a, b); }
unsigned int sep_angle_share(seppart *p, int a, int b) { return share(p[a].angle, p[b].angle, SEP_ANGLE, a, b); }
unsigned int sep_dihed_share(seppart *p, int a, int b) { return share(p[a].dihed, p[b].dihed, SEP_DIHED, a, b); }
unsigned int sep_bonded(seppart *p, int a, int b)
{
    return sep_bond_share(p, a, b) + sep_angle_share(p, a, b) + sep_dihed_share(p, a, b);
}
