/* sep_force.c -- dispatchers for the pair-force entry points.
 *
 * Mirrors the control flow of sep_force_pairs / sep_force_lj / sep_force_dpd (reference
 * source/sepprfrc.c:226-301, 743-780) and sep_coulomb_sf (source/sepcoulomb.c:5-18): cutoff check,
 * brute vs. list, rebuild when sys->neighb_flag is set, then the force evaluation -- which here is a
 * device kernel.  The caller's function pointer is recognised by address and turned into kernel
 * parameters; arbitrary host callbacks cannot run on the device and are rejected.
 */
#include "sep_host.h"

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
    p.cf = cf; p.eps = 1.0; p.sigma = 1.0; p.aw = 1.0;
    if (fun == sep_lj) p.shift = 0.0;
    else if (fun == sep_lj_shift) p.shift = -SEP_LJCF2;    /* u + SEP_LJCF2, whatever cf is (source/sepmisc.c:142) */
    else if (fun == sep_wca) p.shift = -1.0;
    else {
        sep_error("sep_force_pairs: this pair function cannot run on the device "
                  "(supported: sep_lj, sep_lj_shift, sep_wca; or use sep_force_lj)");
        return SEP_FAILURE;
    }

    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    if (sys->neighb_update == SEP_BRUTE) {
        brute_omp_warning(sys);
        /* brute path accumulates epot (source/sepprfrc.c:64) */
        sepb_check(sepgpu_force_lj(b->gpu, &gs, types, &p, opt, 0), "sep_force_pairs");
    } else {
        rebuild_if_flagged(b, sys, &gs, opt);
        /* list path ASSIGNS epot (source/sepprfrc.c:222) */
        sepb_check(sepgpu_force_lj(b->gpu, &gs, types, &p, opt, 1), "sep_force_pairs");
    }
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
