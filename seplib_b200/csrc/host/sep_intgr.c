/* sep_intgr.c -- thermostat and integrator entry points (reference source/sepintgr.c:18-88,
 * 149-198, 296-345).  The arithmetic runs in sepgpu_intgr.cu; this file keeps the host-visible
 * bookkeeping the reference does in the same calls: sys->tnow, sys->neighb_flag,
 * sys->nupdate_neighb, sys->max_dist2, the caller's alpha, and the write-back of atoms[]. */
#include "sep_host.h"
#include <math.h>

static int alpha_slot(sep_binding *b, double *alpha)
{
    for (int k = 0; k < 4; k++)
        if (b->alpha_ptr[k] == alpha) return k;
    for (int k = 0; k < 4; k++)
        if (!b->alpha_ptr[k]) { b->alpha_ptr[k] = alpha; b->alpha_seen[k] = *alpha - 1.0; return k; }
    /* more than four thermostats: recycle slot 3 */
    b->alpha_ptr[3] = alpha; b->alpha_seen[3] = *alpha - 1.0;
    return 3;
}

void sep_nosehoover(sepatom *ptr, double temp0, double *alpha, const double tau, sepsys *sys)
{
    sepdd_allow();
    sep_binding *b = sepb_prepare(ptr, sys);
    const int slot = alpha_slot(b, alpha);
    if (*alpha != b->alpha_seen[slot]) {                 /* first use, or the caller changed it */
        sepb_check(sepgpu_set_alpha(b->gpu, slot, *alpha), "sep_nosehoover");
        b->alpha_seen[slot] = *alpha;
    }
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_nosehoover(b->gpu, &gs, temp0, slot, tau), "sep_nosehoover");
    sepb_dev_newer(b, SEPB_F | SEPB_A);
    if (sep_sync_mode() != SEP_SYNC_LAZY) sepb_pull_scalars(b, sys, NULL, NULL);   /* refreshes *alpha */
    if (sep_sync_mode() == SEP_SYNC_FULL) sepb_download(b, SEPB_F);
}

void _sep_nosehoover_type(seppart *ptr, char type, double Td, double *alpha, const double Q, sepsys *sys)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_nosehoover_type(b->gpu, &gs, type, Td, alpha, Q), "_sep_nosehoover_type");
    sepb_dev_newer(b, SEPB_F | SEPB_A);
    if (sep_sync_mode() == SEP_SYNC_FULL) sepb_download(b, SEPB_F);
}

static void after_integrator(sep_binding *b, sepsys *sys, sepret *ret, int count_update)
{
    sepgpu_scalars s;
    b->last_ret = ret;
    sepb_pull_scalars(b, sys, ret, &s);
    sepb_dev_newer(b, SEPB_X | SEPB_V | SEPB_F | SEPB_A | SEPB_CN | SEPB_CR);
    if (s.neighb_flag) {                                  /* source/sepintgr.c:72-84 */
        sys->neighb_flag = 1;
        if (count_update) sys->nupdate_neighb++;
        b->dev_dirty |= SEPB_XN;
    }
    if (sepb_eager(b)) sepb_download(b, ~0u);
}

void sep_leapfrog(seppart *ptr, sepsys *sys, sepret *retval)
{
    sepdd_allow();
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_leapfrog(b->gpu, &gs), "sep_leapfrog");
    after_integrator(b, sys, retval, 1);
    sys->tnow += sys->dt;                                 /* source/sepintgr.c:86 */
}

void sep_verlet_dpd(seppart *ptr, double lambda, int stepnow, sepsys *sys, sepret *retval)
{
    sep_binding *b = sepb_find(ptr);
    if (b && !b->dpd_state_on_device) { b->dpd_state_on_device = 1; b->host_dirty |= SEPB_PV | SEPB_PA; }
    b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_verlet_dpd(b->gpu, &gs, lambda, stepnow), "sep_verlet_dpd");
    sepb_dev_newer(b, SEPB_PV | SEPB_PA);
    after_integrator(b, sys, retval, 0);                  /* does not count list updates (:336-342) */
}

/* ---- stochastic integrators (reference source/sepintgr.c:89-146, 235-293) ---------------------------------------
 * The Gaussian numbers come from the reference's generator -- polar Box-Muller on glibc rand() with the second
 * deviate cached (source/sepmisc.c:1131-1160) -- drawn HERE, in the reference's order (atom by atom, x y z), so a
 * program seeded the same way sees the same noise as with the reference.  The device applies them. */
double sep_randn(void)
{
    static int have_spare = 0;
    static double spare = 0.0;
    if (have_spare) { have_spare = 0; return spare; }
    double x1, x2, w;
    do {
        x1 = 2.0 * sep_rand() - 1.0;
        x2 = 2.0 * sep_rand() - 1.0;
        w = x1 * x1 + x2 * x2;
    } while (w >= 1.0 || w == 0.0);
    w = sqrt((-2.0 * log(w)) / w);
    spare = x2 * w;
    have_spare = 1;
    return x1 * w;
}

static double *draw_noise(sep_binding *b, const seppart *ptr, long npart)
{
    if (b->noise_cap < (size_t)npart) {
        free(b->noise);
        b->noise = malloc(sizeof(double) * 4 * (size_t)npart);
        if (!b->noise) sep_error("%s: Couldn't allocate memory", (char *)__func__);
        b->noise_cap = (size_t)npart;
    }
    for (long n = 0; n < npart; n++) {
        for (int k = 0; k < 3; k++) b->noise[4 * n + k] = sep_randn();
        b->noise[4 * n + 3] = ptr[n].ldiff;
    }
    return b->noise;
}

void sep_fp(seppart *ptr, double temp_desired, sepsys *sys, sepret *retval)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_fp(b->gpu, &gs, temp_desired, draw_noise(b, ptr, sys->npart)), "sep_fp");
    after_integrator(b, sys, retval, 1);                  /* sys->tnow is not advanced by sep_fp in the reference either */
}

void sep_langevinGJF(sepatom *ptr, double temp0, double alpha, sepsys *sys, sepret *retval)
{
    sep_binding *b = sepb_prepare(ptr, sys);
    sepgpu_sys gs;
    sepb_fill_sys(sys, &gs);
    sepb_check(sepgpu_langevin_gjf(b->gpu, &gs, temp0, alpha, draw_noise(b, ptr, sys->npart)), "sep_langevinGJF");
    after_integrator(b, sys, retval, 1);
    sys->tnow += sys->dt;                                 /* source/sepintgr.c:144 */
}

void sep_set_ldiff(sepatom *ptr, char type, double ldiff, sepsys sys)
{
    for (long n = 0; n < sys.npart; n++)
        if (ptr[n].type == type) ptr[n].ldiff = ldiff;    /* host-only datum: it travels with the noise every step */
}

/* sep_periodic on the HOST copy, for user code that calls it directly (source/sepintgr.c:18-40) */
double sep_periodic(sepatom *atoms, unsigned n, sepsys *sys)
{
    double d2 = 0.0;
    sepatom *a = &atoms[n];
    for (int k = 0; k < 3; k++) {
        if (a->x[k] > sys->length[k]) { a->x[k] -= sys->length[k]; a->cross_neighb[k]++; a->crossings[k]++; }
        else if (a->x[k] < 0.0) { a->x[k] += sys->length[k]; a->cross_neighb[k]--; a->crossings[k]--; }
        const double ri = (a->x[k] + a->cross_neighb[k] * sys->length[k]) - a->xn[k];
        d2 += ri * ri;
    }
    return d2;
}
