/* The prg1-style NVT Lennard-Jones loop written against include/sep.h (reference prgs/prg1.c:52-70 is the model), with a
 * clock around it: what an unchanged seplib program gets per step through the sep_* API -- on one GPU, or on N with
 * SEP_NGPU=N (bench.py records it at N = 2 as e2e_sep_ngpu).  Prints one line:
 *   natoms N steps K warm W seconds_warm (first hot call: device start-up, fork, upload; + W-1 steps) seconds_loop (K steps,
 *   sepret refreshed after every call) seconds_download (atoms[] back on the host) epot/N ekin/N rebuilds */
#define _POSIX_C_SOURCE 200809L
#include "sep.h"
#include <time.h>

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void step(sepatom *atoms, sepsys *sys, sepret *ret, double temp, double *alpha)
{
    sep_reset_retval(ret);
    sep_reset_force(atoms, sys);
    sep_force_pairs(atoms, "AA", 2.5, sep_lj_shift, sys, ret, SEP_ALL);
    sep_nosehoover(atoms, temp, alpha, 0.01, sys);
    sep_leapfrog(atoms, sys, ret);
}

int main(int argc, char **argv)
{
    const int nside = argc > 1 ? atoi(argv[1]) : 100, nsteps = argc > 2 ? atoi(argv[2]) : 1000, nwarm = argc > 3 ? atoi(argv[3]) : 300;
    const int natoms = nside * nside * nside;
    const double dens = 0.8, dt = 0.005, temp = 1.0;
    const double lbox = pow(natoms / dens, 1.0 / 3.0);
    double alpha = 0.1;
    sepret ret;
    sepatom *atoms = sep_init(natoms, 0);
    sepsys sys = sep_sys_setup(lbox, lbox, lbox, 2.5, dt, natoms, SEP_LLIST_NEIGHBLIST);
    sep_set_lattice(atoms, sys);
    sep_set_vel_seed(atoms, temp, 42, sys);
    const double t0 = now();
    for (int n = 0; n < nwarm; n++) step(atoms, &sys, &ret, temp, &alpha);
    const double t1 = now();
    const int nb1 = (int)sys.nupdate_neighb;
    for (int n = 0; n < nsteps; n++) step(atoms, &sys, &ret, temp, &alpha);
    const double t2 = now();
    const double mom = sep_eval_mom(atoms, natoms);      /* a library reader: brings atoms[] back to the host */
    const double t3 = now();
    printf("natoms %d steps %d warm %d seconds_warm %.6f seconds_loop %.6f seconds_download %.6f epot_per_atom %.10f ekin_per_atom %.10f "
           "rebuilds %d momentum %1.3e\n", natoms, nsteps, nwarm, t1 - t0, t2 - t1, t3 - t2, ret.epot / natoms, ret.ekin / natoms,
           (int)sys.nupdate_neighb - nb1, mom);
    sep_close(atoms, natoms);
    sep_free_sys(&sys);
    return 0;
}
