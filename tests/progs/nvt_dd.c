/* A prg1-style NVT Lennard-Jones run written against include/sep.h (reference prgs/prg1.c:52-83 is the model), large
 * enough for a slab decomposition: the SAME binary runs on one GPU, or on N with SEP_NGPU=N (tests/test_gpu_dd.py).
 * Prints epot/N, ekin/N, the thermostat variable, the momentum and the rebuild count every 20 steps, writes final.xyz. */
#include "sep.h"

int main(int argc, char **argv)
{
    const int nside = argc > 1 ? atoi(argv[1]) : 24, nsteps = argc > 2 ? atoi(argv[2]) : 200;
    const int natoms = nside * nside * nside;
    const double dens = 0.8, dt = 0.005, temp = 1.0;
    const double lbox = pow(natoms / dens, 1.0 / 3.0);
    double alpha = 0.1;
    sepret ret;
    sepatom *atoms = sep_init(natoms, SEP_NEIGHB);
    sepsys sys = sep_sys_setup(lbox, lbox, lbox, 2.5, dt, natoms, SEP_LLIST_NEIGHBLIST);
    sep_set_lattice(atoms, sys);
    sep_set_vel_seed(atoms, temp, 42, sys);
    FILE *log = fopen("steps.log", "w");            /* opened before the first hot call: only one copy may write it */
    for (int n = 0; n < nsteps; n++) {
        sep_reset_retval(&ret);
        sep_reset_force(atoms, &sys);
        sep_force_pairs(atoms, "AA", 2.5, sep_lj_shift, &sys, &ret, SEP_ALL);
        sep_nosehoover(atoms, temp, &alpha, 0.1, &sys);
        sep_leapfrog(atoms, &sys, &ret);
        if (n % 20 == 0) {
            const double mom = sep_eval_mom(atoms, natoms);
            sep_pressure_tensor(&ret, &sys);
            printf("%d %.10f %.10f %.10f %.8f %1.3e %d\n", n, ret.epot / natoms, ret.ekin / natoms, alpha, ret.p, mom, (int)sys.nupdate_neighb);
            fprintf(log, "%d\n", n);
            /* a host-side edit between hot calls: every copy makes it, each uploads its own atoms */
            if (n == 100) for (int i = 0; i < natoms; i++) atoms[i].v[0] *= 1.0 + 1e-3;
        }
    }
    fclose(log);
    sep_save_xyz(atoms, "A", "final.xyz", "w", &sys);
    sep_close(atoms, natoms);
    sep_free_sys(&sys);
    return 0;
}
