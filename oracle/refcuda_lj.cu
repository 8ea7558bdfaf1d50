// oracle/refcuda_lj.cu -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
//
// Times the reference's OWN CUDA Lennard-Jones path (reference cuda/sepcuda*.cu, single precision, its public
// sep_cuda_* API) on the GPU this runs on, as a second stated baseline next to the reference's CPU path
// ("competitor").  The loop is the one of the reference's benchmark program
// cuda/tgpu_0.cu:17-29 -- reset, list update, pair force, leapfrog, list check every second step -- with the number of
// steps taken from the command line instead of the 100 000 hard-wired there, and a device-synchronised clock around it.
// Compiled by oracle/Makefile (target refcuda) against the reference sources where they lie; nothing of them is copied.
//
//   refcuda_lj <start.xyz> <steps> [warmup]      prints: natoms N steps K seconds S
#include "sepcuda.h"

#include <chrono>

static void run(sepcupart *ptr, sepcusys *sptr, int nloops)
{
    for (int n = 0; n < nloops; n++) {
        sep_cuda_reset_iteration(ptr);
        sep_cuda_update_neighblist(ptr, 2.5);
        sep_cuda_force_lj(ptr);
        sep_cuda_integrate_leapfrog(ptr);
        if (n % 2 == 0) sep_cuda_check_neighblist(ptr, sptr->skin);
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <start.xyz> <steps> [warmup]\n", argv[0]);
        return 2;
    }
    const int steps = atoi(argv[2]), warmup = argc > 3 ? atoi(argv[3]) : 200;
    sepcupart *ptr = sep_cuda_load_xyz(argv[1]);
    sepcusys *sptr = sep_cuda_sys_setup(ptr);
    run(ptr, sptr, warmup);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "device error after the warm-up\n"); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    run(ptr, sptr, steps);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "device error in the timed loop\n"); return 1; }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("natoms %u steps %d seconds %.6f\n", ptr->npart, steps, secs);
    sep_cuda_free_memory(ptr);
    return 0;
}
