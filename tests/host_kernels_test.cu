// host_kernels_test.cu -- test infrastructure: runs the per-atom arithmetic of the device integrators
// (seplib_b200/csrc/gpu/sepgpu_intgr_atom.cuh, __host__ __device__) on the CPU so that tests/test_cpu_kernels.py can
// compare it with the oracle and the reference's golden loops without a GPU.  Built on demand by the test with
//   nvcc -shared -Xcompiler -fPIC -Iinclude -Iseplib_b200/csrc/gpu tests/host_kernels_test.cu -o tests/_build/libhostk.so
// Nothing here is part of the product.
#include "sepgpu_intgr_atom.cuh"

// one sep_fp (gjf = 0) or sep_langevinGJF (gjf = 1) step over n atoms with the arrays in the reference's layout
// (3 doubles / ints per atom); noise4 = {g0, g1, g2, ldiff} per atom; clpack = packed crossings since the list build.
// out[0] = sum m v^2, out[1] = max displacement^2 of this call.
extern "C" void hostk_stoch_step(int gjf, int n, double *x, double *v, const double *f, const double *m, const double *noise4,
                                 double *prevf, double *randn, const double *xn, int *cross_neighb, int *crossings, int *clpack,
                                 const double *L, double dt, double temp, double alpha, double *out)
{
    const double cc = exp(-alpha * dt);
    double sum = 0.0, mx = 0.0;
    for (int i = 0; i < n; i++) {
        d4 X = {x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.0}, V = {v[3 * i], v[3 * i + 1], v[3 * i + 2], m[i]};
        const d4 F = {f[3 * i], f[3 * i + 1], f[3 * i + 2], 0.0};
        const d4 G = {noise4[4 * i], noise4[4 * i + 1], noise4[4 * i + 2], noise4[4 * i + 3]};
        d4 PF = {prevf[3 * i], prevf[3 * i + 1], prevf[3 * i + 2], 0.0}, RN = {randn[3 * i], randn[3 * i + 1], randn[3 * i + 2], 0.0};
        const d4 XN = {xn[3 * i], xn[3 * i + 1], xn[3 * i + 2], 0.0};
        i4 CR = {cross_neighb[3 * i], cross_neighb[3 * i + 1], cross_neighb[3 * i + 2], clpack[i]};
        int cl[3], t[3] = {0, 0, 0};
        bool changed = false;
        unpack_cl(CR.w, cl[0], cl[1], cl[2]);
        const double d2 = gjf ? stoch_atom<true>(X, V, F, G, PF, RN, XN, CR, cl, t, changed, L[0], L[1], L[2], dt, temp, alpha, cc)
                              : stoch_atom<false>(X, V, F, G, PF, RN, XN, CR, cl, t, changed, L[0], L[1], L[2], dt, temp, alpha, cc);
        if (changed) clpack[i] = pack_cl(cl[0], cl[1], cl[2]);
        x[3 * i] = X.x; x[3 * i + 1] = X.y; x[3 * i + 2] = X.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        if (gjf) {
            prevf[3 * i] = PF.x; prevf[3 * i + 1] = PF.y; prevf[3 * i + 2] = PF.z;
            randn[3 * i] = RN.x; randn[3 * i + 1] = RN.y; randn[3 * i + 2] = RN.z;
        }
        cross_neighb[3 * i] = CR.x; cross_neighb[3 * i + 1] = CR.y; cross_neighb[3 * i + 2] = CR.z;
        for (int k = 0; k < 3; k++) crossings[3 * i + k] += t[k];
        sum += V.x * V.x * m[i] + V.y * V.y * m[i] + V.z * V.z * m[i];
        if (d2 > mx) mx = d2;
    }
    out[0] = sum;
    out[1] = mx;
}
