// sepgpu_pair.cuh -- pair-kernel parameter blocks and arithmetic helpers shared by the list kernels
// (sepgpu_force.cu) and the tile kernels (sepgpu_force_tile.cu).
#pragma once

#include "sepgpu_internal.cuh"

#include <math.h>

struct LJDev {
    double cf2, sig2, eps48, eps4, aw, awh, shift;
    int t0, t1;
};

struct BoxDev { double Lx, Ly, Lz; };

// 1/x to ~1 ulp: MUFU.RCP64H seed (relative error <= 2^-23) + two Newton steps (4 DFMA) instead of the
// IEEE division sequence; inputs are r^2 of in-range pairs, far from denormals/inf.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// Same seed, ONE third-order step: y1 = y0 (1 + e + e^2), e = 1 - x y0.  The error goes 2^-23 -> 2^-69,
// below double rounding, in 3 DFMA instead of 4.
__device__ __forceinline__ double fast_rcp3(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double e2 = fma(e, e, e);
    return fma(y, e2, y);
}

__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // Newton: y <- y + y*(0.5 - 0.5*x*y*y)
#pragma unroll
    for (int it = 0; it < 2; it++) {
        double h = 0.5 * y;
        double e = fma(-x * y, h, 0.5);
        y = fma(y, e, y);
    }
    // one more correction step for full double accuracy
    double h = 0.5 * y;
    double e = fma(-x * y, h, 0.5);
    y = fma(y, e, y);
    return y;
}

__device__ __forceinline__ void apply_image(int code, const BoxDev &B, double &dx, double &dy, double &dz)
{
    // code = (sx+1) + 3(sy+1) + 9(sz+1); s = +1 means the reference's sep_Wrap subtracted L
    const int sx = code % 3 - 1, sy = (code / 3) % 3 - 1, sz = code / 9 - 1;
    dx -= sx * B.Lx; dy -= sy * B.Ly; dz -= sz * B.Lz;
}

// ---- Lennard-Jones, Verlet list ------------------------------------------------------------------------
struct PairAcc {
    double fx, fy, fz;                       // per atom, in units of 48 eps
    double u;                                // per thread
    int nin;
    double v[6];                             // per thread: xx xy xz yy yz zz, touched once per atom + on boundary pairs
};

__device__ __forceinline__ void virial_add(double *v, double gx, double gy, double gz, double sx, double sy, double sz)
{
    v[0] = fma(gx, sx, v[0]); v[1] = fma(gx, sy, v[1]); v[2] = fma(gx, sz, v[2]);
    v[3] = fma(gy, sy, v[3]); v[4] = fma(gy, sz, v[4]); v[5] = fma(gz, sz, v[5]);
}

int sepgpu_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags);
