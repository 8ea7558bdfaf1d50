// sepgpu_pair.cuh -- pair-kernel parameter blocks and arithmetic helpers shared by the list kernels
// (sepgpu_force.cu) and the tile kernels (sepgpu_force_tile.cu).
#pragma once

#include "sepgpu_internal.cuh"

#include <math.h>

// (struct LJDev and struct BoxDev live in sepgpu_internal.cuh: the context keeps a copy of the last force call's)

// {force factor, energy} of a tabulated pair function at r2: cubic Lagrange interpolation through the four grid points
// around r2 (uniform grid in r^2).  Error <= (3/128) h^4 max|d4/d(r2)4|; with the 32768-point table the host layer
// builds that is < 1e-12 relative for a Lennard-Jones-like function at r >= 0.8 sigma.  *below: r2 under the table's range.
__device__ __forceinline__ double2 table_eval(const LJDev &P, double r2, bool &below)
{
    const double t = (r2 - P.t_lo) * P.t_inv;
    below = t < 0.0;
    int k = (int)t;
    k = k < 1 ? 1 : (k > P.t_n - 3 ? P.t_n - 3 : k);
    const double w = t - (double)k;
    const double wm = w - 1.0, wp = w + 1.0, w2 = w - 2.0;
    const double c0 = -w * wm * w2 * (1.0 / 6.0), c1 = wp * wm * w2 * 0.5, c2 = -wp * w * w2 * 0.5, c3 = wp * w * wm * (1.0 / 6.0);
    const double2 a = __ldg(P.tab + k - 1), b = __ldg(P.tab + k), c = __ldg(P.tab + k + 1), d = __ldg(P.tab + k + 2);
    return make_double2(c0 * a.x + c1 * b.x + c2 * c.x + c3 * d.x, c0 * a.y + c1 * b.y + c2 * c.y + c3 * d.y);
}


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
