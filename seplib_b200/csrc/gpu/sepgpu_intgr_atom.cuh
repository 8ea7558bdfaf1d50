// sepgpu_intgr_atom.cuh -- per-atom arithmetic of the integrators as __host__ __device__ functions, so that the
// same code the kernels run can be exercised on the CPU (tests/host_kernels_test.cu compares it with the oracle).
#pragma once

#include "sepgpu_internal.cuh"

#include <math.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

__host__ __device__ __forceinline__ int pack_cl(int cx, int cy, int cz) { return (cx + 512) | ((cy + 512) << 10) | ((cz + 512) << 20); }
__host__ __device__ __forceinline__ void unpack_cl(int w, int &cx, int &cy, int &cz)
{
    if (w == 0) { cx = cy = cz = 0; return; }
    cx = (w & 1023) - 512; cy = ((w >> 10) & 1023) - 512; cz = ((w >> 20) & 1023) - 512;
}

// one component of sep_periodic (source/sepintgr.c:21-35); returns the squared displacement term
__host__ __device__ __forceinline__ double periodic_1d(double &x, double L, int &cn, int &cl, int &cross_total, bool &changed, double xn)
{
    if (x > L) { x -= L; cn++; cl++; cross_total = 1; changed = true; }
    else if (x < 0.0) { x += L; cn--; cl--; cross_total = -1; changed = true; }
    const double ri = (x + cn * L) - xn;
    return ri * ri;
}

// One atom of sep_fp (GJF = false, source/sepintgr.c:235-293) or sep_langevinGJF (GJF = true, :89-146).
// g = {gaussian x, y, z, ldiff}; pf / rn = previous force / previous noise (GJF only, updated in place).
// cr = {cross_neighb x, y, z, packed crossings since the list build}; t[] receives the crossing of this step.
// Returns the squared displacement from the last-list-build position xn.
template <bool GJF>
__host__ __device__ __forceinline__ double stoch_atom(d4 &x, d4 &v, const d4 &f, const d4 &g, d4 &pf, d4 &rn, const d4 &xn, i4 &cr,
                                                      int cl[3], int t[3], bool &changed, double Lx, double Ly, double Lz,
                                                      double dt, double temp, double alpha, double cc)
{
    const double m = v.w;
    double d2 = 0.0;
    if (!GJF) {
        const double im = 1.0 / m;
        const double fric = temp / g.w;                              // :245  (g.w = ldiff)
        const double gaussfac = sqrt(24 * temp * fric / dt);         // :246
        const double fac = sqrt(1.0 / 12.0);
        const double ax = g.x * fac * gaussfac, ay = g.y * fac * gaussfac, az = g.z * fac * gaussfac;   // :252
        x.x += dt * v.x; v.x += im * dt * (f.x - fric * v.x + ax);   // :254-255
        d2 += periodic_1d(x.x, Lx, cr.x, cl[0], t[0], changed, xn.x);
        x.y += dt * v.y; v.y += im * dt * (f.y - fric * v.y + ay);
        d2 += periodic_1d(x.y, Ly, cr.y, cl[1], t[1], changed, xn.y);
        x.z += dt * v.z; v.z += im * dt * (f.z - fric * v.z + az);
        d2 += periodic_1d(x.z, Lz, cr.z, cl[2], t[2], changed, xn.z);
    } else {
        const double imass = 1.0 / m, imass2 = 0.5 * imass;
        const double fac = sqrt(temp * (1.0 - cc * cc));             // :100
        const double c_ = alpha * dt * imass2;
        const double a = (1.0 - c_) / (1.0 + c_), b = 1.0 / (1.0 + c_);
        v.x = a * v.x + dt * imass2 * (a * pf.x + f.x) + b * imass * rn.x;     // :108
        v.y = a * v.y + dt * imass2 * (a * pf.y + f.y) + b * imass * rn.y;
        v.z = a * v.z + dt * imass2 * (a * pf.z + f.z) + b * imass * rn.z;
        pf.x = f.x; pf.y = f.y; pf.z = f.z;                                    // :111
        rn.x = fac * g.x; rn.y = fac * g.y; rn.z = fac * g.z;                   // :115
        x.x += b * dt * v.x + b * dt * dt * imass2 * f.x + b * dt * imass2 * rn.x;   // :117
        x.y += b * dt * v.y + b * dt * dt * imass2 * f.y + b * dt * imass2 * rn.y;
        x.z += b * dt * v.z + b * dt * dt * imass2 * f.z + b * dt * imass2 * rn.z;
        d2 += periodic_1d(x.x, Lx, cr.x, cl[0], t[0], changed, xn.x);
        d2 += periodic_1d(x.y, Ly, cr.y, cl[1], t[1], changed, xn.y);
        d2 += periodic_1d(x.z, Lz, cr.z, cl[2], t[2], changed, xn.z);
    }
    return d2;
}
