// sepgpu_force2.cu -- shifted-force Coulomb and DPD pair kernels.
//
// Stand-ins for sep_coulomb_sf_{neighb,brute} (reference source/sepcoulomb.c:96-160, 20-94) and
// sep_dpdforce_{neighb,brute} (source/sepprfrc.c:1007-1133, 1135-1231).  Same ownership scheme as
// the Lennard-Jones kernel in sepgpu_force.cu: full list, TPA lanes per atom, register accumulation,
// shuffle reduction, one store per atom, per-block partial rows for energy and virial.
#include "sepgpu_internal.cuh"

#include <math.h>
#include <float.h>

#define FORCE_BLOCK 128
#define FORCE_MAX_GRID (148 * 16)

struct BoxC { double Lx, Ly, Lz; };

int sepgpu_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags);
int sepgpu_ensure_dpd(sepgpu_ctx *c);
int sepgpu_need_global_rows(sepgpu_ctx *c, const sepgpu_sys *sys);
int sepgpu_dd_refresh_charges(sepgpu_ctx *c);
int sepgpu_dd_halo_update(sepgpu_ctx *c, const sepgpu_sys *sys);

__device__ __forceinline__ double rsqrt_nr(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 3; it++) {            // y <- y + y*(0.5 - 0.5 x y^2)
        const double h = 0.5 * y;
        const double e = fma(-x * y, h, 0.5);
        y = fma(y, e, y);
    }
    return y;
}

__device__ __forceinline__ void apply_image_c(int code, const BoxC &B, double &dx, double &dy, double &dz)
{
    const int sx = code % 3 - 1, sy = (code / 3) % 3 - 1, sz = code / 9 - 1;
    dx -= sx * B.Lx; dy -= sy * B.Ly; dz -= sz * B.Lz;
}

__global__ void k_sort_charges(const double *__restrict__ z, const int *__restrict__ order, double *__restrict__ zs, int n)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) zs[s] = z[order[s]];
}

// ---- Coulomb, Verlet list ----------------------------------------------------------------------------------------
template <int TPA, bool STORE>
__global__ void __launch_bounds__(FORCE_BLOCK)
k_coulomb_list(const d4 *__restrict__ xs, const double *__restrict__ zs, const unsigned *__restrict__ nbr,
               const int *__restrict__ cnt, const int *__restrict__ order, d4 *__restrict__ f4, int n, int npad,
               double cf, BoxC B, double *__restrict__ partial, double *fij, int nmol)
{
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    const int sub = threadIdx.x % TPA;
    const int groups_per_block = FORCE_BLOCK / TPA;
    const double cf2 = cf * cf, icf2 = 1.0 / cf2, icf = 1.0 / cf;
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;

    for (int s0 = blockIdx.x * groups_per_block; s0 < n; s0 += gridDim.x * groups_per_block) {
        const int s = s0 + threadIdx.x / TPA;
        const bool valid = s < n;
        double fx = 0.0, fy = 0.0, fz = 0.0;
        if (valid) {
            const d4 pi = xs[s];
            const double zi = zs[s];
            const int m = cnt[s];
#pragma unroll 2
            for (int k = sub; k < m; k += TPA) {
                const unsigned e = nbr[nbr_index(k, s, npad)];
                const int j = (int)(e & SEPGPU_INDEX_MASK);
                const d4 pj = xs[j];
                const double zj = zs[j];
                double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const int code = (int)(e >> SEPGPU_SHIFT_BITS);
                if (code != 13) apply_image_c(code, B, dx, dy, dz);
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                // the reference skips list owners with |z| < DBL_EPSILON (source/sepcoulomb.c:106); a pair
                // with either charge below that contributes at most ~1e-16 z, so the product test is enough
                const double zizj = zi * zj;
                if (r2 < cf2 && zizj != 0.0) {
                    const double rinv = rsqrt_nr(r2);
                    const double r = r2 * rinv;
                    const double ft = zizj * (rinv * rinv - icf2) * rinv;       // :121
                    const double gx = ft * dx, gy = ft * dy, gz = ft * dz;
                    fx += gx; fy += gy; fz += gz;
                    acc[1] += zizj * (rinv + (r - cf) * icf2 - icf);            // :150
                    acc[2] = fma(gx, dx, acc[2]); acc[3] = fma(gx, dy, acc[3]); acc[4] = fma(gx, dz, acc[4]);
                    acc[5] = fma(gy, dy, acc[5]); acc[6] = fma(gy, dz, acc[6]); acc[7] = fma(gz, dz, acc[7]);
                    if (fij) {                                                   // :138-147 (no mi != mj test here)
                        const int mi = tag_mol(pi.w), mj = tag_mol(pj.w);
                        if (mi != -1 && mj != -1) {
                            double *t = fij + ((size_t)mi * nmol + mj) * 3;
                            atomicAdd(t, gx); atomicAdd(t + 1, gy); atomicAdd(t + 2, gz);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = TPA / 2; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (valid && sub == 0) {
            const int i = order[s];
            if (STORE) { d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0.0; f4[i] = o; }
            else { d4 o = f4[i]; o.x += fx; o.y += fy; o.z += fz; f4[i] = o; }
        }
    }
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

// ---- Coulomb, Verlet list, second kernel (option coulomb_kernel = 2) ---------------------------------------------------
// Same sums as k_coulomb_list, organised like the Lennard-Jones list kernel (sepgpu_force.cu):
//  * the charge rides in .w of a per-call copy xq[s] = {x, y, z, z_s} of the sorted coordinates, so a pair costs ONE
//    32-byte gather instead of a position gather plus a charge gather (k_make_xq: 64 B/atom per call);
//  * four list entries per 128-bit streaming load, issued one chunk ahead;
//  * two pairs per straight-line block, branch-free: out-of-range pairs run the same arithmetic with the partner
//    charge selected to zero (the reference's |z| < DBL_EPSILON skip, source/sepcoulomb.c:106, removes exact zeros);
//  * 1/r from the MUFU seed and ONE third-order step (5 FP64 operations instead of 12);
//  * z_i is applied once per atom, and the virial comes from the full-list identity
//        sum_ij g_ij (x) d_ij = 2 sum_i F_i (x) x_i - sum_ij g_ij (x) S_ij      (see lj_pair in sepgpu_force.cu).
// No molecule-pair table here: contexts that carry Fij stay on k_coulomb_list.
struct CoulDev { double cf2, icf2, twoicf; };

__device__ __forceinline__ double rsqrt_3rd(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // x y^2 = 1 - e  =>  1/sqrt(x) = y (1 + e/2 + 3 e^2/8 + O(e^3));  |e| <= 2^-19 leaves < 2^-58
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5) * e;
    return fma(y, p, y);
}

struct CoulAcc {
    double fx, fy, fz, ua;                   // per atom, in units of z_i
    double v[6];                             // per thread
};

__device__ __forceinline__ void coul_virial(double *v, double gx, double gy, double gz, double sx, double sy, double sz)
{
    v[0] = fma(gx, sx, v[0]); v[1] = fma(gx, sy, v[1]); v[2] = fma(gx, sz, v[2]);
    v[3] = fma(gy, sy, v[3]); v[4] = fma(gy, sz, v[4]); v[5] = fma(gz, sz, v[5]);
}

__device__ __forceinline__ void coul_pair2(const d4 &pi, const d4 &p0, const d4 &p1, unsigned e0, unsigned e1, bool valid1,
                                           const CoulDev &P, const BoxC &B, CoulAcc &A)
{
    double dx0 = pi.x - p0.x, dy0 = pi.y - p0.y, dz0 = pi.z - p0.z;
    double dx1 = pi.x - p1.x, dy1 = pi.y - p1.y, dz1 = pi.z - p1.z;
    const int c0 = (int)(e0 >> SEPGPU_SHIFT_BITS), c1 = valid1 ? (int)(e1 >> SEPGPU_SHIFT_BITS) : 13;
    const bool shifted = (c0 != 13) | (c1 != 13);
    if (shifted) { apply_image_c(c0, B, dx0, dy0, dz0); apply_image_c(c1, B, dx1, dy1, dz1); }
    const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
    const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
    const bool in0 = __double_as_longlong(r20) < __double_as_longlong(P.cf2);                 // r2 < cf2, :118
    const bool in1 = (__double_as_longlong(r21) < __double_as_longlong(P.cf2)) && valid1;
    const double q0 = in0 ? p0.w : 0.0, q1 = in1 ? p1.w : 0.0;                               // partner charge, masked
    const double y0 = rsqrt_3rd(r20), y1 = rsqrt_3rd(r21);
    const double f0 = q0 * (fma(y0, y0, -P.icf2) * y0);             // z_j (1/r^2 - 1/cf^2)/r, :121
    const double f1 = q1 * (fma(y1, y1, -P.icf2) * y1);
    const double ra = r20 * y0, rb = r21 * y1;                      // r
    const double u0 = q0 * (y0 + fma(ra, P.icf2, -P.twoicf));       // z_j (1/r + (r - cf)/cf^2 - 1/cf), :150
    const double u1 = q1 * (y1 + fma(rb, P.icf2, -P.twoicf));
    A.fx = fma(f0, dx0, A.fx); A.fy = fma(f0, dy0, A.fy); A.fz = fma(f0, dz0, A.fz);
    A.fx = fma(f1, dx1, A.fx); A.fy = fma(f1, dy1, A.fy); A.fz = fma(f1, dz1, A.fz);
    A.ua += u0; A.ua += u1;
    if (shifted) {                                                  // boundary-crossing pairs: - g (x) S
        const double zi = pi.w;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        apply_image_c(c0, B, sx, sy, sz);
        double g = zi * f0;
        coul_virial(A.v, g * dx0, g * dy0, g * dz0, sx, sy, sz);
        sx = sy = sz = 0.0;
        apply_image_c(c1, B, sx, sy, sz);
        g = zi * f1;
        coul_virial(A.v, g * dx1, g * dy1, g * dz1, sx, sy, sz);
    }
}

__global__ void k_make_xq(const d4 *__restrict__ xs, const double *__restrict__ z, const int *__restrict__ order,
                          d4 *__restrict__ xq, int n)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    d4 p = xs[s];
    p.w = z[order[s]];
    xq[s] = p;
}

// MINB: CTAs per SM the register budget is cut for
template <bool STORE, int MINB>
__global__ void __launch_bounds__(FORCE_BLOCK, MINB)
k_coulomb_list2(const d4 *__restrict__ xq, const unsigned *__restrict__ nbr, const int *__restrict__ cnt,
                const int *__restrict__ order, d4 *__restrict__ f4, int n, int npad, int atoms_per_cta,
                CoulDev P, BoxC B, double *__restrict__ partial)
{
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    CoulAcc A;
#pragma unroll
    for (int q = 0; q < 6; q++) A.v[q] = 0.0;
    double usum = 0.0;
    const int first = blockIdx.x * atoms_per_cta;
    const int last = min(n, first + atoms_per_cta);
    const uint4 *nbrv = reinterpret_cast<const uint4 *>(nbr);

    for (int s = first + (int)threadIdx.x; s < last; s += FORCE_BLOCK) {
        const d4 pi = xq[s];
        const int m = cnt[s];
        const int nch = (m + 3) >> 2;
        const uint4 *row = nbrv + s;
        A.fx = A.fy = A.fz = 0.0; A.ua = 0.0;
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (nch > 0) cur = __ldcs(row);
        for (int c = 0; c < nch; c++) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c + 1 < nch) nxt = __ldcs(row + (size_t)(c + 1) * npad);
            const int left = m - 4 * c;                          // >= 1 valid entries in this chunk
            {
                const bool v1 = left > 1;
                const d4 p0 = xq[cur.x & SEPGPU_INDEX_MASK];
                const d4 p1 = xq[(v1 ? cur.y : cur.x) & SEPGPU_INDEX_MASK];
                coul_pair2(pi, p0, p1, cur.x, cur.y, v1, P, B, A);
            }
            if (left > 2) {
                const bool v3 = left > 3;
                const d4 p2 = xq[cur.z & SEPGPU_INDEX_MASK];
                const d4 p3 = xq[(v3 ? cur.w : cur.z) & SEPGPU_INDEX_MASK];
                coul_pair2(pi, p2, p3, cur.z, cur.w, v3, P, B, A);
            }
            cur = nxt;
        }
        const double zi = pi.w;
        const double fx = zi * A.fx, fy = zi * A.fy, fz = zi * A.fz;
        usum = fma(zi, A.ua, usum);
        const int i = order[s];
        if (STORE) { d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0.0; f4[i] = o; }
        else { d4 o = f4[i]; o.x += fx; o.y += fy; o.z += fz; f4[i] = o; }
        coul_virial(A.v, fx + fx, fy + fy, fz + fz, pi.x, pi.y, pi.z);      // 2 F_i (x) x_i
    }
    double acc[SEPGPU_NPART_F];
    acc[0] = 0.0;
    acc[1] = usum;
#pragma unroll
    for (int q = 0; q < 6; q++) acc[2 + q] = A.v[q];
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
    }
}

__device__ __forceinline__ int share_tab_c(const int *__restrict__ tab, int width, int a, int b)
{
    for (int k = 0; k < width; k++) {
        int ta = tab[a * width + k], tb = tab[b * width + k];
        if (ta == -1 || tb == -1) break;
        if (ta == b || tb == a) return 1;
    }
    return 0;
}

// ---- Coulomb, all pairs ---------------------------------------------------------------------------------------------
template <bool STORE>
__global__ void __launch_bounds__(FORCE_BLOCK)
k_coulomb_brute(const d4 *__restrict__ x4, const double *__restrict__ z, d4 *__restrict__ f4, int n, double cf,
                BoxC B, unsigned opt, const int *__restrict__ eb, const int *__restrict__ ea,
                const int *__restrict__ ed, double *__restrict__ partial, double *fij, int nmol)
{
    __shared__ d4 tile[FORCE_BLOCK];
    __shared__ double ztile[FORCE_BLOCK];
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    const int i = blockIdx.x * FORCE_BLOCK + threadIdx.x;
    const bool valid = i < n;
    d4 pi; pi.x = pi.y = pi.z = 0; pi.w = 0;
    double zi = 0.0;
    if (valid) { pi = x4[i]; zi = z[i]; }
    const int mi = tag_mol(pi.w);
    const double cf2 = cf * cf, icf2 = 1.0 / cf2, icf = 1.0 / cf;
    const double hx = 0.5 * B.Lx, hy = 0.5 * B.Ly, hz = 0.5 * B.Lz;
    double fx = 0, fy = 0, fz = 0;
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    for (int j0 = 0; j0 < n; j0 += FORCE_BLOCK) {
        __syncthreads();
        if (j0 + threadIdx.x < n) { tile[threadIdx.x] = x4[j0 + threadIdx.x]; ztile[threadIdx.x] = z[j0 + threadIdx.x]; }
        __syncthreads();
        const int lim = min(FORCE_BLOCK, n - j0);
        if (!valid) continue;
        for (int t = 0; t < lim; t++) {
            const int j = j0 + t;
            if (j == i) continue;
            const d4 pj = tile[t];
            const double zj = ztile[t];
            // the reference tests only the lower-index atom n of the pair (source/sepcoulomb.c:30)
            const double zlow = i < j ? zi : zj;
            if (fabs(zlow) < DBL_EPSILON) continue;
            if (opt == SEPGPU_EXCL_SAME_MOL) { if (mi == tag_mol(pj.w) && mi != -1) continue; }
            else if (opt == SEPGPU_EXCL_BONDED) {
                const int a = min(i, j), b = max(i, j);
                if (share_tab_c(eb, 10, a, b) + share_tab_c(ea, 10, a, b) + share_tab_c(ed, 20, a, b) == 1) continue;   // :37
            }
            const double dx = wrap_exact(pi.x - pj.x, B.Lx, hx);
            const double dy = wrap_exact(pi.y - pj.y, B.Ly, hy);
            const double dz = wrap_exact(pi.z - pj.z, B.Lz, hz);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (r2 < cf2) {
                const double zizj = zi * zj;
                const double r = sqrt(r2);
                const double ft = zizj * (1.0 / r2 - icf2) / r;
                const double gx = ft * dx, gy = ft * dy, gz = ft * dz;
                fx += gx; fy += gy; fz += gz;
                acc[1] += zizj * (1.0 / r + (r - cf) * icf2 - icf);
                acc[2] = fma(gx, dx, acc[2]); acc[3] = fma(gx, dy, acc[3]); acc[4] = fma(gx, dz, acc[4]);
                acc[5] = fma(gy, dy, acc[5]); acc[6] = fma(gy, dz, acc[6]); acc[7] = fma(gz, dz, acc[7]);
                if (fij) {                                                       // source/sepcoulomb.c:68-82
                    const int mj = tag_mol(pj.w);
                    if (mi != -1 && mj != -1) {
                        double *t = fij + ((size_t)mi * nmol + mj) * 3;
                        atomicAdd(t, gx); atomicAdd(t + 1, gy); atomicAdd(t + 2, gz);
                    }
                }
            }
        }
    }
    if (valid) {
        if (STORE) { d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0; f4[i] = o; }
        else { d4 o = f4[i]; o.x += fx; o.y += fy; o.z += fz; f4[i] = o; }
    }
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

template <int TPA>
static void launch_coulomb(sepgpu_ctx *c, int grid, bool store, double cf, const BoxC &B)
{
    if (store) k_coulomb_list<TPA, true><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->zs, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, cf, B, c->partial, c->fij, c->nmol);
    else       k_coulomb_list<TPA, false><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->zs, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, cf, B, c->partial, c->fij, c->nmol);
}

extern "C" int sepgpu_coulomb_sf(sepgpu_ctx *c, const sepgpu_sys *sys, double cf, unsigned opt)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    BoxC B; B.Lx = sys->length[0]; B.Ly = sys->length[1]; B.Lz = sys->length[2];
    const bool store = c->f_zero;
    if (sys->neighb_update == 0) {
        if (opt == SEPGPU_EXCL_BONDED && !c->have_excl) {
            sepgpu_set_error("coulomb_sf: SEP_EXCL_BONDED needs the partner tables");
            return SEPGPU_ESTATE;
        }
        const int grid = (c->n + FORCE_BLOCK - 1) / FORCE_BLOCK;
        ktimer_begin(c, &c->t_coul);
        if (store) k_coulomb_brute<true><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->x4, c->z, c->f4, c->n, cf, B, opt, c->excl_bond, c->excl_angle, c->excl_dihed, c->partial, c->fij, c->nmol);
        else       k_coulomb_brute<false><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->x4, c->z, c->f4, c->n, cf, B, opt, c->excl_bond, c->excl_angle, c->excl_dihed, c->partial, c->fij, c->nmol);
        ktimer_end(c, &c->t_coul);
        KERNEL_CHECK();
        c->f_zero = false;
        return sepgpu_finalize_force(c, grid, 0.5, 4);
    }
    // list mode: the reference neither builds nor checks the list here and ignores opt
    // (source/sepcoulomb.c:8-16); it reuses whatever the preceding sep_force_pairs left behind.
    if (!c->list_valid) {
        // The reference walks whatever list the last sep_force_pairs left behind, however stale (a program may call
        // sep_coulomb_sf before sep_force_pairs in a step).  Here the skin trigger has already invalidated that list: build
        // a fresh one with the exclusion rule of the last build -- a superset of what the stale list still guarantees.
        if (c->list_gen == 0) {
            sepgpu_set_error("coulomb_sf: no neighbour list (call sep_force_pairs first, as the reference requires)");
            return SEPGPU_ESTATE;
        }
        int rcb = sepgpu_neighb_build(c, sys, c->list_opt);
        if (rcb) return rcb;
    }
    if (c->list_f16) {                      // the list kernels below walk global-index rows
        int rcb = sepgpu_need_global_rows(c, sys);
        if (rcb) return rcb;
    }
    if (c->dd) {
        // decomposed run: per-row charges follow the rebuild (own and halo rows), halo coordinates of this step into xs
        if (!sepgpu_dd_refresh_charges(c)) {
            sepgpu_set_error("coulomb_sf: decomposed runs take the charges of all atoms by global id (sepgpu_dd_set_charges)");
            return SEPGPU_ESTATE;
        }
        int rch = sepgpu_dd_halo_update(c, sys);
        if (rch) return rch;
    }
    if ((c->coulomb_kernel == 2 || c->dd) && !c->fij) {
        if (!c->xq) CUDA_TRY(cudaMalloc((void **)&c->xq, sizeof(d4) * (size_t)c->ncap));
        CoulDev P; P.cf2 = cf * cf; P.icf2 = 1.0 / P.cf2; P.twoicf = 2.0 / cf;
        // contiguous ranges of the sorted atoms per CTA, as in the Lennard-Jones list kernel
        int grid = 148 * 64;
        int apc = (c->n + grid - 1) / grid;
        apc = ((apc + FORCE_BLOCK - 1) / FORCE_BLOCK) * FORCE_BLOCK;
        grid = (c->n + apc - 1) / apc;
        ktimer_begin(c, &c->t_coul);
        k_make_xq<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->xs, c->z, c->order, c->xq, c->n);
#define C2_LAUNCH(MB)                                                                                                                                            \
        do {                                                                                                                                                     \
            if (store) k_coulomb_list2<true, MB><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xq, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, apc, P, B, c->partial);  \
            else       k_coulomb_list2<false, MB><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xq, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, apc, P, B, c->partial); \
        } while (0)
        C2_LAUNCH(4);                       // measured on B200: 4 CTAs/SM (no spills) 1.51 ms, 5: 1.54, 6: 1.55 (water, 1.12 M atoms)
#undef C2_LAUNCH
        ktimer_end(c, &c->t_coul);
        KERNEL_CHECK();
        c->f_zero = false;
        return sepgpu_finalize_force(c, grid, 0.5, 4);
    }
    if (!c->zs) CUDA_TRY(cudaMalloc((void **)&c->zs, sizeof(double) * (size_t)c->n));
    k_sort_charges<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->z, c->order, c->zs, c->n);
    const int tpa = c->tpa;
    const long long gpb = FORCE_BLOCK / tpa;
    long long want = ((long long)c->n + gpb - 1) / gpb;
    const int grid = (int)(want < FORCE_MAX_GRID ? want : FORCE_MAX_GRID);
    ktimer_begin(c, &c->t_coul);
    switch (tpa) {
    case 1: launch_coulomb<1>(c, grid, store, cf, B); break;
    case 2: launch_coulomb<2>(c, grid, store, cf, B); break;
    case 4: launch_coulomb<4>(c, grid, store, cf, B); break;
    case 8: launch_coulomb<8>(c, grid, store, cf, B); break;
    case 16: launch_coulomb<16>(c, grid, store, cf, B); break;
    default: launch_coulomb<32>(c, grid, store, cf, B); break;
    }
    ktimer_end(c, &c->t_coul);
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, grid, 0.5, 4);
}

// ---- DPD ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// pair-symmetric counter-based uniform in [0,1): both ends of a pair draw the same number, so the
// random force obeys Newton's third law without communication.  Replaces the glibc rand() stream of
// source/sepprfrc.c:1069, which no parallel evaluation order can reproduce.
__device__ __forceinline__ double dpd_uniform(unsigned long long seed, unsigned long long step, unsigned i, unsigned j)
{
    const unsigned lo = i < j ? i : j, hi = i < j ? j : i;
    if (seed == SEPGPU_DPD_SEED_FIXED) return 0.75;      // parity fixture: the reference with rand() interposed to a constant
    unsigned long long h = mix64(seed ^ (step * 0xD1342543DE82EF95ULL));
    h = mix64(h ^ (((unsigned long long)lo << 32) | hi));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

struct DpdParams { double cf2, aij, gamma, sigma, isqrtdt, facchk; int t0, t1; unsigned long long seed, step; };

// MODE 0: Verlet list (sorted indices, image codes); MODE 1: all pairs on x4
template <int MODE, bool STORE>
__global__ void __launch_bounds__(FORCE_BLOCK)
k_dpd(const d4 *__restrict__ xs, const d4 *__restrict__ x4, const d4 *__restrict__ pv4,
      const unsigned *__restrict__ nbr, const int *__restrict__ cnt, const int *__restrict__ order,
      d4 *__restrict__ f4, int n, int npad, DpdParams P, BoxC B, double *__restrict__ partial)
{
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    const int s = blockIdx.x * FORCE_BLOCK + threadIdx.x;
    if (s < n) {
        const int i = MODE == 0 ? order[s] : s;
        const d4 pi = MODE == 0 ? xs[s] : x4[s];
        const d4 vi = pv4[i];
        const int ti = tag_type(pi.w);
        const int m = MODE == 0 ? cnt[s] : n;
        double fx = 0, fy = 0, fz = 0;
        if (ti == P.t0 || ti == P.t1) {
            for (int k = 0; k < m; k++) {
                int j, jo; d4 pj; double dx, dy, dz;
                if (MODE == 0) {
                    const unsigned e = nbr[nbr_index(k, s, npad)];
                    j = (int)(e & SEPGPU_INDEX_MASK); jo = order[j]; pj = xs[j];
                    dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
                    const int code = (int)(e >> SEPGPU_SHIFT_BITS);
                    if (code != 13) apply_image_c(code, B, dx, dy, dz);
                } else {
                    j = k; jo = k; if (j == s) continue; pj = x4[j];
                    dx = wrap_exact(pi.x - pj.x, B.Lx, 0.5 * B.Lx);
                    dy = wrap_exact(pi.y - pj.y, B.Ly, 0.5 * B.Ly);
                    dz = wrap_exact(pi.z - pj.z, B.Lz, 0.5 * B.Lz);
                }
                const int tj = tag_type(pj.w);
                if (!((ti == P.t0 && tj == P.t1) || (ti == P.t1 && tj == P.t0))) continue;
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (!(r2 < P.cf2)) continue;
                const double dij = sqrt(r2), w = 1.0 - dij;                     // source/sepprfrc.c:1058-1059
                const double rx = dx / dij, ry = dy / dij, rz = dz / dij;
                const d4 vj = pv4[jo];
                const double dotrv = rx * (vi.x - vj.x) + ry * (vi.y - vj.y) + rz * (vi.z - vj.z);
                const double xi = (dpd_uniform(P.seed, P.step, (unsigned)i, (unsigned)jo) - 0.5) * P.facchk;
                const double mag = P.aij * w - P.gamma * w * w * dotrv + P.sigma * w * P.isqrtdt * xi;  // fC+fD+fR along rhat
                fx += mag * rx; fy += mag * ry; fz += mag * rz;
                acc[0] += 0.5 * P.aij * w * w;                                   // :1084
                if (MODE == 1) {                                                 // brute path keeps the conservative virial (:1199-1201)
                    const double cx = P.aij * w * rx, cy = P.aij * w * ry, cz = P.aij * w * rz;
                    acc[2] += cx * dx; acc[3] += cx * dy; acc[4] += cx * dz;
                    acc[5] += cy * dy; acc[6] += cy * dz; acc[7] += cz * dz;
                }
            }
        }
        if (STORE) { d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0; f4[i] = o; }
        else { d4 o = f4[i]; o.x += fx; o.y += fy; o.z += fz; f4[i] = o; }
    }
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

extern "C" int sepgpu_force_dpd(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], double cf,
                                double aij, double temp, double sigma, unsigned opt,
                                unsigned long long seed, unsigned long long step)
{
    if (!c || !sys || !types) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    int rc = sepgpu_ensure_dpd(c);
    if (rc) return rc;
    DpdParams P;
    P.cf2 = cf * cf; P.aij = aij; P.gamma = sigma * sigma / (2.0 * temp); P.sigma = sigma;
    P.isqrtdt = 1.0 / sqrt(sys->dt); P.facchk = 2.0 * sqrt(3.0);
    P.t0 = (unsigned char)types[0]; P.t1 = (unsigned char)types[1];
    P.seed = seed; P.step = step;
    BoxC B; B.Lx = sys->length[0]; B.Ly = sys->length[1]; B.Lz = sys->length[2];
    const bool store = c->f_zero;
    const int grid = (c->n + FORCE_BLOCK - 1) / FORCE_BLOCK;
    if (grid > SEPGPU_MAX_BLOCKS_PARTIAL) { sepgpu_set_error("force_dpd: system too large"); return SEPGPU_EINVAL; }
    if (sys->neighb_update == 0) {
        if (store) k_dpd<1, true><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->x4, c->pv4, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, P, B, c->partial);
        else       k_dpd<1, false><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->x4, c->pv4, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, P, B, c->partial);
        KERNEL_CHECK();
        c->f_zero = false;
        return sepgpu_finalize_force(c, grid, 0.5, 0);        // brute: epot += (source/sepprfrc.c:1196)
    }
    c->need_atom_rows = true;                                                    // DPD walks global-index rows
    if (!c->list_valid && (rc = sepgpu_neighb_build(c, sys, opt))) return rc;    // :1021-1031
    if (c->list_f16 && (rc = sepgpu_need_global_rows(c, sys))) return rc;
    if (store) k_dpd<0, true><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->x4, c->pv4, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, P, B, c->partial);
    else       k_dpd<0, false><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->xs, c->x4, c->pv4, c->nbr, c->cnt, c->order, c->f4, c->n, c->npad, P, B, c->partial);
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, grid, 0.5, 1);            // list: epot assigned (:1132), no virial (:1089-1116)
}
