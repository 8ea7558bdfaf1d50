// sepgpu_intgr.cu -- fused thermostat + integrator kernels.
//
// Stand-ins for sep_nosehoover / _sep_nosehoover_type (reference source/sepintgr.c:149-198),
// sep_leapfrog + sep_periodic + the skin trigger + sep_set_xn (source/sepintgr.c:18-88,
// source/sepmisc.c:525-533), sep_verlet_dpd (source/sepintgr.c:296-345), sep_reset_momentum
// (source/sepmisc.c:1173-1192) and the coordinate rescale of sep_compress_box (:1009-1010).
//
// The reference makes two passes for the thermostat (sum m v^2, then f -= alpha m v) and a third for
// leapfrog.  Here the integrator kernel of step n emits sum m v^2 for the thermostat of step n+1, the
// multiplier update is a one-thread kernel, and f -= alpha m v is applied inside the next integrator
// kernel -- one streaming pass over the atoms per time step (HBM-bound: 144 B read + 128 B written).
#include "sepgpu_internal.cuh"
#include "sepgpu_intgr_atom.cuh"

#define INTGR_BLOCK 256
#define INTGR_MAX_GRID (148 * 8)

struct IntgrParams {
    double Lx, Ly, Lz;
    double dt, skin;
    int n;
    int alpha_slot;     // -1: no pending thermostat
    int alpha_type;     // -1: all atoms
    int f_zero;         // force array logically zero
    int write_xs;       // maintain the cell-sorted copy
};

// sep_nosehoover's multiplier update (source/sepintgr.c:157-161) as one function with pinned rounding: with option
// step_fold every thread of k_integrate<.., true> evaluates it on the same inputs and k_finalize_both stores the same value
struct NhFold { double temp0, tau, npart; };
__device__ __forceinline__ double nh_alpha_next(double alpha, double sum_mv2, const NhFold &N, double dt)
{
    const double ekin = __ddiv_rn(__dmul_rn(0.5, sum_mv2), N.npart);
    const double temp = __dmul_rn(0.666667, ekin);
    const double rate = __ddiv_rn(dt, __dmul_rn(N.tau, N.tau));
    return __dadd_rn(alpha, __dmul_rn(rate, __dsub_rn(__ddiv_rn(temp, N.temp0), 1.0)));
}

template <bool DPD, bool NHFOLD>
__global__ void __launch_bounds__(INTGR_BLOCK)
k_integrate(d4 *__restrict__ x4, d4 *__restrict__ v4, d4 *__restrict__ f4, const d4 *__restrict__ xn4,
            i4 *__restrict__ cr4, int *__restrict__ crossings, const int *__restrict__ rank,
            d4 *__restrict__ xs, d4 *__restrict__ pv4, d4 *__restrict__ pa4, const DevScalars *__restrict__ scal,
            IntgrParams P, double lambda, int stepnow, double *__restrict__ partial, NhFold N)
{
    __shared__ double red[SEPGPU_NPART_I * (INTGR_BLOCK / 32)];
    double acc[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) acc[q] = 0.0;
    double alpha = P.alpha_slot >= 0 ? scal->alpha[P.alpha_slot] : 0.0;
    if (NHFOLD) alpha = nh_alpha_next(alpha, scal->sum_mv2, N, P.dt);
    const double dt = P.dt;

    for (int i = blockIdx.x * INTGR_BLOCK + threadIdx.x; i < P.n; i += gridDim.x * INTGR_BLOCK) {
        d4 x = x4[i], v = v4[i], f;
        if (P.f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
        const double m = v.w;
        bool f_changed = P.f_zero != 0;
        if (P.alpha_slot >= 0 && (P.alpha_type < 0 || tag_type(x.w) == P.alpha_type)) {
            // sep_nosehoover second pass: f[k] -= alpha*m*v[k]   (source/sepintgr.c:163-166)
            // (_sep_nosehoover_type multiplies alpha*v*m, :193 -- same value up to rounding order)
            const double am = alpha * m;
            f.x -= am * v.x; f.y -= am * v.y; f.z -= am * v.z;
            f_changed = true;
        }
        const double ax = f.x / m, ay = f.y / m, az = f.z / m;       // :53
        double ux, uy, uz;                                            // velocity entering ekin / kin_P
        if (!DPD) {
            v.x += ax * dt; v.y += ay * dt; v.z += az * dt;           // :54
            x.x += v.x * dt; x.y += v.y * dt; x.z += v.z * dt;        // :55
            ux = v.x - 0.5 * ax * dt; uy = v.y - 0.5 * ay * dt; uz = v.z - 0.5 * az * dt;   // :57
        } else {
            d4 pa = pa4[i];
            if (stepnow > 0) {                                         // source/sepintgr.c:311-312
                v.x += 0.5 * dt * (ax + pa.x); v.y += 0.5 * dt * (ay + pa.y); v.z += 0.5 * dt * (az + pa.z);
            }
            x.x += v.x * dt + 0.5 * dt * dt * ax;                      // :314
            x.y += v.y * dt + 0.5 * dt * dt * ay;
            x.z += v.z * dt + 0.5 * dt * dt * az;
            d4 pv; pv.x = v.x + lambda * dt * ax; pv.y = v.y + lambda * dt * ay; pv.z = v.z + lambda * dt * az; pv.w = 0;
            pa.x = ax; pa.y = ay; pa.z = az;
            pv4[i] = pv; pa4[i] = pa;
            ux = v.x; uy = v.y; uz = v.z;                              // :321
        }
        acc[0] += ux * ux * m; acc[0] += uy * uy * m; acc[0] += uz * uz * m;      // :58
        acc[1] += ux * ux * m; acc[2] += ux * uy * m; acc[3] += ux * uz * m;      // :66-68 (symmetric)
        acc[4] += uy * uy * m; acc[5] += uy * uz * m; acc[6] += uz * uz * m;
        acc[8] += (v.x * v.x + v.y * v.y + v.z * v.z) * m;           // for the next sep_nosehoover
        acc[9] += v.x * m; acc[10] += v.y * m; acc[11] += v.z * m;

        i4 cr = cr4[i];
        const d4 xn = xn4[i];
        int clx, cly, clz; unpack_cl(cr.w, clx, cly, clz);
        int tx = 0, ty = 0, tz = 0; bool changed = false;
        double d2 = 0.0;
        d2 += periodic_1d(x.x, P.Lx, cr.x, clx, tx, changed, xn.x);
        d2 += periodic_1d(x.y, P.Ly, cr.y, cly, ty, changed, xn.y);
        d2 += periodic_1d(x.z, P.Lz, cr.z, clz, tz, changed, xn.z);
        acc[7] = fmax(acc[7], d2);                                    // :62

        x4[i] = x; v4[i] = v;
        if (f_changed) f4[i] = f;
        if (changed) {
            cr.w = pack_cl(clx, cly, clz);
            cr4[i] = cr;
            if (tx) crossings[3 * i] += tx;
            if (ty) crossings[3 * i + 1] += ty;
            if (tz) crossings[3 * i + 2] += tz;
        }
        if (P.write_xs) {
            d4 u; u.x = x.x + clx * P.Lx; u.y = x.y + cly * P.Ly; u.z = x.z + clz * P.Lz; u.w = x.w;
            xs[rank[i]] = u;
        }
    }
    // block reduction: sums for all but slot 7 (max)
    const double mymax = acc[7];
    acc[7] = 0.0;
    block_sum<SEPGPU_NPART_I, INTGR_BLOCK>(acc, red);
    __syncthreads();
    double wm = warp_max(mymax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int w = 0; w < INTGR_BLOCK / 32; w++) mx = fmax(mx, red[w]);
        acc[7] = mx;
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++) partial[blockIdx.x * SEPGPU_NPART_I + q] = acc[q];
    }
}

// one block: reduce the partial rows, update sepret/sepsys scalars, evaluate the skin trigger
__global__ void __launch_bounds__(256)
k_finalize_intgr(const double *__restrict__ partial, int nrows, DevScalars *scal, double skin, int mode, double *comm, int resets,
                 int rank, int nranks)
{
    // resets: bit0 a sep_reset_retval is pending, bit1 a sep_reset_force (max_dist2 <- 0) is pending
    if (mode != 1 && threadIdx.x == 0) {
        if (resets & 1) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
        if (resets & 2) scal->max_dist2 = 0.0;
    }
    // mode 0: single GPU.  mode 1: decomposed run, phase A -- local sums into comm[0..11], local max into
    // comm[12 + rank] (the other ranks' slots zero), so that ONE sum all-reduce over 12 + nranks doubles also
    // delivers every rank's maximum.  mode 2: phase B -- apply comm.
    if (mode == 2) {
        if (threadIdx.x == 0) {
            scal->ekin += 0.5 * comm[0];
            const double K[9] = {comm[1], comm[2], comm[3], comm[2], comm[4], comm[5], comm[3], comm[5], comm[6]};
            for (int k = 0; k < 9; k++) scal->kin_P[k] += K[k];
            double gmx = 0.0;
            for (int r = 0; r < nranks; r++) gmx = fmax(gmx, comm[12 + r]);
            if (gmx > scal->max_dist2) scal->max_dist2 = gmx;
            scal->sum_mv2 = comm[8];
            scal->mom[0] = comm[9]; scal->mom[1] = comm[10]; scal->mom[2] = comm[11];
            scal->neighb_flag = sqrt(scal->max_dist2) > skin * 0.5 ? 1 : 0;
        }
        return;
    }
    __shared__ double red[SEPGPU_NPART_I * 8];
    double v[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) v[q] = 0.0;
    double mx = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 256) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++)
            if (q != 7) v[q] += partial[r * SEPGPU_NPART_I + q];
        mx = fmax(mx, partial[r * SEPGPU_NPART_I + 7]);
    }
    block_sum<SEPGPU_NPART_I, 256>(v, red);
    __syncthreads();
    double wm = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; w++) mx = fmax(mx, red[w]);
        if (mode == 1) {
            for (int q = 0; q < SEPGPU_NPART_I; q++) comm[q] = v[q];
            comm[7] = 0.0;
            for (int r = 0; r < nranks; r++) comm[12 + r] = r == rank ? mx : 0.0;
            return;
        }
        scal->ekin += 0.5 * v[0];                                     // source/sepintgr.c:87
        const double K[9] = {v[1], v[2], v[3], v[2], v[4], v[5], v[3], v[5], v[6]};
        for (int k = 0; k < 9; k++) scal->kin_P[k] += K[k];
        if (mx > scal->max_dist2) scal->max_dist2 = mx;               // :62
        scal->sum_mv2 = v[8];
        scal->mom[0] = v[9]; scal->mom[1] = v[10]; scal->mom[2] = v[11];
        scal->neighb_flag = sqrt(scal->max_dist2) > skin * 0.5 ? 1 : 0;   // :72
    }
}

// Option step_fold, single GPU: the step's last force reduction (k_finalize_force), the Nose-Hoover multiplier update
// (k_nh_update) and the integrator's own reduction (k_finalize_intgr, mode 0) in ONE kernel, applied in that order.
// fflags < 0: no force reduction pending; nh_slot < 0: no multiplier update pending.
#define FIN_BOTH_THREADS 1024
__global__ void __launch_bounds__(FIN_BOTH_THREADS)
k_finalize_both(const double *__restrict__ fpartial, int fnrows, double fscale, int fflags,
                const double *__restrict__ ipartial, int inrows, DevScalars *scal, double skin, int resets,
                int nh_slot, NhFold N, double dt)
{
    __shared__ double red[SEPGPU_NPART_I * (FIN_BOTH_THREADS / 32)];
    // ---- force rows (source/sepprfrc.c:222 and friends; see k_finalize_force) ----
    double f[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] = 0.0;
    if (fflags >= 0) {
        for (int r = threadIdx.x; r < fnrows; r += FIN_BOTH_THREADS) {
#pragma unroll
            for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] += fpartial[(size_t)r * SEPGPU_NPART_F + q];
        }
        block_sum<SEPGPU_NPART_F, FIN_BOTH_THREADS>(f, red);
        __syncthreads();
    }
    // ---- integrator rows ----
    double v[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) v[q] = 0.0;
    double mx = 0.0;
    for (int r = threadIdx.x; r < inrows; r += FIN_BOTH_THREADS) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++)
            if (q != 7) v[q] += ipartial[r * SEPGPU_NPART_I + q];
        mx = fmax(mx, ipartial[r * SEPGPU_NPART_I + 7]);
    }
    block_sum<SEPGPU_NPART_I, FIN_BOTH_THREADS>(v, red);
    __syncthreads();
    const double wm = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < FIN_BOTH_THREADS / 32; w++) mx = fmax(mx, red[w]);
        if (fflags >= 0) {                                             // what k_finalize_force does
            if (fflags & 8) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
            const double e = f[0] * fscale, ec = f[1] * fscale;
            if (fflags & 1) scal->epot = e; else scal->epot += e;
            if (fflags & 4) { scal->epot += ec; scal->ecoul += ec; }
            const double xx = f[2] * fscale, xy = f[3] * fscale, xz = f[4] * fscale;
            const double yy = f[5] * fscale, yz = f[6] * fscale, zz = f[7] * fscale;
            const double Pm[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
            for (int k = 0; k < 9; k++) { scal->pot_P[k] += Pm[k]; if (fflags & 2) scal->pot_P_bond[k] += Pm[k]; }
        }
        if (nh_slot >= 0)                                              // what k_nh_update (mode 0) does, before sum_mv2 moves on
            scal->alpha[nh_slot] = nh_alpha_next(scal->alpha[nh_slot], scal->sum_mv2, N, dt);
        // what k_finalize_intgr (mode 0) does
        if (resets & 1) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
        if (resets & 2) scal->max_dist2 = 0.0;
        scal->ekin += 0.5 * v[0];
        const double K[9] = {v[1], v[2], v[3], v[2], v[4], v[5], v[3], v[5], v[6]};
        for (int k = 0; k < 9; k++) scal->kin_P[k] += K[k];
        if (mx > scal->max_dist2) scal->max_dist2 = mx;
        scal->sum_mv2 = v[8];
        scal->mom[0] = v[9]; scal->mom[1] = v[10]; scal->mom[2] = v[11];
        scal->neighb_flag = sqrt(scal->max_dist2) > skin * 0.5 ? 1 : 0;
    }
}

// k_finalize_both on several CTAs (options step_fold + fin_multi): CTA b sums fixed chunks of the force rows and of the
// integrator rows into stage row b (8 + 12 doubles), the CTA that draws the last ticket adds the stage rows in index
// order and applies them exactly as k_finalize_both does.  Fixed chunks, fixed final order: deterministic.
#define FBM_THREADS 256
#define FBM_MAX_CTAS 64
#define FBM_W (SEPGPU_NPART_F + SEPGPU_NPART_I)
__global__ void __launch_bounds__(FBM_THREADS)
k_finalize_both_multi(const double *__restrict__ fpartial, int fnrows, double fscale, int fflags,
                      const double *__restrict__ ipartial, int inrows, double *__restrict__ stage, unsigned *ticket,
                      DevScalars *scal, double skin, int resets, int nh_slot, NhFold N, double dt)
{
    __shared__ double red[SEPGPU_NPART_I * (FBM_THREADS / 32)];
    __shared__ double tot[FBM_W];
    __shared__ int s_last;
    const int G = gridDim.x;
    double f[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] = 0.0;
    if (fflags >= 0) {
        const int chunk = (fnrows + G - 1) / G, r0 = blockIdx.x * chunk, r1 = min(fnrows, r0 + chunk);
        for (int r = r0 + (int)threadIdx.x; r < r1; r += FBM_THREADS) {
#pragma unroll
            for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] += fpartial[(size_t)r * SEPGPU_NPART_F + q];
        }
        block_sum<SEPGPU_NPART_F, FBM_THREADS>(f, red);
        __syncthreads();
    }
    double v[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) v[q] = 0.0;
    double mx = 0.0;
    {
        const int chunk = (inrows + G - 1) / G, r0 = blockIdx.x * chunk, r1 = min(inrows, r0 + chunk);
        for (int r = r0 + (int)threadIdx.x; r < r1; r += FBM_THREADS) {
#pragma unroll
            for (int q = 0; q < SEPGPU_NPART_I; q++)
                if (q != 7) v[q] += ipartial[r * SEPGPU_NPART_I + q];
            mx = fmax(mx, ipartial[r * SEPGPU_NPART_I + 7]);
        }
    }
    block_sum<SEPGPU_NPART_I, FBM_THREADS>(v, red);
    __syncthreads();
    const double wm = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < FBM_THREADS / 32; w++) mx = fmax(mx, red[w]);
        v[7] = mx;
        double *row = stage + (size_t)blockIdx.x * FBM_W;
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) row[q] = f[q];
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++) row[SEPGPU_NPART_F + q] = v[q];
        __threadfence();                                            // stage row visible before the ticket is drawn
        s_last = atomicAdd(ticket, 1u) == (unsigned)G - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < FBM_W) {
        const int q = threadIdx.x;
        double a = 0.0;
        for (int b = 0; b < G; b++) {
            const double x = __ldcg(stage + (size_t)b * FBM_W + q);
            if (q == SEPGPU_NPART_F + 7) a = fmax(a, x); else a += x;
        }
        tot[q] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0;
        const double *F = tot, *V = tot + SEPGPU_NPART_F;
        if (fflags >= 0) {
            if (fflags & 8) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
            const double e = F[0] * fscale, ec = F[1] * fscale;
            if (fflags & 1) scal->epot = e; else scal->epot += e;
            if (fflags & 4) { scal->epot += ec; scal->ecoul += ec; }
            const double xx = F[2] * fscale, xy = F[3] * fscale, xz = F[4] * fscale;
            const double yy = F[5] * fscale, yz = F[6] * fscale, zz = F[7] * fscale;
            const double Pm[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
            for (int k = 0; k < 9; k++) { scal->pot_P[k] += Pm[k]; if (fflags & 2) scal->pot_P_bond[k] += Pm[k]; }
        }
        if (nh_slot >= 0) scal->alpha[nh_slot] = nh_alpha_next(scal->alpha[nh_slot], scal->sum_mv2, N, dt);
        if (resets & 1) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
        if (resets & 2) scal->max_dist2 = 0.0;
        scal->ekin += 0.5 * V[0];
        const double K[9] = {V[1], V[2], V[3], V[2], V[4], V[5], V[3], V[5], V[6]};
        for (int k = 0; k < 9; k++) scal->kin_P[k] += K[k];
        if (V[7] > scal->max_dist2) scal->max_dist2 = V[7];
        scal->sum_mv2 = V[8];
        scal->mom[0] = V[9]; scal->mom[1] = V[10]; scal->mom[2] = V[11];
        scal->neighb_flag = sqrt(scal->max_dist2) > skin * 0.5 ? 1 : 0;
    }
}

// Decomposed run with peer memory: ONE kernel reduces this rank's partial rows, stores the 12 sums into every
// rank's gather table over NVLink (slot = seq & 1, row = my rank), raises the per-sender flag, waits for all
// senders' flags and then adds the rows in rank order -- every rank computes bit-identical totals and so takes
// the same rebuild decision.  Replaces finalize(phase A) + NCCL all-reduce + finalize(phase B).
// Slot reuse is safe: nobody can be two refreshes ahead of a rank whose contribution it still needs.
// FOLD (option step_fold): the step's last force reduction and the Nose-Hoover multiplier update ride along, applied in
// the order the separate kernels would (force sums, multiplier, then the integrator's sums).
struct FoldArgs { const double *fpartial; int fnrows; double fscale; int fflags; int nh_slot; NhFold N; double dt; };

template <bool FOLD>
__global__ void __launch_bounds__(256)
k_finalize_intgr_p2p(const double *__restrict__ partial, int nrows, DevScalars *scal, double skin, int resets, GatherDev G, FoldArgs F)
{
    __shared__ double red[SEPGPU_NPART_I * 8];
    __shared__ double mine[SEPGPU_NPART_I];
    if (FOLD) {
        if (F.fflags >= 0) {
            double f[SEPGPU_NPART_F];
#pragma unroll
            for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] = 0.0;
            for (int r = threadIdx.x; r < F.fnrows; r += 256) {
#pragma unroll
                for (int q = 0; q < SEPGPU_NPART_F; q++) f[q] += F.fpartial[(size_t)r * SEPGPU_NPART_F + q];
            }
            block_sum<SEPGPU_NPART_F, 256>(f, red);
            if (threadIdx.x == 0) {
                if (F.fflags & 8) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
                const double e = f[0] * F.fscale, ec = f[1] * F.fscale;
                if (F.fflags & 1) scal->epot = e; else scal->epot += e;
                if (F.fflags & 4) { scal->epot += ec; scal->ecoul += ec; }
                const double xx = f[2] * F.fscale, xy = f[3] * F.fscale, xz = f[4] * F.fscale;
                const double yy = f[5] * F.fscale, yz = f[6] * F.fscale, zz = f[7] * F.fscale;
                const double Pm[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
                for (int k = 0; k < 9; k++) { scal->pot_P[k] += Pm[k]; if (F.fflags & 2) scal->pot_P_bond[k] += Pm[k]; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0 && F.nh_slot >= 0)            // sum_mv2 is still the previous step's global value here
            scal->alpha[F.nh_slot] = nh_alpha_next(scal->alpha[F.nh_slot], scal->sum_mv2, F.N, F.dt);
    }
    if (threadIdx.x == 0) {
        if (resets & 1) { scal->epot = 0; scal->ecoul = 0; scal->ekin = 0; for (int k = 0; k < 9; k++) { scal->pot_P[k] = 0; scal->kin_P[k] = 0; scal->pot_P_bond[k] = 0; } }
        if (resets & 2) scal->max_dist2 = 0.0;
    }
    double v[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) v[q] = 0.0;
    double mx = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 256) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++)
            if (q != 7) v[q] += partial[r * SEPGPU_NPART_I + q];
        mx = fmax(mx, partial[r * SEPGPU_NPART_I + 7]);
    }
    block_sum<SEPGPU_NPART_I, 256>(v, red);
    __syncthreads();
    double wm = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; w++) mx = fmax(mx, red[w]);
        for (int q = 0; q < SEPGPU_NPART_I; q++) mine[q] = v[q];
        mine[7] = mx;
    }
    __syncthreads();
    const int slot = (int)(G.seq & 1);
    const int t = threadIdx.x;
    if (t < G.nranks) {                                   // thread t serves peer t
        unsigned char *base = G.bases[t];
        double *row = reinterpret_cast<double *>(base + G.gather_off) + ((size_t)slot * G.nranks + G.rank) * SEPGPU_GATHER_W;
        for (int q = 0; q < SEPGPU_NPART_I; q++) row[q] = mine[q];
        __threadfence_system();
        unsigned long long *fl = reinterpret_cast<unsigned long long *>(base + G.gflag_off) + (size_t)slot * G.nranks + G.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(G.seq) : "memory");
        // ... and waits for sender t
        const unsigned long long *my = reinterpret_cast<const unsigned long long *>(G.bases[G.rank] + G.gflag_off) + (size_t)slot * G.nranks + t;
        const long long t0 = clock64();
        unsigned long long got;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(my) : "memory");
            if (got >= G.seq) break;
            if (clock64() - t0 > SEPGPU_SPIN_LIMIT) { scal->error = SEPGPU_ENCCL; scal->error_where = 5; break; }
        } while (true);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *tab = reinterpret_cast<const double *>(G.bases[G.rank] + G.gather_off) + (size_t)slot * G.nranks * SEPGPU_GATHER_W;
        double s[SEPGPU_NPART_I];
        for (int q = 0; q < SEPGPU_NPART_I; q++) s[q] = 0.0;
        double gmx = 0.0;
        for (int r = 0; r < G.nranks; r++) {
            for (int q = 0; q < SEPGPU_NPART_I; q++) {
                const double x = __ldcg(tab + (size_t)r * SEPGPU_GATHER_W + q);
                if (q == 7) gmx = fmax(gmx, x); else s[q] += x;
            }
        }
        scal->ekin += 0.5 * s[0];
        const double K[9] = {s[1], s[2], s[3], s[2], s[4], s[5], s[3], s[5], s[6]};
        for (int k = 0; k < 9; k++) scal->kin_P[k] += K[k];
        if (gmx > scal->max_dist2) scal->max_dist2 = gmx;
        scal->sum_mv2 = s[8];
        scal->mom[0] = s[9]; scal->mom[1] = s[10]; scal->mom[2] = s[11];
        scal->neighb_flag = sqrt(scal->max_dist2) > skin * 0.5 ? 1 : 0;
    }
}

// xn <- x, cross_neighb <- 0 (source/sepintgr.c:76-82)
__global__ void k_set_xn(const d4 *__restrict__ x4, d4 *__restrict__ xn4, i4 *__restrict__ cr4, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 x = x4[i]; x.w = 0.0;
    xn4[i] = x;
    i4 c = cr4[i]; c.x = c.y = c.z = 0;
    cr4[i] = c;
}

int sepgpu_reset_xn(sepgpu_ctx *c)
{
    k_set_xn<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->x4, c->xn4, c->cr4, c->n_own);
    KERNEL_CHECK();
    return 0;
}

int sepgpu_ensure_dpd(sepgpu_ctx *c);
// domain-decomposition hooks (sepgpu_dd.cu)
double *sepgpu_dd_comm(sepgpu_ctx *c);
void sepgpu_dd_positions_moved(sepgpu_ctx *c);
int sepgpu_dd_before_positions_change(sepgpu_ctx *c);
int sepgpu_dd_allreduce(sepgpu_ctx *c, double *sum_buf, int nsum, double *max_buf, int nmax);
void sepgpu_dd_rank(sepgpu_ctx *c, int *rank, int *nranks);
bool sepgpu_dd_gather_next(sepgpu_ctx *c, GatherDev *g);
int sepgpu_dd_uses_p2p(sepgpu_ctx *c);

__global__ void k_partial2_to_comm(const double *__restrict__ partial, int nrows, double *comm)
{
    __shared__ double red[2 * 8];
    double v[2] = {0.0, 0.0};
    for (int r = threadIdx.x; r < nrows; r += 256) { v[0] += partial[2 * r]; v[1] += partial[2 * r + 1]; }
    block_sum<2, 256>(v, red);
    if (threadIdx.x == 0) { comm[0] = v[0]; comm[1] = v[1]; }
}
__global__ void k_comm_to_mv2(DevScalars *scal, const double *comm) { scal->sum_mv2 = comm[0]; }

int sepgpu_spec_force_launch(sepgpu_ctx *c);

// reads the scalar block (the rebuild trigger) after the finaliser that has just been queued; returns < 0 on error
static int sepgpu_spec_force_launch_after(sepgpu_ctx *c)
{
    if (c->spec.on && c->spec.streak >= 3 && (!c->dd || c->spec.on == 2)) {
        if (!c->flag_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&c->flag_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fin, cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventRecord(c->ev_fin, c->stream));
        const int rc = sepgpu_spec_force_launch(c);
        if (rc) return rc < 0 ? rc : -1;
        if (c->spec.launched) {
            CUDA_TRY(cudaStreamWaitEvent(c->flag_stream, c->ev_fin, 0));
            CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->flag_stream));
            CUDA_TRY(cudaStreamSynchronize(c->flag_stream));
            return 0;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

static int run_integrator(sepgpu_ctx *c, const sepgpu_sys *sys, bool dpd, double lambda, int stepnow)
{
    CUDA_TRY(cudaSetDevice(c->device));
    if (dpd) { int rc = sepgpu_ensure_dpd(c); if (rc) return rc; }
    IntgrParams P;
    P.Lx = sys->length[0]; P.Ly = sys->length[1]; P.Lz = sys->length[2];
    P.dt = sys->dt; P.skin = sys->skin; P.n = c->n_own;
    P.alpha_slot = c->pending_alpha_slot;
    P.alpha_type = c->pending_alpha_type;
    P.f_zero = c->f_zero ? 1 : 0;
    P.write_xs = (sys->neighb_update != 0 && c->list_valid) ? 1 : 0;
    long long want = ((long long)c->n_own + INTGR_BLOCK - 1) / INTGR_BLOCK;
    const int grid = (int)(want < INTGR_MAX_GRID ? want : INTGR_MAX_GRID);
    // option step_fold (single GPU, leapfrog): what the force routine and sep_nosehoover left pending is folded in here
    NhFold N; N.temp0 = 1.0; N.tau = 1.0; N.npart = 1.0;
    // decomposed runs fold into the peer-memory finaliser only (the NCCL path keeps its three kernels)
    const bool fold = !dpd && (!c->dd || sepgpu_dd_uses_p2p(c)) && (c->fin_pending.active || c->nh_pending.active);
    if (!fold && (c->fin_pending.active || c->nh_pending.active)) { int rs = sepgpu_settle(c); if (rs) return rs; }
    const bool fold_nh = fold && c->nh_pending.active && c->nh_pending.slot == c->pending_alpha_slot && c->pending_alpha_type < 0;
    if (fold && c->nh_pending.active && !fold_nh) { int rs = sepgpu_nh_update_now(c); if (rs) return rs; }
    if (fold_nh) { N.temp0 = c->nh_pending.temp0; N.tau = c->nh_pending.tau; N.npart = (double)c->n_global; }
    // the integrator's partial rows must not land on force rows that are still waiting for their reduction
    double *ipartial = fold ? c->partial + (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * SEPGPU_NPART_F + 2048 : c->partial;
    if (c->dd) { int rw = sepgpu_dd_before_positions_change(c); if (rw) return rw; }
    ktimer_begin(c, &c->t_intgr);
    if (dpd)
        k_integrate<true, false><<<grid, INTGR_BLOCK, 0, c->stream>>>(c->x4, c->v4, c->f4, c->xn4, c->cr4, c->crossings,
            c->rank, c->xs, c->pv4, c->pa4, c->scal, P, lambda, stepnow, c->partial, N);
    else if (fold_nh)
        k_integrate<false, true><<<grid, INTGR_BLOCK, 0, c->stream>>>(c->x4, c->v4, c->f4, c->xn4, c->cr4, c->crossings,
            c->rank, c->xs, c->pv4, c->pa4, c->scal, P, lambda, stepnow, ipartial, N);
    else
        k_integrate<false, false><<<grid, INTGR_BLOCK, 0, c->stream>>>(c->x4, c->v4, c->f4, c->xn4, c->cr4, c->crossings,
            c->rank, c->xs, c->pv4, c->pa4, c->scal, P, lambda, stepnow, ipartial, N);
    const int resets = (c->ret_reset_pending ? 1 : 0) | (c->maxd_reset_pending ? 2 : 0);
    c->ret_reset_pending = false; c->maxd_reset_pending = false;
    GatherDev gd;
    if (fold && c->dd) {
        if (!sepgpu_dd_gather_next(c, &gd)) { sepgpu_set_error("step_fold: the peer-memory path went away"); return SEPGPU_ESTATE; }
        const bool ff = c->fin_pending.active;
        FoldArgs F;
        F.fpartial = c->partial; F.fnrows = ff ? c->fin_pending.nrows : 0; F.fscale = ff ? c->fin_pending.scale : 0.0;
        F.fflags = ff ? c->fin_pending.flags : -1; F.nh_slot = fold_nh ? c->nh_pending.slot : -1; F.N = N; F.dt = sys->dt;
        k_finalize_intgr_p2p<true><<<1, 256, 0, c->stream>>>(ipartial, grid, c->scal, sys->skin, resets, gd, F);
        c->fin_pending.active = false;
        c->nh_pending.active = false;
    } else if (fold) {
        const bool ff = c->fin_pending.active;
        if (c->fin_multi) {
            if (!c->fin_ticket) {
                CUDA_TRY(cudaMalloc((void **)&c->fin_ticket, sizeof(unsigned)));
                CUDA_TRY(cudaMemsetAsync(c->fin_ticket, 0, sizeof(unsigned), c->stream));
            }
            const int most = ff && c->fin_pending.nrows > grid ? c->fin_pending.nrows : grid;
            int per = (most + FBM_MAX_CTAS - 1) / FBM_MAX_CTAS;
            if (per < 8) per = 8;
            const int ctas = (most + per - 1) / per;
            double *stage = c->partial + (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * SEPGPU_NPART_F;      // 64 x 20 doubles, below ipartial
            k_finalize_both_multi<<<ctas, FBM_THREADS, 0, c->stream>>>(c->partial, ff ? c->fin_pending.nrows : 0, ff ? c->fin_pending.scale : 0.0,
                ff ? c->fin_pending.flags : -1, ipartial, grid, stage, c->fin_ticket, c->scal, sys->skin, resets,
                fold_nh ? c->nh_pending.slot : -1, N, sys->dt);
        } else
        k_finalize_both<<<1, FIN_BOTH_THREADS, 0, c->stream>>>(c->partial, ff ? c->fin_pending.nrows : 0, ff ? c->fin_pending.scale : 0.0,
            ff ? c->fin_pending.flags : -1, ipartial, grid, c->scal, sys->skin, resets, fold_nh ? c->nh_pending.slot : -1, N, sys->dt);
        c->fin_pending.active = false;
        c->nh_pending.active = false;
    } else if (c->dd && sepgpu_dd_gather_next(c, &gd)) {
        FoldArgs F0; F0.fpartial = NULL; F0.fnrows = 0; F0.fscale = 0.0; F0.fflags = -1; F0.nh_slot = -1; F0.N = N; F0.dt = 0.0;
        k_finalize_intgr_p2p<false><<<1, 256, 0, c->stream>>>(c->partial, grid, c->scal, sys->skin, resets, gd, F0);
    } else if (c->dd) {
        double *comm = sepgpu_dd_comm(c);
        int drank = 0, dn = 1;
        sepgpu_dd_rank(c, &drank, &dn);
        k_finalize_intgr<<<1, 256, 0, c->stream>>>(c->partial, grid, c->scal, sys->skin, 1, comm, 0, drank, dn);
        int rcd = sepgpu_dd_allreduce(c, comm, 12 + dn, NULL, 0);
        if (rcd) return rcd;
        k_finalize_intgr<<<1, 32, 0, c->stream>>>(c->partial, grid, c->scal, sys->skin, 2, comm, resets, drank, dn);
    } else {
        k_finalize_intgr<<<1, 256, 0, c->stream>>>(c->partial, grid, c->scal, sys->skin, 0, NULL, resets, 0, 1);
    }
    ktimer_end(c, &c->t_intgr);
    KERNEL_CHECK();
    if (c->pending_alpha_slot >= 0 || c->f_zero) c->f_zero = false;   // f4 now holds the force that was used
    c->pending_alpha_slot = -1;
    c->pending_alpha_type = -1;
    c->mv2_valid = true;
    c->moved_since_build = true;
    if (!P.write_xs) c->xs_current = false;
    sepgpu_dd_positions_moved(c);

    // the trigger is needed by the host before the next force call: small D2H + stream sync per step.  With a force launch
    // sent ahead (sepgpu_spec_force_launch) the copy runs on a stream of its own, right behind the finaliser and beside
    // that launch, so that the host decides while the device already computes.
    {
        int rs = sepgpu_spec_force_launch_after(c);
        if (rs < 0) return rs;
    }
    if (c->dd && c->scal_host->error == SEPGPU_ENCCL) {
        sepgpu_set_error("decomposed step: a neighbour's data never arrived (peer-memory wait %d timed out: 1 migration counts, "
                         "2 migration records, 3 halo unpack, 4 halo wait in the force kernel, 5 all-gather of the step's sums)",
                         c->scal_host->error_where);
        return SEPGPU_ENCCL;
    }
    if (c->scal_host->error == SEPGPU_ETABLE) {
        sepgpu_set_error("a pair came closer than the tabulated pair function reaches (raise its resolution towards r = 0: SEP_TABLE_RMIN)");
        return SEPGPU_ETABLE;
    }
    c->scal_cache_valid = !c->dd;                          // sepgpu_read_scalars right after this call needs no second copy
    c->scal_cache_seq = c->api_seq;
    if (c->scal_host->neighb_flag) {
        c->spec.cancelled = true;                          // (the launch sent ahead has seen the flag on the device and left)
        k_set_xn<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->x4, c->xn4, c->cr4, c->n_own);
        KERNEL_CHECK();
        c->list_valid = false;
    }
    return 0;
}

// ---- stochastic integrators ---------------------------------------------------------
// sep_fp (Brownian dynamics at the Fokker-Planck level, source/sepintgr.c:235-293) and sep_langevinGJF
// (Gronbech-Jensen/Farago Langevin integrator, :89-146).  Both draw one Gaussian number per atom and component from
// the reference's serial generator (sep_randn on glibc rand(), source/sepmisc.c:1131-1160); the HOST layer produces
// that stream in the reference's order and hands it over as four doubles per atom {g0, g1, g2, ldiff}, so the
// device arithmetic sees exactly the numbers the reference would (32 B per atom per step over PCIe -- these
// integrators are for small systems).  GJF keeps its previous force and previous noise per atom on the device.
template <bool GJF>
__global__ void __launch_bounds__(INTGR_BLOCK)
k_integrate_stoch(d4 *__restrict__ x4, d4 *__restrict__ v4, const d4 *__restrict__ f4, const d4 *__restrict__ xn4,
                  i4 *__restrict__ cr4, int *__restrict__ crossings, const int *__restrict__ rank, d4 *__restrict__ xs,
                  const d4 *__restrict__ noise, d4 *__restrict__ prevf4, d4 *__restrict__ randn4,
                  IntgrParams P, double temp, double alpha, double *__restrict__ partial)
{
    __shared__ double red[SEPGPU_NPART_I * (INTGR_BLOCK / 32)];
    double acc[SEPGPU_NPART_I];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_I; q++) acc[q] = 0.0;
    const double dt = P.dt;
    const double cc = exp(-alpha * dt);

    for (int i = blockIdx.x * INTGR_BLOCK + threadIdx.x; i < P.n; i += gridDim.x * INTGR_BLOCK) {
        d4 x = x4[i], v = v4[i], f;
        if (P.f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
        const double m = v.w;
        const d4 g = noise[i];
        i4 cr = cr4[i];
        const d4 xn = xn4[i];
        int cl[3]; unpack_cl(cr.w, cl[0], cl[1], cl[2]);
        int t[3] = {0, 0, 0}; bool changed = false;
        d4 pf, rn;
        if (GJF) { pf = prevf4[i]; rn = randn4[i]; } else { pf.x = pf.y = pf.z = pf.w = 0.0; rn = pf; }
        const double d2 = stoch_atom<GJF>(x, v, f, g, pf, rn, xn, cr, cl, t, changed, P.Lx, P.Ly, P.Lz, dt, temp, alpha, cc);
        if (GJF) { prevf4[i] = pf; randn4[i] = rn; }
        const int clx = cl[0], cly = cl[1], clz = cl[2], tx = t[0], ty = t[1], tz = t[2];
        acc[0] += v.x * v.x * m; acc[0] += v.y * v.y * m; acc[0] += v.z * v.z * m;     // :258 / :113
        acc[1] += v.x * v.x * m; acc[2] += v.x * v.y * m; acc[3] += v.x * v.z * m;
        acc[4] += v.y * v.y * m; acc[5] += v.y * v.z * m; acc[6] += v.z * v.z * m;
        acc[7] = fmax(acc[7], d2);
        acc[8] += (v.x * v.x + v.y * v.y + v.z * v.z) * m;
        acc[9] += v.x * m; acc[10] += v.y * m; acc[11] += v.z * m;
        x4[i] = x; v4[i] = v;
        if (changed) {
            cr.w = pack_cl(clx, cly, clz);
            cr4[i] = cr;
            if (tx) crossings[3 * i] += tx;
            if (ty) crossings[3 * i + 1] += ty;
            if (tz) crossings[3 * i + 2] += tz;
        }
        if (P.write_xs) {
            d4 u; u.x = x.x + clx * P.Lx; u.y = x.y + cly * P.Ly; u.z = x.z + clz * P.Lz; u.w = x.w;
            xs[rank[i]] = u;
        }
    }
    const double mymax = acc[7];
    acc[7] = 0.0;
    block_sum<SEPGPU_NPART_I, INTGR_BLOCK>(acc, red);
    __syncthreads();
    double wm = warp_max(mymax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int w = 0; w < INTGR_BLOCK / 32; w++) mx = fmax(mx, red[w]);
        acc[7] = mx;
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_I; q++) partial[blockIdx.x * SEPGPU_NPART_I + q] = acc[q];
    }
}

// noise4: host array, 4 doubles per atom {g0, g1, g2, ldiff}
static int run_stochastic(sepgpu_ctx *c, const sepgpu_sys *sys, bool gjf, double temp, double alpha, const double *noise4)
{
    if (c->dd) { sepgpu_set_error("stochastic integrators are not available in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    int rc = sepgpu_apply_pending(c);
    if (rc) return rc;
    const size_t bytes = sizeof(d4) * (size_t)c->n_own;
    if ((rc = sepgpu_ensure_stage(c, bytes))) return rc;
    memcpy(c->stage, noise4, bytes);
    CUDA_TRY(cudaMemcpyAsync(c->dstage, c->stage, bytes, cudaMemcpyHostToDevice, c->stream));
    if (gjf && !c->prevf4) {
        CUDA_TRY(cudaMalloc((void **)&c->prevf4, sizeof(d4) * (size_t)c->ncap));
        CUDA_TRY(cudaMalloc((void **)&c->randn4, sizeof(d4) * (size_t)c->ncap));
        CUDA_TRY(cudaMemsetAsync(c->prevf4, 0, sizeof(d4) * (size_t)c->ncap, c->stream));   // sep_init zeroes both, source/sepinit.c:45-46
        CUDA_TRY(cudaMemsetAsync(c->randn4, 0, sizeof(d4) * (size_t)c->ncap, c->stream));
    }
    IntgrParams P;
    P.Lx = sys->length[0]; P.Ly = sys->length[1]; P.Lz = sys->length[2];
    P.dt = sys->dt; P.skin = sys->skin; P.n = c->n_own;
    P.alpha_slot = -1; P.alpha_type = -1;
    P.f_zero = c->f_zero ? 1 : 0;
    P.write_xs = (sys->neighb_update != 0 && c->list_valid) ? 1 : 0;
    long long want = ((long long)c->n_own + INTGR_BLOCK - 1) / INTGR_BLOCK;
    const int grid = (int)(want < INTGR_MAX_GRID ? want : INTGR_MAX_GRID);
    if (gjf)
        k_integrate_stoch<true><<<grid, INTGR_BLOCK, 0, c->stream>>>(c->x4, c->v4, c->f4, c->xn4, c->cr4, c->crossings, c->rank, c->xs,
            (const d4 *)c->dstage, c->prevf4, c->randn4, P, temp, alpha, c->partial);
    else
        k_integrate_stoch<false><<<grid, INTGR_BLOCK, 0, c->stream>>>(c->x4, c->v4, c->f4, c->xn4, c->cr4, c->crossings, c->rank, c->xs,
            (const d4 *)c->dstage, NULL, NULL, P, temp, alpha, c->partial);
    const int resets = (c->ret_reset_pending ? 1 : 0) | (c->maxd_reset_pending ? 2 : 0);
    c->ret_reset_pending = false; c->maxd_reset_pending = false;
    k_finalize_intgr<<<1, 256, 0, c->stream>>>(c->partial, grid, c->scal, sys->skin, 0, NULL, resets, 0, 1);
    KERNEL_CHECK();
    c->mv2_valid = true;
    c->moved_since_build = true;
    if (!P.write_xs) c->xs_current = false;
    CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->scal_host->neighb_flag) {
        k_set_xn<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->x4, c->xn4, c->cr4, c->n_own);
        KERNEL_CHECK();
        c->list_valid = false;
    }
    return 0;
}

extern "C" int sepgpu_fp(sepgpu_ctx *c, const sepgpu_sys *sys, double temp, const double *noise4)
{
    if (!c || !sys || !noise4) return SEPGPU_EINVAL;
    return run_stochastic(c, sys, false, temp, 0.0, noise4);
}

extern "C" int sepgpu_langevin_gjf(sepgpu_ctx *c, const sepgpu_sys *sys, double temp, double alpha, const double *noise4)
{
    if (!c || !sys || !noise4) return SEPGPU_EINVAL;
    return run_stochastic(c, sys, true, temp, alpha, noise4);
}

extern "C" int sepgpu_leapfrog(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    return run_integrator(c, sys, false, 0.0, 0);
}

extern "C" int sepgpu_verlet_dpd(sepgpu_ctx *c, const sepgpu_sys *sys, double lambda, int stepnow)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    return run_integrator(c, sys, true, lambda, stepnow);
}

// ---- thermostat -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INTGR_BLOCK)
k_sum_mv2(const d4 *__restrict__ v4, const d4 *__restrict__ x4, int n, int type, double *__restrict__ partial)
{
    __shared__ double red[2 * (INTGR_BLOCK / 32)];
    double acc[2] = {0.0, 0.0};
    for (int i = blockIdx.x * INTGR_BLOCK + threadIdx.x; i < n; i += gridDim.x * INTGR_BLOCK) {
        if (type >= 0 && tag_type(x4[i].w) != type) continue;
        const d4 v = v4[i];
        acc[0] += (v.x * v.x + v.y * v.y + v.z * v.z) * v.w;
        acc[1] += 1.0;
    }
    block_sum<2, INTGR_BLOCK>(acc, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = acc[0]; partial[2 * blockIdx.x + 1] = acc[1]; }
}

// mode 0: sep_nosehoover (source/sepintgr.c:157-161); mode 1: _sep_nosehoover_type (:180-187)
__global__ void __launch_bounds__(256)
k_nh_update(const double *__restrict__ partial, int nrows, DevScalars *scal, int slot, int mode,
            double temp0, double tau_or_Q, double dt, double npart, double a0, double a1, double a2)
{
    __shared__ double red[2 * 8];
    double v[2] = {0.0, 0.0};
    if (nrows > 0) {
        for (int r = threadIdx.x; r < nrows; r += 256) { v[0] += partial[2 * r]; v[1] += partial[2 * r + 1]; }
        block_sum<2, 256>(v, red);
    }
    if (threadIdx.x == 0) {
        const double sum = nrows > 0 ? v[0] : scal->sum_mv2;
        if (mode == 0) {
            const double ekin = 0.5 * sum / npart;
            const double temp = 0.666667 * ekin;
            scal->alpha[slot] = scal->alpha[slot] + dt / (tau_or_Q * tau_or_Q) * (temp / temp0 - 1.0);
        } else {
            const double g = 3.0 * v[1] - 3.0;
            // history shift: alpha[0]<-alpha[1], alpha[1]<-alpha[2], alpha[2]<-old alpha[0] + ...   (slots 4..6: the
            // caller's sep_nosehoover multipliers live in slots 0..3 and must not be touched)
            scal->alpha[4] = a1;
            scal->alpha[5] = a2;
            scal->alpha[6] = a0 + 2.0 * dt * (sum - g * temp0) / tau_or_Q;
        }
    }
}

__global__ void k_nh_update_fold(DevScalars *scal, int slot, double temp0, double tau, double dt, double npart)
{
    NhFold N; N.temp0 = temp0; N.tau = tau; N.npart = npart;
    scal->alpha[slot] = nh_alpha_next(scal->alpha[slot], scal->sum_mv2, N, dt);
}

int sepgpu_nh_update_now(sepgpu_ctx *c);

extern "C" int sepgpu_nosehoover(sepgpu_ctx *c, const sepgpu_sys *sys, double temp0, int slot, double tau)
{
    if (!c || !sys || slot < 0 || slot > 3) return SEPGPU_EINVAL;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = sepgpu_apply_pending(c);          // two thermostat calls in one step: flush the first
    if (rc) return rc;
    int nrows = 0;
    if (!c->mv2_valid) {
        if (c->fin_pending.active && (rc = sepgpu_settle(c))) return rc;      // k_sum_mv2 writes into c->partial
        long long want = ((long long)c->n_own + INTGR_BLOCK - 1) / INTGR_BLOCK;
        nrows = (int)(want < INTGR_MAX_GRID ? want : INTGR_MAX_GRID);
        k_sum_mv2<<<nrows, INTGR_BLOCK, 0, c->stream>>>(c->v4, c->x4, c->n_own, -1, c->partial);
        if (c->dd) {                       // global sum m v^2 over the ranks
            double *comm = sepgpu_dd_comm(c);
            k_partial2_to_comm<<<1, 256, 0, c->stream>>>(c->partial, nrows, comm);
            if ((rc = sepgpu_dd_allreduce(c, comm, 2, NULL, 0))) return rc;
            k_comm_to_mv2<<<1, 1, 0, c->stream>>>(c->scal, comm);
            nrows = 0;
        }
    }
    if (c->step_fold && (!c->dd || sepgpu_dd_uses_p2p(c)) && nrows == 0 && !c->nh_pending.active) {
        // option step_fold: sum m v^2 of the last integrator call is current, so the update needs no pass of its own --
        // the integrator evaluates it (k_integrate<.., true>) and k_finalize_both stores it
        c->nh_pending.active = true; c->nh_pending.slot = slot; c->nh_pending.temp0 = temp0; c->nh_pending.tau = tau;
        c->nh_dt = sys->dt;
    } else {
        if (c->nh_pending.active && (rc = sepgpu_nh_update_now(c))) return rc;
        k_nh_update<<<1, 256, 0, c->stream>>>(c->partial, nrows, c->scal, slot, 0, temp0, tau, sys->dt, (double)c->n_global, 0, 0, 0);
        KERNEL_CHECK();
    }
    c->pending_alpha_slot = slot;
    c->pending_alpha_type = -1;
    return 0;
}

// the pending multiplier update as a launch of its own (somebody needs alpha, or the force, before the integrator runs)
int sepgpu_nh_update_now(sepgpu_ctx *c)
{
    if (!c->nh_pending.active) return 0;
    c->nh_pending.active = false;
    k_nh_update_fold<<<1, 1, 0, c->stream>>>(c->scal, c->nh_pending.slot, c->nh_pending.temp0, c->nh_pending.tau, c->nh_dt, (double)c->n_global);
    KERNEL_CHECK();
    return 0;
}

extern "C" int sepgpu_nosehoover_type(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double Td,
                                      double alpha3[3], double Q)
{
    if (!c || !sys || !alpha3) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    int rc = sepgpu_apply_pending(c);
    if (rc) return rc;
    long long want = ((long long)c->n_own + INTGR_BLOCK - 1) / INTGR_BLOCK;
    const int nrows = (int)(want < INTGR_MAX_GRID ? want : INTGR_MAX_GRID);
    k_sum_mv2<<<nrows, INTGR_BLOCK, 0, c->stream>>>(c->v4, c->x4, c->n_own, (unsigned char)type, c->partial);
    k_nh_update<<<1, 256, 0, c->stream>>>(c->partial, nrows, c->scal, 0, 1, Td, Q, sys->dt, (double)c->n_global,
                                          alpha3[0], alpha3[1], alpha3[2]);
    KERNEL_CHECK();
    // the history is host-visible state of the caller: read it back (3 doubles)
    CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    alpha3[0] = c->scal_host->alpha[4]; alpha3[1] = c->scal_host->alpha[5]; alpha3[2] = c->scal_host->alpha[6];
    c->pending_alpha_slot = 5;                 // the multiplier applied is alpha[1] of the history (source/sepintgr.c:193)
    c->pending_alpha_type = (unsigned char)type;
    return 0;
}

__global__ void k_apply_alpha(d4 *__restrict__ f4, const d4 *__restrict__ v4, const d4 *__restrict__ x4,
                              const DevScalars *__restrict__ scal, int slot, int type, int n, int f_zero)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 f;
    if (f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
    if (type < 0 || tag_type(x4[i].w) == type) {
        const d4 v = v4[i];
        const double am = scal->alpha[slot] * v.w;
        f.x -= am * v.x; f.y -= am * v.y; f.z -= am * v.z;
    }
    f4[i] = f;
}

int sepgpu_apply_pending(sepgpu_ctx *c)
{
    if (c->pending_alpha_slot < 0) return 0;
    if (c->nh_pending.active) { int rs = sepgpu_nh_update_now(c); if (rs) return rs; }     // the multiplier about to be applied
    k_apply_alpha<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->f4, c->v4, c->x4, c->scal, c->pending_alpha_slot,
                                                            c->pending_alpha_type, c->n_own, c->f_zero ? 1 : 0);
    KERNEL_CHECK();
    c->pending_alpha_slot = -1;
    c->pending_alpha_type = -1;
    c->f_zero = false;
    return 0;
}

// ---- momentum reset / rescale ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INTGR_BLOCK)
k_sum_mom(const d4 *__restrict__ v4, const d4 *__restrict__ x4, int n, int type, double *__restrict__ partial)
{
    __shared__ double red[4 * (INTGR_BLOCK / 32)];
    double acc[4] = {0, 0, 0, 0};
    for (int i = blockIdx.x * INTGR_BLOCK + threadIdx.x; i < n; i += gridDim.x * INTGR_BLOCK) {
        if (tag_type(x4[i].w) != type) continue;
        const d4 v = v4[i];
        acc[0] += v.x * v.w; acc[1] += v.y * v.w; acc[2] += v.z * v.w; acc[3] += v.w;
    }
    block_sum<4, INTGR_BLOCK>(acc, red);
    if (threadIdx.x == 0) for (int q = 0; q < 4; q++) partial[4 * blockIdx.x + q] = acc[q];
}

__global__ void __launch_bounds__(256) k_mom_final(const double *__restrict__ partial, int nrows, DevScalars *scal)
{
    __shared__ double red[4 * 8];
    double v[4] = {0, 0, 0, 0};
    for (int r = threadIdx.x; r < nrows; r += 256) for (int q = 0; q < 4; q++) v[q] += partial[4 * r + q];
    block_sum<4, 256>(v, red);
    if (threadIdx.x == 0) for (int q = 0; q < 4; q++) scal->mom[q] = v[q];
}

__global__ void k_sub_mom(d4 *__restrict__ v4, const d4 *__restrict__ x4, int n, int type, const DevScalars *__restrict__ scal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || tag_type(x4[i].w) != type) return;
    const double mass = scal->mom[3];
    d4 v = v4[i];
    v.x -= scal->mom[0] / mass; v.y -= scal->mom[1] / mass; v.z -= scal->mom[2] / mass;   // source/sepmisc.c:1190
    v4[i] = v;
}

extern "C" int sepgpu_reset_momentum(sepgpu_ctx *c, char type)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    long long want = ((long long)c->n_own + INTGR_BLOCK - 1) / INTGR_BLOCK;
    const int nrows = (int)(want < INTGR_MAX_GRID ? want : INTGR_MAX_GRID);
    k_sum_mom<<<nrows, INTGR_BLOCK, 0, c->stream>>>(c->v4, c->x4, c->n_own, (unsigned char)type, c->partial);
    k_mom_final<<<1, 256, 0, c->stream>>>(c->partial, nrows, c->scal);
    k_sub_mom<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->v4, c->x4, c->n_own, (unsigned char)type, c->scal);
    KERNEL_CHECK();
    c->mv2_valid = false;
    return 0;
}

__global__ void k_scale_x(d4 *__restrict__ x4, d4 *__restrict__ xs, int n, double xi, int scale_xs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 x = x4[i]; x.x *= xi; x.y *= xi; x.z *= xi; x4[i] = x;
    if (scale_xs) { d4 u = xs[i]; u.x *= xi; u.y *= xi; u.z *= xi; xs[i] = u; }
}

extern "C" int sepgpu_scale_positions(sepgpu_ctx *c, double xi)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    k_scale_x<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->x4, c->xs, c->n_own, xi, c->list_valid ? 1 : 0);
    KERNEL_CHECK();
    return 0;
}


// ---- the prg1 loop, driven from C (reference prgs/prg1.c:52-70) ---------------------------------------------------------
extern "C" int sepgpu_md_lj_nvt(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], const sepgpu_ljparam *p, unsigned opt,
                                double temp, int alpha_slot, double tau, int nsteps)
{
    if (!c || !sys || !types || !p || nsteps < 0) return SEPGPU_EINVAL;
    for (int n = 0; n < nsteps; n++) {
        int rc = sepgpu_reset_ret(c);
        if (!rc) rc = sepgpu_reset_force(c);
        if (!rc) rc = sepgpu_force_lj(c, sys, types, p, opt, 1);
        if (!rc) rc = sepgpu_nosehoover(c, sys, temp, alpha_slot, tau);
        if (!rc) rc = sepgpu_leapfrog(c, sys);
        if (rc) return rc;
    }
    return 0;
}
