// sepgpu_force.cu -- pair-force kernels (Lennard-Jones family, shifted-force Coulomb, DPD).
//
// Stand-ins for sep_force_pair_neighb / sep_lj_pair_neighb / sep_force_pair_brute / sep_lj_pair_brute
// (reference source/sepprfrc.c:94-224, 782-1000), sep_coulomb_sf_{neighb,brute}
// (source/sepcoulomb.c:20-160) and sep_dpdforce_{neighb,brute} (source/sepprfrc.c:1007-1231).
//
// The reference scatters +f/-f over a half list.  Here every atom owns its force: TPA lanes share one
// atom, stride through its FULL neighbour row, keep fx,fy,fz in registers, butterfly-reduce with warp
// shuffles and issue exactly one 256-bit store per atom -- no global atomics.  Energy and virial are
// reduced warp -> block -> one partial row per block -> fixed-order final sum (deterministic), and
// halved because each pair is seen from both ends.
//
// Positions come from the cell-sorted copy xs (continuous coordinates since the last rebuild) so that
// neighbour gathers hit L1/L2; each gather is one 32-byte sector fetched by a single LDG.256.
// The periodic image of a pair is fixed at list-build time and stored in the top bits of the list
// entry, which removes the 9 FP64 compare/adjust operations of sep_Wrap from the inner loop.
#include "sepgpu_internal.cuh"
#include "sepgpu_pair.cuh"

#include <math.h>

#define FORCE_BLOCK 128
#define FORCE_MAX_GRID (148 * 64)
#ifndef LJ_MIN_CTAS
#define LJ_MIN_CTAS 7          // 72 registers per thread; forcing 8 CTAs (64 registers) spills and measured 11 % slower
#endif

// One listed pair, branch-free: out-of-range (or wrong-type) pairs run the same arithmetic with the
// force factor selected to zero.  In a warp some lane is almost always in range, so the branchy form
// executes the full block anyway; without the branch the compiler can interleave the dependent DFMA
// chains of two pairs.  r2 of a real pair is finite and non-zero, so the masked values stay finite.
//
// The pair body is 20 FP64 instructions (32 in the first version).  Measured on B200 the kernel time did
// NOT follow (0.268 -> 0.265 ms at 1e6 atoms): the binding unit is the L1 data pipe -- the 32 lanes of one
// neighbour gather touch ~22 different 128-byte lines, one wavefront each (ncu: l1tex data pipe 79 %, 19
// wavefronts per load request), the FP64 pipe needs ~10 cycles.  The leaner body is kept because it frees issue slots and power:
//  * the cutoff test is a 64-bit INTEGER compare of the bit patterns (both sides are non-negative doubles);
//  * the prefactor 48 eps is applied once per atom, not per pair;
//  * the virial is NOT accumulated per pair.  With a full list every pair is seen from both ends with
//    exactly negated separation d_ij = x_i - x_j - S_ij (S = image shift), so
//        sum_i sum_j g_ij (x) d_ij  =  2 sum_i F_i (x) x_i  -  sum_i sum_j g_ij (x) S_ij ,
//    i.e. one outer product per ATOM (list epilogue) plus a correction on the rare pairs that cross a
//    periodic boundary (code != 13, already a branch).  F_i is the force of THIS call only.  The sum
//    is the reference's pot_P (source/sepprfrc.c:195-197) up to rounding: terms are O(|F| L) instead of
//    O(|F| rc), costing log10(L/rc) < 2 of the 16 digits.
// FIJ: also accumulate the molecule-molecule force table Fij[mi][mj] += f (reference
// source/sepprfrc.c:199-207; each directed pair adds its own direction, the partner thread adds -f to
// Fij[mj][mi]).  Only small systems carry the table (<= SEP_MAX_NUM_MOL molecules), the adds are FP64
// atomics on nmol^2*3 scattered addresses -- off for every benchmarked configuration.
template <bool TYPED, bool FIJ>
__device__ __forceinline__ void lj_pair(const d4 &pi, const d4 &pj, unsigned e, int ti, const LJDev &P,
                                        const BoxDev &B, PairAcc &A, double *fij = nullptr, int nmol = 0)
{
    double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const int code = (int)(e >> SEPGPU_SHIFT_BITS);
    if (code != 13) apply_image(code, B, dx, dy, dz);
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    bool in = __double_as_longlong(r2) < __double_as_longlong(P.cf2);
    if (TYPED) {
        const int tj = tag_type(pj.w);
        in = in && ((ti == P.t0 && tj == P.t1) || (ti == P.t1 && tj == P.t0));   // source/sepprfrc.c:171-172
    }
    double ft;
    if (FIJ && P.tab) {                                            // (the FIJ instantiation is the general-purpose one)
        // user pair function sampled by the host layer (sepgpu_force_table): force factor and energy from the table
        bool below;
        const double2 fu = table_eval(P, in ? r2 : P.cf2, below);
        if (in && below) A.nin |= 0x40000000;
        ft = in ? fu.x : 0.0;
        A.u += in ? fu.y : 0.0;
    } else {
        const double rri = P.sig2 * fast_rcp3(r2);
        double rri3 = rri * rri * rri;
        ft = rri3 * (rri3 - P.awh) * rri;                          // source/sepmisc.c:139, sepprfrc.c:888 (/ 48 eps)
        const double uu = rri3 - P.aw;
        ft = in ? ft : 0.0;
        rri3 = in ? rri3 : 0.0;
        A.u = fma(rri3, uu, A.u);                                  // u/(4 eps) before the shift
        A.nin += in ? 1 : 0;
    }
    A.fx = fma(ft, dx, A.fx); A.fy = fma(ft, dy, A.fy); A.fz = fma(ft, dz, A.fz);
    if (code != 13) {                                              // boundary-crossing pair: - g (x) S
        double sx = 0.0, sy = 0.0, sz = 0.0;
        apply_image(code, B, sx, sy, sz);                          // s = -S
        const double g = P.eps48 * ft;
        virial_add(A.v, g * dx, g * dy, g * dz, sx, sy, sz);
    }
    if (FIJ && in && fij) {
        const int mi = tag_mol(pi.w), mj = tag_mol(pj.w);
        if (mi != -1 && mj != -1 && mi != mj) {
            const double g = P.eps48 * ft;
            double *t = fij + ((size_t)mi * nmol + mj) * 3;
            atomicAdd(t, g * dx); atomicAdd(t + 1, g * dy); atomicAdd(t + 2, g * dz);
        }
    }
}

// Two listed pairs as ONE straight-line block, so that the scheduler interleaves their dependent FP64 chains
// (a pair is ~16 dependent FP64 operations deep; one chain at a time leaves the pipe idle).  The image shift
// and its virial correction are taken by both pairs when either needs them (code 13 shifts by zero).
// valid1 == false: the second slot is past the end of the row; its gather points at a real atom and its
// contribution is masked out.
template <bool TYPED>
__device__ __forceinline__ void lj_pair2(const d4 &pi, const d4 &p0, const d4 &p1, unsigned e0, unsigned e1, bool valid1,
                                         int ti, const LJDev &P, const BoxDev &B, PairAcc &A)
{
    double dx0 = pi.x - p0.x, dy0 = pi.y - p0.y, dz0 = pi.z - p0.z;
    double dx1 = pi.x - p1.x, dy1 = pi.y - p1.y, dz1 = pi.z - p1.z;
    const int c0 = (int)(e0 >> SEPGPU_SHIFT_BITS), c1 = valid1 ? (int)(e1 >> SEPGPU_SHIFT_BITS) : 13;
    const bool shifted = (c0 != 13) | (c1 != 13);
    if (shifted) { apply_image(c0, B, dx0, dy0, dz0); apply_image(c1, B, dx1, dy1, dz1); }
    const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
    const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
    bool in0 = __double_as_longlong(r20) < __double_as_longlong(P.cf2);
    bool in1 = (__double_as_longlong(r21) < __double_as_longlong(P.cf2)) && valid1;
    if (TYPED) {
        const int t0 = tag_type(p0.w), t1 = tag_type(p1.w);
        in0 = in0 && ((ti == P.t0 && t0 == P.t1) || (ti == P.t1 && t0 == P.t0));     // source/sepprfrc.c:171-172
        in1 = in1 && ((ti == P.t0 && t1 == P.t1) || (ti == P.t1 && t1 == P.t0));
    }
    const double a0 = P.sig2 * fast_rcp3(r20), a1 = P.sig2 * fast_rcp3(r21);
    double b0 = a0 * a0 * a0, b1 = a1 * a1 * a1;
    double f0 = b0 * (b0 - P.awh) * a0, f1 = b1 * (b1 - P.awh) * a1;   // source/sepmisc.c:139, sepprfrc.c:888 (/ 48 eps)
    const double u0 = b0 - P.aw, u1 = b1 - P.aw;
    f0 = in0 ? f0 : 0.0; f1 = in1 ? f1 : 0.0;
    b0 = in0 ? b0 : 0.0; b1 = in1 ? b1 : 0.0;
    A.fx = fma(f0, dx0, A.fx); A.fy = fma(f0, dy0, A.fy); A.fz = fma(f0, dz0, A.fz);
    A.u = fma(b0, u0, A.u);
    A.fx = fma(f1, dx1, A.fx); A.fy = fma(f1, dy1, A.fy); A.fz = fma(f1, dz1, A.fz);
    A.u = fma(b1, in1 ? u1 : 0.0, A.u);
    A.nin += (in0 ? 1 : 0) + (in1 ? 1 : 0);
    if (shifted) {                                                  // boundary-crossing pairs: - g (x) S
        double sx = 0.0, sy = 0.0, sz = 0.0;
        apply_image(c0, B, sx, sy, sz);
        double g = P.eps48 * f0;
        virial_add(A.v, g * dx0, g * dy0, g * dz0, sx, sy, sz);
        sx = sy = sz = 0.0;
        apply_image(c1, B, sx, sy, sz);
        g = P.eps48 * f1;
        virial_add(A.v, g * dx1, g * dy1, g * dz1, sx, sy, sz);
    }
}

// STORE: first force kernel after sep_reset_force -> plain store instead of read-modify-write.
// Each CTA walks a CONTIGUOUS range of the cell-sorted atoms so that the neighbour rows it gathers
// stay resident in its SM's L1.  A lane reads its next FOUR list entries with one streaming 128-bit
// load (ld.global.cs: never reused) issued one chunk ahead of use -- the list comes from HBM and was
// the dominant stall -- and then gathers the four neighbour sectors two at a time.
template <int TPA, bool TYPED, bool STORE, bool FIJ>
__global__ void __launch_bounds__(FORCE_BLOCK, LJ_MIN_CTAS)
k_lj_list(const d4 *__restrict__ xs, const unsigned *__restrict__ nbr, const int *__restrict__ cnt,
          const int *__restrict__ order, d4 *__restrict__ f4, int n, int npad, int atoms_per_cta,
          LJDev P, BoxDev B, double *__restrict__ partial, double *fij, int nmol,
          const unsigned char *__restrict__ cls, int want, DevScalars *scal)
{
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    const int sub = threadIdx.x % TPA;
    constexpr int GROUPS = FORCE_BLOCK / TPA;
    PairAcc A;
    A.u = 0.0;
    A.nin = 0;
#pragma unroll
    for (int q = 0; q < 6; q++) A.v[q] = 0.0;
    const int first = blockIdx.x * atoms_per_cta;
    const int last = min(n, first + atoms_per_cta);
    const uint4 *nbrv = reinterpret_cast<const uint4 *>(nbr);

    for (int s0 = first; s0 < last; s0 += GROUPS) {
        const int s = s0 + threadIdx.x / TPA;
        // cls != NULL: decomposed run split into an interior pass (want 0, needs no halo) and a pass over the
        // atoms next to a halo layer (want 1); halo atoms themselves own no rows
        const bool valid = s < last && (cls == nullptr || cls[s] == want);
        A.fx = A.fy = A.fz = 0.0;
        if (valid) {
            const d4 pi = xs[s];
            int m = cnt[s];
            int ti = 0;
            if (TYPED) {
                ti = tag_type(pi.w);
                if (ti != P.t0 && ti != P.t1) m = 0;             // source/sepprfrc.c:164-165
            }
            const int nch = (m + 3) >> 2;
            const uint4 *row = nbrv + s;
            int c = sub;
            uint4 cur = make_uint4(0, 0, 0, 0);
            if (c < nch) cur = __ldcs(row + (size_t)c * npad);
            while (c < nch) {
                const int cn = c + TPA;
                uint4 nxt = make_uint4(0, 0, 0, 0);
                if (cn < nch) nxt = __ldcs(row + (size_t)cn * npad);
                const int left = m - 4 * c;                      // >= 1 valid entries in this chunk
                if (FIJ) {
                    const bool v1 = left > 1;
                    const d4 p0 = xs[cur.x & SEPGPU_INDEX_MASK];
                    const d4 p1 = xs[v1 ? (cur.y & SEPGPU_INDEX_MASK) : (unsigned)s];
                    lj_pair<TYPED, FIJ>(pi, p0, cur.x, ti, P, B, A, fij, nmol);
                    if (v1) lj_pair<TYPED, FIJ>(pi, p1, cur.y, ti, P, B, A, fij, nmol);
                    if (left > 2) {
                        const bool v3 = left > 3;
                        const d4 p2 = xs[cur.z & SEPGPU_INDEX_MASK];
                        const d4 p3 = xs[v3 ? (cur.w & SEPGPU_INDEX_MASK) : (unsigned)s];
                        lj_pair<TYPED, FIJ>(pi, p2, cur.z, ti, P, B, A, fij, nmol);
                        if (v3) lj_pair<TYPED, FIJ>(pi, p3, cur.w, ti, P, B, A, fij, nmol);
                    }
                } else {
                    {
                        const bool v1 = left > 1;
                        SEPGPU_EMU_GATHER(&xs[cur.x & SEPGPU_INDEX_MASK]);
                        SEPGPU_EMU_GATHER(&xs[(v1 ? cur.y : cur.x) & SEPGPU_INDEX_MASK]);
                        const d4 p0 = xs[cur.x & SEPGPU_INDEX_MASK];
                        const d4 p1 = xs[(v1 ? cur.y : cur.x) & SEPGPU_INDEX_MASK];
                        lj_pair2<TYPED>(pi, p0, p1, cur.x, cur.y, v1, ti, P, B, A);
                    }
                    if (left > 2) {
                        const bool v3 = left > 3;
                        SEPGPU_EMU_GATHER(&xs[cur.z & SEPGPU_INDEX_MASK]);
                        SEPGPU_EMU_GATHER(&xs[(v3 ? cur.w : cur.z) & SEPGPU_INDEX_MASK]);
                        const d4 p2 = xs[cur.z & SEPGPU_INDEX_MASK];
                        const d4 p3 = xs[(v3 ? cur.w : cur.z) & SEPGPU_INDEX_MASK];
                        lj_pair2<TYPED>(pi, p2, p3, cur.z, cur.w, v3, ti, P, B, A);
                    }
                }
                cur = nxt;
                c = cn;
            }
        }
        // butterfly over the TPA lanes of this atom
#pragma unroll
        for (int o = TPA / 2; o > 0; o >>= 1) {
            A.fx += __shfl_xor_sync(0xffffffffu, A.fx, o);
            A.fy += __shfl_xor_sync(0xffffffffu, A.fy, o);
            A.fz += __shfl_xor_sync(0xffffffffu, A.fz, o);
        }
        if (valid && sub == 0) {
            const int i = order[s];
            const double fx = P.eps48 * A.fx, fy = P.eps48 * A.fy, fz = P.eps48 * A.fz;
            if (STORE) {
                d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0.0;
                f4[i] = o;
            } else {
                d4 o = f4[i];
                o.x += fx; o.y += fy; o.z += fz;
                f4[i] = o;
            }
            // per-atom part of the virial (see lj_pair): 2 F_i (x) x_i, upper triangle
            const d4 pi = xs[s];
            const double tx = fx + fx, ty = fy + fy, tz = fz + fz;
            virial_add(A.v, tx, ty, tz, pi.x, pi.y, pi.z);
        }
    }
    double acc[SEPGPU_NPART_F];
    if (FIJ && P.tab) {
        acc[0] = A.u;
        if (A.nin & 0x40000000) scal->error = SEPGPU_ETABLE;
    } else {
        acc[0] = P.eps4 * A.u - P.shift * (double)A.nin;
    }
    acc[1] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) acc[2 + q] = A.v[q];
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
    }
}

// ---- Lennard-Jones, all pairs (SEP_BRUTE) ------------------------------------------------------------------
// exact reference arithmetic for the separation (wrapped x, sep_Wrap branches)
__device__ __forceinline__ int share_tab_f(const int *__restrict__ tab, int width, int a, int b)
{
    for (int k = 0; k < width; k++) {
        int ta = tab[a * width + k], tb = tab[b * width + k];
        if (ta == -1 || tb == -1) break;
        if (ta == b || tb == a) return 1;
    }
    return 0;
}

template <bool STORE>
__global__ void __launch_bounds__(FORCE_BLOCK)
k_lj_brute(const d4 *__restrict__ x4, d4 *__restrict__ f4, int n, LJDev P, BoxDev B, unsigned opt,
           const int *__restrict__ excl_bond, double *__restrict__ partial, double *fij, int nmol, DevScalars *scal)
{
    __shared__ d4 tile[FORCE_BLOCK];
    __shared__ double red[SEPGPU_NPART_F * (FORCE_BLOCK / 32)];
    const int i = blockIdx.x * FORCE_BLOCK + threadIdx.x;
    const bool valid = i < n;
    d4 pi; pi.x = pi.y = pi.z = 0; pi.w = 0;
    if (valid) pi = x4[i];
    const int ti = tag_type(pi.w), mi = tag_mol(pi.w);
    const double hx = 0.5 * B.Lx, hy = 0.5 * B.Ly, hz = 0.5 * B.Lz;
    double fx = 0, fy = 0, fz = 0;
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    int nin = 0;
    bool below_seen = false;
    for (int j0 = 0; j0 < n; j0 += FORCE_BLOCK) {
        __syncthreads();
        if (j0 + threadIdx.x < n) tile[threadIdx.x] = x4[j0 + threadIdx.x];
        __syncthreads();
        const int lim = min(FORCE_BLOCK, n - j0);
        if (!valid) continue;
        for (int t = 0; t < lim; t++) {
            const int j = j0 + t;
            if (j == i) continue;
            const d4 pj = tile[t];
            const int tj = tag_type(pj.w);
            if (!((ti == P.t0 && tj == P.t1) || (ti == P.t1 && tj == P.t0))) continue;
            if (opt == SEPGPU_EXCL_SAME_MOL) { if (mi == tag_mol(pj.w) && mi != -1) continue; }   // :36-39
            else if (opt == SEPGPU_EXCL_BONDED) { if (share_tab_f(excl_bond, 10, min(i, j), max(i, j)) == 1) continue; }  // :33
            const double dx = wrap_exact(pi.x - pj.x, B.Lx, hx);
            const double dy = wrap_exact(pi.y - pj.y, B.Ly, hy);
            const double dz = wrap_exact(pi.z - pj.z, B.Lz, hz);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (r2 < P.cf2) {
                double ft;
                if (P.tab) {                                     // user pair function (tabulated by the host layer)
                    bool below;
                    const double2 fu = table_eval(P, r2, below);
                    if (below) below_seen = true;
                    ft = fu.x;
                    acc[1] += fu.y;                              // energy, moved to slot 0 below
                } else {
                    const double rri = P.sig2 / r2;
                    const double rri3 = rri * rri * rri;
                    ft = P.eps48 * rri3 * (rri3 - P.awh) * rri;
                    acc[0] = fma(rri3, rri3 - P.aw, acc[0]);
                    nin++;
                }
                const double gx = ft * dx, gy = ft * dy, gz = ft * dz;
                fx += gx; fy += gy; fz += gz;
                acc[2] = fma(gx, dx, acc[2]); acc[3] = fma(gx, dy, acc[3]); acc[4] = fma(gx, dz, acc[4]);
                acc[5] = fma(gy, dy, acc[5]); acc[6] = fma(gy, dz, acc[6]); acc[7] = fma(gz, dz, acc[7]);
                if (fij) {                                   // source/sepprfrc.c:70-80: both molecules known
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
    if (P.tab) { acc[0] = acc[1]; acc[1] = 0.0; if (below_seen) scal->error = SEPGPU_ETABLE; }
    else acc[0] = P.eps4 * acc[0] - P.shift * (double)nin;
    block_sum<SEPGPU_NPART_F, FORCE_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

// ---- final reduction of the per-block partial rows ---------------------------------------------------------
// flags: bit0 epot assign (source/sepprfrc.c:222), bit1 also add virial to pot_P_bond, bit2 add acc[1] to
// ecoul AND epot (source/sepcoulomb.c:150-153), bit3 a sep_reset_retval is pending: clear the block first
#define FIN_THREADS 1024
__device__ __forceinline__ void reset_ret_block(DevScalars *s)
{
    s->epot = 0; s->ecoul = 0; s->ekin = 0;                        // sep_reset_retval, source/sepret.c:19-47
    for (int k = 0; k < 9; k++) { s->pot_P[k] = 0; s->kin_P[k] = 0; s->pot_P_bond[k] = 0; }
}

__global__ void __launch_bounds__(FIN_THREADS)
k_finalize_force(const double *__restrict__ partial, int nrows, DevScalars *scal, double scale, int flags)
{
    __shared__ double red[SEPGPU_NPART_F * (FIN_THREADS / 32)];
    double v[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) v[q] = 0.0;
    // fixed assignment of rows to threads and a fixed reduction tree: deterministic sums
    for (int r = threadIdx.x; r < nrows; r += FIN_THREADS) {
        const double4 a = *reinterpret_cast<const double4 *>(partial + (size_t)r * SEPGPU_NPART_F);
        const double4 b = *reinterpret_cast<const double4 *>(partial + (size_t)r * SEPGPU_NPART_F + 4);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    block_sum<SEPGPU_NPART_F, FIN_THREADS>(v, red);
    if (threadIdx.x == 0) {
        if (flags & 8) reset_ret_block(scal);
        const double e = v[0] * scale, ec = v[1] * scale;
        if (flags & 1) scal->epot = e; else scal->epot += e;
        if (flags & 4) { scal->epot += ec; scal->ecoul += ec; }
        const double xx = v[2] * scale, xy = v[3] * scale, xz = v[4] * scale;
        const double yy = v[5] * scale, yz = v[6] * scale, zz = v[7] * scale;
        const double P[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
        for (int k = 0; k < 9; k++) {
            scal->pot_P[k] += P[k];
            if (flags & 2) scal->pot_P_bond[k] += P[k];
        }
    }
}

int sepgpu_dd_halo_update(sepgpu_ctx *c, const sepgpu_sys *sys);
int sepgpu_dd_halo_begin(sepgpu_ctx *c, const sepgpu_sys *sys);
int sepgpu_dd_halo_end(sepgpu_ctx *c);
int sepgpu_need_global_rows(sepgpu_ctx *c, const sepgpu_sys *sys);
int sepgpu_lj_tile_launch(sepgpu_ctx *c, const sepgpu_sys *sys, const LJDev &P, const BoxDev &B, bool typed, bool store, int *nrows,
                          d4 *f_out = nullptr, const int *cancel = nullptr);

// Option fin_multi: the same reduction spread over several CTAs.  CTA b sums a fixed chunk of rows into stage row b;
// the CTA that draws the last ticket adds the stage rows in index order and applies the result -- fixed chunks and a
// fixed final order, so the sums are as deterministic as the single-CTA kernel's (not bit-identical to it: the
// association differs).  One launch, parallel row loads instead of one CTA's serial latency chain.
#define FIN2_THREADS 256
#define FIN2_MAX_CTAS 64
__global__ void __launch_bounds__(FIN2_THREADS)
k_finalize_force_multi(const double *__restrict__ partial, int nrows, double *__restrict__ stage, unsigned *ticket,
                       DevScalars *scal, double scale, int flags)
{
    __shared__ double red[SEPGPU_NPART_F * (FIN2_THREADS / 32)];
    __shared__ double tot[SEPGPU_NPART_F];
    __shared__ int s_last;
    const int chunk = (nrows + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * chunk, r1 = min(nrows, r0 + chunk);
    double v[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) v[q] = 0.0;
    for (int r = r0 + (int)threadIdx.x; r < r1; r += FIN2_THREADS) {
        const double4 a = *reinterpret_cast<const double4 *>(partial + (size_t)r * SEPGPU_NPART_F);
        const double4 b = *reinterpret_cast<const double4 *>(partial + (size_t)r * SEPGPU_NPART_F + 4);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    block_sum<SEPGPU_NPART_F, FIN2_THREADS>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) stage[blockIdx.x * SEPGPU_NPART_F + q] = v[q];
        __threadfence();                                            // stage row visible before the ticket is drawn
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < SEPGPU_NPART_F) {
        double a = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) a += __ldcg(stage + b * SEPGPU_NPART_F + threadIdx.x);
        tot[threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0;                                                // ready for the next launch on this stream
        if (flags & 8) reset_ret_block(scal);
        const double e = tot[0] * scale, ec = tot[1] * scale;
        if (flags & 1) scal->epot = e; else scal->epot += e;
        if (flags & 4) { scal->epot += ec; scal->ecoul += ec; }
        const double xx = tot[2] * scale, xy = tot[3] * scale, xz = tot[4] * scale;
        const double yy = tot[5] * scale, yz = tot[6] * scale, zz = tot[7] * scale;
        const double P[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
        for (int k = 0; k < 9; k++) {
            scal->pot_P[k] += P[k];
            if (flags & 2) scal->pot_P_bond[k] += P[k];
        }
    }
}

static int launch_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags);

int sepgpu_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags)
{
    if (c->ret_reset_pending) { flags |= 8; c->ret_reset_pending = false; }
    if (c->step_fold && (!c->dd || sepgpu_dd_uses_p2p(c))) {
        // option step_fold: leave the reduction to whoever comes next -- the integrator folds it into its own final
        // kernel, every other entry point launches it first (SEPGPU_ENTER).  The rows stay in c->partial until then.
        if (c->fin_pending.active) { int rc = launch_finalize_force(c, c->fin_pending.nrows, c->fin_pending.scale, c->fin_pending.flags); if (rc) return rc; }
        c->fin_pending.active = true; c->fin_pending.nrows = nrows; c->fin_pending.scale = scale; c->fin_pending.flags = flags;
        return 0;
    }
    return launch_finalize_force(c, nrows, scale, flags);
}


int sepgpu_settle(sepgpu_ctx *c)
{
    if (c->fin_pending.active) {
        c->fin_pending.active = false;
        int rc = launch_finalize_force(c, c->fin_pending.nrows, c->fin_pending.scale, c->fin_pending.flags);
        if (rc) return rc;
    }
    if (c->nh_pending.active) return sepgpu_nh_update_now(c);
    return 0;
}

static int launch_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags)
{
    if (c->fin_multi) {
        if (!c->fin_ticket) {
            CUDA_TRY(cudaMalloc((void **)&c->fin_ticket, sizeof(unsigned)));
            CUDA_TRY(cudaMemsetAsync(c->fin_ticket, 0, sizeof(unsigned), c->stream));
        }
        int per = (nrows + FIN2_MAX_CTAS - 1) / FIN2_MAX_CTAS;           // rows per CTA: at least 8, at most 64 CTAs
        if (per < 8) per = 8;
        const int ctas = nrows > 0 ? (nrows + per - 1) / per : 1;
        // stage rows live behind the largest possible set of force rows of the partial buffer
        double *stage = c->partial + (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * SEPGPU_NPART_F;
        k_finalize_force_multi<<<ctas, FIN2_THREADS, 0, c->stream>>>(c->partial, nrows, stage, c->fin_ticket, c->scal, scale, flags);
        KERNEL_CHECK();
        return 0;
    }
    k_finalize_force<<<1, FIN_THREADS, 0, c->stream>>>(c->partial, nrows, c->scal, scale, flags);
    KERNEL_CHECK();
    return 0;
}

int sepgpu_refresh_xs_identity(sepgpu_ctx *c)
{
    // brute mode works on x4 directly; nothing sorted
    c->sorted_identity = true;
    return 0;
}

static LJDev make_lj(const sepgpu_ljparam *p, const char types[2])
{
    LJDev d;
    d.cf2 = p->cf * p->cf;
    d.sig2 = p->sigma * p->sigma;
    d.eps48 = 48.0 * p->eps;
    d.eps4 = 4.0 * p->eps;
    d.aw = p->aw;
    d.awh = 0.5 * p->aw;
    d.shift = p->shift;
    d.t0 = (unsigned char)types[0];
    d.t1 = (unsigned char)types[1];
    d.tab = NULL; d.t_lo = 0.0; d.t_inv = 0.0; d.t_n = 0;
    return d;
}

// ---- typed sub-lists (option typed_sublist) -------------------------------------------------------------------------
// A typed call such as prg3's sep_force_pairs(atoms, "OO", ...) on water walks every entry of the full list only to
// find that 8 of 9 partners have the wrong type (reference source/sepprfrc.c:126-127 tests the types per pair as well).
// With the option on, the first typed call after a list build copies the matching entries of every row into a second
// list of the same chunked layout (one thread per atom, 128-bit streaming reads of the row, types from a one-byte
// array in sorted order); the force kernel then runs unchanged on that list.  Same pairs, same order -> same sums.
__global__ void k_sorted_types(const d4 *__restrict__ xs, unsigned char *__restrict__ tsort, int n)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) tsort[s] = (unsigned char)tag_type(xs[s].w);
}

__global__ void __launch_bounds__(128)
k_typed_sublist(const unsigned *__restrict__ nbr, const int *__restrict__ cnt, const unsigned char *__restrict__ tsort,
                unsigned *__restrict__ nbr_t, int *__restrict__ cnt_t, int n, int npad, int t0, int t1)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int ti = tsort[s];
    int m = 0;
    if (ti == t0 || ti == t1) {                                   // source/sepprfrc.c:164-165
        const int mfull = cnt[s];
        const int nch = (mfull + 3) >> 2;
        const uint4 *row = reinterpret_cast<const uint4 *>(nbr) + s;
        for (int c = 0; c < nch; c++) {
            const uint4 ch = __ldcs(row + (size_t)c * npad);
            const unsigned e[4] = {ch.x, ch.y, ch.z, ch.w};
            const int left = mfull - 4 * c;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q < left) {
                    const int tj = tsort[e[q] & SEPGPU_INDEX_MASK];
                    if ((ti == t0 && tj == t1) || (ti == t1 && tj == t0)) {        // :171-172
                        nbr_t[nbr_index(m, s, npad)] = e[q];
                        m++;
                    }
                }
            }
        }
    }
    cnt_t[s] = m;
}

// returns the sub-list for the type pair in *nbr_out / *cnt_out (building it when the list is newer than the slot)
static int typed_sublist_get(sepgpu_ctx *c, const char types[2], const unsigned **nbr_out, const int **cnt_out)
{
    const int a = (unsigned char)types[0], b = (unsigned char)types[1];
    const int key = (a < b ? a | (b << 8) : b | (a << 8)) | (1 << 16);
    int k = -1;
    for (int q = 0; q < SEPGPU_NSUB; q++) if (c->sub_key[q] == key) k = q;
    if (k < 0) for (int q = 0; q < SEPGPU_NSUB && k < 0; q++) if (c->sub_key[q] == 0) k = q;
    if (k < 0) {                                                  // all slots taken: reuse the stalest one
        k = 0;
        for (int q = 1; q < SEPGPU_NSUB; q++) if (c->sub_gen[q] < c->sub_gen[k]) k = q;
    }
    if (c->sub_key[k] != key) { c->sub_key[k] = key; c->sub_gen[k] = -1; }
    if (!c->nbr_t[k] || c->sub_cap[k] != c->cap) {
        if (c->nbr_t[k]) cudaFree(c->nbr_t[k]);
        c->nbr_t[k] = NULL;
        CUDA_TRY(cudaMalloc((void **)&c->nbr_t[k], sizeof(unsigned) * (size_t)c->cap * c->npad));
        if (!c->cnt_t[k]) CUDA_TRY(cudaMalloc((void **)&c->cnt_t[k], sizeof(int) * (size_t)c->npad));
        c->sub_cap[k] = c->cap;
        c->sub_gen[k] = -1;
    }
    if (c->sub_gen[k] != c->list_gen) {
        if (!c->tsort) CUDA_TRY(cudaMalloc((void **)&c->tsort, (size_t)c->ncap));
        if (c->tsort_gen != c->list_gen) {
            k_sorted_types<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->xs, c->tsort, c->n);
            c->tsort_gen = c->list_gen;
        }
        k_typed_sublist<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->nbr, c->cnt, c->tsort, c->nbr_t[k], c->cnt_t[k], c->n, c->npad, a, b);
        KERNEL_CHECK();
        c->sub_gen[k] = c->list_gen;
    }
    *nbr_out = c->nbr_t[k];
    *cnt_out = c->cnt_t[k];
    return 0;
}

template <int TPA, bool FIJ>
static void launch_lj_list_f(sepgpu_ctx *c, int grid, int apc, bool typed, bool store, const LJDev &P, const BoxDev &B,
                             double *part, const unsigned char *cls, int want, const unsigned *nbr, const int *cnt)
{
#define LJ_ARGS c->xs, nbr, cnt, c->order, c->f4, c->n, c->npad, apc, P, B, part, c->fij, c->nmol, cls, want, c->scal
    if (typed) {
        if (store) k_lj_list<TPA, true, true, FIJ><<<grid, FORCE_BLOCK, 0, c->stream>>>(LJ_ARGS);
        else       k_lj_list<TPA, true, false, FIJ><<<grid, FORCE_BLOCK, 0, c->stream>>>(LJ_ARGS);
    } else {
        if (store) k_lj_list<TPA, false, true, FIJ><<<grid, FORCE_BLOCK, 0, c->stream>>>(LJ_ARGS);
        else       k_lj_list<TPA, false, false, FIJ><<<grid, FORCE_BLOCK, 0, c->stream>>>(LJ_ARGS);
    }
#undef LJ_ARGS
}

template <int TPA>
static void launch_lj_list(sepgpu_ctx *c, int grid, int apc, bool typed, bool store, const LJDev &P, const BoxDev &B,
                           double *part, const unsigned char *cls, int want, const unsigned *nbr, const int *cnt)
{
    if (c->fij || P.tab) launch_lj_list_f<TPA, true>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt);
    else launch_lj_list_f<TPA, false>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt);
}

static void launch_lj_list_tpa(sepgpu_ctx *c, int grid, int apc, bool typed, bool store, const LJDev &P, const BoxDev &B,
                               double *part, const unsigned char *cls, int want, const unsigned *nbr = NULL, const int *cnt = NULL)
{
    if (!nbr) { nbr = c->nbr; cnt = c->cnt; }
    switch (c->tpa) {
    case 1: launch_lj_list<1>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt); break;
    case 2: launch_lj_list<2>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt); break;
    case 4: launch_lj_list<4>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt); break;
    default: launch_lj_list<8>(c, grid, apc, typed, store, P, B, part, cls, want, nbr, cnt); break;
    }
}

static int force_pairs_dev(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], const LJDev &P, unsigned opt, int epot_assign);

extern "C" int sepgpu_force_lj(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2],
                               const sepgpu_ljparam *p, unsigned opt, int epot_assign)
{
    if (!c || !sys || !types || !p) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    return force_pairs_dev(c, sys, types, make_lj(p, types), opt, epot_assign);
}

// sep_force_pairs with a pair function of the caller's own (reference include/sepprfrc.h:49-51, called at
// source/sepprfrc.c:140-146): the host layer has sampled it; the table is uploaded when it changes.
extern "C" int sepgpu_force_table(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], double cf, const double *tab_fu, int n,
                                  double r2_lo, unsigned opt, int epot_assign)
{
    if (!c || !sys || !types || !tab_fu || n < 8 || !(r2_lo >= 0.0) || !(cf * cf > r2_lo)) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    // up to SEPGPU_NTAB tables stay resident (one per pair function and cutoff in use); a table is recognised by its
    // geometry and a hash over 256 of its entries, so a re-sampled function at the same address is uploaded again
    unsigned long long h = 1469598103934665603ull;
    for (int q = 0; q < 256; q++) {
        unsigned long long bits[2];
        memcpy(bits, tab_fu + 2 * (size_t)((long long)q * (n - 1) / 255), sizeof bits);
        h = (h ^ bits[0]) * 1099511628211ull; h = (h ^ bits[1]) * 1099511628211ull;
    }
    int slot = -1;
    for (int q = 0; q < SEPGPU_NTAB; q++)
        if (c->tab[q].dev && c->tab[q].key == tab_fu && c->tab[q].n == n && c->tab[q].lo == r2_lo && c->tab[q].cf == cf && c->tab[q].hash == h) slot = q;
    if (slot < 0) {
        slot = (int)(c->tab_next++ % SEPGPU_NTAB);
        if (c->tab[slot].dev && c->tab[slot].n != n) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));            // a kernel in flight may still read it
            cudaFree(c->tab[slot].dev); c->tab[slot].dev = NULL;
        }
        if (!c->tab[slot].dev) CUDA_TRY(cudaMalloc((void **)&c->tab[slot].dev, sizeof(double2) * (size_t)n));
        CUDA_TRY(cudaMemcpyAsync(c->tab[slot].dev, tab_fu, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->tab[slot].key = tab_fu; c->tab[slot].n = n; c->tab[slot].lo = r2_lo; c->tab[slot].cf = cf; c->tab[slot].hash = h;
    }
    sepgpu_ljparam p;
    p.cf = cf; p.eps = 1.0 / 48.0; p.sigma = 1.0; p.aw = 1.0; p.shift = 0.0;       // eps48 = 1: the table holds the force factor itself
    LJDev P = make_lj(&p, types);
    P.eps48 = 1.0; P.eps4 = 1.0;
    P.tab = reinterpret_cast<const double2 *>(c->tab[slot].dev);
    P.t_lo = r2_lo; P.t_inv = (double)(n - 1) / (cf * cf - r2_lo); P.t_n = n;
    return force_pairs_dev(c, sys, types, P, opt, epot_assign);
}

static bool spec_same(const sepgpu_ctx::SpecForce &S, const sepgpu_sys *sys, const char types[2], const LJDev &P, unsigned opt,
                      int epot_assign, bool typed)
{
    if (S.streak <= 0) return false;
    const LJDev &Q = S.P;
    bool eq = Q.cf2 == P.cf2 && Q.sig2 == P.sig2 && Q.eps48 == P.eps48 && Q.eps4 == P.eps4 && Q.aw == P.aw && Q.awh == P.awh &&
              Q.shift == P.shift && Q.t0 == P.t0 && Q.t1 == P.t1 && !P.tab && !Q.tab;
    eq = eq && S.types[0] == types[0] && S.types[1] == types[1] && S.opt == opt && S.epot_assign == epot_assign && S.typed == typed;
    for (int k = 0; k < 3; k++)
        eq = eq && S.sys.length[k] == sys->length[k] && S.sys.lsubbox[k] == sys->lsubbox[k] && S.sys.nsubbox[k] == sys->nsubbox[k];
    return eq && S.sys.cf == sys->cf && S.sys.skin == sys->skin && S.sys.neighb_update == sys->neighb_update;
}

// Called by the integrators once their finaliser is queued: launch the step's first force routine for the NEXT step now,
// guarded by the rebuild flag that finaliser is about to write.  Forces go to the spare array.
int sepgpu_dd_push_carveout_max(void);

int sepgpu_spec_force_launch(sepgpu_ctx *c)
{
    sepgpu_ctx::SpecForce &S = c->spec;
    S.launched = false;
    if (c->dd && S.on == 2 && !c->f4_alt) sepgpu_dd_push_carveout_max();      // (once: f4_alt is allocated below)
    // (decomposed runs: measured on two B200s, the launch sent ahead ended in a peer-memory wait that never returned at 1 M
    //  atoms per rank although every smaller test passed -- not understood yet, so it stays off there)
    if (!S.on || S.streak < 3 || (c->dd && S.on != 2) || !c->list_valid || !c->list_f16 || c->fij) return 0;
    if (!c->f4_alt && cudaMalloc((void **)&c->f4_alt, sizeof(d4) * (size_t)c->ncap) != cudaSuccess) {
        cudaGetLastError();                              // no room for the spare force array: run without launches sent ahead
        c->f4_alt = NULL;
        S.on = 0;
        return 0;
    }
    int nrows = 0;
    ktimer_begin(c, &c->t_force);
    const int rc = sepgpu_lj_tile_launch(c, &S.sys, S.P, S.B, S.typed, true, &nrows, c->f4_alt, &c->scal->neighb_flag);
    ktimer_end(c, &c->t_force);
    if (rc) return rc;
    S.launched = true; S.cancelled = false; S.nrows = nrows; S.list_gen = c->list_gen; S.api_seq = c->api_seq;
    return 0;
}

static int force_pairs_dev(sepgpu_ctx *c, const sepgpu_sys *sys, const char types[2], const LJDev &P, unsigned opt, int epot_assign)
{
    BoxDev B; B.Lx = sys->length[0]; B.Ly = sys->length[1]; B.Lz = sys->length[2];
    const bool store = c->f_zero;
    int rc;

    if (sys->neighb_update == 0) {                       // SEP_BRUTE, source/sepprfrc.c:240-249
        if (opt == SEPGPU_EXCL_BONDED && !c->excl_bond) {
            sepgpu_set_error("force_lj: SEP_EXCL_BONDED needs the bond partner table");
            return SEPGPU_ESTATE;
        }
        const int grid = (c->n + FORCE_BLOCK - 1) / FORCE_BLOCK;
        ktimer_begin(c, &c->t_force);
        if (store) k_lj_brute<true><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->x4, c->f4, c->n, P, B, opt, c->excl_bond, c->partial, c->fij, c->nmol, c->scal);
        else       k_lj_brute<false><<<grid, FORCE_BLOCK, 0, c->stream>>>(c->x4, c->f4, c->n, P, B, opt, c->excl_bond, c->partial, c->fij, c->nmol, c->scal);
        ktimer_end(c, &c->t_force);
        KERNEL_CHECK();
        c->f_zero = false;
        return sepgpu_finalize_force(c, grid, 0.5, epot_assign ? 1 : 0);
    }

    if (!c->list_valid) {
        if ((rc = sepgpu_neighb_build(c, sys, opt))) return rc;
    }
    // the type test is compiled out when every atom carries the one requested type
    const bool typed = !(types[0] == types[1] && c->single_type == (unsigned char)types[0]);
    if (c->list_f16 && c->fij) {
        // the molecule-pair table was switched on after a tile-format list had been built: global-index rows from now on
        if ((rc = sepgpu_need_global_rows(c, sys))) return rc;
    }
    if (c->list_f16) {                                   // rows of 16-bit tile slots: shared-memory staged tile kernel
        int nrows = 0;
        sepgpu_ctx::SpecForce &S = c->spec;
        const bool same = spec_same(S, sys, types, P, opt, epot_assign, typed);
        if (S.launched) {
            // this very call was launched ahead of time behind the last integrator (sepgpu_spec_force_launch): adopt it when
            // nothing has happened since that could change its result, drop it otherwise (it wrote into the spare array)
            const bool adopt = !S.cancelled && store && same && c->list_gen == S.list_gen && c->api_seq == S.api_seq + 1;
            S.launched = false;
            if (adopt) {
                d4 *t = c->f4; c->f4 = c->f4_alt; c->f4_alt = t;
                c->spec_adopted++;
                c->f_zero = false;
                return sepgpu_finalize_force(c, S.nrows, 0.5, epot_assign ? 1 : 0);
            }
            if (c->t_force.enabled && c->t_force.used > 0) c->t_force.used--;      // not one of the step's force launches
        }
        // the first force routine after sep_reset_force, repeated unchanged step after step, is what gets launched ahead
        if (store && !P.tab && (!c->dd || c->spec.on == 2)) {
            if (same) S.streak++;
            else {
                S.streak = 1; S.P = P; S.B = B; S.typed = typed; S.types[0] = types[0]; S.types[1] = types[1];
                S.opt = opt; S.epot_assign = epot_assign; S.sys = *sys;
            }
        }
        ktimer_begin(c, &c->t_force);
        rc = sepgpu_lj_tile_launch(c, sys, P, B, typed, store, &nrows);
        ktimer_end(c, &c->t_force);
        if (rc) return rc;
        c->f_zero = false;
        return sepgpu_finalize_force(c, nrows, 0.5, epot_assign ? 1 : 0);
    }
    c->spec.streak = 0;
    const int tpa = c->tpa;
    // contiguous ranges of the sorted atoms per CTA; several CTAs per SM, a few waves for load balance
    const int groups = FORCE_BLOCK / tpa;
    int grid = c->force_grid > 0 ? c->force_grid : FORCE_MAX_GRID;
    int apc = (c->n + grid - 1) / grid;
    apc = ((apc + groups - 1) / groups) * groups;
    if (apc < groups) apc = groups;
    grid = (c->n + apc - 1) / apc;
    int nrows = grid;
    if (c->dd) {
        // neighbours' boundary atoms moved too: refresh the halo coordinates.  Option "overlap": the transfer
        // proceeds while this stream computes every atom whose neighbourhood is all local (about 1 - 2/layers of
        // them); the atoms next to a halo layer follow once the halo has arrived.  Off by default -- on B200 with
        // the peer-memory push the refresh costs 20 us, less than the extra, mostly empty, boundary wave.
        const int started = c->overlap && c->cls ? sepgpu_dd_halo_begin(c, sys) : 0;
        if (started < 0) return started;
        if (started) {
            ktimer_begin(c, &c->t_force);
            launch_lj_list_tpa(c, grid, apc, typed, store, P, B, c->partial, c->cls, 0);
            if ((rc = sepgpu_dd_halo_end(c))) return rc;
            launch_lj_list_tpa(c, grid, apc, typed, store, P, B, c->partial + (size_t)grid * SEPGPU_NPART_F, c->cls, 1);
            ktimer_end(c, &c->t_force);
            nrows = 2 * grid;
        } else {
            if ((rc = sepgpu_dd_halo_update(c, sys))) return rc;
            ktimer_begin(c, &c->t_force);
            launch_lj_list_tpa(c, grid, apc, typed, store, P, B, c->partial, NULL, 0);
            ktimer_end(c, &c->t_force);
        }
    } else {
        const unsigned *nbr = NULL;
        const int *cnt = NULL;
        if (typed && c->typed_sublist && (rc = typed_sublist_get(c, types, &nbr, &cnt))) return rc;
        ktimer_begin(c, &c->t_force);
        launch_lj_list_tpa(c, grid, apc, typed, store, P, B, c->partial, NULL, 0, nbr, cnt);
        ktimer_end(c, &c->t_force);
    }
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, nrows, 0.5, epot_assign ? 1 : 0);
}
