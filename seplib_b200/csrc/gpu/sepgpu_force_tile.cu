// sepgpu_force_tile.cu -- Lennard-Jones pair forces from shared-memory staged neighbour tiles.
//
// Stand-in for sep_force_pair_neighb / sep_lj_pair_neighb (reference source/sepprfrc.c:94-224, 782-923),
// the default list kernel of this library.
//
// The list kernel of round 1 (k_lj_list, sepgpu_force.cu) gathered every listed neighbour with one 32-byte
// load from global memory per lane; the 32 lanes of such a gather touch ~22 different 128-byte lines and
// L1 serves one line per wavefront, so the kernel ran at the speed of the L1 tag stage (ncu: data pipe 79 %,
// 19 wavefronts per request) with the FP64 pipe half idle.  Here one CTA owns one TILE of the cell grid
// (sepgpu_tile.cuh): it copies the current coordinates of the tile's (bx+2) x (R+2) x 3 candidate cells --
// each a contiguous run of the cell-sorted array, read with coalesced 256-bit loads, shifted to the periodic
// image the tile sees -- into shared memory ONCE, and every thread then walks its atom's row of 16-bit SLOTS
// (positions in that staged order, written by k_build_tile) and fetches its neighbours with one LDS.128
// (x, y) + one LDS.64 (z): no global gathers, no per-pair image arithmetic, 2 bytes of list per pair.
// One thread per atom, register accumulation, one 256-bit store per atom, no atomics; energy and virial go
// warp shuffle -> block -> one partial row per CTA -> fixed-order final sum (deterministic).
#include "sepgpu_internal.cuh"
#include "sepgpu_pair.cuh"
#include "sepgpu_tile.cuh"

int sepgpu_dd_halo_update(sepgpu_ctx *c, const sepgpu_sys *sys);

// Two listed pairs as ONE straight-line block, so that the scheduler interleaves their dependent FP64 chains.
// The pair body is the 20-instruction body of lj_pair2 (sepgpu_force.cu): integer cutoff compare, 1/r^2 from
// the MUFU seed and one third-order step, 48 eps applied per atom, no per-pair virial (see the note there:
// sum g (x) d = 2 sum_i F_i (x) x_i - sum over boundary-crossing pairs g (x) S).
// valid1 == false: the second slot is past the end of the row; it re-reads the first partner and is masked out.
template <bool TYPED>
__device__ __forceinline__ void lj_tile_pair2(double xi, double yi, double zi, int ti, const double2 *__restrict__ XY,
                                              const double *__restrict__ Z, const unsigned char *__restrict__ CODE,
                                              const unsigned char *__restrict__ TYPE, unsigned e0, unsigned e1, bool valid1,
                                              const LJDev &P, const BoxDev &B, PairAcc &A)
{
    const unsigned j0 = e0 & TILE_SLOT_MASK, j1 = valid1 ? (e1 & TILE_SLOT_MASK) : j0;
    const double2 a0 = XY[j0], a1 = XY[j1];
    const double z0 = Z[j0], z1 = Z[j1];
    const double dx0 = xi - a0.x, dy0 = yi - a0.y, dz0 = zi - z0;
    const double dx1 = xi - a1.x, dy1 = yi - a1.y, dz1 = zi - z1;
    const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
    const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
    bool in0 = __double_as_longlong(r20) < __double_as_longlong(P.cf2);
    bool in1 = (__double_as_longlong(r21) < __double_as_longlong(P.cf2)) && valid1;
    if (TYPED) {
        const int t0 = TYPE[j0], t1 = TYPE[j1];
        in0 = in0 && ((ti == P.t0 && t0 == P.t1) || (ti == P.t1 && t0 == P.t0));     // source/sepprfrc.c:126-127
        in1 = in1 && ((ti == P.t0 && t1 == P.t1) || (ti == P.t1 && t1 == P.t0));
    }
    const double q0 = P.sig2 * fast_rcp3(r20), q1 = P.sig2 * fast_rcp3(r21);
    double b0 = q0 * q0 * q0, b1 = q1 * q1 * q1;
    double f0 = b0 * (b0 - P.awh) * q0, f1 = b1 * (b1 - P.awh) * q1;   // source/sepmisc.c:139, sepprfrc.c:888 (/ 48 eps)
    const double u0 = b0 - P.aw, u1 = b1 - P.aw;
    f0 = in0 ? f0 : 0.0; f1 = in1 ? f1 : 0.0;
    b0 = in0 ? b0 : 0.0; b1 = in1 ? b1 : 0.0;
    A.fx = fma(f0, dx0, A.fx); A.fy = fma(f0, dy0, A.fy); A.fz = fma(f0, dz0, A.fz);
    A.u = fma(b0, u0, A.u);
    A.fx = fma(f1, dx1, A.fx); A.fy = fma(f1, dy1, A.fy); A.fz = fma(f1, dz1, A.fz);
    A.u = fma(b1, u1, A.u);
    A.nin += (in0 ? 1 : 0) + (in1 ? 1 : 0);
    if ((e0 | (valid1 ? e1 : 0u)) & TILE_SLOT_IMAGE) {              // boundary-crossing pairs: - g (x) S
        double sx = 0.0, sy = 0.0, sz = 0.0;
        apply_image(CODE[j0], B, sx, sy, sz);                       // s = -S; code 13 shifts by zero
        double g = P.eps48 * f0;
        virial_add(A.v, g * dx0, g * dy0, g * dz0, sx, sy, sz);
        sx = sy = sz = 0.0;
        apply_image(CODE[j1], B, sx, sy, sz);
        g = P.eps48 * f1;
        virial_add(A.v, g * dx1, g * dy1, g * dz1, sx, sy, sz);
    }
}

#define LJT_MIN_CTAS 3

// STORE: first force kernel after sep_reset_force -> plain store instead of read-modify-write.
template <bool TYPED, bool STORE>
__global__ void __launch_bounds__(TILE_THREADS, LJT_MIN_CTAS)
k_lj_tile(const d4 *__restrict__ xs, const uint4 *__restrict__ nbr, const int *__restrict__ cnt,
          const int *__restrict__ order, const int *__restrict__ cell_start, d4 *__restrict__ f4,
          CellGrid G, int R, int npad, int stage_cap, LJDev P, BoxDev B, double *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nslot = stage_cap + TILE_PAD;
    double2 *XY = reinterpret_cast<double2 *>(smem_raw);                          // [nslot]
    double *Z = reinterpret_cast<double *>(XY + nslot);                           // [nslot]
    unsigned char *CODE = reinterpret_cast<unsigned char *>(Z + nslot);           // [nslot]
    unsigned char *TYPE = CODE + nslot;                                           // [nslot] (typed calls only)
    __shared__ TileLayout T;
    __shared__ double red[SEPGPU_NPART_F * (TILE_THREADS / 32)];

    bool live = true;
    if (G.dd) {                                                      // slab run: halo layers own no forces
        int x0, cy0, cz;
        key_cell(blockIdx.x * R * G.bx, G, x0, cy0, cz);
        if (cz == 0 || cz == G.nz - 1) live = false;
    }
    if (live) live = tile_layout(T, G, R, cell_start);
    if (!live || T.total > stage_cap) {                              // (the second cannot happen: the builder sized stage_cap)
        if (threadIdx.x < SEPGPU_NPART_F) partial[blockIdx.x * SEPGPU_NPART_F + threadIdx.x] = 0.0;
        return;
    }
    // ---- stage the current coordinates of every candidate of the tile once ----
    const int total = T.total;
    for (int q = threadIdx.x; q < total + TILE_PAD; q += TILE_THREADS) {
        double2 a = make_double2(1e9, 1e9);                          // pad slots: far away, never in range
        double z = 1e9;
        unsigned code = 13, type = 0;
        if (q < total) {
            const int c = tile_cell_of_slot(T, q);
            const d4 p = xs[T.beg[c] + (q - T.off[c])];
            code = T.code[c];
            a.x = p.x; a.y = p.y; z = p.z;
            if (code != 13) {
                const int wx = (int)(code % 3u) - 1, wy = (int)((code / 3u) % 3u) - 1, wz = (int)(code / 9u) - 1;
                a.x += wx * B.Lx; a.y += wy * B.Ly; z += wz * B.Lz;
            }
            if (TYPED) type = (unsigned)tag_type(p.w);
        }
        XY[q] = a; Z[q] = z; CODE[q] = (unsigned char)code;
        if (TYPED) TYPE[q] = (unsigned char)type;
    }
    __syncthreads();

    PairAcc A;
    A.u = 0.0;
    A.nin = 0;
#pragma unroll
    for (int q = 0; q < 6; q++) A.v[q] = 0.0;
    const int a0 = T.a0, nhome = T.nhome;
    for (int ab = threadIdx.x; ab < nhome; ab += TILE_THREADS) {
        const int s = a0 + ab;
        const d4 pi = xs[s];
        int m = cnt[s];
        int ti = 0;
        if (TYPED) {
            ti = tag_type(pi.w);
            if (ti != P.t0 && ti != P.t1) m = 0;                     // source/sepprfrc.c:119-120
        }
        A.fx = A.fy = A.fz = 0.0;
        const int nch = (m + 7) >> 3;
        const uint4 *row = nbr + s;
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (nch > 0) cur = __ldcs(row);
        for (int c = 0; c < nch; c++) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c + 1 < nch) nxt = __ldcs(row + (size_t)(c + 1) * npad);
            const int left = m - 8 * c;                              // >= 1 valid entries in this chunk
            lj_tile_pair2<TYPED>(pi.x, pi.y, pi.z, ti, XY, Z, CODE, TYPE, cur.x & 0xffffu, cur.x >> 16, left > 1, P, B, A);
            if (left > 2) lj_tile_pair2<TYPED>(pi.x, pi.y, pi.z, ti, XY, Z, CODE, TYPE, cur.y & 0xffffu, cur.y >> 16, left > 3, P, B, A);
            if (left > 4) lj_tile_pair2<TYPED>(pi.x, pi.y, pi.z, ti, XY, Z, CODE, TYPE, cur.z & 0xffffu, cur.z >> 16, left > 5, P, B, A);
            if (left > 6) lj_tile_pair2<TYPED>(pi.x, pi.y, pi.z, ti, XY, Z, CODE, TYPE, cur.w & 0xffffu, cur.w >> 16, left > 7, P, B, A);
            cur = nxt;
        }
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
        // per-atom part of the virial: 2 F_i (x) x_i, upper triangle
        virial_add(A.v, fx + fx, fy + fy, fz + fz, pi.x, pi.y, pi.z);
    }
    double acc[SEPGPU_NPART_F];
    acc[0] = P.eps4 * A.u - P.shift * (double)A.nin;
    acc[1] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) acc[2 + q] = A.v[q];
    block_sum<SEPGPU_NPART_F, TILE_THREADS>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
    }
}

size_t sepgpu_tile_force_smem(int stage_cap)
{
    return (size_t)(stage_cap + TILE_PAD) * (sizeof(double2) + sizeof(double) + 2);
}

// Launches the tile kernel on the context's current tile-format list; returns the number of partial rows.
int sepgpu_lj_tile_launch(sepgpu_ctx *c, const LJDev &P, const BoxDev &B, bool typed, bool store, int *nrows)
{
    const CellGrid G = c->tile_grid;
    const int R = c->tile_R;
    const int grid = c->tile_count;
    const int stage_cap = c->tile_stage_used;
    const size_t smem = sepgpu_tile_force_smem(stage_cap);
#define LJT_LAUNCH(TY, ST)                                                                                                   \
    do {                                                                                                                     \
        CUDA_TRY(cudaFuncSetAttribute(k_lj_tile<TY, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        k_lj_tile<TY, ST><<<grid, TILE_THREADS, smem, c->stream>>>(c->xs, reinterpret_cast<const uint4 *>(c->nbr), c->cnt,   \
            c->order, c->cell_start, c->f4, G, R, c->npad, stage_cap, P, B, c->partial);                                     \
    } while (0)
    if (typed) { if (store) LJT_LAUNCH(true, true); else LJT_LAUNCH(true, false); }
    else { if (store) LJT_LAUNCH(false, true); else LJT_LAUNCH(false, false); }
#undef LJT_LAUNCH
    KERNEL_CHECK();
    *nrows = grid;
    return 0;
}
