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
int sepgpu_dd_before_positions_change(sepgpu_ctx *c);

// One listed pair.  Coordinates arrive divided by sigma, so 1/r^2 needs no rescaling; 48 eps / sigma is applied once per
// atom.  19 FP64 instructions: 3 sub, 3 for r^2, 3 for 1/r^2 (MUFU seed + one third-order step), 2 for its cube,
// 3 for the force factor, 2 for the energy, 3 force accumulations.  The cutoff test is a 64-bit INTEGER compare of
// the bit patterns (both sides are non-negative doubles) and selects r^2 itself: an out-of-range (or wrong-type, or
// padding) partner continues with r^2 = 1e300, whose inverse cube underflows to exactly zero -- no further selects.
// No per-pair virial: with a full list  sum g (x) d = 2 sum_i F_i (x) x_i - sum over boundary-crossing pairs g (x) S
// (see lj_pair in sepgpu_force.cu); IMAGE tiles (those that touch a face of the box) add the second term.
template <bool TYPED, bool IMAGE, bool TABLE>
__device__ __forceinline__ void lj_tile_pair(double xi, double yi, double zi, int ti, const double2 *__restrict__ XY,
                                             const double *__restrict__ Z, const unsigned char *__restrict__ CODE,
                                             const unsigned char *__restrict__ TYPE, const double *__restrict__ SHIFT,
                                             unsigned e, const LJDev &P, double &fx, double &fy, double &fz, double &u, int &nin, double *v)
{
    const unsigned j = IMAGE ? (e & TILE_SLOT_MASK) : e;
    const double2 a = XY[j];
    const double z = Z[j];
    const double dx = xi - a.x, dy = yi - a.y, dz = zi - z;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    bool in = __double_as_longlong(r2) < __double_as_longlong(P.cf2);
    if (TYPED) {
        const int tj = TYPE[j];
        in = in && ((ti == P.t0 && tj == P.t1) || (ti == P.t1 && tj == P.t0));       // source/sepprfrc.c:126-127
    }
    double f;
    if (TABLE) {                                                     // user pair function, sampled by the host layer
        bool below;
        const double2 fu = table_eval(P, in ? r2 : P.cf2, below);
        if (in && below) nin |= 0x40000000;                          // closer than the table reaches: reported by the host
        f = in ? fu.x : 0.0;
        u += in ? fu.y : 0.0;
    } else {
        r2 = in ? r2 : 1e300;
        nin += in ? 1 : 0;
        const double q = fast_rcp3(r2);
        const double b = q * q * q;
        f = b * (b - P.awh) * q;                                     // source/sepmisc.c:139, sepprfrc.c:888 (/ 48 eps)
        u = fma(b, b - P.aw, u);
    }
    fx = fma(f, dx, fx); fy = fma(f, dy, fy); fz = fma(f, dz, fz);
    if (IMAGE) {
        if (e & TILE_SLOT_IMAGE) {                                   // boundary-crossing pair: - g (x) S
            const double *sh = SHIFT + 3 * CODE[j];                  // -S of the partner's image, from the CTA's 27-entry table
            const double g = P.eps48 * f;                            // (eps48 carries the 1/sigma of the scaled coordinates)
            virial_add(v, g * dx, g * dy, g * dz, sh[0], sh[1], sh[2]);
        }
    }
}

// the whole row of one atom
template <bool TYPED, bool IMAGE, bool TABLE>
__device__ __forceinline__ void lj_tile_row(double xi, double yi, double zi, int ti, int m, const uint4 *__restrict__ row, int npad,
                                            const double2 *__restrict__ XY, const double *__restrict__ Z,
                                            const unsigned char *__restrict__ CODE, const unsigned char *__restrict__ TYPE,
                                            const double *__restrict__ SHIFT, const LJDev &P, double &fx, double &fy, double &fz, double &u, int &nin, double *v)
{
    // rows are padded to whole chunks of 8 with the tile's far-away pad slot: no tail handling
    const int nch = (m + 7) >> 3;
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (nch > 0) cur = __ldcs(row);
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (c + 1 < nch) nxt = __ldcs(row + (size_t)(c + 1) * npad);
#define LJT_PAIR(E) lj_tile_pair<TYPED, IMAGE, TABLE>(xi, yi, zi, ti, XY, Z, CODE, TYPE, SHIFT, (E), P, fx, fy, fz, u, nin, v)
        LJT_PAIR(cur.x & 0xffffu); LJT_PAIR(cur.x >> 16);
        LJT_PAIR(cur.y & 0xffffu); LJT_PAIR(cur.y >> 16);
        LJT_PAIR(cur.z & 0xffffu); LJT_PAIR(cur.z >> 16);
        LJT_PAIR(cur.w & 0xffffu); LJT_PAIR(cur.w >> 16);
#undef LJT_PAIR
        cur = nxt;
    }
}


// STORE: first force kernel after sep_reset_force -> plain store instead of read-modify-write.
// tile_hdr / tile_src: the staging tables the list builder left behind (per tile: home range, staged count, image flag;
// per slot: sorted index | image code), so that the kernel starts copying at once -- no cell arithmetic here.
// MINB: CTAs per SM the register budget is cut for (3: 72 registers, 4: 56)
// TABLE: the pair function is a table sampled from a user callback (sep_force_pairs with a function of the caller's own)
template <bool TYPED, bool STORE, int MINB, bool TABLE>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
k_lj_tile(const d4 *__restrict__ xs, const uint4 *__restrict__ nbr, const int *__restrict__ cnt,
          const int *__restrict__ order, const int4 *__restrict__ tile_hdr, const unsigned *__restrict__ tile_src,
          d4 *__restrict__ f4, int stride, int npad, int stage_cap, LJDev P, BoxDev B, double isig,
          double *__restrict__ partial, HaloArgs H, DevScalars *scal)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nslot = stage_cap + TILE_PAD;
    double2 *XY = reinterpret_cast<double2 *>(smem_raw);                          // [nslot]
    double *Z = reinterpret_cast<double *>(XY + nslot);                           // [nslot]
    unsigned char *CODE = reinterpret_cast<unsigned char *>(Z + nslot);           // [nslot]
    unsigned char *TYPE = CODE + nslot;                                           // [nslot] (typed calls only)
    __shared__ double red[SEPGPU_NPART_F * (TILE_THREADS / 32)];
    __shared__ double SHIFT[27 * 3];                                              // -S per image code

    // The first round of staging entries is requested together with the header: they lie inside this tile's stride of
    // tile_src whatever `total` turns out to be, and the two loads then share one trip to memory instead of two.
    constexpr int UN = 4, UN2 = 2 * UN;
    // decomposed runs start one brick layer up, so that the tiles that have to wait for the neighbours' coordinates run
    // at the end of the launch, when those have long arrived, instead of spinning on an SM at its start
    int tile = (int)blockIdx.x + H.rot;
    if (tile >= (int)gridDim.x) tile -= (int)gridDim.x;
    const unsigned *src = tile_src + (size_t)tile * stride;
    unsigned e[UN2];
#pragma unroll
    for (int u = 0; u < UN2; u++) {
        const int q = threadIdx.x + u * TILE_THREADS;
        e[u] = q < stride ? __ldg(src + q) : 0xffffffffu;
    }
    const int4 hdr = tile_hdr[tile];
    const int a0 = hdr.x, nhome = hdr.y, total = hdr.z;
    const bool image = (hdr.w & 1) != 0;
    // slab runs on the peer-memory path: the neighbours store their boundary coordinates straight into this rank's
    // receive buffers every step and raise a flag.  Only tiles next to a halo layer wait for it -- every other CTA of
    // this launch computes while the transfer is still on its way -- and they read the halo atoms from those buffers.
    const bool halo = (hdr.w & 2) != 0 && H.seq != 0;
    // speculative launch (sepgpu_spec_force_launch): the integrator's finaliser has just asked for a list rebuild -> nothing to do
    const bool cancelled = H.cancel != nullptr && *reinterpret_cast<const volatile int *>(H.cancel) != 0;
    if (cancelled || nhome == 0 || total > stage_cap) {              // (the last cannot happen: the builder sized stage_cap)
        if (threadIdx.x < SEPGPU_NPART_F) partial[tile * SEPGPU_NPART_F + threadIdx.x] = 0.0;
        return;
    }
    if (halo) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            unsigned long long f0, f1;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f0) : "l"(H.flags) : "memory");
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f1) : "l"(H.flags + 1) : "memory");
                if (f0 >= H.seq && f1 >= H.seq) break;
                if (clock64() - t0 > SEPGPU_SPIN_LIMIT) { scal->error = SEPGPU_ENCCL; scal->error_where = 4; break; }
                __nanosleep(100);
            } while (true);
        }
        __syncthreads();
    }
    // ---- stage the current coordinates of every candidate of the tile once; eight entries per thread and round, their
    // coordinates in two groups of four independent loads; image shift and 1/sigma on the way ----
    {
        for (int base = threadIdx.x; base < total; base += TILE_THREADS * UN2) {
            if (base != (int)threadIdx.x) {
#pragma unroll
                for (int u = 0; u < UN2; u++) {
                    const int q = base + u * TILE_THREADS;
                    e[u] = q < total ? __ldg(src + q) : 0xffffffffu;
                }
            }
            // (all eight coordinates in one round, as separate (x,y) / z loads, was measured slower: 0.234 against 0.231 ms)
#pragma unroll
            for (int g = 0; g < 2; g++) {
                d4 p[UN];
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    const int q = base + (g * UN + u) * TILE_THREADS;
                    if (q >= total) e[g * UN + u] = 0xffffffffu;
                    if (e[g * UN + u] == 0xffffffffu) continue;
                    const int j = (int)(e[g * UN + u] & SEPGPU_INDEX_MASK);
                    int k = -1;
                    if (halo) k = order[j] - H.n_own;                 // >= 0: a halo atom, k-th in arrival order
                    if (k >= 0) {
                        const double2 *sp = reinterpret_cast<const double2 *>(k < H.n0 ? H.in0 + k : H.in1 + (k - H.n0));
                        const double2 a = __ldcg(sp), b = __ldcg(sp + 1);       // written by another GPU: never through L1
                        p[u].x = a.x; p[u].y = a.y; p[u].z = b.x; p[u].w = b.y;
                    } else {
                        p[u] = xs[j];
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    const unsigned ee = e[g * UN + u];
                    if (ee != 0xffffffffu) {
                        const int q = base + (g * UN + u) * TILE_THREADS;
                        const int code = (int)(ee >> SEPGPU_SHIFT_BITS);
                        double sx = 0.0, sy = 0.0, sz = 0.0;
                        if (code != 13) apply_image(code, B, sx, sy, sz);               // s = -S
                        XY[q] = make_double2((p[u].x - sx) * isig, (p[u].y - sy) * isig);
                        Z[q] = (p[u].z - sz) * isig;
                        if (image) CODE[q] = (unsigned char)code;
                        if (TYPED) TYPE[q] = (unsigned char)tag_type(p[u].w);
                    }
                }
            }
        }
        if (threadIdx.x < 27) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
            apply_image((int)threadIdx.x, B, sx, sy, sz);
            SHIFT[3 * threadIdx.x] = sx; SHIFT[3 * threadIdx.x + 1] = sy; SHIFT[3 * threadIdx.x + 2] = sz;
        }
        if (threadIdx.x < TILE_PAD) {                                // pad slots: far away, never in range
            XY[total + threadIdx.x] = make_double2(1e9, 1e9);
            Z[total + threadIdx.x] = 1e9;
            CODE[total + threadIdx.x] = 13;
            if (TYPED) TYPE[total + threadIdx.x] = 0;
        }
    }
    __syncthreads();

    double tot[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) tot[q] = 0.0;
    // one pass per TILE_THREADS home atoms (one pass unless the tile is unusually full)
    for (int base = 0; base < nhome; base += TILE_THREADS) {
        double acc[SEPGPU_NPART_F];
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
        const int ab = base + threadIdx.x;
        if (ab < nhome) {
            const int s = a0 + ab;
            const d4 pi = xs[s];
            int m = cnt[s];
            int ti = 0;
            if (TYPED) {
                ti = tag_type(pi.w);
                if (ti != P.t0 && ti != P.t1) m = 0;                 // source/sepprfrc.c:119-120
            }
            const double xi = pi.x * isig, yi = pi.y * isig, zi = pi.z * isig;
            double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0;
            int nin = 0;
            if (image) lj_tile_row<TYPED, true, TABLE>(xi, yi, zi, ti, m, nbr + s, npad, XY, Z, CODE, TYPE, SHIFT, P, fx, fy, fz, u, nin, acc + 2);
            else       lj_tile_row<TYPED, false, TABLE>(xi, yi, zi, ti, m, nbr + s, npad, XY, Z, CODE, TYPE, SHIFT, P, fx, fy, fz, u, nin, acc + 2);
            const int i = order[s];
            fx *= P.eps48; fy *= P.eps48; fz *= P.eps48;
            if (STORE) {
                d4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0.0;
                f4[i] = o;
            } else {
                d4 o = f4[i];
                o.x += fx; o.y += fy; o.z += fz;
                f4[i] = o;
            }
            // per-atom part of the virial: 2 F_i (x) x_i, upper triangle
            virial_add(acc + 2, fx + fx, fy + fy, fz + fz, pi.x, pi.y, pi.z);
            if (TABLE) {
                if (nin & 0x40000000) scal->error = SEPGPU_ETABLE;
                acc[0] = u;
            } else {
                acc[0] = P.eps4 * u - P.shift * (double)nin;
            }
        }
        block_sum<SEPGPU_NPART_F, TILE_THREADS>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < SEPGPU_NPART_F; q++) tot[q] += acc[q];
        }
        __syncthreads();                                             // red is reused by the next pass
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[tile * SEPGPU_NPART_F + q] = tot[q];
    }
}

size_t sepgpu_tile_force_smem(int stage_cap)
{
    return (size_t)(stage_cap + TILE_PAD) * (sizeof(double2) + sizeof(double) + 2);
}

// Launches the tile kernel on the context's current tile-format list; returns the number of partial rows.
int sepgpu_lj_tile_launch(sepgpu_ctx *c, const sepgpu_sys *sys, const LJDev &P, const BoxDev &B, bool typed, bool store, int *nrows,
                          d4 *f_out, const int *cancel)
{
    // decomposed run: peer-memory path -> the kernel waits for and reads the neighbours' coordinates itself;
    // otherwise refresh the halo entries of xs first
    HaloArgs H;
    memset(&H, 0, sizeof H);
    if (c->dd) {
        int rh = sepgpu_dd_halo_args(c, sys, &H);
        if (rh < 0) return rh;
        if (rh == 0 && (rh = sepgpu_dd_halo_update(c, sys))) return rh;
        // launch sent ahead in a decomposed run (spec_force=2, experimental): this grid and my push become runnable on the
        // same event; ordering the grid behind the push means my neighbours' waiting tiles never depend on a push that
        // sits behind my own waiting tiles (docs/ROUND_NOTES.md, open issue; not verified on hardware)
        if (cancel && c->spec.on == 2 && (rh = sepgpu_dd_before_positions_change(c))) return rh;
    }
    H.cancel = cancel;
    d4 *const f4 = f_out ? f_out : c->f4;
    const int grid = c->tile_count;
    if (c->dd && H.seq != 0) {
        const CellGrid &G = c->tile_grid;
        const int per_layer = G.nbx * G.nby * (BRICK_YZ * BRICK_YZ / c->tile_R);      // tiles in one layer of bricks
        if (per_layer > 0 && per_layer < grid) H.rot = per_layer;
    }
    const int stage_cap = c->tile_stage_used;
    const size_t smem = sepgpu_tile_force_smem(stage_cap);
    // the kernel works on coordinates divided by sigma: cutoff and force prefactor follow
    const double sigma = P.tab ? 1.0 : sqrt(P.sig2), isig = 1.0 / sigma;
    LJDev Ps = P;
    Ps.cf2 = P.cf2 / (sigma * sigma); Ps.sig2 = 1.0; Ps.eps48 = P.eps48 * isig;
#define LJT_LAUNCH3(TY, ST, MB)                                                                                                  \
    do {                                                                                                                         \
        if (P.tab) {                                                                                                             \
            CUDA_TRY(cudaFuncSetAttribute(k_lj_tile<TY, ST, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            k_lj_tile<TY, ST, 3, true><<<grid, TILE_THREADS, smem, c->stream>>>(c->xs, reinterpret_cast<const uint4 *>(c->nbr), c->cnt, \
                c->order, c->tile_hdr, c->tile_src, f4, c->tile_stride, c->npad, stage_cap, Ps, B, isig, c->partial, H, c->scal);       \
            break;                                                                                                               \
        }                                                                                                                        \
        CUDA_TRY(cudaFuncSetAttribute(k_lj_tile<TY, ST, MB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        k_lj_tile<TY, ST, MB, false><<<grid, TILE_THREADS, smem, c->stream>>>(c->xs, reinterpret_cast<const uint4 *>(c->nbr), c->cnt,   \
            c->order, c->tile_hdr, c->tile_src, f4, c->tile_stride, c->npad, stage_cap, Ps, B, isig, c->partial, H, c->scal);       \
    } while (0)
#define LJT_LAUNCH(TY, ST) LJT_LAUNCH3(TY, ST, 3)        // 3 CTAs per SM (72 registers); 4 spill and were measured slower
    if (typed) { if (store) LJT_LAUNCH(true, true); else LJT_LAUNCH(true, false); }
    else { if (store) LJT_LAUNCH(false, true); else LJT_LAUNCH(false, false); }
#undef LJT_LAUNCH
#undef LJT_LAUNCH3
    KERNEL_CHECK();
    *nrows = grid;
    return 0;
}
