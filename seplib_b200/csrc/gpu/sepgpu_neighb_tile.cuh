// sepgpu_neighb_tile.cuh -- the fast Verlet-list builder (included by sepgpu_neighb.cu after the
// shared definitions BuildParams / pair_exact / excluded).
//
// One CTA owns one TILE (sepgpu_tile.cuh): bx x R home cells whose atoms are contiguous in the cell-sorted
// array; one THREAD owns one home atom.  All (bx+2) x (R+2) x 3 candidate cells of the tile are staged ONCE
// into shared memory as FP32 positions already shifted to the right periodic image.  After a single barrier
// every thread sweeps, for each of the 9 (dy,dz) rows, the 3 cells around its own cell in two passes: a
// branch-free pass that tests 32 candidates into a bit mask, and a pass over the set bits that appends
// entries.  Every candidate is read from HBM once per CTA instead of once per atom, all lanes work on
// different atoms, and there is no barrier inside the sweep.
// Candidates inside the FP32 error band of the cutoff take the exact FP64 test (pair_exact), so the
// resulting pair set equals the reference's bit for bit.
//
// F16 = true : rows of 16-bit tile slots (sepgpu_tile.cuh) for the tile force kernels, 8 per 128-bit chunk.
// F16 = false: rows of 32-bit entries  sorted index | image code << 26, 4 per chunk, for the kernels that
//              gather from global memory (DPD, the molecule-pair table, small grids).
// Rows are assembled in a register window (RowWriter16 / RowWriter32) and leave as whole 128-bit chunks.
#pragma once

#include "sepgpu_tile.cuh"

#include <cuda_pipeline.h>

// bits [max(lo,0), min(hi,32)) of a 32-bit mask
__device__ __forceinline__ unsigned bit_range(int lo, int hi)
{
    lo = lo < 0 ? 0 : lo;
    hi = hi > 32 ? 32 : hi;
    if (hi <= lo) return 0u;
    const unsigned upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return upto & ~((1u << lo) - 1u);
}

// packed FP32 (two lanes of one 64-bit register pair per instruction: FADD2 / FFMA2 on sm_100)
#ifdef SEPGPU_EMU
static inline float2 f2_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 f2_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline int imin3(int a, int b, int c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
static inline unsigned shift_in_sign(unsigned m, float d) { return (m << 1) | (__float_as_uint(d) >> 31); }
#else
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ int imin3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
// (m << 1) | sign(d): one funnel shift
__device__ __forceinline__ unsigned shift_in_sign(unsigned m, float d) { return __funnelshift_l(__float_as_uint(d), m, 1); }
#endif

// 32-bit rows (four entries per 128-bit chunk): the last four entries ride in registers, a whole chunk leaves with one
// store through a pointer that walks the row (no address arithmetic from the atom index at store time)
struct RowWriter32 {
    unsigned v0, v1, v2, v3;
    uint4 *chunk;
    __device__ __forceinline__ void init(unsigned *nbr, int s)
    {
        v0 = v1 = v2 = v3 = 0u;
        chunk = reinterpret_cast<uint4 *>(nbr) + s;
    }
    // entry number `count` of the row
    __device__ __forceinline__ void push(unsigned e, int count, int npad)
    {
        v0 = v1; v1 = v2; v2 = v3; v3 = e;
        if ((count & 3) == 3) { *chunk = make_uint4(v0, v1, v2, v3); chunk += npad; }
    }
    // the last, partial chunk: its entries sit in the upper registers
    __device__ __forceinline__ void finish(int count)
    {
        const int k = count & 3;
        if (k == 0) return;
        if (k == 1) *chunk = make_uint4(v3, 0u, 0u, 0u);
        else if (k == 2) *chunk = make_uint4(v2, v3, 0u, 0u);
        else *chunk = make_uint4(v1, v2, v3, 0u);
    }
};

// 16-bit rows (eight tile slots per chunk): the same scheme, entries shifted in from the top, 16 bits at a time
struct RowWriter16 {
    unsigned v0, v1, v2, v3;
    uint4 *chunk;
    __device__ __forceinline__ void init(unsigned *nbr, int s)
    {
        v0 = v1 = v2 = v3 = 0u;
        chunk = reinterpret_cast<uint4 *>(nbr) + s;
    }
    __device__ __forceinline__ void push(unsigned e, int count, int npad)
    {
        v0 = (v0 >> 16) | (v1 << 16); v1 = (v1 >> 16) | (v2 << 16); v2 = (v2 >> 16) | (v3 << 16);      // (one PRMT each)
        v3 = (v3 >> 16) | (e << 16);
        if ((count & 7) == 7) { *chunk = make_uint4(v0, v1, v2, v3); chunk += npad; }
    }
    // the last, partial chunk is filled up with the tile's far-away pad slot
    __device__ __forceinline__ void finish(int count, unsigned pad, int npad)
    {
        for (int k = count; k & 7; k++) push(pad, k, npad);
    }
};

template <unsigned OPT, bool F16>
#ifndef BUILD_MINB
#define BUILD_MINB 4
#endif
__global__ void __launch_bounds__(TILE_THREADS, BUILD_MINB)
k_build_tile(const d4 *__restrict__ xs, const float4 *__restrict__ xf, const int *__restrict__ order,
             const int *__restrict__ cell_start, const int *__restrict__ excl_bond,
             const int *__restrict__ excl_angle, const int *__restrict__ excl_dihed,
             unsigned *__restrict__ nbr, int *__restrict__ cnt, DevScalars *scal, BuildParams P, int R, int stage_cap,
             int home_cap, int4 *__restrict__ tile_hdr, unsigned *__restrict__ tile_src)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // candidates as four arrays (x, y, z, list entry), so that one 128-bit read brings one coordinate of FOUR candidates
    // for the packed FP32 tests below; each [stage_cap + TILE_PAD], both multiples of 32
    const int SC = stage_cap + TILE_PAD;
    float *candX = reinterpret_cast<float *>(smem_raw);
    float *candY = candX + SC;
    float *candZ = candY + SC;
    unsigned *candW = reinterpret_cast<unsigned *>(candZ + SC);
    int *home_order = reinterpret_cast<int *>(candW + SC);                     // [home_cap] original index of the home atoms
    int *cand_mol = home_order + home_cap;                                     // [SC] (SAME_MOL only)
    __shared__ TileLayout T;
    __shared__ int s_red[3];

    const CellGrid G = P.G;
    {
        int x0, cy0, cz;
        key_cell(blockIdx.x * R * G.bx, G, x0, cy0, cz);
        if (G.dd && x0 < G.nx && cy0 < G.ny && (cz == 0 || cz == G.nz - 1)) {     // halo layer: its atoms own no rows
            const int key0 = blockIdx.x * R * G.bx;
            const int b = cell_start[key0], e = cell_start[key0 + R * G.bx];
            for (int q = b + threadIdx.x; q < e; q += TILE_THREADS) cnt[q] = 0;
            if (threadIdx.x == 0) tile_hdr[blockIdx.x] = make_int4(0, 0, 0, 0);
            return;
        }
    }
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    if (!tile_layout(T, G, R, cell_start)) {
        if (threadIdx.x == 0) tile_hdr[blockIdx.x] = make_int4(0, 0, 0, 0);
        return;
    }
    const int total = T.total;
    if (total > stage_cap || total > TILE_MAX_SLOTS) {             // host grows the staging buffer (or shrinks the tile) and relaunches
        if (threadIdx.x == 0) atomicMax(&scal->stage_needed, total);
        return;
    }
    if (threadIdx.x == 0) {
        atomicMax(&scal->stage_used, total);
        // what the force kernels need to know about this tile: home atoms [a0, a0 + nhome), staged atoms, image flag
        // flags: bit 0 some candidate is a periodic image, bit 1 (slab runs) some candidate is a halo atom
        tile_hdr[blockIdx.x] = make_int4(T.a0, T.nhome, total, T.any_image | ((G.dd && (T.cz == 1 || T.cz == G.nz - 2)) ? 2 : 0));
    }
    // ---- stage every candidate of the tile once: thread q takes slots q, q + 288, ...; four loads in flight per
    // thread, then image shift, list entry, and the scatter into the four arrays ----
    {
        if (threadIdx.x < TILE_PAD) {                                          // padding: never in range
            const int q = total + threadIdx.x;
            candX[q] = 1e18f; candY[q] = 1e18f; candZ[q] = 1e18f; candW[q] = 0u;
        }
        // (the reference-style half-list length below compares original indices inside the own cell)
        if (T.nhome <= home_cap)
            for (int k = threadIdx.x; k < T.nhome; k += TILE_THREADS) home_order[k] = order[T.a0 + k];
        for (int qb = threadIdx.x; qb < total; qb += 4 * TILE_THREADS) {
            float4 f[4];
            unsigned ent[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = qb + u * TILE_THREADS;
                if (q < total) {
                    const int cc = tile_cell_of_slot(T, q);
                    const int j = T.beg[cc] + (q - T.off[cc]);
                    ent[u] = (unsigned)j | ((unsigned)T.code[cc] << SEPGPU_SHIFT_BITS);
                    f[u] = xf[j];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = qb + u * TILE_THREADS;
                if (q < total) {
                    const unsigned code = ent[u] >> SEPGPU_SHIFT_BITS;
                    const int wx = (int)(code % 3u) - 1, wy = (int)((code / 3u) % 3u) - 1, wz = (int)(code / 9u) - 1;
                    if (OPT == SEPGPU_EXCL_SAME_MOL) {                  // atoms outside molecules (-1) never match anyone
                        const int mj = __float_as_int(f[u].w);
                        cand_mol[q] = mj == -1 ? 0x3fffffff : mj;
                    }
                    candX[q] = f[u].x + wx * P.fLx; candY[q] = f[u].y + wy * P.fLy; candZ[q] = f[u].z + wz * P.fLz;
                    candW[q] = ent[u];
                    // the staging order of this tile, for the force kernels: slot -> sorted index | image code
                    tile_src[(size_t)blockIdx.x * stage_cap + q] = ent[u];
                }
            }
        }
    }
    __syncthreads();

    const int ncc = T.ncc, nry = T.nry, a0 = T.a0, nhome = T.nhome, nh = R * G.bx;
    int blk_max = 0, blk_half = 0, blk_sum = 0;
    for (int ab = 0; ab < nhome; ab += TILE_THREADS) {
        const int s = a0 + ab + threadIdx.x;
        if (s < a0 + nhome) {
            const int h = tile_home_cell(T, s, nh);                // my home cell inside the tile
            const int hx = h % G.bx, hy = h / G.bx;
            const float4 fi = xf[s];
            const int mol_i = __float_as_int(fi.w);
            const float2 nxi = make_float2(-fi.x, -fi.x), nyi = make_float2(-fi.y, -fi.y), nzi = make_float2(-fi.z, -fi.z);
            const float2 nhi = make_float2(-P.fcut_hi, -P.fcut_hi);
            const int band_bits = __float_as_int(P.fband);
            const int mol_x = mol_i == -1 ? 0x3ffffffe : mol_i;
            int count = 0, half_count = 0;
            RowWriter16 W16;
            W16.init(nbr, s);
            RowWriter32 W;
            W.init(nbr, s);
            const bool img_tile = F16 && T.any_image != 0;
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                const int oy = r % 3 - 1, oz = r / 3 - 1;
                const bool half_row = (oz == 1) || (oz == 0 && oy == 1);
                const bool centre_row = (oz == 0 && oy == 0);
                const int c0 = ((oz + 1) * nry + (hy + oy + 1)) * ncc + hx;   // candidate cells c0 (ox=-1), c0+1 (own column), c0+2 (ox=+1)
                int wlo = T.off[c0], whi = T.off[c0 + 3];
                const int cut_a = T.off[c0 + 1];
                const int self_q = centre_row ? cut_a + (s - T.beg[c0 + 1]) : -1;
                const int cut_b = T.off[c0 + 2];
                const bool im0 = img_tile && T.code[c0] != 13, im1 = img_tile && T.code[c0 + 1] != 13, im2 = img_tile && T.code[c0 + 2] != 13;
                if (P.xwindow) {
                    // Cells are sorted along x inside (k_cell_finalize), so the three cells of a row are ONE list ascending
                    // in the staged x.  Only candidates with |x_j - x_i| <= reach can be within the list cutoff, where
                    // reach^2 = cut^2 - (distance to the row's y slab)^2 - (distance to its z slab)^2; both ends of that
                    // window are found by bisection.  All in FP32 with a margin far above its rounding (1e-3 in reach^2
                    // plus 1e-3 in x against a band of ~1e-4): a dropped candidate could not have passed the mask test.
                    const float dy = oy == 0 ? 0.f : (oy > 0 ? fmaxf((T.cy0 + hy + 1) * P.flsy - fi.y, 0.f) : fmaxf(fi.y - (T.cy0 + hy) * P.flsy, 0.f));
                    const float dz = oz == 0 ? 0.f : (oz > 0 ? fmaxf((P.zcell0 + T.cz + 1) * P.flsz - fi.z, 0.f) : fmaxf(fi.z - (P.zcell0 + T.cz) * P.flsz, 0.f));
                    const float reach2 = P.fcut_hi * 1.001f - dy * dy - dz * dz;
                    if (reach2 <= 0.f) continue;
                    const float reach = sqrtf(reach2) + 1e-3f;
                    const float xlo = fi.x - reach, xhi = fi.x + reach;
                    int lo = wlo, hi = whi;                       // first candidate with x >= xlo
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (candX[mid] < xlo) lo = mid + 1; else hi = mid; }
                    wlo = lo;
                    hi = whi;                                     // first candidate with x > xhi
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (candX[mid] <= xhi) lo = mid + 1; else hi = mid; }
                    whi = lo;
                }
#pragma unroll 1
                for (int q0 = wlo & ~3; q0 < whi; q0 += 32) {
#ifdef SEPGPU_EMU
                    sepgpu_emu_counter[0] += 32;         // CPU kernel emulator only: candidates tested (work statistics)
#endif
                    // 32 candidates, eight at a time: d = r^2 - cut_hi in packed FP32 (two candidates per instruction);
                    // the sign bit of d IS the mask bit (funnel-shifted in), and the smallest signed integer among the
                    // bit patterns of the d's belongs to the accepted candidate closest to the cutoff -- which tells
                    // whether anyone sits in the FP32 error band [cut_lo, cut_hi) that needs the exact test.
                    unsigned mask = 0, band = 0, same = 0;
                    int dmin = 0x7fffffff, ntest = 0;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (g > 0 && q0 + 8 * g >= whi) break;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float4 X = *reinterpret_cast<const float4 *>(candX + q0 + 8 * g + 4 * h);
                            const float4 Y = *reinterpret_cast<const float4 *>(candY + q0 + 8 * g + 4 * h);
                            const float4 Z = *reinterpret_cast<const float4 *>(candZ + q0 + 8 * g + 4 * h);
                            const float2 dx0 = f2_add(make_float2(X.x, X.y), nxi), dx1 = f2_add(make_float2(X.z, X.w), nxi);
                            const float2 dy0 = f2_add(make_float2(Y.x, Y.y), nyi), dy1 = f2_add(make_float2(Y.z, Y.w), nyi);
                            const float2 dz0 = f2_add(make_float2(Z.x, Z.y), nzi), dz1 = f2_add(make_float2(Z.z, Z.w), nzi);
                            const float2 d0 = f2_fma(dz0, dz0, f2_fma(dy0, dy0, f2_fma(dx0, dx0, nhi)));
                            const float2 d1 = f2_fma(dz1, dz1, f2_fma(dy1, dy1, f2_fma(dx1, dx1, nhi)));
                            mask = shift_in_sign(mask, d0.x); mask = shift_in_sign(mask, d0.y);
                            mask = shift_in_sign(mask, d1.x); mask = shift_in_sign(mask, d1.y);
                            dmin = imin3(dmin, __float_as_int(d0.x), __float_as_int(d0.y));
                            dmin = imin3(dmin, __float_as_int(d1.x), __float_as_int(d1.y));
                            if (OPT == SEPGPU_EXCL_SAME_MOL) {               // same molecule <=> (mol_j ^ mol_i) - 1 is negative
                                const int4 M = *reinterpret_cast<const int4 *>(cand_mol + q0 + 8 * g + 4 * h);
                                same = shift_in_sign(same, __int_as_float((M.x ^ mol_x) - 1));
                                same = shift_in_sign(same, __int_as_float((M.y ^ mol_x) - 1));
                                same = shift_in_sign(same, __int_as_float((M.z ^ mol_x) - 1));
                                same = shift_in_sign(same, __int_as_float((M.w ^ mol_x) - 1));
                            }
                        }
                        ntest += 8;
                    }
                    mask = __brev(mask) >> (32 - ntest);                 // bit b = candidate q0 + b
                    if (OPT == SEPGPU_EXCL_SAME_MOL) mask &= ~(__brev(same) >> (32 - ntest));      // source/sepprfrc.c:673-674
                    if (q0 < wlo) mask &= ~((1u << (wlo - q0)) - 1u);    // (the block starts on a multiple of four)
                    if (dmin < 0 && (dmin & 0x7fffffff) <= band_bits) {
                        // someone within the error band below cut_hi: find them (rare)
                        for (int b = 0; b < ntest; b++) {
                            const float ex = candX[q0 + b] - fi.x, ey = candY[q0 + b] - fi.y, ez = candZ[q0 + b] - fi.z;
                            const float e = fmaf(ez, ez, fmaf(ey, ey, fmaf(ex, ex, -P.fcut_hi)));
                            if (e < 0.f && e >= -P.fband) band |= 1u << b;
                        }
                    }
                    const int nvalid = whi - q0;
                    if (nvalid < 32) mask &= (1u << nvalid) - 1u;
                    if ((unsigned)(self_q - q0) < 32u) mask &= ~(1u << (self_q - q0));
                    band &= mask;
                    // rare slow filters first, so that the append loop below is branch-light:
                    // candidates inside the FP32 error band take the exact FP64 test ...
                    while (band) {
                        const int b = __ffs(band) - 1;
                        band &= band - 1;
                        int code;
                        const int j = (int)(candW[q0 + b] & SEPGPU_INDEX_MASK);
                        // (under the prefilter preconditions the image pair_exact picks equals the cell image)
                        if (!pair_exact(xs[s], xs[j], P, code)) mask &= ~(1u << b);
                    }
                    // ... and the exclusion rules remove their pairs from the mask
                    if (OPT == SEPGPU_EXCL_BONDED) {
                        unsigned m2 = mask;
                        while (m2) {
                            const int b = __ffs(m2) - 1;
                            m2 &= m2 - 1;
                            const int j = (int)(candW[q0 + b] & SEPGPU_INDEX_MASK);
                            if (excluded<OPT>(mol_i, OPT == SEPGPU_EXCL_SAME_MOL ? cand_mol[q0 + b] : 0, s, j, order,
                                              excl_bond, excl_angle, excl_dihed)) mask &= ~(1u << b);
                        }
                    }
                    // reference half-list length: cells of the half stencil, or (centre row) everything
                    // stored after my own position -- my own cell with j2 > j1 and the ox = +1 cell
                    if (half_row) half_count += __popc(mask);
                    else if (centre_row) {
                        // the ox = +1 cell counts whole, my own cell by atom index (slots of a cell are ordered along x)
                        half_count += __popc(mask & bit_range(cut_b - q0, 32));
                        unsigned own = mask & bit_range(cut_a - q0, cut_b - q0);
                        if (own) {
                            if (nhome <= home_cap) {
                                const int my_i = home_order[s - a0];
                                while (own) {
                                    const int b = __ffs(own) - 1;
                                    own &= own - 1;
                                    half_count += home_order[(int)(candW[q0 + b] & SEPGPU_INDEX_MASK) - a0] > my_i;
                                }
                            } else {
                                const int my_i = order[s];
                                while (own) {
                                    const int b = __ffs(own) - 1;
                                    own &= own - 1;
                                    half_count += order[candW[q0 + b] & SEPGPU_INDEX_MASK] > my_i;
                                }
                            }
                        }
                    }
                    const int nacc = __popc(mask);
                    if (count + nacc <= P.cap) {
                        if (F16) {
                            if (img_tile) {
                                unsigned imgbits = 0;                    // candidates of this block that sit in a periodic image
                                if (im0) imgbits |= bit_range(wlo - q0, cut_a - q0);
                                if (im1) imgbits |= bit_range(cut_a - q0, cut_b - q0);
                                if (im2) imgbits |= bit_range(cut_b - q0, whi - q0);
                                while (mask) {
                                    const int b = __ffs(mask) - 1;
                                    mask &= mask - 1;
                                    W16.push((unsigned)(q0 + b) | (((imgbits >> b) & 1u) << 15), count, P.npad);
                                    count++;
                                }
                            } else {
                                while (mask) {
                                    const int b = __ffs(mask) - 1;
                                    mask &= mask - 1;
                                    W16.push((unsigned)(q0 + b), count, P.npad);
                                    count++;
                                }
                            }
                        } else {
                            // (32-bit rows are long where they are used -- 400 entries per atom in water -- and leave as whole
                            //  128-bit chunks: single 4-byte stores were measured 50 % slower there)
                            while (mask) {
                                const int b = __ffs(mask) - 1;
                                mask &= mask - 1;
                                W.push(candW[q0 + b], count, P.npad);
                                count++;
                            }
                        }
                    } else {
                        count += nacc;                                   // overflow: the host grows the list and rebuilds
                    }
                }
            }
            if (count <= P.cap) {
                if (F16) W16.finish(count, (unsigned)total, P.npad);
                else W.finish(count);
            }
            cnt[s] = min(count, P.cap);
            blk_max = max(blk_max, count); blk_half = max(blk_half, half_count); blk_sum += count;
        }
    }
    // block statistics: warp reduce, then shared atomics, then three global atomics per CTA
    for (int o = 16; o > 0; o >>= 1) {
        blk_max = max(blk_max, __shfl_xor_sync(0xffffffffu, blk_max, o));
        blk_half = max(blk_half, __shfl_xor_sync(0xffffffffu, blk_half, o));
        blk_sum += __shfl_xor_sync(0xffffffffu, blk_sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_red[0], blk_max); atomicMax(&s_red[1], blk_half); atomicAdd(&s_red[2], blk_sum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMax(&scal->max_neighb, s_red[0]);
        atomicMax(&scal->max_half, s_red[1]);
        atomicAdd((unsigned long long *)&scal->npairs_listed, (unsigned long long)s_red[2]);
    }
}
