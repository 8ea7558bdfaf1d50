// sepgpu_neighb_tile.cuh -- the fast Verlet-list builder (included by sepgpu_neighb.cu after the
// shared definitions CellGrid / BuildParams / pair_exact / excluded).
//
// One CTA owns one x-run of a brick: G.bx consecutive home cells of one x-row, whose atoms are
// contiguous in the cell-sorted array; one THREAD owns one home atom.  All 9 x (G.bx+2) candidate
// cells of the tile are staged ONCE into shared memory as FP32 positions already shifted to the right
// periodic image, with the finished list entry (sorted index | image code) in .w.  After a single
// barrier every thread sweeps, for each of the 9 (dy,dz) rows, the 3 cells around its own cell in two
// passes: a branch-free pass that tests 32 candidates into a bit mask, and a pass over the set bits
// that appends entries.  Every candidate is read from HBM once per CTA instead of once per atom, all
// lanes work on different atoms, and there is no barrier inside the sweep.
// Candidates inside the FP32 error band of the cutoff take the exact FP64 test (pair_exact), so the
// resulting pair set equals the reference's bit for bit.
#pragma once

#define TILE2_MAXCX 8
#define TILE2_THREADS 160
#define TILE2_PAD 32
#define TILE2_NCELL (9 * (TILE2_MAXCX + 2))

// bits [max(lo,0), min(hi,32)) of a 32-bit mask
__device__ __forceinline__ unsigned bit_range(int lo, int hi)
{
    lo = lo < 0 ? 0 : lo;
    hi = hi > 32 ? 32 : hi;
    if (hi <= lo) return 0u;
    const unsigned upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return upto & ~((1u << lo) - 1u);
}

// PAIR (option pair_tile): rows are written for PAIRS of sorted atoms (2t, 2t+1) instead of
// atoms -- row 2t holds the union of both atoms' neighbours, each entry flagged with the atom(s) it does NOT
// belong to (SEPGPU_PT_SKIP_A / _B), and cnt[2t+1] = -1.  The pair-tile force kernel (k_lj_pairtile) gathers
// every listed neighbour once for two atoms.  A pair whose atoms fall into different x-runs (different CTAs here)
// stays two single rows.  Membership flags are each atom's own accepted set, so per-atom pair sets -- and the
// reference's half-list length -- are exactly those of the per-atom list.
// SPATIAL (option cell_order = 1): slots of a cell follow a space-filling curve instead of the atom index.
// PRUNE (option build_prune = 1): candidate cells whose nearest point is beyond the cutoff are not swept.
template <unsigned OPT, bool PAIR, bool SPATIAL, bool PRUNE>
__global__ void __launch_bounds__(TILE2_THREADS, (PAIR ? 6 : 9))      // 40 registers for the per-atom variants, as measured in round 1
k_build_tile2(const d4 *__restrict__ xs, const float4 *__restrict__ xf, const int *__restrict__ order,
              const int *__restrict__ cell_start, const int *__restrict__ excl_bond,
              const int *__restrict__ excl_angle, const int *__restrict__ excl_dihed,
              unsigned *__restrict__ nbr, int *__restrict__ cnt, DevScalars *scal, BuildParams P, int stage_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *cand = reinterpret_cast<float4 *>(smem_raw);                       // [stage_cap + TILE2_PAD]
    int *cand_mol = reinterpret_cast<int *>(cand + stage_cap + TILE2_PAD);     // [stage_cap + TILE2_PAD] (SAME_MOL only)
    __shared__ int s_off[TILE2_NCELL + 1];     // staged offset of candidate cell (row r, cc): index r*(ncx+2)+cc
    __shared__ int s_beg[TILE2_NCELL];         // first sorted index of that cell
    __shared__ signed char s_w[TILE2_NCELL][3];   // periodic image (wx, wy, wz) of that cell
    __shared__ int s_home[TILE2_MAXCX + 1];
    __shared__ int s_red[3];

    const CellGrid G = P.G;
    int x0, cy, cz;
    key_cell(blockIdx.x * G.bx, G, x0, cy, cz);
    if (x0 >= G.nx || cy >= G.ny || cz >= G.nz) return;          // padding of the brick grid
    const int ncx = min(G.bx, G.nx - x0);
    const int ncc = ncx + 2, ncell = 9 * ncc;
    const int key0 = blockIdx.x * G.bx;
    if (G.dd && (cz == 0 || cz == G.nz - 1)) {                   // halo layer: its atoms own no rows
        const int b = cell_start[key0], e = cell_start[key0 + ncx];
        for (int q = b + threadIdx.x; q < e; q += TILE2_THREADS) cnt[q] = 0;
        return;
    }
    if (threadIdx.x <= ncx) s_home[threadIdx.x] = cell_start[key0 + threadIdx.x];
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    if (threadIdx.x < ncell) {
        const int r = threadIdx.x / ncc, cc = threadIdx.x % ncc;
        const int oy = r % 3 - 1, oz = r / 3 - 1;
        int mx = x0 - 1 + cc, wx = 0, my = cy + oy, wy = 0, mz = cz + oz, wz = 0;
        if (mx >= G.nx) { mx -= G.nx; wx = 1; } else if (mx < 0) { mx += G.nx; wx = -1; }
        if (my == G.ny) { my = 0; wy = 1; } else if (my == -1) { my = G.ny - 1; wy = -1; }
        if (G.dd) {                       // slab: no wrap in the local layer index; the image follows the global layer
            const int gl = G.zoff + mz;
            wz = gl < 0 ? -1 : (gl >= G.nzg ? 1 : 0);
        } else if (mz == G.nz) { mz = 0; wz = 1; } else if (mz == -1) { mz = G.nz - 1; wz = -1; }
        const int key = cell_key(mx, my, mz, G);
        const int b = cell_start[key];
        s_beg[threadIdx.x] = b;
        s_off[threadIdx.x] = cell_start[key + 1] - b;             // length for now
        s_w[threadIdx.x][0] = (signed char)wx; s_w[threadIdx.x][1] = (signed char)wy; s_w[threadIdx.x][2] = (signed char)wz;
    }
    __syncthreads();
    const int a0 = s_home[0], nhome = s_home[ncx] - a0;
    if (nhome == 0) return;
    if (threadIdx.x < 32) {                                       // exclusive scan of <= 90 lengths by one warp
        int carry = 0;
        for (int base = 0; base < ncell; base += 32) {
            const int idx = base + threadIdx.x;
            const int v = idx < ncell ? s_off[idx] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
            if (idx < ncell) s_off[idx] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (threadIdx.x == 0) s_off[ncell] = carry;
    }
    __syncthreads();
    const int total = s_off[ncell];
    if (total > stage_cap) {                                      // host grows the staging buffer and relaunches
        if (threadIdx.x == 0) atomicMax(&scal->stage_needed, total);
        return;
    }
    // ---- stage every candidate of the tile once ----
    for (int q = threadIdx.x; q < total + TILE2_PAD; q += TILE2_THREADS) {
        float4 f = make_float4(1e18f, 1e18f, 1e18f, 0.f);         // padding: never in range
        if (q < total) {
            int lo = 0, hi = ncell - 1;                            // last cell with s_off <= q
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= q) lo = mid; else hi = mid - 1; }
            const int j = s_beg[lo] + (q - s_off[lo]);
            f = xf[j];
            if (OPT == SEPGPU_EXCL_SAME_MOL) cand_mol[q] = __float_as_int(f.w);
            const int wx = s_w[lo][0], wy = s_w[lo][1], wz = s_w[lo][2];
            f.x += wx * P.fLx; f.y += wy * P.fLy; f.z += wz * P.fLz;
            const unsigned code = (unsigned)(wx + 1) + 3u * (unsigned)(wy + 1) + 9u * (unsigned)(wz + 1);
            f.w = __uint_as_float((unsigned)j | (code << SEPGPU_SHIFT_BITS));
        }
        cand[q] = f;
    }
    __syncthreads();

    int blk_max = 0, blk_half = 0, blk_sum = 0, blk_rows = 0;
    // PAIR: the thread <-> atom mapping starts at the even slot at or below a0, so that the two atoms of a globally
    // aligned pair sit in neighbouring lanes (even, odd) of one warp
    const int a_base = PAIR ? (a0 & ~1) : a0;
    const int span = a0 + nhome - a_base;
    for (int ab = 0; ab < span; ab += TILE2_THREADS) {
        const int s = a_base + ab + threadIdx.x;
        if (s >= a0 && s < a0 + nhome) {
            int h = 0;                                   // my home cell inside the tile
            for (int q = 1; q < ncx; q++) h += (s >= s_home[q]);
            bool paired = false;                         // PAIR: my partner s ^ 1 is a home atom of this tile too
            int h_lo = h, h_hi = h;
            if (PAIR) {
                const int sp = s ^ 1;
                paired = sp >= a0 && sp < a0 + nhome;
                if (paired) {
                    int hp = 0;
                    for (int q = 1; q < ncx; q++) hp += (sp >= s_home[q]);
                    h_lo = min(h, hp); h_hi = max(h, hp);
                }
            }
            const unsigned pm = 3u << (threadIdx.x & 30);           // the two lanes of my pair
            int own_total = 0;
            const float4 fi = xf[s];
            const int mol_i = __float_as_int(fi.w);
            int count = 0, half_count = 0;
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                const int oy = r % 3 - 1, oz = r / 3 - 1;
                const bool half_row = (oz == 1) || (oz == 0 && oy == 1);
                const bool centre_row = (oz == 0 && oy == 0);
                const int c0 = r * ncc + h;              // candidate cells c0 (ox=-1), c0+1 (own column), c0+2 (ox=+1)
                int wlo = s_off[c0], whi = s_off[c0 + 3];
                const int cut_a = s_off[c0 + 1];
                const int self_q = centre_row ? cut_a + (s - s_beg[c0 + 1]) : -1;
                if (PRUNE) {
                    // Nearest point of the row's cells, from my distances to the faces of my own cell (FP32, clamped at 0).
                    // Conservative: the limit carries a 1e-3 margin, far above FP32 rounding of positions and faces, so a
                    // dropped cell cannot hold a candidate the mask pass would have accepted.
                    const float dyl = fmaxf(fi.y - cy * P.flsy, 0.f), dyh = fmaxf((cy + 1) * P.flsy - fi.y, 0.f);
                    const float dzl = fmaxf(fi.z - cz * P.flsz, 0.f), dzh = fmaxf((cz + 1) * P.flsz - fi.z, 0.f);
                    const float dy = oy == 0 ? 0.f : (oy > 0 ? dyh : dyl), dz = oz == 0 ? 0.f : (oz > 0 ? dzh : dzl);
                    const float dyz2 = dy * dy + dz * dz;
                    const float lim = P.fcut_hi * 1.001f;
                    const float dxl = fmaxf(fi.x - (x0 + h) * P.flsx, 0.f), dxh = fmaxf((x0 + h + 1) * P.flsx - fi.x, 0.f);
                    if (dyz2 > lim) { whi = wlo; }                                        // whole row out of reach
                    else {
                        if (dyz2 + dxl * dxl > lim) wlo = cut_a;                              // ox = -1 cell out of reach
                        if (dyz2 + dxh * dxh > lim) whi = s_off[c0 + 2];                      // ox = +1 cell out of reach
                    }
                }
                // PAIR: both lanes of a pair sweep the union of their two windows in the same 32-candidate steps
                int plo = PAIR ? s_off[r * ncc + h_lo] : wlo, phi = PAIR ? s_off[r * ncc + h_hi + 3] : whi;
                if (PAIR && PRUNE) {
                    plo = wlo; phi = whi;
                    if (paired) {
                        const int olo = __shfl_xor_sync(pm, wlo, 1), ohi = __shfl_xor_sync(pm, whi, 1);
                        if (whi <= wlo) { plo = olo; phi = ohi; }                  // my range is empty: follow my partner's
                        else if (ohi > olo) { plo = min(wlo, olo); phi = max(whi, ohi); }
                    }
                }
#pragma unroll 1
                for (int q0 = plo; q0 < phi; q0 += 32) {
#ifdef SEPGPU_EMU
                    sepgpu_emu_counter[0] += 32;         // CPU kernel emulator only: candidates tested (work statistics)
#endif
                    unsigned mask = 0, band = 0;
#pragma unroll
                    for (int b = 0; b < 32; b++) {
                        const float4 fj = cand[q0 + b];
                        const float dx = fi.x - fj.x, dy = fi.y - fj.y, dz = fi.z - fj.z;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (r2 <= P.fcut_hi) mask |= 1u << b;
                        if (r2 >= P.fcut_lo) band |= 1u << b;
                    }
                    if (PAIR) {
                        mask &= bit_range(wlo - q0, whi - q0);          // my own 3-cell stencil only
                    } else {
                        const int nvalid = whi - q0;
                        if (nvalid < 32) mask &= (1u << nvalid) - 1u;
                    }
                    if ((unsigned)(self_q - q0) < 32u) mask &= ~(1u << (self_q - q0));
                    band &= mask;
                    // rare slow filters first, so that the append loop below is branch-light:
                    // candidates inside the FP32 error band take the exact FP64 test ...
                    while (band) {
                        const int b = __ffs(band) - 1;
                        band &= band - 1;
                        int code;
                        const int j = (int)(__float_as_uint(cand[q0 + b].w) & SEPGPU_INDEX_MASK);
                        // (under the prefilter preconditions the image pair_exact picks equals the cell image)
                        if (!pair_exact(xs[s], xs[j], P, code)) mask &= ~(1u << b);
                    }
                    // ... and the exclusion rules remove their pairs from the mask
                    if (OPT != SEPGPU_ALL) {
                        unsigned m2 = mask;
                        while (m2) {
                            const int b = __ffs(m2) - 1;
                            m2 &= m2 - 1;
                            const int j = (int)(__float_as_uint(cand[q0 + b].w) & SEPGPU_INDEX_MASK);
                            if (excluded<OPT>(mol_i, OPT == SEPGPU_EXCL_SAME_MOL ? cand_mol[q0 + b] : 0, s, j, order,
                                              excl_bond, excl_angle, excl_dihed)) mask &= ~(1u << b);
                        }
                    }
                    // reference half-list length: cells of the half stencil, or (centre row) everything
                    // stored after my own position -- my own cell with j2 > j1 and the ox = +1 cell
                    if (half_row) half_count += __popc(mask);
                    else if (SPATIAL && centre_row) {
                        // slots of a cell are not in index order: the ox = +1 cell counts whole, my own cell by atom index
                        const int cut_b = s_off[c0 + 2];
                        half_count += __popc(mask & bit_range(cut_b - q0, whi - q0));
                        unsigned own = mask & bit_range(cut_a - q0, cut_b - q0);
                        const int my_i = order[s];
                        while (own) {
                            const int b = __ffs(own) - 1;
                            own &= own - 1;
                            half_count += order[__float_as_uint(cand[q0 + b].w) & SEPGPU_INDEX_MASK] > my_i;
                        }
                    } else if (centre_row) {
                        const int d = self_q + 1 - q0;                   // first bit that counts
                        half_count += __popc(d <= 0 ? mask : (d >= 32 ? 0u : mask & ~((1u << d) - 1u)));
                    }
                    if (PAIR) own_total += __popc(mask);
                    if (PAIR && paired) {
                        const unsigned other = __shfl_xor_sync(pm, mask, 1);
                        if (!(threadIdx.x & 1)) {                        // the even atom owns the pair's row
                            unsigned u = mask | other;
                            const int nacc = __popc(u);
                            if (count + nacc <= P.cap) {
                                while (u) {
                                    const int b = __ffs(u) - 1;
                                    u &= u - 1;
                                    unsigned e = __float_as_uint(cand[q0 + b].w);
                                    if (!(mask >> b & 1u)) e |= SEPGPU_PT_SKIP_A;
                                    if (!(other >> b & 1u)) e |= SEPGPU_PT_SKIP_B;
                                    nbr[nbr_index(count, s, P.npad)] = e;
                                    count++;
                                }
                            } else {
                                count += nacc;
                            }
                        }
                        continue;
                    }
                    const int nacc = __popc(mask);
                    if (count + nacc <= P.cap) {
                        while (mask) {
                            const int b = __ffs(mask) - 1;
                            mask &= mask - 1;
                            nbr[nbr_index(count, s, P.npad)] = __float_as_uint(cand[q0 + b].w);
                            count++;
                        }
                    } else {
                        count += nacc;                                   // overflow: the host grows the list and rebuilds
                    }
                }
            }
            if (PAIR && paired && (threadIdx.x & 1)) cnt[s] = -1;        // second atom of a pair: its entries live in row s - 1
            else cnt[s] = min(count, P.cap);
            blk_max = max(blk_max, count); blk_half = max(blk_half, half_count); blk_sum += PAIR ? own_total : count;
            if (PAIR) blk_rows += count;
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
    if (PAIR) {                                  // statistics only: entries actually stored (union rows)
        for (int o = 16; o > 0; o >>= 1) blk_rows += __shfl_xor_sync(0xffffffffu, blk_rows, o);
        if ((threadIdx.x & 31) == 0 && blk_rows) atomicAdd((unsigned long long *)&scal->row_entries, (unsigned long long)blk_rows);
    }
}
