// sepgpu_neighb.cu -- GPU cell binning + Verlet-list build.
//
// Stands in for sep_make_celllist + sep_make_neighblist_from_llist{,_nonbonded,_excl_same_mol}
// (reference source/sepprfrc.c:394-415, 419-513, 517-603, 606-700).  The reference walks a linked
// cell list serially and stores each pair once (half list); here
//   1. atoms are binned with the reference's exact FP64 expression (int)(x/lsubbox) and
//      counting-sorted by cell; inside a cell the order is ascending x, ties by atom index (deterministic),
//      and cells are laid out brick-major (bricks of BX x 4 x 4 cells, x fastest inside a brick) so
//      that atoms close in space are close in memory in all three directions,
//   2. a tiled kernel sweeps the 27 surrounding cells and writes a transposed FULL list.
// Acceptance reproduces the reference bit for bit: r2 = ((0+dx*dx)+dy*dy)+dz*dz from wrapped
// positions with the sep_Wrap branches, no FMA contraction, r2 < (cf+skin)^2.  r2 is symmetric in
// (i,j) bit for bit, so the full list is exactly the reference's half list read both ways.
// An FP32 prefilter classifies candidates that are not within a rigorous error band of the cutoff;
// only band candidates pay for the exact FP64 test.
#include "sepgpu_internal.cuh"

#include "sepgpu_tile.cuh"

#define BUILD_WARPS 8

// ---- binning ------------------------------------------------------------------------------------------
__global__ void k_cell_count(const d4 *__restrict__ x4, int n, double lsx, double lsy, double lsz,
                             CellGrid G, int *__restrict__ cell_of, int *__restrict__ cell_cnt, DevScalars *scal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 p = x4[i];
    // source/sepprfrc.c:404-406, IEEE division then truncation
    int cx = (int)__ddiv_rn(p.x, lsx), cy = (int)__ddiv_rn(p.y, lsy), cz = (int)__ddiv_rn(p.z, lsz);
    const int nzg = G.dd ? G.nzg : G.nz;
    if (cx < 0 || cx >= G.nx || cy < 0 || cy >= G.ny || cz < 0 || cz >= nzg || !(p.x == p.x)) {
        // The reference does not clamp: it forms the linear index cx + cy*nx + cz*nx*ny and, while that
        // stays inside the head array, files the atom under whatever cell the index aliases to
        // (source/sepprfrc.c:404-412; e.g. sep_set_lattice's +1.0 offset in prg6 puts atoms beyond L).
        // Reproduce that, and make the host redo this build with the exact (minimum-image) builder,
        // because an aliased atom breaks the cell-image == minimum-image premise of the fast one.
        const long long lin = (long long)cx + (long long)cy * G.nx + (long long)cz * G.nx * G.ny;
        if (G.dd || !(p.x == p.x) || lin < 0 || lin >= (long long)G.nx * G.ny * nzg) {
            scal->error = SEPGPU_ECELL;
            cx = min(max(cx, 0), G.nx - 1); cy = min(max(cy, 0), G.ny - 1); cz = min(max(cz, 0), nzg - 1);
        } else {
            cx = (int)(lin % G.nx); cy = (int)((lin / G.nx) % G.ny); cz = (int)(lin / ((long long)G.nx * G.ny));
            scal->aliased_seen = 1;            // "aliased atom seen" flag of this build
        }
    }
    if (G.dd) {                         // global layer -> local layer of this slab
        cz -= G.zoff;
        if (cz < 0) cz += G.nzg; else if (cz >= G.nzg) cz -= G.nzg;
        if (cz >= G.nz) { scal->error = SEPGPU_ECELL; cz = G.nz - 1; }
    }
    const int key = cell_key(cx, cy, cz, G);
    cell_of[i] = key;
    atomicAdd(&cell_cnt[key], 1);
}

// three-kernel exclusive scan over the cell counters (also clears them for the scatter pass)
#define SCAN_BLOCK 1024
#define SCAN_ITEMS 2
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_local(int *__restrict__ cnt, int *__restrict__ start, int *__restrict__ block_sum, int ncell)
{
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int base = (blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS], tot = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) { v[q] = base + q < ncell ? cnt[base + q] : 0; tot += v[q]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
        if (lane == 31) block_sum[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = wsum[wid] + incl - tot;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++)
        if (base + q < ncell) { start[base + q] = run; run += v[q]; cnt[base + q] = 0; }
}

__global__ void __launch_bounds__(1024) k_scan_blocks(int *__restrict__ block_sum, int nblocks)
{
    // single block: exclusive scan of up to 1024*k block totals
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < nblocks ? block_sum[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + wsum[wid] + incl - v;
        if (idx < nblocks) block_sum[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sum[nblocks] = carry;      // grand total
}

__global__ void k_scan_apply(int *__restrict__ start, const int *__restrict__ block_sum, int ncell, int nblocks)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ncell) start[idx] += block_sum[idx / (SCAN_BLOCK * SCAN_ITEMS)];
    if (idx == 0) start[ncell] = block_sum[nblocks];
}

// exclusive scan of cnt[0..n) into start[0..n], start[n] = total; cnt is cleared.  scratch: n/2048 + 2 ints
int sepgpu_exclusive_scan(cudaStream_t st, int *cnt, int *start, int *scratch, int n)
{
    const int nb = (n + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
    k_scan_local<<<nb, SCAN_BLOCK, 0, st>>>(cnt, start, scratch, n);
    k_scan_blocks<<<1, 1024, 0, st>>>(scratch, nb);
    k_scan_apply<<<(n + 255) / 256 + 1, 256, 0, st>>>(start, scratch, n, nb);
    KERNEL_CHECK();
    return 0;
}

__global__ void k_cell_scatter(const int *__restrict__ cell_of, int n, const int *__restrict__ cell_start,
                               int *__restrict__ cell_cnt, int *__restrict__ tmp_slot)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    tmp_slot[cell_start[c] + atomicAdd(&cell_cnt[c], 1)] = i;
}

// Order inside a cell: ascending x (FP32 image of the coordinate, ties by atom index) -- deterministic, independent of the
// atomic arrival order, and it makes every x-row of cells a list sorted along x, which lets the tiled builder restrict
// its sweep of a row to the window |x_j - x_i| <= reach (sepgpu_neighb_tile.cuh).  One warp per cell: keys staged in
// shared memory, rank by counting; then the sorted copies are written.
#define FIN_WARPS 8
#define FIN_CAP 256            // atoms of one cell ranked from shared memory; fuller cells take the global loop
__global__ void __launch_bounds__(FIN_WARPS * 32)
k_cell_finalize(const int *__restrict__ tmp_slot, const int *__restrict__ cell_start, const d4 *__restrict__ x4,
                int nkey, int *__restrict__ order, int *__restrict__ rank,
                d4 *__restrict__ xs, float4 *__restrict__ xf, i4 *__restrict__ cr4,
                unsigned char *__restrict__ cls, CellGrid G)
{
    __shared__ unsigned s_kx[FIN_WARPS][FIN_CAP];
    __shared__ int s_ki[FIN_WARPS][FIN_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int key = blockIdx.x * FIN_WARPS + w;
    if (key >= nkey) return;
    const int b = cell_start[key], m = cell_start[key + 1] - b;
    if (m == 0) return;
    const bool staged = m <= FIN_CAP;
    if (staged) {
        for (int k = lane; k < m; k += 32) {
            const int i = tmp_slot[b + k];
            s_ki[w][k] = i;
            s_kx[w][k] = __float_as_uint((float)x4[i].x);        // x >= 0: the bit pattern orders like the value
        }
        __syncwarp();
    }
    unsigned char cl = 0;
    if (cls) {             // slab decomposition: does this cell's neighbourhood reach into a halo layer?
        int cx, cy, cz;
        key_cell(key, G, cx, cy, cz);
        cl = (cz == 0 || cz == G.nz - 1) ? 2 : ((cz == 1 || cz == G.nz - 2) ? 1 : 0);
    }
    for (int k = lane; k < m; k += 32) {
        int i, r = 0;
        if (staged) {
            i = s_ki[w][k];
            const unsigned kx = s_kx[w][k];
            for (int q = 0; q < m; q++) {
                const unsigned kq = s_kx[w][q];
                r += (kq < kx) || (kq == kx && s_ki[w][q] < i);
            }
        } else {
            i = tmp_slot[b + k];
            const unsigned kx = __float_as_uint((float)x4[i].x);
            for (int q = 0; q < m; q++) {
                const int iq = tmp_slot[b + q];
                const unsigned kq = __float_as_uint((float)x4[iq].x);
                r += (kq < kx) || (kq == kx && iq < i);
            }
        }
        const int s = b + r;
        order[s] = i;
        rank[i] = s;
        const d4 p = x4[i];
        xs[s] = p;
        xf[s] = make_float4((float)p.x, (float)p.y, (float)p.z, __int_as_float(tag_mol(p.w)));
        cr4[i].w = 0;          // crossings since list build
        if (cls) cls[s] = cl;
    }
}

// ---- exclusion predicates (source/sepprfrc.c:703-740), original atom indices ----------------------------
__device__ __forceinline__ int share_tab(const int *__restrict__ tab, int width, int a, int b)
{
    for (int k = 0; k < width; k++) {
        int ta = tab[a * width + k], tb = tab[b * width + k];
        if (ta == -1 || tb == -1) break;
        if (ta == b || tb == a) return 1;
    }
    return 0;
}

struct BuildParams {
    double Lx, Ly, Lz;
    double cut2;
    float fLx, fLy, fLz, fcut_lo, fcut_hi, fband;       // fband: width of the FP32 error band below fcut_hi
    float flsy, flsz;       // cell widths along y and z (window of the tiled builder)
    int zcell0;             // global cell layer of local layer 0 (slab runs: zoff, else 0)
    int xwindow;            // tiled builder: restrict each row to the x window within reach (cells at least as wide as the cutoff)
    CellGrid G;
    int n, npad, cap;
    unsigned opt;
    int prefilter;
};

// exact reference test; returns accept and the image code chosen by the sep_Wrap branches
__device__ __forceinline__ bool pair_exact(const d4 &a, const d4 &b, const BuildParams &P, int &code)
{
    double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    int sx = 0, sy = 0, sz = 0;
    if (dx > 0.5 * P.Lx) { dx = __dsub_rn(dx, P.Lx); sx = 1; } else if (dx < -0.5 * P.Lx) { dx = __dadd_rn(dx, P.Lx); sx = -1; }
    if (dy > 0.5 * P.Ly) { dy = __dsub_rn(dy, P.Ly); sy = 1; } else if (dy < -0.5 * P.Ly) { dy = __dadd_rn(dy, P.Ly); sy = -1; }
    if (dz > 0.5 * P.Lz) { dz = __dsub_rn(dz, P.Lz); sz = 1; } else if (dz < -0.5 * P.Lz) { dz = __dadd_rn(dz, P.Lz); sz = -1; }
    double r2 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(dx, dx)), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    code = (sx + 1) + 3 * (sy + 1) + 9 * (sz + 1);
    return r2 < P.cut2;
}

template <unsigned OPT>
__device__ __forceinline__ bool excluded(int mol_i, int mol_j, int s, int j, const int *__restrict__ order,
                                         const int *__restrict__ eb, const int *__restrict__ ea,
                                         const int *__restrict__ ed)
{
    if (OPT == SEPGPU_EXCL_SAME_MOL) return !(mol_i == -1 || mol_i != mol_j);                      // :673-674
    if (OPT == SEPGPU_EXCL_BONDED) {
        const int io = order[s], jo = order[j];
        return share_tab(eb, 10, io, jo) + share_tab(ea, 10, io, jo) + share_tab(ed, 20, io, jo) != 0;   // :583
    }
    return false;
}

// ---- exact fallback: one warp per atom, FP64 test for every candidate -------------------------------------
// Used for small or degenerate grids (< 4 cells in a direction, skin enlarged past the cell width).
__global__ void __launch_bounds__(BUILD_WARPS * 32)
k_build_list(const d4 *__restrict__ xs, const float4 *__restrict__ xf, const int *__restrict__ order,
             const int *__restrict__ cell_of, const int *__restrict__ cell_start,
             const int *__restrict__ excl_bond, const int *__restrict__ excl_angle,
             const int *__restrict__ excl_dihed, unsigned *__restrict__ nbr, int *__restrict__ cnt,
             DevScalars *scal, BuildParams P)
{
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * BUILD_WARPS + (threadIdx.x >> 5);
    if (s >= P.n) return;
    const unsigned lt_mask = (1u << lane) - 1u;
    const CellGrid G = P.G;

    const int i_orig = order[s];
    int cx, cy, cz;
    key_cell(cell_of[i_orig], G, cx, cy, cz);
    const d4 pi = xs[s];
    const int mol_i = tag_mol(pi.w);

    int count = 0, half_count = 0;

    for (int oz = -1; oz <= 1; oz++) {
        int mz = cz + oz;
        if (mz == G.nz) mz = 0; else if (mz == -1) mz = G.nz - 1;
        for (int oy = -1; oy <= 1; oy++) {
            int my = cy + oy;
            if (my == G.ny) my = 0; else if (my == -1) my = G.ny - 1;
            for (int ox = -1; ox <= 1; ox++) {
                int mx = cx + ox;
                if (mx == G.nx) mx = 0; else if (mx == -1) mx = G.nx - 1;
                const int m2 = cell_key(mx, my, mz, G);
                const int jb = cell_start[m2], je = cell_start[m2 + 1];
                // reference half-stencil membership (source/sepprfrc.c:424-426), for the half-list length
                const bool in_half = (oz == 1) || (oz == 0 && (oy == 1 || (oy == 0 && ox == 1)));
                const bool same_cell = (ox == 0 && oy == 0 && oz == 0);
                for (int j0 = jb; j0 < je; j0 += 32) {
                    const int j = j0 + lane;
                    bool ok = false; int code = 13;
                    if (j < je && j != s) {
                        const d4 pj = xs[j];
                        ok = pair_exact(pi, pj, P, code);
                        if (ok && P.opt == SEPGPU_EXCL_SAME_MOL) ok = !excluded<SEPGPU_EXCL_SAME_MOL>(mol_i, tag_mol(pj.w), s, j, order, excl_bond, excl_angle, excl_dihed);
                        if (ok && P.opt == SEPGPU_EXCL_BONDED) ok = !excluded<SEPGPU_EXCL_BONDED>(0, 0, s, j, order, excl_bond, excl_angle, excl_dihed);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int pos = count + __popc(mask & lt_mask);
                        if (pos < P.cap) nbr[nbr_index(pos, s, P.npad)] = (unsigned)j | ((unsigned)code << SEPGPU_SHIFT_BITS);
                    }
                    count += __popc(mask);
                    if (in_half) half_count += __popc(mask);
                    else if (same_cell) {
                        // same cell: the reference keeps j2 > j1 (original indices)
                        half_count += __popc(__ballot_sync(0xffffffffu, ok && order[j < je ? j : s] > i_orig));
                    }
                }
            }
        }
    }
    if (lane == 0) {
        cnt[s] = min(count, P.cap);
        atomicMax(&scal->max_neighb, count);
        atomicMax(&scal->max_half, half_count);
        atomicAdd((unsigned long long *)&scal->npairs_listed, (unsigned long long)count);
    }
}

// ---- tiled build: the fast path is k_build_tile (sepgpu_neighb_tile.cuh) ---------------------------------------------
#include "sepgpu_neighb_tile.cuh"

__global__ void k_build_begin(DevScalars *s) { s->max_neighb = 0; s->max_half = 0; s->npairs_listed = 0; s->stage_needed = 0; s->stage_used = 0; s->aliased_seen = 0; }
__global__ void k_build_end(DevScalars *s) { s->nbuild += 1; }

int sepgpu_dd_before_build(sepgpu_ctx *c, const sepgpu_sys *sys, int *zoff, int *nz_local);

static int estimate_cap(const sepgpu_ctx *c, const sepgpu_sys *sys)
{
    const double vol = sys->length[0] * sys->length[1] * sys->length[2];
    const double rc = sys->cf + sys->skin;
    const double expect = 4.18879020478639 * rc * rc * rc * (double)c->n_global / vol;
    int cap = (int)(expect * 1.5) + 24;
    if (cap > c->n_global) cap = (int)c->n_global;
    return (cap + 7) & ~7;
}

// Tile shape for a mean cell occupancy: brick x-extent bx (1, 2, 4, 8) and x-runs per tile R (1, 2, 4), as many home
// atoms as one pass of TILE_THREADS threads takes, within the staging limits (16-bit slots, shared memory).
static void choose_tile(double mean_per_cell, int nx, int ny, int *bx_out, int *R_out)
{
    const double budget = 0.94 * TILE_THREADS;
    int bx = 1;
    while (bx < TILE_MAXBX && 2 * bx * mean_per_cell <= budget && 2 * bx <= nx) bx *= 2;
    int R = 1;
    while (R < TILE_MAXR && 2 * R * bx * mean_per_cell <= budget && 2 * R <= ny &&
           3.0 * (2 * R + 2) * (bx + 2) * mean_per_cell * 1.3 <= 0.8 * TILE_MAX_SLOTS) R *= 2;
    *bx_out = bx; *R_out = R;
}

extern "C" int sepgpu_neighb_build(sepgpu_ctx *c, const sepgpu_sys *sys, unsigned opt)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    const int nx = sys->nsubbox[0], ny = sys->nsubbox[1], nz = sys->nsubbox[2];
    if (nx < 1 || ny < 1 || nz < 1) { sepgpu_set_error("neighb_build: empty cell grid"); return SEPGPU_EINVAL; }
    if (opt < SEPGPU_ALL || opt > SEPGPU_EXCL_SAME_MOL) { sepgpu_set_error("neighb_build: bad opt %u", opt); return SEPGPU_EINVAL; }
    if (opt == SEPGPU_EXCL_BONDED && !c->have_excl) {
        sepgpu_set_error("neighb_build: SEP_EXCL_BONDED needs the bond/angle/dihed partner tables");
        return SEPGPU_ESTATE;
    }
    CellGrid G;
    G.nx = nx; G.ny = ny; G.nz = nz;
    G.dd = 0; G.zoff = 0; G.nzg = nz;
    if (c->dd) {
        // migrate atoms that left the slab, refresh the halo layers, then build on owned + halo atoms
        int rcd = sepgpu_dd_before_build(c, sys, &G.zoff, &G.nz);
        if (rcd) return rcd;
        G.dd = 1;
    }
    const double mean_per_cell = (double)c->n / ((double)nx * ny * G.nz);
    int R = 1;
    choose_tile(mean_per_cell, nx, ny, &G.bx, &R);
    if (c->tile_R_max > 0) while (R > c->tile_R_max) R /= 2;       // a tile overflowed the slot range earlier: smaller tiles
    G.nbx = (nx + G.bx - 1) / G.bx; G.nby = (ny + BRICK_YZ - 1) / BRICK_YZ; G.nbz = (G.nz + BRICK_YZ - 1) / BRICK_YZ;
    const long long nkey_ll = (long long)G.nbx * G.nby * G.nbz * G.bx * BRICK_YZ * BRICK_YZ;
    if (nkey_ll > (1LL << 30)) { sepgpu_set_error("neighb_build: too many cells"); return SEPGPU_EINVAL; }
    const int nkey = (int)nkey_ll;
    const int scan_blocks = (nkey + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
    if (nkey > c->ncell_cap) {
        if (c->cell_cnt) cudaFree(c->cell_cnt);
        if (c->cell_start) cudaFree(c->cell_start);
        c->cell_cnt = c->cell_start = NULL;
        // cell_cnt doubles as scratch for the scan's block totals (stored after the counters)
        CUDA_TRY(cudaMalloc((void **)&c->cell_cnt, sizeof(int) * ((size_t)nkey + 1 + scan_blocks + 1024)));
        CUDA_TRY(cudaMalloc((void **)&c->cell_start, sizeof(int) * ((size_t)nkey + 1)));
        c->ncell_cap = nkey;
    }
    int *block_sum = c->cell_cnt + nkey + 1;
    if (c->cap == 0) c->cap = estimate_cap(c, sys);

    const int B = 256, Gn = (c->n + B - 1) / B;
    bool force_exact = false;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (!c->nbr) CUDA_TRY(cudaMalloc((void **)&c->nbr, sizeof(unsigned) * (size_t)c->cap * c->npad));

        // Rows of 16-bit tile slots for the tile force kernels unless a consumer of global-index rows has been seen
        // (DPD, the molecule-pair table) or the context asks for them (option tile_list = 0)
        const bool f16 = c->tile_list && !c->need_atom_rows && !c->fij && !c->have_charge;
        bool built_f16 = false;
        ktimer_begin(c, &c->t_build);
        CUDA_TRY(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * ((size_t)nkey + 1), c->stream));
        k_build_begin<<<1, 1, 0, c->stream>>>(c->scal);
        k_cell_count<<<Gn, B, 0, c->stream>>>(c->x4, c->n, sys->lsubbox[0], sys->lsubbox[1], sys->lsubbox[2],
                                              G, c->cell_of, c->cell_cnt, c->scal);
        if (sepgpu_exclusive_scan(c->stream, c->cell_cnt, c->cell_start, block_sum, nkey)) return SEPGPU_ECUDA;
        k_cell_scatter<<<Gn, B, 0, c->stream>>>(c->cell_of, c->n, c->cell_start, c->cell_cnt, c->tmp_slot);
        if (c->dd && !c->cls) CUDA_TRY(cudaMalloc((void **)&c->cls, (size_t)c->ncap));
        k_cell_finalize<<<(nkey + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, c->stream>>>(c->tmp_slot, c->cell_start, c->x4, nkey,
                                                 c->order, c->rank, c->xs, c->xf, c->cr4, c->dd ? c->cls : NULL, G);
        BuildParams P;
        P.Lx = sys->length[0]; P.Ly = sys->length[1]; P.Lz = sys->length[2];
        const double cut = sys->cf + sys->skin;
        P.cut2 = cut * cut;                              // sep_Sq(sys->cf + sys->skin), :432
        P.fLx = (float)P.Lx; P.fLy = (float)P.Ly; P.fLz = (float)P.Lz;
        P.G = G;
        P.flsy = (float)sys->lsubbox[1]; P.flsz = (float)sys->lsubbox[2];
        P.zcell0 = G.dd ? G.zoff : 0;
        P.n = c->n; P.npad = c->npad; P.cap = c->cap; P.opt = opt;
        // FP32 prefilter: |r2_f32 - r2_exact| <= 2*sqrt(3)*cut * 4*Lmax*2^-24 (+ accumulation rounding);
        // the band below is 5x that bound.  It also needs cell image == minimum image, which holds when
        // every dimension has >= 4 cells and cut < 2 cells (< L/2).
        const double Lmax = fmax(P.Lx, fmax(P.Ly, P.Lz));
        const double wmin = fmin(sys->lsubbox[0], fmin(sys->lsubbox[1], sys->lsubbox[2]));
        const double band = 4.2e-6 * cut * Lmax + 2e-6 * P.cut2;
        P.prefilter = c->prefilter && !force_exact && nx >= 4 && ny >= 4 && nz >= 4 && cut < 1.95 * wmin && band < 0.05 * P.cut2;
        if (c->dd && !P.prefilter) { sepgpu_set_error("neighb_build: decomposed runs need >= 4 cells per direction"); return SEPGPU_EINVAL; }
        // the window argument needs every candidate within the cutoff to sit in the 27 cells around the atom anyway: cells
        // at least as wide as the cutoff (not the case after sep_set_skin enlarged the skin past the cell width)
        // Measured on B200 (profiles/r02_build_window_ab.txt): the per-lane windows halve the candidate tests but cost the
        // warp its shared (broadcast) candidate reads; that pays from ~24 atoms per cell on (water -13 %, butane -6 %) and
        // loses 3 % at 17 atoms per cell (the 1 M-atom Lennard-Jones fluid).  build_window = 2 forces it on.
        P.xwindow = cut <= wmin && (c->build_window == 2 || (c->build_window == 1 && mean_per_cell >= 24.0)) ? 1 : 0;
        P.fcut_lo = (float)(P.cut2 - band);
        P.fcut_hi = (float)(P.cut2 + band);
        P.fband = (float)(2.0 * band) * 1.0001f;
        const int ntile = nkey / (G.bx * R);
        if (P.prefilter) {
            if (c->tile_stage_cap == 0) {
                // 3 x (R+2) rows x (bx+2) cells x mean occupancy, with head-room for density fluctuations
                c->tile_stage_cap = ((int)(3.0 * (R + 2) * (G.bx + 2) * mean_per_cell * 1.3) + 96 + 31) & ~31;
            }
            const int stage_cap = c->tile_stage_cap;
            if ((size_t)ntile > c->tile_hdr_cap) {
                if (c->tile_hdr) cudaFree(c->tile_hdr);
                c->tile_hdr = NULL;
                CUDA_TRY(cudaMalloc((void **)&c->tile_hdr, sizeof(int4) * (size_t)ntile));
                c->tile_hdr_cap = (size_t)ntile;
            }
            if ((size_t)ntile * stage_cap > c->tile_src_cap) {
                if (c->tile_src) cudaFree(c->tile_src);
                c->tile_src = NULL;
                CUDA_TRY(cudaMalloc((void **)&c->tile_src, sizeof(unsigned) * (size_t)ntile * stage_cap));
                c->tile_src_cap = (size_t)ntile * stage_cap;
            }
            const int home_cap = stage_cap / 2;
            const size_t smem = (size_t)(stage_cap + TILE_PAD) * (sizeof(float4) + (opt == SEPGPU_EXCL_SAME_MOL ? sizeof(int) : 0)) + sizeof(int) * (size_t)home_cap;
            if (smem > 200 * 1024) { sepgpu_set_error("neighb_build: cell occupancy too high for the tiled builder"); return SEPGPU_EINVAL; }
#define TILE_ARGS c->xs, c->xf, c->order, c->cell_start, c->excl_bond, c->excl_angle, c->excl_dihed, c->nbr, c->cnt, c->scal, P, R, stage_cap, home_cap, c->tile_hdr, c->tile_src
#define TILE_LAUNCH(O, F)                                                                                                        \
    do {                                                                                                                         \
        CUDA_TRY(cudaFuncSetAttribute(k_build_tile<O, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        k_build_tile<O, F><<<ntile, TILE_THREADS, smem, c->stream>>>(TILE_ARGS);                                                 \
    } while (0)
            if (opt == SEPGPU_ALL) { if (f16) TILE_LAUNCH(SEPGPU_ALL, true); else TILE_LAUNCH(SEPGPU_ALL, false); }
            else if (opt == SEPGPU_EXCL_SAME_MOL) { if (f16) TILE_LAUNCH(SEPGPU_EXCL_SAME_MOL, true); else TILE_LAUNCH(SEPGPU_EXCL_SAME_MOL, false); }
            else { if (f16) TILE_LAUNCH(SEPGPU_EXCL_BONDED, true); else TILE_LAUNCH(SEPGPU_EXCL_BONDED, false); }
            built_f16 = f16;
#undef TILE_LAUNCH
#undef TILE_ARGS
        } else {
            k_build_list<<<(c->n + BUILD_WARPS - 1) / BUILD_WARPS, BUILD_WARPS * 32, 0, c->stream>>>(
                c->xs, c->xf, c->order, c->cell_of, c->cell_start, c->excl_bond, c->excl_angle, c->excl_dihed,
                c->nbr, c->cnt, c->scal, P);
        }
        k_build_end<<<1, 1, 0, c->stream>>>(c->scal);
        ktimer_end(c, &c->t_build);
        KERNEL_CHECK();

        // capacity / error check (one small D2H per rebuild)
        CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (c->scal_host->error == SEPGPU_ECELL) {
            sepgpu_set_error("neighb_build: atom outside the cell grid (position not in [0,L))");
            return SEPGPU_ECELL;
        }
        if (c->scal_host->max_half >= 3000) {         // SEP_NEIGHB, source/sepprfrc.c:499-501
            sepgpu_set_error("Too many neighbours");
            return SEPGPU_ENEIGHB;
        }
        if (c->scal_host->aliased_seen && P.prefilter) {      // an atom outside [0,L) was filed under an aliased cell
            force_exact = true;
            c->scal_host->nbuild -= 1;
            CUDA_TRY(cudaMemcpyAsync(&c->scal->nbuild, &c->scal_host->nbuild, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            continue;
        }
        if (c->scal_host->stage_needed > 0) {       // a tile needed a larger staging buffer: grow (or shrink the tile), rebuild
            if (c->scal_host->stage_needed > TILE_MAX_SLOTS) {
                if (R == 1) { sepgpu_set_error("neighb_build: cell occupancy too high for the tiled builder"); return SEPGPU_EINVAL; }
                R /= 2; c->tile_R_max = R; c->tile_stage_cap = 0;
            } else {
                c->tile_stage_cap = (c->scal_host->stage_needed * 5 / 4 + 63) & ~31;
                if (c->tile_stage_cap > TILE_MAX_SLOTS) c->tile_stage_cap = TILE_MAX_SLOTS;
            }
            c->scal_host->nbuild -= 1;
            CUDA_TRY(cudaMemcpyAsync(&c->scal->nbuild, &c->scal_host->nbuild, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            continue;
        }
        if (c->scal_host->max_neighb <= c->cap) {
            c->list_valid = true; c->list_opt = opt; c->sorted_identity = false; c->xs_current = true;
            c->zs_valid = false;
            c->list_f16 = built_f16;
            c->tile_grid = G; c->tile_R = R; c->tile_count = ntile;
            c->tile_stage_used = (c->scal_host->stage_used + 31) & ~31;
            c->tile_stride = c->tile_stage_cap;
            c->moved_since_build = false;
            c->list_gen++;
            c->grid_n[0] = nx; c->grid_n[1] = ny; c->grid_n[2] = nz;
            return 0;
        }
        // grow and retry (nbuild was bumped once too often; undo)
        c->scal_host->nbuild -= 1;
        CUDA_TRY(cudaMemcpyAsync(&c->scal->nbuild, &c->scal_host->nbuild, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        cudaFree(c->nbr); c->nbr = NULL;
        c->cap = (c->scal_host->max_neighb * 5 / 4 + 15) & ~7;
    }
    sepgpu_set_error("neighb_build: neighbour capacity did not converge");
    return SEPGPU_ENEIGHB;
}

// A consumer of global-index rows (list Coulomb, DPD, the molecule-pair table) met a list of tile slots: this context
// builds global-index rows from now on, and the current list is rebuilt in that format.  When atoms have moved since
// the list was built, the rebuild point also becomes the reference point of the skin trigger (xn <- x, cross_neighb <- 0,
// what the reference does at a rebuild, source/sepintgr.c:76-82): a list built here must be valid until every atom has
// moved skin/2 from HERE.
int sepgpu_reset_xn(sepgpu_ctx *c);
int sepgpu_need_global_rows(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    c->need_atom_rows = true;
    if (!c->list_valid || !c->list_f16) return 0;
    const bool moved = c->moved_since_build;
    int rc = sepgpu_neighb_build(c, sys, c->list_opt);
    if (rc) return rc;
    if (moved) return sepgpu_reset_xn(c);
    return 0;
}

// ---- pair export -----------------------------------------------------------------------------------------------
__global__ void k_export_pairs(const unsigned *__restrict__ nbr, const int *__restrict__ cnt,
                               const int *__restrict__ order, int n, int npad, int *__restrict__ out,
                               long long max_pairs, unsigned long long *counter, const int *__restrict__ gid, int n_own)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = order[s];
    if (i >= n_own) return;                      // halo atoms own no rows
    const int m = cnt[s];
    if (gid) i = gid[i];
    for (int k = 0; k < m; k++) {
        int j = order[nbr[nbr_index(k, s, npad)] & SEPGPU_INDEX_MASK];
        if (gid) j = gid[j];                     // decomposed run: global ids; a cross-rank pair is emitted by one rank only
        if (i < j) {
            unsigned long long p = atomicAdd(counter, 1ULL);
            if ((long long)p < max_pairs) { out[2 * p] = i; out[2 * p + 1] = j; }
        }
    }
}

// rows of 16-bit tile slots: the tile's staging table maps slots back to sorted indices
__global__ void __launch_bounds__(TILE_THREADS)
k_export_pairs_tile(const unsigned short *__restrict__ nbr16, const int *__restrict__ cnt, const int *__restrict__ order,
                    const int4 *__restrict__ tile_hdr, const unsigned *__restrict__ tile_src, int stride, int npad,
                    int *__restrict__ out, long long max_pairs, unsigned long long *counter, const int *__restrict__ gid, int n_own)
{
    const int4 hdr = tile_hdr[blockIdx.x];
    const unsigned *src = tile_src + (size_t)blockIdx.x * stride;
    for (int ab = threadIdx.x; ab < hdr.y; ab += TILE_THREADS) {
        const int s = hdr.x + ab;
        int i = order[s];
        if (i >= n_own) continue;
        if (gid) i = gid[i];
        const int m = cnt[s];
        for (int k = 0; k < m; k++) {
            const int q = nbr16[nbr16_index(k, s, npad)] & TILE_SLOT_MASK;
            int j = order[src[q] & SEPGPU_INDEX_MASK];
            if (gid) j = gid[j];
            if (i < j) {
                unsigned long long p = atomicAdd(counter, 1ULL);
                if ((long long)p < max_pairs) { out[2 * p] = i; out[2 * p + 1] = j; }
            }
        }
    }
}

extern "C" long long sepgpu_get_pairs(sepgpu_ctx *c, int *pairs, long long max_pairs)
{
    if (!c || !pairs || max_pairs <= 0) return SEPGPU_EINVAL;
    if (!c->list_valid) { sepgpu_set_error("get_pairs: no list"); return SEPGPU_ESTATE; }
    if (cudaSetDevice(c->device) != cudaSuccess) return SEPGPU_ECUDA;
    int *dout; unsigned long long *dcount;
    if (cudaMalloc((void **)&dout, sizeof(int) * 2 * (size_t)max_pairs) != cudaSuccess) return SEPGPU_ECUDA;
    if (cudaMalloc((void **)&dcount, sizeof(unsigned long long)) != cudaSuccess) { cudaFree(dout); return SEPGPU_ECUDA; }
    cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), c->stream);
    if (c->list_f16) k_export_pairs_tile<<<c->tile_count, TILE_THREADS, 0, c->stream>>>(reinterpret_cast<const unsigned short *>(c->nbr), c->cnt, c->order, c->tile_hdr, c->tile_src, c->tile_stride, c->npad, dout, max_pairs, dcount, c->dd ? c->gid : NULL, c->n_own);
    else k_export_pairs<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->nbr, c->cnt, c->order, c->n, c->npad, dout, max_pairs, dcount, c->dd ? c->gid : NULL, c->n_own);
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, dcount, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    long long np = (long long)h;
    if (np <= max_pairs) cudaMemcpy(pairs, dout, sizeof(int) * 2 * (size_t)np, cudaMemcpyDeviceToHost);
    cudaFree(dout); cudaFree(dcount);
    if (cudaGetLastError() != cudaSuccess) return SEPGPU_ECUDA;
    if (np > max_pairs) { sepgpu_set_error("get_pairs: %lld pairs exceed buffer %lld", np, max_pairs); return SEPGPU_EINVAL; }
    return np;
}
