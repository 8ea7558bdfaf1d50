// sepgpu_neighb.cu -- GPU cell binning + warp-cooperative Verlet-list build.
//
// Stands in for sep_make_celllist + sep_make_neighblist_from_llist{,_nonbonded,_excl_same_mol}
// (reference source/sepprfrc.c:394-415, 419-513, 517-603, 606-700).  The reference walks a linked
// cell list serially and stores each pair once (half list); here
//   1. atoms are binned with the reference's exact FP64 expression (int)(x/lsubbox) and
//      counting-sorted by cell (ascending atom index inside a cell => deterministic),
//   2. one warp per atom sweeps the 27 surrounding cells of the sorted array and compacts the
//      accepted partners with ballot/popc into a transposed FULL list.
// Acceptance reproduces the reference bit for bit: r2 = ((0+dx*dx)+dy*dy)+dz*dz from wrapped
// positions with the sep_Wrap branches, no FMA contraction, r2 < (cf+skin)^2.  r2 is symmetric in
// (i,j) bit for bit, so the full list is exactly the reference's half list read both ways.
// An FP32 prefilter classifies candidates that are not within a rigorous error band of the cutoff;
// only band candidates pay for the exact FP64 test.
#include "sepgpu_internal.cuh"

#define BUILD_WARPS 8

// ---- binning ------------------------------------------------------------------------------------------
__global__ void k_cell_count(const d4 *__restrict__ x4, int n, double lsx, double lsy, double lsz,
                             int nx, int ny, int nz, int *__restrict__ cell_of,
                             int *__restrict__ cell_cnt, DevScalars *scal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 p = x4[i];
    // source/sepprfrc.c:404-406, IEEE division then truncation
    int cx = (int)__ddiv_rn(p.x, lsx), cy = (int)__ddiv_rn(p.y, lsy), cz = (int)__ddiv_rn(p.z, lsz);
    if (cx < 0 || cx >= nx || cy < 0 || cy >= ny || cz < 0 || cz >= nz || !(p.x == p.x)) {
        scal->error = SEPGPU_ECELL;
        cx = min(max(cx, 0), nx - 1); cy = min(max(cy, 0), ny - 1); cz = min(max(cz, 0), nz - 1);
    }
    int c = cx + cy * nx + cz * nx * ny;
    cell_of[i] = c;
    atomicAdd(&cell_cnt[c], 1);
}

// single-block exclusive scan over the cells; also clears the counters for the scatter pass
__global__ void __launch_bounds__(1024) k_cell_scan(int *__restrict__ cell_cnt, int *__restrict__ cell_start, int ncell)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        int idx = base + threadIdx.x;
        int v = idx < ncell ? cell_cnt[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        int excl = carry + wsum[wid] + incl - v;
        if (idx < ncell) { cell_start[idx] = excl; cell_cnt[idx] = 0; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_start[ncell] = carry;
}

__global__ void k_cell_scatter(const int *__restrict__ cell_of, int n, const int *__restrict__ cell_start,
                               int *__restrict__ cell_cnt, int *__restrict__ tmp_slot)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    tmp_slot[cell_start[c] + atomicAdd(&cell_cnt[c], 1)] = i;
}

// rank inside the cell by atom index (removes the atomic arrival order), then write the sorted copies
__global__ void k_cell_finalize(const int *__restrict__ tmp_slot, const int *__restrict__ cell_of,
                                const int *__restrict__ cell_start, const d4 *__restrict__ x4,
                                int n, int *__restrict__ order, int *__restrict__ rank,
                                d4 *__restrict__ xs, float4 *__restrict__ xf, i4 *__restrict__ cr4)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = tmp_slot[t];
    int c = cell_of[i];
    int b = cell_start[c], e = cell_start[c + 1];
    int r = 0;
    for (int q = b; q < e; q++) r += tmp_slot[q] < i;
    int s = b + r;
    order[s] = i;
    rank[i] = s;
    d4 p = x4[i];
    xs[s] = p;
    xf[s] = make_float4((float)p.x, (float)p.y, (float)p.z, __int_as_float(tag_mol(p.w)));
    cr4[i].w = 0;          // crossings since list build
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
    float fLx, fLy, fLz, fcut_lo, fcut_hi;
    int nx, ny, nz;
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

    const int i_orig = order[s];
    const int c = cell_of[i_orig];
    const int cx = c % P.nx, cy = (c / P.nx) % P.ny, cz = c / (P.nx * P.ny);
    const d4 pi = xs[s];
    const float4 fi = xf[s];
    const int mol_i = __float_as_int(fi.w);

    int count = 0, half_count = 0;

    for (int oz = -1; oz <= 1; oz++) {
        int mz = cz + oz; int wz = 0;
        if (mz == P.nz) { mz = 0; wz = 1; } else if (mz == -1) { mz = P.nz - 1; wz = -1; }
        for (int oy = -1; oy <= 1; oy++) {
            int my = cy + oy; int wy = 0;
            if (my == P.ny) { my = 0; wy = 1; } else if (my == -1) { my = P.ny - 1; wy = -1; }
            for (int ox = -1; ox <= 1; ox++) {
                int mx = cx + ox; int wx = 0;
                if (mx == P.nx) { mx = 0; wx = 1; } else if (mx == -1) { mx = P.nx - 1; wx = -1; }
                const int m2 = mx + my * P.nx + mz * P.nx * P.ny;
                const int jb = cell_start[m2], je = cell_start[m2 + 1];
                // reference half-stencil membership (source/sepprfrc.c:424-426), for the half-list length
                const bool in_half = (oz == 1) || (oz == 0 && (oy == 1 || (oy == 0 && ox == 1)));
                const bool same_cell = (ox == 0 && oy == 0 && oz == 0);
                // image of the candidate cell relative to the home atom: x_i - x_j is shifted by +w*L
                const float sxf = fi.x - wx * P.fLx, syf = fi.y - wy * P.fLy, szf = fi.z - wz * P.fLz;
                const int cell_code = (wx + 1) + 3 * (wy + 1) + 9 * (wz + 1);

                for (int j0 = jb; j0 < je; j0 += 32) {
                    const int j = j0 + lane;
                    bool ok = false; int code = cell_code; int j_orig = -1;
                    if (j < je && j != s) {
                        bool need_exact = !P.prefilter;
                        if (P.prefilter) {
                            const float4 fj = xf[j];
                            const float dx = sxf - fj.x, dy = syf - fj.y, dz = szf - fj.z;
                            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (r2 < P.fcut_lo) ok = true;
                            else if (r2 <= P.fcut_hi) need_exact = true;
                            if (ok && P.opt == SEPGPU_EXCL_SAME_MOL) {
                                const int mol_j = __float_as_int(fj.w);
                                if (!(mol_i == -1 || mol_i != mol_j)) ok = false;   // :673-674
                            }
                        }
                        if (need_exact) {
                            const d4 pj = xs[j];
                            ok = pair_exact(pi, pj, P, code);
                            if (ok && P.opt == SEPGPU_EXCL_SAME_MOL) {
                                const int mol_j = tag_mol(pj.w);
                                if (!(mol_i == -1 || mol_i != mol_j)) ok = false;
                            }
                        }
                        if (ok && P.opt == SEPGPU_EXCL_BONDED) {
                            j_orig = order[j];
                            const int b = share_tab(excl_bond, 10, i_orig, j_orig) +
                                          share_tab(excl_angle, 10, i_orig, j_orig) +
                                          share_tab(excl_dihed, 20, i_orig, j_orig);
                            if (b != 0) ok = false;                                  // :583
                        }
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int pos = count + __popc(mask & lt_mask);
                        if (pos < P.cap) nbr[(size_t)pos * P.npad + s] = (unsigned)j | ((unsigned)code << SEPGPU_SHIFT_BITS);
                    }
                    count += __popc(mask);
                    if (in_half) half_count += __popc(mask);
                    else if (same_cell) {
                        // same cell: the reference keeps j2 > j1 (original indices); sorted order inside a
                        // cell is ascending in the original index, so j > s is the same condition
                        half_count += __popc(__ballot_sync(0xffffffffu, ok && j > s));
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

// ---- tiled build: the fast path ------------------------------------------------------------------------------
// One CTA owns TILE_CX consecutive home cells of one x-row (their atoms are contiguous in the sorted
// array) and one THREAD owns one home atom.  For each of the 9 (dy,dz) rows the TILE_CX+2 candidate
// cells are staged once into shared memory as FP32 positions already shifted to the right periodic
// image; a thread then sweeps the 3 cells around its own cell.  Compared with one warp per atom this
// reads every candidate once per CTA instead of once per atom and keeps all 32 lanes busy on different
// atoms.  Candidates inside the FP32 error band of the cutoff take the exact FP64 test (pair_exact).
#define TILE_CX 4
#define TILE_THREADS 128
#define TILE_STAGE 1024

__global__ void __launch_bounds__(TILE_THREADS)
k_build_tile(const d4 *__restrict__ xs, const float4 *__restrict__ xf, const int *__restrict__ order,
             const int *__restrict__ cell_start, const int *__restrict__ excl_bond,
             const int *__restrict__ excl_angle, const int *__restrict__ excl_dihed,
             unsigned *__restrict__ nbr, int *__restrict__ cnt, DevScalars *scal, BuildParams P)
{
    __shared__ float4 cand[TILE_STAGE];
    __shared__ int s_off[TILE_CX + 3];      // staged offset of candidate cell cc (cc = 0..ncx+1), +1 end marker
    __shared__ int s_jbase[TILE_CX + 2];    // sorted index of a candidate = staged position + s_jbase[cc]
    __shared__ int s_wx[TILE_CX + 2];       // x image of candidate cell cc
    __shared__ int s_home[TILE_CX + 1];     // sorted-index boundaries of the home cells
    __shared__ int s_red[3];

    const int nseg = (P.nx + TILE_CX - 1) / TILE_CX;
    const int seg = blockIdx.x % nseg;
    const int cy = (blockIdx.x / nseg) % P.ny;
    const int cz = blockIdx.x / (nseg * P.ny);
    const int x0 = seg * TILE_CX;
    const int ncx = min(TILE_CX, P.nx - x0);
    const int row_home = (cy + cz * P.ny) * P.nx;
    if (threadIdx.x <= ncx) s_home[threadIdx.x] = cell_start[row_home + x0 + threadIdx.x];
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    __syncthreads();
    const int a0 = s_home[0], nhome = s_home[ncx] - a0;
    int blk_max = 0, blk_half = 0, blk_sum = 0;

    for (int ab = 0; ab < nhome; ab += TILE_THREADS) {
        const int s = a0 + ab + threadIdx.x;
        const bool active = s < a0 + nhome;
        int h = 0;                                   // my home cell inside the tile
        float4 fi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            for (int q = 1; q < ncx; q++) h += (s >= s_home[q]);
            fi = xf[s];
        }
        const int mol_i = __float_as_int(fi.w);
        int count = 0, half_count = 0;
        int i_orig = -1;

        for (int r = 0; r < 9; r++) {
            const int oy = r % 3 - 1, oz = r / 3 - 1;
            int my = cy + oy, wy = 0, mz = cz + oz, wz = 0;
            if (my == P.ny) { my = 0; wy = 1; } else if (my == -1) { my = P.ny - 1; wy = -1; }
            if (mz == P.nz) { mz = 0; wz = 1; } else if (mz == -1) { mz = P.nz - 1; wz = -1; }
            const int row = (my + mz * P.ny) * P.nx;
            __syncthreads();                         // previous row fully consumed
            if (threadIdx.x == 0) {
                int off = 0;
                for (int cc = 0; cc < ncx + 2; cc++) {
                    int mx = x0 - 1 + cc, wx = 0;
                    if (mx >= P.nx) { mx -= P.nx; wx = 1; } else if (mx < 0) { mx += P.nx; wx = -1; }
                    const int b = cell_start[row + mx], e = cell_start[row + mx + 1];
                    s_off[cc] = off; s_jbase[cc] = b - off; s_wx[cc] = wx;
                    off += e - b;
                }
                s_off[ncx + 2] = off;
            }
            __syncthreads();
            const int total = s_off[ncx + 2];
            const float shy = wy * P.fLy, shz = wz * P.fLz;
            const bool half_row = (oz == 1) || (oz == 0 && oy == 1);
            const int code_yz = 3 * (wy + 1) + 9 * (wz + 1);

            for (int base = 0; base < total; base += TILE_STAGE) {
                const int lim = min(total - base, TILE_STAGE);
                if (base > 0) __syncthreads();
                for (int q = threadIdx.x; q < lim; q += TILE_THREADS) {
                    const int g = base + q;
                    int cc = 0;
                    for (int t = 1; t < ncx + 2; t++) cc += (g >= s_off[t]);
                    float4 f = xf[g + s_jbase[cc]];
                    f.x += s_wx[cc] * P.fLx; f.y += shy; f.z += shz;
                    cand[q] = f;
                }
                __syncthreads();
                if (active) {
                    for (int cc = h; cc < h + 3; cc++) {
                        const int lo = max(s_off[cc], base) - base, hi = min(s_off[cc + 1], base + lim) - base;
                        const int jb = s_jbase[cc] + base;
                        const int code_c = (s_wx[cc] + 1) + code_yz;
                        const int ox = cc - 1 - h;
                        const bool in_half = half_row || (oz == 0 && oy == 0 && ox == 1);
                        const bool same_cell = (oz == 0 && oy == 0 && ox == 0);
                        for (int q = lo; q < hi; q++) {
                            const float4 fj = cand[q];
                            const float dx = fi.x - fj.x, dy = fi.y - fj.y, dz = fi.z - fj.z;
                            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (r2 > P.fcut_hi) continue;
                            const int j = q + jb;
                            if (j == s) continue;
                            int code = code_c;
                            bool ok = true;
                            if (r2 >= P.fcut_lo) ok = pair_exact(xs[s], xs[j], P, code);      // inside the error band
                            if (ok && P.opt == SEPGPU_EXCL_SAME_MOL) {
                                const int mol_j = __float_as_int(fj.w);
                                if (!(mol_i == -1 || mol_i != mol_j)) ok = false;             // :673-674
                            } else if (ok && P.opt == SEPGPU_EXCL_BONDED) {
                                if (i_orig < 0) i_orig = order[s];
                                const int j_orig = order[j];
                                if (share_tab(excl_bond, 10, i_orig, j_orig) + share_tab(excl_angle, 10, i_orig, j_orig) +
                                    share_tab(excl_dihed, 20, i_orig, j_orig) != 0) ok = false;  // :583
                            }
                            if (!ok) continue;
                            if (count < P.cap) nbr[(size_t)count * P.npad + s] = (unsigned)j | ((unsigned)code << SEPGPU_SHIFT_BITS);
                            count++;
                            half_count += (in_half || (same_cell && j > s)) ? 1 : 0;
                        }
                    }
                }
            }
        }
        if (active) {
            cnt[s] = min(count, P.cap);
            blk_max = max(blk_max, count); blk_half = max(blk_half, half_count); blk_sum += count;
        }
    }
    // block statistics: warp reduce, then three atomics per warp
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

__global__ void k_build_begin(DevScalars *s) { s->max_neighb = 0; s->max_half = 0; s->npairs_listed = 0; }
__global__ void k_build_end(DevScalars *s) { s->nbuild += 1; s->neighb_flag = 0; }

static int estimate_cap(const sepgpu_ctx *c, const sepgpu_sys *sys)
{
    const double vol = sys->length[0] * sys->length[1] * sys->length[2];
    const double rc = sys->cf + sys->skin;
    const double expect = 4.18879020478639 * rc * rc * rc * (double)c->n / vol;
    int cap = (int)(expect * 1.5) + 24;
    if (cap > c->n) cap = c->n;
    return (cap + 7) & ~7;
}

extern "C" int sepgpu_neighb_build(sepgpu_ctx *c, const sepgpu_sys *sys, unsigned opt)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    CUDA_TRY(cudaSetDevice(c->device));
    const int nx = sys->nsubbox[0], ny = sys->nsubbox[1], nz = sys->nsubbox[2];
    if (nx < 1 || ny < 1 || nz < 1) { sepgpu_set_error("neighb_build: empty cell grid"); return SEPGPU_EINVAL; }
    if (opt == SEPGPU_EXCL_BONDED && !c->have_excl) {
        sepgpu_set_error("neighb_build: SEP_EXCL_BONDED needs the bond/angle/dihed partner tables");
        return SEPGPU_ESTATE;
    }
    const long long ncell_ll = (long long)nx * ny * nz;
    if (ncell_ll > (1LL << 30)) { sepgpu_set_error("neighb_build: too many cells"); return SEPGPU_EINVAL; }
    const int ncell = (int)ncell_ll;
    if (ncell > c->ncell_cap) {
        if (c->cell_cnt) cudaFree(c->cell_cnt);
        if (c->cell_start) cudaFree(c->cell_start);
        c->cell_cnt = c->cell_start = NULL;
        CUDA_TRY(cudaMalloc((void **)&c->cell_cnt, sizeof(int) * ((size_t)ncell + 1)));
        CUDA_TRY(cudaMalloc((void **)&c->cell_start, sizeof(int) * ((size_t)ncell + 1)));
        c->ncell_cap = ncell;
    }
    if (c->cap == 0) c->cap = estimate_cap(c, sys);

    const int B = 256, G = (c->n + B - 1) / B;
    for (int attempt = 0; attempt < 6; attempt++) {
        if (!c->nbr) CUDA_TRY(cudaMalloc((void **)&c->nbr, sizeof(unsigned) * (size_t)c->cap * c->npad));

        ktimer_begin(c, &c->t_build);
        CUDA_TRY(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * ((size_t)ncell + 1), c->stream));
        k_build_begin<<<1, 1, 0, c->stream>>>(c->scal);
        k_cell_count<<<G, B, 0, c->stream>>>(c->x4, c->n, sys->lsubbox[0], sys->lsubbox[1], sys->lsubbox[2],
                                             nx, ny, nz, c->cell_of, c->cell_cnt, c->scal);
        k_cell_scan<<<1, 1024, 0, c->stream>>>(c->cell_cnt, c->cell_start, ncell);
        k_cell_scatter<<<G, B, 0, c->stream>>>(c->cell_of, c->n, c->cell_start, c->cell_cnt, c->tmp_slot);
        k_cell_finalize<<<G, B, 0, c->stream>>>(c->tmp_slot, c->cell_of, c->cell_start, c->x4, c->n,
                                                c->order, c->rank, c->xs, c->xf, c->cr4);
        BuildParams P;
        P.Lx = sys->length[0]; P.Ly = sys->length[1]; P.Lz = sys->length[2];
        const double cut = sys->cf + sys->skin;
        P.cut2 = cut * cut;                              // sep_Sq(sys->cf + sys->skin), :432
        P.fLx = (float)P.Lx; P.fLy = (float)P.Ly; P.fLz = (float)P.Lz;
        P.nx = nx; P.ny = ny; P.nz = nz;
        P.n = c->n; P.npad = c->npad; P.cap = c->cap; P.opt = opt;
        // FP32 prefilter: |r2_f32 - r2_exact| <= 2*sqrt(3)*cut * 4*Lmax*2^-24 (+ accumulation rounding);
        // the band below is 5x that bound.  It also needs cell image == minimum image, which holds when
        // every dimension has >= 4 cells and cut < 2 cells (< L/2).
        const double Lmax = fmax(P.Lx, fmax(P.Ly, P.Lz));
        const double wmin = fmin(sys->lsubbox[0], fmin(sys->lsubbox[1], sys->lsubbox[2]));
        const double band = 4.2e-6 * cut * Lmax + 2e-6 * P.cut2;
        P.prefilter = c->prefilter && nx >= 4 && ny >= 4 && nz >= 4 && cut < 1.95 * wmin && band < 0.05 * P.cut2;
        P.fcut_lo = (float)(P.cut2 - band);
        P.fcut_hi = (float)(P.cut2 + band);
        if (P.prefilter) {
            const int nseg = (nx + TILE_CX - 1) / TILE_CX;
            k_build_tile<<<nseg * ny * nz, TILE_THREADS, 0, c->stream>>>(
                c->xs, c->xf, c->order, c->cell_start, c->excl_bond, c->excl_angle, c->excl_dihed,
                c->nbr, c->cnt, c->scal, P);
        } else {
            // small or degenerate grids (< 4 cells in a direction, oversized skin): exact FP64 test for
            // every candidate, one warp per atom
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
        if (c->scal_host->max_neighb <= c->cap) {
            c->list_valid = true; c->list_opt = opt; c->sorted_identity = false; c->xs_current = true;
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

// ---- pair export -----------------------------------------------------------------------------------------------
__global__ void k_export_pairs(const unsigned *__restrict__ nbr, const int *__restrict__ cnt,
                               const int *__restrict__ order, int n, int npad, int *__restrict__ out,
                               long long max_pairs, unsigned long long *counter)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = order[s];
    const int m = cnt[s];
    for (int k = 0; k < m; k++) {
        const int j = order[nbr[(size_t)k * npad + s] & SEPGPU_INDEX_MASK];
        if (i < j) {
            unsigned long long p = atomicAdd(counter, 1ULL);
            if ((long long)p < max_pairs) { out[2 * p] = i; out[2 * p + 1] = j; }
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
    k_export_pairs<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->nbr, c->cnt, c->order, c->n, c->npad, dout, max_pairs, dcount);
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
