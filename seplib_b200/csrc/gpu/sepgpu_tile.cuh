// sepgpu_tile.cuh -- the cell grid and the TILE decomposition shared by the Verlet-list builder, the
// tile force kernels and the pair export.
//
// Cells are the reference's grid (sys->nsubbox, source/sepinit.c:257-276) laid out brick-major: bricks of
// bx x 4 x 4 cells, x fastest inside a brick, so that an x-run of bx cells -- and R consecutive x-runs -- are one
// contiguous range of the cell-sorted atom arrays.
//
// A TILE is R = 1, 2 or 4 consecutive x-runs of one brick layer: bx x R x 1 home cells.  One CTA owns one tile.
// Its candidate cells are the (bx+2) x (R+2) x 3 cells around it; their atoms, shifted to the periodic image
// the tile sees, are staged ONCE into shared memory in a fixed order (z row, y row, x cell, position inside
// the cell).  The position of a staged atom in that order is its SLOT.  The builder writes neighbour rows as
// 16-bit slots (bit 15: the partner sits in a periodic image), the force kernel stages the same cells in the
// same order from the current coordinates and gathers neighbours from shared memory by slot -- no global
// gathers, no per-pair image arithmetic, half the list bytes.  Slots are valid until the next list build
// (cell membership, and therefore the order, only changes there).
#pragma once

#include "sepgpu_internal.cuh"

#define BRICK_YZ 4

__host__ __device__ __forceinline__ int cell_key(int cx, int cy, int cz, const CellGrid &G)
{
    const int bxi = cx / G.bx, lx = cx % G.bx;
    const int byi = cy / BRICK_YZ, ly = cy % BRICK_YZ;
    const int bzi = cz / BRICK_YZ, lz = cz % BRICK_YZ;
    return (((bzi * G.nby + byi) * G.nbx + bxi) * (BRICK_YZ * BRICK_YZ) + lz * BRICK_YZ + ly) * G.bx + lx;
}

__host__ __device__ __forceinline__ void key_cell(int key, const CellGrid &G, int &cx, int &cy, int &cz)
{
    const int lx = key % G.bx; key /= G.bx;
    const int ly = key % BRICK_YZ; key /= BRICK_YZ;
    const int lz = key % BRICK_YZ; key /= BRICK_YZ;
    const int bxi = key % G.nbx; key /= G.nbx;
    const int byi = key % G.nby; const int bzi = key / G.nby;
    cx = bxi * G.bx + lx; cy = byi * BRICK_YZ + ly; cz = bzi * BRICK_YZ + lz;
}

#define TILE_MAXBX   8
#define TILE_MAXR    4
#define TILE_THREADS 288                                              // 9 warps; three CTAs per SM
#define TILE_MAXCELLS ((TILE_MAXBX + 2) * (TILE_MAXR + 2) * 3)        // 180 candidate cells
#define TILE_MAXHOME  (TILE_MAXBX * TILE_MAXR)                        // 32 home cells
#define TILE_PAD 32                                                   // far-away pad slots behind the staged atoms
#define TILE_SLOT_MASK  0x7fffu
#define TILE_SLOT_IMAGE 0x8000u                                       // the partner was staged with a periodic shift
#define TILE_MAX_SLOTS  0x7fff

// per-CTA description of a tile, in shared memory
struct TileLayout {
    int off[TILE_MAXCELLS + 1];        // slot of the first atom of candidate cell c = (sz*nry + sy)*ncc + cc
    int beg[TILE_MAXCELLS];            // first sorted index of that cell
    unsigned char code[TILE_MAXCELLS]; // periodic image the tile sees it in: (wx+1) + 3(wy+1) + 9(wz+1); 13 = none
    int home[TILE_MAXHOME + 1];        // first sorted index of home cell h = hy*bx + hx
    int ncx, nry, ncc, ncell;          // home cells along x, staged y rows (home rows + 2), staged cells along x, all cells
    int a0, nhome, total;              // home atoms [a0, a0 + nhome), staged atoms
    int cy0, cz, x0;
    int any_image;                     // some candidate cell is a periodic image
};

// Fills T for tile blockIdx.x; every thread of the CTA must call it (it synchronises).  Returns false --
// uniformly -- for tiles that lie in the padding of the brick grid or hold no home atoms.
// Needs blockDim.x >= TILE_MAXCELLS.
__device__ __forceinline__ bool tile_layout(TileLayout &T, const CellGrid &G, int R, const int *__restrict__ cell_start)
{
    const int tid = threadIdx.x;
    const int key0 = blockIdx.x * R * G.bx;
    int x0, cy0, cz;
    key_cell(key0, G, x0, cy0, cz);
    if (x0 >= G.nx || cy0 >= G.ny || cz >= G.nz) return false;            // padding of the brick grid
    const int ncx = min(G.bx, G.nx - x0);
    const int nhy = min(R, G.ny - cy0);                                    // home rows inside the grid
    const int nry = nhy + 2, ncc = ncx + 2, ncell = 3 * nry * ncc;
    if (tid <= R * G.bx) T.home[tid] = cell_start[key0 + tid];
    if (tid < ncell) {
        const int cc = tid % ncc, sy = (tid / ncc) % nry, sz = tid / (ncc * nry);
        int mx = x0 - 1 + cc, wx = 0, my = cy0 - 1 + sy, wy = 0, mz = cz - 1 + sz, wz = 0;
        if (mx >= G.nx) { mx -= G.nx; wx = 1; } else if (mx < 0) { mx += G.nx; wx = -1; }
        if (my >= G.ny) { my -= G.ny; wy = 1; } else if (my < 0) { my += G.ny; wy = -1; }
        if (G.dd) {                       // slab: no wrap in the local layer index; the image follows the global layer
            const int gl = G.zoff + mz;
            wz = gl < 0 ? -1 : (gl >= G.nzg ? 1 : 0);
            mz = min(max(mz, 0), G.nz - 1);                                // (halo tiles never get here with mz out of range)
        } else if (mz >= G.nz) { mz -= G.nz; wz = 1; } else if (mz < 0) { mz += G.nz; wz = -1; }
        const int key = cell_key(mx, my, mz, G);
        const int b = cell_start[key];
        T.beg[tid] = b;
        T.off[tid] = cell_start[key + 1] - b;                              // length for now
        T.code[tid] = (unsigned char)((wx + 1) + 3 * (wy + 1) + 9 * (wz + 1));
    }
    if (tid == 0) {
        T.ncx = ncx; T.nry = nry; T.ncc = ncc; T.ncell = ncell;
        T.cy0 = cy0; T.cz = cz; T.x0 = x0;
    }
    __syncthreads();
    if (tid < 32) {                                                        // exclusive scan of <= 180 lengths by one warp
        int carry = 0, img = 0;
        for (int base = 0; base < ncell; base += 32) {
            const int idx = base + tid;
            const int v = idx < ncell ? T.off[idx] : 0;
            if (idx < ncell && v > 0 && T.code[idx] != 13) img = 1;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
            if (idx < ncell) T.off[idx] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        img = __any_sync(0xffffffffu, img);
        if (tid == 0) {
            T.off[ncell] = carry; T.total = carry;
            T.a0 = T.home[0]; T.nhome = T.home[R * G.bx] - T.home[0];
            T.any_image = img;
        }
    }
    __syncthreads();
    return T.nhome > 0;
}

// candidate cell holding slot q (last cell whose first slot is <= q)
__device__ __forceinline__ int tile_cell_of_slot(const TileLayout &T, int q)
{
    int lo = 0, hi = T.ncell - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (T.off[mid] <= q) lo = mid; else hi = mid - 1; }
    return lo;
}

// home cell of sorted atom s (a home atom of this tile): h = hy*bx + hx
__device__ __forceinline__ int tile_home_cell(const TileLayout &T, int s, int nh)
{
    int h = 0;
    for (int q = 1; q < nh; q++) h += (s >= T.home[q]);
    return h;
}

// 16-bit rows: 8 entries per 128-bit chunk, chunks transposed over the sorted atoms (one coalesced 512-byte
// read per warp and chunk)
__host__ __device__ __forceinline__ size_t nbr16_index(int k, int s, int npad)
{
    return ((size_t)(k >> 3) * npad + s) * 8 + (k & 7);
}
