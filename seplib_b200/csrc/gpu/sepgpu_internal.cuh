// sepgpu_internal.cuh -- device-side data model of seplib-b200 (sm_100a only).
//
// HBM layout (all FP64, 32-byte vectors so that one atom == one DRAM sector and every per-atom
// access is a single 256-bit LDG/STG, which sm_100a has natively):
//
//   primary state, ORIGINAL atom order (host indices, topology indices stay valid):
//     x4[i]  = {x, y, z, tag}       wrapped position; tag = type | (molindex+1)<<8 as raw bits
//     v4[i]  = {vx, vy, vz, m}
//     f4[i]  = {fx, fy, fz, 0}
//     xn4[i] = {xn, yn, zn, 0}      position at last list build (reference: seppart.xn)
//     cr4[i] = {cross_neighb[3], packed crossings since the list was built}
//     crossings[3i..]               total boundary crossings (written only when an atom wraps)
//   cell-sorted copy used by the list build and the force kernels:
//     xs[s]  = {xu, yu, zu, tag}    xu = x + (crossings since list build)*L : continuous between
//                                   rebuilds, so the image shift stored with a list entry stays valid
//     order[s] = i, rank[i] = s
//   Verlet list: FULL list (i->j and j->i), transposed in chunks of 4 entries so that a lane reads its
//   next four entries with one 128-bit load and a warp reads 512 contiguous bytes:
//     nbr[((k/4)*npad + s)*4 + k%4] = j_sorted | shift_code << 26,   cnt[s]
//
// Reference counterparts are cited at each kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "sepgpu.h"

struct __align__(32) d4 { double x, y, z, w; };
struct __align__(16) i4 { int x, y, z, w; };

#define SEPGPU_SHIFT_BITS 26
#define SEPGPU_INDEX_MASK ((1u << SEPGPU_SHIFT_BITS) - 1u)
#define SEPGPU_MAX_ATOMS  (1u << SEPGPU_SHIFT_BITS)
// number of doubles one block writes as its partial result
#define SEPGPU_NPART_F 8      // force kernels : e, ecoul, 6 virial
#define SEPGPU_NPART_I 12     // integrators   : sum m vh^2, 6 kin_P, max d2, sum m v^2, 3 momentum
#define SEPGPU_MAX_BLOCKS_PARTIAL 65536
#define SEPGPU_NSUB 4         // typed sub-lists kept per list build

// device-resident scalar block (one per context)
struct DevScalars {
    double epot, ecoul, ekin;
    double pot_P[9], kin_P[9], pot_P_bond[9];
    double max_dist2;
    double sum_mv2;
    double alpha[8];           // 0..3: sep_nosehoover multipliers (host slots); 4..6: history of _sep_nosehoover_type
    double mom[4];             // sum m v (3) + sum m, per sepgpu_reset_momentum
    int neighb_flag;           // skin trigger fired in the LAST integrator call
    int nbuild;
    int error;
    int max_neighb;
    long long npairs_listed;
    int stage_needed;          // list build: largest candidate count a tile wanted to stage (host grows the buffer)
    int max_half;              // longest reference-style half list at the last build
    int aliased_seen;          // list build: an atom outside [0,L) was filed under an aliased cell (reference behaviour)
    int stage_used;            // list build: largest candidate count a tile staged (sizes the tile force kernels' shared memory)
    int error_where;           // which bounded wait gave up (SEPGPU_ENCCL): 1 migration counts, 2 migration records, 3 halo unpack,
                               // 4 halo wait inside the tile force kernel, 5 the finaliser's all-gather
};

struct KernelTimer {
    cudaEvent_t start[64], stop[64];
    int used;
    float total_ms;
    int launches;
    bool enabled;
};

struct CellGrid {
    int nx, ny, nz;         // reference cell grid (sys->nsubbox)
    int bx;                 // brick extent along x (1,2,4,8); y and z extents are 4 cells (sepgpu_tile.cuh)
    int nbx, nby, nbz;      // bricks per direction
    // slab decomposition along z (sepgpu_dd.cu): nz above is the number of LOCAL layers (owned layers
    // plus one halo layer on each side); local layer l holds global layer zoff + l (mod nzg)
    int dd, zoff, nzg;
};

struct DDState;            // slab domain decomposition (sepgpu_dd.cu); NULL when the context is not decomposed
struct FeedState;          // sampler feeds (sepgpu_feeds.cu); NULL until a feed is asked for

// parameters of one Lennard-Jones-family pair call, as the kernels take them (sepgpu_pair.cuh has the arithmetic)
struct LJDev {
    double cf2, sig2, eps48, eps4, aw, awh, shift;
    int t0, t1;
    // tabulated pair function (user callbacks of sep_force_pairs): n samples {f, u} on a uniform r^2 grid [t_lo, t_lo + (n-1)/t_inv];
    // NULL = the Lennard-Jones family above
    const double2 *tab;
    double t_lo, t_inv;
    int t_n;
};
struct BoxDev { double Lx, Ly, Lz; };

struct sepgpu_ctx {
    int n;                 // atoms in the sorted arrays: owned + halo (== n_own when not decomposed)
    int n_own;             // atoms this context integrates
    int ncap;              // allocation size of the per-atom arrays
    long long n_global;    // atoms of the whole system (sep_nosehoover divides by it)
    int npad;              // ncap rounded up to 32: row stride of the neighbour list
    DDState *dd;
    FeedState *feeds;
    int *gid;              // global atom id per local atom (decomposed runs only)
    int device;
    cudaStream_t stream;

    d4 *x4, *v4, *f4, *xn4, *pv4, *pa4;
    d4 *x0;                // tether positions (sep_set_x0), allocated on first upload
    d4 *prevf4, *randn4;   // sep_langevinGJF: previous force and previous noise per atom, allocated on first use
    i4 *cr4;
    int *crossings;        // 3n
    double *z;             // charges
    char *type;
    int *molindex;
    double *zs;            // charges in cell-sorted order (valid while zs_valid)
    bool zs_valid;
    int *excl_bond, *excl_angle, *excl_dihed;   // partner tables, n*10 / n*10 / n*20
    bool have_excl;
    bool have_charge;
    bool have_dpd;         // pv4/pa4 allocated

    // cell-sorted side
    d4 *xs;
    float4 *xf;            // cell-relative FP32 copy for the list prefilter (w = raw cell index)
    int *order, *rank;
    int *cell_of;          // [n] cell index of atom i
    unsigned char *cls;    // [n] decomposed runs: class of SORTED atom s: 0 interior, 1 next to a halo layer, 2 halo
    int *cell_cnt, *cell_start;   // [ncell_cap+1]
    int *tmp_slot;         // [n]
    int ncell_cap;
    int grid_n[3];         // cell grid the current list was built on

    // Verlet list
    unsigned *nbr;
    int *cnt;
    int cap;               // rows allocated (max neighbours per atom)
    bool list_valid;
    unsigned list_opt;
    bool sorted_identity;  // brute mode: xs is x4 in original order
    bool need_atom_rows;   // a consumer of global-index rows (list Coulomb, DPD, the molecule-pair table) has been seen
    bool list_f16;         // the list holds rows of 16-bit tile slots (sepgpu_tile.cuh) for the tile force kernels
    int build_window;      // option: x-window sweep in the tiled list builder (default 1)
    int tile_list;         // option: build 16-bit tile rows when no consumer needs global-index rows (default 1)
    CellGrid tile_grid;    // grid / tile shape of the current list (every build, both formats)
    int tile_R, tile_count, tile_stage_used, tile_R_max;
    int4 *tile_hdr;        // per tile: first home atom (sorted index), home atoms, staged atoms, image flag
    unsigned *tile_src;    // per tile and slot: sorted index | image code << 26 (the staging order), stride tile_stride
    size_t tile_hdr_cap, tile_src_cap;
    int tile_stride;
    bool moved_since_build; // an integrator ran since the list was built
    long long list_gen;    // bumped by every successful list build (keys the derived lists below)

    // typed sub-lists (option typed_sublist): for a typed Lennard-Jones call ("OO" in water) the entries of the
    // full list whose partner has the wanted type, same chunked layout, rebuilt lazily once per list build
    unsigned *nbr_t[SEPGPU_NSUB];
    int *cnt_t[SEPGPU_NSUB];
    int sub_key[SEPGPU_NSUB];        // t0 | t1 << 8 with t0 <= t1; 0 = slot unused
    long long sub_gen[SEPGPU_NSUB];  // list_gen the slot was filled for
    int sub_cap[SEPGPU_NSUB];        // rows allocated (== cap at allocation time)
    unsigned char *tsort;            // type char per sorted slot
    long long tsort_gen;
    d4 *xq;                          // coulomb_kernel 2: {continuous x, y, z, charge} per sorted slot, refreshed per call

    // topology
    unsigned *blist, *alist, *dlist;
    unsigned nb, na, nd;
    int *atom_bond_ptr, *atom_bond_idx;     // inverse topology (CSR): term*4+role
    int *atom_angle_ptr, *atom_angle_idx;
    int *atom_dihed_ptr, *atom_dihed_idx;
    double *blengths, *angles, *dihedrals;

    // molecule-molecule force table (reference sepmolinfo.Fij), nmol*nmol*3 doubles; NULL unless enabled
    double *fij;
    int nmol;

    // scalars and partial sums
    DevScalars *scal;            // device
    DevScalars *scal_host;       // pinned mirror
    double *partial;             // [SEPGPU_MAX_BLOCKS_PARTIAL * 16]

    // state flags
    bool f_zero;                 // sep_reset_force seen, no force kernel since: first kernel stores, not adds
    int  pending_alpha_slot;     // >=0: f -= alpha[slot] m v still to be applied by the integrator
    int  pending_alpha_type;     // -1 all atoms, else restrict to this type char
    bool xs_current;             // xs matches x4 (brute mode / after host put)
    bool ret_reset_pending;      // sep_reset_retval seen: the next kernel that touches the scalar block clears it first
    bool maxd_reset_pending;     // sep_reset_force seen: the next integrator starts max_dist2 from zero

    // staging
    void *stage; size_t stage_bytes;     // pinned host
    void *dstage; size_t dstage_bytes;   // device

    int single_type;             // type char shared by all atoms, or -1 (mixed)
    bool mv2_valid;              // scal->sum_mv2 matches the velocities now in v4

    // options
    int tpa;                     // lanes per atom in list force kernels
    int prefilter;               // FP32 prefilter in list build (1) or exact FP64 everywhere (0)
    int force_grid;              // CTAs of the list force kernel (0 = default)
    int tile_stage_cap;          // candidates the tiled list builder can stage per CTA (grows on demand)
    // option step_fold: the final reduction of the last force routine of a step and the Nose-Hoover multiplier update are
    // not launched on their own but folded into the integrator (k_integrate<.., NHFOLD> + k_finalize_both): 3 kernels per
    // Lennard-Jones step instead of 5.  Anything else that enters the library first settles what is pending.
    int step_fold;
    struct { bool active; int nrows; double scale; int flags; } fin_pending;
    struct { bool active; int slot; double temp0, tau; } nh_pending;
    double nh_dt;
    int fin_multi;               // multi-CTA final reduction of the force partial rows (0 = off, default)
    unsigned *fin_ticket;        // its ticket counter
    int coulomb_kernel;          // 1: first list Coulomb kernel (hardware-verified default); 2: k_coulomb_list2
    int typed_sublist;           // typed Lennard-Jones calls walk a per-type sub-list (0 = off, default)
    int overlap;                 // decomposed runs: halo refresh beside an interior-only force pass (default 0: measured slower,
                                 // the boundary pass is a nearly empty wave that costs more than the 20 us refresh)

    // tabulated pair function (sepgpu_force_table): device copy and what it was made from
#define SEPGPU_NTAB 4
    struct { void *dev; const void *key; int n; double lo, cf; unsigned long long hash; } tab[SEPGPU_NTAB];
    unsigned tab_next;

    const int *host_rows;          // sepgpu_set_host_rows

    // Speculative force launch (option spec_force): after an integrator call the first force call of the previous steps
    // is launched again at once, behind the finaliser and guarded by the device-side rebuild flag, into a second force
    // array; the host reads the flag beside it on its own stream.  The next sepgpu_force_lj with the same arguments adopts
    // the launch (swaps the force arrays) instead of launching -- the device never waits for the host's decision.
    struct SpecForce {
        int on, streak;
        bool launched, cancelled;
        LJDev P; BoxDev B; bool typed;
        char types[2]; unsigned opt; int epot_assign;
        sepgpu_sys sys;
        int nrows, list_gen;
        unsigned long long api_seq;
    } spec;
    d4 *f4_alt;
    int *put_changed;              // device flag: an upload that may leave the list alone found different values (sepgpu_put_fields)
    bool put_check;                // ... and that flag has to be read before the upload returns
    d4 *f_side;                    // forces of ONE bonded term kind on their own (sepgpu_bonded_side: the reference's sep_omp_bond family)
    double *partial_side;          // ... and the block sums that kernel leaves behind (not used)
    long long spec_adopted;        // launches adopted so far (tests, sepgpu_get_option)
    long long feed_calls, get_calls;   // sampler feeds / per-atom downloads served so far (tests, sepgpu_get_option)
    bool scal_cache_valid;         // scal_host holds the block as the last integrator's finaliser left it ...
    unsigned long long scal_cache_seq;     // ... and nothing but readers has entered the library since
    unsigned long long api_seq;    // entries into the library that may change state (SEPGPU_ENTER; readers take themselves out)
    cudaStream_t flag_stream;
    cudaEvent_t ev_fin;

    // measurement
    cudaEvent_t ev0, ev1;
    KernelTimer t_force, t_build, t_intgr, t_coul, t_bonded, t_halo, t_migr;
    void *flush_buf; size_t flush_bytes;
};

// CPU kernel emulator (tests/emu): scattered loads marked with this macro are counted as warp-wide requests and the
// 128-byte lines they touch; compiles to nothing in the product
#ifdef SEPGPU_EMU
#define SEPGPU_EMU_GATHER(p) emu::record_gather(p)
#else
#define SEPGPU_EMU_GATHER(p) ((void)0)
#endif

// ---- error plumbing -------------------------------------------------------------------------------
void sepgpu_set_error(const char *fmt, ...);
#define CUDA_TRY(call)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            sepgpu_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                  \
                             cudaGetErrorString(_e));                                       \
            return SEPGPU_ECUDA;                                                            \
        }                                                                                   \
    } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())
// entry points: select the device and, with option step_fold, launch whatever an earlier call left pending
int sepgpu_settle(sepgpu_ctx *c);
int sepgpu_nh_update_now(sepgpu_ctx *c);
int sepgpu_dd_uses_p2p(sepgpu_ctx *c);             // sepgpu_dd.cu: decomposed run on the peer-memory path
#define SEPGPU_BENIGN(c) ((c)->api_seq--)          /* after SEPGPU_ENTER in entries that only read or set lazy flags */
#define SEPGPU_ENTER(c)                                                                     \
    do {                                                                                    \
        (c)->api_seq++;                                                                     \
        CUDA_TRY(cudaSetDevice((c)->device));                                               \
        if ((c)->fin_pending.active || (c)->nh_pending.active) {                            \
            int _rs = sepgpu_settle(c);                                                     \
            if (_rs) return _rs;                                                            \
        }                                                                                   \
    } while (0)

// ---- small device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ long long tag_bits(double w) { return __double_as_longlong(w); }
__device__ __forceinline__ int tag_type(double w) { return (int)(tag_bits(w) & 0xff); }
__device__ __forceinline__ int tag_mol(double w) { return (int)((tag_bits(w) >> 8) & 0xffffffffLL) - 1; }
__device__ __forceinline__ double make_tag(int type, int molindex)
{
    long long b = (long long)(type & 0xff) | ((long long)(unsigned)(molindex + 1) << 8);
    return __longlong_as_double(b);
}

// position of entry k of sorted atom s in the chunked, transposed neighbour array
__host__ __device__ __forceinline__ size_t nbr_index(int k, int s, int npad)
{
    return ((size_t)(k >> 2) * npad + s) * 4 + (k & 3);
}

// sep_Wrap (include/sepmisc.h:81-85), exact branch form
__device__ __forceinline__ double wrap_exact(double d, double len, double half)
{
    if (d > half) d -= len;
    else if (d < -half) d += len;
    return d;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block reduction of NV per-thread values; result valid in thread 0.  Fixed tree => deterministic.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* NV * BLOCK/32 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double s = warp_sum(v[q]);
        if (lane == 0) smem[q * (BLOCK / 32) + wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double s = lane < BLOCK / 32 ? smem[q * (BLOCK / 32) + lane] : 0.0;
            v[q] = warp_sum(s);
        }
    }
}

// decomposed runs, peer-memory all-gather of the integrator's per-rank sums (sepgpu_dd.cu sets it up,
// k_finalize_intgr_p2p in sepgpu_intgr.cu uses it)
#define SEPGPU_GATHER_W 16
// Bound of the peer-memory spin waits, in SM clocks (~60 s).  A peer may legitimately be late by seconds (host work
// between steps on one rank); only a dead peer takes longer, and then the sticky device error is set instead of hanging.
#define SEPGPU_SPIN_LIMIT 120000000000LL
struct GatherDev {
    unsigned char **bases;       // device array [nranks]: every rank's shared block as mapped here (own block included)
    size_t gather_off;           // double gather[2][nranks][SEPGPU_GATHER_W]
    size_t gflag_off;            // unsigned long long gflag[2][nranks]
    unsigned long long seq;
    int rank, nranks;
};

// decomposed runs on the peer-memory path: where a kernel finds the neighbours' freshly pushed boundary coordinates
// (sepgpu_dd.cu fills it; the tile force kernels wait on the flags themselves and read the buffers directly)
struct HaloArgs {
    const d4 *in0, *in1;                 // my from-hi / from-lo receive buffers (written by the neighbours over NVLink)
    int n0, n_own;                       // halo atoms from hi (the ones from lo follow); local atoms n_own.. are the halo
    const unsigned long long *flags;     // [0] raised by the hi neighbour, [1] by the lo neighbour
    unsigned long long seq;              // refresh number to wait for; 0 = nothing to wait for (not decomposed)
    int rot;                             // CTA b works on tile (b + rot) mod gridDim: the brick layers next to the halo come last
    const int *cancel;                   // speculative launch: every CTA leaves at once when this flag is set (rebuild requested)
};
int sepgpu_dd_halo_args(sepgpu_ctx *c, const sepgpu_sys *sys, HaloArgs *out);

// internal cross-file entry points
int sepgpu_ensure_stage(sepgpu_ctx *c, size_t bytes);
int sepgpu_apply_pending(sepgpu_ctx *c);          // flush a deferred thermostat update into f4
int sepgpu_refresh_xs_identity(sepgpu_ctx *c);    // brute mode: xs <- x4
void ktimer_begin(sepgpu_ctx *c, KernelTimer *t);
void ktimer_end(sepgpu_ctx *c, KernelTimer *t);
