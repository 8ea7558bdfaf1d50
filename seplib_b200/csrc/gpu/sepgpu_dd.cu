// sepgpu_dd.cu -- spatial domain decomposition across the GPUs of one box: one process per GPU,
// slabs of whole cell layers along z, halo exchange and migration over NVLink with NCCL point-to-point.
//
// The reference has nothing of the kind (single address space, OpenMP only).
// What must be preserved is its semantics on GLOBAL quantities:
//   * the neighbour-pair set: every rank bins its owned + halo atoms with the reference's exact
//     expression on the GLOBAL cell grid (source/sepprfrc.c:404-406) and accepts pairs with the exact
//     test, so the union over ranks is bit-identical to the single-GPU / reference set;
//   * sep_nosehoover's temperature (sum m v^2 over ALL atoms / global npart, source/sepintgr.c:152-159)
//     and the skin trigger (max displacement over ALL atoms, :72): one small all-reduce per step.
//
// Layout: rank r owns global cell layers [z0,z1); its local cell grid has (z1-z0)+2 layers, the first
// and last holding copies ("halo") of the neighbouring ranks' boundary layers.  Local per-atom arrays
// hold the owned atoms first, halo atoms after them.  Every time step the owners send the continuous
// coordinates of their boundary-layer atoms (32 B per atom) to both neighbours; at a list rebuild atoms
// that changed slab migrate with their full state and the halo membership is re-established.
// Forces use the full list, so each rank computes the force on its own atoms only and nothing flows
// back.  Compaction uses prefix sums (no atomics), so a run is reproducible bit for bit.
//
// NCCL is opened with dlopen at first use: the single-GPU library has no NCCL dependency.
#include "sepgpu_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

int sepgpu_exclusive_scan(cudaStream_t st, int *cnt, int *start, int *scratch, int n);
int sepgpu_dd_before_positions_change(sepgpu_ctx *c);

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load(void)
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (int k = 0; k < 2 && !g_nccl.lib; k++) g_nccl.lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.lib) { sepgpu_set_error("domain decomposition: cannot dlopen libnccl.so.2 (%s)", dlerror()); return SEPGPU_ENCCL; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name); if (!g_nccl.field) { sepgpu_set_error("NCCL symbol %s missing", name); return SEPGPU_ENCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}

#define NCCL_TRY(call)                                                                         \
    do {                                                                                       \
        ncclResult_t _r = (call);                                                              \
        if (_r != ncclSuccess) {                                                               \
            sepgpu_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); \
            return SEPGPU_ENCCL;                                                               \
        }                                                                                      \
    } while (0)

#define REC_D4 4        // migration record: x4, v4, xn4, {cross_neighb[3], crossings[3], gid, list-crossings}

struct DDState {
    int rank, nranks;
    ncclComm_t comm;
    int nzg, z0, z1;            // global layers, owned range [z0,z1)
    int lo_rank, hi_rank;
    double *comm_buf;           // device scratch for all-reduces [64]
    double *comm_host;          // pinned
    // classification + scans
    int *flag[5], *pos[5];      // stay, to_lo, to_hi, halo_lo, halo_hi
    int *scan_scratch;
    int *counts_dev, *counts_host;     // [16]
    // double buffers for compaction
    d4 *x4b, *v4b, *xn4b; i4 *cr4b; int *crossb, *gidb;
    // transfer buffers (device)
    d4 *send[2], *recv[2];      // [bufcap * REC_D4]
    size_t bufcap;              // atoms per direction
    int *send_idx[2];           // per-step halo send lists (local indices of owned atoms), lo and hi
    int n_send[2], n_recv[2];   // halo atoms sent to lo/hi, received from hi/lo (in that order)
    bool halo_current;
    // second stream: the per-step halo refresh (NCCL path) runs beside the interior force pass
    cudaStream_t stream2;
    cudaEvent_t ev_ready, ev_halo;
    bool halo_inflight;
    bool push_pending;          // a push kernel on stream2 may still be reading x4 / cr4
    // peer-memory halo push (NVLink P2P through CUDA IPC).  Every rank owns one block
    //   [from-hi buffer: bufcap d4][from-lo buffer: bufcap d4][flags: 2 x u64]
    // that both neighbours map; the pack kernel of a neighbour stores straight into it and then raises the flag.
    bool p2p;
    unsigned char *ipc_base;
    size_t ipc_flag_off;
    void *peer_base[2];               // mapping of the lo / hi neighbour's block (the same pointer when they coincide)
    d4 *peer_dst[2];                  // lo neighbour's from-hi buffer, hi neighbour's from-lo buffer
    unsigned long long *peer_flag[2];
    d4 *p2p_recv[2];                  // my from-hi / from-lo buffers
    unsigned long long *p2p_flag;     // my flags: [0] raised by the hi neighbour, [1] by the lo neighbour
    size_t peer_cap[2];               // atoms the lo / hi neighbour's buffers hold
    unsigned long long seq;           // halo refreshes so far (identical on all ranks)
    unsigned int *done_ctr;
    // migration through the same blocks (peer-memory path): count mailboxes and "records delivered" flags in the header
    size_t mig_off;                   // MigBox box[2] (written by the hi / lo neighbour), then u64 flag2[2]
    unsigned long long mseq;          // rebuilds so far (identical on all ranks)
    // all ranks' blocks (for the all-gather of the integrator sums)
    void *peer_all[64];
    unsigned char **bases_dev;
    size_t gather_off, gflag_off;
    unsigned long long gseq;
    // bonded terms: global id -> local row of the last rebuild
    int *gmap;
    int gmap_gen;
    // charges by global id (they never change): the local copy follows ownership and halo membership at every rebuild
    double *zglob;
    int z_gen;
};

extern "C" int sepgpu_dd_unique_id(void *out128)
{
    if (!out128) return SEPGPU_EINVAL;
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

template <typename T>
static int dmalloc(T **p, size_t count)
{
    CUDA_TRY(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
    CUDA_TRY(cudaMemset(*p, 0, sizeof(T) * (count ? count : 1)));
    return 0;
}

// ---- peer-memory set-up ---------------------------------------------------------------------------------------
// One cudaMalloc block per rank, its IPC handle all-gathered over the NCCL communicator, the two neighbours'
// blocks mapped with cudaIpcOpenMemHandle (NVLink peer access).  All ranks agree on the outcome: when any
// rank cannot export or map a block, every rank keeps the NCCL send/recv path.  SEPGPU_DD_P2P=0 disables it.
static int p2p_setup(sepgpu_ctx *c)
{
    DDState *d = c->dd;
    d->p2p = false;
    const char *env = getenv("SEPGPU_DD_P2P");
    int want = !(env && env[0] == '0');
    // block layout: a fixed-size header (flags, gather table) followed by the two receive buffers.  Ranks own
    // different numbers of layers, so the buffer size is per rank and travels with the handle.
    const size_t buf_bytes = d->bufcap * sizeof(d4);
    d->ipc_flag_off = 0;
    d->gather_off = 256;
    d->gflag_off = d->gather_off + sizeof(double) * 2 * 64 * SEPGPU_GATHER_W;
    d->mig_off = 20480;                                   // behind the gather table and its flags (end at 17664)
    const size_t hdr = 32768;
    const size_t total = hdr + 2 * buf_bytes;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    int ok = want;
    if (ok) {
        if (cudaMalloc((void **)&d->ipc_base, total < ((size_t)4 << 20) ? ((size_t)4 << 20) : total) != cudaSuccess) { ok = 0; d->ipc_base = NULL; }
        else if (cudaMemset(d->ipc_base, 0, total) != cudaSuccess || cudaIpcGetMemHandle(&mine, d->ipc_base) != cudaSuccess) ok = 0;
        cudaGetLastError();
    }
    // all-gather {handle, ok}
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 16;
    unsigned char *dev = NULL, *host = (unsigned char *)calloc(d->nranks, rec);
    CUDA_TRY(cudaMalloc((void **)&dev, rec * d->nranks));
    memcpy(host + rec * d->rank, &mine, sizeof mine);
    host[rec * d->rank + sizeof mine] = (unsigned char)ok;
    { unsigned long long bb = buf_bytes; memcpy(host + rec * d->rank + sizeof mine + 8, &bb, 8); }
    CUDA_TRY(cudaMemcpy(dev + rec * d->rank, host + rec * d->rank, rec, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllGather(dev + rec * d->rank, dev, rec, ncclChar, d->comm, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(host, dev, rec * d->nranks, cudaMemcpyDeviceToHost));
    int all_ok = 1;
    for (int r = 0; r < d->nranks; r++) all_ok &= host[rec * r + sizeof mine];
    int mapped = all_ok;
    if (all_ok) {
        for (int r = 0; r < d->nranks && mapped; r++) {
            if (r == d->rank) { d->peer_all[r] = d->ipc_base; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host + rec * r, sizeof h);
            if (cudaIpcOpenMemHandle(&d->peer_all[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                d->peer_all[r] = NULL;
                mapped = 0;
            }
        }
        d->peer_base[0] = d->peer_all[d->lo_rank];
        d->peer_base[1] = d->peer_all[d->hi_rank];
    }
    // second agreement round: did everybody map both neighbours?
    double *flag = d->comm_buf + 150;
    double v = mapped ? 0.0 : 1.0;
    CUDA_TRY(cudaMemcpy(flag, &v, sizeof v, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclSum, d->comm, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(&v, flag, sizeof v, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    if (v != 0.0) {
        free(host);
        for (int r = 0; r < d->nranks; r++)
            if (r != d->rank && d->peer_all[r]) { cudaIpcCloseMemHandle(d->peer_all[r]); d->peer_all[r] = NULL; }
        d->peer_base[0] = d->peer_base[1] = NULL;
        return 0;                                         // NCCL path on every rank
    }
    if (dmalloc(&d->bases_dev, (size_t)d->nranks)) return SEPGPU_ECUDA;
    CUDA_TRY(cudaMemcpy(d->bases_dev, d->peer_all, sizeof(void *) * d->nranks, cudaMemcpyHostToDevice));
    d->gseq = 0;
    unsigned long long peer_bytes[2];
    memcpy(&peer_bytes[0], host + rec * d->lo_rank + sizeof mine + 8, 8);
    memcpy(&peer_bytes[1], host + rec * d->hi_rank + sizeof mine + 8, 8);
    d->p2p_recv[0] = (d4 *)(d->ipc_base + hdr);
    d->p2p_recv[1] = (d4 *)(d->ipc_base + hdr + buf_bytes);
    d->p2p_flag = (unsigned long long *)(d->ipc_base + d->ipc_flag_off);
    // what I send "down" is the lo neighbour's data "from above" and the other way round
    d->peer_dst[0] = (d4 *)((unsigned char *)d->peer_base[0] + hdr);
    d->peer_flag[0] = (unsigned long long *)((unsigned char *)d->peer_base[0] + d->ipc_flag_off);
    d->peer_dst[1] = (d4 *)((unsigned char *)d->peer_base[1] + hdr + peer_bytes[1]);
    d->peer_flag[1] = (unsigned long long *)((unsigned char *)d->peer_base[1] + d->ipc_flag_off) + 1;
    d->peer_cap[0] = peer_bytes[0] / sizeof(d4);
    d->peer_cap[1] = peer_bytes[1] / sizeof(d4);
    free(host);
    if (dmalloc(&d->done_ctr, 1)) return SEPGPU_ECUDA;
    d->seq = 0;
    d->p2p = true;
    return 0;
}

extern "C" int sepgpu_dd_init(sepgpu_ctx *c, int rank, int nranks, const void *id128, const sepgpu_sys *sys,
                              long long n_global)
{
    if (!c || !id128 || !sys || nranks < 2 || rank < 0 || rank >= nranks) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("dd_init: already decomposed"); return SEPGPU_ESTATE; }
    int rc = nccl_load();
    if (rc) return rc;
    SEPGPU_ENTER(c);
    const int nzg = sys->nsubbox[2];
    if (nzg < 2 * nranks || nzg < 4) {
        sepgpu_set_error("dd_init: %d cell layers along z cannot be split over %d ranks", nzg, nranks);
        return SEPGPU_EINVAL;
    }
    DDState *d = (DDState *)calloc(1, sizeof(DDState));
    d->rank = rank; d->nranks = nranks; d->nzg = nzg;
    d->z0 = (int)((long long)rank * nzg / nranks);
    d->z1 = (int)((long long)(rank + 1) * nzg / nranks);
    d->lo_rank = (rank + nranks - 1) % nranks;
    d->hi_rank = (rank + 1) % nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&d->comm, nranks, id, rank));
    const size_t cap = (size_t)c->ncap;
    // a boundary layer holds about n_own / (owned layers); leave generous head-room
    d->bufcap = cap / (size_t)(d->z1 - d->z0) * 2 + 4096;
    if (d->bufcap > cap) d->bufcap = cap;
    if (nranks > 64) { sepgpu_set_error("dd_init: at most 64 ranks"); return SEPGPU_EINVAL; }
    if (dmalloc(&d->comm_buf, 160)) return SEPGPU_ECUDA;
    CUDA_TRY(cudaMallocHost((void **)&d->comm_host, sizeof(double) * 64));
    for (int k = 0; k < 5; k++) if (dmalloc(&d->flag[k], cap + 1) || dmalloc(&d->pos[k], cap + 1)) return SEPGPU_ECUDA;
    if (dmalloc(&d->scan_scratch, cap / 2048 + 1030) || dmalloc(&d->counts_dev, 16)) return SEPGPU_ECUDA;
    CUDA_TRY(cudaMallocHost((void **)&d->counts_host, sizeof(int) * 16));
    if (dmalloc(&d->x4b, cap) || dmalloc(&d->v4b, cap) || dmalloc(&d->xn4b, cap) || dmalloc(&d->cr4b, cap) ||
        dmalloc(&d->crossb, 3 * cap) || dmalloc(&d->gidb, cap)) return SEPGPU_ECUDA;
    for (int k = 0; k < 2; k++)
        if (dmalloc(&d->send[k], d->bufcap * REC_D4) || dmalloc(&d->recv[k], d->bufcap * REC_D4) ||
            dmalloc(&d->send_idx[k], d->bufcap)) return SEPGPU_ECUDA;
    if (!c->gid && dmalloc(&c->gid, cap)) return SEPGPU_ECUDA;
    c->dd = d;
    if ((rc = p2p_setup(c))) { c->dd = NULL; return rc; }
    CUDA_TRY(cudaStreamCreateWithFlags(&d->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
    c->dd = d;
    c->n_global = n_global;
    return 0;
}

extern "C" int sepgpu_dd_set_owned(sepgpu_ctx *c, int n_own)
{
    if (!c || n_own < 0 || n_own > c->ncap) return SEPGPU_EINVAL;
    c->n_own = n_own;
    c->n = n_own;
    c->list_valid = false;
    return 0;
}

extern "C" int sepgpu_dd_layers(sepgpu_ctx *c, int *z0, int *z1, int *n_own, int *n_halo)
{
    if (!c || !c->dd) return SEPGPU_EINVAL;
    if (z0) *z0 = c->dd->z0;
    if (z1) *z1 = c->dd->z1;
    if (n_own) *n_own = c->n_own;
    if (n_halo) *n_halo = c->n - c->n_own;
    return 0;
}

void sepgpu_dd_destroy(sepgpu_ctx *c)
{
    DDState *d = c->dd;
    if (!d) return;
    if (d->stream2) { cudaStreamSynchronize(d->stream2); cudaStreamDestroy(d->stream2); }
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < d->nranks; r++)
        if (r != d->rank && d->peer_all[r]) cudaIpcCloseMemHandle(d->peer_all[r]);
    if (d->bases_dev) cudaFree(d->bases_dev);
    if (d->ipc_base) cudaFree(d->ipc_base);
    if (d->done_ctr) cudaFree(d->done_ctr);
    if (d->gmap) cudaFree(d->gmap);
    if (d->zglob) cudaFree(d->zglob);
    if (d->ev_ready) cudaEventDestroy(d->ev_ready);
    if (d->ev_halo) cudaEventDestroy(d->ev_halo);
    if (d->comm) g_nccl.CommDestroy(d->comm);
    void *ptrs[] = {d->comm_buf, d->scan_scratch, d->counts_dev, d->x4b, d->v4b, d->xn4b, d->cr4b, d->crossb, d->gidb,
                    d->send[0], d->send[1], d->recv[0], d->recv[1], d->send_idx[0], d->send_idx[1],
                    d->flag[0], d->flag[1], d->flag[2], d->flag[3], d->flag[4], d->pos[0], d->pos[1], d->pos[2], d->pos[3], d->pos[4]};
    for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++) if (ptrs[i]) cudaFree(ptrs[i]);
    if (d->comm_host) cudaFreeHost(d->comm_host);
    if (d->counts_host) cudaFreeHost(d->counts_host);
    free(d);
    c->dd = NULL;
}

double *sepgpu_dd_comm(sepgpu_ctx *c) { return c->dd->comm_buf; }

// peer-memory all-gather available?  Fills the kernel argument and advances the sequence number.
bool sepgpu_dd_gather_next(sepgpu_ctx *c, GatherDev *g)
{
    DDState *d = c->dd;
    if (!d->p2p) return false;
    d->gseq++;
    g->bases = d->bases_dev; g->gather_off = d->gather_off; g->gflag_off = d->gflag_off;
    g->seq = d->gseq; g->rank = d->rank; g->nranks = d->nranks;
    return true;
}
int sepgpu_dd_uses_p2p(sepgpu_ctx *c) { return c->dd && c->dd->p2p ? 1 : 0; }
int sepgpu_dd_push_carveout_max(void);
int sepgpu_dd_nglobal(sepgpu_ctx *c) { return c->dd ? (int)c->n_global : c->n; }

// ---- global id -> local row (own and halo), for the bonded terms: rebuilt lazily after every list build -------------
__global__ void k_dd_gmap_fill(int *gmap, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) gmap[i] = -1; }
__global__ void k_dd_gmap_scatter(const int *__restrict__ gid, int first, int last, int *gmap)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < last) gmap[gid[i]] = i;
}

int sepgpu_dd_gmap(sepgpu_ctx *c, const int **gmap)
{
    DDState *d = c->dd;
    const int ng = (int)c->n_global;
    if (!d->gmap) { if (dmalloc(&d->gmap, (size_t)ng)) return SEPGPU_ECUDA; d->gmap_gen = -1; }
    if (d->gmap_gen != c->list_gen) {
        // an atom can be own AND halo only in a slab that is its own neighbour through the periodic boundary (never with
        // >= 2 ranks); own rows are written last so that they win
        k_dd_gmap_fill<<<(ng + 255) / 256, 256, 0, c->stream>>>(d->gmap, ng);
        if (c->n > c->n_own)
            k_dd_gmap_scatter<<<(c->n - c->n_own + 255) / 256, 256, 0, c->stream>>>(c->gid, c->n_own, c->n, d->gmap);
        if (c->n_own)
            k_dd_gmap_scatter<<<(c->n_own + 255) / 256, 256, 0, c->stream>>>(c->gid, 0, c->n_own, d->gmap);
        KERNEL_CHECK();
        d->gmap_gen = c->list_gen;
    }
    *gmap = d->gmap;
    return 0;
}

// ---- charges in decomposed runs (sep_coulomb_sf) -------------------------------------------------------------------
// Every rank holds the charges of ALL atoms by global id (8 B per atom; they are constants of the run); the per-row array
// the Coulomb kernels read is refilled from it after every rebuild, for own and halo rows alike -- nothing to migrate.
extern "C" int sepgpu_dd_set_charges(sepgpu_ctx *c, const double *z_global)
{
    if (!c || !z_global) return SEPGPU_EINVAL;
    if (!c->dd) { sepgpu_set_error("dd_set_charges: call sepgpu_dd_init first"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    DDState *d = c->dd;
    const size_t ng = (size_t)c->n_global;
    if (!d->zglob && dmalloc(&d->zglob, ng)) return SEPGPU_ECUDA;
    CUDA_TRY(cudaMemcpyAsync(d->zglob, z_global, sizeof(double) * ng, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    bool any = false;
    for (size_t i = 0; i < ng && !any; i++) any = z_global[i] != 0.0;
    c->have_charge = any;                     // lists in global-index rows from the next build on
    c->zs_valid = false;
    c->list_valid = false;
    d->z_gen = -1;
    return 0;
}

__global__ void k_dd_gather_charges(const double *__restrict__ zglob, const int *__restrict__ gid, double *__restrict__ z, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = zglob[gid[i]];
}

// the per-row charges of the current ownership / halo membership; returns 1 when there are global charges at all
int sepgpu_dd_refresh_charges(sepgpu_ctx *c)
{
    DDState *d = c->dd;
    if (!d || !d->zglob) return 0;
    if (d->z_gen != c->list_gen) {
        if (c->n) k_dd_gather_charges<<<(c->n + 255) / 256, 256, 0, c->stream>>>(d->zglob, c->gid, c->z, c->n);
        KERNEL_CHECK();
        d->z_gen = c->list_gen;
        c->zs_valid = false;
    }
    return 1;
}

void sepgpu_dd_rank(sepgpu_ctx *c, int *rank, int *nranks) { *rank = c->dd->rank; *nranks = c->dd->nranks; }

int sepgpu_dd_allreduce(sepgpu_ctx *c, double *sum_buf, int nsum, double *max_buf, int nmax)
{
    DDState *d = c->dd;
    NCCL_TRY(g_nccl.GroupStart());
    if (nsum > 0) NCCL_TRY(g_nccl.AllReduce(sum_buf, sum_buf, (size_t)nsum, ncclDouble, ncclSum, d->comm, c->stream));
    if (nmax > 0) NCCL_TRY(g_nccl.AllReduce(max_buf, max_buf, (size_t)nmax, ncclDouble, ncclMax, d->comm, c->stream));
    NCCL_TRY(g_nccl.GroupEnd());
    return 0;
}

// ---- rebuild: classification, migration, halo membership -------------------------------------------------
__global__ void k_dd_classify(const d4 *__restrict__ x4, int n_own, double lsz, int nzg, int z0, int z1,
                              int *f_stay, int *f_lo, int *f_hi, DevScalars *scal)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int cz = (int)__ddiv_rn(x4[i].z, lsz);                 // reference binning (source/sepprfrc.c:406)
    int stay = 0, lo = 0, hi = 0;
    if (cz >= z0 && cz < z1) stay = 1;
    else if (cz == (z0 - 1 + nzg) % nzg) lo = 1;
    else if (cz == z1 % nzg) hi = 1;
    else scal->error = SEPGPU_ECELL;                             // moved more than one layer, or left the box
    f_stay[i] = stay; f_lo[i] = lo; f_hi[i] = hi;
}

__global__ void k_dd_halo_flags(const d4 *__restrict__ x4, int n_own, double lsz, int z0, int z1, int *f_hlo, int *f_hhi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int cz = (int)__ddiv_rn(x4[i].z, lsz);
    f_hlo[i] = cz == z0 ? 1 : 0;
    f_hhi[i] = cz == z1 - 1 ? 1 : 0;
}

__device__ __forceinline__ d4 pack_aux(const i4 &cr, const int *crossings, int gid)
{
    d4 a;
    a.x = __longlong_as_double(((long long)(unsigned)cr.x) | ((long long)(unsigned)cr.y << 32));
    a.y = __longlong_as_double(((long long)(unsigned)cr.z) | ((long long)(unsigned)cr.w << 32));
    a.z = __longlong_as_double(((long long)(unsigned)crossings[0]) | ((long long)(unsigned)crossings[1] << 32));
    a.w = __longlong_as_double(((long long)(unsigned)crossings[2]) | ((long long)(unsigned)gid << 32));
    return a;
}
__device__ __forceinline__ void unpack_aux(const d4 &a, i4 &cr, int *crossings, int &gid)
{
    long long b;
    b = __double_as_longlong(a.x); cr.x = (int)(b & 0xffffffffLL); cr.y = (int)(b >> 32);
    b = __double_as_longlong(a.y); cr.z = (int)(b & 0xffffffffLL); cr.w = (int)(b >> 32);
    b = __double_as_longlong(a.z); crossings[0] = (int)(b & 0xffffffffLL); crossings[1] = (int)(b >> 32);
    b = __double_as_longlong(a.w); crossings[2] = (int)(b & 0xffffffffLL); gid = (int)(b >> 32);
}

// stayers -> compacted copy; leavers -> send records (order = ascending local index: prefix sums)
__global__ void k_dd_split(const d4 *__restrict__ x4, const d4 *__restrict__ v4, const d4 *__restrict__ xn4,
                           const i4 *__restrict__ cr4, const int *__restrict__ crossings, const int *__restrict__ gid,
                           int n_own, const int *__restrict__ pos_stay, const int *__restrict__ pos_lo,
                           const int *__restrict__ pos_hi, d4 *x4b, d4 *v4b, d4 *xn4b, i4 *cr4b, int *crossb, int *gidb,
                           d4 *send_lo, d4 *send_hi, size_t bufcap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int ps = pos_stay[i], pl = pos_lo[i], ph = pos_hi[i];
    if (pos_stay[i + 1] > ps) {
        x4b[ps] = x4[i]; v4b[ps] = v4[i]; xn4b[ps] = xn4[i]; cr4b[ps] = cr4[i];
        crossb[3 * ps] = crossings[3 * i]; crossb[3 * ps + 1] = crossings[3 * i + 1]; crossb[3 * ps + 2] = crossings[3 * i + 2];
        gidb[ps] = gid[i];
    } else {
        const bool to_lo = pos_lo[i + 1] > pl;
        const int p = to_lo ? pl : ph;
        if ((size_t)p >= bufcap) return;
        d4 *rec = (to_lo ? send_lo : send_hi) + (size_t)p * REC_D4;
        rec[0] = x4[i]; rec[1] = v4[i]; rec[2] = xn4[i];
        rec[3] = pack_aux(cr4[i], crossings + 3 * i, gid[i]);
    }
}

__global__ void k_dd_unpack_migrants(const d4 *__restrict__ recv, int nrec, int dst0, d4 *x4, d4 *v4, d4 *xn4, i4 *cr4,
                                     int *crossings, int *gid)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrec) return;
    const d4 *rec = recv + (size_t)k * REC_D4;
    const int i = dst0 + k;
    x4[i] = rec[0]; v4[i] = rec[1]; xn4[i] = rec[2];
    i4 cr; int g;
    unpack_aux(rec[3], cr, crossings + 3 * i, g);
    cr4[i] = cr; gid[i] = g;
}

// halo membership at a rebuild: wrapped position + tag and the global id; also records the send list
__global__ void k_dd_pack_halo(const d4 *__restrict__ x4, const int *__restrict__ gid, int n_own,
                               const int *__restrict__ pos, d4 *send, int *send_idx, size_t bufcap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int p = pos[i];
    if (pos[i + 1] == p || (size_t)p >= bufcap) return;
    send[2 * (size_t)p] = x4[i];
    d4 a; a.x = __longlong_as_double((long long)gid[i]); a.y = a.z = a.w = 0.0;
    send[2 * (size_t)p + 1] = a;
    send_idx[p] = i;
}

__global__ void k_dd_unpack_halo(const d4 *__restrict__ recv, int nrec, int dst0, d4 *x4, d4 *v4, i4 *cr4, int *gid)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrec) return;
    const int i = dst0 + k;
    x4[i] = recv[2 * (size_t)k];
    gid[i] = (int)__double_as_longlong(recv[2 * (size_t)k + 1].x);
    d4 v; v.x = v.y = v.z = 0.0; v.w = 1.0; v4[i] = v;
    i4 z; z.x = z.y = z.z = z.w = 0; cr4[i] = z;
}

static int read_counts(sepgpu_ctx *c, int n)
{
    DDState *d = c->dd;
    CUDA_TRY(cudaMemcpyAsync(d->counts_host, d->counts_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

__global__ void k_dd_totals(const int *p0, const int *p1, const int *p2, int n, int *out)
{
    out[0] = p0[n]; out[1] = p1[n]; if (p2) out[2] = p2[n];
}

// exchange with both neighbours: what goes "down" (to lo_rank) arrives as the receiver's data "from above"
static int exchange(sepgpu_ctx *c, const void *send_lo, size_t n_lo, const void *send_hi, size_t n_hi,
                    void *recv_from_hi, size_t n_from_hi, void *recv_from_lo, size_t n_from_lo, size_t elem_bytes,
                    cudaStream_t st = 0)
{
    DDState *d = c->dd;
    if (!st) st = c->stream;
    NCCL_TRY(g_nccl.GroupStart());
    if (n_lo) NCCL_TRY(g_nccl.Send(send_lo, n_lo * elem_bytes, ncclChar, d->lo_rank, d->comm, st));
    if (n_hi) NCCL_TRY(g_nccl.Send(send_hi, n_hi * elem_bytes, ncclChar, d->hi_rank, d->comm, st));
    if (n_from_hi) NCCL_TRY(g_nccl.Recv(recv_from_hi, n_from_hi * elem_bytes, ncclChar, d->hi_rank, d->comm, st));
    if (n_from_lo) NCCL_TRY(g_nccl.Recv(recv_from_lo, n_from_lo * elem_bytes, ncclChar, d->lo_rank, d->comm, st));
    NCCL_TRY(g_nccl.GroupEnd());
    return 0;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- migration and halo membership through peer memory (no NCCL call, one host read per rebuild) ----------------
// Round 1: every rank classifies its atoms, counts who leaves towards which neighbour and who stays in a boundary layer,
// and stores those two numbers into the neighbour's mailbox; it then waits for the neighbours' numbers.  By the time a
// neighbour has posted its counts it has finished the previous step, so its receive buffers are free.
// Round 2: one kernel compacts the stayers, writes the leavers' full records and the boundary stayers' (position, id)
// STRAIGHT INTO the neighbours' receive buffers over NVLink and raises their "delivered" flags; a second kernel waits for
// both neighbours' flags and unpacks.  The atoms a rank has just sent away are exactly the halo atoms it still needs from
// that side, so they are appended to its halo locally and never travel back.
struct MigBox { unsigned long long seq; int n_mig, n_bnd; int pad[4]; };      // 32 bytes

// Classification and the five exclusive prefix sums it feeds (stayers, leavers to lo / hi, boundary stayers lo / hi) in
// three launches: every thread classifies its atoms on the fly, five warp scans run side by side.
#define S5_BLOCK 512
#define S5_ITEMS 2
struct Pos5 { int *p[5]; };

__device__ __forceinline__ void dd_classify(const d4 *__restrict__ x4, int i, double lsz, int nzg, int z0, int z1, int f[5], DevScalars *scal)
{
    const int cz = (int)__ddiv_rn(x4[i].z, lsz);                 // reference binning (source/sepprfrc.c:406)
    f[0] = f[1] = f[2] = 0;
    if (cz >= z0 && cz < z1) f[0] = 1;
    else if (cz == (z0 - 1 + nzg) % nzg) f[1] = 1;
    else if (cz == z1 % nzg) f[2] = 1;
    else scal->error = SEPGPU_ECELL;                             // moved more than one layer, or left the box
    f[3] = f[0] && cz == z0; f[4] = f[0] && cz == z1 - 1;
}

__global__ void __launch_bounds__(S5_BLOCK)
k_dd_scan5_local(const d4 *__restrict__ x4, int n, double lsz, int nzg, int z0, int z1, Pos5 P, int *__restrict__ block_sum,
                 int nblocks, DevScalars *scal)
{
    __shared__ int wsum[5][S5_BLOCK / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int base = (blockIdx.x * S5_BLOCK + threadIdx.x) * S5_ITEMS;
    int f[S5_ITEMS][5], tot[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < S5_ITEMS; q++) {
        if (base + q < n) dd_classify(x4, base + q, lsz, nzg, z0, z1, f[q], scal);
        else f[q][0] = f[q][1] = f[q][2] = f[q][3] = f[q][4] = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) tot[k] += f[q][k];
    }
    int incl[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int v = tot[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        incl[k] = v;
        if (lane == 31) wsum[k][wid] = v;
    }
    __syncthreads();
    if (wid < 5) {                                               // warp k scans the warp totals of array k
        const int w = lane < S5_BLOCK / 32 ? wsum[wid][lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        if (lane < S5_BLOCK / 32) wsum[wid][lane] = wi - w;
        if (lane == S5_BLOCK / 32 - 1) block_sum[wid * (nblocks + 1) + blockIdx.x] = wi;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int run = wsum[k][wid] + incl[k] - tot[k];
#pragma unroll
        for (int q = 0; q < S5_ITEMS; q++)
            if (base + q < n) { P.p[k][base + q] = run; run += f[q][k]; }
    }
}

// one block: exclusive scan of the block totals of each of the five arrays (<= 1024 blocks per pass), grand totals at [nblocks]
__global__ void __launch_bounds__(1024) k_dd_scan5_blocks(int *__restrict__ block_sum, int nblocks)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < 5; k++) {
        int *bs = block_sum + k * (nblocks + 1);
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base < nblocks; base += 1024) {
            const int idx = base + threadIdx.x;
            const int v = idx < nblocks ? bs[idx] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) wsum[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                const int w = wsum[lane];
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
                wsum[lane] = wi - w;
            }
            __syncthreads();
            const int excl = carry + wsum[wid] + incl - v;
            if (idx < nblocks) bs[idx] = excl;
            __syncthreads();
            if (threadIdx.x == 1023) carry = excl + v;
            __syncthreads();
        }
        if (threadIdx.x == 0) bs[nblocks] = carry;
        __syncthreads();
    }
}

__global__ void k_dd_scan5_apply(Pos5 P, const int *__restrict__ block_sum, int n, int nblocks)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int b = i / (S5_BLOCK * S5_ITEMS);
#pragma unroll
        for (int k = 0; k < 5; k++) P.p[k][i] += block_sum[k * (nblocks + 1) + b];
    }
    if (i < 5) P.p[i][n] = block_sum[i * (nblocks + 1) + nblocks];
}

// counts_dev: [0] stay [1] to_lo [2] to_hi [3] bnd_lo [4] bnd_hi | from the hi neighbour: [5] migrants [6] boundary stayers |
//             from the lo neighbour: [7] migrants [8] boundary stayers
__global__ void k_dd_post_counts(const int *p_stay, const int *p_lo, const int *p_hi, const int *p_blo, const int *p_bhi, int n,
                                 MigBox *lo_box_from_above, MigBox *hi_box_from_below, const MigBox *my_boxes,
                                 unsigned long long seq, int *counts, DevScalars *scal)
{
    if (threadIdx.x != 0) return;
    counts[0] = p_stay[n]; counts[1] = p_lo[n]; counts[2] = p_hi[n]; counts[3] = p_blo[n]; counts[4] = p_bhi[n];
    lo_box_from_above->n_mig = counts[1]; lo_box_from_above->n_bnd = counts[3];
    hi_box_from_below->n_mig = counts[2]; hi_box_from_below->n_bnd = counts[4];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&lo_box_from_above->seq), "l"(seq) : "memory");
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&hi_box_from_below->seq), "l"(seq) : "memory");
    const long long t0 = clock64();
    while (ld_acquire_sys(&my_boxes[0].seq) < seq || ld_acquire_sys(&my_boxes[1].seq) < seq) {
        if (clock64() - t0 > SEPGPU_SPIN_LIMIT) { scal->error = SEPGPU_ENCCL; scal->error_where = 1; break; }
        __nanosleep(100);
    }
    const volatile MigBox *b = my_boxes;
    counts[5] = b[0].n_mig; counts[6] = b[0].n_bnd; counts[7] = b[1].n_mig; counts[8] = b[1].n_bnd;
}

// peer_lo / peer_hi: the neighbours' receive buffers: [n_to * REC_D4 migrant records][n_bnd * 2 boundary records]
// keep_lo / keep_hi: (position, id) of my own leavers, which stay with me as halo atoms from that side
__global__ void k_dd_split2(const d4 *__restrict__ x4, const d4 *__restrict__ v4, const d4 *__restrict__ xn4,
                            const i4 *__restrict__ cr4, const int *__restrict__ crossings, const int *__restrict__ gid,
                            int n_own, const int *__restrict__ pos_stay, const int *__restrict__ pos_lo,
                            const int *__restrict__ pos_hi, const int *__restrict__ pos_blo, const int *__restrict__ pos_bhi,
                            d4 *x4b, d4 *v4b, d4 *xn4b, i4 *cr4b, int *crossb, int *gidb,
                            d4 *peer_lo, d4 *peer_hi, int n_to_lo, int n_to_hi, d4 *keep_lo, d4 *keep_hi,
                            int *send_idx_lo, int *send_idx_hi,
                            unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long seq, unsigned int *done)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_own) {
        const int ps = pos_stay[i];
        if (pos_stay[i + 1] > ps) {
            const d4 x = x4[i];
            x4b[ps] = x; v4b[ps] = v4[i]; xn4b[ps] = xn4[i]; cr4b[ps] = cr4[i];
            crossb[3 * ps] = crossings[3 * i]; crossb[3 * ps + 1] = crossings[3 * i + 1]; crossb[3 * ps + 2] = crossings[3 * i + 2];
            const int g = gid[i];
            gidb[ps] = g;
            d4 a; a.x = __longlong_as_double((long long)g); a.y = a.z = a.w = 0.0;
            const int bl = pos_blo[i], bh = pos_bhi[i];
            if (pos_blo[i + 1] > bl) {                   // boundary layer towards lo: halo atom of the lo neighbour
                d4 *rec = peer_lo + (size_t)n_to_lo * REC_D4 + 2 * (size_t)bl;
                rec[0] = x; rec[1] = a;
                send_idx_lo[bl] = ps;
            }
            if (pos_bhi[i + 1] > bh) {
                d4 *rec = peer_hi + (size_t)n_to_hi * REC_D4 + 2 * (size_t)bh;
                rec[0] = x; rec[1] = a;
                send_idx_hi[bh] = ps;
            }
        } else {
            const int pl = pos_lo[i];
            const bool to_lo = pos_lo[i + 1] > pl;
            const int p = to_lo ? pl : pos_hi[i];
            d4 *rec = (to_lo ? peer_lo : peer_hi) + (size_t)p * REC_D4;
            const d4 x = x4[i];
            const int g = gid[i];
            rec[0] = x; rec[1] = v4[i]; rec[2] = xn4[i];
            rec[3] = pack_aux(cr4[i], crossings + 3 * i, g);
            d4 *kp = (to_lo ? keep_lo : keep_hi) + 2 * (size_t)p;
            d4 a; a.x = __longlong_as_double((long long)g); a.y = a.z = a.w = 0.0;
            kp[0] = x; kp[1] = a;
        }
    }
    // The block's stores are ordered before thread 0's system-scope fence by the barrier (cumulativity): one fence per
    // block publishes them; the last block to finish raises both neighbours' flags
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(done, 1u);
        if (t == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_lo), "l"(seq) : "memory");
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_hi), "l"(seq) : "memory");
        }
    }
}

__device__ __forceinline__ d4 ld_peer_d4(const d4 *p)             // written by another GPU: never through L1
{
    const double2 *s = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldcg(s), b = __ldcg(s + 1);
    d4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    return v;
}

// Work items: [migrants from hi][migrants from lo][halo from hi: its boundary stayers, my leavers to hi][halo from lo: ...]
__global__ void k_dd_unpack2(const d4 *from_hi, const d4 *from_lo, const d4 *__restrict__ keep_lo, const d4 *__restrict__ keep_hi,
                             int n_stay, int m_hi, int m_lo, int b_hi, int b_lo, int n_to_hi, int n_to_lo, int nb_lo, int nb_hi,
                             d4 *x4, d4 *v4, d4 *xn4, i4 *cr4, int *crossings, int *gid, int *send_idx_lo, int *send_idx_hi,
                             const unsigned long long *flags2, unsigned long long seq, DevScalars *scal)
{
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags2) < seq || ld_acquire_sys(flags2 + 1) < seq) {
            if (clock64() - t0 > SEPGPU_SPIN_LIMIT) { scal->error = SEPGPU_ENCCL; scal->error_where = 2; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_own = n_stay + m_hi + m_lo;
    if (k < m_hi + m_lo) {                                        // migrants: full records
        const bool lo = k >= m_hi;
        const int q = lo ? k - m_hi : k;
        const d4 *rec = (lo ? from_lo : from_hi) + (size_t)q * REC_D4;
        const int i = n_stay + k;
        x4[i] = ld_peer_d4(rec); v4[i] = ld_peer_d4(rec + 1); xn4[i] = ld_peer_d4(rec + 2);
        i4 cr; int g;
        unpack_aux(ld_peer_d4(rec + 3), cr, crossings + 3 * i, g);
        cr4[i] = cr; gid[i] = g;
        // an atom that arrived from below sits in my lowest layer: the lo neighbour needs it as halo (and likewise above)
        if (lo) send_idx_lo[nb_lo + q] = i; else send_idx_hi[nb_hi + q] = i;
        return;
    }
    k -= m_hi + m_lo;
    const int h_hi = b_hi + n_to_hi, h_lo = b_lo + n_to_lo;
    if (k >= h_hi + h_lo) return;
    const bool lo = k >= h_hi;
    const int q = lo ? k - h_hi : k;
    const int nb = lo ? b_lo : b_hi;
    d4 x, a;
    if (q < nb) {                                                 // the neighbour's boundary stayers
        const d4 *rec = (lo ? from_lo + (size_t)m_lo * REC_D4 : from_hi + (size_t)m_hi * REC_D4) + 2 * (size_t)q;
        x = ld_peer_d4(rec); a = ld_peer_d4(rec + 1);
    } else {                                                      // the atoms I have just handed to that neighbour
        const d4 *rec = (lo ? keep_lo : keep_hi) + 2 * (size_t)(q - nb);
        x = rec[0]; a = rec[1];
    }
    const int i = n_own + k;
    x4[i] = x;
    gid[i] = (int)__double_as_longlong(a.x);
    d4 v; v.x = v.y = v.z = 0.0; v.w = 1.0; v4[i] = v;
    i4 z; z.x = z.y = z.z = z.w = 0; cr4[i] = z;
}

static int before_build_p2p(sepgpu_ctx *c, const sepgpu_sys *sys, int *zoff, int *nz_local)
{
    DDState *d = c->dd;
    const double lsz = sys->lsubbox[2];
    const int B = 256;
    int n_own = c->n_own;
    const int G = n_own ? (n_own + B - 1) / B : 1;
    int rc;
    d->mseq++;
    MigBox *my_boxes = (MigBox *)(d->ipc_base + d->mig_off);
    MigBox *lo_box = (MigBox *)((unsigned char *)d->peer_base[0] + d->mig_off);            // lo neighbour's "from above" box
    MigBox *hi_box = (MigBox *)((unsigned char *)d->peer_base[1] + d->mig_off) + 1;        // hi neighbour's "from below" box
    unsigned long long *my_flags2 = (unsigned long long *)(d->ipc_base + d->mig_off + 2 * sizeof(MigBox));
    unsigned long long *lo_flag2 = (unsigned long long *)((unsigned char *)d->peer_base[0] + d->mig_off + 2 * sizeof(MigBox));
    unsigned long long *hi_flag2 = (unsigned long long *)((unsigned char *)d->peer_base[1] + d->mig_off + 2 * sizeof(MigBox)) + 1;
    ktimer_begin(c, &c->t_migr);
    {
        Pos5 P5;
        for (int k = 0; k < 5; k++) P5.p[k] = d->pos[k];
        const int nb5 = n_own ? (n_own + S5_BLOCK * S5_ITEMS - 1) / (S5_BLOCK * S5_ITEMS) : 1;
        int *bs = d->flag[0];                                    // the flag arrays are free on this path: scratch for 5 x (nb5 + 1) block totals
        k_dd_scan5_local<<<nb5, S5_BLOCK, 0, c->stream>>>(c->x4, n_own, lsz, d->nzg, d->z0, d->z1, P5, bs, nb5, c->scal);
        k_dd_scan5_blocks<<<1, 1024, 0, c->stream>>>(bs, nb5);
        k_dd_scan5_apply<<<(n_own + 255) / 256 + 1, 256, 0, c->stream>>>(P5, bs, n_own, nb5);
    }
    k_dd_post_counts<<<1, 32, 0, c->stream>>>(d->pos[0], d->pos[1], d->pos[2], d->pos[3], d->pos[4], n_own, lo_box, hi_box, my_boxes,
                                             d->mseq, d->counts_dev, c->scal);
    if ((rc = read_counts(c, 9))) return rc;
    const int *h = d->counts_host;
    const int n_stay = h[0], n_to_lo = h[1], n_to_hi = h[2], nb_lo = h[3], nb_hi = h[4];
    const int m_hi = h[5], b_hi = h[6], m_lo = h[7], b_lo = h[8];          // arriving from hi / lo: migrants, boundary stayers
    const size_t need_lo = (size_t)n_to_lo * REC_D4 + 2 * (size_t)nb_lo, need_hi = (size_t)n_to_hi * REC_D4 + 2 * (size_t)nb_hi;
    const int n_new = n_stay + m_hi + m_lo;
    const int n_halo = b_hi + n_to_hi + b_lo + n_to_lo;
    if (need_lo > d->peer_cap[0] || need_hi > d->peer_cap[1] || (size_t)(nb_lo + m_lo) > d->bufcap || (size_t)(nb_hi + m_hi) > d->bufcap ||
        2 * (size_t)n_to_lo > d->bufcap * REC_D4 || 2 * (size_t)n_to_hi > d->bufcap * REC_D4 || n_new + n_halo > c->ncap) {
        sepgpu_set_error("decomposed rebuild: migration / halo exceeds the buffers (stay %d, out %d/%d, in %d/%d, halo %d, cap %d)",
                         n_stay, n_to_lo, n_to_hi, m_hi, m_lo, n_halo, c->ncap);
        return SEPGPU_EINVAL;
    }
    k_dd_split2<<<G, B, 0, c->stream>>>(c->x4, c->v4, c->xn4, c->cr4, c->crossings, c->gid, n_own, d->pos[0], d->pos[1], d->pos[2],
        d->pos[3], d->pos[4], d->x4b, d->v4b, d->xn4b, d->cr4b, d->crossb, d->gidb, d->peer_dst[0], d->peer_dst[1], n_to_lo, n_to_hi,
        d->send[0], d->send[1], d->send_idx[0], d->send_idx[1], lo_flag2, hi_flag2, d->mseq, d->done_ctr);
    { d4 *t; i4 *ti; int *tn;
      t = c->x4; c->x4 = d->x4b; d->x4b = t;  t = c->v4; c->v4 = d->v4b; d->v4b = t;  t = c->xn4; c->xn4 = d->xn4b; d->xn4b = t;
      ti = c->cr4; c->cr4 = d->cr4b; d->cr4b = ti;  tn = c->crossings; c->crossings = d->crossb; d->crossb = tn;
      tn = c->gid; c->gid = d->gidb; d->gidb = tn; }
    const int nwork = m_hi + m_lo + n_halo;
    k_dd_unpack2<<<nwork ? (nwork + B - 1) / B : 1, B, 0, c->stream>>>(d->p2p_recv[0], d->p2p_recv[1], d->send[0], d->send[1],
        n_stay, m_hi, m_lo, b_hi, b_lo, n_to_hi, n_to_lo, nb_lo, nb_hi, c->x4, c->v4, c->xn4, c->cr4, c->crossings, c->gid,
        d->send_idx[0], d->send_idx[1], my_flags2, d->mseq, c->scal);
    ktimer_end(c, &c->t_migr);
    KERNEL_CHECK();
    d->n_send[0] = nb_lo + m_lo; d->n_send[1] = nb_hi + m_hi;
    d->n_recv[0] = b_hi + n_to_hi; d->n_recv[1] = b_lo + n_to_lo;           // from hi, from lo
    c->n_own = n_new;
    c->n = n_new + n_halo;
    d->halo_current = true;
    d->halo_inflight = false;                            // a refresh pushed before this rebuild (a force launch sent ahead) is obsolete
    *zoff = d->z0 - 1;
    *nz_local = (d->z1 - d->z0) + 2;
    return 0;
}

int sepgpu_dd_before_build(sepgpu_ctx *c, const sepgpu_sys *sys, int *zoff, int *nz_local)
{
    DDState *d = c->dd;
    if (!c->f_zero) {
        sepgpu_set_error("decomposed rebuild: forces from an earlier call of this step would be lost "
                         "(call the list-building force routine first after sep_reset_force)");
        return SEPGPU_ESTATE;
    }
    if (sys->nsubbox[2] != d->nzg) { sepgpu_set_error("decomposed run: the cell grid along z changed"); return SEPGPU_ESTATE; }
    { int rw = sepgpu_dd_before_positions_change(c); if (rw) return rw; }
    if (d->p2p) return before_build_p2p(c, sys, zoff, nz_local);
    const double lsz = sys->lsubbox[2];
    const int B = 256;
    int n_own = c->n_own;
    int G = (n_own + B - 1) / B;
    int rc;
    ktimer_begin(c, &c->t_migr);
    // 1. who stays, who leaves
    if (G) k_dd_classify<<<G, B, 0, c->stream>>>(c->x4, n_own, lsz, d->nzg, d->z0, d->z1, d->flag[0], d->flag[1], d->flag[2], c->scal);
    for (int k = 0; k < 3; k++)
        if ((rc = sepgpu_exclusive_scan(c->stream, d->flag[k], d->pos[k], d->scan_scratch, n_own))) return rc;
    k_dd_totals<<<1, 1, 0, c->stream>>>(d->pos[0], d->pos[1], d->pos[2], n_own, d->counts_dev);
    // 2. neighbours learn how many atoms arrive
    if ((rc = exchange(c, d->counts_dev + 1, 1, d->counts_dev + 2, 1, d->counts_dev + 4, 1, d->counts_dev + 5, 1, sizeof(int)))) return rc;
    if ((rc = read_counts(c, 8))) return rc;
    const int n_stay = d->counts_host[0], n_to_lo = d->counts_host[1], n_to_hi = d->counts_host[2];
    const int n_from_hi = d->counts_host[4], n_from_lo = d->counts_host[5];
    if ((size_t)n_to_lo > d->bufcap || (size_t)n_to_hi > d->bufcap || (size_t)n_from_hi > d->bufcap || (size_t)n_from_lo > d->bufcap ||
        n_stay + n_from_hi + n_from_lo > c->ncap) {
        sepgpu_set_error("decomposed rebuild: migration exceeds the buffers (stay %d, out %d/%d, in %d/%d, cap %d)",
                         n_stay, n_to_lo, n_to_hi, n_from_hi, n_from_lo, c->ncap);
        return SEPGPU_EINVAL;
    }
    // 3. compact the stayers into the second buffer set, pack the leavers, swap buffers
    if (G) k_dd_split<<<G, B, 0, c->stream>>>(c->x4, c->v4, c->xn4, c->cr4, c->crossings, c->gid, n_own, d->pos[0], d->pos[1], d->pos[2],
                                               d->x4b, d->v4b, d->xn4b, d->cr4b, d->crossb, d->gidb, d->send[0], d->send[1], d->bufcap);
    { d4 *t; i4 *ti; int *tn;
      t = c->x4; c->x4 = d->x4b; d->x4b = t;  t = c->v4; c->v4 = d->v4b; d->v4b = t;  t = c->xn4; c->xn4 = d->xn4b; d->xn4b = t;
      ti = c->cr4; c->cr4 = d->cr4b; d->cr4b = ti;  tn = c->crossings; c->crossings = d->crossb; d->crossb = tn;
      tn = c->gid; c->gid = d->gidb; d->gidb = tn; }
    if ((rc = exchange(c, d->send[0], (size_t)n_to_lo * REC_D4, d->send[1], (size_t)n_to_hi * REC_D4,
                       d->recv[0], (size_t)n_from_hi * REC_D4, d->recv[1], (size_t)n_from_lo * REC_D4, sizeof(d4)))) return rc;
    if (n_from_hi) k_dd_unpack_migrants<<<(n_from_hi + B - 1) / B, B, 0, c->stream>>>(d->recv[0], n_from_hi, n_stay, c->x4, c->v4, c->xn4, c->cr4, c->crossings, c->gid);
    if (n_from_lo) k_dd_unpack_migrants<<<(n_from_lo + B - 1) / B, B, 0, c->stream>>>(d->recv[1], n_from_lo, n_stay + n_from_hi, c->x4, c->v4, c->xn4, c->cr4, c->crossings, c->gid);
    n_own = n_stay + n_from_hi + n_from_lo;
    G = (n_own + B - 1) / B;
    // 4. halo membership of the new owned set
    if (G) k_dd_halo_flags<<<G, B, 0, c->stream>>>(c->x4, n_own, lsz, d->z0, d->z1, d->flag[3], d->flag[4]);
    for (int k = 3; k < 5; k++)
        if ((rc = sepgpu_exclusive_scan(c->stream, d->flag[k], d->pos[k], d->scan_scratch, n_own))) return rc;
    k_dd_totals<<<1, 1, 0, c->stream>>>(d->pos[3], d->pos[4], NULL, n_own, d->counts_dev + 8);
    if ((rc = exchange(c, d->counts_dev + 8, 1, d->counts_dev + 9, 1, d->counts_dev + 10, 1, d->counts_dev + 11, 1, sizeof(int)))) return rc;
    if ((rc = read_counts(c, 12))) return rc;
    d->n_send[0] = d->counts_host[8]; d->n_send[1] = d->counts_host[9];
    d->n_recv[0] = d->counts_host[10]; d->n_recv[1] = d->counts_host[11];          // from hi, from lo
    const int n_halo = d->n_recv[0] + d->n_recv[1];
    if (d->p2p && ((size_t)d->n_send[0] > d->peer_cap[0] || (size_t)d->n_send[1] > d->peer_cap[1])) {
        sepgpu_set_error("decomposed rebuild: halo (%d / %d atoms) exceeds a neighbour's receive buffer", d->n_send[0], d->n_send[1]);
        return SEPGPU_EINVAL;
    }
    if ((size_t)d->n_send[0] > d->bufcap || (size_t)d->n_send[1] > d->bufcap || (size_t)d->n_recv[0] > d->bufcap ||
        (size_t)d->n_recv[1] > d->bufcap || n_own + n_halo > c->ncap) {
        sepgpu_set_error("decomposed rebuild: halo exceeds the buffers (own %d, halo %d, cap %d, buf %zu)", n_own, n_halo, c->ncap, d->bufcap);
        return SEPGPU_EINVAL;
    }
    if (G) {
        k_dd_pack_halo<<<G, B, 0, c->stream>>>(c->x4, c->gid, n_own, d->pos[3], d->send[0], d->send_idx[0], d->bufcap);
        k_dd_pack_halo<<<G, B, 0, c->stream>>>(c->x4, c->gid, n_own, d->pos[4], d->send[1], d->send_idx[1], d->bufcap);
    }
    if ((rc = exchange(c, d->send[0], (size_t)d->n_send[0] * 2, d->send[1], (size_t)d->n_send[1] * 2,
                       d->recv[0], (size_t)d->n_recv[0] * 2, d->recv[1], (size_t)d->n_recv[1] * 2, sizeof(d4)))) return rc;
    if (d->n_recv[0]) k_dd_unpack_halo<<<(d->n_recv[0] + B - 1) / B, B, 0, c->stream>>>(d->recv[0], d->n_recv[0], n_own, c->x4, c->v4, c->cr4, c->gid);
    if (d->n_recv[1]) k_dd_unpack_halo<<<(d->n_recv[1] + B - 1) / B, B, 0, c->stream>>>(d->recv[1], d->n_recv[1], n_own + d->n_recv[0], c->x4, c->v4, c->cr4, c->gid);
    ktimer_end(c, &c->t_migr);
    KERNEL_CHECK();
    c->n_own = n_own;
    c->n = n_own + n_halo;
    d->halo_current = true;
    d->halo_inflight = false;
    *zoff = d->z0 - 1;
    *nz_local = (d->z1 - d->z0) + 2;
    return 0;
}

// ---- every step: refresh the halo coordinates ---------------------------------------------------------------

// both directions in one launch
__global__ void k_dd_pack_xu2(const d4 *__restrict__ x4, const i4 *__restrict__ cr4, const int *__restrict__ idx0, int n0,
                              const int *__restrict__ idx1, int n1, double Lx, double Ly, double Lz,
                              d4 *__restrict__ out0, d4 *__restrict__ out1)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n0 + n1) return;
    const bool second = k >= n0;
    if (second) k -= n0;
    const int i = second ? idx1[k] : idx0[k];
    d4 x = x4[i];
    const int w = cr4[i].w;
    if (w != 0) {
        x.x += ((w & 1023) - 512) * Lx; x.y += (((w >> 10) & 1023) - 512) * Ly; x.z += (((w >> 20) & 1023) - 512) * Lz;
    }
    (second ? out1 : out0)[k] = x;
}

__global__ void k_dd_unpack_xu2(const d4 *__restrict__ in0, int n0, const d4 *__restrict__ in1, int n1, int first_local,
                                const int *__restrict__ rank, d4 *__restrict__ xs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n0 + n1) return;
    xs[rank[first_local + k]] = k < n0 ? in0[k] : in1[k - n0];
}

// ---- peer-memory variant: the pack kernel stores into the neighbours' buffers over NVLink ---------------------
__global__ void k_dd_push_xu2(const d4 *__restrict__ x4, const i4 *__restrict__ cr4, const int *__restrict__ idx0, int n0,
                              const int *__restrict__ idx1, int n1, double Lx, double Ly, double Lz,
                              d4 *out0, d4 *out1, unsigned long long *flag0, unsigned long long *flag1,
                              unsigned long long seq, unsigned int *done)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n0 + n1) {
        const bool second = k >= n0;
        if (second) k -= n0;
        const int i = second ? idx1[k] : idx0[k];
        d4 x = x4[i];
        const int w = cr4[i].w;
        if (w != 0) {
            x.x += ((w & 1023) - 512) * Lx; x.y += (((w >> 10) & 1023) - 512) * Ly; x.z += (((w >> 20) & 1023) - 512) * Lz;
        }
        (second ? out1 : out0)[k] = x;
    }
    // The block's stores are ordered before thread 0's system-scope fence by the barrier (cumulativity): one fence per
    // block publishes them; the last block to finish raises both flags
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(done, 1u);
        if (t == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag0), "l"(seq) : "memory");
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag1), "l"(seq) : "memory");
        }
    }
}

// Hypothesis for the open issue with launches sent ahead in decomposed runs (docs/ROUND_NOTES.md): the push kernel asks for
// the default shared-memory carve-out, the tile force kernel for the largest; an SM changes its carve-out only when it is
// empty, and a force grid whose ~200 waiting tiles sit on every SM then keeps the push -- which those tiles wait for, on the
// other GPU -- from ever being placed.  With the same preference the push can share an SM with them.  (Used with
// spec_force=2 only; not verified on hardware.)
int sepgpu_dd_push_carveout_max(void)
{
    return cudaFuncSetAttribute(k_dd_push_xu2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess ? 0 : SEPGPU_ECUDA;
}

// waits until both neighbours have delivered refresh number `seq`, then scatters into the halo slots of xs.
// The spin is bounded (SEPGPU_SPIN_LIMIT): a neighbour that never arrives sets the sticky device error instead of hanging.
__global__ void k_dd_wait_unpack_xu2(const d4 *in0, int n0, const d4 *in1, int n1, int first_local,
                                     const int *__restrict__ rank, d4 *__restrict__ xs,
                                     const unsigned long long *flags, unsigned long long seq, DevScalars *scal)
{
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags) < seq || ld_acquire_sys(flags + 1) < seq) {
            if (clock64() - t0 > SEPGPU_SPIN_LIMIT) { scal->error = SEPGPU_ENCCL; scal->error_where = 3; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n0 + n1) return;
    const double2 *src = reinterpret_cast<const double2 *>(k < n0 ? in0 + k : in1 + (k - n0));
    const double2 a = __ldcg(src), b = __ldcg(src + 1);         // written by another GPU: never through L1
    d4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    xs[rank[first_local + k]] = v;
}

// The refresh as two halves, so that a caller can put work that needs no halo between them:
//   begin : second stream waits for everything already queued on the main stream (the integrator that moved
//           the atoms), then packs, exchanges with both neighbours and scatters into the halo slots of xs;
//   end   : the main stream waits for that.
// Returns 1 from begin when a transfer was started, 0 when the halo was current already.
int sepgpu_dd_halo_begin(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    DDState *d = c->dd;
    if (d->halo_current || d->halo_inflight) return 0;
    const int B = 256;
    if (d->p2p) {
        // The push runs on the second stream, launched BEFORE the force kernel that follows on the main stream: both
        // are on the device at once, my remote stores and the flag travel while my own force tiles already compute.
        // (The next integrator waits for it: sepgpu_dd_before_positions_change.)
        d->seq++;
        const int nsend = d->n_send[0] + d->n_send[1];
        if (cudaEventRecord(d->ev_ready, c->stream) != cudaSuccess || cudaStreamWaitEvent(d->stream2, d->ev_ready, 0) != cudaSuccess) {
            sepgpu_set_error("halo refresh: stream hand-over failed");
            return SEPGPU_ECUDA;
        }
        k_dd_push_xu2<<<nsend ? (nsend + B - 1) / B : 1, B, 0, d->stream2>>>(c->x4, c->cr4, d->send_idx[0], d->n_send[0],
            d->send_idx[1], d->n_send[1], sys->length[0], sys->length[1], sys->length[2], d->peer_dst[0], d->peer_dst[1],
            d->peer_flag[0], d->peer_flag[1], d->seq, d->done_ctr);
        KERNEL_CHECK();
        CUDA_TRY(cudaEventRecord(d->ev_halo, d->stream2));
        d->push_pending = true;
        d->halo_inflight = true;
        return 1;
    }
    cudaStream_t st = d->stream2;
    if (cudaEventRecord(d->ev_ready, c->stream) != cudaSuccess || cudaStreamWaitEvent(st, d->ev_ready, 0) != cudaSuccess) {
        sepgpu_set_error("halo refresh: stream hand-over failed");
        return SEPGPU_ECUDA;
    }
    if (d->n_send[0] + d->n_send[1])
        k_dd_pack_xu2<<<(d->n_send[0] + d->n_send[1] + B - 1) / B, B, 0, st>>>(c->x4, c->cr4, d->send_idx[0], d->n_send[0],
            d->send_idx[1], d->n_send[1], sys->length[0], sys->length[1], sys->length[2], d->send[0], d->send[1]);
    int rc = exchange(c, d->send[0], (size_t)d->n_send[0], d->send[1], (size_t)d->n_send[1],
                      d->recv[0], (size_t)d->n_recv[0], d->recv[1], (size_t)d->n_recv[1], sizeof(d4), st);
    if (rc) return rc;
    if (d->n_recv[0] + d->n_recv[1])
        k_dd_unpack_xu2<<<(d->n_recv[0] + d->n_recv[1] + B - 1) / B, B, 0, st>>>(d->recv[0], d->n_recv[0], d->recv[1], d->n_recv[1],
                                                                               c->n_own, c->rank, c->xs);
    KERNEL_CHECK();
    CUDA_TRY(cudaEventRecord(d->ev_halo, st));
    d->halo_inflight = true;
    return 1;
}

int sepgpu_dd_halo_end(sepgpu_ctx *c)
{
    DDState *d = c->dd;
    if (!d->halo_inflight) return 0;
    if (d->p2p) {
        const int B = 256;
        const int nrecv = d->n_recv[0] + d->n_recv[1];
        k_dd_wait_unpack_xu2<<<nrecv ? (nrecv + B - 1) / B : 1, B, 0, c->stream>>>(d->p2p_recv[0], d->n_recv[0], d->p2p_recv[1],
            d->n_recv[1], c->n_own, c->rank, c->xs, d->p2p_flag, d->seq, c->scal);
        KERNEL_CHECK();
    } else {
        CUDA_TRY(cudaStreamWaitEvent(c->stream, d->ev_halo, 0));
    }
    d->halo_inflight = false;
    d->halo_current = true;
    return 0;
}

int sepgpu_dd_halo_update(sepgpu_ctx *c, const sepgpu_sys *sys)
{
    ktimer_begin(c, &c->t_halo);
    int rc = sepgpu_dd_halo_begin(c, sys);
    if (rc < 0) return rc;
    rc = sepgpu_dd_halo_end(c);
    ktimer_end(c, &c->t_halo);
    return rc;
}

// before anything on the main stream overwrites positions (integrators, migration): the push kernel must have read them
int sepgpu_dd_before_positions_change(sepgpu_ctx *c)
{
    DDState *d = c->dd;
    if (!d || !d->push_pending) return 0;
    CUDA_TRY(cudaStreamWaitEvent(c->stream, d->ev_halo, 0));
    d->push_pending = false;
    return 0;
}

void sepgpu_dd_positions_moved(sepgpu_ctx *c)
{
    if (!c->dd) return;
    c->dd->halo_current = false;
    if (c->dd->p2p) c->dd->halo_inflight = false;        // a pushed refresh nobody unpacked is simply obsolete now
}

// The tile force kernels consume the neighbours' boundary coordinates straight from the receive buffers: push mine
// (if this step's refresh has not gone out yet) and tell the kernel where to wait and read.  Returns 1 when `out` is
// valid, 0 when the caller has to refresh xs the ordinary way (not decomposed on the peer-memory path), < 0 on error.
int sepgpu_dd_halo_args(sepgpu_ctx *c, const sepgpu_sys *sys, HaloArgs *out)
{
    memset(out, 0, sizeof *out);
    DDState *d = c->dd;
    if (!d || !d->p2p) return 0;
    if (!d->halo_current) {
        ktimer_begin(c, &c->t_halo);
        const int rc = sepgpu_dd_halo_begin(c, sys);      // no-op when this step's push is already under way
        ktimer_end(c, &c->t_halo);
        if (rc < 0) return rc;
        out->seq = d->seq;                                // halo_current: the list build has just placed the halo in xs itself
    }
    out->in0 = d->p2p_recv[0]; out->in1 = d->p2p_recv[1];
    out->n0 = d->n_recv[0]; out->n_own = c->n_own;
    out->flags = d->p2p_flag;
    return 1;
}

// ---- scalars: force-derived sums are kept per rank and summed when read -----------------------------------
__global__ void k_dd_gather_force_scalars(const DevScalars *s, double *comm)
{
    const int t = threadIdx.x;
    if (t == 0) { comm[0] = s->epot; comm[1] = s->ecoul; }
    if (t < 9) { comm[2 + t] = s->pot_P[t]; comm[11 + t] = s->pot_P_bond[t]; }
}

int sepgpu_dd_reduce_force_scalars(sepgpu_ctx *c, double *epot, double *ecoul, double *pot_P, double *pot_P_bond)
{
    DDState *d = c->dd;
    k_dd_gather_force_scalars<<<1, 32, 0, c->stream>>>(c->scal, d->comm_buf + 96);
    int rc = sepgpu_dd_allreduce(c, d->comm_buf + 96, 20, NULL, 0);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d->comm_host, d->comm_buf + 96, sizeof(double) * 20, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *epot = d->comm_host[0]; *ecoul = d->comm_host[1];
    memcpy(pot_P, d->comm_host + 2, sizeof(double) * 9);
    memcpy(pot_P_bond, d->comm_host + 11, sizeof(double) * 9);
    return 0;
}
