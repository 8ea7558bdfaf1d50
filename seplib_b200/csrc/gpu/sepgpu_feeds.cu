// sepgpu_feeds.cu -- sampler feeds: the reduced inputs of the run-time samplers, computed where the data lives.
//
// The reference's samplers (source/sepsampler.c:177-240 and the routines it dispatches to) read atoms[] on the host
// at every sample -- the mean-square-displacement sampler at EVERY step, to follow the atoms across the periodic
// boundaries (:537-552).  Behind the sep_* API that is a full download of the 568-byte records per sample.  The
// entry points below return what each sampler actually consumes instead:
//
//   sepgpu_feed_vacf      one block of the velocity autocorrelation (:658-722): lvec rows of v_x stay on the device,
//                         the block's sum over atoms and time origins comes back as lvec numbers
//   sepgpu_feed_msd       sum of dr^2, dr^4, the atom count and the self-intermediate scattering sums (:555-655) from the
//                         device's own crossing counters -- no per-step tracking at all
//   sepgpu_feed_profile   momentum / mass / thermal sums and counts per slab along z (:1416-1525)
//   sepgpu_feed_fourier   the Fourier sums behind the generalised-hydrodynamics correlators (:926-1107), 7 complex
//                         numbers per wave vector
//   sepgpu_feed_radial    the pair-distance histogram per type combination (:361-468), integer counts (exact)
//
// All reductions but the profile bins are fixed trees (same numbers every run); the profile uses FP64 atomics into
// shared-memory bins.  Single-domain contexts only (decomposed runs answer SEPGPU_ESTATE and the host layer falls back
// to its synchronised path).
#include "sepgpu_internal.cuh"

#define FEED_BLOCK 256
#define FEED_NV 16
#define FEED_MAX_CHUNKS 296          // two CTAs per SM of a B200

struct FeedState {
    // velocity autocorrelation
    double *vrows; int v_lvec, v_fill; size_t v_n;
    // mean square displacement origin
    d4 *pos0; int *cr0; bool msd_set;
    // scratch for the reductions (device) and their results (device, then staged to the caller)
    double *part; size_t part_cap;
    double *out; size_t out_cap;
    unsigned long long *hist; size_t hist_cap;
};

int sepgpu_apply_pending(sepgpu_ctx *c);

static FeedState *feeds_of(sepgpu_ctx *c)
{
    if (!c->feeds) c->feeds = (FeedState *)calloc(1, sizeof(FeedState));
    return c->feeds;
}

void sepgpu_feeds_destroy(sepgpu_ctx *c)
{
    FeedState *F = c->feeds;
    if (!F) return;
    if (F->vrows) cudaFree(F->vrows);
    if (F->pos0) cudaFree(F->pos0);
    if (F->cr0) cudaFree(F->cr0);
    if (F->part) cudaFree(F->part);
    if (F->out) cudaFree(F->out);
    if (F->hist) cudaFree(F->hist);
    free(F);
    c->feeds = NULL;
}

static int feed_scratch(FeedState *F, size_t part, size_t out)
{
    if (part > F->part_cap) {
        if (F->part) cudaFree(F->part);
        F->part = NULL; F->part_cap = 0;
        CUDA_TRY(cudaMalloc((void **)&F->part, sizeof(double) * part));
        F->part_cap = part;
    }
    if (out > F->out_cap) {
        if (F->out) cudaFree(F->out);
        F->out = NULL; F->out_cap = 0;
        CUDA_TRY(cudaMalloc((void **)&F->out, sizeof(double) * out));
        F->out_cap = out;
    }
    return 0;
}

static int feed_enter(sepgpu_ctx *c, const char *who)
{
    if (c->dd) { sepgpu_set_error("%s: sampler feeds serve single-domain contexts", who); return SEPGPU_ESTATE; }
    if (c->n_own <= 0) { sepgpu_set_error("%s: no atoms", who); return SEPGPU_EINVAL; }
    c->feed_calls++;
    return 0;
}

static int feed_chunks(int n)
{
    const int want = (n + FEED_BLOCK - 1) / FEED_BLOCK;
    return want < FEED_MAX_CHUNKS ? want : FEED_MAX_CHUNKS;
}

// out[r * stride + q] = sum over chunks of part[(r * nchunks + chunk) * stride + q], chunks in index order
__global__ void k_feed_sum_chunks(const double *__restrict__ part, int nchunks, int stride, double *__restrict__ out)
{
    const int r = blockIdx.x, q = threadIdx.x;
    if (q >= stride) return;
    double s = 0.0;
    for (int ch = 0; ch < nchunks; ch++) s += part[((size_t)r * nchunks + ch) * stride + q];
    out[(size_t)r * stride + q] = s;
}

static int feed_download(sepgpu_ctx *c, const void *dev, void *host, size_t bytes)
{
    int rc = sepgpu_ensure_stage(c, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));                // an earlier asynchronous copy may still read the staging buffer
    CUDA_TRY(cudaMemcpyAsync(c->stage, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(host, c->stage, bytes);
    return 0;
}

// ---- velocity autocorrelation ----------------------------------------------------------------------------------------
__global__ void k_feed_vx_row(const d4 *__restrict__ v4, double *__restrict__ row, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) row[i] = v4[i].x;                                // source/sepsampler.c:694 samples v[0] only
}

// One CTA takes groups of 32 atoms: their lvec samples are staged in shared memory once ([lvec][32], conflict-free), warp w
// owns the lags t = w, w + 8, ... (lane = atom), sums over the time origins, then over the 32 atoms by shuffles, and adds to
// the CTA's own acf[t] (one owner per lag, no race).  The rows are read from HBM exactly once per block.
#define ACF_GROUP 32
__global__ void __launch_bounds__(FEED_BLOCK)
k_feed_acf(const double *__restrict__ rows, int lvec, size_t n, double *__restrict__ part)
{
    extern __shared__ double acf_smem[];
    double *col = acf_smem;                       // [lvec][32]
    double *cacc = acf_smem + (size_t)lvec * ACF_GROUP;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = FEED_BLOCK / 32;
    for (int t = threadIdx.x; t < lvec; t += FEED_BLOCK) cacc[t] = 0.0;
    const size_t ngroups = (n + ACF_GROUP - 1) / ACF_GROUP;
    for (size_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        __syncthreads();
        const size_t i = g * ACF_GROUP + lane;
        for (int t0 = wid; t0 < lvec; t0 += nw) col[t0 * ACF_GROUP + lane] = i < n ? rows[(size_t)t0 * n + i] : 0.0;
        __syncthreads();
        for (int t = wid; t < lvec; t += nw) {
            double s = 0.0;
            for (int t0 = 0; t0 + t < lvec; t0++) s += col[t0 * ACF_GROUP + lane] * col[(t0 + t) * ACF_GROUP + lane];
            s = warp_sum(s);
            if (lane == 0) cacc[t] += s;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < lvec; t += FEED_BLOCK) part[(size_t)t * gridDim.x + blockIdx.x] = cacc[t];
}

extern "C" int sepgpu_feed_vacf(sepgpu_ctx *c, int lvec, double *acf_block, int *completed)
{
    if (!c || lvec <= 0 || !acf_block || !completed) return SEPGPU_EINVAL;
    if (lvec > 768) { sepgpu_set_error("feed_vacf: at most 768 samples per block (shared-memory staging)"); return SEPGPU_EINVAL; }
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    int rc = feed_enter(c, "feed_vacf");
    if (rc) return rc;
    FeedState *F = feeds_of(c);
    const size_t n = (size_t)c->n_own;
    if (!F->vrows || F->v_lvec != lvec || F->v_n != n) {
        if (F->vrows) cudaFree(F->vrows);
        F->vrows = NULL;
        CUDA_TRY(cudaMalloc((void **)&F->vrows, sizeof(double) * n * (size_t)lvec));
        F->v_lvec = lvec; F->v_n = n; F->v_fill = 0;
    }
    k_feed_vx_row<<<(c->n_own + FEED_BLOCK - 1) / FEED_BLOCK, FEED_BLOCK, 0, c->stream>>>(c->v4, F->vrows + (size_t)F->v_fill * n, c->n_own);
    KERNEL_CHECK();
    *completed = 0;
    if (++F->v_fill < lvec) return 0;
    F->v_fill = 0;
    const size_t ngroups = (n + ACF_GROUP - 1) / ACF_GROUP;
    const int nch = ngroups < FEED_MAX_CHUNKS ? (int)ngroups : FEED_MAX_CHUNKS;
    if ((rc = feed_scratch(F, (size_t)lvec * nch, (size_t)lvec))) return rc;
    const size_t smem = sizeof(double) * (size_t)lvec * (ACF_GROUP + 1);
    CUDA_TRY(cudaFuncSetAttribute(k_feed_acf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_feed_acf<<<nch, FEED_BLOCK, smem, c->stream>>>(F->vrows, lvec, n, F->part);
    KERNEL_CHECK();
    k_feed_sum_chunks<<<lvec, 32, 0, c->stream>>>(F->part, nch, 1, F->out);
    KERNEL_CHECK();
    if ((rc = feed_download(c, F->out, acf_block, sizeof(double) * (size_t)lvec))) return rc;
    *completed = 1;
    return 0;
}

// ---- mean square displacement -------------------------------------------------------------------------------------------
__global__ void k_feed_msd_origin(const d4 *__restrict__ x4, const int *__restrict__ crossings, d4 *__restrict__ pos0,
                                  int *__restrict__ cr0, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos0[i] = x4[i];
    cr0[3 * i] = crossings[3 * i]; cr0[3 * i + 1] = crossings[3 * i + 1]; cr0[3 * i + 2] = crossings[3 * i + 2];
}

struct MsdK { double k[FEED_NV - 3]; int nk; };

// per chunk: [0] sum dr^2, [1] sum dr^4, [2] atoms of the type, [3 + i] sum cos(k_i dx)      (source/sepsampler.c:583-600)
__global__ void __launch_bounds__(FEED_BLOCK)
k_feed_msd(const d4 *__restrict__ x4, const int *__restrict__ crossings, const d4 *__restrict__ pos0,
           const int *__restrict__ cr0, int n, int type, double Lx, double Ly, double Lz, MsdK K, double *__restrict__ part)
{
    __shared__ double red[FEED_NV * (FEED_BLOCK / 32)];
    double acc[FEED_NV];
#pragma unroll
    for (int q = 0; q < FEED_NV; q++) acc[q] = 0.0;
    for (int i = blockIdx.x * FEED_BLOCK + threadIdx.x; i < n; i += gridDim.x * FEED_BLOCK) {
        const d4 x = x4[i];
        if (tag_type(x.w) != type) continue;
        const d4 p = pos0[i];
        const double dx = x.x + (crossings[3 * i] - cr0[3 * i]) * Lx - p.x;
        const double dy = x.y + (crossings[3 * i + 1] - cr0[3 * i + 1]) * Ly - p.y;
        const double dz = x.z + (crossings[3 * i + 2] - cr0[3 * i + 2]) * Lz - p.z;
        const double a = dx * dx + dy * dy + dz * dz;
        acc[0] += a; acc[1] += a * a; acc[2] += 1.0;
#pragma unroll
        for (int q = 0; q < FEED_NV - 3; q++)
            if (q < K.nk) acc[3 + q] += cos(K.k[q] * dx);
    }
    block_sum<FEED_NV, FEED_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < FEED_NV; q++) part[(size_t)blockIdx.x * FEED_NV + q] = acc[q];
}

// new_origin != 0: the positions and crossing counters of this call become the origin (displacements of this call are 0)
extern "C" int sepgpu_feed_msd(sepgpu_ctx *c, int new_origin, char type, const double length[3], int nk, const double *k,
                               double *sums, double *fs)
{
    if (!c || !length || !sums || nk < 0 || (nk && (!k || !fs))) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    int rc = feed_enter(c, "feed_msd");
    if (rc) return rc;
    FeedState *F = feeds_of(c);
    const int n = c->n_own;
    if (!F->pos0) {
        CUDA_TRY(cudaMalloc((void **)&F->pos0, sizeof(d4) * (size_t)c->ncap));
        CUDA_TRY(cudaMalloc((void **)&F->cr0, sizeof(int) * 3 * (size_t)c->ncap));
        F->msd_set = false;
    }
    if (new_origin || !F->msd_set) {
        k_feed_msd_origin<<<(n + FEED_BLOCK - 1) / FEED_BLOCK, FEED_BLOCK, 0, c->stream>>>(c->x4, c->crossings, F->pos0, F->cr0, n);
        KERNEL_CHECK();
        F->msd_set = true;
    }
    const int nch = feed_chunks(n);
    if ((rc = feed_scratch(F, (size_t)nch * FEED_NV, FEED_NV))) return rc;
    const int per = FEED_NV - 3;
    int done = 0;
    do {                                                        // wave numbers in groups of 13; the first three sums with every group
        MsdK K;
        K.nk = nk - done < per ? nk - done : per;
        for (int q = 0; q < per; q++) K.k[q] = q < K.nk ? k[done + q] : 0.0;
        k_feed_msd<<<nch, FEED_BLOCK, 0, c->stream>>>(c->x4, c->crossings, F->pos0, F->cr0, n, (int)(unsigned char)type,
                                                       length[0], length[1], length[2], K, F->part);
        KERNEL_CHECK();
        k_feed_sum_chunks<<<1, 32, 0, c->stream>>>(F->part, nch, FEED_NV, F->out);
        KERNEL_CHECK();
        double got[FEED_NV];
        if ((rc = feed_download(c, F->out, got, sizeof got))) return rc;
        sums[0] = got[0]; sums[1] = got[1]; sums[2] = got[2];
        for (int q = 0; q < K.nk; q++) fs[done + q] = got[3 + q];
        done += K.nk;
    } while (done < nk);
    return 0;
}

// ---- profiles along z ---------------------------------------------------------------------------------------------------
// out[0..nb) momentum m v_x, [nb..2nb) mass, [2nb..3nb) m (v_y^2 + v_z^2), [3nb..4nb) atom counts (source/sepsampler.c:1447-1463)
__global__ void __launch_bounds__(FEED_BLOCK)
k_feed_profile(const d4 *__restrict__ x4, const d4 *__restrict__ v4, int n, int type, double dl, int nb, double *__restrict__ out)
{
    extern __shared__ double bins[];
    for (int q = threadIdx.x; q < 4 * nb; q += FEED_BLOCK) bins[q] = 0.0;
    __syncthreads();
    for (int i = blockIdx.x * FEED_BLOCK + threadIdx.x; i < n; i += gridDim.x * FEED_BLOCK) {
        const d4 x = x4[i];
        if (tag_type(x.w) != type) continue;
        const d4 v = v4[i];
        int b = (int)(x.z / dl);
        if (b < 0) b = 0;
        if (b >= nb) b = nb - 1;
        atomicAdd(&bins[b], v.w * v.x);
        atomicAdd(&bins[nb + b], v.w);
        atomicAdd(&bins[2 * nb + b], v.w * v.y * v.y + v.w * v.z * v.z);
        atomicAdd(&bins[3 * nb + b], 1.0);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 4 * nb; q += FEED_BLOCK)
        if (bins[q] != 0.0) atomicAdd(&out[q], bins[q]);
}

extern "C" int sepgpu_feed_profile(sepgpu_ctx *c, char type, double lz, int nbins, double *out4)
{
    if (!c || !out4 || nbins <= 0 || !(lz > 0.0)) return SEPGPU_EINVAL;
    if (nbins > 1024) { sepgpu_set_error("feed_profile: at most 1024 slabs"); return SEPGPU_EINVAL; }
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    int rc = feed_enter(c, "feed_profile");
    if (rc) return rc;
    FeedState *F = feeds_of(c);
    if ((rc = feed_scratch(F, 0, (size_t)4 * nbins))) return rc;
    CUDA_TRY(cudaMemsetAsync(F->out, 0, sizeof(double) * 4 * (size_t)nbins, c->stream));
    const int nch = feed_chunks(c->n_own) < 148 ? feed_chunks(c->n_own) : 148;
    k_feed_profile<<<nch, FEED_BLOCK, sizeof(double) * 4 * (size_t)nbins, c->stream>>>(c->x4, c->v4, c->n_own, (int)(unsigned char)type,
                                                                                      lz / nbins, nbins, F->out);
    KERNEL_CHECK();
    return feed_download(c, F->out, out4, sizeof(double) * 4 * (size_t)nbins);
}

// ---- Fourier sums of the generalised-hydrodynamics sampler ----------------------------------------------------------------
// per wave vector (0, k, 0), with e = exp(i k y_true):  [0,1] sum e   [2,3] sum m e   [4,5] sum m v_x e   [6,7] sum m v_y e
// [8,9] sum (m v^2 / 2) e   [10,11] sum m a_y e   [12,13] sum m v_y^2 e   [14] sum m v^2 / 2   (source/sepsampler.c:926-1000)
__global__ void __launch_bounds__(FEED_BLOCK)
k_feed_fourier(const d4 *__restrict__ x4, const int *__restrict__ crossings, const d4 *__restrict__ v4,
               const d4 *__restrict__ f4, int n, double Ly, const double *__restrict__ kv, double *__restrict__ part)
{
    __shared__ double red[FEED_NV * (FEED_BLOCK / 32)];
    double acc[FEED_NV];
#pragma unroll
    for (int q = 0; q < FEED_NV; q++) acc[q] = 0.0;
    const double k = kv[blockIdx.y];
    for (int i = blockIdx.x * FEED_BLOCK + threadIdx.x; i < n; i += gridDim.x * FEED_BLOCK) {
        const d4 x = x4[i], v = v4[i];
        const double ytrue = x.y + crossings[3 * i + 1] * Ly;   // sep_eval_xtrue
        double sn, cs;
        sincos(k * ytrue, &sn, &cs);
        const double m = v.w;
        const double ekin = 0.5 * m * v.x * v.x + 0.5 * m * v.y * v.y + 0.5 * m * v.z * v.z;
        const double may = f4 ? f4[i].y : 0.0;                  // m a_y with a = f / m (source/sepintgr.c:53)
        const double w[7] = {1.0, m, m * v.x, m * v.y, ekin, may, m * v.y * v.y};
#pragma unroll
        for (int q = 0; q < 7; q++) { acc[2 * q] += w[q] * cs; acc[2 * q + 1] += w[q] * sn; }
        acc[14] += ekin;
    }
    block_sum<FEED_NV, FEED_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < FEED_NV; q++) part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * FEED_NV + q] = acc[q];
}

extern "C" int sepgpu_feed_fourier(sepgpu_ctx *c, double ly, int nwave, const double *k, double *out16)
{
    if (!c || !k || !out16 || nwave <= 0 || nwave > 4096) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    int rc = feed_enter(c, "feed_fourier");
    if (rc) return rc;
    if ((rc = sepgpu_apply_pending(c))) return rc;             // the thermostat's share of f, as SEPGPU_F_A has it
    FeedState *F = feeds_of(c);
    const int nch = feed_chunks(c->n_own) < 64 ? feed_chunks(c->n_own) : 64;
    if ((rc = feed_scratch(F, (size_t)nwave * nch * FEED_NV + (size_t)nwave, (size_t)nwave * FEED_NV))) return rc;
    double *kdev = F->part + (size_t)nwave * nch * FEED_NV;
    if ((rc = sepgpu_ensure_stage(c, sizeof(double) * (size_t)nwave))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(c->stage, k, sizeof(double) * (size_t)nwave);
    CUDA_TRY(cudaMemcpyAsync(kdev, c->stage, sizeof(double) * (size_t)nwave, cudaMemcpyHostToDevice, c->stream));
    k_feed_fourier<<<dim3(nch, nwave), FEED_BLOCK, 0, c->stream>>>(c->x4, c->crossings, c->v4, c->f_zero ? NULL : c->f4, c->n_own, ly, kdev, F->part);
    KERNEL_CHECK();
    k_feed_sum_chunks<<<nwave, 32, 0, c->stream>>>(F->part, nch, FEED_NV, F->out);
    KERNEL_CHECK();
    return feed_download(c, F->out, out16, sizeof(double) * (size_t)nwave * FEED_NV);
}

// ---- radial distribution: pair-distance histogram per type combination ------------------------------------------------------
#define RDF_TILE 128
struct RdfTypes { int n; unsigned char t[16]; };

// grid (tiles, tiles), upper triangle only; every CTA counts into shared-memory bins, then adds them to the 64-bit table
__global__ void __launch_bounds__(RDF_TILE)
k_feed_radial(const d4 *__restrict__ x4, int n, double lbox, double dg, int lvec, RdfTypes T, int ncomb,
              unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned rbins[];
    __shared__ d4 xj[RDF_TILE];
    if (blockIdx.y < blockIdx.x) return;
    for (int q = threadIdx.x; q < lvec * ncomb; q += RDF_TILE) rbins[q] = 0u;
    const int i = blockIdx.x * RDF_TILE + threadIdx.x;
    const int j0 = blockIdx.y * RDF_TILE;
    if (j0 + (int)threadIdx.x < n) xj[threadIdx.x] = x4[j0 + threadIdx.x];
    __syncthreads();
    if (i < n) {
        const d4 xi = x4[i];
        const int ti = tag_type(xi.w);
        const double half = 0.5 * lbox;
        const int jn = n - j0 < RDF_TILE ? n - j0 : RDF_TILE;
        for (int jj = 0; jj < jn; jj++) {
            if (j0 + jj <= i) continue;                         // pairs i < j once (source/sepsampler.c:390-391)
            const d4 p = xj[jj];
            const double dx = wrap_exact(xi.x - p.x, lbox, half);      // the x length wraps all three directions there (:396)
            const double dy = wrap_exact(xi.y - p.y, lbox, half);
            const double dz = wrap_exact(xi.z - p.z, lbox, half);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const int index = (int)(sqrt(r2) / dg);
            if (index >= lvec) continue;
            const int tj = tag_type(p.w);
            int counter = 0;
            for (int a = 0; a < T.n; a++)
                for (int b = a; b < T.n; b++) {
                    if ((ti == T.t[a] && tj == T.t[b]) || (ti == T.t[b] && tj == T.t[a])) atomicAdd(&rbins[index * ncomb + counter], 1u);
                    counter++;
                }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < lvec * ncomb; q += RDF_TILE)
        if (rbins[q]) atomicAdd(&hist[q], (unsigned long long)rbins[q]);
}

// counts[lvec][ncomb] of THIS call (the caller accumulates), ncomb = ntypes (ntypes + 1) / 2 in the reference's order
extern "C" int sepgpu_feed_radial(sepgpu_ctx *c, double lbox, int lvec, int ntypes, const char *types, long long *counts)
{
    if (!c || !types || !counts || lvec <= 0 || ntypes <= 0 || ntypes > 16 || !(lbox > 0.0)) return SEPGPU_EINVAL;
    const int ncomb = ntypes * (ntypes + 1) / 2;
    if ((size_t)lvec * ncomb > 8192) { sepgpu_set_error("feed_radial: at most 8192 bins x type combinations"); return SEPGPU_EINVAL; }
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    int rc = feed_enter(c, "feed_radial");
    if (rc) return rc;
    FeedState *F = feeds_of(c);
    const size_t nb = (size_t)lvec * ncomb;
    if (nb > F->hist_cap) {
        if (F->hist) cudaFree(F->hist);
        F->hist = NULL; F->hist_cap = 0;
        CUDA_TRY(cudaMalloc((void **)&F->hist, sizeof(unsigned long long) * nb));
        F->hist_cap = nb;
    }
    CUDA_TRY(cudaMemsetAsync(F->hist, 0, sizeof(unsigned long long) * nb, c->stream));
    RdfTypes T; T.n = ntypes;
    for (int a = 0; a < 16; a++) T.t[a] = a < ntypes ? (unsigned char)types[a] : 0;
    const int tiles = (c->n_own + RDF_TILE - 1) / RDF_TILE;
    if (tiles > 65535) { sepgpu_set_error("feed_radial: system too large for the all-pairs histogram"); return SEPGPU_EINVAL; }
    k_feed_radial<<<dim3(tiles, tiles), RDF_TILE, sizeof(unsigned) * nb, c->stream>>>(c->x4, c->n_own, lbox, 0.5 * lbox / lvec, lvec, T, ncomb, F->hist);
    KERNEL_CHECK();
    return feed_download(c, F->hist, counts, sizeof(long long) * nb);
}
