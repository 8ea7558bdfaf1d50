// sepgpu_state.cu -- context lifetime, host<->HBM marshalling, scalar block, measurement helpers.
#include "sepgpu_internal.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <functional>

void sepgpu_dd_destroy(sepgpu_ctx *c);
void sepgpu_feeds_destroy(sepgpu_ctx *c);
int sepgpu_dd_reduce_force_scalars(sepgpu_ctx *c, double *epot, double *ecoul, double *pot_P, double *pot_P_bond);

static thread_local char g_err[512] = "";

void sepgpu_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *sepgpu_last_error(void) { return g_err; }

extern "C" int sepgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <typename T>
static int dalloc(T **p, size_t count)
{
    CUDA_TRY(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
    CUDA_TRY(cudaMemset(*p, 0, sizeof(T) * (count ? count : 1)));
    return 0;
}

__global__ void k_init_tags(d4 *x4, d4 *v4, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // defaults of sep_init (source/sepinit.c:34-45): type 'A', m = 1, molindex = -1, z = 0
    x4[i].w = make_tag('A', -1);
    v4[i].w = 1.0;
}


extern "C" int sepgpu_create(sepgpu_ctx **out, size_t npart, int device)
{
    if (!out || npart == 0 || npart >= SEPGPU_MAX_ATOMS) {
        sepgpu_set_error("sepgpu_create: npart=%zu out of range (1..%u)", npart, SEPGPU_MAX_ATOMS - 1);
        return SEPGPU_EINVAL;
    }
    int ndev = sepgpu_device_count();
    if (ndev <= 0) {
        sepgpu_set_error("sepgpu_create: no CUDA device available (seplib-b200 has no CPU path)");
        return SEPGPU_ENODEV;
    }
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % ndev : 0;
    }
    if (device >= ndev) { sepgpu_set_error("sepgpu_create: device %d of %d", device, ndev); return SEPGPU_EINVAL; }
    CUDA_TRY(cudaSetDevice(device));

    sepgpu_ctx *c = (sepgpu_ctx *)calloc(1, sizeof(sepgpu_ctx));
    if (!c) return SEPGPU_EINVAL;
    c->n = (int)npart;
    c->n_own = c->n;
    c->ncap = c->n;
    c->n_global = (long long)npart;
    c->npad = (c->n + 31) & ~31;
    c->device = device;
    c->pending_alpha_slot = -1;
    c->pending_alpha_type = -1;
    c->tpa = 1;
    c->prefilter = 1;
    c->overlap = 0;
    c->single_type = 'A';
    // defaults settled on a B200 in round 2 (scripts/gpu_r2_ab.sh, profiles/r02_optin_ab.txt)
    c->coulomb_kernel = 2;
    c->typed_sublist = 1;
    c->step_fold = 1;
    c->fin_multi = 1;
    c->spec.on = 1;
    c->tile_list = 1;
    c->build_window = 1;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));

    const size_t n = npart;
    if (dalloc(&c->x4, n) || dalloc(&c->v4, n) || dalloc(&c->f4, n) || dalloc(&c->xn4, n) ||
        dalloc(&c->cr4, n) || dalloc(&c->crossings, 3 * n) || dalloc(&c->z, n) ||
        dalloc(&c->xs, n) || dalloc(&c->xf, n) || dalloc(&c->order, n) || dalloc(&c->rank, n) ||
        dalloc(&c->cell_of, n) || dalloc(&c->tmp_slot, n) || dalloc(&c->cnt, (size_t)c->npad) ||
        dalloc(&c->scal, 1) || dalloc(&c->partial, (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * 16))
        return SEPGPU_ECUDA;
    CUDA_TRY(cudaMallocHost((void **)&c->scal_host, sizeof(DevScalars)));
    memset(c->scal_host, 0, sizeof(DevScalars));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));

    k_init_tags<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->x4, c->v4, c->n);
    KERNEL_CHECK();
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->f_zero = true;
    // SEPGPU_OPTS="name=value,name=value": options for programs that only see the sep_* API (prgs/*.c)
    const char *opts = getenv("SEPGPU_OPTS");
    if (opts && *opts) {
        char buf[512];
        strncpy(buf, opts, sizeof buf - 1); buf[sizeof buf - 1] = 0;
        for (char *tok = strtok(buf, ","); tok; tok = strtok(NULL, ",")) {
            char *eq = strchr(tok, '=');
            if (!eq) { sepgpu_set_error("SEPGPU_OPTS: '%s' is not name=value", tok); sepgpu_destroy(c); return SEPGPU_EINVAL; }
            *eq = 0;
            const int rc = sepgpu_set_option(c, tok, atoll(eq + 1));
            if (rc) { sepgpu_destroy(c); return rc; }
        }
    }
    *out = c;
    return 0;
}

static void ktimer_free(KernelTimer *t)
{
    if (!t->enabled) return;
    for (int i = 0; i < 64; i++) { cudaEventDestroy(t->start[i]); cudaEventDestroy(t->stop[i]); }
}

extern "C" void sepgpu_destroy(sepgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    sepgpu_dd_destroy(c);
    sepgpu_feeds_destroy(c);
    if (c->gid) cudaFree(c->gid);
    if (c->f_side) cudaFree(c->f_side);
    if (c->put_changed) cudaFree(c->put_changed);
    if (c->partial_side) cudaFree(c->partial_side);
    if (c->fij) cudaFree(c->fij);
    if (c->cls) cudaFree(c->cls);
    if (c->x0) cudaFree(c->x0);
    if (c->prevf4) cudaFree(c->prevf4);
    for (int k = 0; k < SEPGPU_NSUB; k++) { if (c->nbr_t[k]) cudaFree(c->nbr_t[k]); if (c->cnt_t[k]) cudaFree(c->cnt_t[k]); }
    if (c->tsort) cudaFree(c->tsort);
    if (c->xq) cudaFree(c->xq);
    if (c->fin_ticket) cudaFree(c->fin_ticket);
    for (int q = 0; q < SEPGPU_NTAB; q++) if (c->tab[q].dev) cudaFree(c->tab[q].dev);
    if (c->f4_alt) cudaFree(c->f4_alt);
    if (c->flag_stream) cudaStreamDestroy(c->flag_stream);
    if (c->ev_fin) cudaEventDestroy(c->ev_fin);
    if (c->tile_hdr) cudaFree(c->tile_hdr);
    if (c->tile_src) cudaFree(c->tile_src);
    if (c->randn4) cudaFree(c->randn4);
    void *ptrs[] = {c->x4, c->v4, c->f4, c->xn4, c->pv4, c->pa4, c->cr4, c->crossings, c->z, c->type,
                    c->molindex, c->excl_bond, c->excl_angle, c->excl_dihed, c->zs, c->xs, c->xf, c->order,
                    c->rank, c->cell_of, c->cell_cnt, c->cell_start, c->tmp_slot, c->nbr, c->cnt,
                    c->blist, c->alist, c->dlist, c->atom_bond_ptr, c->atom_bond_idx,
                    c->atom_angle_ptr, c->atom_angle_idx, c->atom_dihed_ptr, c->atom_dihed_idx,
                    c->blengths, c->angles, c->dihedrals, c->scal, c->partial, c->dstage, c->flush_buf};
    for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++)
        if (ptrs[i]) cudaFree(ptrs[i]);
    if (c->scal_host) cudaFreeHost(c->scal_host);
    if (c->stage) cudaFreeHost(c->stage);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    ktimer_free(&c->t_force); ktimer_free(&c->t_build); ktimer_free(&c->t_intgr); ktimer_free(&c->t_coul); ktimer_free(&c->t_bonded); ktimer_free(&c->t_halo); ktimer_free(&c->t_migr);
    cudaStreamDestroy(c->stream);
    free(c);
}

int sepgpu_ensure_stage(sepgpu_ctx *c, size_t bytes)
{
    if (c->stage_bytes < bytes) {
        if (c->stage) cudaFreeHost(c->stage);
        c->stage = NULL; c->stage_bytes = 0;
        CUDA_TRY(cudaMallocHost(&c->stage, bytes));
        c->stage_bytes = bytes;
    }
    if (c->dstage_bytes < bytes) {
        if (c->dstage) cudaFree(c->dstage);
        c->dstage = NULL; c->dstage_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->dstage, bytes));
        c->dstage_bytes = bytes;
    }
    return 0;
}

// ---- packing kernels ----------------------------------------------------------------------------------
__global__ void k_vec3_to_d4(d4 *dst, const double *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 v = dst[i];
    v.x = src[3 * i]; v.y = src[3 * i + 1]; v.z = src[3 * i + 2];
    dst[i] = v;
}
// Uploads that invalidate the neighbour list only when they bring different values: a program that edits one member of
// atoms[] between hot calls (reference prgs/prg5.c:72-76 adds to atoms[i].f) makes the host layer upload every member --
// it cannot know which one changed -- and positions, types, molecule indices and exclusion tables that come back
// unchanged must not cost a list rebuild per step.
__global__ void k_vec3_to_d4_changed(d4 *dst, const double *src, int n, int *changed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 v = dst[i];
    const double a = src[3 * i], b = src[3 * i + 1], c = src[3 * i + 2];
    if (v.x != a || v.y != b || v.z != c) {
        v.x = a; v.y = b; v.z = c;
        dst[i] = v;
        *changed = 1;
    }
}
__global__ void k_set_tag_changed(d4 *x4, const char *types, const int *mols, int n, int *changed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double w = x4[i].w;
    const int t = types ? (int)(unsigned char)types[i] : tag_type(w), m = mols ? mols[i] : tag_mol(w);
    if (t != tag_type(w) || m != tag_mol(w)) {
        x4[i].w = make_tag((unsigned char)t, m);
        *changed = 1;
    }
}
__global__ void k_copy_int_changed(int *dst, const int *src, size_t count, int *changed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int v = src[i];
    if (dst[i] != v) { dst[i] = v; *changed = 1; }
}
__global__ void k_d4_to_vec3(double *dst, const d4 *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
__global__ void k_scalar_to_w(d4 *dst, const double *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i].w = src[i];
}
__global__ void k_w_to_scalar(double *dst, const d4 *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i].w;
}
__global__ void k_set_type(d4 *x4, const char *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x4[i].w = make_tag((unsigned char)src[i], tag_mol(x4[i].w));
}
__global__ void k_get_type(char *dst, const d4 *x4, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (char)tag_type(x4[i].w);
}
__global__ void k_set_mol(d4 *x4, const int *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x4[i].w = make_tag(tag_type(x4[i].w), src[i]);
}
__global__ void k_get_mol(int *dst, const d4 *x4, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = tag_mol(x4[i].w);
}
__global__ void k_int3_to_cr(i4 *dst, const int *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i4 v = dst[i];
    v.x = src[3 * i]; v.y = src[3 * i + 1]; v.z = src[3 * i + 2];
    dst[i] = v;
}
__global__ void k_cr_to_int3(int *dst, const i4 *src, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
__global__ void k_accel(double *dst, const d4 *f4, const d4 *v4, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 f = f4[i]; double m = v4[i].w;
    dst[3 * i] = f.x / m; dst[3 * i + 1] = f.y / m; dst[3 * i + 2] = f.z / m;   // source/sepintgr.c:53
}

struct FieldInfo { size_t elem; int width; };   // host element size in bytes, and how many per atom

static FieldInfo field_info(int field)
{
    switch (field) {
    case SEPGPU_F_X: case SEPGPU_F_V: case SEPGPU_F_F: case SEPGPU_F_XN:
    case SEPGPU_F_PV: case SEPGPU_F_PA: case SEPGPU_F_A: case SEPGPU_F_X0: return {sizeof(double), 3};
    case SEPGPU_F_M: case SEPGPU_F_Z: return {sizeof(double), 1};
    case SEPGPU_F_TYPE: return {1, 1};
    case SEPGPU_F_MOLINDEX: case SEPGPU_F_GID: return {sizeof(int), 1};
    case SEPGPU_F_CROSS_NEIGHB: case SEPGPU_F_CROSSINGS: return {sizeof(int), 3};
    case SEPGPU_F_BOND: case SEPGPU_F_ANGLE: return {sizeof(int), 10};
    case SEPGPU_F_DIHED: return {sizeof(int), 20};
    default: return {0, 0};
    }
}

static int ensure_dpd(sepgpu_ctx *c)
{
    if (c->have_dpd) return 0;
    if (dalloc(&c->pv4, (size_t)c->n) || dalloc(&c->pa4, (size_t)c->n)) return SEPGPU_ECUDA;
    c->have_dpd = true;
    return 0;
}
int sepgpu_ensure_dpd(sepgpu_ctx *c) { return ensure_dpd(c); }

// ---- host <-> device field movement ----------------------------------------------------------------------
// One pass over the caller's array gathers every requested field (parallel host threads; the AoS
// seppart array is 568 B per atom, so a pass per field would re-stream it each time), one H2D copy moves
// the packed block, and a small kernel per field converts to the 32-byte device records.
#include <thread>
#include <vector>

static void parallel_rows(size_t n, const std::function<void(size_t, size_t)> &fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n < 65536 ? 1 : (hw ? (hw > 16 ? 16 : hw) : 4);
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; t++) {
        const size_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        if (b < e) th.emplace_back(fn, b, e);
    }
    for (auto &x : th) x.join();
}

static int put_dispatch(sepgpu_ctx *c, int field, const void *dsrc, const void *hsrc)
{
    const FieldInfo fi = field_info(field);
    const size_t row = fi.elem * fi.width, n = (size_t)c->n_own;
    const int B = 256, G = (c->n_own + B - 1) / B;
    int rc;
    // single domain with a valid list: see k_vec3_to_d4_changed (decomposed runs rebuild collectively: no shortcut there)
    const bool cmp = !c->dd && c->list_valid && c->xs_current;
    if (cmp && !c->put_check && (field == SEPGPU_F_X || field == SEPGPU_F_TYPE || field == SEPGPU_F_MOLINDEX ||
                                 field == SEPGPU_F_BOND || field == SEPGPU_F_ANGLE || field == SEPGPU_F_DIHED)) {
        if (!c->put_changed) CUDA_TRY(cudaMalloc((void **)&c->put_changed, sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(c->put_changed, 0, sizeof(int), c->stream));
        c->put_check = true;
    }
    switch (field) {
    case SEPGPU_F_X:
        if (cmp) { k_vec3_to_d4_changed<<<G, B, 0, c->stream>>>(c->x4, (const double *)dsrc, c->n_own, c->put_changed); break; }
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->x4, (const double *)dsrc, c->n_own);
        c->xs_current = false; c->list_valid = false;
        break;
    case SEPGPU_F_V:
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->v4, (const double *)dsrc, c->n_own);
        c->mv2_valid = false;
        break;
    case SEPGPU_F_F:
        if ((rc = sepgpu_apply_pending(c))) return rc;
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->f4, (const double *)dsrc, c->n_own);
        c->f_zero = false;
        break;
    case SEPGPU_F_XN:
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->xn4, (const double *)dsrc, c->n_own);
        break;
    case SEPGPU_F_X0:
        if (!c->x0 && dalloc(&c->x0, (size_t)c->ncap)) return SEPGPU_ECUDA;
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->x0, (const double *)dsrc, c->n_own);
        break;
    case SEPGPU_F_PV:
        if ((rc = ensure_dpd(c))) return rc;
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->pv4, (const double *)dsrc, c->n_own);
        break;
    case SEPGPU_F_PA:
        if ((rc = ensure_dpd(c))) return rc;
        k_vec3_to_d4<<<G, B, 0, c->stream>>>(c->pa4, (const double *)dsrc, c->n_own);
        break;
    case SEPGPU_F_M:
        k_scalar_to_w<<<G, B, 0, c->stream>>>(c->v4, (const double *)dsrc, c->n_own);
        c->mv2_valid = false;
        break;
    case SEPGPU_F_Z: {
        CUDA_TRY(cudaMemcpyAsync(c->z, dsrc, row * n, cudaMemcpyDeviceToDevice, c->stream));
        const double *hz = (const double *)hsrc;
        bool any = false;
        for (size_t i = 0; i < n && !any; i++) any = hz[i] != 0.0;
        c->have_charge = any;
        c->zs_valid = false;
        break;
    }
    case SEPGPU_F_TYPE: {
        if (cmp) k_set_tag_changed<<<G, B, 0, c->stream>>>(c->x4, (const char *)dsrc, NULL, c->n_own, c->put_changed);
        else k_set_type<<<G, B, 0, c->stream>>>(c->x4, (const char *)dsrc, c->n_own);
        const unsigned char *ht = (const unsigned char *)hsrc;
        int st = ht[0];
        for (size_t i = 1; i < n && st >= 0; i++) if (ht[i] != ht[0]) st = -1;
        c->single_type = st;
        if (cmp) break;
        c->xs_current = false; c->list_valid = false;
        break;
    }
    case SEPGPU_F_MOLINDEX:
        if (cmp) { k_set_tag_changed<<<G, B, 0, c->stream>>>(c->x4, NULL, (const int *)dsrc, c->n_own, c->put_changed); break; }
        k_set_mol<<<G, B, 0, c->stream>>>(c->x4, (const int *)dsrc, c->n_own);
        c->xs_current = false; c->list_valid = false;
        break;
    case SEPGPU_F_CROSS_NEIGHB:
        k_int3_to_cr<<<G, B, 0, c->stream>>>(c->cr4, (const int *)dsrc, c->n_own);
        break;
    case SEPGPU_F_CROSSINGS:
        CUDA_TRY(cudaMemcpyAsync(c->crossings, dsrc, row * n, cudaMemcpyDeviceToDevice, c->stream));
        break;
    case SEPGPU_F_GID:
        if (!c->gid && dalloc(&c->gid, (size_t)c->ncap)) return SEPGPU_ECUDA;
        CUDA_TRY(cudaMemcpyAsync(c->gid, dsrc, row * n, cudaMemcpyDeviceToDevice, c->stream));
        break;
    case SEPGPU_F_BOND: case SEPGPU_F_ANGLE: case SEPGPU_F_DIHED: {
        int **tab = field == SEPGPU_F_BOND ? &c->excl_bond : field == SEPGPU_F_ANGLE ? &c->excl_angle : &c->excl_dihed;
        if (cmp && *tab) {
            const size_t count = (size_t)fi.width * n;
            k_copy_int_changed<<<(unsigned)((count + B - 1) / B), B, 0, c->stream>>>(*tab, (const int *)dsrc, count, c->put_changed);
            break;
        }
        if (!*tab && dalloc(tab, (size_t)fi.width * c->ncap)) return SEPGPU_ECUDA;
        CUDA_TRY(cudaMemcpyAsync(*tab, dsrc, row * n, cudaMemcpyDeviceToDevice, c->stream));
        c->have_excl = c->excl_bond && c->excl_angle && c->excl_dihed;
        c->list_valid = false;
        break;
    }
    }
    KERNEL_CHECK();
    return 0;
}

#define SEPGPU_MAX_FIELDS 24

extern "C" int sepgpu_put_fields(sepgpu_ctx *c, const void *base, size_t stride, int nfields,
                                 const int *fields, const size_t *offsets)
{
    if (!c || !base || nfields <= 0 || nfields > SEPGPU_MAX_FIELDS || !fields) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    const size_t n = (size_t)c->n_own;
    size_t row[SEPGPU_MAX_FIELDS], off[SEPGPU_MAX_FIELDS], hoff[SEPGPU_MAX_FIELDS], total = 0;
    for (int f = 0; f < nfields; f++) {
        const FieldInfo fi = field_info(fields[f]);
        if (!fi.elem || fields[f] == SEPGPU_F_A) { sepgpu_set_error("sepgpu_put: bad field %d", fields[f]); return SEPGPU_EINVAL; }
        row[f] = fi.elem * fi.width;
        hoff[f] = offsets ? offsets[f] : 0;
        off[f] = total;
        total += (row[f] * n + 255) & ~(size_t)255;
    }
    if (stride == 0) { if (nfields != 1) return SEPGPU_EINVAL; stride = row[0]; }
    int rc = sepgpu_ensure_stage(c, total);
    if (rc) return rc;
    // the previous async copy out of the staging buffer must be done before we overwrite it
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const char *src = (const char *)base;
    char *dst = (char *)c->stage;
    const int *rows = c->host_rows;                 // (SEP_NGPU) device atom i lives in host record rows[i]
    if (!rows && nfields == 1 && stride == row[0]) memcpy(dst, src + hoff[0], row[0] * n);
    else parallel_rows(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            const char *rec = src + (rows ? (size_t)rows[i] : i) * stride;
            for (int f = 0; f < nfields; f++) memcpy(dst + off[f] + i * row[f], rec + hoff[f], row[f]);
        }
    });
    CUDA_TRY(cudaMemcpyAsync(c->dstage, c->stage, total, cudaMemcpyHostToDevice, c->stream));
    for (int f = 0; f < nfields; f++)
        if ((rc = put_dispatch(c, fields[f], (const char *)c->dstage + off[f], (const char *)c->stage + off[f]))) return rc;
    if (c->put_check) {                    // did positions / types / molecule indices / exclusion tables come back different?
        int changed = 1;
        c->put_check = false;
        CUDA_TRY(cudaMemcpyAsync(&changed, c->put_changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (changed) { c->xs_current = false; c->list_valid = false; }
    }
    return 0;
}

// Host records are addressed through `rows` from now on (NULL: record i = device atom i).  The array stays the caller's and
// must hold one entry per owned atom at every later put/get.
extern "C" int sepgpu_set_host_rows(sepgpu_ctx *c, const int *rows)
{
    if (!c) return SEPGPU_EINVAL;
    c->host_rows = rows;
    return 0;
}

extern "C" int sepgpu_put(sepgpu_ctx *c, int field, const void *host, size_t stride)
{
    return sepgpu_put_fields(c, host, stride, 1, &field, NULL);
}

// fills ddst (device, packed) for `field`, or returns a device array that already has the packed layout
static int get_prepare(sepgpu_ctx *c, int field, void *ddst, const void **direct)
{
    const FieldInfo fi = field_info(field);
    const size_t row = fi.elem * fi.width, n = (size_t)c->n_own;
    const int B = 256, G = (c->n_own + B - 1) / B;
    int rc;
    *direct = NULL;
    switch (field) {
    case SEPGPU_F_X: k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, c->x4, c->n_own); break;
    case SEPGPU_F_V: k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, c->v4, c->n_own); break;
    case SEPGPU_F_F:
        if ((rc = sepgpu_apply_pending(c))) return rc;
        if (c->f_zero) CUDA_TRY(cudaMemsetAsync(ddst, 0, row * n, c->stream));
        else k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, c->f4, c->n_own);
        break;
    case SEPGPU_F_A:
        if ((rc = sepgpu_apply_pending(c))) return rc;
        if (c->f_zero) CUDA_TRY(cudaMemsetAsync(ddst, 0, row * n, c->stream));
        else k_accel<<<G, B, 0, c->stream>>>((double *)ddst, c->f4, c->v4, c->n_own);
        break;
    case SEPGPU_F_XN: k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, c->xn4, c->n_own); break;
    case SEPGPU_F_X0:
        if (!c->x0) { sepgpu_set_error("sepgpu_get: no tether positions on the device"); return SEPGPU_ESTATE; }
        k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, c->x0, c->n_own);
        break;
    case SEPGPU_F_PV: case SEPGPU_F_PA:
        if (!c->have_dpd) { sepgpu_set_error("sepgpu_get: no DPD state"); return SEPGPU_ESTATE; }
        k_d4_to_vec3<<<G, B, 0, c->stream>>>((double *)ddst, field == SEPGPU_F_PV ? c->pv4 : c->pa4, c->n_own);
        break;
    case SEPGPU_F_M: k_w_to_scalar<<<G, B, 0, c->stream>>>((double *)ddst, c->v4, c->n_own); break;
    case SEPGPU_F_TYPE: k_get_type<<<G, B, 0, c->stream>>>((char *)ddst, c->x4, c->n_own); break;
    case SEPGPU_F_MOLINDEX: k_get_mol<<<G, B, 0, c->stream>>>((int *)ddst, c->x4, c->n_own); break;
    case SEPGPU_F_CROSS_NEIGHB: k_cr_to_int3<<<G, B, 0, c->stream>>>((int *)ddst, c->cr4, c->n_own); break;
    case SEPGPU_F_Z: *direct = c->z; break;
    case SEPGPU_F_CROSSINGS: *direct = c->crossings; break;
    case SEPGPU_F_GID: *direct = c->gid; break;
    case SEPGPU_F_BOND: *direct = c->excl_bond; break;
    case SEPGPU_F_ANGLE: *direct = c->excl_angle; break;
    case SEPGPU_F_DIHED: *direct = c->excl_dihed; break;
    default: return SEPGPU_EINVAL;
    }
    KERNEL_CHECK();
    if (field == SEPGPU_F_Z || field == SEPGPU_F_CROSSINGS || (field >= SEPGPU_F_BOND && field <= SEPGPU_F_GID)) {
        if (!*direct) { sepgpu_set_error("sepgpu_get: field %d not present on device", field); return SEPGPU_ESTATE; }
        CUDA_TRY(cudaMemcpyAsync(ddst, *direct, row * n, cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

extern "C" int sepgpu_get_fields(sepgpu_ctx *c, void *base, size_t stride, int nfields,
                                 const int *fields, const size_t *offsets)
{
    if (!c || !base || nfields <= 0 || nfields > SEPGPU_MAX_FIELDS || !fields) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    c->get_calls++;
    const size_t n = (size_t)c->n_own;
    size_t row[SEPGPU_MAX_FIELDS], off[SEPGPU_MAX_FIELDS], hoff[SEPGPU_MAX_FIELDS], total = 0;
    for (int f = 0; f < nfields; f++) {
        const FieldInfo fi = field_info(fields[f]);
        if (!fi.elem) { sepgpu_set_error("sepgpu_get: bad field %d", fields[f]); return SEPGPU_EINVAL; }
        row[f] = fi.elem * fi.width;
        hoff[f] = offsets ? offsets[f] : 0;
        off[f] = total;
        total += (row[f] * n + 255) & ~(size_t)255;
    }
    if (stride == 0) { if (nfields != 1) return SEPGPU_EINVAL; stride = row[0]; }
    int rc = sepgpu_ensure_stage(c, total);
    if (rc) return rc;
    for (int f = 0; f < nfields; f++) {
        const void *direct;
        if ((rc = get_prepare(c, fields[f], (char *)c->dstage + off[f], &direct))) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(c->stage, c->dstage, total, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    char *dst = (char *)base;
    const char *src = (const char *)c->stage;
    const int *rows = c->host_rows;
    if (!rows && nfields == 1 && stride == row[0]) memcpy(dst + hoff[0], src, row[0] * n);
    else parallel_rows(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            char *rec = dst + (rows ? (size_t)rows[i] : i) * stride;
            for (int f = 0; f < nfields; f++) memcpy(rec + hoff[f], src + off[f] + i * row[f], row[f]);
        }
    });
    return 0;
}

extern "C" int sepgpu_get(sepgpu_ctx *c, int field, void *host, size_t stride)
{
    return sepgpu_get_fields(c, host, stride, 1, &field, NULL);
}


// ---- molecule-molecule force table (reference sepmolinfo.Fij, include/sepstrct.h:94) ---------------------------
extern "C" int sepgpu_fij_enable(sepgpu_ctx *c, int nmol)
{
    if (!c || nmol <= 0) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("fij_enable: not available in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    if (c->fij && c->nmol == nmol) return 0;
    if (c->fij) { cudaFree(c->fij); c->fij = NULL; }
    CUDA_TRY(cudaMalloc((void **)&c->fij, sizeof(double) * 3 * (size_t)nmol * nmol));
    CUDA_TRY(cudaMemsetAsync(c->fij, 0, sizeof(double) * 3 * (size_t)nmol * nmol, c->stream));
    c->nmol = nmol;
    return 0;
}

extern "C" int sepgpu_fij_reset(sepgpu_ctx *c)
{
    if (!c || !c->fij) return SEPGPU_ESTATE;
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaMemsetAsync(c->fij, 0, sizeof(double) * 3 * (size_t)c->nmol * c->nmol, c->stream));
    return 0;
}

__global__ void k_fij_to_float(const double *__restrict__ in, float *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// the table as nmol*nmol*3 floats (the reference's element type), row-major [i][j][k]
extern "C" int sepgpu_fij_get(sepgpu_ctx *c, float *out)
{
    if (!c || !out || !c->fij) return SEPGPU_ESTATE;
    SEPGPU_ENTER(c);
    const size_t cnt = 3 * (size_t)c->nmol * c->nmol;
    int rc = sepgpu_ensure_stage(c, cnt * sizeof(float));
    if (rc) return rc;
    k_fij_to_float<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(c->fij, (float *)c->dstage, cnt);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(c->stage, c->dstage, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out, c->stage, cnt * sizeof(float));
    return 0;
}

// ---- scalars ----------------------------------------------------------------------------------------------
__global__ void k_reset_ret(DevScalars *s)
{
    // sep_reset_retval (source/sepret.c:19-47)
    int t = threadIdx.x;
    if (t == 0) { s->epot = 0; s->ecoul = 0; s->ekin = 0; }
    if (t < 9) { s->pot_P[t] = 0; s->kin_P[t] = 0; s->pot_P_bond[t] = 0; }
}

extern "C" int sepgpu_reset_ret(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    // no launch: the next kernel that accumulates into the scalar block clears it first
    // (sepgpu_flush_resets does it explicitly when the block is read before that)
    c->ret_reset_pending = true;
    return 0;
}

__global__ void k_reset_maxdist(DevScalars *s) { s->max_dist2 = 0.0; }

extern "C" int sepgpu_reset_force(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    // f <- 0 is not written out: the first force kernel after this call stores instead of adding.
    c->f_zero = true;
    c->pending_alpha_slot = -1;       // a pending f -= alpha m v dies with the force it would modify
    c->maxd_reset_pending = true;                         // source/sepmisc.c:399, applied by the next integrator
    return 0;
}

int sepgpu_flush_resets(sepgpu_ctx *c)
{
    if (c->ret_reset_pending) { k_reset_ret<<<1, 32, 0, c->stream>>>(c->scal); c->ret_reset_pending = false; }
    if (c->maxd_reset_pending) { k_reset_maxdist<<<1, 1, 0, c->stream>>>(c->scal); c->maxd_reset_pending = false; }
    KERNEL_CHECK();
    return 0;
}

extern "C" int sepgpu_read_scalars(sepgpu_ctx *c, sepgpu_scalars *out)
{
    if (!c || !out) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    // straight after an integrator call the block is on the host already (it was fetched with the rebuild trigger)
    const bool cached = c->scal_cache_valid && c->scal_cache_seq == c->api_seq && !c->ret_reset_pending && !c->maxd_reset_pending;
    if (!cached) {
        { int rcf = sepgpu_flush_resets(c); if (rcf) return rcf; }
        CUDA_TRY(cudaMemcpyAsync(c->scal_host, c->scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->scal_cache_valid = false;
    }
    const DevScalars *s = c->scal_host;
    if (s->error == SEPGPU_ETABLE) {
        sepgpu_set_error("a pair came closer than the tabulated pair function reaches (SEP_TABLE_RMIN)");
        return SEPGPU_ETABLE;
    }
    out->epot = s->epot; out->ecoul = s->ecoul; out->ekin = s->ekin;
    memcpy(out->pot_P, s->pot_P, sizeof out->pot_P);
    memcpy(out->kin_P, s->kin_P, sizeof out->kin_P);
    memcpy(out->pot_P_bond, s->pot_P_bond, sizeof out->pot_P_bond);
    out->max_dist2 = s->max_dist2;
    out->sum_mv2 = s->sum_mv2;
    memcpy(out->alpha, s->alpha, sizeof out->alpha);
    out->neighb_flag = s->neighb_flag;
    out->nbuild = s->nbuild;
    out->error = s->error;
    out->max_neighb = s->max_neighb;
    out->npairs_listed = s->npairs_listed;
    if (c->dd) {
        // decomposed run: integrator-derived values are already global; force-derived sums are per rank
        int rc = sepgpu_dd_reduce_force_scalars(c, &out->epot, &out->ecoul, out->pot_P, out->pot_P_bond);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int sepgpu_sync(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

__global__ void k_set_alpha(DevScalars *s, int slot, double a) { s->alpha[slot] = a; }

extern "C" int sepgpu_set_alpha(sepgpu_ctx *c, int slot, double alpha)
{
    if (!c || slot < 0 || slot > 3) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    k_set_alpha<<<1, 1, 0, c->stream>>>(c->scal, slot, alpha);
    KERNEL_CHECK();
    return 0;
}

__global__ void k_set_flag(DevScalars *s) { s->neighb_flag = 1; }

extern "C" int sepgpu_request_rebuild(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    c->list_valid = false;
    k_set_flag<<<1, 1, 0, c->stream>>>(c->scal);
    KERNEL_CHECK();
    return 0;
}

static int ktimer_enable(KernelTimer *t)
{
    if (t->enabled) return 0;
    for (int i = 0; i < 64; i++) {
        CUDA_TRY(cudaEventCreate(&t->start[i]));
        CUDA_TRY(cudaEventCreate(&t->stop[i]));
    }
    t->enabled = true; t->used = 0; t->total_ms = 0; t->launches = 0;
    return 0;
}

static void ktimer_drain(KernelTimer *t)
{
    for (int i = 0; i < t->used; i++) {
        float ms = 0;
        cudaEventSynchronize(t->stop[i]);
        cudaEventElapsedTime(&ms, t->start[i], t->stop[i]);
        t->total_ms += ms; t->launches++;
    }
    t->used = 0;
}

void ktimer_begin(sepgpu_ctx *c, KernelTimer *t)
{
    if (!t->enabled) return;
    if (t->used == 64) ktimer_drain(t);
    cudaEventRecord(t->start[t->used], c->stream);
}
void ktimer_end(sepgpu_ctx *c, KernelTimer *t)
{
    if (!t->enabled) return;
    cudaEventRecord(t->stop[t->used], c->stream);
    t->used++;
}

extern "C" int sepgpu_set_option(sepgpu_ctx *c, const char *name, long long value)
{
    if (!c || !name) return SEPGPU_EINVAL;
    if (!strcmp(name, "tpa")) {
        if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16 && value != 32 && value != 0)
            return SEPGPU_EINVAL;
        c->tpa = value ? (int)(value > 8 ? 8 : value) : 1;
        return 0;
    }
    if (!strcmp(name, "prefilter")) { c->prefilter = value != 0; return 0; }
    if (!strcmp(name, "overlap")) { c->overlap = value != 0; return 0; }
    if (!strcmp(name, "step_fold")) { if (!value) { int rs = sepgpu_settle(c); if (rs) return rs; } c->step_fold = value != 0; return 0; }
    if (!strcmp(name, "fin_multi")) { c->fin_multi = value != 0; return 0; }
    if (!strcmp(name, "spec_force")) { c->spec.on = value == 2 ? 2 : value != 0; c->spec.streak = 0; return 0; }   // 2: also decomposed (experimental)
    if (!strcmp(name, "build_window")) { if (value < 0 || value > 2) return SEPGPU_EINVAL; c->build_window = (int)value; return 0; }
    if (!strcmp(name, "tile_list")) { c->tile_list = value != 0; c->list_valid = false; return 0; }
    if (!strcmp(name, "coulomb_kernel")) { if (value != 1 && value != 2) return SEPGPU_EINVAL; c->coulomb_kernel = (int)value; return 0; }
    if (!strcmp(name, "typed_sublist")) { c->typed_sublist = value != 0; return 0; }
    if (!strcmp(name, "force_grid")) { c->force_grid = value > 0 && value <= SEPGPU_MAX_BLOCKS_PARTIAL ? (int)value : 0; return 0; }
    if (!strcmp(name, "time_kernels")) {
        if (value) { if (ktimer_enable(&c->t_force) || ktimer_enable(&c->t_build) || ktimer_enable(&c->t_intgr) ||
                         ktimer_enable(&c->t_coul) || ktimer_enable(&c->t_bonded) || ktimer_enable(&c->t_halo) ||
                         ktimer_enable(&c->t_migr)) return SEPGPU_ECUDA; }
        return 0;
    }
    if (!strcmp(name, "neighb_cap")) {
        if (value < 8) return SEPGPU_EINVAL;
        if (c->nbr) { cudaFree(c->nbr); c->nbr = NULL; }
        c->cap = (int)value; c->list_valid = false;
        return 0;
    }
    sepgpu_set_error("sepgpu_set_option: unknown option '%s'", name);
    return SEPGPU_EINVAL;
}

int sepgpu_dd_uses_p2p(sepgpu_ctx *c);

extern "C" int sepgpu_get_option(sepgpu_ctx *c, const char *name, long long *value)
{
    if (!c || !name || !value) return SEPGPU_EINVAL;
    if (!strcmp(name, "tpa")) *value = c->tpa;
    else if (!strcmp(name, "prefilter")) *value = c->prefilter;
    else if (!strcmp(name, "overlap")) *value = c->overlap;
    else if (!strcmp(name, "force_grid")) *value = c->force_grid;
    else if (!strcmp(name, "neighb_cap")) *value = c->cap;
    else if (!strcmp(name, "coulomb_kernel")) *value = c->coulomb_kernel;
    else if (!strcmp(name, "typed_sublist")) *value = c->typed_sublist;
    else if (!strcmp(name, "tile_list")) *value = c->tile_list;
    else if (!strcmp(name, "build_window")) *value = c->build_window;
    else if (!strcmp(name, "fin_multi")) *value = c->fin_multi;
    else if (!strcmp(name, "step_fold")) *value = c->step_fold;
    else if (!strcmp(name, "spec_force")) *value = c->spec.on;
    else if (!strcmp(name, "spec_adopted")) *value = c->spec_adopted;
    else if (!strcmp(name, "feed_calls")) *value = c->feed_calls;          // sampler feeds served so far (sepgpu_feeds.cu)
    else if (!strcmp(name, "get_calls")) *value = c->get_calls;            // per-atom downloads served so far (sepgpu_get_fields)
    else if (!strcmp(name, "list_f16")) *value = c->list_valid && c->list_f16 ? 1 : 0;       // rows of 16-bit tile slots
    else if (!strcmp(name, "tile_R")) *value = c->tile_R;
    else if (!strcmp(name, "tile_stage")) *value = c->tile_stage_used;
    else if (!strcmp(name, "dd_p2p")) *value = sepgpu_dd_uses_p2p(c);               // decomposed run on the peer-memory path
    else if (!strcmp(name, "max_half")) *value = c->scal_host->max_half;            // longest reference-style half list, last build
    else { sepgpu_set_error("sepgpu_get_option: unknown option '%s'", name); return SEPGPU_EINVAL; }
    return 0;
}

extern "C" int sepgpu_kernel_time(sepgpu_ctx *c, const char *which, float *ms_total, int *launches)
{
    if (!c || !which) return SEPGPU_EINVAL;
    KernelTimer *t = !strcmp(which, "force") ? &c->t_force : !strcmp(which, "build") ? &c->t_build
                   : !strcmp(which, "intgr") ? &c->t_intgr : !strcmp(which, "coulomb") ? &c->t_coul
                   : !strcmp(which, "bonded") ? &c->t_bonded : !strcmp(which, "halo") ? &c->t_halo
                   : !strcmp(which, "migrate") ? &c->t_migr : NULL;
    if (!t || !t->enabled) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    ktimer_drain(t);
    if (ms_total) *ms_total = t->total_ms;
    if (launches) *launches = t->launches;
    t->total_ms = 0; t->launches = 0;
    return 0;
}

extern "C" int sepgpu_timer_start(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    return 0;
}

extern "C" int sepgpu_timer_stop(sepgpu_ctx *c, float *ms)
{
    if (!c || !ms) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

__global__ void k_flush(double *p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = (double)i;
}

extern "C" int sepgpu_flush_l2(sepgpu_ctx *c)
{
    if (!c) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20;           // 256 MiB > 126 MB L2
        CUDA_TRY(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    k_flush<<<148 * 8, 256, 0, c->stream>>>((double *)c->flush_buf, c->flush_bytes / sizeof(double));
    KERNEL_CHECK();
    return 0;
}

// ---- box ceilings -------------------------------------------------------------------------------------------
// 8 independent DFMA chains per thread; 2 flop per DFMA.
__global__ void __launch_bounds__(256) k_fma64(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, cc = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
        a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int sepgpu_peak_fp64(int device, double *tflops)
{
    if (!tflops) return SEPGPU_EINVAL;
    if (sepgpu_device_count() <= 0) { sepgpu_set_error("no CUDA device"); return SEPGPU_ENODEV; }
    CUDA_TRY(cudaSetDevice(device < 0 ? 0 : device));
    const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
    double *out;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        k_fma64<<<blocks, threads>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms; CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return 0;
}

__global__ void k_copy(const double4 *__restrict__ a, double4 *__restrict__ b, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

extern "C" int sepgpu_peak_copy(int device, double *gbs)
{
    if (!gbs) return SEPGPU_EINVAL;
    if (sepgpu_device_count() <= 0) { sepgpu_set_error("no CUDA device"); return SEPGPU_ENODEV; }
    CUDA_TRY(cudaSetDevice(device < 0 ? 0 : device));
    const size_t bytes = (size_t)1 << 30, n = bytes / sizeof(double4);
    double4 *a, *b;
    CUDA_TRY(cudaMalloc(&a, bytes)); CUDA_TRY(cudaMalloc(&b, bytes));
    CUDA_TRY(cudaMemset(a, 1, bytes));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        k_copy<<<148 * 16, 512>>>(a, b, n);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms; CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(a); cudaFree(b);
    *gbs = best;
    return 0;
}
