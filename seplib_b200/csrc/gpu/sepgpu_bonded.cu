// sepgpu_bonded.cu -- bond / angle / dihedral forces without atomics.
//
// Stand-ins for sep_stretch_harmonic, sep_angle_harmonic, sep_angle_cossq and sep_torsion_Ryckaert
// (reference source/sepmol.c:372-414, 469-516, 418-467, 520-587).  The reference loops over terms and
// scatters into 2/3/4 atoms.  Here an inverse topology (CSR: atom -> (term, role)) is built once on the
// host when the lists are set, and one thread per ATOM re-evaluates every term it takes part in and
// adds its own share in ascending term order -- the same order in which the reference's term loop
// touches that atom, so the per-atom force sum is reproduced bit for bit.  Energy, virial and the
// per-term observables (blengths/angles/dihedrals) are emitted by the role-0 atom only.
// Term geometry uses the reference's exact arithmetic: wrapped positions, sep_Wrap branches,
// sep_dot's left-to-right sum (source/seputil.c:393-403), no FMA contraction.
#include "sepgpu_internal.cuh"

#include <stdlib.h>
#include <math.h>

int sepgpu_dd_nglobal(sepgpu_ctx *c);
int sepgpu_dd_halo_update(sepgpu_ctx *c, const sepgpu_sys *sys);
int sepgpu_dd_gmap(sepgpu_ctx *c, const int **gmap);

#define BONDED_BLOCK 128
#define SEPGPU_PI 3.14159265358979      // SEP_PI, include/sepdef.h:40

struct BoxB { double Lx, Ly, Lz; };

// Slab-decomposed runs (sepgpu_dd.cu): the topology lists and the inverse topology are indexed by GLOBAL atom ids, the
// per-atom arrays by local rows.  A thread owns one of the rank's own rows; partners are found through the global-id ->
// row map of the last rebuild and may be halo atoms, whose current coordinates sit in the halo slots of the sorted copy
// xs (continuous coordinates: the wrap of the separation below absorbs the box shifts they may carry).  Every term is
// evaluated by each rank that owns one of its atoms, for that atom only; energy and virial by the owner of its first atom
// (SURVEY.md section 8e).  on == 0: single domain, everything is indexed by atom.
struct BondedDD {
    const int *gid, *gmap, *rank;
    const d4 *xs;
    int n_own, on;
};

__device__ __forceinline__ d4 atom_pos(const d4 *__restrict__ x4, const BondedDD &D, unsigned a, int *missing)
{
    if (!D.on) return x4[a];
    const int row = D.gmap[a];
    if (row < 0) { *missing = 1; return x4[0]; }
    return row < D.n_own ? x4[row] : D.xs[D.rank[row]];
}

__device__ __forceinline__ void diff_wrap(const d4 &a, const d4 &b, const BoxB &B, double r[3])
{
    r[0] = wrap_exact(__dsub_rn(a.x, b.x), B.Lx, 0.5 * B.Lx);
    r[1] = wrap_exact(__dsub_rn(a.y, b.y), B.Ly, 0.5 * B.Ly);
    r[2] = wrap_exact(__dsub_rn(a.z, b.z), B.Lz, 0.5 * B.Lz);
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3])
{
    return __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(a[0], b[0])), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
}

// ---- bonds ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BONDED_BLOCK)
k_bond(const d4 *__restrict__ x4, d4 *__restrict__ f4, int n, const unsigned *__restrict__ blist,
       const int *__restrict__ aptr, const int *__restrict__ aidx, int type, double lbond, double ks,
       BoxB B, int f_zero, double *__restrict__ blengths, double *__restrict__ partial, BondedDD D, DevScalars *scal)
{
    __shared__ double red[SEPGPU_NPART_F * (BONDED_BLOCK / 32)];
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    const int i = blockIdx.x * BONDED_BLOCK + threadIdx.x;
    int missing = 0;
    if (i < n) {
        const int ig = D.on ? D.gid[i] : i;                      // the atom's index in the topology
        const int b0 = aptr[ig], b1 = aptr[ig + 1];
        if (b1 > b0 || f_zero) {
            d4 f;
            if (f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
            for (int e = b0; e < b1; e++) {
                const int term = aidx[e] >> 2, role = aidx[e] & 3;
                if ((int)blist[3 * term + 2] != type) continue;
                const unsigned a = blist[3 * term], b = blist[3 * term + 1];
                double r[3];
                diff_wrap(atom_pos(x4, D, a, &missing), atom_pos(x4, D, b, &missing), B, r);
                const double r2 = dot3(r, r);
                const double dist = sqrt(r2);
                const double ft = -ks * (dist - lbond) / dist;            // source/sepmol.c:394
                const double gx = ft * r[0], gy = ft * r[1], gz = ft * r[2];
                if (role == 0) {
                    f.x += gx; f.y += gy; f.z += gz;
                    acc[0] += 0.5 * ks * (dist - lbond) * (dist - lbond);  // :409
                    acc[2] += gx * r[0]; acc[3] += gx * r[1]; acc[4] += gx * r[2];
                    acc[5] += gy * r[1]; acc[6] += gy * r[2]; acc[7] += gz * r[2];
                    blengths[term] = dist;
                } else {
                    f.x -= gx; f.y -= gy; f.z -= gz;
                }
            }
            f4[i] = f;
        }
    }
    if (missing) scal->error = SEPGPU_ECELL;                       // a bonded partner is neither mine nor in my halo
    block_sum<SEPGPU_NPART_F, BONDED_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

// ---- angles ---------------------------------------------------------------------------------------------------
template <bool COSSQ>
__global__ void __launch_bounds__(BONDED_BLOCK)
k_angle(const d4 *__restrict__ x4, d4 *__restrict__ f4, int n, const unsigned *__restrict__ alist,
        const int *__restrict__ aptr, const int *__restrict__ aidx, int type, double angle0, double kc,
        double cCon, BoxB B, int f_zero, double *__restrict__ angles, double *__restrict__ partial, BondedDD D, DevScalars *scal)
{
    __shared__ double red[SEPGPU_NPART_F * (BONDED_BLOCK / 32)];
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    const int i = blockIdx.x * BONDED_BLOCK + threadIdx.x;
    int missing = 0;
    if (i < n) {
        const int ig = D.on ? D.gid[i] : i;                      // the atom's index in the topology
        const int b0 = aptr[ig], b1 = aptr[ig + 1];
        if (b1 > b0 || f_zero) {
            d4 f;
            if (f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
            for (int e = b0; e < b1; e++) {
                const int term = aidx[e] >> 2, role = aidx[e] & 3;
                if ((int)alist[4 * term + 3] != type) continue;
                const unsigned a = alist[4 * term], b = alist[4 * term + 1], c = alist[4 * term + 2];
                double d1[3], d2[3];
                const d4 xb = atom_pos(x4, D, b, &missing);
                diff_wrap(xb, atom_pos(x4, D, a, &missing), B, d1);        // dr1 = x_b - x_a
                diff_wrap(atom_pos(x4, D, c, &missing), xb, B, d2);        // dr2 = x_c - x_b
                const double c11 = dot3(d1, d1), c12 = dot3(d1, d2), c22 = dot3(d2, d2);
                const double cD = sqrt(c11 * c22);
                double fm, en, ang;
                if (COSSQ) {                                               // source/sepmol.c:447-464
                    const double cc = c12 / cD;
                    fm = -kc * (cc - cCon);
                    en = 0.5 * kc * (cc - cCon) * (cc - cCon);
                    ang = SEPGPU_PI - acos(cc);
                } else {                                                   // source/sepmol.c:496-512
                    ang = SEPGPU_PI - acos(c12 / cD);
                    fm = -kc * (ang - angle0);
                    en = 0.5 * kc * (ang - angle0) * (ang - angle0);
                }
                double g[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double f1 = fm * ((c12 / c11) * d1[k] - d2[k]) / cD;
                    const double f2 = fm * (d1[k] - (c12 / c22) * d2[k]) / cD;
                    g[k] = role == 0 ? f1 : (role == 1 ? (-f1 - f2) : f2);
                }
                f.x += g[0]; f.y += g[1]; f.z += g[2];
                if (role == 0) { acc[0] += en; angles[term] = ang; }
            }
            f4[i] = f;
        }
    }
    if (missing) scal->error = SEPGPU_ECELL;                       // a bonded partner is neither mine nor in my halo
    block_sum<SEPGPU_NPART_F, BONDED_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

// ---- dihedrals --------------------------------------------------------------------------------------------------
struct RBCoef { double g[6]; };

__global__ void __launch_bounds__(BONDED_BLOCK)
k_torsion(const d4 *__restrict__ x4, d4 *__restrict__ f4, int n, const unsigned *__restrict__ dlist,
          const int *__restrict__ aptr, const int *__restrict__ aidx, int type, RBCoef G, BoxB B,
          int f_zero, double *__restrict__ dihedrals, double *__restrict__ partial, BondedDD D, DevScalars *scal)
{
    __shared__ double red[SEPGPU_NPART_F * (BONDED_BLOCK / 32)];
    double acc[SEPGPU_NPART_F];
#pragma unroll
    for (int q = 0; q < SEPGPU_NPART_F; q++) acc[q] = 0.0;
    const double *g = G.g;
    const int i = blockIdx.x * BONDED_BLOCK + threadIdx.x;
    int missing = 0;
    if (i < n) {
        const int ig = D.on ? D.gid[i] : i;                      // the atom's index in the topology
        const int b0 = aptr[ig], b1 = aptr[ig + 1];
        if (b1 > b0 || f_zero) {
            d4 f;
            if (f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
            for (int e = b0; e < b1; e++) {
                const int term = aidx[e] >> 2, role = aidx[e] & 3;
                if ((int)dlist[5 * term + 4] != type) continue;
                const unsigned a = dlist[5 * term], b = dlist[5 * term + 1], c = dlist[5 * term + 2], d = dlist[5 * term + 3];
                double d1[3], d2[3], d3[3];
                const d4 xb = atom_pos(x4, D, b, &missing), xc = atom_pos(x4, D, c, &missing);
                diff_wrap(xb, atom_pos(x4, D, a, &missing), B, d1);
                diff_wrap(xc, xb, B, d2);
                diff_wrap(atom_pos(x4, D, d, &missing), xc, B, d3);
                const double c11 = dot3(d1, d1), c12 = dot3(d1, d2), c13 = dot3(d1, d3);
                const double c22 = dot3(d2, d2), c23 = dot3(d2, d3), c33 = dot3(d3, d3);
                const double cA = c13 * c22 - c12 * c23;                   // source/sepmol.c:555-559
                const double cB1 = c11 * c22 - c12 * c12;
                const double cB2 = c22 * c33 - c23 * c23;
                const double cD = sqrt(cB1 * cB2);
                const double cc = cA / cD;
                const double fm = -(g[1] + (2. * g[2] + (3. * g[3] + (4. * g[4] + 5. * g[5] * cc) * cc) * cc) * cc);
                const double t1 = cA, t2 = c11 * c23 - c12 * c13, t3 = -cB1;
                const double t4 = cB2, t5 = c13 * c23 - c12 * c33, t6 = -cA;
                const double cR1 = c12 / c22, cR2 = c23 / c22;
                double gk[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double f1 = fm * c22 * (t1 * d1[k] + t2 * d2[k] + t3 * d3[k]) / (cD * cB1);
                    const double f2 = fm * c22 * (t4 * d1[k] + t5 * d2[k] + t6 * d3[k]) / (cD * cB2);
                    gk[k] = role == 0 ? f1
                          : role == 1 ? (-(1.0 + cR1) * f1 + cR2 * f2)
                          : role == 2 ? (cR1 * f1 - (1.0 + cR2) * f2)
                          : f2;
                }
                f.x += gk[0]; f.y += gk[1]; f.z += gk[2];
                if (role == 0) {
                    acc[0] += g[0] + (g[1] + (g[2] + (g[3] + (g[4] + g[5] * cc) * cc) * cc) * cc) * cc;
                    dihedrals[term] = SEPGPU_PI - acos(cc);
                }
            }
            f4[i] = f;
        }
    }
    if (missing) scal->error = SEPGPU_ECELL;                       // a bonded partner is neither mine nor in my halo
    block_sum<SEPGPU_NPART_F, BONDED_BLOCK>(acc, red);
    if (threadIdx.x == 0)
        for (int q = 0; q < SEPGPU_NPART_F; q++) partial[blockIdx.x * SEPGPU_NPART_F + q] = acc[q];
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int upload(void **dst, const void *src, size_t bytes)
{
    if (*dst) { cudaFree(*dst); *dst = NULL; }
    CUDA_TRY(cudaMalloc(dst, bytes ? bytes : 4));
    if (bytes) CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

// CSR atom -> (term<<2 | role), entries of one atom in ascending term order
static int build_inverse(int n, const unsigned *list, unsigned nterms, int width, int natoms_per_term,
                         int **d_ptr, int **d_idx)
{
    int *ptr = (int *)calloc((size_t)n + 1, sizeof(int));
    if (!ptr) return SEPGPU_EINVAL;
    for (unsigned t = 0; t < nterms; t++)
        for (int r = 0; r < natoms_per_term; r++) {
            unsigned a = list[(size_t)t * width + r];
            if (a >= (unsigned)n) { free(ptr); sepgpu_set_error("topology: atom index %u out of range", a); return SEPGPU_EINVAL; }
            ptr[a + 1]++;
        }
    for (int i = 0; i < n; i++) ptr[i + 1] += ptr[i];
    const size_t total = (size_t)ptr[n];
    int *idx = (int *)malloc(sizeof(int) * (total ? total : 1));
    int *fill = (int *)calloc((size_t)n, sizeof(int));
    for (unsigned t = 0; t < nterms; t++)
        for (int r = 0; r < natoms_per_term; r++) {
            unsigned a = list[(size_t)t * width + r];
            idx[ptr[a] + fill[a]++] = (int)(t << 2) | r;
        }
    int rc = upload((void **)d_ptr, ptr, sizeof(int) * ((size_t)n + 1));
    if (!rc) rc = upload((void **)d_idx, idx, sizeof(int) * total);
    free(ptr); free(idx); free(fill);
    return rc;
}

extern "C" int sepgpu_set_topology(sepgpu_ctx *c, const unsigned *blist, unsigned nb,
                                   const unsigned *alist, unsigned na,
                                   const unsigned *dlist, unsigned nd)
{
    if (!c) return SEPGPU_EINVAL;
    if (nb >= (1u << 29) || na >= (1u << 29) || nd >= (1u << 29)) return SEPGPU_EINVAL;
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int rc;
    c->nb = nb; c->na = na; c->nd = nd;
    if ((rc = upload((void **)&c->blist, blist, sizeof(unsigned) * 3 * (size_t)nb))) return rc;
    if ((rc = upload((void **)&c->alist, alist, sizeof(unsigned) * 4 * (size_t)na))) return rc;
    if ((rc = upload((void **)&c->dlist, dlist, sizeof(unsigned) * 5 * (size_t)nd))) return rc;
    const int natoms = c->dd ? sepgpu_dd_nglobal(c) : c->n;        // decomposed: lists and inverse topology by global id
    if ((rc = build_inverse(natoms, blist, nb, 3, 2, &c->atom_bond_ptr, &c->atom_bond_idx))) return rc;
    if ((rc = build_inverse(natoms, alist, na, 4, 3, &c->atom_angle_ptr, &c->atom_angle_idx))) return rc;
    if ((rc = build_inverse(natoms, dlist, nd, 5, 4, &c->atom_dihed_ptr, &c->atom_dihed_idx))) return rc;
    if (c->blengths) cudaFree(c->blengths);
    if (c->angles) cudaFree(c->angles);
    if (c->dihedrals) cudaFree(c->dihedrals);
    c->blengths = c->angles = c->dihedrals = NULL;
    CUDA_TRY(cudaMalloc((void **)&c->blengths, sizeof(double) * (nb ? nb : 1)));
    CUDA_TRY(cudaMalloc((void **)&c->angles, sizeof(double) * (na ? na : 1)));
    CUDA_TRY(cudaMalloc((void **)&c->dihedrals, sizeof(double) * (nd ? nd : 1)));
    CUDA_TRY(cudaMemset(c->blengths, 0, sizeof(double) * (nb ? nb : 1)));
    CUDA_TRY(cudaMemset(c->angles, 0, sizeof(double) * (na ? na : 1)));
    CUDA_TRY(cudaMemset(c->dihedrals, 0, sizeof(double) * (nd ? nd : 1)));
    return 0;
}

extern "C" int sepgpu_get_bonded_values(sepgpu_ctx *c, double *blengths, double *angles, double *dihedrals)
{
    if (!c) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("get_bonded_values: per-term values stay on the rank that owns the term's first atom in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (blengths && c->nb) CUDA_TRY(cudaMemcpy(blengths, c->blengths, sizeof(double) * c->nb, cudaMemcpyDeviceToHost));
    if (angles && c->na) CUDA_TRY(cudaMemcpy(angles, c->angles, sizeof(double) * c->na, cudaMemcpyDeviceToHost));
    if (dihedrals && c->nd) CUDA_TRY(cudaMemcpy(dihedrals, c->dihedrals, sizeof(double) * c->nd, cudaMemcpyDeviceToHost));
    return 0;
}

int sepgpu_finalize_force(sepgpu_ctx *c, int nrows, double scale, int flags);

// what the bonded kernels need to know about a decomposed run; also brings the halo coordinates of this step into xs
static int bonded_view(sepgpu_ctx *c, const sepgpu_sys *sys, BondedDD *D, int *nrows)
{
    memset(D, 0, sizeof *D);
    *nrows = c->n;
    if (!c->dd) return 0;
    int rc = sepgpu_dd_halo_update(c, sys);
    if (rc) return rc;
    const int *gmap = NULL;
    if ((rc = sepgpu_dd_gmap(c, &gmap))) return rc;
    D->gid = c->gid; D->gmap = gmap; D->rank = c->rank; D->xs = c->xs; D->n_own = c->n_own; D->on = 1;
    *nrows = c->n_own;
    return 0;
}

static BoxB make_box(const sepgpu_sys *sys)
{
    BoxB B; B.Lx = sys->length[0]; B.Ly = sys->length[1]; B.Lz = sys->length[2];
    return B;
}

extern "C" int sepgpu_stretch_harmonic(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double lbond, double ks)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    if (!c->atom_bond_ptr) { sepgpu_set_error("stretch_harmonic: no topology on the device"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    BondedDD D; int nrows;
    { const int rcv = bonded_view(c, sys, &D, &nrows); if (rcv) return rcv; }
    const int grid = (nrows + BONDED_BLOCK - 1) / BONDED_BLOCK;
    ktimer_begin(c, &c->t_bonded);
    k_bond<<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f4, nrows, c->blist, c->atom_bond_ptr, c->atom_bond_idx,
                                                 type, lbond, ks, make_box(sys), c->f_zero ? 1 : 0, c->blengths, c->partial, D, c->scal);
    ktimer_end(c, &c->t_bonded);
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, grid, 1.0, 2);      // epot +=, pot_P += and pot_P_bond += (source/sepmol.c:403-409)
}

static int run_angle(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double angle0, double k, bool cossq)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    if (!c->atom_angle_ptr) { sepgpu_set_error("angle force: no topology on the device"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    BondedDD D; int nrows;
    { const int rcv = bonded_view(c, sys, &D, &nrows); if (rcv) return rcv; }
    const int grid = (nrows + BONDED_BLOCK - 1) / BONDED_BLOCK;
    const double cCon = cos(SEPGPU_PI - angle0);        // host libm, as the reference (source/sepmol.c:422)
    ktimer_begin(c, &c->t_bonded);
    if (cossq)
        k_angle<true><<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f4, nrows, c->alist, c->atom_angle_ptr, c->atom_angle_idx,
                                                            type, angle0, k, cCon, make_box(sys), c->f_zero ? 1 : 0, c->angles, c->partial, D, c->scal);
    else
        k_angle<false><<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f4, nrows, c->alist, c->atom_angle_ptr, c->atom_angle_idx,
                                                             type, angle0, k, cCon, make_box(sys), c->f_zero ? 1 : 0, c->angles, c->partial, D, c->scal);
    ktimer_end(c, &c->t_bonded);
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, grid, 1.0, 0);      // no virial from angles (partial rows carry zeros)
}

extern "C" int sepgpu_angle_harmonic(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double angle0, double k)
{
    return run_angle(c, sys, type, angle0, k, false);
}

extern "C" int sepgpu_angle_cossq(sepgpu_ctx *c, const sepgpu_sys *sys, int type, double angle0, double k)
{
    return run_angle(c, sys, type, angle0, k, true);
}

extern "C" int sepgpu_torsion_ryckaert(sepgpu_ctx *c, const sepgpu_sys *sys, int type, const double g[6])
{
    if (!c || !sys || !g) return SEPGPU_EINVAL;
    if (!c->atom_dihed_ptr) { sepgpu_set_error("torsion force: no topology on the device"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    BondedDD D; int nrows;
    { const int rcv = bonded_view(c, sys, &D, &nrows); if (rcv) return rcv; }
    const int grid = (nrows + BONDED_BLOCK - 1) / BONDED_BLOCK;
    RBCoef G; for (int k = 0; k < 6; k++) G.g[k] = g[k];
    ktimer_begin(c, &c->t_bonded);
    k_torsion<<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f4, nrows, c->dlist, c->atom_dihed_ptr, c->atom_dihed_idx,
                                                    type, G, make_box(sys), c->f_zero ? 1 : 0, c->dihedrals, c->partial, D, c->scal);
    ktimer_end(c, &c->t_bonded);
    KERNEL_CHECK();
    c->f_zero = false;
    return sepgpu_finalize_force(c, grid, 1.0, 0);
}

// ---- one bonded term kind into an array of its own ------------------------------------------------------------
// The reference's "OpenMP model II" helpers (source/sepomp.c:179-329: sep_omp_bond, sep_omp_angle, sep_omp_torsion) add the
// forces of one term kind to a matrix of the caller's and touch nothing else -- no sepret sums, no atoms[].f.  Same kernels
// as the entries above, pointed at a spare force array and a spare block-sum array; the simulation state does not change.
// kind 0: harmonic bonds (par = {lbond, ks})   1: cos^2 angles (par = {angle0, k}, :218-262)   2: Ryckaert-Bellemans (par = g[6])
// out3[3 * i + k] = force component k of atom i from those terms.
extern "C" int sepgpu_bonded_side(sepgpu_ctx *c, const sepgpu_sys *sys, int kind, int type, const double *par, double *out3)
{
    if (!c || !sys || !par || !out3 || kind < 0 || kind > 2) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("bonded_side: single-domain contexts only"); return SEPGPU_ESTATE; }
    if (!(kind == 0 ? c->atom_bond_ptr : kind == 1 ? c->atom_angle_ptr : c->atom_dihed_ptr)) {
        sepgpu_set_error("bonded_side: no topology on the device");
        return SEPGPU_ESTATE;
    }
    SEPGPU_ENTER(c);
    SEPGPU_BENIGN(c);
    const int n = c->n;
    if (n <= 0) return 0;
    if (!c->f_side) CUDA_TRY(cudaMalloc((void **)&c->f_side, sizeof(d4) * (size_t)c->ncap));
    if (!c->partial_side) CUDA_TRY(cudaMalloc((void **)&c->partial_side, sizeof(double) * (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * 16));
    BondedDD D;
    memset(&D, 0, sizeof D);
    const int grid = (n + BONDED_BLOCK - 1) / BONDED_BLOCK;
    if (grid > SEPGPU_MAX_BLOCKS_PARTIAL) { sepgpu_set_error("bonded_side: system too large"); return SEPGPU_EINVAL; }
    if (kind == 0) {
        k_bond<<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f_side, n, c->blist, c->atom_bond_ptr, c->atom_bond_idx, type, par[0], par[1],
                                                     make_box(sys), 1, c->blengths, c->partial_side, D, c->scal);
    } else if (kind == 1) {
        const double cCon = cos(SEPGPU_PI - par[0]);
        k_angle<true><<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f_side, n, c->alist, c->atom_angle_ptr, c->atom_angle_idx, type, par[0],
                                                            par[1], cCon, make_box(sys), 1, c->angles, c->partial_side, D, c->scal);
    } else {
        RBCoef G; for (int k = 0; k < 6; k++) G.g[k] = par[k];
        k_torsion<<<grid, BONDED_BLOCK, 0, c->stream>>>(c->x4, c->f_side, n, c->dlist, c->atom_dihed_ptr, c->atom_dihed_idx, type, G,
                                                        make_box(sys), 1, c->dihedrals, c->partial_side, D, c->scal);
    }
    KERNEL_CHECK();
    const size_t bytes = sizeof(d4) * (size_t)n;
    int rc = sepgpu_ensure_stage(c, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));                // an earlier asynchronous copy may still read the staging buffer
    CUDA_TRY(cudaMemcpyAsync(c->stage, c->f_side, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const d4 *h = (const d4 *)c->stage;
    for (int i = 0; i < n; i++) { out3[3 * i] = h[i].x; out3[3 * i + 1] = h[i].y; out3[3 * i + 2] = h[i].z; }
    return 0;
}
