// sepgpu_extra.cu -- the callers either side of the hot path:
//   box-changing routines  sep_compress_box*, sep_berendsen*        (reference source/sepmisc.c:892-1083)
//   per-type temperature relaxation  sep_relax_temp                 (source/sepmisc.c:357-390)
//   tethering springs  sep_force_x0 with sep_spring_x0              (source/sepmisc.c:167-181, 645-670)
// All are one streaming pass over the per-atom records (HBM-bound, 64-96 B per atom); sums go through the same
// deterministic block -> partial row -> fixed-order reduction as the integrator.
#include "sepgpu_internal.cuh"

#include <math.h>

#define XB 256
#define X_MAX_GRID (148 * 8)

// ---- box scaling ----------------------------------------------------------------------------------------------
// x <- x * s per direction.  The cell-sorted copy holds x + (crossings since the list was built) * L; the
// reference scales positions and box lengths by DIFFERENT factors in the barostat routines (length by xi,
// positions by xi^(1/3), source/sepmisc.c:897-901), so the sorted copy is rebuilt from x4 and the NEW box
// lengths instead of being scaled.
__global__ void k_scale_box(d4 *__restrict__ x4, const i4 *__restrict__ cr4, const int *__restrict__ rank, d4 *__restrict__ xs,
                            int n, double sx, double sy, double sz, double Lx, double Ly, double Lz, int write_xs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 x = x4[i];
    x.x *= sx; x.y *= sy; x.z *= sz;
    x4[i] = x;
    if (write_xs) {
        const int w = cr4[i].w;                       // packed crossings since the list build (sepgpu_intgr.cu)
        if (w != 0) {
            x.x += ((w & 1023) - 512) * Lx; x.y += (((w >> 10) & 1023) - 512) * Ly; x.z += (((w >> 20) & 1023) - 512) * Lz;
        }
        xs[rank[i]] = x;
    }
}

extern "C" int sepgpu_scale_box(sepgpu_ctx *c, const double scale[3], const double new_length[3])
{
    if (!c || !scale || !new_length) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("scale_box: not available in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    const int write_xs = (c->list_valid && !c->sorted_identity) ? 1 : 0;
    k_scale_box<<<(c->n_own + XB - 1) / XB, XB, 0, c->stream>>>(c->x4, c->cr4, c->rank, c->xs, c->n_own, scale[0], scale[1], scale[2],
                                                               new_length[0], new_length[1], new_length[2], write_xs);
    KERNEL_CHECK();
    if (!write_xs) c->xs_current = false;
    return 0;
}

// ---- sep_relax_temp ---------------------------------------------------------------------------------------------
// Ta = 2 ekin / (3 ntype) of the atoms of `type`;  v *= sqrt(1 + dt/tau (Td/Ta - 1));  then the momentum of
// that type is removed (sep_reset_momentum, source/sepmisc.c:1173-1192).
__global__ void __launch_bounds__(XB)
k_type_mv2(const d4 *__restrict__ v4, const d4 *__restrict__ x4, int n, int type, double *__restrict__ partial)
{
    __shared__ double red[2 * (XB / 32)];
    double acc[2] = {0.0, 0.0};
    for (int i = blockIdx.x * XB + threadIdx.x; i < n; i += gridDim.x * XB) {
        if (tag_type(x4[i].w) != type) continue;
        const d4 v = v4[i];
        // ekin += m * v_k^2 component by component (source/sepmisc.c:366-367)
        acc[0] += v.w * (v.x * v.x) + v.w * (v.y * v.y) + v.w * (v.z * v.z);
        acc[1] += 1.0;
    }
    block_sum<2, XB>(acc, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = acc[0]; partial[2 * blockIdx.x + 1] = acc[1]; }
}

__global__ void __launch_bounds__(256)
k_relax_factor(const double *__restrict__ partial, int nrows, double Td, double dt_over_tau, double *out)
{
    __shared__ double red[2 * 8];
    double v[2] = {0.0, 0.0};
    for (int r = threadIdx.x; r < nrows; r += 256) { v[0] += partial[2 * r]; v[1] += partial[2 * r + 1]; }
    block_sum<2, 256>(v, red);
    if (threadIdx.x == 0) {
        const double ekin = 0.5 * v[0];
        const double Ta = 2.0 * ekin / (3.0 * v[1]);
        out[0] = sqrt(1.0 + dt_over_tau * (Td / Ta - 1.0));
        out[1] = ekin;
    }
}

__global__ void k_scale_v_type(d4 *__restrict__ v4, const d4 *__restrict__ x4, int n, int type, const double *__restrict__ fact)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || tag_type(x4[i].w) != type) return;
    const double s = fact[0];
    d4 v = v4[i];
    v.x *= s; v.y *= s; v.z *= s;
    v4[i] = v;
}

extern "C" int sepgpu_relax_temp(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double Td, double tau, double *ekin_type)
{
    if (!c || !sys || tau == 0.0) return SEPGPU_EINVAL;
    if (c->dd) { sepgpu_set_error("relax_temp: not available in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    long long want = ((long long)c->n_own + XB - 1) / XB;
    const int nrows = (int)(want < X_MAX_GRID ? want : X_MAX_GRID);
    const int t = (unsigned char)type;
    int rc;
    if ((rc = sepgpu_apply_pending(c))) return rc;        // a deferred thermostat term must see the velocities it was computed for
    double *fact = c->partial + (size_t)SEPGPU_MAX_BLOCKS_PARTIAL * 16 - 8;      // two doubles at the end of the scratch rows
    k_type_mv2<<<nrows, XB, 0, c->stream>>>(c->v4, c->x4, c->n_own, t, c->partial);
    k_relax_factor<<<1, 256, 0, c->stream>>>(c->partial, nrows, Td, sys->dt / tau, fact);
    k_scale_v_type<<<(c->n_own + XB - 1) / XB, XB, 0, c->stream>>>(c->v4, c->x4, c->n_own, t, fact);
    KERNEL_CHECK();
    c->mv2_valid = false;
    if (ekin_type) {
        if ((rc = sepgpu_ensure_stage(c, 64))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->stage, fact + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        *ekin_type = *(double *)c->stage;
    }
    return sepgpu_reset_momentum(c, type);
}

// ---- sep_force_x0 with sep_spring_x0 ------------------------------------------------------------------------------
// r = wrap(x0 - x);  f -= ft r  with ft = -k  (source/sepmisc.c:645-668; the routine's energy is computed and
// dropped by the reference, :665, so nothing is added to epot here either)
__global__ void k_force_x0(const d4 *__restrict__ x4, const d4 *__restrict__ x0, d4 *__restrict__ f4, int n, int type,
                           double ft, double Lx, double Ly, double Lz, int f_zero)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d4 x = x4[i];
    const bool mine = tag_type(x.w) == type;
    if (!mine && !f_zero) return;
    d4 f;
    if (f_zero) { f.x = f.y = f.z = f.w = 0.0; } else f = f4[i];
    if (mine) {
        const d4 a = x0[i];
        const double rx = wrap_exact(a.x - x.x, Lx, 0.5 * Lx);
        const double ry = wrap_exact(a.y - x.y, Ly, 0.5 * Ly);
        const double rz = wrap_exact(a.z - x.z, Lz, 0.5 * Lz);
        f.x -= ft * rx; f.y -= ft * ry; f.z -= ft * rz;
    }
    f4[i] = f;
}

extern "C" int sepgpu_force_x0(sepgpu_ctx *c, const sepgpu_sys *sys, char type, double kspring)
{
    if (!c || !sys) return SEPGPU_EINVAL;
    if (!c->x0) { sepgpu_set_error("force_x0: no tether positions on the device (put SEPGPU_F_X0 first)"); return SEPGPU_ESTATE; }
    if (c->dd) { sepgpu_set_error("force_x0: not available in decomposed runs"); return SEPGPU_ESTATE; }
    SEPGPU_ENTER(c);
    k_force_x0<<<(c->n_own + XB - 1) / XB, XB, 0, c->stream>>>(c->x4, c->x0, c->f4, c->n_own, (unsigned char)type, -kspring,
                                                              sys->length[0], sys->length[1], sys->length[2], c->f_zero ? 1 : 0);
    KERNEL_CHECK();
    c->f_zero = false;
    return 0;
}
