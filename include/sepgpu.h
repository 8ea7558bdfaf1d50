/* sepgpu.h -- C ABI of the B200 (sm_100a) device layer of seplib-b200.
 *
 * Plain C: opaque handle, plain pointers and sizes, no C++/torch types.  The host layer
 * (seplib_b200/csrc/host, the sep_* API of include/sep.h) is the only in-tree caller; tests and
 * bench.py bind the same entry points through ctypes.  Every function returns 0 on success or a
 * negative SEPGPU_E* code; sepgpu_last_error() gives the text.  There is NO CPU fallback: every
 * entry point fails with SEPGPU_ENODEV when no CUDA device is usable.
 *
 * Each entry point names the reference interface it stands in for (paths relative to the
 * reference root).
 */
#ifndef SEPGPU_H
#define SEPGPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sepgpu_ctx sepgpu_ctx;

/* error codes */
#define SEPGPU_OK          0
#define SEPGPU_ENODEV     -1   /* no CUDA device / driver */
#define SEPGPU_ECUDA      -2   /* a CUDA runtime call failed */
#define SEPGPU_EINVAL     -3   /* bad argument */
#define SEPGPU_ECELL      -4   /* atom outside the cell grid ("Index larger than array length", source/sepprfrc.c:408) */
#define SEPGPU_ENEIGHB    -5   /* "Too many neighbours" (source/sepprfrc.c:499): a half list reached SEP_NEIGHB */
#define SEPGPU_ESTATE     -6   /* call order problem (e.g. list force without a list) */
#define SEPGPU_ENCCL      -7   /* NCCL failure (domain decomposition) */
#define SEPGPU_ETABLE     -8   /* a pair came closer than a tabulated pair function reaches */

/* per-atom fields for sepgpu_put / sepgpu_get.  Host element = what the reference keeps in seppart
 * (include/sepstrct.h:23-61): 3 doubles for vectors, 1 double for scalars, 3 ints for counters. */
enum {
    SEPGPU_F_X = 0,          /* x[3]   wrapped position              */
    SEPGPU_F_V,              /* v[3]                                  */
    SEPGPU_F_F,              /* f[3]                                  */
    SEPGPU_F_M,              /* m                                     */
    SEPGPU_F_Z,              /* z      point charge                   */
    SEPGPU_F_TYPE,           /* type   (char)                         */
    SEPGPU_F_MOLINDEX,       /* molindex (int)                        */
    SEPGPU_F_XN,             /* xn[3]  position at last list build    */
    SEPGPU_F_CROSS_NEIGHB,   /* cross_neighb[3] (int)                 */
    SEPGPU_F_CROSSINGS,      /* crossings[3]    (int)                 */
    SEPGPU_F_PV,             /* pv[3]  predicted velocity (DPD)       */
    SEPGPU_F_PA,             /* pa[3]  previous acceleration (DPD)    */
    SEPGPU_F_A,              /* a[3]   acceleration (get only; = f/m) */
    SEPGPU_F_BOND,           /* bond[SEP_BOND]   partner table (int)  */
    SEPGPU_F_ANGLE,          /* angle[SEP_ANGLE] partner table (int)  */
    SEPGPU_F_DIHED,          /* dihed[SEP_DIHED] partner table (int)  */
    SEPGPU_F_GID,            /* global atom id (int), decomposed runs */
    SEPGPU_F_X0,             /* x0[3]  tether position (sep_set_x0)   */
    SEPGPU_F_COUNT
};

/* exclusion rule of the list builders (include/sepdef.h:29-36) */
#define SEPGPU_ALL            1
#define SEPGPU_EXCL_BONDED    2
#define SEPGPU_EXCL_SAME_MOL  3

/* pair-function family of sep_force_pairs / sep_force_lj (source/sepmisc.c:115-164,
 * source/sepprfrc.c:782-795).  All are u = 4 eps[(s/r)^12 - aw (s/r)^6] - shift. */
typedef struct {
    double cf;        /* interaction cutoff                         */
    double eps;       /* epsilon                                    */
    double sigma;     /* sigma                                      */
    double aw;        /* weight of the attractive term              */
    double shift;     /* subtracted from u for every in-range pair  */
} sepgpu_ljparam;

/* run parameters the host passes on every call (mirror of the sepsys fields the hot path reads,
 * include/sepstrct.h:104-135) */
typedef struct {
    double length[3];
    double lsubbox[3];
    int    nsubbox[3];
    double cf;               /* maximum cutoff of the system  */
    double skin;
    double dt;
    int    neighb_update;    /* SEP_BRUTE=0 / SEP_NEIGHBLIST=1 / SEP_LLIST_NEIGHBLIST=2 */
} sepgpu_sys;

/* scalar results, device-accumulated with the reference's assign/accumulate rules and copied
 * out by sepgpu_read_scalars (mirror of the sepret/sepsys fields the hot path writes) */
typedef struct {
    double epot, ecoul, ekin;
    double pot_P[9], kin_P[9], pot_P_bond[9];
    double max_dist2;          /* sys->max_dist2                                   */
    double sum_mv2;            /* sum m v^2 of the velocities now on the device     */
    double alpha[4];           /* device copies of thermostat multipliers (slots)  */
    int    neighb_flag;        /* 1 when the skin trigger fired in the last integrator call */
    int    nbuild;             /* list builds executed so far                       */
    int    error;              /* sticky device error (SEPGPU_E*) or 0              */
    int    max_neighb;         /* longest full list seen at the last build          */
    long long npairs_listed;   /* ordered (i->j) entries in the current list        */
} sepgpu_scalars;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* device state for npart atoms; replaces the allocation side of sep_init (source/sepinit.c:15-56) */
int  sepgpu_create(sepgpu_ctx **out, size_t npart, int device);
void sepgpu_destroy(sepgpu_ctx *ctx);
const char *sepgpu_last_error(void);
int  sepgpu_device_count(void);

/* ---- bulk state movement (host <-> HBM) ------------------------------------------------------- */
/* strided host access so the host layer can point straight into its seppart AoS array
 * (stride = sizeof(seppart)) while tests pass packed numpy arrays (stride = element size). */
int sepgpu_put(sepgpu_ctx *ctx, int field, const void *host, size_t stride_bytes);
int sepgpu_get(sepgpu_ctx *ctx, int field, void *host, size_t stride_bytes);
/* several fields of one record array in a single pass over host memory (parallel host threads) and a
 * single PCIe transfer: field f of atom i lives at base + i*stride_bytes + offsets[f] */
int sepgpu_put_fields(sepgpu_ctx *ctx, const void *base, size_t stride_bytes, int nfields,
                      const int *fields, const size_t *offsets);
int sepgpu_get_fields(sepgpu_ctx *ctx, void *base, size_t stride_bytes, int nfields,
                      const int *fields, const size_t *offsets);
/* topology lists in the reference layout (include/sepstrct.h:77-87): blist[3n], alist[4n], dlist[5n] */
int sepgpu_set_topology(sepgpu_ctx *ctx, const unsigned *blist, unsigned nbonds,
                        const unsigned *alist, unsigned nangles,
                        const unsigned *dlist, unsigned ndihedrals);
int sepgpu_get_bonded_values(sepgpu_ctx *ctx, double *blengths, double *angles, double *dihedrals);

/* Convenience: nsteps of the Lennard-Jones NVT loop of the reference's prg1 (prgs/prg1.c:52-70), driven from C -- per step
 * exactly sepgpu_reset_ret, sepgpu_reset_force, sepgpu_force_lj(epot_assign = 1; rebuilds the list when the trigger fired),
 * sepgpu_nosehoover, sepgpu_leapfrog.  Same results as making those calls one by one; no interpreter between them when the
 * caller is not C.  Collective in decomposed runs.  Stops at the first error. */
int sepgpu_md_lj_nvt(sepgpu_ctx *ctx, const sepgpu_sys *sys, const char types[2], const sepgpu_ljparam *p, unsigned opt,
                     double temp, int alpha_slot, double tau, int nsteps);

/* ---- per-step hot path -------------------------------------------------------------------------- */
/* sep_reset_retval (source/sepret.c:19-47) */
int sepgpu_reset_ret(sepgpu_ctx *ctx);
/* sep_reset_force (source/sepmisc.c:393-400): f <- 0 and max_dist2 <- 0 */
int sepgpu_reset_force(sepgpu_ctx *ctx);
/* sep_neighb / sep_neighb_nonbonded / sep_neighb_excl_same_mol (source/sepprfrc.c:347-378):
 * cell binning + Verlet list.  The pair SET equals the reference's bit for bit. */
int sepgpu_neighb_build(sepgpu_ctx *ctx, const sepgpu_sys *sys, unsigned opt);
/* sep_force_pairs list/brute branch and sep_force_lj (source/sepprfrc.c:226-274, 743-780).
 * epot_assign=1 reproduces "retval->epot = epot" (:222), 0 accumulates (:64, :922). */
int sepgpu_force_lj(sepgpu_ctx *ctx, const sepgpu_sys *sys, const char types[2],
                    const sepgpu_ljparam *p, unsigned opt, int epot_assign);
/* sep_force_pairs with a pair function of the caller's own (include/sepprfrc.h:49-51: double (*fun)(double r2, char opt)):
 * the host layer samples fun(r2,'f') and fun(r2,'u') on n uniform points of r^2 in [r2_lo, cf^2]; tab_fu holds the n pairs
 * {f, u}.  The device interpolates with a cubic through the four surrounding samples.  Tile lists and SEP_BRUTE. */
int sepgpu_force_table(sepgpu_ctx *ctx, const sepgpu_sys *sys, const char types[2], double cf, const double *tab_fu, int n,
                       double r2_lo, unsigned opt, int epot_assign);
/* sep_coulomb_sf (source/sepcoulomb.c:5-18) */
int sepgpu_coulomb_sf(sepgpu_ctx *ctx, const sepgpu_sys *sys, double cf, unsigned opt);
/* sep_force_dpd (source/sepprfrc.c:278-301); counter-based pair noise keyed on (seed, step).
 * seed == SEPGPU_DPD_SEED_FIXED: every pair draws u = 0.75 -- what the reference computes when its rand() is interposed to
 * return 3*2^29 (sep_rand() = rand()/(RAND_MAX+1), include/sepmisc.h:60); used to pin the dissipative and random terms to
 * the reference itself (tests/golden/dpd_force_n512.npz). */
#define SEPGPU_DPD_SEED_FIXED 0xFFFFFFFFFFFFFFFFULL
int sepgpu_force_dpd(sepgpu_ctx *ctx, const sepgpu_sys *sys, const char types[2], double cf,
                     double aij, double temp, double sigma, unsigned opt,
                     unsigned long long seed, unsigned long long step);
/* sep_stretch_harmonic / sep_angle_harmonic / sep_angle_cossq / sep_torsion_Ryckaert
 * (source/sepmol.c:372-414, 469-516, 418-467, 520-587) */
int sepgpu_stretch_harmonic(sepgpu_ctx *ctx, const sepgpu_sys *sys, int type, double lbond, double ks);
int sepgpu_angle_harmonic(sepgpu_ctx *ctx, const sepgpu_sys *sys, int type, double angle0, double k);
int sepgpu_angle_cossq(sepgpu_ctx *ctx, const sepgpu_sys *sys, int type, double angle0, double k);
int sepgpu_torsion_ryckaert(sepgpu_ctx *ctx, const sepgpu_sys *sys, int type, const double g[6]);
/* The forces of ONE bonded term kind in an array of their own, nothing else touched (no sums, no change of the atoms' force
 * array): what the reference's OpenMP "model II" helpers compute -- sep_omp_bond (source/sepomp.c:179-213, kind 0, par =
 * {lbond, ks}), sep_omp_angle (:215-262, kind 1, the cos^2 form, par = {angle0, k}), sep_omp_torsion (:265-329, kind 2,
 * par = g[6]).  out3[3 i + k] is written for every atom.  Single-domain contexts. */
int sepgpu_bonded_side(sepgpu_ctx *ctx, const sepgpu_sys *sys, int kind, int type, const double *par, double *out3);
/* sep_nosehoover (source/sepintgr.c:149-168).  alpha lives in device slot `slot` (0..3); the
 * f -= alpha m v update is fused into the next integrator kernel. */
int sepgpu_nosehoover(sepgpu_ctx *ctx, const sepgpu_sys *sys, double temp0, int slot, double tau);
/* _sep_nosehoover_type (source/sepintgr.c:170-198): alpha3 is the caller's 3-slot history (in/out) */
int sepgpu_nosehoover_type(sepgpu_ctx *ctx, const sepgpu_sys *sys, char type, double Td,
                           double alpha3[3], double Q);
int sepgpu_set_alpha(sepgpu_ctx *ctx, int slot, double alpha);
/* sep_leapfrog (+ sep_periodic, skin trigger, sep_set_xn) (source/sepintgr.c:18-88) */
int sepgpu_leapfrog(sepgpu_ctx *ctx, const sepgpu_sys *sys);
/* sep_verlet_dpd (source/sepintgr.c:296-345) */
int sepgpu_verlet_dpd(sepgpu_ctx *ctx, const sepgpu_sys *sys, double lambda, int stepnow);
/* sep_reset_momentum (source/sepmisc.c:1173-1192) and the x-rescale of sep_compress_box (:1009-1010) */
int sepgpu_reset_momentum(sepgpu_ctx *ctx, char type);
int sepgpu_scale_positions(sepgpu_ctx *ctx, double xi);

/* molecule-molecule force table for the molecular pressure tensor: sepmolinfo.Fij (include/sepstrct.h:94),
 * filled by the pair routines (source/sepprfrc.c:199-207, source/sepcoulomb.c:138-147), cleared by
 * sep_reset_force_mol (source/sepmisc.c:403-430).  fij_get returns nmol*nmol*3 floats, [i][j][k]. */
int sepgpu_fij_enable(sepgpu_ctx *ctx, int nmol);
int sepgpu_fij_reset(sepgpu_ctx *ctx);
int sepgpu_fij_get(sepgpu_ctx *ctx, float *out);

/* ---- callers either side of the hot path -------------------------------------- */
/* x <- x * scale[k]; the box becomes new_length.  Device side of sep_compress_box, sep_compress_box_dir,
 * sep_compress_box_dir_length (source/sepmisc.c:994-1083) and sep_berendsen, sep_berendsen_iso (:892-944); the
 * host layer updates sys->length / nsubbox / lsubbox / volume as those routines do. */
int sepgpu_scale_box(sepgpu_ctx *ctx, const double scale[3], const double new_length[3]);
/* sep_relax_temp (source/sepmisc.c:357-390): rescale the velocities of one type towards Td, then remove that
 * type's momentum.  ekin_type (may be NULL) receives the type's kinetic energy before the rescale. */
int sepgpu_relax_temp(sepgpu_ctx *ctx, const sepgpu_sys *sys, char type, double Td, double tau, double *ekin_type);
/* sep_force_x0 with sep_spring_x0 (source/sepmisc.c:167-181, 645-670): harmonic tether of one type to SEPGPU_F_X0 */
int sepgpu_force_x0(sepgpu_ctx *ctx, const sepgpu_sys *sys, char type, double kspring);

/* sep_fp (source/sepintgr.c:235-293) and sep_langevinGJF (:89-146).  noise4: HOST array, four doubles per atom
 * {g0, g1, g2, ldiff}: the Gaussian numbers of this step in the reference's drawing order (atom-major, component-
 * minor) and the atom's seppart.ldiff (used by sep_fp only). */
int sepgpu_fp(sepgpu_ctx *ctx, const sepgpu_sys *sys, double temp, const double *noise4);
int sepgpu_langevin_gjf(sepgpu_ctx *ctx, const sepgpu_sys *sys, double temp, double alpha, const double *noise4);

/* ---- sampler feeds: what the reference's run-time samplers consume, reduced on the device.
 * sep_sample (source/sepsampler.c:177-240) hands the host atoms[] to each sampler; behind the sep_* API that costs a download
 * of the whole array per sample (per STEP for "msd", which follows the atoms across the boundaries itself, :537-552).  The
 * calls below return the sums those samplers form instead.  Single-domain contexts only (SEPGPU_ESTATE otherwise); all are
 * stream-synchronising reads and change no simulation state.  The host layer uses them when SEP_SAMPLER_FEEDS=1. */
/* "vacf" (:658-722): stores v_x of every atom as the next row of a device-resident block of lvec rows; when the block is
 * full, *completed = 1 and acf_block[t] = sum over atoms and time origins t0 of v_x(t0) v_x(t0 + t), t < lvec <= 768. */
int sepgpu_feed_vacf(sepgpu_ctx *ctx, int lvec, double *acf_block, int *completed);
/* "msd" (:555-655): sums[0] = sum |dr|^2, sums[1] = sum |dr|^4, sums[2] = number of atoms, over the atoms of `type`;
 * fs[i] = sum cos(k[i] dx).  dr is the displacement since the last call with new_origin != 0, unwrapped with the device's own
 * crossing counters (seppart.crossings) and the box lengths given. */
int sepgpu_feed_msd(sepgpu_ctx *ctx, int new_origin, char type, const double length[3], int nk, const double *k,
                    double *sums, double *fs);
/* "profs" (:1416-1525): nbins slabs along z over [0, lz); out4 = {sum m v_x}[nbins], {sum m}[nbins], {sum m (v_y^2 + v_z^2)}[nbins],
 * {atoms}[nbins] for the atoms of `type` (FP64 atomics: sums are order-dependent in the last bits).  nbins <= 1024. */
int sepgpu_feed_profile(sepgpu_ctx *ctx, char type, double lz, int nbins, double *out4);
/* "gh" (:926-1107): for each wave vector (0, k[n], 0), with e = exp(i k y_true), out16[16 n + ...] = Re, Im of
 * sum e, sum m e, sum m v_x e, sum m v_y e, sum (m v^2 / 2) e, sum m a_y e, sum m v_y^2 e; then [14] = sum m v^2 / 2, [15] = 0. */
int sepgpu_feed_fourier(sepgpu_ctx *ctx, double ly, int nwave, const double *k, double *out16);
/* "radial" (:361-468): counts[lvec][ncomb] of this configuration, bin width lbox / (2 lvec), every pair i < j once, type
 * combinations (a <= b) in the reference's order; integer counts, identical to the host loop.  All pairs: O(N^2). */
int sepgpu_feed_radial(sepgpu_ctx *ctx, double lbox, int lvec, int ntypes, const char *types, long long *counts);

/* ---- results -------------------------------------------------------------------------------------- */
/* stream-synchronising read of the scalar block */
int sepgpu_read_scalars(sepgpu_ctx *ctx, sepgpu_scalars *out);
int sepgpu_sync(sepgpu_ctx *ctx);
/* current Verlet list as unordered pairs (i<j, original atom indices), for parity checks against
 * ptr[i].neighb (source/sepprfrc.c:493-497).  Returns the pair count or a negative error. */
long long sepgpu_get_pairs(sepgpu_ctx *ctx, int *pairs, long long max_pairs);
/* force the next force call to rebuild (sys->neighb_flag = 1) */
int sepgpu_request_rebuild(sepgpu_ctx *ctx);
/* tuning: lanes cooperating on one atom in the list force kernels (1,2,4,8,16,32; 0 = default) */
int sepgpu_set_option(sepgpu_ctx *ctx, const char *name, long long value);
/* Options (also through the environment, SEPGPU_OPTS="name=value,..." read by sepgpu_create):
 *   tpa, prefilter, overlap, force_grid, time_kernels, neighb_cap      tuning / measurement
 *   tile_list      = 1 | 0     rows of 16-bit tile slots + the shared-memory tile force kernel (default) | global-index rows
 *                              and the gather kernel (contexts with charges, DPD state, the molecule-pair table and small
 *                              grids use global-index rows in any case)
 *   build_window   = 1 | 0 | 2 per-lane x-window in the list builder: by cell occupancy (default) | never | always
 *   coulomb_kernel = 2 | 1     list Coulomb kernel: charge-in-record, branch-free (default) | first version
 *   typed_sublist  = 1 | 0     typed Lennard-Jones calls on global-index rows walk a per-type sub-list (default on)
 *   step_fold      = 1 | 0     the step's last force reduction and the Nose-Hoover multiplier update are folded into the
 *                              integrator's kernels (3 launches per Lennard-Jones step instead of 5; default on)
 *   fin_multi      = 1 | 0     final reductions on several CTAs with a last-block pass (default on)
 *   spec_force     = 1 | 0     the step's first force call is queued for the next step behind the integrator, guarded by the
 *                              device-side rebuild flag, and adopted by the next sepgpu_force_lj (default on; single GPU)
 * sepgpu_get_option also answers "list_f16", "tile_R", "tile_stage", "max_half", "dd_p2p", "spec_adopted". */
int sepgpu_get_option(sepgpu_ctx *ctx, const char *name, long long *value);

/* ---- spatial domain decomposition over the GPUs of one box (no reference counterpart: the reference
 * is single-address-space OpenMP).  One process per GPU; slabs of whole cell layers along z; halo
 * coordinates every step and migration at list rebuilds travel over NVLink with NCCL send/recv; one
 * small all-reduce per step carries sum m v^2 (sep_nosehoover) and the maximum displacement (skin
 * trigger).  Usage: create the context with capacity ncap >= owned + halo atoms, sepgpu_dd_init on
 * every rank with the same id, sepgpu_dd_set_owned(n), sepgpu_put(X, V, GID, ...) for the atoms whose
 * cell layer (int)(z/lsubbox[2]) lies in this rank's range, then the ordinary per-step calls.  All
 * ranks must make the same calls in the same order (they are collective). ---- */
/* put/get address the host records through rows[i] (i = device atom) instead of i; NULL switches back.  Decomposed runs
 * behind the sep_* API use it with the global ids so that each process touches only its own atoms of the full array. */
int sepgpu_set_host_rows(sepgpu_ctx *ctx, const int *rows);
int sepgpu_dd_unique_id(void *out128);
/* decomposed sep_coulomb_sf: the charges of ALL atoms by global id, on every rank (after sepgpu_dd_init) */
int sepgpu_dd_set_charges(sepgpu_ctx *ctx, const double *z_global);
int sepgpu_dd_init(sepgpu_ctx *ctx, int rank, int nranks, const void *id128, const sepgpu_sys *sys, long long n_global);
int sepgpu_dd_set_owned(sepgpu_ctx *ctx, int n_own);
int sepgpu_dd_layers(sepgpu_ctx *ctx, int *z0, int *z1, int *n_own, int *n_halo);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------------- */
/* CUDA-event timing on the context's stream */
int sepgpu_timer_start(sepgpu_ctx *ctx);
int sepgpu_timer_stop(sepgpu_ctx *ctx, float *ms);
/* device time spent in the list-force kernel since the last call (events around each launch) */
int sepgpu_kernel_time(sepgpu_ctx *ctx, const char *which, float *ms_total, int *launches);
/* FP64 FMA-chain and copy microbenchmarks: the box's own FP64 / HBM ceilings */
int sepgpu_peak_fp64(int device, double *tflops);
int sepgpu_peak_copy(int device, double *gbytes_per_s);
/* write `bytes` of HBM to flush L2 between timed iterations */
int sepgpu_flush_l2(sepgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
