/* sep.h -- the seplib C99 API, as served by seplib-b200 (B200 / sm_100a device path).
 *
 * This header is the drop-in boundary: programs written against the reference's include/sep.h
 * (prgs/prg0-9.c) compile against it unchanged and link with -lsep -- prg5 included: its OpenMP "model II"
 * helpers sep_omp_bond / sep_omp_angle / sep_omp_torsion (source/sepomp.c) are served from the device; the pair-force
 * members of that family (sep_omp_pairs, sep_omp_coulomb, sep_omp_dpd_pairs) are not provided.  Type names, field
 * names, field order, constants and prototypes follow the reference so that user code which reads
 * atoms[i].x/v/f, sys.nupdate_neighb or ret.epot directly keeps working; each block cites the
 * reference header it mirrors (paths relative to the reference root).
 *
 * What is different underneath: the per-timestep hot path (pair forces, neighbour list, bonded
 * terms, Coulomb, DPD, thermostat, integrators) runs on the GPU through include/sepgpu.h.  There is
 * no CPU implementation of that path in this library: without a usable CUDA device those calls stop
 * with sep_error().
 *
 * Host/device coherence (env SEP_SYNC, or sep_gpu_set_sync):
 *   auto (default)  atoms[] lives in pages the library protects while the device copy is newer.  The first READ of any
 *                   atoms[i] member by user code faults, the library brings the array up to date and the read continues;
 *                   the first WRITE marks the array as changed by the host and it is uploaded before the next hot call.
 *                   Programs that leave atoms[] alone inside the loop pay nothing; programs that read or write it
 *                   (reference prgs/prg0.c:64 prints atoms[0].f, prgs/prg5.c writes forces) stay correct without any
 *                   change.  sepret / sepsys scalars are refreshed after every hot call.  Arrays that did not come from
 *                   sep_init / sep_init_xyz cannot be protected and are handled as in "step".
 *   step            atoms[] is refreshed from the device at the end of every integrator call;
 *                   sepret / sepsys scalars after every hot call.
 *   lazy            atoms[] is refreshed only by library calls that read it (sep_eval_mom,
 *                   sep_save_xyz, ...), by sep_gpu_sync() and by sep_close(); scalars after
 *                   integrator calls.
 *   full            like step, plus forces are written back after every force call.
 * In step / lazy / full mode, writing into atoms[] from user code between hot calls needs sep_gpu_invalidate(atoms);
 * in auto mode the library notices by itself.
 *
 * Pair functions: sep_force_pairs recognises sep_lj, sep_lj_shift and sep_wca by address; any other
 * double fun(double r2, char opt) is sampled once per (function, cutoff) and interpolated on the device
 * (env SEP_TABLE_N, SEP_TABLE_RMIN; sep_pairs_retabulate()).
 *
 * Several GPUs: env SEP_NGPU=N runs the unchanged program on N GPUs of the node -- the library forks one copy per GPU at
 * the first hot call and decomposes the box into slabs along z (INTEGRATION.md section 5 lists what the program must
 * satisfy).
 */
#ifndef SEP_B200_SEP_H
#define SEP_B200_SEP_H

#include <stdio.h>
#include <stdlib.h>
#include <stdbool.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <float.h>
#include <stdarg.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (include/sepdef.h:16-64) --------------------------------------------------- */
#define SEP_FALSE 0
#define SEP_TRUE  1

#define SEP_BOND  10            /* bonded partners kept per atom */
#define SEP_ANGLE 10
#define SEP_DIHED 20

#define SEP_NEIGHB      3000    /* reference half-list capacity per atom */
#define SEP_NUM_NEIGHB  3000
#define SEP_NO_NEIGHB   0

#define SEP_NEIGHB_ALL            1
#define SEP_ALL                   1
#define SEP_NEIGHB_EXCL_BONDED    2
#define SEP_EXCL_BONDED           2
#define SEP_NEIGHB_EXCL_SAME_MOL  3
#define SEP_EXCL_SAME_MOL         3

#define SEP_MAX_NUM_MOL 10000

#define SEP_PI     3.14159265358979
#define SEP_WCACF  1.12246204830937
#define SEP_LJCF2  0.016316891136

#define SEP_R250_LEN 250
#define SEP_R521_LEN 521
#define SEP_MAXNID   100000
#define SEP_NRETVALS 10

#define SEP_BRUTE            0
#define SEP_NEIGHBLIST       1
#define SEP_LLIST_NEIGHBLIST 2

#define SEP_LEAPFROG   0
#define SEP_NOSEHOOVER 1
#define SEP_ANDERSEN   2
#define SEP_SHAKE      3
#define SEP_SHAKE_MAXIT 1000

#define SEP_SUCCESS 1
#define SEP_FAILURE 1           /* sic: identical in the reference (include/sepdef.h:63-64) */

/* ---- macros (include/sepmisc.h:36-94) ------------------------------------------------------- */
#define SEP_FLUSH  fflush(stdout)
#define SEP_TICTOC clock_t time_then
#define SEP_TIC    time_then = clock()
#define SEP_TOC    ((int)(1000*(clock()-time_then)/CLOCKS_PER_SEC))%1000

#define sep_rand()   ( rand()/(RAND_MAX+1.0) )
#define sep_here(x)  { printf("%d\n", x); SEP_FLUSH; }
#define sep_Sq(x)    ( (x)*(x) )
#define sep_Abs(x)   ( (x) > 0.0 ? (x) : -(x) )
#define sep_Wrap( x, y )  { if ( x > 0.5*y ) x -= y; else if ( x < -0.5*y ) x += y; }
#define sep_Periodic( x, y )  { if ( x > y ) x -= y; else if ( x < 0 ) x += y; }

/* ---- data model (reference include/sepstrct.h:23-204) -------------------------------------------------
 * Programs built against the reference read and write these records directly, so every member keeps the
 * reference's name, type and position (tests/test_cpu_host.py compares sizes and offsets with the compiled
 * reference).  Members this library maintains:
 *   seppart   x v f a m type z   state of one atom (x wrapped into the box);  neighb: host list row, filled only by
 *             sep_gpu_export_neighb;  cross_neighb / crossings: box crossings since the last list build / since
 *             the start;  molindex (-1: none) and the bonded partner tables bond / angle / dihed (-1 terminated);
 *             xtrue x0 xn: unwrapped, tether and last-list-build positions;  pv pa: DPD predictor state.
 *             sigma collid colltime ldiff xp px randn prevf belong to integrators that are not part of this library.
 *   sepmolinfo  topology lists (blist: a b type; alist: a b c type; dlist: a b c d type), their measured
 *             values, and the molecule-pair force table Fij.
 *   sepsys    box, time step, cell grid, neighbour-list switches and the rebuild flag.
 *   sepmol    per-molecule derived data (centre of mass, velocity, member atoms).
 *   sepret    sums returned by the hot calls (energies, pressure tensors). */
typedef struct {
    double x[3], v[3], f[3], a[3], m;
    char type;
    double z;
    int *neighb;
    int cross_neighb[3], crossings[3], molindex, bond[SEP_BOND], angle[SEP_ANGLE], dihed[SEP_DIHED];
    double sigma;
    int *collid;
    double *colltime;
    double ldiff, xtrue[3], x0[3], xn[3], xp[3], px[3], pv[3], pa[3], randn[3], prevf[3];
} seppart;
typedef seppart sepatom;

typedef struct {
    unsigned num_mols, max_nuau;
    int flag_bonds, flag_angles, flag_dihedrals;
    unsigned num_bonds, *blist, num_btypes;
    unsigned num_angles, *alist, num_atypes;
    unsigned num_dihedrals, *dlist, num_dtypes;
    double *blengths, *angles, *dihedrals;
    unsigned flag_Fij;
    float ***Fij, ***Fiajb;
} sepmolinfo;

typedef struct {
    long int npart;
    double length[3], volume;
    int intgr_type;
    double dt, tnow;
    unsigned ndof;
    double max_dist2, cf, lsubbox[3];
    int nsubbox[3];
    double skin;
    unsigned neighb_update, neighb_flag, nupdate_neighb;
    bool omp_flag;
    unsigned int nthreads;
    int fun_cstate;
    sepmolinfo *molptr;
} sep3D;
typedef sep3D sepsys;

typedef struct {
    double m, x[3], xtrue[3], v[3];
    unsigned nuau;
    int *index;
    double ete[3], re2, rg, S[3], s[3], inertia[3][3], w[3];
    int method_w;
    double pel[3];
    char type;
    unsigned nbonds;
    double *blength;
    int shake_flag;
} sepmol;

typedef struct {
    double etot, ekin, epot, ecoul, sumv2;
    double P[3][3], kin_P[3][3], pot_P[3][3], p;
    double P_mol[3][3], kin_P_mol[3][3], pot_P_mol[3][3], p_mol;
    double pot_P_conservative[3][3], pot_P_random[3][3], pot_P_dissipative[3][3], pot_P_bond[3][3];
    double pot_T_mol[3][3], kin_T_mol[3][3], T_mol[3][3], t_mol;
} sepret;

/* ---- setup and teardown (include/sepinit.h:29-84) --------------------------------------------- */
seppart *sep_init(size_t npart, size_t nneighb);
void sep_close(seppart *ptr, size_t npart);
seppart *sep_init_xyz(double *lbox, int *npart, const char *file, char verbose);
sepsys sep_sys_setup(double lengthx, double lengthy, double lengthz,
                     double maxsyscf, double dt, size_t npart, size_t update);
void sep_free_sys(sepsys *ptr);
void sep_set_lattice(seppart *ptr, sepsys sys);
void sep_set_vel(seppart *ptr, double temp, sepsys sys);
void sep_set_vel_seed(seppart *ptr, double temp, unsigned int seed, sepsys sys);
void sep_set_vel_type(seppart *ptr, char type, double temp, unsigned int seed, sepsys sys);

/* ---- pair forces and neighbour list (include/sepprfrc.h:49-93) --------------------------------- */
int sep_force_pairs(seppart *ptr, const char *types, double cf,
                    double (*fun)(double, char), sepsys *sys,
                    sepret *retval, const unsigned opt);
void sep_force_lj(seppart *ptr, const char *types, const double *param,
                  sepsys *sys, sepret *retval, const unsigned opt);
void sep_force_dpd(seppart *ptr, const char *types, const double cf, const double aij,
                   const double temp_desired, const double sigma,
                   sepsys *sys, sepret *retval, const unsigned opt);
void sep_neighb(seppart *ptr, sepsys *sys);
void sep_neighb_nonbonded(seppart *ptr, sepsys *sys);
void sep_neighb_excl_same_mol(seppart *ptr, sepsys *sys);
unsigned int sep_bond_share(seppart *ptr, int j1, int j2);
unsigned int sep_angle_share(seppart *ptr, int j1, int j2);
unsigned int sep_dihed_share(seppart *ptr, int j1, int j2);
unsigned int sep_bonded(seppart *ptr, int i, int j);

/* ---- shifted-force Coulomb (include/sepcoulomb.h:39-53) ----------------------------------------- */
void sep_coulomb_sf(seppart *ptr, double cf, sepsys *sys, sepret *retval, const unsigned opt);

/* ---- thermostat and integrators (include/sepintgr.h:27-90) -------------------------------------- */
double sep_periodic(sepatom *atoms, unsigned n, sepsys *sys);
void sep_leapfrog(seppart *ptr, sepsys *sys, sepret *retval);
void sep_nosehoover(seppart *ptr, double Td, double *alpha, const double Q, sepsys *sys);
void _sep_nosehoover_type(seppart *ptr, char type, double Td, double *alpha, const double Q, sepsys *sys);
void sep_verlet_dpd(seppart *ptr, double lambda, int stepnow, sepsys *sys, sepret *retval);

/* ---- molecules: topology and bonded forces (include/sepmol.h:18-100) ----------------------------- */
void sep_read_topology_file(sepatom *aptr, const char *file, sepsys *sysptr, char opt);
void sep_free_bonds(sepmolinfo *ptr);
void sep_free_angles(sepmolinfo *ptr);
void sep_free_dihedrals(sepmolinfo *ptr);
sepmol *sep_init_mol(sepatom *atom, sepsys *sys);
void sep_free_mol(sepmol *ptr, sepsys *sys);
void sep_stretch_harmonic(sepatom *aptr, int type, const double lbond, const double ks,
                          sepsys *sys, sepret *ret);
void sep_angle_harmonic(sepatom *ptr, int type, const double angle0, const double k,
                        sepsys *sys, sepret *ret);
void sep_angle_cossq(sepatom *ptr, int type, const double angle0, const double k,
                     sepsys *sys, sepret *ret);
void sep_torsion_Ryckaert(sepatom *ptr, int type, const double g[6], sepsys *sys, sepret *ret);
/* OpenMP "model II" helpers of prgs/prg5.c (reference include/sepomp.h:87-123, source/sepomp.c:179-329): the forces of one
 * bonded term kind ADDED to the caller's matrix ftot[npart][3]; atoms[].f and sepret are not touched.  Computed on the
 * device from the current positions; calls from different OpenMP sections are serialised inside the library. */
void sep_omp_bond(double **ftot, seppart *aptr, int type, const double lbond, const double ks, sepsys *sys);
void sep_omp_angle(double **ftot, seppart *ptr, int type, const double angle0, const double k, sepsys *sys);
void sep_omp_torsion(double **ftot, seppart *ptr, int type, const double g[6], sepsys *sys);
void sep_mol_cm(seppart *ptr, sepmol *mol, sepsys *sys);
void sep_mol_eval_xtrue(seppart *ptr, sepmol *mol, sepsys sys);
void sep_mol_spin(sepatom *atom, sepmol *mol, sepsys *sys, bool safe);
void sep_mol_dipoles(seppart *atom, sepmol *mol, sepsys *sys);
void sep_mol_velcm(seppart *atom, sepmol *mol, sepsys *sys);
void sep_eval_mol_pressure_tensor(sepatom *atoms, sepmol *mols, sepret *ret, sepsys *sys);
double sep_average_bondlengths(int type, sepsys *sys);

/* ---- return values (include/sepret.h:36-67) ------------------------------------------------------- */
void sep_reset_retval(sepret *retval);
double sep_get_pressure(sepret *retval, sepsys *sys);
double sep_get_temperature(sepret *retval, sepsys *sys);
void sep_pressure_tensor(sepret *retval, sepsys *sys);
void sep_mol_pressure_tensor(sepatom *atoms, sepmol *mols, sepret *ret, sepsys *sys);

/* ---- miscellaneous runtime (include/sepmisc.h:101-450) ---------------------------------------------- */
void sep_error(char *str, ...);
void sep_warning(char *str, ...);
double sep_lj(double r2, char opt);
double sep_lj_shift(double r2, char opt);
double sep_wca(double r2, char opt);
/* Extension: sep_force_pairs samples a pair function of the caller's own once per (function, cutoff) -- see
 * INTEGRATION.md "Pair functions of your own".  Call this after changing parameters that function reads. */
void sep_pairs_retabulate(void);
void sep_reset_force(seppart *ptr, sepsys *sys);
void sep_reset_force_mol(sepsys *sys);
int sep_nsubbox(double cf, double delta, double lbox);
double sep_box_length(double dens, int npart, int ndim);
int sep_count_type(seppart *ptr, char spec, int npart);
void sep_set_x0(seppart *ptr, int npart);
void sep_set_xn(seppart *ptr, int npart);
void sep_save_xyz(seppart *ptr, const char *partnames, const char *file, char *mode, sepsys *sys);
double sep_eval_mom(seppart *ptr, int npart);
double sep_eval_mom_type(seppart *ptr, char type, int dir, int npart);
void sep_compress_box(sepatom *ptr, double rhoD, double xi, sepsys *sys);
void sep_compress_box_dir(sepatom *ptr, double rhoD, double xi, int dir, sepsys *sys);
void sep_compress_box_dir_length(sepatom *ptr, double length, double xi, int dir, sepsys *sys);
void sep_berendsen(sepatom *ptr, double Pd, double beta, sepret *ret, sepsys *sys);
void sep_berendsen_iso(sepatom *ptr, double Pd, double beta, sepret *ret, sepsys *sys);
void sep_relax_temp(seppart *ptr, char type, double Td, double tau, sepsys *sys);
double sep_randn(void);
void sep_fp(seppart *ptr, double temp_desired, sepsys *sys, sepret *retval);
void sep_langevinGJF(sepatom *ptr, double temp0, double alpha, sepsys *sys, sepret *retval);
void sep_set_ldiff(sepatom *ptr, char type, double ldiff, sepsys sys);
double sep_spring_x0(double r2, char opt);
void sep_force_x0(seppart *ptr, char type, double (*fun)(double, char), sepsys *sys);
void sep_set_charge(seppart *ptr, char type, double z, sepsys sys);
void sep_set_mass(seppart *ptr, char type, double m, sepsys sys);
void sep_set_type(seppart *ptr, char spec, int numb, sepsys *sys);
void sep_set_omp(unsigned nthreads, sepsys *sys);
void sep_set_skin(sepsys *sys, double value);
void sep_set_ndof(size_t ndof, sepsys *sys);
void sep_reset_momentum(seppart *ptr, const char type, sepsys *sys);
double sep_dist_ij(double *r, seppart *ptr, int i, int j, sepsys *sys);
void sep_eval_xtrue(seppart *ptr, sepsys *sys);

/* ---- array helpers used by the example programs (include/separray.h, include/seputil.h) -------------- */
double *sep_vector(size_t length);
int *sep_vector_int(size_t length);
double **sep_matrix(size_t nrow, size_t ncol);
void sep_free_matrix(double **ptr, size_t nrow);
void sep_matrix_set(double **a, size_t nrow, size_t ncol, double value);
float ***sep_tensor_float(size_t nx, size_t ny, size_t nz);
void sep_free_tensor_float(float ***ptr, size_t nx, size_t ny);
double sep_dot(double *a, double *b, int length);
void sep_vector_set(double *vec, size_t length, double value);

/* ---- samplers (reference include/sepsampler.h): host post-processing of the synchronised data.
 * "sacf", "vacf", "msd", "gh", "profs", "radial", "msacf", "mvacf" and "mgh" write the reference's files in the reference's
 * format (seplib_b200/csrc/host/sep_sampler.c); the remaining names are accepted and record nothing.
 * With env SEP_SAMPLER_FEEDS=1 (single-GPU runs) "vacf", "msd", "profs", "gh" and "radial" take their sums from the device
 * (sepgpu_feed_*, include/sepgpu.h) and atoms[] is not downloaded for them; "gh" then leaves atoms[].xtrue untouched.
 * The per-sampler state is private to the library. ----------------------------------------------------- */
typedef struct {
    sepmol *molptr;
    void *impl;
    long unsigned msd_counter;
} sepsampler;
sepsampler sep_init_sampler(void);
void sep_add_sampler(sepsampler *sptr, const char *sampler, sepsys sys, int lvec, ...);
void sep_add_mol_sampler(sepsampler *sptr, sepmol *mols);
void sep_sample(seppart *pptr, sepsampler *sptr, sepret *ret, sepsys sys, unsigned n);
void sep_close_sampler(sepsampler *ptr);

/* ---- seplib-b200 extensions (not in the reference) -------------------------------------------------- */
#define SEP_SYNC_LAZY 0
#define SEP_SYNC_STEP 1
#define SEP_SYNC_FULL 2
#define SEP_SYNC_AUTO 3
void sep_gpu_set_sync(int mode);                 /* overrides env SEP_SYNC                              */
void sep_gpu_sync(seppart *ptr);                 /* device -> atoms[] (everything that changed)         */
void sep_gpu_invalidate(seppart *ptr);           /* atoms[] was edited by the caller: re-upload         */
void sep_gpu_sync_scalars(seppart *ptr, sepsys *sys, sepret *ret);   /* sepret/sepsys scalars only   */
long sep_gpu_export_neighb(seppart *ptr, sepsys *sys, int *pairs, long max_pairs);   /* list as (i<j) pairs */
void *sep_gpu_handle(seppart *ptr);              /* the sepgpu_ctx behind an atom array (or NULL)       */
void sep_gpu_set_dpd_seed(unsigned long long seed);

#ifdef __cplusplus
}
#endif
#endif
