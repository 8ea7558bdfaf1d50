/* sep_oracle.h -- CPU oracle for the seplib hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C (C99, scalar, FP64) restatement of the algorithms on the reference's per-timestep
 * path, written on flat arrays (positions as x[3*i+k]) instead of the reference's AoS structs.
 * Every function cites the reference file:line it restates (paths relative to the reference root).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against the
 * reference itself (oracle/_ref/libsep_ref.so, compiled from the unmodified reference sources by
 * oracle/Makefile) and tests/test_golden.py checks it against committed vectors in tests/golden/
 * that were produced by that reference build (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 * The product (libsep.so) never links or calls it.
 */
#ifndef SEP_ORACLE_H
#define SEP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* exclusion rules: include/sepdef.h:29-36 */
#define ORC_ALL            1
#define ORC_EXCL_BONDED    2
#define ORC_EXCL_SAME_MOL  3

/* pair potential family: source/sepmisc.c:115-164 and source/sepprfrc.c:782-795 */
#define ORC_POT_LJ        0   /* sep_lj        : u = 4(r^-12 - r^-6)              */
#define ORC_POT_LJ_SHIFT  1   /* sep_lj_shift  : + SEP_LJCF2 (0.016316891136)     */
#define ORC_POT_WCA       2   /* sep_wca       : + 1.0                            */
#define ORC_POT_LJ_PARAM  3   /* sep_force_lj  : param = {cf, eps, sigma, aw}     */

#define ORC_SEP_BOND  10
#define ORC_SEP_ANGLE 10
#define ORC_SEP_DIHED 20

typedef struct {
    double epot, ecoul, ekin;
    double pot_P[9], kin_P[9], pot_P_bond[9];
} orc_ret;

/* Per-atom bonded-partner tables as the topology reader fills them (-1 terminated rows),
 * source/sepmol.c:84-95, :208-211, :312-327.  Any pointer may be NULL when opt does not need it. */
typedef struct {
    const int *molindex;   /* [n]            */
    const int *bond;       /* [n*ORC_SEP_BOND]  */
    const int *angle;      /* [n*ORC_SEP_ANGLE] */
    const int *dihed;      /* [n*ORC_SEP_DIHED] */
} orc_topo;

/* sep_Wrap, include/sepmisc.h:81-85 */
double orc_wrap(double d, double len);

/* sep_nsubbox + sep_sys_setup cell geometry, source/sepmisc.c:454-463, source/sepinit.c:257-276 */
void orc_cell_geometry(const double len[3], double cf, double delta, int nsub[3], double lsub[3]);

/* Half Verlet list through the linked-cell list:
 * sep_make_celllist (source/sepprfrc.c:394-415) + sep_make_neighblist_from_llist and its
 * _nonbonded / _excl_same_mol variants (:419-513, :517-603, :606-700).
 * Emits pairs (j1,j2) in the reference's visiting order into pairs[2*k], pairs[2*k+1].
 * Returns the number of pairs, or -1 if capacity (max_pairs) is exceeded, or -2 if some j1
 * collects >= 3000 (SEP_NEIGHB) partners ("Too many neighbours", :499-501). */
long orc_neighb_pairs(int n, const double *x, const double len[3], const int nsub[3],
                      const double lsub[3], double cutoff_plus_skin, unsigned opt,
                      const orc_topo *topo, int *pairs, long max_pairs);

/* All-pairs list used by SEP_NEIGHBLIST, sep_make_neighblist (source/sepprfrc.c:306-343);
 * note it excludes by sep_bond_share only. */
long orc_neighb_pairs_n2(int n, const double *x, const double len[3], double cutoff_plus_skin,
                         unsigned opt, const orc_topo *topo, int *pairs, long max_pairs);

/* sep_force_pair_neighb serial branch (source/sepprfrc.c:161-223) over a half pair list.
 * f is ACCUMULATED into (caller zeroes), ret->pot_P accumulated, ret->epot ASSIGNED (:222)
 * for pot<3 and ACCUMULATED for ORC_POT_LJ_PARAM (sep_lj_pair_neighb, :922). */
void orc_force_pairs_list(int n, const double *x, const char *type, const double len[3],
                          const int *pairs, long npairs, const char types[2], double cf,
                          int pot, const double *ljparam, double *f, orc_ret *ret);

/* sep_force_pair_brute / sep_lj_pair_brute (source/sepprfrc.c:19-91, :925-1000): epot accumulated.
 * EXCL_BONDED here means bond partners only (:33). */
void orc_force_pairs_brute(int n, const double *x, const char *type, const double len[3],
                           const char types[2], double cf, int pot, const double *ljparam,
                           unsigned opt, const orc_topo *topo, double *f, orc_ret *ret);

/* sep_coulomb_sf_neighb / _brute (source/sepcoulomb.c:96-160, :20-94). */
void orc_coulomb_sf_list(int n, const double *x, const double *z, const double len[3],
                         const int *pairs, long npairs, double cf, double *f, orc_ret *ret);
void orc_coulomb_sf_brute(int n, const double *x, const double *z, const double len[3], double cf,
                          unsigned opt, const orc_topo *topo, double *f, orc_ret *ret);

/* Bonded terms, source/sepmol.c:372-414, :469-516, :418-467, :520-587.  Lists have the reference's
 * layout: blist[3n]={a,b,type}, alist[4n]={a,b,c,type}, dlist[5n]={a,b,c,d,type}. */
void orc_stretch_harmonic(const double *x, const double len[3], const unsigned *blist, unsigned nb,
                          int type, double lbond, double ks, double *f, orc_ret *ret, double *blengths);
void orc_angle_harmonic(const double *x, const double len[3], const unsigned *alist, unsigned na,
                        int type, double angle0, double k, double *f, orc_ret *ret, double *angles);
void orc_angle_cossq(const double *x, const double len[3], const unsigned *alist, unsigned na,
                     int type, double angle0, double k, double *f, orc_ret *ret, double *angles);
void orc_torsion_ryckaert(const double *x, const double len[3], const unsigned *dlist, unsigned nd,
                          int type, const double g[6], double *f, orc_ret *ret, double *dihedrals);

/* sep_nosehoover, source/sepintgr.c:149-168.  Returns the updated alpha. */
double orc_nosehoover(int n, const double *v, const double *m, double *f, double temp0,
                      double alpha, double tau, double dt);
/* _sep_nosehoover_type, source/sepintgr.c:170-198 (alpha is a 3-slot history). */
void orc_nosehoover_type(int n, const double *v, const double *m, const char *type, char which,
                         double *f, double Td, double alpha[3], double Q, double dt);

/* sep_leapfrog + sep_periodic + trigger, source/sepintgr.c:18-88.
 * state: x,v,f,m as above; xn[3n], cross_neighb[3n], crossings[3n].
 * *max_dist2 in/out (reset by sep_reset_force, source/sepmisc.c:399).
 * Returns 1 if the rebuild trigger fired (then xn<-x, cross_neighb<-0 were applied). */
int orc_leapfrog(int n, double *x, double *v, const double *f, const double *m, double *a,
                 double *xn, int *cross_neighb, int *crossings, const double len[3], double dt,
                 double skin, double *max_dist2, orc_ret *ret);

/* sep_verlet_dpd, source/sepintgr.c:296-345. */
int orc_verlet_dpd(int n, double *x, double *v, const double *f, const double *m, double *a,
                   double *pv, double *pa, double *xn, int *cross_neighb, int *crossings,
                   const double len[3], double dt, double lambda, int stepnow, double skin,
                   double *max_dist2, orc_ret *ret);

/* ---- callers either side of the hot path ---- */
/* sep_relax_temp + sep_reset_momentum (source/sepmisc.c:357-390, :1173-1192).  Returns the type's kinetic
 * energy before the rescale. */
double orc_relax_temp(int n, double *v, const double *m, const char *type, char which, double Td, double tau, double dt);
/* sep_force_x0 with sep_spring_x0 (source/sepmisc.c:167-181, :645-670): f accumulated, no energy returned */
void orc_force_x0(int n, const double *x, const double *x0, const char *type, char which, const double len[3], double *f);
/* sep_compress_box (source/sepmisc.c:994-1026).  len/nsub/lsub/volume in-out; returns 1 when the box changed. */
int orc_compress_box(int n, double *x, double rhoD, double xi, double len[3], int nsub[3], double lsub[3],
                     double *volume, double cf, double skin, int list_mode);
/* sep_berendsen (iso = 0, z only, :892-914) and sep_berendsen_iso (iso = 1, :918-944); p is ret->p. */
void orc_berendsen(int n, double *x, double Pd, double beta, double p, double dt, int iso, double len[3],
                   int nsub[3], double lsub[3], double *volume, double cf, int list_mode);

/* sep_randn (source/sepmisc.c:1131-1160): polar Box-Muller on glibc rand(), second deviate cached.
 * orc_randn_reset() drops the cached deviate (a fresh process in the reference). */
double orc_randn(void);
void orc_randn_reset(void);
/* sep_fp, source/sepintgr.c:235-293 (state as orc_leapfrog; ldiff[n]; draws 3n numbers from orc_randn) */
int orc_fp(int n, double *x, double *v, const double *f, const double *m, const double *ldiff, double *xn,
           int *cross_neighb, int *crossings, const double len[3], double dt, double temp, double skin,
           double *max_dist2, orc_ret *ret);
/* sep_langevinGJF, source/sepintgr.c:89-146 (prevf[3n], randn[3n] carried between calls) */
int orc_langevin_gjf(int n, double *x, double *v, const double *f, const double *m, double *a, double *prevf, double *randn,
                     double *xn, int *cross_neighb, int *crossings, const double len[3], double dt, double temp,
                     double alpha, double skin, double *max_dist2, orc_ret *ret);

/* sep_dpdforce_neighb (source/sepprfrc.c:1007-1133) with the reference's glibc rand() stream
 * replaced by the product's counter-based pair generator (see orc_dpd_uniform); the reference's
 * stream cannot be reproduced by any parallel evaluation order. */
double orc_dpd_uniform(unsigned long long seed, unsigned long long step, unsigned i, unsigned j);
void orc_dpd_force_list(int n, const double *x, const double *pv, const char *type,
                        const double len[3], const int *pairs, long npairs, const char types[2],
                        double cf, double aij, double temp, double sigma, double dt,
                        unsigned long long seed, unsigned long long step, double *f, orc_ret *ret);

#ifdef __cplusplus
}
#endif
#endif
