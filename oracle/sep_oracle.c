/* sep_oracle.c -- CPU oracle for the seplib hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C99 restatement (scalar FP64, flat arrays) of the reference algorithms named in
 * sep_oracle.h.  Parity status: PINNED against the compiled reference (oracle/_ref) and the
 * committed golden vectors; see sep_oracle.h.  Build with -O2 -fno-fast-math -ffp-contract=off
 * so that every product/sum rounds exactly as written.
 */
#include "sep_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_NEIGHB_LIMIT 3000           /* SEP_NEIGHB, include/sepdef.h:23 */
#define ORC_PI 3.14159265358979         /* SEP_PI, include/sepdef.h:40 (14 digits, on purpose) */
#define ORC_LJCF2 0.016316891136        /* SEP_LJCF2, include/sepdef.h:44 */

/* include/sepmisc.h:81-85 */
double orc_wrap(double d, double len)
{
    if (d > 0.5 * len) d -= len;
    else if (d < -0.5 * len) d += len;
    return d;
}

/* min-image separation and its square, accumulated in the reference's order r2 = ((0+dx^2)+dy^2)+dz^2 */
static double sep_r2(const double *xi, const double *xj, const double len[3], double r[3])
{
    double r2 = 0.0;
    for (int k = 0; k < 3; k++) {
        r[k] = orc_wrap(xi[k] - xj[k], len[k]);
        r2 += r[k] * r[k];
    }
    return r2;
}

void orc_cell_geometry(const double len[3], double cf, double delta, int nsub[3], double lsub[3])
{
    for (int k = 0; k < 3; k++) {
        nsub[k] = (int)(len[k] / (cf + delta));      /* source/sepmisc.c:454-463 */
        lsub[k] = len[k] / nsub[k];                  /* source/sepinit.c:274-276 */
    }
}

/* ---- exclusion predicates, source/sepprfrc.c:703-740 ------------------------------------ */
static int share(const int *tab, int width, int a, int b)
{
    for (int k = 0; k < width; k++) {
        int ta = tab[a * width + k], tb = tab[b * width + k];
        if (ta == -1 || tb == -1) break;
        if (ta == b || tb == a) return 1;
    }
    return 0;
}
static int bond_share(const orc_topo *t, int a, int b) { return share(t->bond, ORC_SEP_BOND, a, b); }
static int bonded_sum(const orc_topo *t, int a, int b)
{
    return share(t->bond, ORC_SEP_BOND, a, b) + share(t->angle, ORC_SEP_ANGLE, a, b) +
           share(t->dihed, ORC_SEP_DIHED, a, b);
}

/* ---- neighbour pairs ------------------------------------------------------------------------ */
long orc_neighb_pairs(int n, const double *x, const double len[3], const int nsub[3],
                      const double lsub[3], double cutoff_plus_skin, unsigned opt,
                      const orc_topo *topo, int *pairs, long max_pairs)
{
    /* half stencil, source/sepprfrc.c:424-426 */
    static const int ox[14] = {0, 1, 1, 0, -1, 0, 1, 1, 0, -1, -1, -1, 0, 1};
    static const int oy[14] = {0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, -1, -1, -1};
    static const int oz[14] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1};

    const int nxy = nsub[0] * nsub[1];
    const int ncell = nxy * nsub[2];
    const double cut2 = cutoff_plus_skin * cutoff_plus_skin;

    int *head = malloc(sizeof(int) * (size_t)ncell);
    int *next = malloc(sizeof(int) * (size_t)n);
    int *count = calloc((size_t)n, sizeof(int));
    long np = 0;
    int status = 0;

    /* sep_make_celllist: head insertion, so chains run from high to low index (:394-415) */
    for (int c = 0; c < ncell; c++) head[c] = -1;
    for (int i = 0; i < n; i++) {
        int c = (int)(x[3 * i] / lsub[0]) + (int)(x[3 * i + 1] / lsub[1]) * nsub[0] +
                (int)(x[3 * i + 2] / lsub[2]) * nxy;
        next[i] = head[c];
        head[c] = i;
    }

    for (int cz = 0; cz < nsub[2] && !status; cz++)
    for (int cy = 0; cy < nsub[1] && !status; cy++)
    for (int cx = 0; cx < nsub[0] && !status; cx++) {
        const int c1 = cz * nxy + cy * nsub[0] + cx;
        for (int j1 = head[c1]; j1 != -1 && !status; j1 = next[j1]) {
            for (int o = 0; o < 14 && !status; o++) {
                int nx_ = cx + ox[o], ny_ = cy + oy[o], nz_ = cz + oz[o];
                if (nx_ == nsub[0]) nx_ = 0; else if (nx_ == -1) nx_ = nsub[0] - 1;
                if (ny_ == nsub[1]) ny_ = 0; else if (ny_ == -1) ny_ = nsub[1] - 1;
                if (nz_ == nsub[2]) nz_ = 0;                 /* z offsets are never negative */
                const int c2 = nz_ * nxy + ny_ * nsub[0] + nx_;
                for (int j2 = head[c2]; j2 != -1; j2 = next[j2]) {
                    if (!(c1 != c2 || j2 > j1)) continue;
                    if (opt == ORC_EXCL_BONDED && bonded_sum(topo, j1, j2) > 0) continue;   /* :583 */
                    if (opt == ORC_EXCL_SAME_MOL &&
                        !(topo->molindex[j1] == -1 || topo->molindex[j1] != topo->molindex[j2]))
                        continue;                                                           /* :673 */
                    double r[3];
                    if (sep_r2(&x[3 * j1], &x[3 * j2], len, r) < cut2) {
                        if (np >= max_pairs) { status = -1; break; }
                        pairs[2 * np] = j1; pairs[2 * np + 1] = j2; np++;
                        count[j1]++;
                    }
                    if (count[j1] == ORC_NEIGHB_LIMIT) { status = -2; break; }
                }
            }
        }
    }
    /* The reference keeps the pairs in per-atom rows ptr[j1].neighb[] and its force loops walk the atoms
     * in index order (source/sepprfrc.c:161-170): regroup by j1 (stable) so that a consumer of this
     * pair array adds forces in exactly the reference's order. */
    if (!status && np > 0) {
        long *start = calloc((size_t)n + 1, sizeof(long));
        int *tmp = malloc(sizeof(int) * 2 * (size_t)np);
        for (long p = 0; p < np; p++) start[pairs[2 * p] + 1]++;
        for (int i = 0; i < n; i++) start[i + 1] += start[i];
        for (long p = 0; p < np; p++) {
            const long q = start[pairs[2 * p]]++;
            tmp[2 * q] = pairs[2 * p]; tmp[2 * q + 1] = pairs[2 * p + 1];
        }
        memcpy(pairs, tmp, sizeof(int) * 2 * (size_t)np);
        free(start); free(tmp);
    }
    free(head); free(next); free(count);
    return status ? status : np;
}

long orc_neighb_pairs_n2(int n, const double *x, const double len[3], double cutoff_plus_skin,
                         unsigned opt, const orc_topo *topo, int *pairs, long max_pairs)
{
    const double cut2 = cutoff_plus_skin * cutoff_plus_skin;
    long np = 0;
    for (int a = 0; a < n - 1; a++)
        for (int b = a + 1; b < n; b++) {
            if (opt == ORC_EXCL_BONDED && bond_share(topo, a, b) == 1) continue;
            if (opt == ORC_EXCL_SAME_MOL && topo->molindex[a] == topo->molindex[b] &&
                topo->molindex[a] != -1) continue;
            double r[3];
            if (sep_r2(&x[3 * a], &x[3 * b], len, r) < cut2) {
                if (np >= max_pairs) return -1;
                pairs[2 * np] = a; pairs[2 * np + 1] = b; np++;
            }
        }
    return np;
}

/* ---- pair potentials ------------------------------------------------------------------------ */
typedef struct { int kind; double eps48, eps4, awh, aw, sig2, shift; } potdef;

static potdef make_pot(int pot, const double *p)
{
    potdef d; memset(&d, 0, sizeof d); d.kind = pot;
    if (pot == ORC_POT_LJ_PARAM) {               /* source/sepprfrc.c:785-795: {cf, eps, sigma, aw} */
        const double cf = p[0], eps = p[1], sigma = p[2], aw = p[3];
        d.shift = 4.0 * eps * (pow(sigma / cf, 12.) - aw * pow(sigma / cf, 6.));
        d.eps48 = 48.0 * eps; d.eps4 = 4.0 * eps; d.awh = 0.5 * aw; d.aw = aw; d.sig2 = sigma * sigma;
    }
    return d;
}

/* returns ft (force/r) and *u for one in-range pair */
static double pot_eval(const potdef *d, double r2, double *u)
{
    if (d->kind == ORC_POT_LJ_PARAM) {           /* source/sepprfrc.c:887-896 */
        double rri = d->sig2 / r2, rri3 = rri * rri * rri;
        *u = d->eps4 * rri3 * (rri3 - d->aw) - d->shift;
        return d->eps48 * rri3 * (rri3 - d->awh) * rri;
    }
    /* source/sepmisc.c:115-164 */
    double rri = 1.0 / r2, rri3 = rri * rri * rri;
    double shift = d->kind == ORC_POT_LJ_SHIFT ? ORC_LJCF2 : (d->kind == ORC_POT_WCA ? 1.0 : 0.0);
    *u = 4.0 * rri3 * (rri3 - 1.0);
    if (d->kind != ORC_POT_LJ) *u = *u + shift;
    return 48.0 * rri3 * (rri3 - 0.5) * rri;
}

static int type_match(char ti, char tj, const char types[2])
{
    return (ti == types[0] && tj == types[1]) || (ti == types[1] && tj == types[0]);
}

void orc_force_pairs_list(int n, const double *x, const char *type, const double len[3],
                          const int *pairs, long npairs, const char types[2], double cf,
                          int pot, const double *ljparam, double *f, orc_ret *ret)
{
    (void)n;
    const potdef d = make_pot(pot, ljparam);
    const double cf2 = cf * cf;
    double epot = 0.0;
    for (long p = 0; p < npairs; p++) {
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (!type_match(type[i], type[j], types)) continue;
        double r[3], u;
        const double r2 = sep_r2(&x[3 * i], &x[3 * j], len, r);
        if (!(r2 < cf2)) continue;
        const double ft = pot_eval(&d, r2, &u);
        double fk[3];
        for (int k = 0; k < 3; k++) {
            fk[k] = ft * r[k];
            f[3 * i + k] += fk[k];
            f[3 * j + k] += -fk[k];
        }
        epot += u;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->pot_P[3 * k + kk] += fk[k] * r[kk];
    }
    if (pot == ORC_POT_LJ_PARAM) ret->epot += epot;   /* source/sepprfrc.c:922 */
    else ret->epot = epot;                             /* source/sepprfrc.c:222 (assignment) */
}

void orc_force_pairs_brute(int n, const double *x, const char *type, const double len[3],
                           const char types[2], double cf, int pot, const double *ljparam,
                           unsigned opt, const orc_topo *topo, double *f, orc_ret *ret)
{
    const potdef d = make_pot(pot, ljparam);
    const double cf2 = cf * cf;
    for (int a = 0; a < n - 1; a++)
        for (int b = a + 1; b < n; b++) {
            if (opt == ORC_EXCL_BONDED && bond_share(topo, a, b) == 1) continue;      /* :33 */
            if (opt == ORC_EXCL_SAME_MOL && topo->molindex[a] == topo->molindex[b] &&
                topo->molindex[a] != -1) continue;
            if (!type_match(type[a], type[b], types)) continue;
            double r[3], u;
            const double r2 = sep_r2(&x[3 * a], &x[3 * b], len, r);
            if (!(r2 < cf2)) continue;
            const double ft = pot_eval(&d, r2, &u);
            double fk[3];
            for (int k = 0; k < 3; k++) {
                fk[k] = ft * r[k];
                f[3 * a + k] += fk[k];
                f[3 * b + k] -= fk[k];
            }
            ret->epot += u;
            for (int k = 0; k < 3; k++)
                for (int kk = 0; kk < 3; kk++) ret->pot_P[3 * k + kk] += fk[k] * r[kk];
        }
}

/* ---- shifted-force Coulomb -------------------------------------------------------------------- */
static void coulomb_pair(const double *x, const double *z, const double len[3], int i, int j,
                         double cf, double cf2, double icf2, double icf, double *f, orc_ret *ret)
{
    double dr[3];
    const double r2 = sep_r2(&x[3 * i], &x[3 * j], len, dr);
    if (!(r2 < cf2)) return;
    const double zizj = z[j] * z[i];
    const double r = sqrt(r2);
    const double ft = zizj * (1.0 / r2 - icf2) / r;
    double fk[3];
    for (int k = 0; k < 3; k++) {
        fk[k] = ft * dr[k];
        f[3 * i + k] += fk[k];
        f[3 * j + k] -= fk[k];
    }
    for (int k = 0; k < 3; k++)
        for (int kk = 0; kk < 3; kk++) ret->pot_P[3 * k + kk] += fk[k] * dr[kk];
    const double ec = zizj * (1.0 / r + (r - cf) * icf2 - icf);
    ret->epot += ec;
    ret->ecoul += ec;
}

void orc_coulomb_sf_list(int n, const double *x, const double *z, const double len[3],
                         const int *pairs, long npairs, double cf, double *f, orc_ret *ret)
{
    (void)n;
    const double cf2 = cf * cf, icf2 = 1.0 / cf2, icf = 1.0 / cf;
    for (long p = 0; p < npairs; p++) {
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (fabs(z[i]) < DBL_EPSILON) continue;          /* only the list owner is tested, :102 */
        coulomb_pair(x, z, len, i, j, cf, cf2, icf2, icf, f, ret);
    }
}

void orc_coulomb_sf_brute(int n, const double *x, const double *z, const double len[3], double cf,
                          unsigned opt, const orc_topo *topo, double *f, orc_ret *ret)
{
    const double cf2 = cf * cf, icf2 = 1.0 / cf2, icf = 1.0 / cf;
    for (int a = 0; a < n - 1; a++) {
        if (fabs(z[a]) < DBL_EPSILON) continue;
        for (int b = a + 1; b < n; b++) {
            if (opt == ORC_EXCL_BONDED && bonded_sum(topo, a, b) == 1) continue;     /* :37, "== 1" */
            if (opt == ORC_EXCL_SAME_MOL && topo->molindex[a] == topo->molindex[b] &&
                topo->molindex[a] != -1) continue;
            coulomb_pair(x, z, len, a, b, cf, cf2, icf2, icf, f, ret);
        }
    }
}

/* ---- bonded terms ------------------------------------------------------------------------------- */
static double dot3(const double *a, const double *b)   /* sep_dot, source/seputil.c:393-403 */
{
    double s = 0.0;
    for (int k = 0; k < 3; k++) s += a[k] * b[k];
    return s;
}

void orc_stretch_harmonic(const double *x, const double len[3], const unsigned *blist, unsigned nb,
                          int type, double lbond, double ks, double *f, orc_ret *ret, double *blengths)
{
    for (unsigned n = 0; n < nb; n++) {
        if ((int)blist[3 * n + 2] != type) continue;
        const unsigned a = blist[3 * n], b = blist[3 * n + 1];
        double r[3];
        const double r2 = sep_r2(&x[3 * a], &x[3 * b], len, r);
        const double dist = sqrt(r2);
        const double ft = -ks * (dist - lbond) / dist;
        double fk[3];
        for (int k = 0; k < 3; k++) {
            fk[k] = ft * r[k];
            f[3 * a + k] += fk[k];
            f[3 * b + k] -= fk[k];
        }
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) {
                ret->pot_P[3 * k + kk] += fk[k] * r[kk];
                ret->pot_P_bond[3 * k + kk] += fk[k] * r[kk];
            }
        ret->epot += 0.5 * ks * (dist - lbond) * (dist - lbond);
        if (blengths) blengths[n] = dist;
    }
}

static void angle_common(const double *x, const double len[3], const unsigned *alist, unsigned na,
                         int type, double angle0, double kc, double *f, orc_ret *ret, double *angles,
                         int cossq)
{
    const double cCon = cos(ORC_PI - angle0);
    for (unsigned n = 0; n < na; n++) {
        if ((int)alist[4 * n + 3] != type) continue;
        const unsigned a = alist[4 * n], b = alist[4 * n + 1], c = alist[4 * n + 2];
        double d1[3], d2[3];
        for (int k = 0; k < 3; k++) {
            d1[k] = orc_wrap(x[3 * b + k] - x[3 * a + k], len[k]);
            d2[k] = orc_wrap(x[3 * c + k] - x[3 * b + k], len[k]);
        }
        const double c11 = dot3(d1, d1), c12 = dot3(d1, d2), c22 = dot3(d2, d2);
        const double cD = sqrt(c11 * c22);
        double fm, e, ang;
        if (cossq) {                                   /* source/sepmol.c:447-464 */
            const double cc = c12 / cD;
            fm = -kc * (cc - cCon);
            e = 0.5 * kc * (cc - cCon) * (cc - cCon);
            ang = ORC_PI - acos(cc);
        } else {                                       /* source/sepmol.c:496-512 */
            ang = ORC_PI - acos(c12 / cD);
            fm = -kc * (ang - angle0);
            e = 0.5 * kc * (ang - angle0) * (ang - angle0);
        }
        for (int k = 0; k < 3; k++) {
            const double f1 = fm * ((c12 / c11) * d1[k] - d2[k]) / cD;
            const double f2 = fm * (d1[k] - (c12 / c22) * d2[k]) / cD;
            f[3 * a + k] += f1;
            f[3 * b + k] += (-f1 - f2);
            f[3 * c + k] += f2;
        }
        ret->epot += e;
        if (angles) angles[n] = ang;
    }
}

void orc_angle_harmonic(const double *x, const double len[3], const unsigned *alist, unsigned na,
                        int type, double angle0, double k, double *f, orc_ret *ret, double *angles)
{
    angle_common(x, len, alist, na, type, angle0, k, f, ret, angles, 0);
}

void orc_angle_cossq(const double *x, const double len[3], const unsigned *alist, unsigned na,
                     int type, double angle0, double k, double *f, orc_ret *ret, double *angles)
{
    angle_common(x, len, alist, na, type, angle0, k, f, ret, angles, 1);
}

void orc_torsion_ryckaert(const double *x, const double len[3], const unsigned *dlist, unsigned nd,
                          int type, const double g[6], double *f, orc_ret *ret, double *dihedrals)
{
    for (unsigned n = 0; n < nd; n++) {
        if ((int)dlist[5 * n + 4] != type) continue;
        const unsigned a = dlist[5 * n], b = dlist[5 * n + 1], c = dlist[5 * n + 2], d = dlist[5 * n + 3];
        double d1[3], d2[3], d3[3];
        for (int k = 0; k < 3; k++) {
            d1[k] = orc_wrap(x[3 * b + k] - x[3 * a + k], len[k]);
            d2[k] = orc_wrap(x[3 * c + k] - x[3 * b + k], len[k]);
            d3[k] = orc_wrap(x[3 * d + k] - x[3 * c + k], len[k]);
        }
        const double c11 = dot3(d1, d1), c12 = dot3(d1, d2), c13 = dot3(d1, d3);
        const double c22 = dot3(d2, d2), c23 = dot3(d2, d3), c33 = dot3(d3, d3);
        const double cA = c13 * c22 - c12 * c23;
        const double cB1 = c11 * c22 - c12 * c12;
        const double cB2 = c22 * c33 - c23 * c23;
        const double cD = sqrt(cB1 * cB2);
        const double cc = cA / cD;
        const double fm = -(g[1] + (2. * g[2] + (3. * g[3] + (4. * g[4] + 5. * g[5] * cc) * cc) * cc) * cc);
        const double t1 = cA, t2 = c11 * c23 - c12 * c13, t3 = -cB1;
        const double t4 = cB2, t5 = c13 * c23 - c12 * c33, t6 = -cA;
        const double cR1 = c12 / c22, cR2 = c23 / c22;
        for (int k = 0; k < 3; k++) {
            const double f1 = fm * c22 * (t1 * d1[k] + t2 * d2[k] + t3 * d3[k]) / (cD * cB1);
            const double f2 = fm * c22 * (t4 * d1[k] + t5 * d2[k] + t6 * d3[k]) / (cD * cB2);
            f[3 * a + k] += f1;
            f[3 * b + k] += (-(1.0 + cR1) * f1 + cR2 * f2);
            f[3 * c + k] += (cR1 * f1 - (1.0 + cR2) * f2);
            f[3 * d + k] += f2;
        }
        ret->epot += g[0] + (g[1] + (g[2] + (g[3] + (g[4] + g[5] * cc) * cc) * cc) * cc) * cc;
        if (dihedrals) dihedrals[n] = ORC_PI - acos(cc);
    }
}

/* ---- thermostat + integrators ------------------------------------------------------------------ */
double orc_nosehoover(int n, const double *v, const double *m, double *f, double temp0,
                      double alpha, double tau, double dt)
{
    double ekin = 0.0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) ekin += v[3 * i + k] * v[3 * i + k] * m[i];
    ekin = 0.5 * ekin / n;
    const double temp = 0.666667 * ekin;               /* literal, source/sepintgr.c:159 */
    alpha = alpha + dt / (tau * tau) * (temp / temp0 - 1.0);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) f[3 * i + k] -= alpha * m[i] * v[3 * i + k];
    return alpha;
}

void orc_nosehoover_type(int n, const double *v, const double *m, const char *type, char which,
                         double *f, double Td, double alpha[3], double Q, double dt)
{
    int ntype = 0; double ekin = 0.0;
    for (int i = 0; i < n; i++)
        if (type[i] == which) {
            ntype++;
            for (int k = 0; k < 3; k++) ekin += v[3 * i + k] * v[3 * i + k] * m[i];
        }
    const double g = 3 * ntype - 3;
    const double tmp = alpha[0];
    alpha[0] = alpha[1];
    alpha[1] = alpha[2];
    alpha[2] = tmp + 2.0 * dt * (ekin - g * Td) / Q;
    for (int i = 0; i < n; i++)
        if (type[i] == which)
            for (int k = 0; k < 3; k++) f[3 * i + k] -= alpha[1] * v[3 * i + k] * m[i];
}

/* sep_periodic, source/sepintgr.c:18-40 */
static double periodic(double *x, const double *xn, int *cn, int *cr, const double len[3])
{
    double d2 = 0.0;
    for (int k = 0; k < 3; k++) {
        if (x[k] > len[k]) { x[k] -= len[k]; cn[k]++; cr[k]++; }
        else if (x[k] < 0.0) { x[k] += len[k]; cn[k]--; cr[k]--; }
        const double ri = (x[k] + cn[k] * len[k]) - xn[k];
        d2 += ri * ri;
    }
    return d2;
}

static int trigger(int n, const double *x, double *xn, int *cross_neighb, double skin, double max_dist2)
{
    if (!(sqrt(max_dist2) > skin * 0.5)) return 0;
    memcpy(xn, x, sizeof(double) * 3 * (size_t)n);
    memset(cross_neighb, 0, sizeof(int) * 3 * (size_t)n);
    return 1;
}

int orc_leapfrog(int n, double *x, double *v, const double *f, const double *m, double *a,
                 double *xn, int *cross_neighb, int *crossings, const double len[3], double dt,
                 double skin, double *max_dist2, orc_ret *ret)
{
    double sumekin = 0.0;
    for (int i = 0; i < n; i++) {
        double vh[3];
        for (int k = 0; k < 3; k++) {
            a[3 * i + k] = f[3 * i + k] / m[i];
            v[3 * i + k] += a[3 * i + k] * dt;
            x[3 * i + k] += v[3 * i + k] * dt;
            vh[k] = v[3 * i + k] - 0.5 * a[3 * i + k] * dt;
            sumekin += vh[k] * vh[k] * m[i];
        }
        const double d2 = periodic(&x[3 * i], &xn[3 * i], &cross_neighb[3 * i], &crossings[3 * i], len);
        if (d2 > *max_dist2) *max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->kin_P[3 * k + kk] += vh[k] * vh[kk] * m[i];
    }
    ret->ekin += 0.5 * sumekin;
    return trigger(n, x, xn, cross_neighb, skin, *max_dist2);
}

int orc_verlet_dpd(int n, double *x, double *v, const double *f, const double *m, double *a,
                   double *pv, double *pa, double *xn, int *cross_neighb, int *crossings,
                   const double len[3], double dt, double lambda, int stepnow, double skin,
                   double *max_dist2, orc_ret *ret)
{
    double sumekin = 0.0;
    for (int i = 0; i < n; i++) {
        double vv[3];
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            a[q] = f[q] / m[i];
            if (stepnow > 0) v[q] += 0.5 * dt * (a[q] + pa[q]);
            x[q] += v[q] * dt + 0.5 * dt * dt * a[q];
            pv[q] = v[q] + lambda * dt * a[q];
            pa[q] = a[q];
            vv[k] = v[q];
            sumekin += vv[k] * vv[k] * m[i];
        }
        const double d2 = periodic(&x[3 * i], &xn[3 * i], &cross_neighb[3 * i], &crossings[3 * i], len);
        if (d2 > *max_dist2) *max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->kin_P[3 * k + kk] += vv[k] * vv[kk] * m[i];
    }
    ret->ekin += 0.5 * sumekin;
    return trigger(n, x, xn, cross_neighb, skin, *max_dist2);
}

/* ---- DPD ------------------------------------------------------------------------------------------- */
static unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* pair-symmetric counter-based uniform in [0,1): replaces sep_rand() at source/sepprfrc.c:1069 */
double orc_dpd_uniform(unsigned long long seed, unsigned long long step, unsigned i, unsigned j)
{
    const unsigned lo = i < j ? i : j, hi = i < j ? j : i;
    if (seed == 0xFFFFFFFFFFFFFFFFULL) return 0.75;     /* every draw = a rand() that always returns 3*2^29 (tests/golden/rand_shim.c) */
    unsigned long long h = mix64(seed ^ (step * 0xD1342543DE82EF95ULL));
    h = mix64(h ^ (((unsigned long long)lo << 32) | hi));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

void orc_dpd_force_list(int n, const double *x, const double *pv, const char *type,
                        const double len[3], const int *pairs, long npairs, const char types[2],
                        double cf, double aij, double temp, double sigma, double dt,
                        unsigned long long seed, unsigned long long step, double *f, orc_ret *ret)
{
    (void)n;
    const double cf2 = cf * cf, isqrtdt = 1.0 / sqrt(dt);
    const double facchk = 2.0 * sqrt(3.0);
    const double gamma = sigma * sigma / (2.0 * temp);
    double epot = 0.0;
    for (long p = 0; p < npairs; p++) {
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (!type_match(type[i], type[j], types)) continue;
        double r[3];
        const double r2 = sep_r2(&x[3 * i], &x[3 * j], len, r);
        if (!(r2 < cf2)) continue;
        const double dij = sqrt(r2), w = 1.0 - dij;
        double rhat[3], dotrv = 0.0;
        for (int k = 0; k < 3; k++) {
            rhat[k] = r[k] / dij;
            dotrv += rhat[k] * (pv[3 * i + k] - pv[3 * j + k]);
        }
        const double xi = (orc_dpd_uniform(seed, step, (unsigned)i, (unsigned)j) - 0.5) * facchk;
        for (int k = 0; k < 3; k++) {
            const double fC = aij * w * rhat[k];
            const double fD = -gamma * w * w * dotrv * rhat[k];
            const double fR = sigma * w * rhat[k] * isqrtdt * xi;
            f[3 * i + k] += fC + fD + fR;
            f[3 * j + k] -= fC + fD + fR;
        }
        epot += 0.5 * aij * w * w;
    }
    ret->epot = epot;                                    /* assignment, source/sepprfrc.c:1132 */
}

/* ---- callers either side of the hot path ---------------------------------- */
static int nsubbox_of(double cf, double delta, double lbox) { return (int)(lbox / (cf + delta)); }   /* source/sepmisc.c:454-463 */

double orc_relax_temp(int n, double *v, const double *m, const char *type, char which, double Td, double tau, double dt)
{
    double ekin = 0.0;
    int ntype = 0;
    for (int i = 0; i < n; i++)
        if (type[i] == which) {
            ntype++;
            for (int k = 0; k < 3; k++) ekin += m[i] * (v[3 * i + k] * v[3 * i + k]);     /* :364-367 */
        }
    ekin = 0.5 * ekin;
    const double Ta = 2.0 * ekin / (3 * ntype);                                            /* :376 */
    const double fact = sqrt(1.0 + (dt / tau) * (Td / Ta - 1.0));                          /* :378 */
    for (int i = 0; i < n; i++)
        if (type[i] == which)
            for (int k = 0; k < 3; k++) v[3 * i + k] *= fact;
    /* sep_reset_momentum, :1173-1192 */
    double mom[3] = {0.0, 0.0, 0.0}, mass = 0.0;
    for (int i = 0; i < n; i++)
        if (type[i] == which) {
            for (int k = 0; k < 3; k++) mom[k] += v[3 * i + k] * m[i];
            mass += m[i];
        }
    for (int i = 0; i < n; i++)
        if (type[i] == which)
            for (int k = 0; k < 3; k++) v[3 * i + k] -= mom[k] / mass;
    return ekin;
}

void orc_force_x0(int n, const double *x, const double *x0, const char *type, char which, const double len[3], double *f)
{
    const double kspring = 500.0;                                                          /* sep_spring_x0, :168 */
    for (int i = 0; i < n; i++) {
        if (type[i] != which) continue;
        for (int k = 0; k < 3; k++) {
            const double r = orc_wrap(x0[3 * i + k] - x[3 * i + k], len[k]);               /* :655-657 */
            const double ft = -kspring;                                                    /* fun(r2,'f') */
            f[3 * i + k] -= ft * r;                                                        /* :660-662 */
        }
    }
}

int orc_compress_box(int n, double *x, double rhoD, double xi, double len[3], int nsub[3], double lsub[3],
                     double *volume, double cf, double skin, int list_mode)
{
    const double density = n / *volume;
    if (fabs(density - rhoD) < 1e-6) return 0;
    if (density > rhoD) xi = 1.0 / xi;
    for (int k = 0; k < 3; k++) len[k] *= xi;
    for (int i = 0; i < 3 * n; i++) x[i] *= xi;
    if (list_mode)
        for (int k = 0; k < 3; k++) {
            nsub[k] = nsubbox_of(cf, skin, len[k]);                                        /* with the skin, :1016 */
            lsub[k] = len[k] / nsub[k];
        }
    *volume = len[0] * len[1] * len[2];
    return 1;
}

void orc_berendsen(int n, double *x, double Pd, double beta, double p, double dt, int iso, double len[3],
                   int nsub[3], double lsub[3], double *volume, double cf, int list_mode)
{
    const double xi = 1 - beta * dt * (Pd - p);
    const double scale = pow(xi, 1.0 / 3.0);            /* positions move by xi^(1/3), lengths by xi (:897-901) */
    for (int k = iso ? 0 : 2; k < 3; k++) {
        len[k] *= xi;
        for (int i = 0; i < n; i++) x[3 * i + k] *= scale;
    }
    *volume = len[0] * len[1] * len[2];
    if (list_mode)
        for (int k = iso ? 0 : 2; k < 3; k++) {
            nsub[k] = nsubbox_of(cf, 0.0, len[k]);                                         /* without the skin, :906 */
            lsub[k] = len[k] / nsub[k];
        }
}

/* ---- stochastic integrators ------------------------------------------------------------------------------------ */
static int randn_cached = 0;
static double randn_spare = 0.0;

void orc_randn_reset(void) { randn_cached = 0; }

double orc_randn(void)
{
    if (randn_cached) { randn_cached = 0; return randn_spare; }
    double x1, x2, w;
    do {
        x1 = 2.0 * (rand() / (RAND_MAX + 1.0)) - 1.0;          /* sep_rand(), include/sepmisc.h:63 */
        x2 = 2.0 * (rand() / (RAND_MAX + 1.0)) - 1.0;
        w = x1 * x1 + x2 * x2;
    } while (w >= 1.0 || w == 0.0);
    w = sqrt((-2.0 * log(w)) / w);
    randn_spare = x2 * w;
    randn_cached = 1;
    return x1 * w;
}

int orc_fp(int n, double *x, double *v, const double *f, const double *m, const double *ldiff, double *xn,
           int *cross_neighb, int *crossings, const double len[3], double dt, double temp, double skin,
           double *max_dist2, orc_ret *ret)
{
    const double fac = sqrt(1.0 / 12.0);
    double sumekin = 0.0;
    for (int i = 0; i < n; i++) {
        double d2 = 0.0;
        const double im = 1.0 / m[i];
        const double fric = temp / ldiff[i];
        const double gaussfac = sqrt(24 * temp * fric / dt);
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            const double a = orc_randn() * fac * gaussfac;
            x[q] += dt * v[q];
            v[q] += im * dt * (f[q] - fric * v[q] + a);
            if (x[q] > len[k]) { x[q] -= len[k]; cross_neighb[q]++; crossings[q]++; }
            else if (x[q] < 0.0) { x[q] += len[k]; cross_neighb[q]--; crossings[q]--; }
            sumekin += v[q] * v[q] * m[i];
            const double ri = (x[q] + cross_neighb[q] * len[k]) - xn[q];
            d2 += ri * ri;
        }
        if (d2 > *max_dist2) *max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->kin_P[3 * k + kk] += v[3 * i + k] * v[3 * i + kk] * m[i];
    }
    ret->ekin += 0.5 * sumekin;
    return trigger(n, x, xn, cross_neighb, skin, *max_dist2);
}

int orc_langevin_gjf(int n, double *x, double *v, const double *f, const double *m, double *a, double *prevf, double *randn,
                     double *xn, int *cross_neighb, int *crossings, const double len[3], double dt, double temp,
                     double alpha, double skin, double *max_dist2, orc_ret *ret)
{
    const double cc = exp(-alpha * dt);
    double sumekin = 0.0;
    for (int i = 0; i < n; i++) {
        const double mass = m[i], imass = 1.0 / mass, imass2 = 0.5 * imass;
        const double fac = sqrt(temp * (1.0 - cc * cc));
        const double c = alpha * dt * imass2;
        const double ca = (1.0 - c) / (1.0 + c), cb = 1.0 / (1.0 + c);
        for (int k = 0; k < 3; k++) {
            const int q = 3 * i + k;
            v[q] = ca * v[q] + dt * imass2 * (ca * prevf[q] + f[q]) + cb * imass * randn[q];
            a[q] = f[q] * imass;
            prevf[q] = f[q];
            sumekin += v[q] * v[q] * mass;
            randn[q] = fac * orc_randn();
            x[q] += cb * dt * v[q] + cb * dt * dt * imass2 * f[q] + cb * dt * imass2 * randn[q];
        }
        const double d2 = periodic(&x[3 * i], &xn[3 * i], &cross_neighb[3 * i], &crossings[3 * i], len);
        if (d2 > *max_dist2) *max_dist2 = d2;
        for (int k = 0; k < 3; k++)
            for (int kk = 0; kk < 3; kk++) ret->kin_P[3 * k + kk] += v[3 * i + k] * v[3 * i + kk] * mass;
    }
    ret->ekin += 0.5 * sumekin;
    return trigger(n, x, xn, cross_neighb, skin, *max_dist2);
}
