"""Synthetic benchmark / scale-test systems of SURVEY.md section 8 (C1-C3), built from in-repo data only.

C1  simple-cubic Lennard-Jones lattice, positions (i + 0.5) a strictly inside [0, L)  (section 8d)
C2  butane: the 4000-atom / 1000-molecule unit cell recorded in tests/golden/butane_n4000.npz (prg2's system,
    evolved by the reference), tiled reps^3 times with atom and molecule indices offset (6^3 -> 864 000 atoms)
C3  water: the 648-atom / 216-molecule compressed cell of tests/golden/water_dense_n648.npz (the state prg3
    ends in, rho = 3.15), tiled reps^3 times (12^3 -> 1 119 744 atoms)

Everything here is host-side numpy; nothing touches the oracle or the reference.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# prg2 literals (reference prgs/prg2.c:24-27, 61-73)
BUTANE = dict(cf=2.5, dt=0.001, temp=4.0, tau=0.1, lbond=0.407, kbond=2074.0, angle=1.90, kangle=400.0,
              rb=(15.5000, 20.3050, -21.9170, -5.1150, 43.8340, -52.6070), types=b"CC")
# prg3 literals (reference prgs/prg3.c:25-36, 65-73)
WATER = dict(cf=2.9, cf_lj=2.5, dt=5.0e-4, temp=3.81, tau=0.01, lbond=0.316, kbond=68421.0, angle=1.97,
             kangle=490.0, types=b"OO")


def lj_lattice(ncell, rho):
    n = ncell ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / ncell
    g = (np.arange(ncell) + 0.5) * a
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    return np.ascontiguousarray(np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)), L


def lj_velocities(n, temp, seed):
    rng = np.random.default_rng(seed)
    v = rng.random((n, 3)) - 0.5
    v -= v.mean(axis=0)
    v *= np.sqrt(temp * 3.0 * (n - 1) / (v * v).sum())
    return np.ascontiguousarray(v)


def _tile_partner(tab, reps3, n0):
    """partner tables (bond[10], angle[10], dihed[20]): -1 stays, indices shift by the copy's atom offset"""
    out = np.empty((reps3 * len(tab), tab.shape[1]), dtype=np.int32)
    for k in range(reps3):
        blk = tab.astype(np.int32).copy()
        blk[blk >= 0] += k * n0
        out[k * len(tab):(k + 1) * len(tab)] = blk
    return out


def _tile_terms(lst, reps3, n0, natoms_cols):
    """topology rows (a, b, [c, [d]], type): atom columns shift, the type column does not"""
    if len(lst) == 0:
        return np.zeros((0, natoms_cols + 1), dtype=np.uint32)
    out = np.empty((reps3 * len(lst), lst.shape[1]), dtype=np.uint32)
    for k in range(reps3):
        blk = lst.astype(np.int64).copy()
        blk[:, :natoms_cols] += k * n0
        out[k * len(lst):(k + 1) * len(lst)] = blk
    return out


def tiled_molecular(fixture, reps):
    """Tile a recorded molecular unit cell reps^3 times (or rx x ry x rz for a tuple).  Returns a dict of numpy arrays."""
    rx, ry, rz = (reps, reps, reps) if np.isscalar(reps) else reps
    g = np.load(os.path.join(GOLDEN, fixture))
    L0 = np.atleast_1d(g["L"]).astype(float)
    if L0.size == 1:
        L0 = np.repeat(L0, 3)
    x0, v0 = g["x0"].copy(), g["v0"]
    n0 = len(x0)
    nmol0 = int(g["molindex"].max()) + 1
    # Molecules that straddle the unit cell's periodic boundary are made whole first (each atom placed at the
    # minimum image of its predecessor in the molecule), otherwise a tiled copy would tear them apart; moving an
    # atom by a lattice vector of the unit cell leaves the periodic crystal unchanged.
    mi = g["molindex"]
    same = np.zeros(n0, dtype=bool)
    same[1:] = (mi[1:] == mi[:-1]) & (mi[1:] >= 0)
    for i in range(1, n0):
        if same[i]:
            d = x0[i] - x0[i - 1]
            x0[i] = x0[i - 1] + (d - L0 * np.round(d / L0))
    reps3 = rx * ry * rz
    x = np.empty((reps3 * n0, 3)); v = np.empty((reps3 * n0, 3))
    mol = np.empty(reps3 * n0, dtype=np.int32)
    k = 0
    for iz in range(rz):
        for iy in range(ry):
            for ix in range(rx):
                x[k * n0:(k + 1) * n0] = x0 + np.array([ix, iy, iz]) * L0
                v[k * n0:(k + 1) * n0] = v0
                mol[k * n0:(k + 1) * n0] = g["molindex"] + k * nmol0
                k += 1

    def per_atom(key, default, dtype):
        a = g[key] if key in g.files else np.full(n0, default)
        return np.ascontiguousarray(np.tile(a, reps3), dtype=dtype)

    Lbig = L0 * np.array([rx, ry, rz])
    x -= Lbig * np.floor(x / Lbig)                  # back into [0, L)
    x[x >= Lbig] = 0.0                              # guard the rounding case x == L
    out = dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), L=Lbig, n=reps3 * n0, nmol=reps3 * nmol0,
               molindex=mol, type=per_atom("type", ord("C"), np.uint8), m=per_atom("m", 1.0, np.float64),
               z=per_atom("z", 0.0, np.float64),
               bond=_tile_partner(g["bond"], reps3, n0), angle=_tile_partner(g["angle"], reps3, n0),
               dihed=_tile_partner(g["dihed"], reps3, n0),
               blist=_tile_terms(g["blist"], reps3, n0, 2), alist=_tile_terms(g["alist"], reps3, n0, 3),
               dlist=_tile_terms(g["dlist"], reps3, n0, 4))
    return out


def butane(reps=6):
    return tiled_molecular("butane_n4000.npz", reps)


def water(reps=12):
    return tiled_molecular("water_dense_n648.npz", reps)


def write_top(w, path):
    """the .top text format sep_read_topology_file parses (reference source/sepmol.c:22-369)"""
    mol = w["molindex"]
    with open(path, "w") as fh:
        fh.write("[ bonds ]\n;mol a b type\n")
        b = w["blist"]
        np.savetxt(fh, np.column_stack([mol[b[:, 0]], b]), fmt="%d")
        if len(w["alist"]):
            fh.write("\n[ angles ]\n;mol a b c type\n")
            a = w["alist"]
            np.savetxt(fh, np.column_stack([mol[a[:, 0]], a]), fmt="%d")
        if len(w["dlist"]):
            fh.write("\n[ dihedrals ]\n;mol a b c d type\n")
            d = w["dlist"]
            np.savetxt(fh, np.column_stack([mol[d[:, 0]], d]), fmt="%d")
