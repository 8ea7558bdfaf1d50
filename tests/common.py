"""Shared test plumbing: bindings for the CPU oracle port (oracle/liboracle.so), for the compiled
reference (oracle/_ref/libsep_ref.so, when it has been built) and synthetic-system generators.

Only tests/ (and bench.py's cpu_baseline leg, __graft_entry__.smoke) may touch oracle/ -- it is the
checker, never the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from seplib_b200 import capi  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsep_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

ALL, EXCL_BONDED, EXCL_SAME_MOL = 1, 2, 3
POT_LJ, POT_LJ_SHIFT, POT_WCA, POT_LJ_PARAM = 0, 1, 2, 3


class OrcRet(C.Structure):
    _fields_ = [("epot", C.c_double), ("ecoul", C.c_double), ("ekin", C.c_double),
                ("pot_P", C.c_double * 9), ("kin_P", C.c_double * 9), ("pot_P_bond", C.c_double * 9)]


class OrcTopo(C.Structure):
    _fields_ = [("molindex", C.c_void_p), ("bond", C.c_void_p), ("angle", C.c_void_p), ("dihed", C.c_void_p)]


_oracle = None


def oracle():
    """liboracle.so, built on demand from oracle/sep_oracle.c (plain gcc, seconds)."""
    global _oracle
    if _oracle is not None:
        return _oracle
    src = os.path.join(ORACLE_DIR, "sep_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(ORACLE_SO, mode=C.RTLD_LOCAL)
    vp, dbl, i32 = C.c_void_p, C.c_double, C.c_int
    lib.orc_wrap.restype = dbl
    lib.orc_wrap.argtypes = [dbl, dbl]
    lib.orc_cell_geometry.argtypes = [vp, dbl, dbl, vp, vp]
    lib.orc_neighb_pairs.restype = C.c_long
    lib.orc_neighb_pairs.argtypes = [i32, vp, vp, vp, vp, dbl, C.c_uint, C.POINTER(OrcTopo), vp, C.c_long]
    lib.orc_neighb_pairs_n2.restype = C.c_long
    lib.orc_neighb_pairs_n2.argtypes = [i32, vp, vp, dbl, C.c_uint, C.POINTER(OrcTopo), vp, C.c_long]
    lib.orc_force_pairs_list.argtypes = [i32, vp, vp, vp, vp, C.c_long, C.c_char_p, dbl, i32, vp, vp, C.POINTER(OrcRet)]
    lib.orc_force_pairs_brute.argtypes = [i32, vp, vp, vp, C.c_char_p, dbl, i32, vp, C.c_uint, C.POINTER(OrcTopo), vp, C.POINTER(OrcRet)]
    lib.orc_coulomb_sf_list.argtypes = [i32, vp, vp, vp, vp, C.c_long, dbl, vp, C.POINTER(OrcRet)]
    lib.orc_coulomb_sf_brute.argtypes = [i32, vp, vp, vp, dbl, C.c_uint, C.POINTER(OrcTopo), vp, C.POINTER(OrcRet)]
    lib.orc_stretch_harmonic.argtypes = [vp, vp, vp, C.c_uint, i32, dbl, dbl, vp, C.POINTER(OrcRet), vp]
    lib.orc_angle_harmonic.argtypes = [vp, vp, vp, C.c_uint, i32, dbl, dbl, vp, C.POINTER(OrcRet), vp]
    lib.orc_angle_cossq.argtypes = [vp, vp, vp, C.c_uint, i32, dbl, dbl, vp, C.POINTER(OrcRet), vp]
    lib.orc_torsion_ryckaert.argtypes = [vp, vp, vp, C.c_uint, i32, vp, vp, C.POINTER(OrcRet), vp]
    lib.orc_nosehoover.restype = dbl
    lib.orc_nosehoover.argtypes = [i32, vp, vp, vp, dbl, dbl, dbl, dbl]
    lib.orc_nosehoover_type.argtypes = [i32, vp, vp, vp, C.c_char, vp, dbl, vp, dbl, dbl]
    lib.orc_leapfrog.restype = i32
    lib.orc_leapfrog.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, vp, C.POINTER(OrcRet)]
    lib.orc_verlet_dpd.restype = i32
    lib.orc_verlet_dpd.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, i32, dbl, vp, C.POINTER(OrcRet)]
    lib.orc_randn.restype = dbl
    lib.orc_fp.restype = i32
    lib.orc_fp.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, vp, C.POINTER(OrcRet)]
    lib.orc_langevin_gjf.restype = i32
    lib.orc_langevin_gjf.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, vp, C.POINTER(OrcRet)]
    lib.orc_relax_temp.restype = dbl
    lib.orc_relax_temp.argtypes = [i32, vp, vp, vp, C.c_char, dbl, dbl, dbl]
    lib.orc_force_x0.argtypes = [i32, vp, vp, vp, C.c_char, vp, vp]
    lib.orc_compress_box.restype = i32
    lib.orc_compress_box.argtypes = [i32, vp, dbl, dbl, vp, vp, vp, vp, dbl, dbl, i32]
    lib.orc_berendsen.argtypes = [i32, vp, dbl, dbl, dbl, dbl, i32, vp, vp, vp, vp, dbl, i32]
    lib.orc_dpd_uniform.restype = dbl
    lib.orc_dpd_uniform.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint, C.c_uint]
    lib.orc_dpd_force_list.argtypes = [i32, vp, vp, vp, vp, vp, C.c_long, C.c_char_p, dbl, dbl, dbl, dbl, dbl,
                                       C.c_ulonglong, C.c_ulonglong, vp, C.POINTER(OrcRet)]
    _oracle = lib
    return lib


def have_ref():
    return os.path.exists(REF_SO)


_ref = None


def ref():
    """The compiled reference (unmodified sources, -O2 -fno-fast-math); None if not built."""
    global _ref
    if _ref is None and have_ref():
        lib = C.CDLL(REF_SO, mode=C.RTLD_LOCAL)
        capi.declare_sep_api(lib)
        _ref = lib
    return _ref


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def dvec3(v):
    return np.ascontiguousarray(v, dtype=np.float64)


class Topo:
    """Per-atom partner tables as the reference's topology reader fills them."""

    def __init__(self, n):
        self.n = n
        self.molindex = np.full(n, -1, dtype=np.int32)
        self.bond = np.full((n, 10), -1, dtype=np.int32)
        self.angle = np.full((n, 10), -1, dtype=np.int32)
        self.dihed = np.full((n, 20), -1, dtype=np.int32)
        self.blist = np.zeros((0, 3), dtype=np.uint32)
        self.alist = np.zeros((0, 4), dtype=np.uint32)
        self.dlist = np.zeros((0, 5), dtype=np.uint32)

    @staticmethod
    def _add(row, val):
        k = int(np.argmax(row == -1))
        assert row[k] == -1, "partner table full"
        row[k] = val

    def add_bond(self, mol, a, b, t=0):
        self.blist = np.vstack([self.blist, np.array([[a, b, t]], dtype=np.uint32)])
        self.molindex[a] = mol
        self.molindex[b] = mol
        self._add(self.bond[a], b)
        self._add(self.bond[b], a)

    def add_angle(self, a, b, c, t=0):
        self.alist = np.vstack([self.alist, np.array([[a, b, c, t]], dtype=np.uint32)])
        for p, q in ((a, b), (a, c), (b, a), (b, c), (c, a), (c, b)):
            self._add(self.angle[p], q)

    def add_dihedral(self, a, b, c, d, t=0):
        self.dlist = np.vstack([self.dlist, np.array([[a, b, c, d, t]], dtype=np.uint32)])
        q = (a, b, c, d)
        for r in range(4):
            for s in range(4):
                if s != r:
                    self._add(self.dihed[q[r]], q[s])

    def c_struct(self):
        t = OrcTopo()
        t.molindex = self.molindex.ctypes.data
        t.bond = self.bond.ctypes.data
        t.angle = self.angle.ctypes.data
        t.dihed = self.dihed.ctypes.data
        return t


def chain_topology(nmol, nuau):
    """Linear chains of nuau atoms (butane: 4): bonds, angles and dihedrals along the chain, vectorised."""
    n = nmol * nuau
    t = Topo(n)
    base = np.arange(nmol, dtype=np.uint32)[:, None] * nuau
    t.molindex[:] = np.repeat(np.arange(nmol, dtype=np.int32), nuau)

    def rows(k):
        return (base + np.arange(nuau - k, dtype=np.uint32)[None, :]).reshape(-1)

    if nuau >= 2:
        a = rows(1)
        t.blist = np.stack([a, a + 1, np.zeros_like(a)], axis=1).astype(np.uint32)
    if nuau >= 3:
        a = rows(2)
        t.alist = np.stack([a, a + 1, a + 2, np.zeros_like(a)], axis=1).astype(np.uint32)
    if nuau >= 4:
        a = rows(3)
        t.dlist = np.stack([a, a + 1, a + 2, a + 3, np.zeros_like(a)], axis=1).astype(np.uint32)
    _fill_partner_tables(t)
    return t


def _fill_partner_tables(t):
    """Partner tables from the term lists, in the reference reader's order (source/sepmol.c:84-95, 208-211, 312-327)."""
    fill = np.zeros(t.n, dtype=np.int64)
    for a, b, _ in t.blist:
        t.bond[a, fill[a]] = b; fill[a] += 1
        t.bond[b, fill[b]] = a; fill[b] += 1
    fill[:] = 0
    for a, b, c, _ in t.alist:
        for p, q in ((a, b), (a, c), (b, a), (b, c), (c, a), (c, b)):
            t.angle[p, fill[p]] = q; fill[p] += 1
    fill[:] = 0
    for row in t.dlist:
        q = row[:4]
        for r in range(4):
            for s in range(4):
                if s != r:
                    t.dihed[q[r], fill[q[r]]] = q[s]; fill[q[r]] += 1


def lattice(ncell, rho, jitter=0.0, seed=1):
    """Simple cubic lattice of ncell^3 atoms strictly inside [0,L), optional jitter."""
    n = ncell ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / ncell
    g = (np.arange(ncell) + 0.5) * a
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape) * a
        pos = np.mod(pos, L)
        pos[pos >= L] = 0.0
    return np.ascontiguousarray(pos), L


def velocities(n, temp, seed=2, m=None):
    """Uniform(-1/2,1/2) velocities, drift removed, rescaled to temp as sep_set_vel_seed does (source/sepinit.c:148-171)."""
    rng = np.random.default_rng(seed)
    v = rng.random((n, 3)) - 0.5
    m = np.ones(n) if m is None else m
    v -= (v * m[:, None]).sum(axis=0) / m.sum()
    sekin = (v * v * m[:, None]).sum()
    v *= np.sqrt(3 * n * temp / sekin)
    return np.ascontiguousarray(v)


def pair_set(pairs):
    """Canonical sorted array of (min,max) rows (multiset preserved)."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    p = np.sort(p, axis=1)
    order = np.lexsort((p[:, 1], p[:, 0]))
    return p[order]


def oracle_pairs(x, L, cf, skin, opt=ALL, topo=None, grid_skin=0.25, max_pairs=None):
    orc = oracle()
    n = len(x)
    length = dvec3([L, L, L] if np.isscalar(L) else L)
    nsub = np.zeros(3, dtype=np.int32)
    lsub = np.zeros(3)
    orc.orc_cell_geometry(ptr(length), cf, grid_skin, ptr(nsub), ptr(lsub))
    if max_pairs is None:
        vol = float(np.prod(length))
        max_pairs = int(1.5 * n * (2.1 * (cf + skin) ** 3 * n / vol) + 4096)
    buf = np.empty((max_pairs, 2), dtype=np.int32)
    tp = topo.c_struct() if topo is not None else OrcTopo()
    xx = np.ascontiguousarray(x, dtype=np.float64)
    np_ = orc.orc_neighb_pairs(n, ptr(xx), ptr(length), ptr(nsub), ptr(lsub), cf + skin, opt, C.byref(tp), ptr(buf), max_pairs)
    assert np_ >= 0, f"oracle pair build failed ({np_})"
    return buf[:np_].copy()


def rel_force_err(f, fref):
    """max_i |df_i| / max(f_rms, 1)  -- SURVEY.md section 8c normalisation."""
    d = np.linalg.norm(f - fref, axis=1).max()
    frms = np.sqrt((fref * fref).sum(axis=1).mean())
    return d / max(frms, 1.0)


# ---------------------------------------------------------------------------------------------------
# driving the compiled reference (or our libsep.so: same ABI) through the seplib API
# ---------------------------------------------------------------------------------------------------
class ApiSystem:
    """A seplib system (sep_init + sep_sys_setup) on `lib`, with numpy views of the host seppart array."""

    def __init__(self, lib, x, L, cf, dt, update=capi.SEP_LLIST_NEIGHBLIST, v=None, types=None, m=None, z=None,
                 nneighb=3000):
        self.lib = lib
        self.n = len(x)
        length = [L] * 3 if np.isscalar(L) else list(L)
        self.atoms = lib.sep_init(self.n, nneighb)
        self.sys = lib.sep_sys_setup(length[0], length[1], length[2], cf, dt, self.n, update)
        self.ret = capi.SepRet()
        self.view = capi.atoms_view(self.atoms, self.n)
        self.view["x"][:] = x
        self.view["xn"][:] = 0.0
        self.view["pa"][:] = 0.0
        self.view["pv"][:] = 0.0
        if v is not None:
            self.view["v"][:] = v
        if types is not None:
            self.view["type"][:] = types
        if m is not None:
            self.view["m"][:] = m
        if z is not None:
            self.view["z"][:] = z
        self.closed = False

    @classmethod
    def from_xyz(cls, lib, xyz, top, cf, dt, update):
        self = cls.__new__(cls)
        self.lib = lib
        lbox = (C.c_double * 3)()
        npart = C.c_int()
        self.atoms = lib.sep_init_xyz(lbox, C.byref(npart), xyz.encode(), b"q")
        self.n = npart.value
        self.sys = lib.sep_sys_setup(lbox[0], lbox[1], lbox[2], cf, dt, self.n, update)
        self.ret = capi.SepRet()
        self.view = capi.atoms_view(self.atoms, self.n)
        self.view["xn"][:] = 0.0
        if top:
            lib.sep_read_topology_file(self.atoms, top.encode(), C.byref(self.sys), b"q")
        self.closed = False
        return self

    # shorthand
    @property
    def S(self):
        return C.byref(self.sys)

    @property
    def R(self):
        return C.byref(self.ret)

    def fun(self, name):
        return C.cast(getattr(self.lib, name), C.c_void_p)

    def neighb_pairs(self):
        """Half list as stored in ptr[i].neighb (reference only: our library does not fill host rows)."""
        out = []
        for i in range(self.n):
            row = self.atoms[i].neighb
            k = 0
            while row[k] != -1:
                out.append((i, row[k]))
                k += 1
        return np.array(out, dtype=np.int32).reshape(-1, 2)

    def topo(self):
        t = Topo(self.n)
        t.molindex[:] = self.view["molindex"]
        t.bond[:] = self.view["bond"]
        t.angle[:] = self.view["angle"]
        t.dihed[:] = self.view["dihed"]
        mp = self.sys.molptr.contents
        if mp.flag_bonds:
            t.blist = np.ctypeslib.as_array(mp.blist, shape=(mp.num_bonds, 3)).copy() if mp.num_bonds else t.blist
        if mp.flag_angles and mp.num_angles:
            t.alist = np.ctypeslib.as_array(mp.alist, shape=(mp.num_angles, 4)).copy()
        if mp.flag_dihedrals and mp.num_dihedrals:
            t.dlist = np.ctypeslib.as_array(mp.dlist, shape=(mp.num_dihedrals, 5)).copy()
        return t

    def ret_arrays(self):
        r = self.ret
        return dict(epot=r.epot, ecoul=r.ecoul, ekin=r.ekin, pot_P=np.array(r.pot_P).reshape(9).copy(),
                    kin_P=np.array(r.kin_P).reshape(9).copy(), pot_P_bond=np.array(r.pot_P_bond).reshape(9).copy())

    def close(self):
        if not self.closed:
            self.lib.sep_close(self.atoms, self.n)
            self.closed = True


# ---------------------------------------------------------------------------------------------------------------
# drivers of the "next" rows of SURVEY.md section 8f (box-changing callers, per-type relaxation, tether springs).
# The same code runs against the reference build (to record tests/golden/next_rows.npz) and against libsep.so.
# ---------------------------------------------------------------------------------------------------------------
def _final(s, rec):
    s.lib.sep_eval_mom(s.atoms, s.n)             # a host reader: brings atoms[] up to date in every sync mode
    rec["x"] = s.view["x"].copy(); rec["v"] = s.view["v"].copy()
    rec["length"] = np.array(s.sys.length[:]); rec["nsubbox"] = np.array(s.sys.nsubbox[:]); rec["volume"] = s.sys.volume
    return rec


def drive_compress(lib, x, v, L, steps=24, every=3, xi=0.99, rho_target=1.1):
    """prg3's box compression (reference prgs/prg3.c:76-79) on an LJ fluid in list mode; the cell grid goes from
    4 to 3 cells per side on the way."""
    s = ApiSystem(lib, x, L, 2.5, 0.005, v=v, nneighb=3000)
    alpha = C.c_double(0.1)
    fun = s.fun("sep_lj_shift")
    tr = []
    for n in range(steps):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", 2.5, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, 1.0, C.byref(alpha), 0.1, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        if n % every == 0:
            lib.sep_compress_box(s.atoms, rho_target, xi, s.S)
        tr.append((s.ret.epot, s.ret.ekin, s.sys.length[0], s.sys.nsubbox[0], s.sys.volume))
    rec = _final(s, {"traj": np.array(tr)})
    s.close()
    return rec


def drive_berendsen(lib, x, v, L, steps=30, iso=False, update=capi.SEP_BRUTE):
    """prg7's loop (reference prgs/prg7.c:37-56): LJ + Nose-Hoover + leapfrog + Berendsen barostat every step."""
    s = ApiSystem(lib, x, L, 2.5, 0.005, v=v, update=update, nneighb=3000)
    alpha = C.c_double(0.1)
    fun = s.fun("sep_lj_shift")
    tr = []
    for n in range(steps):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", 2.5, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, 0.8, C.byref(alpha), 0.1, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        (lib.sep_berendsen_iso if iso else lib.sep_berendsen)(s.atoms, 5.91, 0.1, s.R, s.S)
        tr.append((s.ret.epot, s.ret.ekin, s.ret.p, s.sys.length[2], s.sys.volume, s.sys.nsubbox[2]))
    rec = _final(s, {"traj": np.array(tr)})
    s.close()
    return rec


def drive_slit(lib, x, v, L, steps=30):
    """prg8's loop (reference prgs/prg8.c:41-66): fluid 'F' between tethered wall atoms 'W' -- three typed pair
    calls on one list, sep_force_x0 with sep_spring_x0, leapfrog, sep_relax_temp on the wall."""
    types = np.where(x[:, 2] < 2.2, ord("W"), ord("F")).astype(np.uint8)
    s = ApiSystem(lib, x, L, 2.5, 0.005, v=v, types=types, nneighb=3000)
    lib.sep_set_x0(s.atoms, s.n)
    lj, wca, spring = s.fun("sep_lj_shift"), s.fun("sep_wca"), s.fun("sep_spring_x0")
    rc_ww = 2.0 ** (1.0 / 6.0)
    tr = []
    for n in range(steps):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"FF", 2.5, lj, s.S, s.R, 1)
        lib.sep_force_pairs(s.atoms, b"WF", 2.5, lj, s.S, s.R, 1)
        lib.sep_force_pairs(s.atoms, b"WW", rc_ww, wca, s.S, s.R, 1)
        lib.sep_force_x0(s.atoms, b"W", spring, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        lib.sep_relax_temp(s.atoms, b"W", 1.4, 0.01, s.S)
        tr.append((s.ret.epot, s.ret.ekin))
    rec = _final(s, {"traj": np.array(tr), "types": types})
    s.close()
    return rec


# ---- the same three loops on the oracle (CPU): pins its restatements of the section-8f rows to the golden vectors ----
class OracleLoop:
    """State of one system stepped with oracle primitives the way the reference's API steps it."""

    def __init__(self, x, v, L, cf, dt, types=None, list_mode=True, skin=0.25):
        self.orc = oracle()
        self.n = len(x)
        self.x = np.ascontiguousarray(x, dtype=np.float64).copy(); self.v = np.ascontiguousarray(v, dtype=np.float64).copy()
        self.m = np.ones(self.n); self.a = np.zeros((self.n, 3)); self.f = np.zeros((self.n, 3))
        self.types = np.full(self.n, ord("A"), dtype=np.uint8) if types is None else np.ascontiguousarray(types, dtype=np.uint8)
        self.xn = np.zeros((self.n, 3)); self.cn = np.zeros((self.n, 3), dtype=np.int32); self.cr = np.zeros((self.n, 3), dtype=np.int32)
        self.len = dvec3([L] * 3); self.cf, self.dt, self.skin, self.list_mode = cf, dt, skin, list_mode
        self.nsub = np.zeros(3, dtype=np.int32); self.lsub = np.zeros(3)
        self.orc.orc_cell_geometry(ptr(self.len), cf, 0.25, ptr(self.nsub), ptr(self.lsub))     # sep_sys_setup: fixed 0.25
        self.volume = C.c_double(float(np.prod(self.len)))
        self.flag, self.pairs, self.maxd2 = 1, None, C.c_double(0.0)
        self.ret = OrcRet()

    def reset(self):
        self.ret = OrcRet(); self.f[:] = 0.0; self.maxd2.value = 0.0

    def pair_force(self, types, cf, pot):
        if self.list_mode:
            if self.flag:
                cap = int(40 * self.n * (self.cf + self.skin) ** 3 * self.n / self.volume.value / 8 + 65536)
                buf = np.empty((cap, 2), dtype=np.int32)
                k = self.orc.orc_neighb_pairs(self.n, ptr(self.x), ptr(self.len), ptr(self.nsub), ptr(self.lsub), self.cf + self.skin,
                                              ALL, C.byref(OrcTopo()), ptr(buf), cap)
                assert k >= 0
                self.pairs = np.ascontiguousarray(buf[:k]); self.flag = 0
            self.orc.orc_force_pairs_list(self.n, ptr(self.x), ptr(self.types), ptr(self.len), ptr(self.pairs), len(self.pairs),
                                          types, cf, pot, None, ptr(self.f), C.byref(self.ret))
        else:
            self.orc.orc_force_pairs_brute(self.n, ptr(self.x), ptr(self.types), ptr(self.len), types, cf, pot, None, ALL,
                                           C.byref(OrcTopo()), ptr(self.f), C.byref(self.ret))

    def leapfrog(self):
        self.flag |= self.orc.orc_leapfrog(self.n, ptr(self.x), ptr(self.v), ptr(self.f), ptr(self.m), ptr(self.a), ptr(self.xn),
                                           ptr(self.cn), ptr(self.cr), ptr(self.len), self.dt, self.skin, C.byref(self.maxd2), C.byref(self.ret))

    def pressure(self):
        kin, pot = np.array(self.ret.kin_P[:]), np.array(self.ret.pot_P[:])
        P = (kin + pot) / self.volume.value
        return (P[0] + P[4] + P[8]) / 3.0

    def record(self):
        return {"x": self.x.copy(), "v": self.v.copy(), "length": np.array(self.len), "nsubbox": self.nsub.copy(),
                "volume": self.volume.value}


def oracle_compress(x, v, L, steps=24, every=3, xi=0.99, rho_target=1.1):
    o = OracleLoop(x, v, L, 2.5, 0.005)
    alpha, tr = 0.1, []
    for n in range(steps):
        o.reset(); o.pair_force(b"AA", 2.5, POT_LJ_SHIFT)
        alpha = o.orc.orc_nosehoover(o.n, ptr(o.v), ptr(o.m), ptr(o.f), 1.0, alpha, 0.1, o.dt)
        o.leapfrog()
        if n % every == 0:
            o.orc.orc_compress_box(o.n, ptr(o.x), rho_target, xi, ptr(o.len), ptr(o.nsub), ptr(o.lsub), C.byref(o.volume), o.cf, o.skin, 1)
        tr.append((o.ret.epot, o.ret.ekin, o.len[0], o.nsub[0], o.volume.value))
    rec = o.record(); rec["traj"] = np.array(tr)
    return rec


def oracle_berendsen(x, v, L, steps=30, iso=False, list_mode=False):
    o = OracleLoop(x, v, L, 2.5, 0.005, list_mode=list_mode)
    alpha, tr = 0.1, []
    for n in range(steps):
        o.reset(); o.pair_force(b"AA", 2.5, POT_LJ_SHIFT)
        alpha = o.orc.orc_nosehoover(o.n, ptr(o.v), ptr(o.m), ptr(o.f), 0.8, alpha, 0.1, o.dt)
        o.leapfrog()
        p = o.pressure()
        o.orc.orc_berendsen(o.n, ptr(o.x), 5.91, 0.1, p, o.dt, 1 if iso else 0, ptr(o.len), ptr(o.nsub), ptr(o.lsub),
                            C.byref(o.volume), o.cf, 1 if list_mode else 0)
        tr.append((o.ret.epot, o.ret.ekin, p, o.len[2], o.volume.value, o.nsub[2]))
    rec = o.record(); rec["traj"] = np.array(tr)
    return rec


def oracle_slit(x, v, L, steps=30):
    types = np.where(x[:, 2] < 2.2, ord("W"), ord("F")).astype(np.uint8)
    o = OracleLoop(x, v, L, 2.5, 0.005, types=types)
    x0 = o.x.copy()
    tr = []
    for n in range(steps):
        o.reset()
        o.pair_force(b"FF", 2.5, POT_LJ_SHIFT); o.pair_force(b"WF", 2.5, POT_LJ_SHIFT); o.pair_force(b"WW", 2.0 ** (1.0 / 6.0), POT_WCA)
        o.orc.orc_force_x0(o.n, ptr(o.x), ptr(x0), ptr(o.types), b"W", ptr(o.len), ptr(o.f))
        o.leapfrog()
        o.orc.orc_relax_temp(o.n, ptr(o.v), ptr(o.m), ptr(o.types), b"W", 1.4, 0.01, o.dt)
        tr.append((o.ret.epot, o.ret.ekin))
    rec = o.record(); rec["traj"] = np.array(tr); rec["types"] = types
    return rec


# ---- stochastic integrators (section 8f rank 3): sep_fp (prg9's loop) and sep_langevinGJF --------------------------
_libc = C.CDLL(None)
STOCH_SEED = 20261017


def drive_stochastic(lib, x, v, L, which, steps=30):
    """prg9's loop (reference prgs/prg9.c:41-63): brute-force pair forces with cutoff SEP_WCACF, then sep_fp -- or
    sep_langevinGJF with friction 1.0 -- at T = 1.12.  The C library's rand() is seeded right before the loop; the
    library under test draws its Gaussian numbers from it."""
    cf = 1.12246204830937
    s = ApiSystem(lib, x, L, cf, 0.001, v=v, update=capi.SEP_BRUTE, nneighb=0)
    fun = s.fun("sep_lj_shift")
    _libc.srand(STOCH_SEED)
    tr = []
    for n in range(steps):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        if which == "fp":
            lib.sep_fp(s.atoms, 1.12, s.S, s.R)
        else:
            lib.sep_langevinGJF(s.atoms, 1.12, 1.0, s.S, s.R)
        tr.append((s.ret.epot, s.ret.ekin, s.sys.tnow))
    rec = _final(s, {"traj": np.array(tr)})
    s.close()
    return rec


def oracle_stochastic(x, v, L, which, steps=30):
    cf = 1.12246204830937
    o = OracleLoop(x, v, L, cf, 0.001, list_mode=False)
    ldiff = np.ones(o.n); prevf = np.zeros((o.n, 3)); randn = np.zeros((o.n, 3))
    o.orc.orc_randn_reset()
    _libc.srand(STOCH_SEED)
    tr, tnow = [], 0.0
    for n in range(steps):
        o.reset(); o.pair_force(b"AA", cf, POT_LJ_SHIFT)
        if which == "fp":
            o.orc.orc_fp(o.n, ptr(o.x), ptr(o.v), ptr(o.f), ptr(o.m), ptr(ldiff), ptr(o.xn), ptr(o.cn), ptr(o.cr), ptr(o.len),
                         o.dt, 1.12, o.skin, C.byref(o.maxd2), C.byref(o.ret))
        else:
            o.orc.orc_langevin_gjf(o.n, ptr(o.x), ptr(o.v), ptr(o.f), ptr(o.m), ptr(o.a), ptr(prevf), ptr(randn), ptr(o.xn), ptr(o.cn),
                                   ptr(o.cr), ptr(o.len), o.dt, 1.12, 1.0, o.skin, C.byref(o.maxd2), C.byref(o.ret))
            tnow += o.dt
        tr.append((o.ret.epot, o.ret.ekin, tnow))
    rec = o.record(); rec["traj"] = np.array(tr)
    return rec


def write_molecular_start_files(dirpath):
    """prg2 reads prg1.xyz / prg1.top, prg3 reads prg2.xyz / prg2.top (reference naming, prgs/prg2.c:30, prg3.c:39):
    written from the recorded butane / water states (tests/golden/*.npz) in the reference's own file formats."""
    for stem, fix, tname in (("prg1", "butane_n4000.npz", "C"), ("prg2", "water_n648.npz", None)):
        g = np.load(os.path.join(GOLDEN, fix))
        L = np.atleast_1d(g["L"]).astype(float)
        n = len(g["x0"])
        types = g["type"] if tname is None else np.full(n, ord(tname), dtype=np.uint8)
        m = g["m"] if "m" in g else np.ones(n)
        z = g["z"] if "z" in g else np.zeros(n)
        with open(os.path.join(str(dirpath), f"{stem}.xyz"), "w") as fh:
            fh.write(f"{n}\n{L[0]:.6f} {L[1]:.6f} {L[2]:.6f}\n")
            for i in range(n):
                fh.write("%c %.15f %.15f %.15f %.15f %.15f %.15f %.15f %.15f\n" % (
                    chr(types[i]), *g["x0"][i], *g["v0"][i], m[i], z[i]))
        with open(os.path.join(str(dirpath), f"{stem}.top"), "w") as fh:
            fh.write("[ bonds ]\n;generated\n")
            for (a, b, t) in g["blist"]:
                fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
            fh.write("\n[ angles ]\n;generated\n")
            for (a, b, c, t) in g["alist"]:
                fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
            if len(g["dlist"]):
                fh.write("\n[ dihedrals ]\n;generated\n")
                for (a, b, c, d, t) in g["dlist"]:
                    fh.write(f"{g['molindex'][a]} {a} {b} {c} {d} {t}\n")


def drive_nvt_1000(lib, x, v, L, nsteps=1000, every=100, cf=2.5, dt=0.005, temp=1.0, tau=0.1, sync_lazy=False):
    """prg1-style NVT loop through the sep_* API (reference prgs/prg1.c:52-83 with the metric's rc = 2.5) on `lib` -- the
    compiled reference when the fixture is recorded, libsep.so on the GPU when it is checked.  Every `every` steps:
    [step, epot/N, ekin/N, T, p, nupdate_neighb] exactly as prg1 derives them (sep_pressure_tensor, ekin*2/3N)."""
    s = ApiSystem(lib, x, L, cf, dt, v=v, nneighb=3000 if not hasattr(lib, "sep_gpu_sync") else 0)
    alpha = C.c_double(0.1)
    fun = s.fun("sep_lj_shift")
    rows = []
    n = s.n
    for step in range(nsteps + 1):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), tau, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        if step % every == 0:
            lib.sep_pressure_tensor(s.R, s.S)
            rows.append([step, s.ret.epot / n, s.ret.ekin / n, s.ret.ekin * 2.0 / (3.0 * n), s.ret.p, s.sys.nupdate_neighb])
    s.close()
    return np.array(rows)
