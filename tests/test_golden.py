"""Golden vectors produced by the REFERENCE (tests/golden/make_golden.py, run against
oracle/_ref/libsep_ref.so) checked two ways:

  * not-gpu tests pin the CPU oracle port (oracle/sep_oracle.c) to them -- the oracle is built with the
    same -O2 -fno-fast-math -ffp-contract=off, so agreement is expected to the last bit for per-atom
    quantities (asserted to 1e-13 relative to stay robust against libm differences between hosts);
  * gpu tests check the CUDA path (through the sepgpu C ABI) against the same vectors: neighbour pair
    sets bit-exact, forces max|df|/max(f_rms,1) <= 1e-10, sums rel <= 1e-10.
"""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

FT = 1e-10
EXACT = 1e-13


def load(name):
    return np.load(os.path.join(cm.GOLDEN, name))


def relerr(a, b, floor=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), floor)


def Lvec(g):
    L = np.atleast_1d(g["L"]).astype(np.float64)
    return cm.dvec3(L if L.size == 3 else [float(L[0])] * 3)


def topo_from(g, n):
    t = cm.Topo(n)
    t.molindex[:] = g["molindex"]; t.bond[:] = g["bond"]; t.angle[:] = g["angle"]; t.dihed[:] = g["dihed"]
    t.blist = np.ascontiguousarray(g["blist"], dtype=np.uint32)
    t.alist = np.ascontiguousarray(g["alist"], dtype=np.uint32)
    t.dlist = np.ascontiguousarray(g["dlist"], dtype=np.uint32)
    return t


# ====================================================================================================
# CPU: oracle port vs reference golden vectors
# ====================================================================================================
def orc_list_force(x, types, length, pairs, tsel, cf, pot, par=None):
    orc = cm.oracle()
    n = len(x)
    f = np.zeros((n, 3)); ret = cm.OrcRet()
    pp = np.ascontiguousarray(pairs, dtype=np.int32)
    pa = cm.dvec3(par) if par is not None else None
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, cf, pot,
                             cm.ptr(pa) if pa is not None else None, cm.ptr(f), C.byref(ret))
    return f, ret


def test_oracle_lj_golden():
    g = load("lj_n1000.npz")
    x = np.ascontiguousarray(g["x0"]); v = np.ascontiguousarray(g["v0"])
    n = len(x); L = float(g["L"]); cf, skin, dt = float(g["cf"]), float(g["skin"]), float(g["dt"])
    length = Lvec(g)
    types = np.full(n, ord("A"), dtype=np.uint8)
    raw = cm.oracle_pairs(x, L, cf, skin)                 # reference visiting order (summation order!)
    assert np.array_equal(cm.pair_set(raw), g["pairs"])
    f, ret = orc_list_force(x, types, length, raw, b"AA", cf, cm.POT_LJ_SHIFT)
    assert relerr(f, g["f_pairs"]) <= EXACT
    assert abs(ret.epot - float(g["epot"])) <= EXACT * abs(float(g["epot"]))
    assert relerr(np.array(ret.pot_P[:]), g["pot_P"]) <= EXACT
    # other pair functions
    f2, r2 = orc_list_force(x, types, length, raw, b"AA", 2.5, cm.POT_LJ)
    assert relerr(f2, g["f_lj"]) <= EXACT and abs(r2.epot - float(g["epot_lj"])) <= EXACT * abs(float(g["epot_lj"]))
    f3, r3 = orc_list_force(x, types, length, raw, b"AA", 2.0 ** (1.0 / 6.0), cm.POT_WCA)
    assert relerr(f3, g["f_wca"]) <= EXACT and abs(r3.epot - float(g["epot_wca"])) <= EXACT * abs(float(g["epot_wca"]))
    par = g["ljparam"]
    f4, r4 = orc_list_force(x, types, length, raw, b"AA", par[0], cm.POT_LJ_PARAM, par)
    assert relerr(f4, g["f_ljparam"]) <= EXACT and abs(r4.epot - float(g["epot_ljparam"])) <= EXACT * abs(float(g["epot_ljparam"]))
    # thermostat + leapfrog
    orc = cm.oracle()
    m = np.ones(n)
    alpha = orc.orc_nosehoover(n, cm.ptr(v), cm.ptr(m), cm.ptr(f), float(g["temp"]), float(g["alpha0"]), float(g["tau"]), dt)
    assert abs(alpha - float(g["alpha1"])) <= EXACT * abs(alpha)
    assert relerr(f, g["f_nh"]) <= EXACT
    xn = np.ascontiguousarray(g["xn0"]); cn = np.ascontiguousarray(g["cn0"], dtype=np.int32)
    cr = np.ascontiguousarray(g["cr0"], dtype=np.int32); a = np.zeros((n, 3)); md2 = C.c_double(0.0)
    flag = orc.orc_leapfrog(n, cm.ptr(x), cm.ptr(v), cm.ptr(f), cm.ptr(m), cm.ptr(a), cm.ptr(xn), cm.ptr(cn), cm.ptr(cr),
                            cm.ptr(length), dt, skin, C.byref(md2), C.byref(ret))
    assert np.array_equal(x, g["x1"]) and np.array_equal(v, g["v1"])
    assert np.array_equal(cr, g["cr1"]) and np.array_equal(cn, g["cn1"]) and np.array_equal(xn, g["xn1"])
    assert flag == int(g["neighb_flag1"]) and md2.value == float(g["max_dist2"])
    assert abs(ret.ekin - float(g["ekin"])) <= EXACT * float(g["ekin"])
    assert relerr(np.array(ret.kin_P[:]), g["kin_P"]) <= EXACT


def test_oracle_butane_golden():
    g = load("butane_n4000.npz")
    x = np.ascontiguousarray(g["x0"]); n = len(x)
    length = Lvec(g); cf = float(g["cf"])
    t = topo_from(g, n)
    types = np.full(n, ord("C"), dtype=np.uint8)
    orc = cm.oracle()
    raw = cm.oracle_pairs(x, length, cf, 0.25, opt=cm.EXCL_SAME_MOL, topo=t)
    assert np.array_equal(cm.pair_set(raw), g["pairs_same_mol"])
    raw_nb = cm.oracle_pairs(x, length, cf, 0.25, opt=cm.EXCL_BONDED, topo=t)
    assert np.array_equal(cm.pair_set(raw_nb), g["pairs_nonbonded"])
    f, ret = orc_list_force(x, types, length, raw, b"CC", cf, cm.POT_LJ_SHIFT)
    assert relerr(f, g["f_lj"]) <= EXACT and abs(ret.epot - float(g["epot_lj"])) <= EXACT * abs(float(g["epot_lj"]))
    bl = np.zeros(len(t.blist)); an = np.zeros(len(t.alist)); di = np.zeros(len(t.dlist))
    orc.orc_stretch_harmonic(cm.ptr(x), cm.ptr(length), cm.ptr(t.blist), len(t.blist), 0, 0.407, 2074.0, cm.ptr(f), C.byref(ret), cm.ptr(bl))
    assert relerr(f, g["f_bond"]) <= EXACT and abs(ret.epot - float(g["epot_bond"])) <= EXACT * abs(float(g["epot_bond"]))
    assert relerr(np.array(ret.pot_P_bond[:]), g["pot_P_bond"]) <= EXACT and np.array_equal(bl, g["blengths"])
    orc.orc_angle_harmonic(cm.ptr(x), cm.ptr(length), cm.ptr(t.alist), len(t.alist), 0, 1.90, 400.0, cm.ptr(f), C.byref(ret), cm.ptr(an))
    assert relerr(f, g["f_angle"]) <= EXACT and abs(ret.epot - float(g["epot_angle"])) <= EXACT * abs(float(g["epot_angle"]))
    assert relerr(an, g["angles"]) <= EXACT
    rb = cm.dvec3(g["rb"])
    orc.orc_torsion_ryckaert(cm.ptr(x), cm.ptr(length), cm.ptr(t.dlist), len(t.dlist), 0, cm.ptr(rb), cm.ptr(f), C.byref(ret), cm.ptr(di))
    assert relerr(f, g["f_torsion"]) <= EXACT and abs(ret.epot - float(g["epot_torsion"])) <= EXACT * abs(float(g["epot_torsion"]))
    assert relerr(di, g["dihedrals"]) <= EXACT
    f2 = np.zeros((n, 3)); r2 = cm.OrcRet()
    orc.orc_angle_cossq(cm.ptr(x), cm.ptr(length), cm.ptr(t.alist), len(t.alist), 0, 1.90, 400.0, cm.ptr(f2), C.byref(r2), cm.ptr(an))
    assert relerr(f2, g["f_cossq"]) <= EXACT and abs(r2.epot - float(g["epot_cossq"])) <= EXACT * abs(float(g["epot_cossq"]))


@pytest.mark.parametrize("fixture", ["water_n648.npz", "water_dense_n648.npz"])
def test_oracle_water_golden(fixture):
    g = load(fixture)
    x = np.ascontiguousarray(g["x0"]); n = len(x)
    length = Lvec(g); cf = float(g["cf"])
    t = topo_from(g, n)
    types = np.ascontiguousarray(g["type"], dtype=np.uint8); z = np.ascontiguousarray(g["z"])
    orc = cm.oracle()
    f = np.zeros((n, 3)); ret = cm.OrcRet(); tp = t.c_struct()
    orc.orc_force_pairs_brute(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), b"OO", 2.5, cm.POT_LJ_SHIFT, None,
                              cm.EXCL_SAME_MOL, C.byref(tp), cm.ptr(f), C.byref(ret))
    assert relerr(f, g["f_lj"]) <= EXACT and abs(ret.epot - float(g["epot_lj"])) <= EXACT * abs(float(g["epot_lj"]))
    orc.orc_stretch_harmonic(cm.ptr(x), cm.ptr(length), cm.ptr(t.blist), len(t.blist), 0, 0.316, 68421.0, cm.ptr(f), C.byref(ret), None)
    assert relerr(f, g["f_bond"]) <= EXACT
    orc.orc_angle_cossq(cm.ptr(x), cm.ptr(length), cm.ptr(t.alist), len(t.alist), 0, 1.97, 490.0, cm.ptr(f), C.byref(ret), None)
    assert relerr(f, g["f_angle"]) <= EXACT
    orc.orc_coulomb_sf_brute(n, cm.ptr(x), cm.ptr(z), cm.ptr(length), cf, cm.EXCL_SAME_MOL, C.byref(tp), cm.ptr(f), C.byref(ret))
    assert relerr(f, g["f_coul"]) <= EXACT
    assert abs(ret.epot - float(g["epot_coul"])) <= EXACT * abs(float(g["epot_coul"]))
    assert abs(ret.ecoul - float(g["ecoul"])) <= EXACT * abs(float(g["ecoul"]))
    assert relerr(np.array(ret.pot_P[:]), g["pot_P_total"]) <= EXACT
    f2 = np.zeros((n, 3)); r2 = cm.OrcRet()
    orc.orc_coulomb_sf_brute(n, cm.ptr(x), cm.ptr(z), cm.ptr(length), cf, cm.EXCL_BONDED, C.byref(tp), cm.ptr(f2), C.byref(r2))
    assert relerr(f2, g["f_coul_bonded"]) <= EXACT and abs(r2.ecoul - float(g["ecoul_bonded"])) <= EXACT * abs(float(g["ecoul_bonded"]))


def test_oracle_verlet_dpd_golden():
    g = load("dpd_n512.npz")
    x = np.ascontiguousarray(g["x0"]); v = np.ascontiguousarray(g["v0"]); n = len(x)
    L = float(g["L"]); length = cm.dvec3([L] * 3); dt = float(g["dt"])
    orc = cm.oracle()
    m = np.ones(n); a = np.zeros((n, 3)); pv = np.zeros((n, 3)); pa = np.zeros((n, 3)); xn = np.zeros((n, 3))
    cn = np.zeros((n, 3), dtype=np.int32); cr = np.zeros((n, 3), dtype=np.int32)
    for step in range(2):
        f = np.ascontiguousarray(g[f"f{step}"]); ret = cm.OrcRet(); md2 = C.c_double(0.0)
        flag = orc.orc_verlet_dpd(n, cm.ptr(x), cm.ptr(v), cm.ptr(f), cm.ptr(m), cm.ptr(a), cm.ptr(pv), cm.ptr(pa), cm.ptr(xn),
                                  cm.ptr(cn), cm.ptr(cr), cm.ptr(length), dt, 0.5, step, 0.25, C.byref(md2), C.byref(ret))
        assert np.array_equal(x, g[f"x{step + 1}"]) and np.array_equal(v, g[f"v{step + 1}"])
        assert np.array_equal(pv, g[f"pv{step + 1}"]) and np.array_equal(pa, g[f"pa{step + 1}"])
        assert np.array_equal(cr, g[f"cr{step + 1}"]) and flag == int(g[f"flag{step + 1}"])
        assert abs(ret.ekin - float(g[f"ekin{step + 1}"])) <= EXACT * ret.ekin


SEED_FIXED = 0xFFFFFFFFFFFFFFFF          # SEPGPU_DPD_SEED_FIXED: every pair draws u = 0.75 (include/sepgpu.h)


def _dpd_pairs(x, L, cf, tag):
    n = len(x); length = cm.dvec3([L] * 3)
    if tag == "list":
        return np.ascontiguousarray(cm.oracle_pairs(x, L, cf, 0.25), dtype=np.int32)
    buf = np.empty((n * n, 2), dtype=np.int32); tp = cm.OrcTopo()
    k = cm.oracle().orc_neighb_pairs_n2(n, cm.ptr(x), cm.ptr(length), cf, cm.ALL, C.byref(tp), cm.ptr(buf), n * n)
    return np.ascontiguousarray(buf[:k])


@pytest.mark.parametrize("tag", ["list", "brute"])
def test_oracle_dpd_force_golden(tag):
    """sep_force_dpd as the REFERENCE computes it when its rand() is interposed to a constant (tests/golden/rand_shim.c):
    conservative + dissipative + random terms of sep_dpdforce_neighb / _brute (source/sepprfrc.c:1007-1231)."""
    g = load("dpd_force_n512.npz")
    x = np.ascontiguousarray(g[tag + "_x"]); pv = np.ascontiguousarray(g[tag + "_pv"]); L = float(g[tag + "_L"]); n = len(x)
    types = np.full(n, ord("A"), dtype=np.uint8); length = cm.dvec3([L] * 3)
    pairs = _dpd_pairs(x, L, float(g["cf"]), tag)
    f = np.zeros((n, 3)); ret = cm.OrcRet()
    cm.oracle().orc_dpd_force_list(n, cm.ptr(x), cm.ptr(pv), cm.ptr(types), cm.ptr(length), cm.ptr(pairs), len(pairs), b"AA",
                                   float(g["cf"]), float(g["aij"]), float(g["temp"]), float(g["sigma"]), float(g["dt"]),
                                   SEED_FIXED, 0, cm.ptr(f), C.byref(ret))
    assert cm.rel_force_err(f, g[tag + "_f"]) <= 1e-12
    assert abs(ret.epot - float(g[tag + "_epot"])) <= 1e-12 * abs(ret.epot)


# ====================================================================================================
# GPU: CUDA path vs reference golden vectors
# ====================================================================================================
def gpu_sys(g, n, neighb_update=capi.SEP_LLIST_NEIGHBLIST, cf=None, skin=0.25):
    L = Lvec(g)
    return capi.make_sys(L, float(g["cf"]) if cf is None else cf, float(g["dt"]), neighb_update=neighb_update, skin=skin)


def gpu_put_topology(s, t):
    s.put(capi.F_MOLINDEX, t.molindex)
    s.put(capi.F_BOND, t.bond); s.put(capi.F_ANGLE, t.angle); s.put(capi.F_DIHED, t.dihed)
    s.set_topology(t.blist, t.alist, t.dlist)


@pytest.mark.gpu
def test_gpu_lj_golden():
    g = load("lj_n1000.npz")
    n = len(g["x0"])
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_XN, g["xn0"])
    s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    sys_ = gpu_sys(g, n)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    assert np.array_equal(cm.pair_set(s.pairs()), g["pairs"])
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_pairs"]) <= FT
    sc = s.scalars()
    assert abs(sc.epot - float(g["epot"])) <= FT * abs(float(g["epot"]))
    assert relerr(np.array(sc.pot_P[:]), g["pot_P"]) <= FT
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    s.call("sepgpu_nosehoover", C.byref(sys_), float(g["temp"]), 0, float(g["tau"]))
    assert cm.rel_force_err(s.get(capi.F_F), g["f_nh"]) <= FT          # flushes the deferred thermostat term
    # get(F) applied the pending update; re-arm it through a second identical sequence on a fresh context
    s.close()
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_XN, g["xn0"])
    s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    s.call("sepgpu_nosehoover", C.byref(sys_), float(g["temp"]), 0, float(g["tau"]))
    s.call("sepgpu_leapfrog", C.byref(sys_))
    sc = s.scalars()
    assert abs(sc.alpha[0] - float(g["alpha1"])) <= 1e-12 * abs(float(g["alpha1"]))
    assert np.abs(s.get(capi.F_X) - g["x1"]).max() <= 1e-12
    assert np.abs(s.get(capi.F_V) - g["v1"]).max() <= 1e-12
    assert cm.rel_force_err(s.get(capi.F_F), g["f_nh"]) <= FT
    assert np.array_equal(s.get(capi.F_CROSSINGS), g["cr1"])
    assert np.array_equal(s.get(capi.F_CROSS_NEIGHB), g["cn1"])
    assert sc.neighb_flag == int(g["neighb_flag1"])
    assert abs(sc.max_dist2 - float(g["max_dist2"])) <= 1e-12 * float(g["max_dist2"])
    assert abs(sc.ekin - float(g["ekin"])) <= FT * float(g["ekin"])
    assert relerr(np.array(sc.kin_P[:]), g["kin_P"]) <= FT
    # the other pair functions
    for kind, key, rc in (("lj", "lj", 2.5), ("wca", "wca", 2.0 ** (1.0 / 6.0))):
        s2 = capi.System(n); s2.put(capi.F_X, g["x0"])
        s2.call("sepgpu_reset_ret"); s2.call("sepgpu_reset_force")
        pp = capi.lj_param(rc, kind=kind)
        s2.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(pp), 1, 1)
        assert cm.rel_force_err(s2.get(capi.F_F), g[f"f_{key}"]) <= FT
        assert abs(s2.scalars().epot - float(g[f"epot_{key}"])) <= FT * abs(float(g[f"epot_{key}"]))
        s2.close()
    par = g["ljparam"]
    s2 = capi.System(n); s2.put(capi.F_X, g["x0"])
    s2.call("sepgpu_reset_ret"); s2.call("sepgpu_reset_force")
    pp = capi.lj_param(par[0], eps=par[1], sigma=par[2], aw=par[3], kind="param")
    s2.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(pp), 1, 0)
    assert cm.rel_force_err(s2.get(capi.F_F), g["f_ljparam"]) <= FT
    assert abs(s2.scalars().epot - float(g["epot_ljparam"])) <= FT * abs(float(g["epot_ljparam"]))
    s2.close(); s.close()


@pytest.mark.gpu
def test_gpu_lj_trajectory_golden():
    """41 further reference steps: epot, ekin, pressure, alpha and the rebuild count follow the reference
    (trajectory tolerance grows with the step count: chaotic amplification of rounding differences)."""
    g = load("lj_n1000.npz")
    n = len(g["x0"])
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_XN, g["xn0"])
    s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    sys_ = gpu_sys(g, n)
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    vol = float(g["L"]) ** 3
    traj = g["traj"]
    nup0 = None
    for step in range(41):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
        s.call("sepgpu_nosehoover", C.byref(sys_), float(g["temp"]), 0, float(g["tau"]))
        s.call("sepgpu_leapfrog", C.byref(sys_))
        if step == 0:
            continue
        sc = s.scalars()
        ref = traj[step - 1]
        tol = 1e-9 * (step + 1)
        assert abs(sc.epot - ref[0]) <= tol * abs(ref[0])
        assert abs(sc.ekin - ref[1]) <= tol * abs(ref[1])
        pr = (sum(sc.kin_P[k] + sc.pot_P[k] for k in (0, 4, 8)) / vol) / 3.0
        assert abs(pr - ref[2]) <= tol * max(abs(ref[2]), 1.0)
        assert abs(sc.alpha[0] - ref[3]) <= tol * max(abs(ref[3]), 1e-2)
    assert np.abs(s.get(capi.F_X) - g["x41"]).max() <= 1e-6
    s.close()


@pytest.mark.gpu
def test_gpu_butane_golden():
    g = load("butane_n4000.npz")
    n = len(g["x0"])
    t = topo_from(g, n)
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_TYPE, np.full(n, ord("C"), dtype=np.uint8))
    s.put(capi.F_XN, g["xn0"]); s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    gpu_put_topology(s, t)
    sys_ = gpu_sys(g, n)
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
    assert np.array_equal(cm.pair_set(s.pairs()), g["pairs_same_mol"])
    s.call("sepgpu_force_lj", C.byref(sys_), b"CC", C.byref(p), cm.EXCL_SAME_MOL, 1)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_lj"]) <= FT
    assert abs(s.scalars().epot - float(g["epot_lj"])) <= FT * abs(float(g["epot_lj"]))
    s.call("sepgpu_stretch_harmonic", C.byref(sys_), 0, 0.407, 2074.0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_bond"]) <= FT
    sc = s.scalars()
    assert abs(sc.epot - float(g["epot_bond"])) <= FT * abs(float(g["epot_bond"]))
    assert relerr(np.array(sc.pot_P_bond[:]), g["pot_P_bond"]) <= FT
    assert relerr(np.array(sc.pot_P[:]), g["pot_P_bond_total"]) <= FT
    s.call("sepgpu_angle_harmonic", C.byref(sys_), 0, 1.90, 400.0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_angle"]) <= FT
    assert abs(s.scalars().epot - float(g["epot_angle"])) <= FT * abs(float(g["epot_angle"]))
    rb = (C.c_double * 6)(*g["rb"])
    s.call("sepgpu_torsion_ryckaert", C.byref(sys_), 0, rb)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_torsion"]) <= FT
    assert abs(s.scalars().epot - float(g["epot_torsion"])) <= FT * abs(float(g["epot_torsion"]))
    bl, an, di = s.bonded_values()
    assert relerr(bl, g["blengths"]) <= 1e-13 and relerr(an, g["angles"]) <= 1e-12 and relerr(di, g["dihedrals"]) <= 1e-12
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    s.call("sepgpu_nosehoover", C.byref(sys_), float(g["temp"]), 0, 0.1)
    s.call("sepgpu_leapfrog", C.byref(sys_))
    sc = s.scalars()
    assert np.abs(s.get(capi.F_X) - g["x1"]).max() <= 1e-11 and np.abs(s.get(capi.F_V) - g["v1"]).max() <= 1e-10
    assert abs(sc.alpha[0] - float(g["alpha1"])) <= 1e-12 * abs(float(g["alpha1"]))
    assert abs(sc.ekin - float(g["ekin"])) <= FT * float(g["ekin"])
    # bonded-partner exclusion list and the cos^2 angle form
    s.put(capi.F_X, g["x0"])
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_BONDED)
    assert np.array_equal(cm.pair_set(s.pairs()), g["pairs_nonbonded"])
    s.call("sepgpu_force_lj", C.byref(sys_), b"CC", C.byref(p), cm.EXCL_BONDED, 1)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_lj_nonbonded"]) <= FT
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_angle_cossq", C.byref(sys_), 0, 1.90, 400.0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_cossq"]) <= FT
    assert abs(s.scalars().epot - float(g["epot_cossq"])) <= FT * abs(float(g["epot_cossq"]))
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["water_n648.npz", "water_dense_n648.npz"])
def test_gpu_water_golden(fixture):
    g = load(fixture)
    n = len(g["x0"])
    t = topo_from(g, n)
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_TYPE, g["type"]); s.put(capi.F_M, g["m"])
    s.put(capi.F_Z, g["z"]); s.put(capi.F_XN, g["xn0"]); s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    gpu_put_topology(s, t)
    sys_ = gpu_sys(g, n, neighb_update=capi.SEP_BRUTE)
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_lj"]) <= FT
    assert abs(s.scalars().epot - float(g["epot_lj"])) <= FT * abs(float(g["epot_lj"]))
    s.call("sepgpu_stretch_harmonic", C.byref(sys_), 0, 0.316, 68421.0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_bond"]) <= FT
    s.call("sepgpu_angle_cossq", C.byref(sys_), 0, 1.97, 490.0)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_angle"]) <= FT
    s.call("sepgpu_coulomb_sf", C.byref(sys_), float(g["cf"]), cm.EXCL_SAME_MOL)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_coul"]) <= FT
    sc = s.scalars()
    assert abs(sc.epot - float(g["epot_coul"])) <= FT * abs(float(g["epot_coul"]))
    assert abs(sc.ecoul - float(g["ecoul"])) <= FT * abs(float(g["ecoul"]))
    assert relerr(np.array(sc.pot_P[:]), g["pot_P_total"]) <= FT
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    s.call("sepgpu_nosehoover", C.byref(sys_), float(g["temp"]), 0, 0.01)
    s.call("sepgpu_leapfrog", C.byref(sys_))
    sc = s.scalars()
    assert np.abs(s.get(capi.F_X) - g["x1"]).max() <= 1e-11 and np.abs(s.get(capi.F_V) - g["v1"]).max() <= 1e-9
    assert abs(sc.ekin - float(g["ekin"])) <= FT * float(g["ekin"])
    s.put(capi.F_X, g["x0"])
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_coulomb_sf", C.byref(sys_), float(g["cf"]), cm.EXCL_BONDED)
    assert cm.rel_force_err(s.get(capi.F_F), g["f_coul_bonded"]) <= FT
    assert abs(s.scalars().ecoul - float(g["ecoul_bonded"])) <= FT * abs(float(g["ecoul_bonded"]))
    s.close()


@pytest.mark.gpu
def test_gpu_verlet_dpd_golden():
    g = load("dpd_n512.npz")
    n = len(g["x0"]); L = float(g["L"])
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"])
    sys_ = capi.make_sys([L] * 3, 1.0, float(g["dt"]))
    for step in range(2):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.put(capi.F_F, g[f"f{step}"])
        s.call("sepgpu_verlet_dpd", C.byref(sys_), 0.5, step)
        sc = s.scalars()
        assert np.abs(s.get(capi.F_X) - g[f"x{step + 1}"]).max() <= 1e-13
        assert np.abs(s.get(capi.F_V) - g[f"v{step + 1}"]).max() <= 1e-13
        assert np.abs(s.get(capi.F_PV) - g[f"pv{step + 1}"]).max() <= 1e-13
        assert np.abs(s.get(capi.F_PA) - g[f"pa{step + 1}"]).max() <= 1e-12
        assert np.array_equal(s.get(capi.F_CROSSINGS), g[f"cr{step + 1}"])
        assert sc.neighb_flag == int(g[f"flag{step + 1}"])
        assert abs(sc.ekin - float(g[f"ekin{step + 1}"])) <= FT * sc.ekin
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["list", "brute"])
def test_gpu_dpd_force_golden(tag):
    """k_dpd against the reference's own sep_force_dpd (rand() interposed to a constant; SEPGPU_DPD_SEED_FIXED draws the
    same number on the device): conservative, dissipative and random terms, list and brute variants."""
    g = load("dpd_force_n512.npz")
    x = g[tag + "_x"]; pv = g[tag + "_pv"]; L = float(g[tag + "_L"]); n = len(x)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_PV, pv)
    sys_ = capi.make_sys([L] * 3, float(g["cf"]), float(g["dt"]),
                         neighb_update=capi.SEP_LLIST_NEIGHBLIST if tag == "list" else capi.SEP_BRUTE)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_dpd", C.byref(sys_), b"AA", float(g["cf"]), float(g["aij"]), float(g["temp"]), float(g["sigma"]),
           cm.ALL, SEED_FIXED, 0)
    assert cm.rel_force_err(s.get(capi.F_F), g[tag + "_f"]) <= FT
    assert abs(s.scalars().epot - float(g[tag + "_epot"])) <= FT * abs(float(g[tag + "_epot"]))
    s.close()


@pytest.mark.gpu
def test_gpu_nvt_1000_steps_follow_the_reference():
    """1000 steps of the prg1-style NVT loop through the sep_* API of libsep.so against the REFERENCE's own run of the same
    loop (tests/golden/nvt1000_n4096.npz): epot/N, T and p every 100 steps.  Bounds from SURVEY.md section 8c -- the
    reference's own -O2 and -Ofast builds differ by 3.6e-6 in epot at step 1000: 1e-5 relative up to step 500, 1e-4 at
    step 1000 (p: of its thermal scale rho*T), and the list-update counter equal over the first 200 steps."""
    g = load("nvt1000_n4096.npz")
    lib = capi.load()
    rows = cm.drive_nvt_1000(lib, np.ascontiguousarray(g["x0"]), np.ascontiguousarray(g["v0"]), float(g["L"]))
    ref = g["rows"]
    assert rows.shape == ref.shape
    for got, want in zip(rows, ref):
        step = int(want[0])
        tol = 1e-5 if step <= 500 else 1e-4
        assert abs(got[1] - want[1]) <= tol * abs(want[1]), (step, got, want)            # epot/N
        assert abs(got[3] - want[3]) <= tol * abs(want[3]), (step, got, want)            # T
        assert abs(got[4] - want[4]) <= 10 * tol * 0.8, (step, got, want)                # p (fluctuates around 1.6; scale rho*T)
        if step <= 200:
            assert got[5] == want[5], (step, got, want)                                  # nupdate_neighb
    assert abs(rows[-1, 5] - ref[-1, 5]) <= 2


def test_oracle_next_rows_golden():
    """Section-8f rows on the oracle (orc_compress_box, orc_berendsen, orc_relax_temp, orc_force_x0 inside the same
    loops the reference ran, tests/common.py) against tests/golden/next_rows.npz.  Trajectories of 12-30 steps:
    per-step sums to 1e-11 relative, final positions and velocities to 1e-10."""
    g = load("next_rows.npz")
    runs = (("compress", cm.oracle_compress(g["c_x0"], g["c_v0"], float(g["c_L"]))),
            ("beriso", cm.oracle_berendsen(g["c_x0"], g["c_v0"], float(g["c_L"]), steps=12, iso=True, list_mode=True)),
            ("ber", cm.oracle_berendsen(g["b_x0"], g["b_v0"], float(g["b_L"]))),
            ("slit", cm.oracle_slit(g["c_x0"], g["c_v0"], float(g["c_L"]))))
    for name, rec in runs:
        ref = g[name + "_traj"]
        assert rec["traj"].shape == ref.shape, name
        scale = np.abs(ref).max(axis=0)
        assert (np.abs(rec["traj"] - ref).max(axis=0) <= 1e-11 * scale).all(), (name, np.abs(rec["traj"] - ref).max(axis=0) / scale)
        assert np.abs(rec["x"] - g[name + "_x"]).max() <= 1e-10, (name, np.abs(rec["x"] - g[name + "_x"]).max())
        assert np.abs(rec["v"] - g[name + "_v"]).max() <= 1e-10, name
        assert np.array_equal(rec["nsubbox"], g[name + "_nsubbox"]) and abs(rec["volume"] - float(g[name + "_volume"])) <= 1e-12 * rec["volume"]


def test_oracle_stochastic_integrators_golden():
    """sep_fp and sep_langevinGJF on the oracle (orc_fp, orc_langevin_gjf, orc_randn on the same glibc rand() stream)
    against the reference's recorded 30-step loops: identical noise, so the trajectories agree to rounding."""
    g = load("next_rows.npz")
    for which in ("fp", "gjf"):
        rec = cm.oracle_stochastic(g["b_x0"], g["b_v0"], float(g["b_L"]), which)
        ref = g[which + "_traj"]
        scale = np.maximum(np.abs(ref).max(axis=0), 1e-300)
        assert (np.abs(rec["traj"] - ref).max(axis=0) <= 1e-11 * scale).all(), (which, np.abs(rec["traj"] - ref).max(axis=0) / scale)
        assert np.abs(rec["x"] - g[which + "_x"]).max() <= 1e-11, which
        assert np.abs(rec["v"] - g[which + "_v"]).max() <= 1e-11, which


def test_host_randn_stream_equals_the_reference():
    """libsep.so's sep_randn (host, feeds the device integrators) reproduces the reference's Gaussian stream"""
    ref = cm.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libsep_ref.so not built")
    ours = capi.load()
    ref.sep_randn.restype = C.c_double
    ours.sep_randn.restype = C.c_double
    libc = C.CDLL(None)
    libc.srand(99); a = [ref.sep_randn() for _ in range(2000)]
    libc.srand(99); b = [ours.sep_randn() for _ in range(2000)]
    assert a == b
