"""GPU parity, part 2: list-mode shifted-force Coulomb, DPD pair force (counter-based noise shared with
the oracle), per-type Nose-Hoover, momentum reset, box compression, edge cases, and the reference-facing
sep_* API (host seppart[] buffers through libsep.so) against the reference's own golden trajectory."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

pytestmark = pytest.mark.gpu
FT = 1e-10


def tiled_water(reps=2):
    """golden water box (648 atoms) tiled reps^3 times so that the cell grid has >= 4 cells per side."""
    g = np.load(os.path.join(cm.GOLDEN, "water_n648.npz"))
    L0 = np.atleast_1d(g["L"]).astype(float)
    n0 = len(g["x0"])
    xs, ts, zs, ms, mols = [], [], [], [], []
    k = 0
    for iz in range(reps):
        for iy in range(reps):
            for ix in range(reps):
                xs.append(g["x0"] + np.array([ix, iy, iz]) * L0)
                ts.append(g["type"]); zs.append(g["z"]); ms.append(g["m"])
                mols.append(g["molindex"] + k * (g["molindex"].max() + 1))
                k += 1
    x = np.ascontiguousarray(np.concatenate(xs))
    return (x, np.concatenate(ts).astype(np.uint8), np.concatenate(zs), np.concatenate(ms),
            np.concatenate(mols).astype(np.int32), L0 * reps, n0)


def test_coulomb_list_matches_oracle():
    """prg3-style water in list mode: sep_force_pairs('OO', EXCL_SAME_MOL) builds the list, sep_coulomb_sf
    reuses it (reference source/sepcoulomb.c:8-16) with the full 2.9 cutoff."""
    x, types, z, m, mol, L, _ = tiled_water(2)
    n = len(x)
    cf, skin = 2.9, 0.25
    t = cm.Topo(n); t.molindex[:] = mol
    pairs = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin, opt=cm.EXCL_SAME_MOL, topo=t, max_pairs=4_000_000), dtype=np.int32)
    orc = cm.oracle(); length = cm.dvec3(L)
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pairs), len(pairs), b"OO", 2.5,
                             cm.POT_LJ_SHIFT, None, cm.ptr(fref), C.byref(rref))
    orc.orc_coulomb_sf_list(n, cm.ptr(x), cm.ptr(z), cm.ptr(length), cm.ptr(pairs), len(pairs), cf, cm.ptr(fref), C.byref(rref))
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types); s.put(capi.F_Z, z); s.put(capi.F_M, m); s.put(capi.F_MOLINDEX, mol)
    sys_ = capi.make_sys(L, cf, 5e-4, skin=skin)
    assert sys_.nsubbox[0] >= 4
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
    assert np.array_equal(cm.pair_set(s.pairs(4_000_000)), cm.pair_set(pairs))
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
    s.call("sepgpu_coulomb_sf", C.byref(sys_), cf, cm.EXCL_SAME_MOL)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FT
    assert abs(sc.ecoul - rref.ecoul) <= FT * abs(rref.ecoul)
    assert abs(sc.epot - rref.epot) <= FT * abs(rref.epot)
    assert np.abs(np.array(sc.pot_P[:]) - np.array(rref.pot_P[:])).max() <= FT * np.abs(np.array(rref.pot_P[:])).max()
    s.close()


@pytest.mark.parametrize("mode", ["list", "brute"])
def test_dpd_force_matches_oracle(mode):
    """Groot-Warren DPD (prg6 parameters).  The pair noise is a counter-based generator keyed on
    (seed, step, min(i,j), max(i,j)) implemented identically in the oracle, so forces agree to rounding;
    parity with the reference's glibc rand() stream is statistical only."""
    ncell = 12 if mode == "list" else 6
    x, L = cm.lattice(ncell, 3.0, jitter=0.35, seed=13)
    n = len(x)
    pv = cm.velocities(n, 1.0, seed=14)
    types = np.full(n, ord("A"), dtype=np.uint8)
    cf, skin, dt, aij, temp, sigma = 1.0, 0.25, 0.02, 25.0, 1.0, 3.0
    length = cm.dvec3([L] * 3)
    if mode == "list":
        pairs = cm.oracle_pairs(x, L, cf, skin)
    else:
        orc0 = cm.oracle(); buf = np.empty((n * n, 2), dtype=np.int32); tp = cm.OrcTopo()
        k = orc0.orc_neighb_pairs_n2(n, cm.ptr(x), cm.ptr(length), cf, cm.ALL, C.byref(tp), cm.ptr(buf), n * n)
        pairs = buf[:k]
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    orc = cm.oracle()
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    orc.orc_dpd_force_list(n, cm.ptr(x), cm.ptr(pv), cm.ptr(types), cm.ptr(length), cm.ptr(pairs), len(pairs), b"AA",
                           cf, aij, temp, sigma, dt, 1234, 7, cm.ptr(fref), C.byref(rref))
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_PV, pv)
    sys_ = capi.make_sys([L] * 3, cf, dt, skin=skin,
                         neighb_update=capi.SEP_LLIST_NEIGHBLIST if mode == "list" else capi.SEP_BRUTE)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_dpd", C.byref(sys_), b"AA", cf, aij, temp, sigma, cm.ALL, 1234, 7)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= 1e-9
    assert abs(sc.epot - rref.epot) <= 1e-10 * abs(rref.epot)
    assert np.abs(f.sum(axis=0)).max() <= 1e-8 * np.abs(f).sum()          # Newton's third law incl. the noise
    s.close()


def test_nosehoover_type_and_momentum_reset():
    x, L = cm.lattice(8, 0.8, jitter=0.1, seed=3)
    n = len(x)
    rng = np.random.default_rng(4)
    types = np.where(rng.random(n) < 0.5, ord("B"), ord("A")).astype(np.uint8)
    m = np.where(types == ord("B"), 2.5, 1.0)
    v = cm.velocities(n, 1.3, seed=5, m=m)
    f0 = rng.normal(size=(n, 3))
    orc = cm.oracle()
    fref = f0.copy(); alpha = np.array([0.05, 0.07, 0.02])
    orc.orc_nosehoover_type(n, cm.ptr(v), cm.ptr(m), cm.ptr(types), b"B", cm.ptr(fref), 1.1, cm.ptr(alpha), 10.0, 0.005)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_V, v); s.put(capi.F_TYPE, types); s.put(capi.F_M, m)
    s.call("sepgpu_reset_force"); s.put(capi.F_F, f0)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    a3 = (C.c_double * 3)(0.05, 0.07, 0.02)
    s.call("sepgpu_nosehoover_type", C.byref(sys_), b"B", 1.1, a3, 10.0)
    assert np.abs(np.array(a3[:]) - alpha).max() <= 1e-13 * np.abs(alpha).max()
    assert np.abs(s.get(capi.F_F) - fref).max() <= 1e-12 * np.abs(fref).max()
    # sep_reset_momentum on one species (reference source/sepmisc.c:1173-1192)
    s.call("sepgpu_reset_momentum", b"B")
    vg = s.get(capi.F_V)
    sel = types == ord("B")
    vref = v.copy()
    vref[sel] -= (v[sel] * m[sel, None]).sum(axis=0) / m[sel].sum()
    assert np.abs(vg - vref).max() <= 1e-13
    assert np.abs((vg[sel] * m[sel, None]).sum(axis=0)).max() <= 1e-10
    s.close()


def test_nosehoover_and_type_thermostat_in_the_same_step():
    """sep_nosehoover's multiplier (device slot 0..3) must survive a _sep_nosehoover_type call in the same step: the
    type thermostat keeps its 3-value history in slots of its own (reference source/sepintgr.c:149-198)."""
    x, L = cm.lattice(8, 0.8, jitter=0.1, seed=3)
    n = len(x)
    rng = np.random.default_rng(14)
    types = np.where(rng.random(n) < 0.5, ord("B"), ord("A")).astype(np.uint8)
    m = np.where(types == ord("B"), 2.5, 1.0)
    v = cm.velocities(n, 1.3, seed=15, m=m)
    f0 = rng.normal(size=(n, 3))
    orc = cm.oracle()
    fref = f0.copy(); hist = np.array([0.05, 0.07, 0.02])
    a_ref = orc.orc_nosehoover(n, cm.ptr(v), cm.ptr(m), cm.ptr(fref), 1.0, 0.3, 0.1, 0.005)
    orc.orc_nosehoover_type(n, cm.ptr(v), cm.ptr(m), cm.ptr(types), b"B", cm.ptr(fref), 1.1, cm.ptr(hist), 10.0, 0.005)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_V, v); s.put(capi.F_TYPE, types); s.put(capi.F_M, m)
    s.call("sepgpu_reset_force"); s.put(capi.F_F, f0)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    s.call("sepgpu_set_alpha", 0, 0.3)
    s.call("sepgpu_set_alpha", 1, 0.77)                    # a second caller-owned multiplier, not used in this step
    s.call("sepgpu_nosehoover", C.byref(sys_), 1.0, 0, 0.1)
    a3 = (C.c_double * 3)(0.05, 0.07, 0.02)
    s.call("sepgpu_nosehoover_type", C.byref(sys_), b"B", 1.1, a3, 10.0)
    assert np.abs(np.array(a3[:]) - hist).max() <= 1e-13 * np.abs(hist).max()
    assert np.abs(s.get(capi.F_F) - fref).max() <= 1e-12 * np.abs(fref).max()
    sc = s.scalars()
    assert abs(sc.alpha[0] - a_ref) <= 1e-13 * abs(a_ref)
    assert sc.alpha[1] == 0.77
    s.close()


def test_error_paths():
    """Atom outside [0,L) -> the reference's 'Index larger than array length' class of error; bonded
    exclusion without partner tables; list force without anything to build from is fine (auto build)."""
    x, L = cm.lattice(8, 0.8, jitter=0.05, seed=1)
    x[-1, 2] = L * 1.5           # linear cell index beyond the grid (an x overshoot would alias like the reference's)
    s = capi.System(len(x)); s.put(capi.F_X, x)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    rc = s.lib.sepgpu_neighb_build(s.ctx, C.byref(sys_), 1)
    assert rc == -4                                                     # SEPGPU_ECELL
    s.close()
    x, L = cm.lattice(8, 0.8, jitter=0.05, seed=1)
    s = capi.System(len(x)); s.put(capi.F_X, x)
    assert s.lib.sepgpu_neighb_build(s.ctx, C.byref(sys_), cm.EXCL_BONDED) == -6     # SEPGPU_ESTATE
    assert s.lib.sepgpu_coulomb_sf(s.ctx, C.byref(sys_), 2.5, 1) == -6               # no list yet
    s.close()


def test_positions_beyond_box_alias_like_the_reference():
    """sep_set_lattice offsets atoms by +1.0 (reference source/sepinit.c:333-345), which in prg6 puts some at
    x > L.  The reference files them under whatever cell the unclamped linear index hits
    (source/sepprfrc.c:404-412); the oracle restates that, and the device build must list the same pairs."""
    x, L = cm.lattice(16, 0.8, jitter=0.1, seed=12)
    x[::37, 0] += L * 0.04          # a few atoms slightly beyond the box in x
    x[5::41, 1] += L * 0.03
    inside = x[:, 2] < L - 3.0      # keep z overshoots out: they would leave the head array
    x = np.ascontiguousarray(x[inside])
    ref_pairs = cm.pair_set(cm.oracle_pairs(x, L, 2.5, 0.25))
    s = capi.System(len(x)); s.put(capi.F_X, x)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    assert np.array_equal(cm.pair_set(s.pairs()), ref_pairs)
    s.close()


def test_small_grid_exact_fallback_and_capacity_growth():
    """3 cells per side (prg1's own size) takes the exact warp-per-atom builder; a deliberately tiny
    neighbour capacity must grow transparently."""
    x, L = cm.lattice(9, 0.7, jitter=0.2, seed=8)           # L = 10.1 -> 3 cells of 3.38
    s = capi.System(len(x)); s.put(capi.F_X, x)
    s.call("sepgpu_set_option", b"neighb_cap", 8)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    assert sys_.nsubbox[0] == 3
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(cm.oracle_pairs(x, L, 2.5, 0.25)))
    s.close()


def test_c_driven_loop_equals_the_calls_one_by_one():
    """sepgpu_md_lj_nvt (the prg1 loop driven from C) against the same five calls per step made from here: bit-identical
    state and sums after 30 steps with several list rebuilds."""
    x, L = cm.lattice(12, 0.8, jitter=0.05, seed=41)
    n = len(x)
    v = cm.velocities(n, 2.5, seed=42)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    p = capi.lj_param(2.5, kind="lj_shift")
    out = []
    for c_loop in (False, True):
        s = capi.System(n); s.put(capi.F_X, x); s.put(capi.F_V, v)
        s.call("sepgpu_set_alpha", 0, 0.05)
        if c_loop:
            s.call("sepgpu_md_lj_nvt", C.byref(sys_), b"AA", C.byref(p), 1, 1.0, 0, 0.1, 30)
        else:
            for _ in range(30):
                s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
                s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
                s.call("sepgpu_nosehoover", C.byref(sys_), 1.0, 0, 0.1)
                s.call("sepgpu_leapfrog", C.byref(sys_))
        sc = s.scalars()
        out.append((s.get(capi.F_X), s.get(capi.F_V), sc.epot, sc.ekin, sc.alpha[0], sc.nbuild))
        s.close()
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:] and a[5] >= 3


# ---- the reference-facing API: host seppart[] buffers through libsep.so -----------------------------------
PAIRFUN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_char)


def _user_lj_shift(r2, opt):
    """what a user would write after source/sepmisc.c:131-145 -- a pair function the library does NOT know by address"""
    rri = 1.0 / r2; rri3 = rri * rri * rri
    if opt == b"f":
        return 48.0 * rri3 * (rri3 - 0.5) * rri
    return 4.0 * rri3 * (rri3 - 1.0) + 0.016316891136


def _user_morse(r2, opt):
    r = r2 ** 0.5; e = np.exp(-2.0 * (r - 1.1))
    if opt == b"f":
        return float(2.0 * 2.0 * 0.7 * e * (e - 1.0) / r)           # -u'(r) / r
    return float(0.7 * (e - 1.0) ** 2 - 0.7)


@pytest.mark.parametrize("update", [capi.SEP_LLIST_NEIGHBLIST, capi.SEP_BRUTE])
def test_sep_force_pairs_with_a_pair_function_of_the_callers_own(update):
    """sep_force_pairs takes ANY double fun(double r2, char opt) (reference include/sepprfrc.h:49-51, called per pair at
    source/sepprfrc.c:140-146).  The library samples it once and interpolates on the device.  (1) a Python re-statement
    of sep_lj_shift must reproduce the reference's recorded forces / energy of the same step; (2) a Morse function
    against a numpy sum over the oracle's pair list.  Tolerance 1e-9 relative (cubic interpolation)."""
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    lib = capi.load()
    lib.sep_gpu_set_sync(1)
    L, cf, dt = float(g["L"]), float(g["cf"]), float(g["dt"])
    s = cm.ApiSystem(lib, g["x0"], L, cf, dt, v=g["v0"], update=update, nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    lj, morse = PAIRFUN(_user_lj_shift), PAIRFUN(_user_morse)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"AA", cf, C.cast(lj, C.c_void_p), s.S, s.R, 1)
    lib.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], g["f_pairs"]) <= 1e-9
    assert abs(s.ret.epot - float(g["epot"])) <= 1e-9 * abs(float(g["epot"]))
    # Morse, cut at 2.0: expectation from the positions, in numpy
    x = g["x0"]
    d = x[:, None, :] - x[None, :, :]; d -= L * np.round(d / L)
    r2 = (d * d).sum(axis=2); np.fill_diagonal(r2, 1e30)
    inr = r2 < 4.0
    r = np.sqrt(np.where(inr, r2, 1.0)); e = np.exp(-2.0 * (r - 1.1))
    ft = np.where(inr, 2.0 * 2.0 * 0.7 * e * (e - 1.0) / r, 0.0)
    f_exp = (ft[:, :, None] * d).sum(axis=1)
    u_exp = 0.5 * np.where(inr, 0.7 * (e - 1.0) ** 2 - 0.7, 0.0).sum()
    P_exp = 0.5 * np.einsum("ij,ija,ijb->ab", ft, d, d)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"AA", 2.0, C.cast(morse, C.c_void_p), s.S, s.R, 1)
    lib.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], f_exp) <= 1e-9
    assert abs(s.ret.epot - u_exp) <= 1e-9 * abs(u_exp)
    assert np.abs(np.array(s.ret.pot_P).reshape(3, 3) - P_exp).max() <= 1e-8 * np.abs(P_exp).max()
    s.close()


def test_pair_closer_than_the_table_is_an_error_not_an_extrapolation():
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    x = g["x0"].copy(); L = float(g["L"])
    x[1] = x[0] + np.array([0.2, 0.0, 0.0])                # closer than 0.15 cf
    s = capi.System(len(x)); s.put(capi.F_X, x)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    n = 4096; lo = (0.15 * 2.5) ** 2
    r2 = lo + (6.25 - lo) * np.arange(n) / (n - 1)
    tab = np.empty((n, 2)); tab[:, 0] = 48.0 * r2 ** -7 - 24.0 * r2 ** -4; tab[:, 1] = 4.0 * (r2 ** -6 - r2 ** -3)
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    rc = s.lib.sepgpu_force_table(s.ctx, C.byref(sys_), b"AA", 2.5, tab.ctypes.data_as(C.POINTER(C.c_double)), n, lo, 1, 1)
    assert rc == 0
    sc = capi.GpuScalars()
    assert s.lib.sepgpu_read_scalars(s.ctx, C.byref(sc)) == -8      # SEPGPU_ETABLE
    s.close()


@pytest.mark.parametrize("sync", [1, 0])        # SEP_SYNC_STEP, SEP_SYNC_LAZY
def test_sep_api_lj_loop_matches_reference_golden(sync):
    """The prg1 loop written against include/sep.h, run through libsep.so on host buffers, reproduces the
    reference's own recorded step (forces, positions, sepret, sys flags) and its 40-step trajectory."""
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    lib = capi.load()
    lib.sep_gpu_set_sync(sync)
    L, cf, dt = float(g["L"]), float(g["cf"]), float(g["dt"])
    s = cm.ApiSystem(lib, g["x0"], L, cf, dt, v=g["v0"], nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    alpha = C.c_double(float(g["alpha0"]))
    fun = s.fun("sep_lj_shift")
    temp, tau = float(g["temp"]), float(g["tau"])

    def step():
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), tau, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)

    step()
    if sync == 0:
        lib.sep_gpu_sync(s.atoms)
    assert np.abs(s.view["x"] - g["x1"]).max() <= 1e-12 and np.abs(s.view["v"] - g["v1"]).max() <= 1e-12
    assert cm.rel_force_err(s.view["f"], g["f_nh"]) <= FT
    assert np.array_equal(s.view["crossings"], g["cr1"]) and np.array_equal(s.view["cross_neighb"], g["cn1"])
    assert abs(alpha.value - float(g["alpha1"])) <= 1e-12 * abs(float(g["alpha1"]))
    assert abs(s.ret.epot - float(g["epot"])) <= FT * abs(float(g["epot"]))
    assert abs(s.ret.ekin - float(g["ekin"])) <= FT * float(g["ekin"])
    assert int(s.sys.neighb_flag) == int(g["neighb_flag1"])
    assert abs(s.sys.max_dist2 - float(g["max_dist2"])) <= 1e-12 * float(g["max_dist2"])
    assert abs(s.sys.tnow - dt) <= 1e-15
    nup0 = int(s.sys.nupdate_neighb)
    traj = g["traj"]
    for k in range(40):
        step()
        lib.sep_pressure_tensor(s.R, s.S)
        tol = 1e-9 * (k + 2)
        assert abs(s.ret.epot - traj[k, 0]) <= tol * abs(traj[k, 0])
        assert abs(s.ret.ekin - traj[k, 1]) <= tol * abs(traj[k, 1])
        assert abs(s.ret.p - traj[k, 2]) <= tol * max(abs(traj[k, 2]), 1.0)
        assert abs(alpha.value - traj[k, 3]) <= tol * max(abs(traj[k, 3]), 1e-2)
        if k == 0:
            nup0 = int(s.sys.nupdate_neighb)
    # list rebuild count over the last 39 steps equals the reference's
    assert int(s.sys.nupdate_neighb) - nup0 == int(traj[-1, 4] - traj[0, 4])
    mom = lib.sep_eval_mom(s.atoms, s.n)          # syncs atoms[] in lazy mode
    assert abs(mom) < 1e-10
    assert np.abs(s.view["x"] - g["x41"]).max() <= 1e-6
    s.close()
    lib.sep_gpu_set_sync(1)


def test_sep_api_butane_and_water_steps():
    """prg2 (butane, list mode, EXCL_SAME_MOL + bonded terms) and prg3 (water, SEP_BRUTE, SF Coulomb) force
    sequences through the sep_* API on host buffers against the reference's recorded forces."""
    lib = capi.load()
    lib.sep_gpu_set_sync(1)
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    n = len(g["x0"]); L = g["L"]
    s = cm.ApiSystem(lib, g["x0"], L, float(g["cf"]), float(g["dt"]), v=g["v0"], types=np.full(n, ord("C"), dtype=np.uint8), nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    # topology through our own .top reader: write the file from the golden lists
    top = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"sepb200_butane_{os.getpid()}.top")
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;generated for the test\n")
        for (a, b, t) in g["blist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
        fh.write("\n[ angles ]\n;generated\n")
        for (a, b, c, t) in g["alist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
        fh.write("\n[ dihedrals ]\n;generated\n")
        for (a, b, c, d, t) in g["dlist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {d} {t}\n")
    lib.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q")
    os.unlink(top)
    assert np.array_equal(s.view["molindex"], g["molindex"]) and np.array_equal(s.view["bond"], g["bond"])
    assert np.array_equal(s.view["angle"], g["angle"]) and np.array_equal(s.view["dihed"], g["dihed"])
    rb = (C.c_double * 6)(*g["rb"])
    alpha = C.c_double(float(g["alpha0"]))
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    lib.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
    lib.sep_angle_harmonic(s.atoms, 0, 1.90, 400.0, s.S, s.R)
    lib.sep_torsion_Ryckaert(s.atoms, 0, rb, s.S, s.R)
    assert abs(s.ret.epot - float(g["epot_torsion"])) <= FT * abs(float(g["epot_torsion"]))
    lib.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], g["f_torsion"]) <= FT
    lib.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), 0.1, s.S)
    lib.sep_leapfrog(s.atoms, s.S, s.R)
    assert np.abs(s.view["x"] - g["x1"]).max() <= 1e-11 and np.abs(s.view["v"] - g["v1"]).max() <= 1e-10
    assert abs(s.ret.ekin - float(g["ekin"])) <= FT * float(g["ekin"])
    s.close()


def _write_top(g, tag):
    top = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"sepb200_{tag}_{os.getpid()}.top")
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;generated for the test\n")
        for (a, b, t) in g["blist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
        fh.write("\n[ angles ]\n;generated\n")
        for (a, b, c, t) in g["alist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
        if len(g["dlist"]):
            fh.write("\n[ dihedrals ]\n;generated\n")
            for (a, b, c, d, t) in g["dlist"]:
                fh.write(f"{g['molindex'][a]} {a} {b} {c} {d} {t}\n")
    return top


def _fij_host(s):
    mp = s.sys.molptr.contents
    nm = mp.num_mols
    rows = C.cast(mp.Fij, C.POINTER(C.POINTER(C.POINTER(C.c_float))))
    out = np.zeros((nm, nm, 3), dtype=np.float32)
    for i in range(nm):
        ri = rows[i]
        for j in range(nm):
            out[i, j] = ri[j][0:3]
    return out


# The reference accumulates Fij in float (include/sepstrct.h:94); the device accumulates in FP64 and rounds
# once on read-out, so the table agrees to float rounding of a sum of O(10) terms.
FIJ_TOL = 2e-5


@pytest.mark.parametrize("sync", [1, 0])
def test_sep_api_molecular_pressure_tensor(sync):
    """sep_init_mol / sep_reset_force_mol / sep_mol_pressure_tensor (source/sepmol.c:913-963) on the recorded
    butane (list mode, EXCL_SAME_MOL) and water (brute LJ + brute SF Coulomb) states, against the reference's
    own Fij table and P_mol (tests/golden/molpress.npz)."""
    lib = capi.load()
    lib.sep_gpu_set_sync(sync)
    gm = np.load(os.path.join(cm.GOLDEN, "molpress.npz"))
    # butane
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    n = len(g["x0"])
    s = cm.ApiSystem(lib, g["x0"], g["L"], 2.5, 0.001, v=g["v0"], types=np.full(n, ord("C"), dtype=np.uint8), nneighb=0)
    s.view["crossings"][:] = g["cr0"]
    top = _write_top(g, "mpb")
    lib.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q"); os.unlink(top)
    mols = lib.sep_init_mol(s.atoms, s.S)
    assert s.sys.molptr.contents.flag_Fij == 1 and s.sys.molptr.contents.num_mols == 1000
    for rep in range(2):                    # second pass: the reset really clears the device table
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S); lib.sep_reset_force_mol(s.S)
        lib.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
        lib.sep_mol_pressure_tensor(s.atoms, mols, s.R, s.S)
        F = _fij_host(s)
        ref = np.zeros_like(F)
        idx = gm["butane_fij_idx"]
        ref[idx[:, 0], idx[:, 1]] = gm["butane_fij_val"]
        scale = np.abs(ref).max()
        assert np.abs(F - ref).max() <= FIJ_TOL * scale
        assert np.array_equal(np.abs(F).sum(axis=2) > 0, np.abs(ref).sum(axis=2) > 0)
        kin = np.array([list(r) for r in s.ret.kin_P_mol]); pot = np.array([list(r) for r in s.ret.pot_P_mol])
        assert np.abs(kin - gm["butane_kin_P_mol"]).max() <= 1e-10 * np.abs(gm["butane_kin_P_mol"]).max()
        assert np.abs(pot - gm["butane_pot_P_mol"]).max() <= FIJ_TOL * np.abs(gm["butane_pot_P_mol"]).max()
        assert abs(s.ret.p_mol - float(gm["butane_p_mol"])) <= FIJ_TOL * abs(float(gm["butane_p_mol"]))
    lib.sep_free_mol(mols, s.S)
    s.close()
    # water
    g = np.load(os.path.join(cm.GOLDEN, "water_n648.npz"))
    s = cm.ApiSystem(lib, g["x0"], g["L"], 2.9, 5.0e-4, update=capi.SEP_BRUTE, v=g["v0"], types=g["type"], m=g["m"],
                     z=g["z"], nneighb=0)
    s.view["crossings"][:] = g["cr0"]
    top = _write_top(g, "mpw")
    lib.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q"); os.unlink(top)
    mols = lib.sep_init_mol(s.atoms, s.S)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S); lib.sep_reset_force_mol(s.S)
    lib.sep_force_pairs(s.atoms, b"OO", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    lib.sep_coulomb_sf(s.atoms, 2.9, s.S, s.R, 3)
    lib.sep_mol_pressure_tensor(s.atoms, mols, s.R, s.S)
    F = _fij_host(s)
    ref = gm["water_fij"]
    assert np.abs(F - ref).max() <= FIJ_TOL * np.abs(ref).max()
    pot = np.array([list(r) for r in s.ret.pot_P_mol])
    assert np.abs(pot - gm["water_pot_P_mol"]).max() <= 5 * FIJ_TOL * np.abs(gm["water_pot_P_mol"]).max()
    assert abs(s.ret.p_mol - float(gm["water_p_mol"])) <= 5 * FIJ_TOL * abs(float(gm["water_p_mol"]))
    lib.sep_free_mol(mols, s.S)
    s.close()
    lib.sep_gpu_set_sync(1)


def test_fij_list_equals_brute_and_sums_to_molecular_force():
    """C-ABI: the molecule-molecule force table filled by the list kernels (source/sepprfrc.c:199-207,
    source/sepcoulomb.c:138-147) equals the one filled by the brute kernels (source/sepprfrc.c:70-84,
    source/sepcoulomb.c:68-82), and its row sums are the total pair force on each molecule."""
    x, types, z, m, mol, L, _ = tiled_water(2)
    n = len(x); nmol = int(mol.max()) + 1
    cf, skin = 2.9, 0.25
    p = capi.lj_param(2.5, kind="lj_shift")
    tabs = []
    for update in (capi.SEP_LLIST_NEIGHBLIST, capi.SEP_BRUTE):
        s = capi.System(n)
        s.put(capi.F_X, x); s.put(capi.F_TYPE, types); s.put(capi.F_Z, z); s.put(capi.F_M, m); s.put(capi.F_MOLINDEX, mol)
        sys_ = capi.make_sys(L, cf, 5e-4, neighb_update=update, skin=skin)
        s.call("sepgpu_fij_enable", nmol)
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        if update != capi.SEP_BRUTE:
            s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
        s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
        s.call("sepgpu_coulomb_sf", C.byref(sys_), cf, cm.EXCL_SAME_MOL)
        tab = np.zeros((nmol, nmol, 3), dtype=np.float32)
        s.call("sepgpu_fij_get", tab.ctypes.data_as(C.c_void_p))
        f = s.get(capi.F_F)
        fmol = np.zeros((nmol, 3)); np.add.at(fmol, mol, f)
        assert np.abs(tab.astype(np.float64).sum(axis=1) - fmol).max() <= 1e-4 * np.abs(fmol).max()
        assert np.abs(tab + tab.transpose(1, 0, 2)).max() <= 1e-5 * np.abs(tab).max()     # Fij = -Fji
        # reset clears it
        s.call("sepgpu_fij_reset")
        t0 = np.ones((nmol, nmol, 3), dtype=np.float32)
        s.call("sepgpu_fij_get", t0.ctypes.data_as(C.c_void_p))
        assert not t0.any()
        tabs.append(tab)
        s.close()
    assert np.abs(tabs[0] - tabs[1]).max() <= 1e-5 * np.abs(tabs[1]).max()
