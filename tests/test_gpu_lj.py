"""GPU parity: Lennard-Jones hot path (list build, list force, brute force, thermostat, leapfrog)
against the CPU oracle on identical inputs, through the sepgpu C ABI.

Tolerances: neighbour pair sets bit-exact; per-particle forces
max|df|/max(f_rms,1) <= 1e-10; scalar sums (epot, virial, ekin) rel <= 1e-10."""
import ctypes as C

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

pytestmark = pytest.mark.gpu

FTOL = 1e-10
STOL = 1e-10


def make_lj(ncell=12, rho=0.8, jitter=0.12, seed=3, temp=1.2):
    x, L = cm.lattice(ncell, rho, jitter=jitter, seed=seed)
    v = cm.velocities(len(x), temp, seed=seed + 1)
    return x, v, L


def gpu_system(x, v=None, types=None, m=None):
    s = capi.System(len(x))
    s.put(capi.F_X, x)
    if v is not None:
        s.put(capi.F_V, v)
    if types is not None:
        s.put(capi.F_TYPE, types)
    if m is not None:
        s.put(capi.F_M, m)
    return s


def oracle_force(x, types, L, pairs, tsel, cf, pot, ljp=None):
    orc = cm.oracle()
    n = len(x)
    f = np.zeros((n, 3))
    ret = cm.OrcRet()
    length = cm.dvec3([L] * 3)
    pp = np.ascontiguousarray(pairs, dtype=np.int32)
    lj = cm.dvec3(ljp) if ljp is not None else None
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, cf, pot,
                             cm.ptr(lj) if lj is not None else None, cm.ptr(f), C.byref(ret))
    return f, ret


@pytest.mark.parametrize("skin,prefilter", [(0.25, 1), (0.25, 0), (1.0, 1), (1.0, 0)])
def test_neighbour_pair_set_exact(skin, prefilter):
    """skin 1.0 after setup reproduces the reference quirk: cells stay (cf+0.25) wide, so pairs more
    than one cell apart are missing from the list -- and must be missing here too."""
    x, v, L = make_lj(ncell=14)
    cf = 2.5
    ref_pairs = cm.pair_set(cm.oracle_pairs(x, L, cf, skin))
    s = gpu_system(x)
    s.call("sepgpu_set_option", b"prefilter", prefilter)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    got = cm.pair_set(s.pairs())
    assert got.shape == ref_pairs.shape
    assert np.array_equal(got, ref_pairs)
    sc = s.scalars()
    assert sc.npairs_listed == 2 * len(ref_pairs)
    s.close()


@pytest.mark.parametrize("tpa", [1, 2, 4, 8, 32])
def test_list_force_matches_oracle(tpa):
    x, v, L = make_lj(ncell=12)
    n = len(x)
    cf, skin = 2.5, 0.25
    types = np.full(n, ord("A"), dtype=np.uint8)
    pairs = cm.oracle_pairs(x, L, cf, skin)
    fref, rref = oracle_force(x, types, L, pairs, b"AA", cf, cm.POT_LJ_SHIFT)
    s = gpu_system(x)
    s.call("sepgpu_set_option", b"tpa", tpa)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    p = capi.lj_param(cf, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
    f = s.get(capi.F_F)
    sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FTOL
    assert abs(sc.epot - rref.epot) <= STOL * abs(rref.epot)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= STOL * np.abs(Pref).max()
    s.close()


def test_typed_pairs_and_accumulation():
    """Two species, three sep_force_pairs-style calls (AA, AB, BB) accumulate forces and virial; epot is
    ASSIGNED by each list call (reference source/sepprfrc.c:222), so only the last call's energy stays."""
    x, v, L = make_lj(ncell=10, seed=11)
    n = len(x)
    cf, skin = 2.5, 0.25
    rng = np.random.default_rng(5)
    types = np.where(rng.random(n) < 0.4, ord("B"), ord("A")).astype(np.uint8)
    pairs = cm.oracle_pairs(x, L, cf, skin)
    orc = cm.oracle()
    fref = np.zeros((n, 3)); rref = cm.OrcRet(); length = cm.dvec3([L] * 3)
    pp = np.ascontiguousarray(pairs, dtype=np.int32)
    s = gpu_system(x, types=types)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    for tsel, rc_, pot, kind in ((b"AA", 2.5, cm.POT_LJ_SHIFT, "lj_shift"), (b"AB", 2.0, cm.POT_LJ, "lj"),
                                 (b"BB", 2 ** (1 / 6), cm.POT_WCA, "wca")):
        orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, rc_, pot,
                                 None, cm.ptr(fref), C.byref(rref))
        p = capi.lj_param(rc_, kind=kind)
        s.call("sepgpu_force_lj", C.byref(sys_), tsel, C.byref(p), cm.ALL, 1)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FTOL
    assert abs(sc.epot - rref.epot) <= STOL * max(abs(rref.epot), 1.0)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= STOL * np.abs(Pref).max()
    s.close()


def test_force_lj_param_variant():
    """sep_force_lj: param = {cf, eps, sigma, aw} in code order, epot accumulates (source/sepprfrc.c:785, 922)."""
    x, v, L = make_lj(ncell=10, seed=21)
    n = len(x)
    types = np.full(n, ord("A"), dtype=np.uint8)
    par = [2.2, 0.8, 1.05, 0.7]
    cf, skin = 2.5, 0.25
    pairs = cm.oracle_pairs(x, L, cf, skin)
    fref, rref = oracle_force(x, types, L, pairs, b"AA", par[0], cm.POT_LJ_PARAM, par)
    s = gpu_system(x)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    p = capi.lj_param(par[0], eps=par[1], sigma=par[2], aw=par[3], kind="param")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 0)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FTOL
    assert abs(sc.epot - rref.epot) <= STOL * abs(rref.epot)
    s.close()


def test_brute_force_matches_oracle():
    """prg0's configuration: 216 atoms, SEP_BRUTE."""
    x, v, L = make_lj(ncell=6, seed=31)
    n = len(x)
    types = np.full(n, ord("A"), dtype=np.uint8)
    orc = cm.oracle()
    fref = np.zeros((n, 3)); rref = cm.OrcRet(); length = cm.dvec3([L] * 3)
    par = cm.dvec3([2.5, 1.0, 1.0, 1.0])
    tp = cm.OrcTopo()
    orc.orc_force_pairs_brute(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), b"AA", 2.5, cm.POT_LJ_PARAM, cm.ptr(par),
                              cm.ALL, C.byref(tp), cm.ptr(fref), C.byref(rref))
    s = gpu_system(x)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, neighb_update=capi.SEP_BRUTE)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    p = capi.lj_param(2.5, kind="param")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 0)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FTOL
    assert abs(sc.epot - rref.epot) <= STOL * abs(rref.epot)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= STOL * np.abs(Pref).max()
    s.close()


def oracle_md(x, v, L, nsteps, cf=2.5, skin=0.25, dt=0.005, temp0=1.0, tau=0.1, alpha=0.1, thermostat=True):
    """The prg1 loop on the oracle: reset -> (rebuild) -> force -> nosehoover -> leapfrog."""
    orc = cm.oracle()
    n = len(x)
    x = x.copy(); v = v.copy()
    m = np.ones(n); types = np.full(n, ord("A"), dtype=np.uint8)
    xn = np.zeros((n, 3)); cn = np.zeros((n, 3), dtype=np.int32); cr = np.zeros((n, 3), dtype=np.int32)
    a = np.zeros((n, 3))
    length = cm.dvec3([L] * 3)
    flag, nbuild = 1, 0
    pairs = None
    log = []
    for step in range(nsteps):
        ret = cm.OrcRet()
        f = np.zeros((n, 3))
        maxd2 = C.c_double(0.0)
        if flag:
            pairs = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin), dtype=np.int32)
            flag = 0; nbuild += 1
        orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pairs), len(pairs), b"AA", cf,
                                 cm.POT_LJ_SHIFT, None, cm.ptr(f), C.byref(ret))
        if thermostat:
            alpha = orc.orc_nosehoover(n, cm.ptr(v), cm.ptr(m), cm.ptr(f), temp0, alpha, tau, dt)
        flag = orc.orc_leapfrog(n, cm.ptr(x), cm.ptr(v), cm.ptr(f), cm.ptr(m), cm.ptr(a), cm.ptr(xn), cm.ptr(cn),
                                cm.ptr(cr), cm.ptr(length), dt, skin, C.byref(maxd2), C.byref(ret))
        log.append((ret.epot, ret.ekin, np.array(ret.pot_P[:]).copy(), np.array(ret.kin_P[:]).copy(), alpha, flag, maxd2.value))
    return x, v, f, cr, log, nbuild


@pytest.mark.parametrize("thermostat", [True, False])
def test_md_trajectory_matches_oracle(thermostat):
    """60 steps of the prg1/prg4 loop: positions, velocities, per-step epot/ekin/virial, thermostat
    multiplier, rebuild trigger steps and boundary crossings all follow the oracle."""
    x0, v0, L = make_lj(ncell=10, jitter=0.05, seed=41, temp=1.0)
    nsteps, cf, skin, dt = 60, 2.5, 0.25, 0.005
    xr, vr, fr, crr, log, nbuild_ref = oracle_md(x0, v0, L, nsteps, cf, skin, dt, thermostat=thermostat)
    s = gpu_system(x0, v0)
    sys_ = capi.make_sys([L] * 3, cf, dt, skin=skin)
    p = capi.lj_param(cf, kind="lj_shift")
    s.call("sepgpu_set_alpha", 0, 0.1)
    flag = 1
    for step in range(nsteps):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        if flag:
            s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
        if thermostat:
            s.call("sepgpu_nosehoover", C.byref(sys_), 1.0, 0, 0.1)
        s.call("sepgpu_leapfrog", C.byref(sys_))
        sc = s.scalars()
        flag = sc.neighb_flag
        epot, ekin, potP, kinP, alpha, rflag, maxd2 = log[step]
        tol = 1e-9 * (1 + step)          # chaotic growth of rounding differences over the trajectory
        assert flag == rflag, f"trigger differs at step {step}"
        assert abs(sc.epot - epot) <= tol * abs(epot)
        assert abs(sc.ekin - ekin) <= tol * abs(ekin)
        assert abs(sc.max_dist2 - maxd2) <= tol * max(maxd2, 1e-3)
        if thermostat:
            assert abs(sc.alpha[0] - alpha) <= tol * max(abs(alpha), 1e-3)
        assert np.abs(np.array(sc.pot_P[:]) - potP).max() <= tol * np.abs(potP).max()
        assert np.abs(np.array(sc.kin_P[:]) - kinP).max() <= tol * np.abs(kinP).max()
    assert s.scalars().nbuild == nbuild_ref
    x = s.get(capi.F_X); v = s.get(capi.F_V); f = s.get(capi.F_F); cr = s.get(capi.F_CROSSINGS)
    assert np.abs(x - xr).max() <= 1e-7
    assert np.abs(v - vr).max() <= 1e-7
    assert cm.rel_force_err(f, fr) <= 1e-6
    assert np.array_equal(cr, crr)
    s.close()
