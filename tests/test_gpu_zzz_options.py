"""Kernel options (sepgpu_set_option / SEPGPU_OPTS) and the list formats behind them:
  tile_list=1 (default)  rows of 16-bit tile slots + the shared-memory staged tile kernel; 0 = global-index rows + gather kernel
  coulomb_kernel=2       second list Coulomb kernel          typed_sublist=1   per-type sub-lists for typed Lennard-Jones calls
  fin_multi=1            multi-CTA final reductions          step_fold=1       force reduction + Nose-Hoover update folded
                                                                               into the integrator's kernels
Same parity bar everywhere: forces 1e-10 of the oracle / the reference's golden vectors, pair sets bit-exact, sums 1e-10,
and agreement between alternative kernels to rounding over runs with many rebuilds.  All of these run on a B200
(scripts/gpu_r2_full.sh; A/B records in profiles/r02_optin_ab.txt); the CPU suite runs the same file on the kernel emulator
(tests/test_cpu_emu.py).  spec_force=1 (default): the step's first force call launched ahead of the host's rebuild decision.
"""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi
from test_gpu_more import tiled_water

pytestmark = [pytest.mark.gpu]

FT = 1e-10


def _water_system(reps, opts):
    x, types, z, m, mol, L, _ = tiled_water(reps)
    n = len(x)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types); s.put(capi.F_Z, z); s.put(capi.F_M, m); s.put(capi.F_MOLINDEX, mol)
    for k, v in opts.items():
        s.call("sepgpu_set_option", k.encode(), v)
    return s, x, types, z, mol, L


def _water_forces(s, L, cf, skin, lj=True, coulomb=True):
    sys_ = capi.make_sys(L, cf, 5e-4, skin=skin)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
    p = capi.lj_param(2.5, kind="lj_shift")
    if lj:
        s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
    if coulomb:
        s.call("sepgpu_coulomb_sf", C.byref(sys_), cf, cm.EXCL_SAME_MOL)
    return s.get(capi.F_F), s.scalars()


@pytest.mark.parametrize("opts", [{"coulomb_kernel": 1, "typed_sublist": 0}, {"coulomb_kernel": 2, "typed_sublist": 0},
                                  {"coulomb_kernel": 1, "typed_sublist": 1}, {"coulomb_kernel": 2, "typed_sublist": 1}],
                         ids=["first-kernels", "coulomb2", "sublist", "both"])
def test_water_forces_with_optin_kernels_match_oracle(opts):
    """prg3-style water step (typed 'OO' Lennard-Jones + shifted-force Coulomb on the same list) against the oracle
    (reference source/sepprfrc.c:94-224, source/sepcoulomb.c:96-160)."""
    cf, skin = 2.9, 0.25
    s, x, types, z, mol, L = _water_system(2, opts)
    n = len(x)
    t = cm.Topo(n); t.molindex[:] = mol
    pairs = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin, opt=cm.EXCL_SAME_MOL, topo=t, max_pairs=4_000_000), dtype=np.int32)
    orc = cm.oracle(); length = cm.dvec3(L)
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pairs), len(pairs), b"OO", 2.5,
                             cm.POT_LJ_SHIFT, None, cm.ptr(fref), C.byref(rref))
    orc.orc_coulomb_sf_list(n, cm.ptr(x), cm.ptr(z), cm.ptr(length), cm.ptr(pairs), len(pairs), cf, cm.ptr(fref), C.byref(rref))
    f, sc = _water_forces(s, L, cf, skin)
    assert cm.rel_force_err(f, fref) <= FT
    assert abs(sc.ecoul - rref.ecoul) <= FT * abs(rref.ecoul)
    assert abs(sc.epot - rref.epot) <= FT * abs(rref.epot)
    assert np.abs(np.array(sc.pot_P[:]) - np.array(rref.pot_P[:])).max() <= FT * np.abs(np.array(rref.pot_P[:])).max()
    s.close()


def test_optin_kernels_agree_with_the_default_ones_over_steps():
    """The same water system stepped with the default kernels and with both options on: per-step forces and sums agree
    to rounding while the list is rebuilt along the way (the sub-list must follow every rebuild)."""
    cf, skin, dt = 2.9, 0.25, 5e-4
    runs = []
    for opts in ({"coulomb_kernel": 1, "typed_sublist": 0}, {"coulomb_kernel": 2, "typed_sublist": 1}):
        s, x, types, z, mol, L = _water_system(2, opts)
        rng = np.random.default_rng(3)
        s.put(capi.F_M, np.full(len(x), 1e6))                       # no bonded terms here: very heavy atoms, near-ballistic motion
        s.put(capi.F_V, rng.normal(0.0, 40.0, size=x.shape))        # fast atoms: a rebuild every ~3 steps
        sys_ = capi.make_sys(L, cf, dt, skin=skin)
        p = capi.lj_param(2.5, kind="lj_shift")
        rec = []
        for step in range(8):
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            if step == 0 or s.scalars().neighb_flag:
                s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
            s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
            s.call("sepgpu_coulomb_sf", C.byref(sys_), cf, cm.EXCL_SAME_MOL)
            f = s.get(capi.F_F); sc = s.scalars()
            rec.append((f.copy(), sc.epot, sc.ecoul, np.array(sc.pot_P[:]), sc.nbuild))
            s.call("sepgpu_leapfrog", C.byref(sys_))
        runs.append(rec)
        s.close()
    assert runs[0][-1][4] >= 2, "the test is meant to cross a rebuild"
    for (f0, e0, c0, p0, b0), (f1, e1, c1, p1, b1) in zip(*runs):
        assert b0 == b1
        assert cm.rel_force_err(f1, f0) <= 1e-9
        assert abs(e1 - e0) <= 1e-9 * abs(e0) and abs(c1 - c0) <= 1e-9 * abs(c0)
        assert np.abs(p1 - p0).max() <= 1e-9 * np.abs(p0).max()


def test_typed_sublists_for_three_type_pairs():
    """Two species, AA / AB / BB calls with different cutoffs and potentials (prg8-style): every call walks its own
    sub-list; epot is assigned by each list call (reference source/sepprfrc.c:222)."""
    x, L = cm.lattice(10, 0.8, jitter=0.1, seed=11)
    n = len(x)
    cf, skin = 2.5, 0.25
    rng = np.random.default_rng(5)
    types = np.where(rng.random(n) < 0.4, ord("B"), ord("A")).astype(np.uint8)
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin), dtype=np.int32)
    orc = cm.oracle()
    fref = np.zeros((n, 3)); rref = cm.OrcRet(); length = cm.dvec3([L] * 3)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types)
    s.call("sepgpu_set_option", b"typed_sublist", 1)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    for rep in range(2):                                             # second round reuses the cached sub-lists
        fref[:] = 0.0
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        rref = cm.OrcRet()
        for tsel, rc_, pot, kind in ((b"AA", 2.5, cm.POT_LJ_SHIFT, "lj_shift"), (b"AB", 2.0, cm.POT_LJ, "lj"),
                                     (b"BB", 2 ** (1 / 6), cm.POT_WCA, "wca")):
            orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, rc_, pot,
                                     None, cm.ptr(fref), C.byref(rref))
            p = capi.lj_param(rc_, kind=kind)
            s.call("sepgpu_force_lj", C.byref(sys_), tsel, C.byref(p), cm.ALL, 1)
        f = s.get(capi.F_F); sc = s.scalars()
        assert cm.rel_force_err(f, fref) <= FT
        assert abs(sc.epot - rref.epot) <= 1e-10 * max(abs(rref.epot), 1.0)
        P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
        assert np.abs(P - Pref).max() <= 1e-10 * np.abs(Pref).max()
    s.close()


def test_options_through_the_environment(tmp_path):
    """SEPGPU_OPTS carries the options to programs that only see the sep_* API; an unknown name is an error."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import os\n"
            "from seplib_b200 import capi\n"
            "if os.environ.get('SEPGPU_EMU_LIB'): capi.LIB_PATH = os.environ['SEPGPU_EMU_LIB']\n"
            "try:\n"
            "    s = capi.System(64)\n"
            "    print('created')\n"
            "except Exception as e:\n"
            "    print('refused', e)\n") % (cm.ROOT, os.path.join(cm.ROOT, "tests"))
    for opts, want in (("coulomb_kernel=2,typed_sublist=1", "created"), ("no_such_option=1", "refused")):
        env = dict(os.environ); env["SEPGPU_OPTS"] = opts
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert want in r.stdout, (opts, r.stdout, r.stderr)


# ---- tile lists: rows of 16-bit tile slots + the shared-memory staged tile kernel (the default Lennard-Jones path) ---
def _opt(s, name):
    v = C.c_longlong(-1)
    s.call("sepgpu_get_option", name.encode(), C.byref(v))
    return v.value

def _lj(ncell, seed=3, jitter=0.12):
    x, L = cm.lattice(ncell, 0.8, jitter=jitter, seed=seed)
    return x, L


@pytest.mark.parametrize("skin", [0.25, 1.0])
@pytest.mark.parametrize("ncell", [14, 18])
def test_tile_list_holds_exactly_the_reference_pair_set(skin, ncell):
    """Slots decode to the reference's pair set bit for bit -- also with the skin-1.0 quirk (cells narrower than the
    list cutoff) and on grids whose size is not a multiple of the brick (padding tiles, clipped tiles)."""
    x, L = _lj(ncell)
    cf = 2.5
    ref_pairs = cm.pair_set(cm.oracle_pairs(x, L, cf, skin))
    s = capi.System(len(x))
    s.put(capi.F_X, x)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert _opt(s, "list_f16") == 1
    got = cm.pair_set(s.pairs())
    assert got.shape == ref_pairs.shape and np.array_equal(got, ref_pairs)
    assert s.scalars().npairs_listed == 2 * len(ref_pairs)
    s.close()


@pytest.mark.parametrize("opts", ["", "tile_list=0"])
def test_pairs_at_the_list_cutoff_are_decided_in_double_precision(opts):
    """The builder tests candidates in packed FP32 and sends everything within the FP32 error band of the list cutoff to
    the exact FP64 test (reference arithmetic, source/sepprfrc.c:437-452).  600 pairs are placed at (cf + skin)(1 +- eps),
    eps from 1e-15 to 1e-4: the listed set must still equal the oracle's bit for bit, with pairs on both sides of it."""
    x, L = _lj(14, seed=31)
    n = len(x)
    cf, skin = 2.5, 0.25
    rng = np.random.default_rng(17)
    idx = rng.permutation(n)[:1200]
    eps = 10.0 ** rng.uniform(-15, -4, 600) * rng.choice([-1.0, 1.0], 600)
    u = rng.normal(size=(600, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    x = x.copy()
    x[idx[600:]] = x[idx[:600]] + (cf + skin) * (1.0 + eps)[:, None] * u
    x -= L * np.floor(x / L)
    ref_pairs = cm.pair_set(cm.oracle_pairs(x, L, cf, skin))
    planted = {(min(a, b), max(a, b)) for a, b in zip(idx[:600], idx[600:])}
    listed = {tuple(p) for p in ref_pairs.tolist()}
    inside = len(planted & listed)
    assert 200 < inside < 400                      # the oracle itself puts them on both sides
    s = capi.System(n)
    s.put(capi.F_X, x)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("="); s.call("sepgpu_set_option", k.encode(), int(v))
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    got = cm.pair_set(s.pairs())
    assert got.shape == ref_pairs.shape and np.array_equal(got, ref_pairs)
    s.close()


@pytest.mark.parametrize("typed,skin", [(False, 0.25), (True, 0.25), (False, 1.0)])
def test_tile_forces_match_oracle(typed, skin):
    x, L = _lj(12, seed=9)
    n = len(x)
    cf = 2.5
    rng = np.random.default_rng(5)
    types = (np.where(rng.random(n) < 0.4, ord("B"), ord("A")) if typed else np.full(n, ord("A"))).astype(np.uint8)
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin), dtype=np.int32)
    orc = cm.oracle(); length = cm.dvec3([L] * 3)
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    calls = ((b"AA", 2.5, cm.POT_LJ_SHIFT, "lj_shift"), (b"AB", 2.0, cm.POT_LJ, "lj"), (b"BB", 2 ** (1 / 6), cm.POT_WCA, "wca")) \
        if typed else ((b"AA", 2.5, cm.POT_LJ_SHIFT, "lj_shift"),)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types)
    sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    for tsel, rc_, pot, kind in calls:
        orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, rc_, pot,
                                 None, cm.ptr(fref), C.byref(rref))
        p = capi.lj_param(rc_, kind=kind)
        s.call("sepgpu_force_lj", C.byref(sys_), tsel, C.byref(p), cm.ALL, 1)
    assert _opt(s, "list_f16") == 1
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FT
    assert abs(sc.epot - rref.epot) <= 1e-10 * max(abs(rref.epot), 1.0)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= 1e-10 * np.abs(Pref).max()
    s.close()


@pytest.mark.parametrize("typed", [False, True])
def test_tabulated_pair_function_on_tile_lists(typed):
    """sepgpu_force_table (what sep_force_pairs calls for a pair function of the caller's own) on a tile-format list:
    a table sampled from the shifted LJ function and from a Morse function, against the oracle's LJ forces and a numpy
    sum over the oracle's pairs.  Bound 1e-9: four-point cubic interpolation on 65536 points."""
    x, L = _lj(12, seed=9)
    n = len(x)
    cf = 2.5
    rng = np.random.default_rng(5)
    types = (np.where(rng.random(n) < 0.4, ord("B"), ord("A")) if typed else np.full(n, ord("A"))).astype(np.uint8)
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, 0.25), dtype=np.int32)
    orc = cm.oracle(); length = cm.dvec3([L] * 3)
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    tsel = b"AB" if typed else b"AA"
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), tsel, cf, cm.POT_LJ_SHIFT,
                             None, cm.ptr(fref), C.byref(rref))
    nt = 65536; lo = (0.15 * cf) ** 2

    def table(rc, ffun, ufun):
        r2 = lo + (rc * rc - lo) * np.arange(nt) / (nt - 1)
        t = np.empty((nt, 2)); t[:, 0] = ffun(r2); t[:, 1] = ufun(r2)
        return t

    lj = table(cf, lambda r2: 48.0 * r2 ** -7 - 24.0 * r2 ** -4, lambda r2: 4.0 * (r2 ** -6 - r2 ** -3) + 0.016316891136)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types)
    sys_ = capi.make_sys([L] * 3, cf, 0.005)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_table", C.byref(sys_), tsel, cf, lj.ctypes.data_as(C.POINTER(C.c_double)), nt, lo, cm.ALL, 1)
    assert _opt(s, "list_f16") == 1
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= 1e-9
    assert abs(sc.epot - rref.epot) <= 1e-9 * max(abs(rref.epot), 1.0)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= 1e-9 * np.abs(Pref).max()
    # Morse cut at 2.0 on the same list, accumulated on top (epot_assign = 0): expectation in numpy over the listed pairs
    D, a, r0, rc = 0.7, 2.0, 1.1, 2.0
    mo = table(rc, lambda r2: 2 * a * D * np.exp(-a * (np.sqrt(r2) - r0)) * (np.exp(-a * (np.sqrt(r2) - r0)) - 1) / np.sqrt(r2),
               lambda r2: D * (np.exp(-a * (np.sqrt(r2) - r0)) - 1) ** 2 - D)
    i, j = pp[:, 0], pp[:, 1]
    d = x[i] - x[j]; d -= L * np.round(d / L)
    r2 = (d * d).sum(axis=1)
    sel = r2 < rc * rc
    if typed:
        ti, tj = types[i], types[j]
        sel &= ((ti == ord("A")) & (tj == ord("B"))) | ((ti == ord("B")) & (tj == ord("A")))
    r = np.sqrt(r2[sel]); e = np.exp(-a * (r - r0))
    ft = 2 * a * D * e * (e - 1) / r
    fexp = fref.copy()
    np.add.at(fexp, i[sel], ft[:, None] * d[sel]); np.add.at(fexp, j[sel], -ft[:, None] * d[sel])
    uexp = rref.epot + (D * (e - 1) ** 2 - D).sum()
    s.call("sepgpu_force_table", C.byref(sys_), tsel, rc, mo.ctypes.data_as(C.POINTER(C.c_double)), nt, lo, cm.ALL, 0)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fexp) <= 1e-9
    assert abs(sc.epot - uexp) <= 1e-9 * max(abs(uexp), 1.0)
    s.close()


def test_force_launch_sent_ahead_changes_nothing():
    """spec_force (default on): after three identical steps the step's first force call is launched behind the integrator's
    finaliser, guarded by the device-side rebuild flag, into a spare force array, and the next sepgpu_force_lj adopts it.
    Same program with the option off: every step's sums, the rebuild steps and the final state are bit-identical -- also
    across a host-side upload between integrator and force call (the launch must be dropped), steps with a second force
    call on top, and forces read back in between."""
    x, L = _lj(12, seed=51, jitter=0.05)
    n = len(x)
    v = cm.velocities(n, 3.0, seed=52)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    p = capi.lj_param(2.5, kind="lj_shift")
    p2 = capi.lj_param(2.0, eps=0.05, sigma=1.0, aw=1.0, kind="param")
    runs = {}
    for on in (0, 1):
        s = capi.System(n)
        s.put(capi.F_X, x); s.put(capi.F_V, v)
        s.call("sepgpu_set_option", b"spec_force", on)
        s.call("sepgpu_set_alpha", 0, 0.0)
        rec = []
        for step in range(48):
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            if step == 20:                                   # host edit between the integrator and the force call
                vv = s.get(capi.F_V); vv[:, 0] *= 1.001; s.put(capi.F_V, vv)
            s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
            if 30 <= step < 34:                              # a second routine on top in some steps
                s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p2), cm.ALL, 0)
            f = s.get(capi.F_F) if step % 8 == 5 else None
            s.call("sepgpu_nosehoover", C.byref(sys_), 2.0, 0, 0.1)
            s.call("sepgpu_leapfrog", C.byref(sys_))
            sc = s.scalars()
            rec.append((sc.epot, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.max_dist2, sc.neighb_flag, sc.nbuild,
                        None if f is None else f.tobytes()))
        runs[on] = (rec, s.get(capi.F_X).copy(), s.get(capi.F_V).copy(), s.get(capi.F_F).copy(), _opt(s, "spec_adopted"))
        s.close()
    assert runs[0][4] == 0 and runs[1][4] >= 20, (runs[0][4], runs[1][4])
    assert runs[0][0] == runs[1][0]
    for k in (1, 2, 3):
        assert np.array_equal(runs[0][k], runs[1][k])
    assert runs[0][0][-1][6] >= 4                              # several rebuilds: launches that left at the flag


def test_tile_trajectory_follows_the_global_row_kernels():
    """24 NVT steps (prg1-style loop through the C ABI) with the tile kernels (default) and with global-index rows +
    the gather kernel (tile_list = 0): same rebuild steps, energies equal to rounding growth."""
    x, L = _lj(12, seed=21, jitter=0.05)
    n = len(x)
    v = cm.velocities(n, 3.0, seed=22)
    cf, skin, dt = 2.5, 0.25, 0.005
    runs = []
    for on in (0, 1):
        s = capi.System(n)
        s.put(capi.F_X, x); s.put(capi.F_V, v)
        s.call("sepgpu_set_option", b"tile_list", on)
        sys_ = capi.make_sys([L] * 3, cf, dt, skin=skin)
        p = capi.lj_param(cf, kind="lj_shift")
        s.call("sepgpu_set_alpha", 0, 0.0)
        rec = []
        for step in range(24):
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            if step == 0 or s.scalars().neighb_flag:
                s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
                assert _opt(s, "list_f16") == on
            s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
            s.call("sepgpu_nosehoover", C.byref(sys_), 3.0, 0, 0.1)
            s.call("sepgpu_leapfrog", C.byref(sys_))
            sc = s.scalars()
            rec.append((sc.epot, sc.ekin, np.array(sc.pot_P[:]), sc.nbuild))
        runs.append((rec, s.get(capi.F_X)))
        s.close()
    (r0, x0), (r1, x1) = runs
    assert r0[-1][3] >= 3
    for k, ((e0, k0, p0, b0), (e1, k1, p1, b1)) in enumerate(zip(r0, r1)):
        assert b0 == b1, k
        assert abs(e1 - e0) <= 1e-9 * (k + 1) * abs(e0) and abs(k1 - k0) <= 1e-9 * (k + 1) * abs(k0)
        assert np.abs(p1 - p0).max() <= 1e-9 * (k + 1) * np.abs(p0).max()
    d = x1 - x0
    d -= L * np.round(d / L)
    assert np.abs(d).max() <= 1e-8


def test_tile_list_falls_back_to_global_rows_for_coulomb_and_dpd():
    """List Coulomb and DPD walk global-index rows: a charged system never gets tile rows, and a DPD call on a context
    that holds tile rows rebuilds the list -- results equal the tile_list = 0 path."""
    # water: charges present -> global-index rows although the option is on
    s, x, types, z, mol, L = _water_system(2, {"tile_list": 1})
    f1, sc1 = _water_forces(s, L, 2.9, 0.25)
    assert _opt(s, "list_f16") == 0
    s.close()
    s, *_ = _water_system(2, {"tile_list": 0})
    f0, sc0 = _water_forces(s, L, 2.9, 0.25)
    s.close()
    assert np.array_equal(f0, f1) and sc0.epot == sc1.epot
    # DPD after a Lennard-Jones call that built tile rows
    x, Lb = cm.lattice(12, 3.0, jitter=0.35, seed=13)
    n = len(x)
    pv = cm.velocities(n, 1.0, seed=14)
    res = []
    for on in (0, 1):
        s = capi.System(n)
        s.put(capi.F_X, x); s.put(capi.F_PV, pv)
        s.call("sepgpu_set_option", b"tile_list", on)
        sys_ = capi.make_sys([Lb] * 3, 1.0, 0.02, skin=0.25)
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        p = capi.lj_param(1.0, kind="lj")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
        assert _opt(s, "list_f16") == on
        s.call("sepgpu_force_dpd", C.byref(sys_), b"AA", 1.0, 25.0, 1.0, 3.0, cm.ALL, 7, 3)
        assert _opt(s, "list_f16") == 0
        res.append((s.get(capi.F_F), s.scalars().epot))
        s.close()
    assert cm.rel_force_err(res[1][0], res[0][0]) <= 1e-10 and abs(res[1][1] - res[0][1]) <= 1e-10 * abs(res[0][1])


def test_tile_list_with_exclusions_butane_golden():
    """prg2-style butane against what the REFERENCE computed (tests/golden/butane_n4000.npz): same-molecule and
    bonded-partner exclusions through the tile builder and the tile kernel
    (reference source/sepprfrc.c:517-700, 703-740)."""
    import test_golden as tg
    g = tg.load("butane_n4000.npz")
    n = len(g["x0"])
    t = tg.topo_from(g, n)
    s = capi.System(n)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_TYPE, np.full(n, ord("C"), dtype=np.uint8))
    tg.gpu_put_topology(s, t)
    sys_ = tg.gpu_sys(g, n)
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    for opt, pairs, fkey, ekey in ((cm.EXCL_SAME_MOL, "pairs_same_mol", "f_lj", "epot_lj"), (cm.EXCL_BONDED, "pairs_nonbonded", "f_lj_nonbonded", None)):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_neighb_build", C.byref(sys_), opt)
        assert _opt(s, "list_f16") == 1
        assert np.array_equal(cm.pair_set(s.pairs()), g[pairs])
        s.call("sepgpu_force_lj", C.byref(sys_), b"CC", C.byref(p), opt, 1)
        assert cm.rel_force_err(s.get(capi.F_F), g[fkey]) <= FT
        if ekey:
            assert abs(s.scalars().epot - float(g[ekey])) <= FT * abs(float(g[ekey]))
            assert tg.relerr(np.array(s.scalars().pot_P[:]), g["pot_P_lj"]) <= FT
    s.close()


def test_tile_list_index_order_unrelated_to_position_and_half_list_length():
    """Atom indices shuffled against positions: pair set and forces unchanged, and the reference-style half-list length
    (the SEP_NEIGHB = 3000 error condition, source/sepprfrc.c:499-501) is the same in both list formats."""
    x, L = _lj(14, seed=4)
    x = np.ascontiguousarray(x[np.random.default_rng(8).permutation(len(x))])
    n = len(x)
    cf, skin = 2.5, 0.25
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, cf, skin), dtype=np.int32)
    types = np.full(n, ord("A"), dtype=np.uint8)
    orc = cm.oracle(); length = cm.dvec3([L] * 3)
    fref = np.zeros((n, 3)); rref = cm.OrcRet()
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), b"AA", cf, cm.POT_LJ_SHIFT,
                             None, cm.ptr(fref), C.byref(rref))
    half = {}
    for on in (0, 1):
        s = capi.System(n)
        s.put(capi.F_X, x)
        s.call("sepgpu_set_option", b"tile_list", on)
        sys_ = capi.make_sys([L] * 3, cf, 0.005, skin=skin)
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
        assert _opt(s, "list_f16") == on
        assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(pp))
        half[on] = _opt(s, "max_half")
        p = capi.lj_param(cf, kind="lj_shift")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
        f = s.get(capi.F_F); sc = s.scalars()
        assert cm.rel_force_err(f, fref) <= FT
        assert abs(sc.epot - rref.epot) <= 1e-10 * abs(rref.epot)
        s.close()
    assert half[0] == half[1]
    assert half[0] >= 1


def test_tile_list_with_exclusions_water():
    s, x, types, z, mol, L = _water_system(2, {})
    s.put(capi.F_Z, np.zeros(len(x)))                         # uncharged copy: tile rows
    n = len(x)
    t = cm.Topo(n); t.molindex[:] = mol
    pairs = cm.oracle_pairs(x, L, 2.9, 0.25, opt=cm.EXCL_SAME_MOL, topo=t, max_pairs=4_000_000)
    sys_ = capi.make_sys(L, 2.9, 5e-4, skin=0.25)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.EXCL_SAME_MOL)
    assert np.array_equal(cm.pair_set(s.pairs(4_000_000)), cm.pair_set(pairs))
    s.close()


@pytest.mark.parametrize("window", [0, 2])
def test_build_window_changes_nothing_but_the_work(window):
    """Cells are sorted along x inside; with build_window the builder sweeps each row of cells only inside the x window
    within reach of the atom (default: from 24 atoms per cell on; 2 forces it).  Pair set bit-exact against the oracle,
    entry count and the reference-style half-list length equal with and without -- also with same-molecule exclusion."""
    x, L = _lj(14, seed=6, jitter=0.3)
    ref_pairs = cm.pair_set(cm.oracle_pairs(x, L, 2.5, 0.25))
    s = capi.System(len(x)); s.put(capi.F_X, x)
    s.call("sepgpu_set_option", b"build_window", window)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert np.array_equal(cm.pair_set(s.pairs()), ref_pairs)
    sc = s.scalars()
    assert sc.npairs_listed == 2 * len(ref_pairs)
    half = _opt(s, "max_half")
    s.close()
    t = capi.System(len(x)); t.put(capi.F_X, x)
    t.call("sepgpu_set_option", b"tile_list", 0)
    t.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert _opt(t, "max_half") == half
    t.close()
    # water cell with same-molecule exclusion
    w, xw, types, z, mol, Lw = _water_system(2, {"build_window": window})
    n = len(xw)
    tp = cm.Topo(n); tp.molindex[:] = mol
    pairs = cm.oracle_pairs(xw, Lw, 2.9, 0.25, opt=cm.EXCL_SAME_MOL, topo=tp, max_pairs=4_000_000)
    sysw = capi.make_sys(Lw, 2.9, 5e-4, skin=0.25)
    w.call("sepgpu_neighb_build", C.byref(sysw), cm.EXCL_SAME_MOL)
    assert np.array_equal(cm.pair_set(w.pairs(4_000_000)), cm.pair_set(pairs))
    w.close()


# ---- corner cases of the list formats ----------------------------------------------------------------------------------
def test_tile_list_capacity_growth_and_small_grids():
    """A deliberately tiny row capacity grows transparently; a 3-cell grid takes the exact builder and therefore
    global-index rows and the gather kernel."""
    x, L = _lj(13, seed=12)
    s = capi.System(len(x)); s.put(capi.F_X, x)
    s.call("sepgpu_set_option", b"neighb_cap", 8)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert _opt(s, "list_f16") == 1 and _opt(s, "neighb_cap") > 8
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, 2.5, 0.25), dtype=np.int32)
    assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(pp))
    types = np.full(len(x), ord("A"), dtype=np.uint8)
    fref = np.zeros((len(x), 3)); rref = cm.OrcRet(); length = cm.dvec3([L] * 3)
    cm.oracle().orc_force_pairs_list(len(x), cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), b"AA", 2.5,
                                     cm.POT_LJ_SHIFT, None, cm.ptr(fref), C.byref(rref))
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
    assert cm.rel_force_err(s.get(capi.F_F), fref) <= FT
    assert abs(s.scalars().epot - rref.epot) <= FT * abs(rref.epot)
    s.close()
    # 3 cells per side: exact warp-per-atom builder, global-index rows, gather kernel
    x, L = cm.lattice(9, 0.7, jitter=0.2, seed=8)
    s = capi.System(len(x)); s.put(capi.F_X, x)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005)
    assert sys_.nsubbox[0] == 3
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert _opt(s, "list_f16") == 0
    assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(cm.oracle_pairs(x, L, 2.5, 0.25)))
    s.close()


def test_tile_list_dense_cells_shrink_the_tile():
    """High density (DPD-like, 3 atoms per unit volume, cells of 1.25): more atoms per cell -> smaller bricks / tiles,
    same pair set."""
    x, Lb = cm.lattice(14, 3.0, jitter=0.35, seed=17)
    s = capi.System(len(x)); s.put(capi.F_X, x)
    sys_ = capi.make_sys([Lb] * 3, 1.0, 0.02, skin=0.25)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert _opt(s, "list_f16") == 1
    assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(cm.oracle_pairs(x, Lb, 1.0, 0.25)))
    s.close()


def test_typed_sublist_follows_capacity_growth():
    x, L = _lj(12, seed=2)
    n = len(x)
    types = np.where(np.random.default_rng(1).random(n) < 0.5, ord("B"), ord("A")).astype(np.uint8)
    pp = np.ascontiguousarray(cm.oracle_pairs(x, L, 2.5, 0.25), dtype=np.int32)
    fref = np.zeros((n, 3)); rref = cm.OrcRet(); length = cm.dvec3([L] * 3)
    cm.oracle().orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), b"AB", 2.5,
                                     cm.POT_LJ_SHIFT, None, cm.ptr(fref), C.byref(rref))
    s = capi.System(n); s.put(capi.F_X, x); s.put(capi.F_TYPE, types)
    s.call("sepgpu_set_option", b"typed_sublist", 1)
    s.call("sepgpu_set_option", b"neighb_cap", 8)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AB", C.byref(p), cm.ALL, 1)          # builds (and grows) the list itself
    assert _opt(s, "neighb_cap") > 8
    assert cm.rel_force_err(s.get(capi.F_F), fref) <= FT
    assert abs(s.scalars().epot - rref.epot) <= FT * abs(rref.epot)
    s.close()


def test_multi_cta_finalisation_of_the_force_sums():
    """fin_multi: epot / virial / ecoul of a water step (five force routines, assign and accumulate flags) equal the
    single-CTA finalisation to rounding, repeatedly (the ticket counter must re-arm)."""
    res = []
    for on in (0, 1):
        s, x, types, z, mol, L = _water_system(2, {"fin_multi": on})
        out = []
        for rep in range(3):
            f, sc = _water_forces(s, L, 2.9, 0.25)
            out.append((sc.epot, sc.ecoul, np.array(sc.pot_P[:])))
        res.append((f, out))
        s.close()
    assert np.array_equal(res[0][0], res[1][0])                      # forces do not pass through the finaliser
    for (e0, c0, p0), (e1, c1, p1) in zip(res[0][1], res[1][1]):
        assert abs(e1 - e0) <= 1e-12 * abs(e0) and abs(c1 - c0) <= 1e-12 * abs(c0)
        assert np.abs(p1 - p0).max() <= 1e-12 * np.abs(p0).max()
    assert res[1][1][0][0] == res[1][1][2][0]                        # same input, same sums, call after call


# ---- step_fold: last force reduction + Nose-Hoover update folded into the integrator's kernels -------------------
def _nvt_loop(opts, nsteps, peek):
    """prg1-style loop through the C ABI; peek: read the scalar block / the forces between the force call and the
    integrator on some steps (which must settle whatever is pending)"""
    x, L = _lj(12, seed=31, jitter=0.05)
    n = len(x)
    v = cm.velocities(n, 2.5, seed=32)
    rng = np.random.default_rng(4)
    types = np.where(rng.random(n) < 0.5, ord("B"), ord("A")).astype(np.uint8)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_V, v); s.put(capi.F_TYPE, types)
    for k, val in opts.items():
        s.call("sepgpu_set_option", k.encode(), val)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    s.call("sepgpu_set_alpha", 0, 0.05)
    rec = []
    for step in range(nsteps):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        for tsel, rc_, kind in ((b"AA", 2.5, "lj_shift"), (b"AB", 2.0, "lj"), (b"BB", 2 ** (1 / 6), "wca")):
            p = capi.lj_param(rc_, kind=kind)
            s.call("sepgpu_force_lj", C.byref(sys_), tsel, C.byref(p), cm.ALL, 1)
        mid = None
        if peek and step % 3 == 1:
            mid = (s.scalars().epot, None)                      # scalars before the thermostat
        s.call("sepgpu_nosehoover", C.byref(sys_), 2.5, 0, 0.05)
        if peek and step % 3 == 2:
            sc = s.scalars()
            mid = (sc.epot, sc.alpha[0], s.get(capi.F_F).copy())   # multiplier and thermostatted forces before the integrator
        s.call("sepgpu_leapfrog", C.byref(sys_))
        sc = s.scalars()
        rec.append((sc.epot, sc.ekin, sc.alpha[0], np.array(sc.pot_P[:]), np.array(sc.kin_P[:]), sc.max_dist2, sc.neighb_flag, sc.nbuild, mid))
    xf, vf, ff = s.get(capi.F_X), s.get(capi.F_V), s.get(capi.F_F)
    s.close()
    return rec, xf, vf, ff


@pytest.mark.parametrize("peek", [False, True])
def test_step_fold_equals_the_unfolded_step(peek):
    """12 NVT steps with three typed force calls per step: every scalar, the multiplier, the rebuild steps and the final
    state equal the default sequence of kernels to rounding; reading scalars or forces between the calls settles what is
    pending and sees the same values."""
    a = _nvt_loop({"step_fold": 0, "fin_multi": 0}, 12, peek)
    b = _nvt_loop({"step_fold": 1, "fin_multi": 1 if peek else 0}, 12, peek)
    assert a[0][-1][7] >= 2
    for k, (ra, rb) in enumerate(zip(a[0], b[0])):
        tol = 1e-11 * (k + 1)
        assert ra[6] == rb[6] and ra[7] == rb[7], k
        assert abs(ra[0] - rb[0]) <= tol * abs(ra[0]) and abs(ra[1] - rb[1]) <= tol * abs(ra[1]), k
        assert abs(ra[2] - rb[2]) <= tol * max(abs(ra[2]), 1e-3), k
        assert np.abs(ra[3] - rb[3]).max() <= tol * np.abs(ra[3]).max() and np.abs(ra[4] - rb[4]).max() <= tol * np.abs(ra[4]).max()
        assert abs(ra[5] - rb[5]) <= tol * ra[5]
        if ra[8] is not None:
            assert abs(ra[8][0] - rb[8][0]) <= tol * abs(ra[8][0])
            if ra[8][1] is not None:
                assert abs(ra[8][1] - rb[8][1]) <= tol * max(abs(ra[8][1]), 1e-3)
                assert cm.rel_force_err(rb[8][2], ra[8][2]) <= 1e-10
    assert np.abs(a[1] - b[1]).max() <= 1e-9 and np.abs(a[2] - b[2]).max() <= 1e-9
    assert cm.rel_force_err(b[3], a[3]) <= 1e-9


def test_step_fold_water_step_and_option_switch_off():
    """prg3-style step (typed LJ, bond, angle, Coulomb, thermostat, leapfrog) with step_fold; switching the option off
    settles what is pending."""
    res = []
    for on in (0, 1):
        s, x, types, z, mol, L = _water_system(2, {"step_fold": on})
        sys_ = capi.make_sys(L, 2.9, 5e-4, skin=0.25)
        rng = np.random.default_rng(2)
        s.put(capi.F_V, rng.normal(0.0, 1.0, size=x.shape))
        s.call("sepgpu_set_alpha", 0, 0.0)
        out = []
        for step in range(4):
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            p = capi.lj_param(2.5, kind="lj_shift")
            s.call("sepgpu_force_lj", C.byref(sys_), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
            s.call("sepgpu_coulomb_sf", C.byref(sys_), 2.9, cm.EXCL_SAME_MOL)
            s.call("sepgpu_nosehoover", C.byref(sys_), 3.8, 0, 0.01)
            if step == 3 and on:
                s.call("sepgpu_set_option", b"step_fold", 0)        # settles the pending reduction and multiplier update
            s.call("sepgpu_leapfrog", C.byref(sys_))
            sc = s.scalars()
            out.append((sc.epot, sc.ecoul, sc.ekin, sc.alpha[0]))
        res.append(out)
        s.close()
    for k, (ra, rb) in enumerate(zip(*res)):
        for va, vb in zip(ra, rb):
            assert abs(va - vb) <= 1e-11 * (k + 1) * max(abs(va), 1e-3), (k, ra, rb)


def test_step_fold_launches_three_kernels_per_lennard_jones_step():
    """Counted by the CPU kernel emulator (the symbol does not exist in libsep.so): 5 launches per steady-state step by
    default -- force, its reduction, multiplier update, integrator, its reduction -- and 3 with step_fold."""
    lib = capi.load()
    if not hasattr(lib, "sepgpu_emu_launches"):
        pytest.skip("launch counter of the CPU kernel emulator")
    lib.sepgpu_emu_launches.restype = C.c_longlong
    x, L = _lj(12, seed=3, jitter=0.05)
    v = cm.velocities(len(x), 1.0, seed=4)
    per_step = {}
    for on in (0, 1):
        s = capi.System(len(x)); s.put(capi.F_X, x); s.put(capi.F_V, v)
        s.call("sepgpu_set_option", b"step_fold", on)
        sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
        p = capi.lj_param(2.5, kind="lj_shift")
        s.call("sepgpu_set_alpha", 0, 0.1)
        for step in range(5):
            n0 = lib.sepgpu_emu_launches()
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
            s.call("sepgpu_nosehoover", C.byref(sys_), 1.0, 0, 0.1)
            s.call("sepgpu_leapfrog", C.byref(sys_))
            per_step[on] = lib.sepgpu_emu_launches() - n0
        s.close()
    assert per_step == {0: 5, 1: 3}, per_step
