"""Edge cases of the pair path through the sepgpu C ABI, against the CPU oracle on identical inputs: type selections
without atoms, boxes with three different edge lengths, a half-empty box (most cells empty, the others full), atoms exactly
on the box faces, systems of one and two atoms.  The reference walks the same cases through the same loops
(source/sepprfrc.c:94-224 list, :226-343 brute, :347-513 cell + list build); nothing special-cases them there, so the
device must not either.

Tolerances as everywhere: pair sets bit-exact, forces 1e-10, sums 1e-10.
Sorts after the other GPU files: written after the round's GPU budget was spent (emulator-checked, tests/test_cpu_emu.py)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

pytestmark = pytest.mark.gpu

FTOL = 1e-10
STOL = 1e-10


def _system(x, types=None):
    s = capi.System(len(x))
    s.put(capi.F_X, np.ascontiguousarray(x))
    if types is not None:
        s.put(capi.F_TYPE, types)
    return s


def _oracle_list_force(x, types, length, pairs, tsel, cf, pot):
    orc = cm.oracle()
    f = np.zeros((len(x), 3)); ret = cm.OrcRet()
    lv = cm.dvec3(length)
    pp = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    orc.orc_force_pairs_list(len(x), cm.ptr(x), cm.ptr(types), cm.ptr(lv), cm.ptr(pp), len(pp), tsel, cf, pot, None,
                             cm.ptr(f), C.byref(ret))
    return f, ret


def _oracle_brute_force(x, types, length, tsel, cf):
    orc = cm.oracle()
    f = np.zeros((len(x), 3)); ret = cm.OrcRet()
    lv = cm.dvec3(length); par = cm.dvec3([cf, 1.0, 1.0, 1.0]); tp = cm.OrcTopo()
    orc.orc_force_pairs_brute(len(x), cm.ptr(x), cm.ptr(types), cm.ptr(lv), tsel, cf, cm.POT_LJ_PARAM, cm.ptr(par), cm.ALL,
                              C.byref(tp), cm.ptr(f), C.byref(ret))
    return f, ret


def _check_list_step(x, types, length, tsel=b"AA", cf=2.5, skin=0.25):
    """list build + one list force call against the oracle: pair set, forces, energy, virial"""
    pairs = cm.oracle_pairs(x, length, cf, skin, max_pairs=80 * len(x) + 4096)
    fref, rref = _oracle_list_force(x, types, length, pairs, tsel, cf, cm.POT_LJ_SHIFT)
    s = _system(x, types)
    sys_ = capi.make_sys(list(length), cf, 0.005, skin=skin)
    s.call("sepgpu_neighb_build", C.byref(sys_), cm.ALL)
    assert np.array_equal(cm.pair_set(s.pairs()), cm.pair_set(pairs))
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    p = capi.lj_param(cf, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), tsel, C.byref(p), cm.ALL, 1)
    f = s.get(capi.F_F); sc = s.scalars()
    assert cm.rel_force_err(f, fref) <= FTOL
    assert abs(sc.epot - rref.epot) <= STOL * max(abs(rref.epot), 1.0)
    P = np.array(sc.pot_P[:]); Pref = np.array(rref.pot_P[:])
    assert np.abs(P - Pref).max() <= STOL * max(np.abs(Pref).max(), 1.0)
    s.close()
    return len(pairs)


def _box_lattice(cells, a, jitter, seed):
    """cells[0] x cells[1] x cells[2] atoms on a lattice of spacing a (per direction), jittered, strictly inside the box"""
    a = np.asarray(a, dtype=float)
    g = [(np.arange(cells[k]) + 0.5) * a[k] for k in range(3)]
    z, y, x = np.meshgrid(g[2], g[1], g[0], indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    length = a * np.asarray(cells)
    rng = np.random.default_rng(seed)
    pos = np.mod(pos + rng.uniform(-jitter, jitter, size=pos.shape) * a, length)
    pos[pos >= length] = 0.0
    return np.ascontiguousarray(pos), length


@pytest.mark.parametrize("mode", ["list", "brute"])
def test_type_selection_without_atoms(mode):
    """sep_force_pairs(.., "XX", ..) and "AB" in a system of A atoms only: the loops run and find nothing
    (source/sepprfrc.c:123-126, :262-265) -- forces stay exactly as they were, the call's energy is zero"""
    x, L = cm.lattice(10, 0.8, jitter=0.1, seed=71)
    n = len(x)
    types = np.full(n, ord("A"), dtype=np.uint8)
    s = _system(x, types)
    brute = mode == "brute"
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, **({"neighb_update": capi.SEP_BRUTE} if brute else {"skin": 0.25}))
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), b"XX", C.byref(p), cm.ALL, 1)
    sc = s.scalars()
    assert np.all(s.get(capi.F_F) == 0.0) and sc.epot == 0.0 and np.all(np.array(sc.pot_P[:]) == 0.0)
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
    f1 = s.get(capi.F_F); e1 = s.scalars().epot; P1 = np.array(s.scalars().pot_P[:])
    assert np.abs(f1).max() > 0.0 and e1 != 0.0
    s.call("sepgpu_force_lj", C.byref(sys_), b"AB", C.byref(p), cm.ALL, 1)          # accumulates nothing
    assert np.array_equal(s.get(capi.F_F), f1)
    assert np.array_equal(np.array(s.scalars().pot_P[:]), P1)
    s.close()


@pytest.mark.parametrize("cells,a", [((9, 12, 16), (1.30, 1.05, 0.95)), ((20, 8, 11), (0.9, 1.4, 1.1))])
def test_box_with_three_different_edges(cells, a):
    """cell counts and cell edges differ per direction (source/sepinit.c:60-75 computes them per direction)"""
    x, length = _box_lattice(cells, a, 0.12, seed=72)
    types = np.full(len(x), ord("A"), dtype=np.uint8)
    sys_ = capi.make_sys(list(length), 2.5, 0.005, skin=0.25)
    assert len({int(sys_.nsubbox[k]) for k in range(3)}) > 1
    assert _check_list_step(x, types, length) > 0


def test_half_empty_box_and_two_species():
    """a liquid slab: every cell above z = L/2 is empty, the ones below are full; A/B mixture, the AB call"""
    x, L = cm.lattice(14, 0.8, jitter=0.1, seed=73)
    x = np.ascontiguousarray(x[x[:, 2] < 0.5 * L])
    rng = np.random.default_rng(74)
    types = np.where(rng.random(len(x)) < 0.35, ord("B"), ord("A")).astype(np.uint8)
    for tsel in (b"AA", b"AB", b"BB"):
        assert _check_list_step(x, types, [L] * 3, tsel=tsel) > 0


def test_atoms_exactly_on_the_box_faces():
    """coordinates 0.0 and the largest double below L fall into the first and the last cell (source/sepprfrc.c:404-406) and
    are each other's neighbours through the periodic image"""
    x, L = cm.lattice(12, 0.8, jitter=0.1, seed=75)
    below = np.nextafter(L, 0.0)
    x[0] = [0.0, 0.0, 0.0]
    x[1] = [below, below, L - 1.0]
    x[2] = [0.0, below, 1.2]
    x[3] = [below, 0.0, L - 2.1]
    # keep the planted atoms from sitting on top of lattice neighbours
    d = x[None, :4, :] - x[4:, None, :]
    d -= L * np.round(d / L)
    keep = np.ones(len(x), dtype=bool)
    keep[4:] = (np.linalg.norm(d, axis=2) > 0.8).all(axis=1)
    x = np.ascontiguousarray(x[keep])
    types = np.full(len(x), ord("A"), dtype=np.uint8)
    _check_list_step(x, types, [L] * 3)
    pairs = cm.pair_set(cm.oracle_pairs(x, L, 2.5, 0.25))
    assert ((pairs[:, 0] == 0) & (pairs[:, 1] == 1)).any(), "atoms 0 and 1 are neighbours through all three faces"
    assert ((pairs[:, 0] == 0) & (pairs[:, 1] == 3)).any() and ((pairs[:, 0] == 1) & (pairs[:, 1] == 2)).any()


def test_one_and_two_atoms():
    """a lone atom feels nothing; two atoms whose minimum image crosses the boundary, brute and list mode"""
    L = 12.0
    types1 = np.full(1, ord("A"), dtype=np.uint8)
    for update in (capi.SEP_BRUTE, None):
        s = _system(np.array([[3.0, 4.0, 5.0]]), types1)
        sys_ = capi.make_sys([L] * 3, 2.5, 0.005, **({"neighb_update": update} if update is not None else {"skin": 0.25}))
        p = capi.lj_param(2.5, kind="lj_shift")
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
        assert np.all(s.get(capi.F_F) == 0.0) and s.scalars().epot == 0.0
        s.close()
    x = np.array([[0.3, 6.0, 11.8], [11.6, 6.2, 0.4]])           # |dr| = (0.7, 0.2, 0.6) through two faces
    types = np.full(2, ord("A"), dtype=np.uint8)
    fref, rref = _oracle_brute_force(x, types, [L] * 3, b"AA", 2.5)
    assert abs(fref[0] + fref[1]).max() <= 1e-12 * np.abs(fref).max() and np.abs(fref).max() > 1.0
    s = _system(x, types)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, neighb_update=capi.SEP_BRUTE)
    p = capi.lj_param(2.5, kind="param")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 0)
    f = s.get(capi.F_F); sc = s.scalars()
    assert np.abs(f - fref).max() <= FTOL * np.abs(fref).max()
    assert abs(sc.epot - rref.epot) <= STOL * abs(rref.epot)
    s.close()
    _check_list_step(x, types, [L] * 3)
