"""prg5's helpers through the sep_* API: sep_omp_bond / sep_omp_angle / sep_omp_torsion (reference include/sepomp.h:87-123,
source/sepomp.c:179-329) add the forces of ONE bonded term kind to a matrix of the caller's and touch nothing else.  Here
they are served by the device (sepgpu_bonded_side).  Checked on the reference's evolved 4000-atom butane cell against the
compiled reference's own sep_omp_* on the same host arrays (oracle/_ref), and for the properties the reference's loop has:
accumulation into the matrix, atoms[].f and sepret untouched.

Sorts after the other GPU files: written after the round's GPU budget was spent (emulator-checked, tests/test_cpu_emu.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
import test_gpu_prgs as prgs
from seplib_b200 import capi

pytestmark = pytest.mark.gpu

DPP = C.POINTER(C.POINTER(C.c_double))


def _declare(lib):
    lib.sep_matrix.restype = DPP
    lib.sep_matrix.argtypes = [C.c_size_t, C.c_size_t]
    lib.sep_free_matrix.argtypes = [DPP, C.c_size_t]
    lib.sep_matrix_set.argtypes = [DPP, C.c_size_t, C.c_size_t, C.c_double]
    lib.sep_omp_bond.argtypes = [DPP, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.sep_omp_angle.argtypes = [DPP, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.sep_omp_torsion.argtypes = [DPP, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    for f in (lib.sep_free_matrix, lib.sep_matrix_set, lib.sep_omp_bond, lib.sep_omp_angle, lib.sep_omp_torsion):
        f.restype = None


def _matrix(m, n):
    return np.array([[m[i][k] for k in range(3)] for i in range(n)])


def _butane(lib, g):
    n = len(g["x0"])
    s = cm.ApiSystem(lib, g["x0"], g["L"], float(g["cf"]), float(g["dt"]), v=g["v0"], types=np.full(n, ord("C"), dtype=np.uint8),
                     nneighb=0 if lib is not cm.ref() else 3000)
    top = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"sepb200_omp_{os.getpid()}.top")
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
    return s


def _three_kinds(lib, s, g, n):
    """prg5's two sections (prgs/prg5.c:58-70): bonds + angles into one matrix, torsions into the other"""
    rb = (C.c_double * 6)(*g["rb"])
    f1, f2 = lib.sep_matrix(n, 3), lib.sep_matrix(n, 3)
    lib.sep_matrix_set(f1, n, 3, 0.0)
    lib.sep_omp_bond(f1, s.atoms, 0, 0.407, 2074.0, s.S)
    fb = _matrix(f1, n)
    lib.sep_omp_angle(f1, s.atoms, 0, 1.90, 425.0, s.S)
    fba = _matrix(f1, n)
    lib.sep_matrix_set(f2, n, 3, 0.0)
    lib.sep_omp_torsion(f2, s.atoms, 0, rb, s.S)
    ft = _matrix(f2, n)
    lib.sep_omp_torsion(f2, s.atoms, 0, rb, s.S)                  # a second call ADDS to the matrix (source/sepomp.c:321-324)
    ft2 = _matrix(f2, n)
    lib.sep_free_matrix(f1, n); lib.sep_free_matrix(f2, n)
    return fb, fba, ft, ft2


def test_sep_omp_helpers_match_the_reference_and_touch_nothing_else():
    lib = capi.load()
    _declare(lib)
    lib.sep_gpu_set_sync(1)
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    n = len(g["x0"])
    s = _butane(lib, g)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    lib.sep_gpu_sync(s.atoms)
    f_before = s.view["f"].copy(); epot_before = s.ret.epot; P_before = np.array(s.ret.pot_P).copy()
    fb, fba, ft, ft2 = _three_kinds(lib, s, g, n)
    assert np.abs(fb).max() > 1.0 and np.abs(ft).max() > 1.0
    assert np.abs(fb.sum(axis=0)).max() <= 1e-9 * np.abs(fb).max() * np.sqrt(n)           # each kind sums to zero over the atoms
    assert np.allclose(ft2, 2.0 * ft, rtol=1e-14, atol=0.0)
    lib.sep_gpu_sync(s.atoms)
    assert np.array_equal(s.view["f"], f_before) and s.ret.epot == epot_before
    assert np.array_equal(np.array(s.ret.pot_P), P_before)
    # the same three kinds through the ordinary entries: the difference of atoms[].f before and after each call
    lib.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
    lib.sep_gpu_sync(s.atoms)
    d = s.view["f"] - f_before
    assert np.abs(d - fb).max() <= 1e-10 * np.abs(fb).max()
    s.close()
    if cm.have_ref():
        r = cm.ref()
        _declare(r)
        sr = _butane(r, g)
        rb_, rba, rt, _ = _three_kinds(r, sr, g, n)
        for ours, theirs in ((fb, rb_), (fba, rba), (ft, rt)):
            assert np.abs(ours - theirs).max() <= 1e-12 * np.abs(theirs).max()
        sr.close()


def test_uploads_that_change_nothing_keep_the_neighbour_list():
    """A program that edits one member of atoms[] between hot calls makes the host layer upload all of them (a write fault
    cannot tell which member it was): positions, types, molecule indices that come back with the same values must not
    invalidate the list (prg5 would rebuild it every step); a single changed coordinate must."""
    x, L = cm.lattice(10, 0.8, jitter=0.1, seed=91)
    n = len(x)
    types = np.full(n, ord("A"), dtype=np.uint8)
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_TYPE, types)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    p = capi.lj_param(2.5, kind="lj_shift")

    def force():
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
        return s.get(capi.F_F), s.scalars().nbuild

    f0, nb0 = force()
    s.put(capi.F_X, s.get(capi.F_X)); s.put(capi.F_TYPE, types); s.put(capi.F_MOLINDEX, s.get(capi.F_MOLINDEX))
    f1, nb1 = force()
    assert nb1 == nb0 and np.array_equal(f1, f0)
    x2 = s.get(capi.F_X); x2[7, 0] += 1e-3
    s.put(capi.F_X, x2)
    f2, nb2 = force()
    assert nb2 == nb0 + 1 and np.abs(f2 - f0).max() > 0.0
    types2 = types.copy(); types2[3] = ord("B")
    s.put(capi.F_TYPE, types2)
    f3, nb3 = force()
    assert nb3 == nb2 + 1 and np.all(f3[3] == 0.0)                     # atom 3 left the "AA" selection
    s.close()


def test_prg5_omp_model_two(tmp_path):
    """prgs/prg5.c, unchanged and compiled with -fopenmp: butane with the bonded forces taken through sep_omp_bond /
    sep_omp_angle / sep_omp_torsion from two OpenMP sections into matrices of its own, added to atoms[i].f BY THE PROGRAM
    between the hot calls (the page-protected atoms[] notices), box compression + momentum reset every 10 steps.
    Deterministic (velocities come from the start file): the printed temperature against the same program linked with the
    compiled reference (tests/golden/prg5.ref.out, tests/golden/make_prg_outputs.sh).  columns: n t T"""
    cm.write_molecular_start_files(tmp_path)
    got, _ = prgs.run_prg("prg5", tmp_path=tmp_path)
    ref = prgs.golden("prg5.ref.out")
    assert got.shape == ref.shape == (10, 3)
    assert abs(got[0, 2] - ref[0, 2]) <= 1.1e-3, (got[0], ref[0])            # step 0: printed precision
    assert abs(got[1, 2] - ref[1, 2]) <= 0.02, (got[1], ref[1])              # step 100: rounding has grown, not decorrelated
    assert abs(got[2:, 2].mean() - ref[2:, 2].mean()) < 0.08                 # thermostatted temperature (4.0), 8 samples of +-0.06



def test_editing_forces_on_the_host_every_step_does_not_rebuild_the_list_every_step():
    """prg5's pattern through the sep_* API in the default coherence mode (page-protected atoms[]): the program adds to
    atoms[i].f between the force call and the integrator, every step.  Same number of list rebuilds as without the edit."""
    lib = capi.load()
    _declare(lib)
    lib.sep_gpu_set_sync(3)                                   # SEP_SYNC_AUTO
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    counts = []
    for edit in (False, True):
        s = _butane(lib, g)
        alpha = C.c_double(0.1)
        for _ in range(30):
            lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
            lib.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
            if edit:
                s.view["f"][:, 0] += 1e-9
            lib.sep_nosehoover(s.atoms, C.c_double(4.0), C.byref(alpha), C.c_double(0.1), s.S)
            lib.sep_leapfrog(s.atoms, s.S, s.R)
        counts.append(int(s.sys.nupdate_neighb))
        s.close()
    assert counts[0] == counts[1] and 1 <= counts[0] <= 10, counts
