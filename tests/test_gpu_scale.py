"""Full-size checks at the BASELINE.json configurations (C1 1 M LJ atoms, C2 864 000-atom
butane, C3 1.12 M-atom water).  The CPU oracle cannot run these sizes in seconds, so parity is carried by
size-independent properties:

  * the list kernels against the brute kernels of the same library -- the brute kernels restate the reference's
    all-pairs arithmetic (sep_Wrap branches, no FMA in r^2) and share nothing with the cell grid / list builder;
    both are pinned to the oracle and the reference's golden vectors at small size (test_golden.py);
  * Newton's third law: the total force of every pair routine and every bonded routine vanishes;
  * the neighbour list is symmetric, rebuilds are idempotent, momentum is conserved over a run and the
    NVE energy drift stays inside the bound SURVEY.md section 8c derives from the reference's own spread.

Tolerances: forces 1e-10 of max(f_rms, 1) per atom (FT), scalar sums 1e-10 relative.
"""
import ctypes as C

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi
from seplib_b200 import workloads as wl

pytestmark = pytest.mark.gpu

FT = 1e-10


def _molecular_system(w, cf, dt, update):
    s = capi.System(w["n"])
    s.put(capi.F_X, w["x"]); s.put(capi.F_V, w["v"]); s.put(capi.F_TYPE, w["type"]); s.put(capi.F_M, w["m"])
    s.put(capi.F_Z, w["z"]); s.put(capi.F_MOLINDEX, w["molindex"])
    s.put(capi.F_BOND, w["bond"]); s.put(capi.F_ANGLE, w["angle"]); s.put(capi.F_DIHED, w["dihed"])
    s.set_topology(w["blist"], w["alist"], w["dlist"])
    return s, capi.make_sys(w["L"], cf, dt, neighb_update=update)


def _net(f):
    return np.abs(f.sum(axis=0)).max()


def test_c1_lj_1m_list_equals_brute_and_conserves():
    x, L = wl.lj_lattice(100, 0.8)
    n = len(x)
    v = wl.lj_velocities(n, 1.0, 7)
    rc, dt = 2.5, 0.005
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_V, v)
    sys_ = capi.make_sys([L] * 3, rc, dt)
    p = capi.lj_param(rc, kind="lj_shift")

    def forces():
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)

    # thermalise off the lattice (NVE), tracking the conserved quantity: after sep_leapfrog the scalar block holds
    # epot(t) from the force call and the reference's ekin(t) (mean of the two half-step velocities squared)
    etot = []
    for step in range(120):
        forces()
        s.call("sepgpu_leapfrog", C.byref(sys_))
        if step in (19, 119):
            sc = s.scalars(); etot.append((sc.epot + sc.ekin) / n)
    forces()
    sc = s.scalars()
    print("C1 etot/N at step 19 and 119:", etot)
    assert abs(etot[1] - etot[0]) <= 5e-4, etot                  # leapfrog fluctuation at dt = 0.005, no drift
    vv = s.get(capi.F_V)
    assert np.abs(vv.sum(axis=0)).max() <= 1e-9                  # momentum: started at zero
    assert sc.nbuild >= 2
    # list forces, energy and virial on the evolved state ...
    f_list = s.get(capi.F_F); sc_list = s.scalars()
    assert _net(f_list) <= 1e-9 * np.abs(f_list).sum(axis=0).max()
    # ... a rebuild on the same positions gives the same list and the same numbers (idempotence)
    npairs = sc_list.npairs_listed
    s.call("sepgpu_request_rebuild")
    forces()
    sc2 = s.scalars()
    assert sc2.nbuild == sc_list.nbuild + 1 and sc2.npairs_listed <= npairs and sc2.npairs_listed % 2 == 0
    f_list2 = s.get(capi.F_F)
    assert cm.rel_force_err(f_list2, f_list) <= FT and abs(sc2.epot - sc_list.epot) <= FT * abs(sc_list.epot)
    # ... and the all-pairs kernel (reference arithmetic, no grid, no list) agrees
    xw = s.get(capi.F_X)
    b = capi.System(n)
    b.put(capi.F_X, xw)
    bsys = capi.make_sys([L] * 3, rc, dt, neighb_update=capi.SEP_BRUTE)
    b.call("sepgpu_reset_ret"); b.call("sepgpu_reset_force")
    b.call("sepgpu_force_lj", C.byref(bsys), b"AA", C.byref(p), 1, 1)
    f_brute = b.get(capi.F_F); sc_b = b.scalars()
    b.close()
    assert cm.rel_force_err(f_list2, f_brute) <= FT
    assert abs(sc2.epot - sc_b.epot) <= FT * abs(sc_b.epot)
    pl, pb = np.array(sc2.pot_P[:]), np.array(sc_b.pot_P[:])
    assert np.abs(pl - pb).max() <= FT * np.abs(pb).max()
    s.close()


def test_c2_butane_864k_list_equals_brute():
    w = wl.butane(6)
    P = wl.BUTANE
    assert w["n"] == 864_000 and w["nmol"] == 216_000
    s, sys_ = _molecular_system(w, P["cf"], P["dt"], capi.SEP_LLIST_NEIGHBLIST)
    p = capi.lj_param(P["cf"], kind="lj_shift")
    rb = (C.c_double * 6)(*P["rb"])
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), P["types"], C.byref(p), cm.EXCL_SAME_MOL, 1)
    f_lj = s.get(capi.F_F); sc_lj = s.scalars()
    assert _net(f_lj) <= 1e-9 * np.abs(f_lj).sum(axis=0).max()
    # periodic replication: every copy of the unit cell feels the unit cell's forces (reference golden, 4000 atoms)
    g = np.load(wl.GOLDEN + "/butane_n4000.npz")
    for k in (0, 77, 215):
        assert cm.rel_force_err(f_lj[k * 4000:(k + 1) * 4000], g["f_lj"]) <= FT
    assert abs(sc_lj.epot - 216 * float(g["epot_lj"])) <= FT * abs(216 * float(g["epot_lj"]))
    s.call("sepgpu_stretch_harmonic", C.byref(sys_), 0, P["lbond"], P["kbond"])
    s.call("sepgpu_angle_harmonic", C.byref(sys_), 0, P["angle"], P["kangle"])
    s.call("sepgpu_torsion_ryckaert", C.byref(sys_), 0, rb)
    f_all = s.get(capi.F_F); sc = s.scalars()
    assert _net(f_all) <= 1e-9 * np.abs(f_all).sum(axis=0).max()
    for k in (0, 100, 215):
        assert cm.rel_force_err(f_all[k * 4000:(k + 1) * 4000], g["f_torsion"]) <= FT
    assert abs(sc.epot - 216 * float(g["epot_torsion"])) <= FT * abs(216 * float(g["epot_torsion"]))
    s.close()
    # all-pairs kernel with the same-molecule rule
    b, bsys = _molecular_system(w, P["cf"], P["dt"], capi.SEP_BRUTE)
    b.call("sepgpu_reset_ret"); b.call("sepgpu_reset_force")
    b.call("sepgpu_force_lj", C.byref(bsys), P["types"], C.byref(p), cm.EXCL_SAME_MOL, 1)
    f_b = b.get(capi.F_F); sc_b = b.scalars()
    b.close()
    assert cm.rel_force_err(f_lj, f_b) <= FT and abs(sc_lj.epot - sc_b.epot) <= FT * abs(sc_b.epot)
    assert np.abs(np.array(sc_lj.pot_P[:]) - np.array(sc_b.pot_P[:])).max() <= FT * np.abs(np.array(sc_b.pot_P[:])).max()


def test_c3_water_1p1m_list_matches_unit_cell_reference():
    w = wl.water(12)
    P = wl.WATER
    assert w["n"] == 1_119_744
    s, sys_ = _molecular_system(w, P["cf"], P["dt"], capi.SEP_LLIST_NEIGHBLIST)
    assert sys_.nsubbox[0] == 22
    p = capi.lj_param(P["cf_lj"], kind="lj_shift")
    g = np.load(wl.GOLDEN + "/water_dense_n648.npz")
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(sys_), P["types"], C.byref(p), cm.EXCL_SAME_MOL, 1)
    f = s.get(capi.F_F)
    for k in (0, 555, 1727):
        assert cm.rel_force_err(f[k * 648:(k + 1) * 648], g["f_lj"]) <= FT
    s.call("sepgpu_stretch_harmonic", C.byref(sys_), 0, P["lbond"], P["kbond"])
    s.call("sepgpu_angle_cossq", C.byref(sys_), 0, P["angle"], P["kangle"])
    s.call("sepgpu_coulomb_sf", C.byref(sys_), P["cf"], cm.EXCL_SAME_MOL)
    f = s.get(capi.F_F); sc = s.scalars()
    assert _net(f) <= 1e-9 * np.abs(f).sum(axis=0).max()
    # the periodic images of the 648-atom cell ARE the tiled system: the reference's brute-force result on the
    # unit cell (tests/golden/water_dense_n648.npz) must reappear in every copy
    for k in (0, 864, 1727):
        assert cm.rel_force_err(f[k * 648:(k + 1) * 648], g["f_coul"]) <= FT
    assert abs(sc.ecoul - 1728 * float(g["ecoul"])) <= FT * abs(1728 * float(g["ecoul"]))
    assert abs(sc.epot - 1728 * float(g["epot_coul"])) <= FT * abs(1728 * float(g["epot_coul"]))
    assert np.abs(np.array(sc.pot_P[:]) - 1728 * g["pot_P_total"]).max() <= FT * np.abs(1728 * g["pot_P_total"]).max()
    s.close()
