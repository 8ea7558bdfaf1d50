"""The benchmark / scale-test systems of seplib_b200/workloads.py on the CPU:
the tiled molecular systems are exact periodic replications of the recorded unit cells (oracle forces on a 2^3
tiling reproduce the reference's unit-cell forces in every copy), molecules are whole, and the .top text that
bench.py feeds to sep_read_topology_file parses back to the same topology in libsep.so and in the reference."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi
from seplib_b200 import workloads as wl


def _bond_lengths(w):
    b = w["blist"]
    d = w["x"][b[:, 0]] - w["x"][b[:, 1]]
    d -= w["L"] * np.round(d / w["L"])
    return np.sqrt((d * d).sum(axis=1))


def test_tiled_butane_replicates_the_unit_cell():
    w = wl.butane(2)
    g = np.load(os.path.join(wl.GOLDEN, "butane_n4000.npz"))
    assert w["n"] == 8 * 4000 and w["nmol"] == 8 * 1000 and w["x"].min() >= 0 and (w["x"] < w["L"]).all()
    r = _bond_lengths(w)
    assert r.max() < 0.7                                      # no molecule torn apart by the tiling
    # oracle LJ forces with the same-molecule exclusion on the tiled system == reference forces of the unit cell
    n = w["n"]
    t = cm.Topo(n); t.molindex[:] = w["molindex"]
    pairs = np.ascontiguousarray(cm.oracle_pairs(w["x"], w["L"], 2.5, 0.25, opt=cm.EXCL_SAME_MOL, topo=t, max_pairs=3_000_000),
                                 dtype=np.int32)
    f = np.zeros((n, 3)); ret = cm.OrcRet()
    cm.oracle().orc_force_pairs_list(n, cm.ptr(w["x"]), cm.ptr(w["type"]), cm.ptr(cm.dvec3(w["L"])), cm.ptr(pairs), len(pairs),
                                     b"CC", 2.5, cm.POT_LJ_SHIFT, None, cm.ptr(f), C.byref(ret))
    for k in range(8):
        assert cm.rel_force_err(f[k * 4000:(k + 1) * 4000], g["f_lj"]) <= 1e-11
    assert abs(ret.epot - 8 * float(g["epot_lj"])) <= 1e-11 * abs(8 * float(g["epot_lj"]))


def test_tiled_water_is_whole_and_neutral():
    w = wl.water(2)
    assert w["n"] == 8 * 648 and w["nmol"] == 8 * 216
    assert _bond_lengths(w).max() < 0.4
    assert abs(w["z"].sum()) < 1e-9
    mol = w["molindex"]
    assert (np.bincount(mol) == 3).all()                      # three atoms per molecule, indices offset per copy


def test_top_text_round_trip(tmp_path):
    ref = cm.ref()
    ours = capi.load()
    w = wl.butane(2)
    top = str(tmp_path / "b.top")
    wl.write_top(w, top)
    got = {}
    for name, lib in (("ours", ours), ("ref", ref)):
        if lib is None:
            continue
        s = cm.ApiSystem(lib, w["x"], w["L"], 2.5, 0.001, v=w["v"], types=w["type"], nneighb=0 if name == "ours" else 1)
        lib.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q")
        t = s.topo()
        got[name] = (t.molindex.copy(), t.bond.copy(), t.angle.copy(), t.dihed.copy(), t.blist.copy(), t.alist.copy(), t.dlist.copy())
        s.close()
    mol, bond, angle, dihed, bl, al, dl = got["ours"]
    assert np.array_equal(mol, w["molindex"]) and np.array_equal(bond, w["bond"])
    assert np.array_equal(angle, w["angle"]) and np.array_equal(dihed, w["dihed"])
    assert np.array_equal(bl, w["blist"]) and np.array_equal(al, w["alist"]) and np.array_equal(dl, w["dlist"])
    if "ref" in got:
        for a, b in zip(got["ours"], got["ref"]):
            assert np.array_equal(a, b)
