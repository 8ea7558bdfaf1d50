"""CPU-only checks (no compute calls into the device layer):
  * libsep.so loads and exports every symbol include/sep.h and include/sepgpu.h declare;
  * host logic of the sep_* layer (system setup, lattice, seeded velocities, .top reader, pair functions,
    pressure tensor) against the compiled reference when oracle/_ref is present;
  * the oracle port against the reference on fresh random inputs, including the quirks (skin enlarged
    after setup, bonded-exclusion scan, sep_coulomb_sf list skipping uncharged owners);
  * the device layer refuses to run without a GPU instead of falling back.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built (make -C oracle ref)")


def test_library_exports_every_declared_symbol(lib):
    for name in capi.SEPGPU_SYMBOLS + capi.SEP_SYMBOLS:
        assert hasattr(lib, name), f"libsep.so does not export {name}"


def test_symbol_lists_cover_the_headers():
    inc = os.path.join(cm.ROOT, "include")
    decl = re.compile(r"^\s*(?:[A-Za-z_][\w\s\*]*?[\s\*])(_?sep(?:gpu)?_\w+)\s*\(", re.M)
    for header, names in (("sepgpu.h", capi.SEPGPU_SYMBOLS), ("sep.h", capi.SEP_SYMBOLS)):
        text = open(os.path.join(inc, header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        found = {m for m in decl.findall(text) if not m.startswith("sep_Sq")}
        found -= {"sep_rand", "sep_here", "sep_Abs", "sep_Wrap", "sep_Periodic"}
        missing = sorted(found - set(names))
        assert not missing, f"{header}: declared but not in the exported-symbol list: {missing}"


def test_no_cpu_fallback_without_device(lib):
    if lib.sepgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    assert lib.sepgpu_create(C.byref(ctx), 100, -1) == -1            # SEPGPU_ENODEV
    assert b"no CUDA device" in lib.sepgpu_last_error()


def test_struct_layout_matches_reference_abi():
    assert C.sizeof(capi.SepPart) == 568 and C.sizeof(capi.SepRet) == 1000 and C.sizeof(capi.SepSys) == 168
    assert capi.SepPart.xn.offset == 400 and capi.SepPart.pv.offset == 472 and capi.SepPart.molindex.offset == 152


@needs_ref
def test_host_setup_matches_reference(lib):
    r = cm.ref()
    for (L, cf, n, upd) in ((11.26, 1.1225, 1000, 2), (23.2, 2.5, 10000, 2), (6.46, 2.5, 216, 0)):
        a = lib.sep_sys_setup(L, L * 1.1, L * 0.9, cf, 0.005, n, upd)
        b = r.sep_sys_setup(L, L * 1.1, L * 0.9, cf, 0.005, n, upd)
        for fld in ("npart", "volume", "dt", "tnow", "ndof", "cf", "skin", "neighb_update", "neighb_flag", "nupdate_neighb"):
            assert getattr(a, fld) == getattr(b, fld), fld
        assert list(a.length) == list(b.length) and list(a.nsubbox) == list(b.nsubbox) and list(a.lsubbox) == list(b.lsubbox)
    # lattice + seeded velocities: same glibc rand() stream, same arithmetic
    n, L = 1000, (1000 / 0.7) ** (1 / 3)
    pa, pb = lib.sep_init(n, 0), r.sep_init(n, 1)
    sa, sb = lib.sep_sys_setup(L, L, L, 2.5, 0.005, n, 2), r.sep_sys_setup(L, L, L, 2.5, 0.005, n, 2)
    lib.sep_set_lattice(pa, sa); r.sep_set_lattice(pb, sb)
    lib.sep_set_vel_seed(pa, 1.0, 42, sa); r.sep_set_vel_seed(pb, 1.0, 42, sb)
    va, vb = capi.atoms_view(pa, n), capi.atoms_view(pb, n)
    for fld in ("x", "v", "pv", "m", "type", "molindex", "bond", "angle", "dihed"):
        assert np.array_equal(va[fld], vb[fld]), fld
    lib.sep_close(pa, n); r.sep_close(pb, n)
    # pair functions
    for name in ("sep_lj", "sep_lj_shift", "sep_wca"):
        fa, fb = getattr(lib, name), getattr(r, name)
        for f in (fa, fb):
            f.restype = C.c_double; f.argtypes = [C.c_double, C.c_char]
        for r2 in (0.8, 1.0, 1.2599, 2.0, 6.25):
            for opt in (b"f", b"u"):
                assert fa(r2, opt) == fb(r2, opt)


@needs_ref
def test_topology_reader_matches_reference(lib, tmp_path):
    r = cm.ref()
    t = cm.chain_topology(50, 4)
    top = tmp_path / "chains.top"
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;comment\n")
        for (a, b, ty) in t.blist:
            fh.write(f"{t.molindex[a]} {a} {b} {ty}\n")
        fh.write("\n[ angles ]\n;comment\n")
        for (a, b, c, ty) in t.alist:
            fh.write(f"{t.molindex[a]} {a} {b} {c} {ty}\n")
        fh.write("\n[ dihedrals ]\n;comment\n")
        for (a, b, c, d, ty) in t.dlist:
            fh.write(f"{t.molindex[a]} {a} {b} {c} {d} {ty}\n")
    n = 200
    out = []
    for L_ in (lib, r):
        p = L_.sep_init(n, 1 if L_ is r else 0)
        s = L_.sep_sys_setup(20.0, 20.0, 20.0, 2.5, 0.001, n, 2)
        L_.sep_read_topology_file(p, str(top).encode(), C.byref(s), b"q")
        v = capi.atoms_view(p, n)
        mp = s.molptr.contents
        out.append(dict(mol=v["molindex"].copy(), bond=v["bond"].copy(), angle=v["angle"].copy(), dihed=v["dihed"].copy(),
                        nums=(mp.num_mols, mp.num_bonds, mp.num_btypes, mp.num_angles, mp.num_atypes, mp.num_dihedrals, mp.num_dtypes),
                        blist=np.ctypeslib.as_array(mp.blist, shape=(mp.num_bonds * 3,)).copy(),
                        alist=np.ctypeslib.as_array(mp.alist, shape=(mp.num_angles * 4,)).copy(),
                        dlist=np.ctypeslib.as_array(mp.dlist, shape=(mp.num_dihedrals * 5,)).copy()))
    a, b = out
    assert a["nums"] == b["nums"]
    for k in ("mol", "bond", "angle", "dihed", "blist", "alist", "dlist"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["bond"], t.bond) and np.array_equal(a["dihed"], t.dihed)


@needs_ref
@pytest.mark.parametrize("skin,opt", [(0.25, 1), (1.0, 1), (0.25, 3), (0.25, 2)])
def test_oracle_pairs_and_forces_vs_reference_live(skin, opt):
    """Fresh random system through the reference API and through the oracle port: identical pair sets
    (read from ptr[i].neighb), identical forces and sums."""
    r = cm.ref()
    x, L = cm.lattice(12, 0.85, jitter=0.3, seed=int(skin * 100) + opt)
    n = len(x)
    rng = np.random.default_rng(opt)
    types = np.where(rng.random(n) < 0.3, ord("B"), ord("A")).astype(np.uint8)
    topo = cm.chain_topology(n // 4, 4) if opt != 1 else None
    s = cm.ApiSystem(r, x, L, 2.5, 0.005, types=types)
    if topo is not None:
        s.view["molindex"][:] = topo.molindex; s.view["bond"][:] = topo.bond
        s.view["angle"][:] = topo.angle; s.view["dihed"][:] = topo.dihed
    r.sep_set_skin(s.S, skin)
    r.sep_reset_retval(s.R); r.sep_reset_force(s.atoms, s.S)
    r.sep_force_pairs(s.atoms, b"AB", 2.5, s.fun("sep_lj_shift"), s.S, s.R, opt)
    ref_pairs = s.neighb_pairs()
    raw = cm.oracle_pairs(x, L, 2.5, skin, opt=opt, topo=topo)
    assert np.array_equal(raw, ref_pairs)                     # same pairs in the same (row) order
    orc = cm.oracle()
    f = np.zeros((n, 3)); ret = cm.OrcRet(); length = cm.dvec3([L] * 3)
    pp = np.ascontiguousarray(raw, dtype=np.int32)
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(pp), len(pp), b"AB", 2.5, cm.POT_LJ_SHIFT,
                             None, cm.ptr(f), C.byref(ret))
    assert np.array_equal(f, s.view["f"])
    assert ret.epot == s.ret.epot
    assert np.array_equal(np.array(ret.pot_P[:]), np.array(s.ret.pot_P).reshape(9))
    s.close()


def _edge_systems():
    """the inputs of tests/test_gpu_zzzz_edge.py: three different box edges, a half-empty box, atoms exactly on the box faces"""
    a = np.array([1.30, 1.05, 0.95]); cells = (9, 12, 16)
    g = [(np.arange(cells[k]) + 0.5) * a[k] for k in range(3)]
    z, y, x = np.meshgrid(g[2], g[1], g[0], indexing="ij")
    length = a * np.array(cells)
    rng = np.random.default_rng(72)
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    pos = np.mod(pos + rng.uniform(-0.12, 0.12, size=pos.shape) * a, length)
    pos[pos >= length] = 0.0
    yield "edges", np.ascontiguousarray(pos), length
    x, L = cm.lattice(14, 0.8, jitter=0.1, seed=73)
    yield "slab", np.ascontiguousarray(x[x[:, 2] < 0.5 * L]), np.array([L] * 3)
    x, L = cm.lattice(12, 0.8, jitter=0.1, seed=75)
    below = np.nextafter(L, 0.0)
    x[0] = [0.0, 0.0, 0.0]; x[1] = [below, below, L - 1.0]; x[2] = [0.0, below, 1.2]; x[3] = [below, 0.0, L - 2.1]
    d = x[None, :4, :] - x[4:, None, :]
    d -= L * np.round(d / L)
    keep = np.ones(len(x), dtype=bool)
    keep[4:] = (np.linalg.norm(d, axis=2) > 0.8).all(axis=1)
    yield "faces", np.ascontiguousarray(x[keep]), np.array([L] * 3)


@needs_ref
def test_oracle_edge_cases_vs_reference_live():
    """The oracle against the compiled reference on the edge-case inputs the GPU parity tests use (same pairs in the same
    row order, bit-identical forces and sums), so that those GPU tests are pinned to the reference and not to the port."""
    r = cm.ref()
    orc = cm.oracle()
    for name, x, length in _edge_systems():
        n = len(x)
        rng = np.random.default_rng(len(name))
        types = np.where(rng.random(n) < 0.35, ord("B"), ord("A")).astype(np.uint8)
        s = cm.ApiSystem(r, x, length, 2.5, 0.005, types=types)
        for tsel in (b"AA", b"AB", b"XX"):
            r.sep_reset_retval(s.R); r.sep_reset_force(s.atoms, s.S)
            r.sep_force_pairs(s.atoms, tsel, 2.5, s.fun("sep_lj_shift"), s.S, s.R, 1)
            ref_pairs = s.neighb_pairs()
            raw = cm.oracle_pairs(x, length, 2.5, 0.25, max_pairs=80 * n + 4096)
            assert np.array_equal(raw, ref_pairs), name
            f = np.zeros((n, 3)); ret = cm.OrcRet(); lv = cm.dvec3(length)
            pp = np.ascontiguousarray(raw, dtype=np.int32)
            orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(lv), cm.ptr(pp), len(pp), tsel, 2.5, cm.POT_LJ_SHIFT,
                                     None, cm.ptr(f), C.byref(ret))
            assert np.array_equal(f, s.view["f"]), (name, tsel)
            assert ret.epot == s.ret.epot, (name, tsel)
            assert np.array_equal(np.array(ret.pot_P[:]), np.array(s.ret.pot_P).reshape(9)), (name, tsel)
            if tsel == b"XX":
                assert not f.any() and ret.epot == 0.0
        s.close()


@needs_ref
def test_oracle_coulomb_list_and_nosehoover_type_vs_reference_live():
    r = cm.ref()
    x, L = cm.lattice(10, 0.9, jitter=0.3, seed=77)
    n = len(x)
    rng = np.random.default_rng(7)
    z = rng.choice([-1.0, 0.0, 0.5, 1.0], size=n)
    types = np.where(rng.random(n) < 0.5, ord("B"), ord("A")).astype(np.uint8)
    m = np.where(types == ord("B"), 2.0, 1.0)
    v = cm.velocities(n, 1.0, seed=3, m=m)
    s = cm.ApiSystem(r, x, L, 2.5, 0.005, v=v, types=types, m=m, z=z)
    r.sep_reset_retval(s.R); r.sep_reset_force(s.atoms, s.S)
    r.sep_force_pairs(s.atoms, b"AA", 2.5, s.fun("sep_lj"), s.S, s.R, 1)
    r.sep_coulomb_sf(s.atoms, 2.2, s.S, s.R, 1)
    a3 = (C.c_double * 3)(0.01, 0.02, 0.03)
    r._sep_nosehoover_type(s.atoms, b"B", 1.2, a3, 5.0, s.S)
    orc = cm.oracle()
    raw = np.ascontiguousarray(cm.oracle_pairs(x, L, 2.5, 0.25), dtype=np.int32)
    f = np.zeros((n, 3)); ret = cm.OrcRet(); length = cm.dvec3([L] * 3)
    orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(raw), len(raw), b"AA", 2.5, cm.POT_LJ, None, cm.ptr(f), C.byref(ret))
    orc.orc_coulomb_sf_list(n, cm.ptr(x), cm.ptr(z), cm.ptr(length), cm.ptr(raw), len(raw), 2.2, cm.ptr(f), C.byref(ret))
    al = np.array([0.01, 0.02, 0.03])
    orc.orc_nosehoover_type(n, cm.ptr(v), cm.ptr(m), cm.ptr(types), b"B", cm.ptr(f), 1.2, cm.ptr(al), 5.0, 0.005)
    assert np.abs(f - s.view["f"]).max() <= 1e-12 * np.abs(f).max()
    assert abs(ret.ecoul - s.ret.ecoul) <= 1e-13 * abs(ret.ecoul) and abs(ret.epot - s.ret.epot) <= 1e-13 * abs(ret.epot)
    assert np.array_equal(al, np.array(a3[:]))
    s.close()


def test_dpd_uniform_is_pair_symmetric_and_uniform():
    orc = cm.oracle()
    u = np.array([orc.orc_dpd_uniform(99, 5, i, j) for i in range(60) for j in range(i + 1, 60)])
    assert orc.orc_dpd_uniform(99, 5, 3, 17) == orc.orc_dpd_uniform(99, 5, 17, 3)
    assert orc.orc_dpd_uniform(99, 5, 3, 17) != orc.orc_dpd_uniform(99, 6, 3, 17)
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1.0 / 12.0) < 0.01
