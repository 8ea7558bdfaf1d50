"""Samplers (host post-processing): libsep.so's sep_add_sampler / sep_sample
against the reference's, fed the SAME sequence of states.  The reference build (oracle/_ref/libsep_ref.so) runs
the simulation on the CPU and samples into one directory; after every step the same atoms[] / sepret / sepsys
are handed to libsep.so's sampler, which writes into another directory.  The files must agree to the printed
precision (%f, 1e-6).  No GPU is involved: the sampler code is host C and the arrays are not device-bound."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi


class RefSampler(C.Structure):              # the reference's sepsampler, opaque (include/sepsampler.h:217-239, 248 bytes)
    _fields_ = [("blob", C.c_char * 512)]


class OurSampler(C.Structure):              # include/sep.h
    _fields_ = [("molptr", C.c_void_p), ("impl", C.c_void_p), ("msd_counter", C.c_ulong)]


def _bind(lib, cls):
    lib.sep_init_sampler.restype = cls
    lib.sep_init_sampler.argtypes = []
    lib.sep_add_mol_sampler.argtypes = [C.POINTER(cls), C.c_void_p]
    lib.sep_sample.argtypes = [C.POINTER(capi.SepPart), C.POINTER(cls), C.POINTER(capi.SepRet), capi.SepSys, C.c_uint]
    lib.sep_close_sampler.argtypes = [C.POINTER(cls)]
    lib.sep_add_sampler.restype = None
    lib.sep_add_sampler.argtypes = None     # variadic: explicit ctypes objects at the call sites


def _add(lib, sampler, name, sys_, lvec, *rest):
    lib.sep_add_sampler(C.byref(sampler), name, sys_, C.c_int(lvec), *rest)


def _compare_dirs(a, b, files):
    for f in files:
        pa, pb = os.path.join(a, f), os.path.join(b, f)
        assert os.path.exists(pa), f"reference did not write {f}"
        assert os.path.exists(pb), f"libsep.so did not write {f}"
        if f == "radial_info.dat":
            assert open(pa).read() == open(pb).read()
            continue
        A, B = np.loadtxt(pa, ndmin=2), np.loadtxt(pb, ndmin=2)
        assert A.shape == B.shape and A.size > 0, (f, A.shape, B.shape)
        assert np.nanmax(np.abs(A)) > 0, f
        ok = np.isclose(A, B, rtol=0, atol=2e-6) | (np.isnan(A) & np.isnan(B))
        assert ok.all(), (f, np.nanmax(np.abs(A - B)))


@pytest.fixture
def libs():
    ref = cm.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libsep_ref.so not built")
    ours = capi.load()
    return ref, ours


def test_atomic_samplers_match_reference(libs, tmp_path):
    ref, ours = libs
    da, db = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(da); os.makedirs(db)
    x, L = cm.lattice(8, 0.8, jitter=0.05, seed=31)
    v = cm.velocities(len(x), 1.2, seed=32)
    s = cm.ApiSystem(ref, x, L, 2.5, 0.005, v=v, nneighb=3000)
    s.view["type"][: len(x) // 4] = ord("B")                 # two species for the partial g(r)
    _bind(ref, RefSampler); _bind(ours, OurSampler)
    cwd = os.getcwd()
    try:
        os.chdir(da)
        sr = ref.sep_init_sampler()
        so = ours.sep_init_sampler()
        for lib, smp, d in ((ref, sr, da), (ours, so, db)):
            os.chdir(d)
            _add(lib, smp, b"vacf", s.sys, 20, C.c_double(1.0))
            _add(lib, smp, b"sacf", s.sys, 10, C.c_double(0.5))
            _add(lib, smp, b"msd", s.sys, 15, C.c_double(1.5), C.c_int(3), C.c_int(ord("A")))
            _add(lib, smp, b"profs", s.sys, 10, C.c_int(ord("A")), C.c_int(2))
            _add(lib, smp, b"radial", s.sys, 50, C.c_int(50), C.c_char_p(b"AB"))
            _add(lib, smp, b"gh", s.sys, 10, C.c_double(0.5), C.c_int(3))
        fun = s.fun("sep_lj_shift")
        alpha = C.c_double(0.1)
        for n in range(650):
            ref.sep_reset_retval(s.R); ref.sep_reset_force(s.atoms, s.S)
            ref.sep_force_pairs(s.atoms, b"AA", 2.5, fun, s.S, s.R, 1)
            ref.sep_force_pairs(s.atoms, b"AB", 2.5, fun, s.S, s.R, 1)
            ref.sep_force_pairs(s.atoms, b"BB", 2.5, fun, s.S, s.R, 1)
            ref.sep_nosehoover(s.atoms, 1.2, C.byref(alpha), 0.1, s.S)
            ref.sep_leapfrog(s.atoms, s.S, s.R)
            os.chdir(da); ref.sep_sample(s.atoms, C.byref(sr), s.R, s.sys, n)
            os.chdir(db); ours.sep_sample(s.atoms, C.byref(so), s.R, s.sys, n)
        ours.sep_close_sampler(C.byref(so))
    finally:
        os.chdir(cwd)
        s.close()
    _compare_dirs(da, db, ["vacf.dat", "sacf.dat", "msd-k.dat", "msd.dat", "msd-gaussparam.dat", "msd-incoherent.dat",
                           "profs.dat", "radial_info.dat", "radial.dat", "gh-wavevector.dat", "gh-trans-momentum-acf.dat",
                           "gh-long-momentum-acf.dat", "gh-rho-acf.dat", "gh-energy-acf.dat", "gh-rho-energy-ccf.dat",
                           "gh-rho-long-momentum-ccf.dat", "gh-energy-long-momentum-ccf.dat", "gh-energy-rho-ccf.dat",
                           "gh-long-momentum-rho-ccf.dat", "gh-long-momentum-energy-ccf.dat", "gh-X-acf.dat"])


def test_log_spaced_msd_matches_reference(libs, tmp_path):
    ref, ours = libs
    da, db = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(da); os.makedirs(db)
    x, L = cm.lattice(6, 0.8, jitter=0.05, seed=33)
    v = cm.velocities(len(x), 1.0, seed=34)
    s = cm.ApiSystem(ref, x, L, 2.5, 0.005, v=v, update=capi.SEP_BRUTE, nneighb=0)
    _bind(ref, RefSampler); _bind(ours, OurSampler)
    cwd = os.getcwd()
    try:
        sr, so = ref.sep_init_sampler(), ours.sep_init_sampler()
        os.chdir(da); _add(ref, sr, b"msd", s.sys, 0, C.c_double(0.5), C.c_int(2), C.c_int(ord("A")))
        os.chdir(db); _add(ours, so, b"msd", s.sys, 0, C.c_double(0.5), C.c_int(2), C.c_int(ord("A")))
        fun = s.fun("sep_lj_shift")
        for n in range(400):
            ref.sep_reset_retval(s.R); ref.sep_reset_force(s.atoms, s.S)
            ref.sep_force_pairs(s.atoms, b"AA", 2.5, fun, s.S, s.R, 1)
            ref.sep_leapfrog(s.atoms, s.S, s.R)
            os.chdir(da); ref.sep_sample(s.atoms, C.byref(sr), s.R, s.sys, n)
            os.chdir(db); ours.sep_sample(s.atoms, C.byref(so), s.R, s.sys, n)
        ours.sep_close_sampler(C.byref(so))
    finally:
        os.chdir(cwd)
        s.close()
    _compare_dirs(da, db, ["msd-k.dat", "msd.dat", "msd-gaussparam.dat", "msd-incoherent.dat"])


def test_molecular_samplers_match_reference(libs, tmp_path):
    ref, ours = libs
    da, db = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(da); os.makedirs(db)
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    top = str(tmp_path / "butane.top")
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;generated\n")
        for (a, b, t) in g["blist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
        fh.write("\n[ angles ]\n;generated\n")
        for (a, b, c, t) in g["alist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
        fh.write("\n[ dihedrals ]\n;generated\n")
        for (a, b, c, d, t) in g["dlist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {d} {t}\n")
    n = len(g["x0"])
    s = cm.ApiSystem(ref, g["x0"], g["L"], 2.5, 0.001, v=g["v0"], types=np.full(n, ord("C"), dtype=np.uint8), nneighb=3000)
    ref.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q")
    mols = ref.sep_init_mol(s.atoms, s.S)
    _bind(ref, RefSampler); _bind(ours, OurSampler)
    rb = (C.c_double * 6)(*g["rb"])
    alpha = C.c_double(0.1)
    cwd = os.getcwd()
    try:
        sr, so = ref.sep_init_sampler(), ours.sep_init_sampler()
        ref.sep_add_mol_sampler(C.byref(sr), mols); ours.sep_add_mol_sampler(C.byref(so), mols)
        for lib, smp, d in ((ref, sr, da), (ours, so, db)):
            os.chdir(d)
            _add(lib, smp, b"msacf", s.sys, 8, C.c_double(0.04))
            _add(lib, smp, b"mvacf", s.sys, 8, C.c_double(0.04))
            _add(lib, smp, b"mgh", s.sys, 8, C.c_double(0.04), C.c_int(2), C.c_int(1))
        fun = s.fun("sep_lj_shift")
        for step in range(90):
            ref.sep_reset_retval(s.R); ref.sep_reset_force(s.atoms, s.S); ref.sep_reset_force_mol(s.S)
            ref.sep_force_pairs(s.atoms, b"CC", 2.5, fun, s.S, s.R, 3)
            ref.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
            ref.sep_angle_harmonic(s.atoms, 0, 1.90, 400.0, s.S, s.R)
            ref.sep_torsion_Ryckaert(s.atoms, 0, rb, s.S, s.R)
            ref.sep_nosehoover(s.atoms, 4.0, C.byref(alpha), 0.1, s.S)
            ref.sep_leapfrog(s.atoms, s.S, s.R)
            os.chdir(da); ref.sep_sample(s.atoms, C.byref(sr), s.R, s.sys, step)
            os.chdir(db); ours.sep_sample(s.atoms, C.byref(so), s.R, s.sys, step)
        ours.sep_close_sampler(C.byref(so))
    finally:
        os.chdir(cwd)
        s.close()
    _compare_dirs(da, db, ["msacf.dat", "mvacf.dat", "mgh-wavevector.dat"] + MGH_FILES)


# mgh-energy-acf.dat is not compared: the reference never initialises sepmgh.avekin (sep_mgh_init,
# source/sepsampler.c:1110-1190, unlike sep_gh_init :835), so its energy fluctuation is taken about heap garbage
MGH_FILES = ["mgh-trans-momentum-acf.dat", "mgh-long-momentum-acf.dat", "mgh-rho-acf.dat",
             "mgh-trans-angmomentum-acf.dat", "mgh-long-angmomentum-acf.dat", "mgh-momentum-angmomentum-ccf.dat",
             "mgh-dipole-acf.dat"]


def test_mgh_unsafe_mode_on_water_matches_reference(libs, tmp_path):
    """charged, non-linear molecules: dipoles and angular velocities (w = I^-1 s) enter the 'unsafe' mgh sampler"""
    ref, ours = libs
    da, db = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(da); os.makedirs(db)
    g = np.load(os.path.join(cm.GOLDEN, "water_dense_n648.npz"))
    top = str(tmp_path / "water.top")
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;generated\n")
        for (a, b, t) in g["blist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
        fh.write("\n[ angles ]\n;generated\n")
        for (a, b, c, t) in g["alist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
    s = cm.ApiSystem(ref, g["x0"], g["L"], 2.9, 5.0e-4, update=capi.SEP_BRUTE, v=g["v0"], types=g["type"], m=g["m"], z=g["z"], nneighb=0)
    ref.sep_read_topology_file(s.atoms, top.encode(), s.S, b"q")
    mols = ref.sep_init_mol(s.atoms, s.S)
    _bind(ref, RefSampler); _bind(ours, OurSampler)
    alpha = (C.c_double * 3)(0.1, 0.0, 0.0)
    cwd = os.getcwd()
    try:
        sr, so = ref.sep_init_sampler(), ours.sep_init_sampler()
        ref.sep_add_mol_sampler(C.byref(sr), mols); ours.sep_add_mol_sampler(C.byref(so), mols)
        for lib, smp, d in ((ref, sr, da), (ours, so, db)):
            os.chdir(d)
            _add(lib, smp, b"mgh", s.sys, 6, C.c_double(0.015), C.c_int(2), C.c_int(0))
        fun = s.fun("sep_lj_shift")
        for step in range(40):
            ref.sep_reset_retval(s.R); ref.sep_reset_force(s.atoms, s.S)
            ref.sep_force_pairs(s.atoms, b"OO", 2.5, fun, s.S, s.R, 3)
            ref.sep_stretch_harmonic(s.atoms, 0, 0.316, 68421.0, s.S, s.R)
            ref.sep_angle_cossq(s.atoms, 0, 1.97, 490.0, s.S, s.R)
            ref.sep_coulomb_sf(s.atoms, 2.9, s.S, s.R, 3)
            ref.sep_nosehoover(s.atoms, 3.81, alpha, 0.01, s.S)
            ref.sep_leapfrog(s.atoms, s.S, s.R)
            os.chdir(da); ref.sep_sample(s.atoms, C.byref(sr), s.R, s.sys, step)
            os.chdir(db); ours.sep_sample(s.atoms, C.byref(so), s.R, s.sys, step)
        ours.sep_close_sampler(C.byref(so))
    finally:
        os.chdir(cwd)
        s.close()
    _compare_dirs(da, db, ["mgh-wavevector.dat"] + MGH_FILES + ["mgh-X-cf.dat"])
