#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libsep_ref.so, compiled from the
unmodified sources under /root/reference with -O2 -fno-fast-math -ffp-contract=off).

Run in the build container only (needs /root/reference for the molecular start files):
    make -C oracle ref && python tests/golden/make_golden.py

Each fixture stores an evolved (thermalised) state -- not a copy of any reference file -- and what the
reference computes from it: the neighbour pair set read from ptr[i].neighb, forces after each force
routine, the sepret sums, and the state after sep_nosehoover + sep_leapfrog.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common as cm  # noqa: E402
from seplib_b200 import capi  # noqa: E402

REFROOT = "/root/reference"


def snap_state(s):
    v = s.view
    return dict(x=v["x"].copy(), v=v["v"].copy(), f=v["f"].copy(), xn=v["xn"].copy(),
                cross_neighb=v["cross_neighb"].copy(), crossings=v["crossings"].copy())


def lj_fixture(lib):
    """prg1/prg4-style LJ: 1000 atoms, rho 0.8, rc 2.5, NH; 150 reference steps to leave the lattice."""
    x, L = cm.lattice(10, 0.8)
    v = cm.velocities(len(x), 1.2, seed=77)
    cf, dt, skin, temp, tau = 2.5, 0.005, 0.25, 1.2, 0.1
    s = cm.ApiSystem(lib, x, L, cf, dt, v=v)
    alpha = C.c_double(0.1)
    fun = s.fun("sep_lj_shift")
    for _ in range(150):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), tau, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
    out = dict(L=L, cf=cf, dt=dt, skin=skin, temp=temp, tau=tau, alpha0=alpha.value)
    st = snap_state(s)
    out.update({"x0": st["x"], "v0": st["v"], "xn0": st["xn"], "cn0": st["cross_neighb"], "cr0": st["crossings"],
                "neighb_flag0": s.sys.neighb_flag})
    # one fully recorded step, list rebuilt first
    s.sys.neighb_flag = 1
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
    out["pairs"] = cm.pair_set(s.neighb_pairs()).astype(np.int32)
    out["f_pairs"] = s.view["f"].copy()
    r = s.ret_arrays(); out["epot"] = r["epot"]; out["pot_P"] = r["pot_P"]
    lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), tau, s.S)
    out["f_nh"] = s.view["f"].copy(); out["alpha1"] = alpha.value
    lib.sep_leapfrog(s.atoms, s.S, s.R)
    st = snap_state(s)
    out.update({"x1": st["x"], "v1": st["v"], "xn1": st["xn"], "cn1": st["cross_neighb"], "cr1": st["crossings"]})
    r = s.ret_arrays(); out["ekin"] = r["ekin"]; out["kin_P"] = r["kin_P"]
    out["max_dist2"] = s.sys.max_dist2; out["neighb_flag1"] = s.sys.neighb_flag
    # 40 more steps: aggregate trajectory
    traj = []
    for _ in range(40):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), tau, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        lib.sep_pressure_tensor(s.R, s.S)
        traj.append([s.ret.epot, s.ret.ekin, s.ret.p, alpha.value, s.sys.nupdate_neighb])
    out["traj"] = np.array(traj)
    out["x41"] = s.view["x"].copy()
    # the sep_force_lj variant and the other pair functions on the x0 state
    s2 = cm.ApiSystem(lib, out["x0"], L, cf, dt, v=out["v0"])
    for name, fn in (("lj", "sep_lj"), ("wca", "sep_wca")):
        lib.sep_reset_retval(s2.R); lib.sep_reset_force(s2.atoms, s2.S)
        rc = 2.5 if name == "lj" else 2.0 ** (1.0 / 6.0)
        lib.sep_force_pairs(s2.atoms, b"AA", rc, s2.fun(fn), s2.S, s2.R, 1)
        out[f"f_{name}"] = s2.view["f"].copy(); out[f"epot_{name}"] = s2.ret.epot
        out[f"pot_P_{name}"] = s2.ret_arrays()["pot_P"]
    par = (C.c_double * 4)(2.2, 0.8, 1.05, 0.7)
    lib.sep_reset_retval(s2.R); lib.sep_reset_force(s2.atoms, s2.S)
    lib.sep_force_lj(s2.atoms, b"AA", par, s2.S, s2.R, 1)
    out["ljparam"] = np.array(par[:]); out["f_ljparam"] = s2.view["f"].copy(); out["epot_ljparam"] = s2.ret.epot
    s.close(); s2.close()
    np.savez_compressed(os.path.join(HERE, "lj_n1000.npz"), **out)
    print("lj_n1000: pairs", len(out["pairs"]), "epot/N", out["epot"] / 1000)


def butane_fixture(lib):
    """prg2's system (4000 united atoms, 1000 butane chains), evolved 60 steps; every force routine recorded."""
    cf, dt, temp = 2.5, 0.001, 4.0
    rb = (C.c_double * 6)(15.5000, 20.3050, -21.9170, -5.1150, 43.8340, -52.6070)
    s = cm.ApiSystem.from_xyz(lib, f"{REFROOT}/test/prg1.xyz", f"{REFROOT}/test/prg1.top", cf, dt, capi.SEP_LLIST_NEIGHBLIST)
    alpha = C.c_double(0.1)
    fun = s.fun("sep_lj_shift")

    def forces():
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"CC", cf, fun, s.S, s.R, 3)
        lib.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
        lib.sep_angle_harmonic(s.atoms, 0, 1.90, 400.0, s.S, s.R)
        lib.sep_torsion_Ryckaert(s.atoms, 0, rb, s.S, s.R)

    for _ in range(60):
        forces()
        lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), 0.1, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
    out = dict(L=np.array(s.sys.length[:]), cf=cf, dt=dt, temp=temp, rb=np.array(rb[:]), alpha0=alpha.value)
    st = snap_state(s)
    out.update({"x0": st["x"], "v0": st["v"], "xn0": st["xn"], "cn0": st["cross_neighb"], "cr0": st["crossings"]})
    t = s.topo()
    out.update(molindex=t.molindex, bond=t.bond, angle=t.angle, dihed=t.dihed, blist=t.blist, alist=t.alist, dlist=t.dlist)
    mp = s.sys.molptr.contents
    s.sys.neighb_flag = 1
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"CC", cf, fun, s.S, s.R, 3)
    out["pairs_same_mol"] = cm.pair_set(s.neighb_pairs()).astype(np.int32)
    out["f_lj"] = s.view["f"].copy(); out["epot_lj"] = s.ret.epot; out["pot_P_lj"] = s.ret_arrays()["pot_P"]
    lib.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
    out["f_bond"] = s.view["f"].copy(); out["epot_bond"] = s.ret.epot
    r = s.ret_arrays(); out["pot_P_bond_total"] = r["pot_P"]; out["pot_P_bond"] = r["pot_P_bond"]
    out["blengths"] = np.ctypeslib.as_array(mp.blengths, shape=(mp.num_bonds,)).copy()
    lib.sep_angle_harmonic(s.atoms, 0, 1.90, 400.0, s.S, s.R)
    out["f_angle"] = s.view["f"].copy(); out["epot_angle"] = s.ret.epot
    out["angles"] = np.ctypeslib.as_array(mp.angles, shape=(mp.num_angles,)).copy()
    lib.sep_torsion_Ryckaert(s.atoms, 0, rb, s.S, s.R)
    out["f_torsion"] = s.view["f"].copy(); out["epot_torsion"] = s.ret.epot
    out["dihedrals"] = np.ctypeslib.as_array(mp.dihedrals, shape=(mp.num_dihedrals,)).copy()
    lib.sep_nosehoover(s.atoms, temp, C.byref(alpha), 0.1, s.S)
    lib.sep_leapfrog(s.atoms, s.S, s.R)
    st = snap_state(s)
    out.update({"x1": st["x"], "v1": st["v"], "alpha1": alpha.value, "ekin": s.ret.ekin})
    # the bonded-exclusion list (bond+angle+dihedral partners) on the x0 state
    s.view["x"][:] = out["x0"]
    s.sys.neighb_flag = 1
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"CC", cf, fun, s.S, s.R, 2)
    out["pairs_nonbonded"] = cm.pair_set(s.neighb_pairs()).astype(np.int32)
    out["f_lj_nonbonded"] = s.view["f"].copy()
    # cos^2 angle potential (prg3 uses it for water; recorded here on the chains as well)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_angle_cossq(s.atoms, 0, 1.90, 400.0, s.S, s.R)
    out["f_cossq"] = s.view["f"].copy(); out["epot_cossq"] = s.ret.epot
    s.close()
    np.savez_compressed(os.path.join(HERE, "butane_n4000.npz"), **out)
    print("butane_n4000: pairs", len(out["pairs_same_mol"]), len(out["pairs_nonbonded"]), "epot/N", out["epot_torsion"] / 4000)


def water_fixture(lib, dense=False):
    """prg3's system (648 atoms, SPC/Fw water, SEP_BRUTE): LJ OO + bonds + cos^2 angles + SF Coulomb.
    dense=True: the compressed state prg3 ends in (rho = 3.15, the reference's cuda/start_water.xyz, which
    carries velocities) -- the unit cell of the C3 benchmark workload."""
    cf, dt, temp = 2.9, 5.0e-4, 3.81
    if dense:
        s = cm.ApiSystem.from_xyz(lib, f"{REFROOT}/cuda/start_water.xyz", f"{REFROOT}/cuda/start_water.top", cf, dt, capi.SEP_BRUTE)
    else:
        s = cm.ApiSystem.from_xyz(lib, f"{REFROOT}/test/prg2.xyz", f"{REFROOT}/test/prg2.top", cf, dt, capi.SEP_BRUTE)
        lib.sep_set_vel_seed(s.atoms, temp, 42, s.sys)
    alpha = (C.c_double * 3)(0.1, 0.0, 0.0)
    fun = s.fun("sep_lj_shift")

    def forces():
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"OO", 2.5, fun, s.S, s.R, 3)
        lib.sep_stretch_harmonic(s.atoms, 0, 0.316, 68421.0, s.S, s.R)
        lib.sep_angle_cossq(s.atoms, 0, 1.97, 490.0, s.S, s.R)
        lib.sep_coulomb_sf(s.atoms, cf, s.S, s.R, 3)

    for _ in range(40):
        forces()
        lib.sep_nosehoover(s.atoms, temp, alpha, 0.01, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
    out = dict(L=np.array(s.sys.length[:]), cf=cf, dt=dt, temp=temp, alpha0=alpha[0])
    st = snap_state(s)
    out.update({"x0": st["x"], "v0": st["v"], "xn0": st["xn"], "cn0": st["cross_neighb"], "cr0": st["crossings"]})
    out["type"] = s.view["type"].copy(); out["m"] = s.view["m"].copy(); out["z"] = s.view["z"].copy()
    t = s.topo()
    out.update(molindex=t.molindex, bond=t.bond, angle=t.angle, dihed=t.dihed, blist=t.blist, alist=t.alist, dlist=t.dlist)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_force_pairs(s.atoms, b"OO", 2.5, fun, s.S, s.R, 3)
    out["f_lj"] = s.view["f"].copy(); out["epot_lj"] = s.ret.epot; out["pot_P_lj"] = s.ret_arrays()["pot_P"]
    lib.sep_stretch_harmonic(s.atoms, 0, 0.316, 68421.0, s.S, s.R)
    out["f_bond"] = s.view["f"].copy(); out["epot_bond"] = s.ret.epot
    lib.sep_angle_cossq(s.atoms, 0, 1.97, 490.0, s.S, s.R)
    out["f_angle"] = s.view["f"].copy(); out["epot_angle"] = s.ret.epot
    lib.sep_coulomb_sf(s.atoms, cf, s.S, s.R, 3)
    out["f_coul"] = s.view["f"].copy(); out["epot_coul"] = s.ret.epot; out["ecoul"] = s.ret.ecoul
    out["pot_P_total"] = s.ret_arrays()["pot_P"]
    lib.sep_nosehoover(s.atoms, temp, alpha, 0.01, s.S)
    lib.sep_leapfrog(s.atoms, s.S, s.R)
    st = snap_state(s)
    out.update({"x1": st["x"], "v1": st["v"], "alpha1": alpha[0], "ekin": s.ret.ekin})
    # brute Coulomb with the bonded rule (the "== 1" quirk, source/sepcoulomb.c:37)
    s.view["x"][:] = out["x0"]
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
    lib.sep_coulomb_sf(s.atoms, cf, s.S, s.R, 2)
    out["f_coul_bonded"] = s.view["f"].copy(); out["ecoul_bonded"] = s.ret.ecoul
    s.close()
    name = "water_dense_n648" if dense else "water_n648"
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, ": L", out["L"], "epot/mol", out["epot_coul"] / 216, "ecoul", out["ecoul"])


def dpd_fixture(lib):
    """sep_verlet_dpd only (the force's glibc rand() stream is not reproducible in parallel): two calls."""
    x, L = cm.lattice(8, 3.0, jitter=0.3, seed=5)
    n = len(x)
    rng = np.random.default_rng(9)
    v = cm.velocities(n, 1.0, seed=6)
    s = cm.ApiSystem(lib, x, L, 1.0, 0.02, v=v)
    out = dict(L=L, dt=0.02, x0=x, v0=v)
    for step in range(2):
        f = rng.normal(size=(n, 3)) * 5.0
        s.view["f"][:] = f
        s.sys.max_dist2 = 0.0
        s.sys.neighb_flag = 0           # record whether THIS call fired the trigger
        lib.sep_reset_retval(s.R)
        lib.sep_verlet_dpd(s.atoms, 0.5, step, s.S, s.R)
        out[f"f{step}"] = f
        out[f"x{step + 1}"] = s.view["x"].copy(); out[f"v{step + 1}"] = s.view["v"].copy()
        out[f"pv{step + 1}"] = s.view["pv"].copy(); out[f"pa{step + 1}"] = s.view["pa"].copy()
        out[f"ekin{step + 1}"] = s.ret.ekin; out[f"flag{step + 1}"] = s.sys.neighb_flag
        out[f"cr{step + 1}"] = s.view["crossings"].copy()
    s.close()
    np.savez_compressed(os.path.join(HERE, "dpd_n512.npz"), **out)
    print("dpd_n512 ok")


def dpd_force_fixture(lib):
    """sep_force_dpd of the REFERENCE with rand() interposed to a constant (tests/golden/rand_shim.c, LD_PRELOAD): every
    pair draws sep_rand() = 0.75.  Pins the conservative, dissipative and random terms (list and brute variants,
    source/sepprfrc.c:1007-1231) to the reference itself; the device and the oracle reproduce the draw with
    SEPGPU_DPD_SEED_FIXED.  Re-executes this script under the shim."""
    import subprocess
    if os.environ.get("SEP_GOLDEN_SHIM") != "1":
        shim = os.path.join(os.path.dirname(HERE), "_build", "librandshim.so")
        os.makedirs(os.path.dirname(shim), exist_ok=True)
        subprocess.check_call(["gcc", "-shared", "-fPIC", "-O2", "-o", shim, os.path.join(HERE, "rand_shim.c")])
        env = dict(os.environ, LD_PRELOAD=shim, SEP_GOLDEN_SHIM="1")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "dpd_force"], env=env)
        return
    assert C.CDLL(None).rand() == 1610612736, "rand() shim not active"
    cf, dt, aij, temp, sigma = 1.0, 0.02, 25.0, 1.0, 3.0
    out = dict(cf=cf, dt=dt, aij=aij, temp=temp, sigma=sigma)
    for tag, ncell, update in (("list", 8, capi.SEP_LLIST_NEIGHBLIST), ("brute", 6, capi.SEP_BRUTE)):
        x, L = cm.lattice(ncell, 3.0, jitter=0.35, seed=41)
        n = len(x)
        pv = cm.velocities(n, 1.0, seed=42)
        s = cm.ApiSystem(lib, x, L, cf, dt, update=update)
        s.view["pv"][:] = pv
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_dpd(s.atoms, b"AA", cf, aij, temp, sigma, s.S, s.R, 1)
        out.update({f"{tag}_x": x, f"{tag}_pv": pv, f"{tag}_L": L, f"{tag}_f": s.view["f"].copy(), f"{tag}_epot": s.ret.epot})
        s.close()
    np.savez_compressed(os.path.join(HERE, "dpd_force_n512.npz"), **out)
    print("dpd_force_n512: list epot", out["list_epot"], "brute epot", out["brute_epot"], "|f|max", np.abs(out["list_f"]).max())


def nvt1000_fixture(lib):
    """1000 steps of the prg1-style NVT loop on the reference (N = 4096, rho 0.8, rc 2.5): epot/N, ekin/N, T, p and the
    list-update counter every 100 steps (trajectories compare on aggregates)."""
    x, L = cm.lattice(16, 0.8, jitter=0.05, seed=61)
    v = cm.velocities(len(x), 1.0, seed=62)
    rows = cm.drive_nvt_1000(lib, x, v, L)
    np.savez_compressed(os.path.join(HERE, "nvt1000_n4096.npz"), x0=x, v0=v, L=L, rows=rows)
    print("nvt1000_n4096:", rows[0], rows[-1])


def fij_array(s):
    """sys->molptr->Fij (float ***) as an (nmol, nmol, 3) float32 array."""
    mp = s.sys.molptr.contents
    nm = mp.num_mols
    rows = C.cast(mp.Fij, C.POINTER(C.POINTER(C.POINTER(C.c_float))))
    out = np.zeros((nm, nm, 3), dtype=np.float32)
    for i in range(nm):
        ri = rows[i]
        for j in range(nm):
            out[i, j] = ri[j][0:3]
    return out


def molpress_fixture(lib):
    """Molecule-molecule force table and molecular pressure tensor (sep_init_mol, sep_reset_force_mol,
    sep_mol_pressure_tensor) on the recorded butane (list mode) and water (brute, with SF Coulomb) states."""
    out = {}
    g = np.load(os.path.join(HERE, "butane_n4000.npz"))
    s = cm.ApiSystem.from_xyz(lib, f"{REFROOT}/test/prg1.xyz", f"{REFROOT}/test/prg1.top", 2.5, 0.001, capi.SEP_LLIST_NEIGHBLIST)
    s.view["x"][:] = g["x0"]; s.view["v"][:] = g["v0"]; s.view["crossings"][:] = g["cr0"]
    mols = lib.sep_init_mol(s.atoms, s.S)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S); lib.sep_reset_force_mol(s.S)
    lib.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    lib.sep_mol_pressure_tensor(s.atoms, mols, s.R, s.S)
    F = fij_array(s)
    nz = np.argwhere(np.abs(F).sum(axis=2) > 0).astype(np.int32)
    out.update(butane_fij_idx=nz, butane_fij_val=F[nz[:, 0], nz[:, 1]], butane_p_mol=s.ret.p_mol,
               butane_P_mol=np.array([list(r) for r in s.ret.P_mol]), butane_kin_P_mol=np.array([list(r) for r in s.ret.kin_P_mol]),
               butane_pot_P_mol=np.array([list(r) for r in s.ret.pot_P_mol]))
    print("molpress butane: nonzero Fij", len(nz), "p_mol", s.ret.p_mol)
    lib.sep_free_mol(mols, s.S)
    s.close()

    g = np.load(os.path.join(HERE, "water_n648.npz"))
    s = cm.ApiSystem.from_xyz(lib, f"{REFROOT}/test/prg2.xyz", f"{REFROOT}/test/prg2.top", 2.9, 5.0e-4, capi.SEP_BRUTE)
    s.view["x"][:] = g["x0"]; s.view["v"][:] = g["v0"]; s.view["crossings"][:] = g["cr0"]
    mols = lib.sep_init_mol(s.atoms, s.S)
    lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S); lib.sep_reset_force_mol(s.S)
    lib.sep_force_pairs(s.atoms, b"OO", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    lib.sep_coulomb_sf(s.atoms, 2.9, s.S, s.R, 3)
    lib.sep_mol_pressure_tensor(s.atoms, mols, s.R, s.S)
    F = fij_array(s)
    out.update(water_fij=F, water_p_mol=s.ret.p_mol, water_P_mol=np.array([list(r) for r in s.ret.P_mol]),
               water_kin_P_mol=np.array([list(r) for r in s.ret.kin_P_mol]),
               water_pot_P_mol=np.array([list(r) for r in s.ret.pot_P_mol]))
    print("molpress water: p_mol", s.ret.p_mol)
    lib.sep_free_mol(mols, s.S)
    s.close()
    np.savez_compressed(os.path.join(HERE, "molpress.npz"), **out)


def next_rows_fixture(lib):
    """Box-changing callers, sep_relax_temp and sep_force_x0, driven by the
    shared loops in tests/common.py on the reference build."""
    out = {}
    x, L = cm.lattice(11, 0.8, jitter=0.08, seed=21)
    v = cm.velocities(len(x), 1.0, seed=22)
    out.update(c_x0=x, c_v0=v, c_L=L)
    for k, val in cm.drive_compress(lib, x, v, L).items():
        out["compress_" + k] = val
    for k, val in cm.drive_berendsen(lib, x, v, L, steps=12, iso=True, update=capi.SEP_LLIST_NEIGHBLIST).items():
        out["beriso_" + k] = val
    xb, Lb = cm.lattice(6, 0.844, jitter=0.05, seed=23)
    vb = cm.velocities(len(xb), 0.728, seed=24)
    out.update(b_x0=xb, b_v0=vb, b_L=Lb)
    for k, val in cm.drive_berendsen(lib, xb, vb, Lb).items():
        out["ber_" + k] = val
    for k, val in cm.drive_slit(lib, x, v, L).items():
        out["slit_" + k] = val
    # stochastic integrators: each in its own process-like state -- the reference's sep_randn caches a deviate, so
    # the two runs draw an even number of deviates (3 * 216 * 30) and leave the cache empty for the next
    for which in ("fp", "gjf"):
        for k, val in cm.drive_stochastic(lib, xb, vb, Lb, which).items():
            out[which + "_" + k] = val
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
    print("next_rows: compress L", out["compress_traj"][0, 2], "->", out["compress_traj"][-1, 2], "cells", out["compress_traj"][0, 3], "->",
          out["compress_traj"][-1, 3], "| berendsen Lz", out["ber_traj"][0, 3], "->", out["ber_traj"][-1, 3], "p", out["ber_traj"][-1, 2],
          "| slit ekin", out["slit_traj"][-1, 1], "walls", int((out["slit_types"] == ord("W")).sum()))


if __name__ == "__main__":
    lib = cm.ref()
    if lib is None:
        sys.exit("oracle/_ref/libsep_ref.so missing: run `make -C oracle ref` first")
    which = sys.argv[1:] or ["lj", "butane", "water", "dpd", "dpd_force", "nvt1000", "molpress", "water_dense", "next_rows"]
    for name in which:
        {"lj": lj_fixture, "butane": butane_fixture, "water": water_fixture, "dpd": dpd_fixture, "dpd_force": dpd_force_fixture, "nvt1000": nvt1000_fixture,
         "molpress": molpress_fixture, "water_dense": lambda l: water_fixture(l, dense=True),
         "next_rows": next_rows_fixture}[name](lib)
