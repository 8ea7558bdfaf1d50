"""Decomposed butane: the checks shared by the emulator run (tests/emu/dd_threads.py butane: ranks are threads) and the
hardware run (tests/dd_check.py with DD_MOL=butane: ranks are processes under torchrun).

The system is the reference's own evolved 4000-atom butane cell (tests/golden/butane_n4000.npz: 5 cell layers -> slabs of
2 and 3).  Every rank uploads the atoms of its slab with their global ids, molecule indices and rebuild-time state, sets the
GLOBAL topology lists, and runs prg2's force sequence (reference prgs/prg2.c:60-72): Lennard-Jones "CC" with same-molecule
exclusion, bond stretch, angle, Ryckaert torsion, Nose-Hoover, leapfrog.  Checked against the golden vectors (what the
REFERENCE computed from this state): accumulated forces after each routine by global id (1e-10), energies summed over the
ranks, positions / velocities / thermostat after the step; then `nsteps` further steps against a single-domain run.
"""
import ctypes as C
import os

import numpy as np

import common as cm
from seplib_b200 import capi

FT = 1e-10
BOND = (0, 0.407, 2074.0)
ANGLE = (0, 1.90, 400.0)


def golden():
    with np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz")) as z:
        return {k: z[k] for k in z.files}          # read everything now: the lazy archive is not safe to share between threads


def setup_rank(s, g, gsys, rank, world, id_bytes):
    """dd_init + upload of this rank's atoms; returns their global ids"""
    n = len(g["x0"])
    nz = gsys.nsubbox[2]
    z0, z1 = capi.dd_slab_range(rank, world, nz)
    cz = np.clip(np.floor(g["x0"][:, 2] / gsys.lsubbox[2]).astype(np.int64), 0, nz - 1)
    mine = np.nonzero((cz >= z0) & (cz < z1))[0].astype(np.int32)
    s.dd_init(rank, world, id_bytes, gsys, n)
    s.dd_set_owned(len(mine))
    s.put(capi.F_X, g["x0"][mine]); s.put(capi.F_V, g["v0"][mine]); s.put(capi.F_GID, mine)
    s.put(capi.F_TYPE, np.full(len(mine), ord("C"), dtype=np.uint8))
    s.put(capi.F_MOLINDEX, np.ascontiguousarray(g["molindex"][mine]))
    s.put(capi.F_XN, g["xn0"][mine]); s.put(capi.F_CROSS_NEIGHB, g["cn0"][mine]); s.put(capi.F_CROSSINGS, g["cr0"][mine])
    s.set_topology(np.ascontiguousarray(g["blist"], dtype=np.uint32), np.ascontiguousarray(g["alist"], dtype=np.uint32),
                   np.ascontiguousarray(g["dlist"], dtype=np.uint32))
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    return mine


def force_sequence(s, g, gsys, p, probe=None):
    """prg2's force calls; probe(tag) is called after each routine (reads are collective in decomposed runs)"""
    rb = (C.c_double * 6)(*g["rb"])
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(gsys), b"CC", C.byref(p), cm.EXCL_SAME_MOL, 1)
    if probe: probe("lj")
    s.call("sepgpu_stretch_harmonic", C.byref(gsys), BOND[0], BOND[1], BOND[2])
    if probe: probe("bond")
    s.call("sepgpu_angle_harmonic", C.byref(gsys), ANGLE[0], ANGLE[1], ANGLE[2])
    if probe: probe("angle")
    s.call("sepgpu_torsion_ryckaert", C.byref(gsys), 0, rb)
    if probe: probe("torsion")


def rank_run(s, g, gsys, rank, world, id_bytes, nsteps):
    """Runs this rank; returns what the checker needs: per-routine (gids, forces, epot) of step 0, the state after step 0,
    the per-step records of the further steps, and the final positions."""
    setup_rank(s, g, gsys, rank, world, id_bytes)
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    temp = float(g["temp"])
    first = {}

    def probe(tag):
        n_own = s.dd_layers()[2]
        first[tag] = (s.get(capi.F_GID)[:n_own].copy(), s.get(capi.F_F)[:n_own].copy(), s.scalars().epot)

    force_sequence(s, g, gsys, p, probe)
    s.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, 0.1)
    s.call("sepgpu_leapfrog", C.byref(gsys))
    sc = s.scalars()
    n_own = s.dd_layers()[2]
    after = (s.get(capi.F_GID)[:n_own].copy(), s.get(capi.F_X)[:n_own].copy(), s.get(capi.F_V)[:n_own].copy(), sc.alpha[0], sc.ekin)
    rec = []
    for _ in range(nsteps):
        force_sequence(s, g, gsys, p)
        s.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, 0.1)
        s.call("sepgpu_leapfrog", C.byref(gsys))
        sc = s.scalars()
        rec.append((sc.epot, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.pot_P_bond[0], sc.neighb_flag, sc.nbuild))
    n_own, n_halo = s.dd_layers()[2:]
    final = (s.get(capi.F_GID)[:n_own].copy(), s.get(capi.F_X)[:n_own].copy(), n_own, n_halo)
    return {"first": first, "after": after, "rec": rec, "final": final}


def single_run(g, gsys, nsteps, device=0):
    n = len(g["x0"])
    s = capi.System(n, device=device)
    s.put(capi.F_X, g["x0"]); s.put(capi.F_V, g["v0"]); s.put(capi.F_TYPE, np.full(n, ord("C"), dtype=np.uint8))
    s.put(capi.F_MOLINDEX, np.ascontiguousarray(g["molindex"]))
    s.put(capi.F_XN, g["xn0"]); s.put(capi.F_CROSS_NEIGHB, g["cn0"]); s.put(capi.F_CROSSINGS, g["cr0"])
    s.set_topology(np.ascontiguousarray(g["blist"], dtype=np.uint32), np.ascontiguousarray(g["alist"], dtype=np.uint32),
                   np.ascontiguousarray(g["dlist"], dtype=np.uint32))
    s.call("sepgpu_set_alpha", 0, float(g["alpha0"]))
    p = capi.lj_param(float(g["cf"]), kind="lj_shift")
    rec = []
    for _ in range(nsteps + 1):
        force_sequence(s, g, gsys, p)
        s.call("sepgpu_nosehoover", C.byref(gsys), float(g["temp"]), 0, 0.1)
        s.call("sepgpu_leapfrog", C.byref(gsys))
        sc = s.scalars()
        rec.append((sc.epot, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.pot_P_bond[0], sc.neighb_flag, sc.nbuild))
    x = s.get(capi.F_X)
    s.close()
    return rec[1:], x


def check(g, results, single_rec, single_x, log=print):
    """results: one dict per rank (rank_run).  Returns True when everything agrees."""
    n = len(g["x0"])
    ok = True
    for tag, fkey, ekey in (("lj", "f_lj", "epot_lj"), ("bond", "f_bond", "epot_bond"), ("angle", "f_angle", "epot_angle"),
                            ("torsion", "f_torsion", "epot_torsion")):
        f = np.full((n, 3), np.nan)
        for r in results:
            gid, fr, _ = r["first"][tag]
            f[gid] = fr
        if np.isnan(f).any():
            log(f"{tag}: atoms missing from the union of the ranks"); ok = False; continue
        err = cm.rel_force_err(f, g[fkey])
        e = results[0]["first"][tag][2]                      # scalars are global (reduced when read)
        if err > FT or abs(e - float(g[ekey])) > FT * abs(float(g[ekey])):
            log(f"{tag}: force error {err:.2e}, epot {e!r} vs reference {float(g[ekey])!r}"); ok = False
    x = np.full((n, 3), np.nan); v = np.full((n, 3), np.nan)
    for r in results:
        gid, xr, vr, alpha, ekin = r["after"]
        x[gid] = xr; v[gid] = vr
        if abs(alpha - float(g["alpha1"])) > 1e-12 * abs(float(g["alpha1"])) or abs(ekin - float(g["ekin"])) > FT * float(g["ekin"]):
            log(f"after the step: alpha {alpha!r} / ekin {ekin!r} vs reference {float(g['alpha1'])!r} / {float(g['ekin'])!r}"); ok = False
    if not (np.abs(x - g["x1"]).max() <= 1e-11 and np.abs(v - g["v1"]).max() <= 1e-10):
        log(f"after the step: max|dx| {np.abs(x - g['x1']).max():.2e} max|dv| {np.abs(v - g['v1']).max():.2e}"); ok = False
    for step, want in enumerate(single_rec):
        tol = 1e-9 * (step + 2)
        for rk, r in enumerate(results):
            got = r["rec"][step]
            for name, a, b in zip(("epot", "ekin", "alpha", "virial", "bond virial"), got[:5], want[:5]):
                if abs(a - b) > tol * max(abs(b), 1e-3):
                    log(f"step {step + 1} rank {rk}: {name} differs: decomposed {a!r} vs single {b!r}"); ok = False
            if got[5] != want[5]:
                log(f"step {step + 1} rank {rk}: rebuild trigger differs"); ok = False
    tot = sum(r["final"][2] for r in results)
    full = np.full((n, 3), np.nan)
    for r in results:
        full[r["final"][0]] = r["final"][1]
    err = np.abs(full - single_x).max() if tot == n else float("nan")
    if tot != n or not err <= 1e-7:
        log(f"final: {tot} of {n} atoms, max|dx| vs single domain {err:.2e}"); ok = False
    builds = single_rec[-1][6] if single_rec else 0
    log(f"dd_mol butane: world={len(results)} n={n} steps=1+{len(single_rec)} builds={builds} own/halo(rank0)={results[0]['final'][2]}/{results[0]['final'][3]} "
        f"max|dx|={err:.2e} -> {'OK' if ok else 'FAIL'}")
    return ok


# ---------------------------------------------------------------------------------------------------------------------
# water (prg3's force sequence, reference prgs/prg3.c:60-72): typed Lennard-Jones on global-index rows, bonds, cos^2
# angles and the shifted-force Coulomb sum in a decomposed run.  The system is the reference's 648-atom water box tiled
# 2 x 2 x 2 (seplib_b200/workloads.tiled_molecular: 4 cell layers -> two slabs of 2); the single-domain run of the same calls is
# pinned to the reference elsewhere (tests/test_golden.py, tests/test_gpu_more.py), the decomposed run is compared with it.
# ---------------------------------------------------------------------------------------------------------------------
WATER = dict(cf=2.9, cf_lj=2.5, dt=5.0e-4, temp=3.81, tau=0.01, lbond=0.316, kbond=68421.0, angle=1.97, kangle=490.0)


def water_system():
    from seplib_b200 import workloads as wl
    t = wl.tiled_molecular("water_n648.npz", 2)              # molecules made whole before tiling; velocities of the fixture
    return dict(x=t["x"], v=t["v"], type=t["type"], z=np.ascontiguousarray(t["z"]), m=np.ascontiguousarray(t["m"]), mol=t["molindex"],
                L=t["L"], blist=t["blist"], alist=t["alist"], dlist=t["dlist"])


def water_forces(s, gsys, p):
    W = WATER
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_force_lj", C.byref(gsys), b"OO", C.byref(p), cm.EXCL_SAME_MOL, 1)
    s.call("sepgpu_stretch_harmonic", C.byref(gsys), 0, W["lbond"], W["kbond"])
    s.call("sepgpu_angle_cossq", C.byref(gsys), 0, W["angle"], W["kangle"])
    s.call("sepgpu_coulomb_sf", C.byref(gsys), W["cf"], cm.EXCL_SAME_MOL)


def water_put(s, w, rows):
    s.put(capi.F_X, w["x"][rows]); s.put(capi.F_V, w["v"][rows]); s.put(capi.F_TYPE, np.ascontiguousarray(w["type"][rows]))
    s.put(capi.F_M, np.ascontiguousarray(w["m"][rows])); s.put(capi.F_MOLINDEX, np.ascontiguousarray(w["mol"][rows]))
    s.set_topology(w["blist"], w["alist"], w["dlist"])
    s.call("sepgpu_set_alpha", 0, 0.1)


def water_rank_run(s, w, gsys, rank, world, id_bytes, nsteps):
    n = len(w["x"])
    nz = gsys.nsubbox[2]
    z0, z1 = capi.dd_slab_range(rank, world, nz)
    cz = np.clip(np.floor(w["x"][:, 2] / gsys.lsubbox[2]).astype(np.int64), 0, nz - 1)
    mine = np.nonzero((cz >= z0) & (cz < z1))[0].astype(np.int32)
    s.dd_init(rank, world, id_bytes, gsys, n)
    s.dd_set_owned(len(mine))
    s.put(capi.F_GID, mine)
    water_put(s, w, mine)
    s.call("sepgpu_dd_set_charges", w["z"].ctypes.data_as(C.POINTER(C.c_double)))
    p = capi.lj_param(WATER["cf_lj"], kind="lj_shift")
    rec = []
    first = None
    for step in range(nsteps):
        water_forces(s, gsys, p)
        if step == 0:
            n_own = s.dd_layers()[2]
            first = (s.get(capi.F_GID)[:n_own].copy(), s.get(capi.F_F)[:n_own].copy())
        s.call("sepgpu_nosehoover", C.byref(gsys), WATER["temp"], 0, WATER["tau"])
        s.call("sepgpu_leapfrog", C.byref(gsys))
        sc = s.scalars()
        rec.append((sc.epot, sc.ecoul, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.neighb_flag, sc.nbuild))
    n_own, n_halo = s.dd_layers()[2:]
    return {"first": first, "rec": rec, "final": (s.get(capi.F_GID)[:n_own].copy(), s.get(capi.F_X)[:n_own].copy(), n_own, n_halo)}


def water_single_run(w, gsys, nsteps, device=0):
    n = len(w["x"])
    s = capi.System(n, device=device)
    water_put(s, w, np.arange(n))
    s.put(capi.F_Z, w["z"])
    p = capi.lj_param(WATER["cf_lj"], kind="lj_shift")
    rec, f0 = [], None
    for step in range(nsteps):
        water_forces(s, gsys, p)
        if step == 0:
            f0 = s.get(capi.F_F).copy()
        s.call("sepgpu_nosehoover", C.byref(gsys), WATER["temp"], 0, WATER["tau"])
        s.call("sepgpu_leapfrog", C.byref(gsys))
        sc = s.scalars()
        rec.append((sc.epot, sc.ecoul, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.neighb_flag, sc.nbuild))
    x = s.get(capi.F_X)
    s.close()
    return rec, f0, x


def water_check(w, results, single, log=print):
    rec1, f1, x1 = single
    n = len(w["x"])
    ok = True
    f = np.full((n, 3), np.nan)
    for r in results:
        f[r["first"][0]] = r["first"][1]
    err = cm.rel_force_err(f, f1) if not np.isnan(f).any() else float("nan")
    if not err <= FT:
        log(f"first step: force error against the single domain {err:.2e}"); ok = False
    for step, want in enumerate(rec1):
        tol = 1e-9 * (step + 2)
        for rk, r in enumerate(results):
            got = r["rec"][step]
            for name, a, b in zip(("epot", "ecoul", "ekin", "alpha", "virial"), got[:5], want[:5]):
                if abs(a - b) > tol * max(abs(b), 1e-3):
                    log(f"step {step} rank {rk}: {name} differs: decomposed {a!r} vs single {b!r}"); ok = False
            if got[5] != want[5]:
                log(f"step {step} rank {rk}: rebuild trigger differs"); ok = False
    tot = sum(r["final"][2] for r in results)
    full = np.full((n, 3), np.nan)
    for r in results:
        full[r["final"][0]] = r["final"][1]
    dx = np.abs(full - x1).max() if tot == n else float("nan")
    if tot != n or not dx <= 1e-7:
        log(f"final: {tot} of {n} atoms, max|dx| {dx:.2e}"); ok = False
    log(f"dd_mol water: world={len(results)} n={n} steps={len(rec1)} builds={rec1[-1][6]} own/halo(rank0)={results[0]['final'][2]}/{results[0]['final'][3]} "
        f"force err={err:.1e} max|dx|={dx:.2e} -> {'OK' if ok else 'FAIL'}")
    return ok
