#!/usr/bin/env python
"""Two-rank slab decomposition on the CPU kernel emulator: the ranks are two THREADS of this process, each driving its
own sepgpu context.  Default: the library's peer-memory path (halo pushed into the neighbour's buffer with release/acquire
flags, integrator sums all-gathered through the peer blocks) -- the emulator hands out in-process "IPC" handles and the
two ranks' kernels run concurrently; SEPGPU_EMU_NO_IPC=1 (or SEPGPU_DD_P2P=0): the NCCL path (tests/emu/fake_nccl.cpp:
halo send/recv every step, all-reduced sums).  Migration at rebuilds goes through the NCCL stand-in in both.  Same checks as
tests/dd_check.py on hardware: union of the ranks' pair sets == single-domain pair set at compared rebuilds, per-step
epot / ekin / alpha / max displacement equal to 1e-9 relative, trigger steps equal, final positions by global id equal
to 1e-7, atom count conserved across migration.

    python tests/emu/dd_threads.py [ncell] [nsteps] [name=value,... options for every context]
"""
import ctypes as C
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from seplib_b200 import capi  # noqa: E402

capi.LIB_PATH = build_emu.build()
import common as cm  # noqa: E402

WORLD = int(os.environ.get("DD_WORLD", "2"))


def main_butane():
    """python tests/emu/dd_threads.py butane [nsteps]: bonded terms in a decomposed run (tests/dd_mol.py)"""
    import dd_mol
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    g = dd_mol.golden()
    gsys = capi.make_sys(list(g["L"]), float(g["cf"]), float(g["dt"]), skin=0.25)
    n = len(g["x0"])
    id_bytes = capi.dd_unique_id()
    barrier = threading.Barrier(WORLD)
    results, errs = [None] * WORLD, []

    def rank_main(rank):
        try:
            s = capi.System(int(1.6 * n / WORLD) + int(3.0 * n / gsys.nsubbox[2]) + 1024, device=0)
            results[rank] = dd_mol.rank_run(s, g, gsys, rank, WORLD, id_bytes, nsteps)
            barrier.wait()
            s.close()
        except BaseException as e:                              # noqa: BLE001 -- reported by the main thread
            import traceback
            sys.stderr.write("rank %d: %s\n" % (rank, traceback.format_exc())); sys.stderr.flush()
            errs.append((rank, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(WORLD)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        print("rank failure:", errs)
        return 1
    rec, x = dd_mol.single_run(g, gsys, nsteps)
    return 0 if dd_mol.check(g, results, rec, x) else 1


def main_water():
    """python tests/emu/dd_threads.py water [nsteps]: typed LJ + bonds + cos^2 angles + Coulomb in a decomposed run"""
    import dd_mol
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    w = dd_mol.water_system()
    gsys = capi.make_sys(list(w["L"]), dd_mol.WATER["cf"], dd_mol.WATER["dt"], skin=0.25)
    n = len(w["x"])
    id_bytes = capi.dd_unique_id()
    barrier = threading.Barrier(WORLD)
    results, errs = [None] * WORLD, []

    def rank_main(rank):
        try:
            s = capi.System(int(1.6 * n / WORLD) + int(3.0 * n / gsys.nsubbox[2]) + 1024, device=0)
            results[rank] = dd_mol.water_rank_run(s, w, gsys, rank, WORLD, id_bytes, nsteps)
            barrier.wait()
            s.close()
        except BaseException as e:                              # noqa: BLE001 -- reported by the main thread
            import traceback
            sys.stderr.write("rank %d: %s\n" % (rank, traceback.format_exc())); sys.stderr.flush()
            errs.append((rank, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(WORLD)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        print("rank failure:", errs)
        return 1
    return 0 if dd_mol.water_check(w, results, dd_mol.water_single_run(w, gsys, nsteps)) else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "butane":
        return main_butane()
    if len(sys.argv) > 1 and sys.argv[1] == "water":
        return main_water()
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    opts = dict(kv.split("=") for kv in sys.argv[3].split(",")) if len(sys.argv) > 3 and sys.argv[3] else {}
    rho, rc, skin, dt, temp, tau = 0.8, 2.5, 0.25, 0.005, 2.0, 0.1
    x, L = cm.lattice(ncell, rho, jitter=0.08, seed=5)
    n = len(x)
    v = cm.velocities(n, temp, seed=6)
    gsys = capi.make_sys([L] * 3, rc, dt, skin=skin)
    nz = gsys.nsubbox[2]
    assert nz >= 2 * WORLD and nz >= 4, "need >= 4 cell layers"
    id_bytes = capi.dd_unique_id()
    p = capi.lj_param(rc, kind="lj_shift")
    barrier = threading.Barrier(WORLD)
    shared = {"pairs": [None] * WORLD, "final": [None] * WORLD, "rec": [None] * WORLD, "err": []}
    check_steps = {0, 1} | set(range(0, nsteps, 7))

    def rank_main(rank):
        try:
            z0, z1 = capi.dd_slab_range(rank, WORLD, nz)
            cz = np.floor(x[:, 2] / gsys.lsubbox[2]).astype(np.int64)
            mine = np.nonzero((cz >= z0) & (cz < z1))[0].astype(np.int32)
            ncap = int(1.6 * n / WORLD) + int(3.0 * n / nz) + 1024
            s = capi.System(ncap, device=0)
            for k, val in opts.items():
                s.call("sepgpu_set_option", k.encode(), int(val))
            s.dd_init(rank, WORLD, id_bytes, gsys, n)
            assert s.dd_layers()[:2] == (z0, z1)
            s.dd_set_owned(len(mine))
            s.put(capi.F_X, x[mine]); s.put(capi.F_V, v[mine]); s.put(capi.F_GID, mine)
            s.call("sepgpu_set_alpha", 0, 0.1)
            rec = []
            for step in range(nsteps):
                s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
                s.call("sepgpu_force_lj", C.byref(gsys), b"AA", C.byref(p), 1, 1)
                pr = None
                if step in check_steps:                        # mid-step reads only here: the other steps run unobserved
                    sc = s.scalars()                           # (collective) -- what option step_fold leaves pending stays pending
                    pr = cm.pair_set(s.pairs(max_pairs=int(sc.npairs_listed) + 16))
                if step == 0:
                    lp = C.c_longlong(-1)
                    s.call("sepgpu_get_option", b"list_f16", C.byref(lp))
                    shared["list_f16"] = lp.value
                    s.call("sepgpu_get_option", b"dd_p2p", C.byref(lp))
                    shared["p2p"] = lp.value
                s.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, tau)
                s.call("sepgpu_leapfrog", C.byref(gsys))
                sc2 = s.scalars()                              # epot and the virial of this step's force call are still in the block
                rec.append((sc2.epot, sc2.ekin, sc2.alpha[0], sc2.max_dist2, sc2.pot_P[0], sc2.neighb_flag, sc2.nbuild, pr))
            _, _, n_own, n_halo = s.dd_layers()
            shared["final"][rank] = (s.get(capi.F_X)[:n_own].copy(), s.get(capi.F_GID)[:n_own].copy(), n_own, n_halo)
            shared["rec"][rank] = rec
            barrier.wait()
            s.close()
        except BaseException as e:                              # noqa: BLE001 -- reported by the main thread
            shared["err"].append((rank, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(WORLD)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if shared["err"]:
        print("rank failure:", shared["err"])
        return 1

    # the same system on one domain
    ref = capi.System(n, device=0)
    ref.put(capi.F_X, x); ref.put(capi.F_V, v)
    ref.call("sepgpu_set_alpha", 0, 0.1)
    ok = True
    checked = 0
    recs = shared["rec"]
    for step in range(nsteps):
        ref.call("sepgpu_reset_ret"); ref.call("sepgpu_reset_force")
        ref.call("sepgpu_force_lj", C.byref(gsys), b"AA", C.byref(p), 1, 1)
        rs = ref.scalars()
        if step in check_steps:
            want = cm.pair_set(ref.pairs())
            got = cm.pair_set(np.concatenate([r[step][7] for r in recs]))
            checked += 1
            if not np.array_equal(got, want):
                print(f"step {step}: pair sets differ: dd {len(got)} vs single {len(want)}")
                ok = False
        ref.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, tau)
        ref.call("sepgpu_leapfrog", C.byref(gsys))
        rs2 = ref.scalars()
        tol = 1e-9 * (step + 1)
        for rk, rec in enumerate(recs):
            e, k, a, md, vir, flag, nb, _ = rec[step]
            for name, got_v, want_v in (("epot", e, rs.epot), ("ekin", k, rs2.ekin), ("alpha", a, rs2.alpha[0]),
                                        ("maxd2", md, rs2.max_dist2), ("virial", vir, rs.pot_P[0])):
                if abs(got_v - want_v) > tol * max(abs(want_v), 1e-3):
                    print(f"step {step} rank {rk}: {name} differs: dd {got_v!r} vs single {want_v!r}")
                    ok = False
            if flag != rs2.neighb_flag:
                print(f"step {step} rank {rk}: trigger differs")
                ok = False
    tot = sum(f[2] for f in shared["final"])
    if tot != n:
        print(f"atom count not conserved: {tot} vs {n}")
        ok = False
    full = np.full((n, 3), np.nan)
    for xo, go, _, _ in shared["final"]:
        full[go] = xo
    err = np.abs(full - ref.get(capi.F_X)).max()
    builds = ref.scalars().nbuild
    ref.close()
    ok = ok and err <= 1e-7 and builds >= 3 and shared.get("list_f16") == int(opts.get("tile_list", 1))
    ok = ok and shared.get("p2p") == (0 if os.environ.get("SEPGPU_EMU_NO_IPC") == "1" or os.environ.get("SEPGPU_DD_P2P") == "0" else 1)
    print(f"dd_threads: world={WORLD} n={n} steps={nsteps} layers={nz} builds={builds} pair-set checks={checked} "
          f"max|dx|={err:.2e} list_f16={shared.get('list_f16')} p2p={shared.get('p2p')} own/halo(rank0)={shared['final'][0][2]}/{shared['final'][0][3]} opts={opts} -> {'OK' if ok else 'FAIL'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
