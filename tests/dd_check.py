#!/usr/bin/env python
"""Multi-GPU check of the slab domain decomposition, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dd_check.py

Every rank integrates its slab of a Lennard-Jones system for NSTEPS steps; rank 0 additionally runs the
SAME system undecomposed on its own GPU.  Checked: the union of the ranks' neighbour pair sets (global
ids) equals the single-GPU set bit for bit at every compared rebuild; per-step epot/ekin/alpha agree to
1e-9 relative; positions by global id agree to 1e-7 after the run; atom count is conserved across
migration.  Exit code 0 = pass.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm  # noqa: E402
from seplib_b200 import capi  # noqa: E402

NSTEPS = int(os.environ.get("DD_STEPS", "60"))
NCELL = int(os.environ.get("DD_NCELL", "28"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if os.environ.get("DD_MOL") in ("butane", "water"):
        run_mol = run_butane if os.environ["DD_MOL"] == "butane" else run_water
        ok = run_mol(rank, world, local, int(os.environ.get("DD_STEPS", "60")))
        dist.destroy_process_group()
        return 0 if ok else 1
    res = run(rank, world, local, NSTEPS, NCELL)
    dist.destroy_process_group()
    return 0 if res["ok"] else 1


def run_butane(rank, world, local, nsteps):
    """Bonded terms in a decomposed run (tests/dd_mol.py): the reference's evolved butane cell over `world` slabs against the
    reference's golden vectors and a single-GPU run on rank 0.  Collective."""
    import dd_mol
    g = dd_mol.golden()
    gsys = capi.make_sys(list(g["L"]), float(g["cf"]), float(g["dt"]), skin=0.25)
    n = len(g["x0"])
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.dd_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    s = capi.System(int(1.6 * n / world) + int(3.0 * n / gsys.nsubbox[2]) + 1024, device=local)
    res = dd_mol.rank_run(s, g, gsys, rank, world, bytes(idt.cpu().numpy().tobytes()), nsteps)
    dist.barrier()
    s.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = True
    if rank == 0:
        rec, x = dd_mol.single_run(g, gsys, nsteps, device=local)
        ok = dd_mol.check(g, gathered, rec, x, log=lambda m: print(m, file=sys.stderr, flush=True))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def run_water(rank, world, local, nsteps):
    """Typed Lennard-Jones on global-index rows, bonds, cos^2 angles and the shifted-force Coulomb sum in a decomposed run
    (tests/dd_mol.py): the reference's water box tiled 2^3 over `world` slabs against a single-GPU run on rank 0.  Collective."""
    import dd_mol
    w = dd_mol.water_system()
    gsys = capi.make_sys(list(w["L"]), dd_mol.WATER["cf"], dd_mol.WATER["dt"], skin=0.25)
    n = len(w["x"])
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.dd_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    s = capi.System(int(1.6 * n / world) + int(3.0 * n / gsys.nsubbox[2]) + 1024, device=local)
    res = dd_mol.water_rank_run(s, w, gsys, rank, world, bytes(idt.cpu().numpy().tobytes()), nsteps)
    dist.barrier()
    s.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = True
    if rank == 0:
        ok = dd_mol.water_check(w, gathered, dd_mol.water_single_run(w, gsys, nsteps, device=local),
                                log=lambda m: print(m, file=sys.stderr, flush=True))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def run(rank, world, local, nsteps, ncell, verbose=True):
    """The check itself, for a process group that is already initialised (bench.py --gpus N runs it before its timed
    region, so that the driver's scaling record carries a decomposed-vs-single-GPU comparison).  Collective."""
    NSTEPS, NCELL = nsteps, ncell
    rho, rc, skin, dt, temp, tau = 0.8, 2.5, 0.25, 0.005, 1.0, 0.1
    x, L = cm.lattice(NCELL, rho, jitter=0.08, seed=5)
    n = len(x)
    v = cm.velocities(n, temp, seed=6)
    gsys = capi.make_sys([L] * 3, rc, dt, skin=skin)
    nz = gsys.nsubbox[2]
    z0, z1 = capi.dd_slab_range(rank, world, nz)
    cz = np.floor(x[:, 2] / gsys.lsubbox[2]).astype(np.int64)
    mine = np.nonzero((cz >= z0) & (cz < z1))[0].astype(np.int32)

    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.dd_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    id_bytes = bytes(idt.cpu().numpy().tobytes())

    ncap = int(1.6 * n / world) + int(3.0 * n / nz) + 1024
    s = capi.System(ncap, device=local)
    s.dd_init(rank, world, id_bytes, gsys, n)
    assert s.dd_layers()[:2] == (z0, z1)
    s.dd_set_owned(len(mine))
    s.put(capi.F_X, x[mine]); s.put(capi.F_V, v[mine]); s.put(capi.F_GID, mine)
    s.call("sepgpu_set_alpha", 0, 0.1)
    p = capi.lj_param(rc, kind="lj_shift")

    ref = None
    if rank == 0:
        ref = capi.System(n, device=local)
        ref.put(capi.F_X, x); ref.put(capi.F_V, v)
        ref.call("sepgpu_set_alpha", 0, 0.1)

    ok = True
    pairs_checked = 0
    for step in range(NSTEPS):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(gsys), b"AA", C.byref(p), 1, 1)
        sc = s.scalars()                                   # collective: sums force scalars over ranks
        rebuilt = sc.nbuild
        if step in (0, 1) or (step % 17 == 0):
            # union of the ranks' pair sets, in global ids
            _, _, n_own, n_halo = s.dd_layers()
            loc = s.pairs(max_pairs=int(sc.npairs_listed) + 16)
            cnt = torch.tensor([len(loc)], device="cuda")
            allc = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(allc, cnt)
            mx = int(max(int(c.item()) for c in allc))
            buf = torch.full((mx, 2), -1, dtype=torch.int32, device="cuda")
            if len(loc):
                buf[:len(loc)] = torch.from_numpy(np.ascontiguousarray(loc)).cuda()
            allb = [torch.zeros_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            if rank == 0:
                got = np.concatenate([b.cpu().numpy()[:int(c.item())] for b, c in zip(allb, allc)])
        s.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, tau)
        s.call("sepgpu_leapfrog", C.byref(gsys))
        sc2 = s.scalars()
        if rank == 0:
            ref.call("sepgpu_reset_ret"); ref.call("sepgpu_reset_force")
            ref.call("sepgpu_force_lj", C.byref(gsys), b"AA", C.byref(p), 1, 1)
            if step in (0, 1) or (step % 17 == 0):
                want = cm.pair_set(ref.pairs())
                same = np.array_equal(cm.pair_set(got), want)
                pairs_checked += 1
                if not same:
                    print(f"step {step}: pair sets differ: dd {len(got)} vs single {len(want)}", flush=True)
                    ok = False
            rs = ref.scalars()
            ref.call("sepgpu_nosehoover", C.byref(gsys), temp, 0, tau)
            ref.call("sepgpu_leapfrog", C.byref(gsys))
            rs2 = ref.scalars()
            tol = 1e-9 * (step + 1)
            for name, a, b in (("epot", sc.epot, rs.epot), ("ekin", sc2.ekin, rs2.ekin), ("alpha", sc2.alpha[0], rs2.alpha[0]),
                               ("maxd2", sc2.max_dist2, rs2.max_dist2), ("virial", sc.pot_P[0], rs.pot_P[0])):
                if abs(a - b) > tol * max(abs(b), 1e-3):
                    print(f"step {step}: {name} differs: dd {a!r} vs single {b!r}", flush=True)
                    ok = False
            if sc2.neighb_flag != rs2.neighb_flag:
                print(f"step {step}: trigger differs", flush=True)
                ok = False
    # final state by global id
    _, _, n_own, n_halo = s.dd_layers()
    xo = s.get(capi.F_X); go = s.get(capi.F_GID)
    tot = torch.tensor([n_own], device="cuda")
    dist.all_reduce(tot)
    if int(tot.item()) != n:
        print(f"atom count not conserved: {int(tot.item())} vs {n}", flush=True)
        ok = False
    full = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    full[torch.from_numpy(go.astype(np.int64)).cuda()] = torch.from_numpy(xo).cuda()
    dist.all_reduce(full)
    res = {"world": world, "natoms": n, "steps": NSTEPS, "cell_layers": int(nz)}
    if rank == 0:
        xr = ref.get(capi.F_X)
        err = np.abs(full.cpu().numpy() - xr).max()
        if verbose:
            print(f"dd_check: world={world} n={n} steps={NSTEPS} layers={nz} builds={rs2.nbuild} "
                  f"pair-set checks={pairs_checked} max|dx|={err:.2e} own/halo(rank0)={n_own}/{n_halo} -> {'OK' if ok and err <= 1e-7 else 'FAIL'}",
                  file=sys.stderr, flush=True)
        ok = ok and err <= 1e-7
        res.update({"list_rebuilds": int(rs2.nbuild), "pair_set_comparisons": pairs_checked, "max_abs_dx_vs_single_gpu": float(err),
                    "epot_per_atom_decomposed": sc.epot / n, "epot_per_atom_single_gpu": rs.epot / n})
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    s.close()
    if ref is not None:
        ref.close()
    res["ok"] = int(flag.item()) == 1
    return res


if __name__ == "__main__":
    sys.exit(main())
