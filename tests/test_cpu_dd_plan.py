"""CPU checks of the multi-GPU host logic (no device): the slab plan that bench.py / dd_check.py use to hand
each rank its atoms, exercised with a real 2-process `gloo` group -- every atom is owned by exactly one
rank, ownership follows the reference's cell expression (int)(z/lsubbox), the id broadcast plumbing
works, and the weak-scaling lattice keeps the per-GPU atom count fixed."""
import importlib.util
import os
import socket

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(cm.ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_weak_lattice_and_slab_partition_cover_every_atom_once():
    b = _bench()
    assert b.weak_lattice_dims(100, 1) == [100, 100, 100]
    assert b.weak_lattice_dims(100, 8) == [200, 200, 200]
    for world in (2, 4, 8):
        side = 12 if world < 8 else 24
        dims = b.weak_lattice_dims(side, world)
        assert dims[0] * dims[1] * dims[2] == world * side ** 3
        Lvec, a = b.lattice_box(dims, 0.8)
        gs = capi.make_sys(Lvec, 2.5, 0.005)
        nz = gs.nsubbox[2]
        assert nz >= 2 * world
        seen = np.zeros(dims[0] * dims[1] * dims[2], dtype=np.int32)
        for r in range(world):
            z0, z1 = capi.dd_slab_range(r, world, nz)
            pos, gid = b.slab_atoms(dims, a, z0, z1, gs.lsubbox[2])
            seen[gid] += 1
            cz = np.floor(pos[:, 2] / gs.lsubbox[2]).astype(int)
            assert ((cz >= z0) & (cz < z1)).all()
            assert (pos >= 0).all() and (pos[:, 0] < Lvec[0]).all() and (pos[:, 2] < Lvec[2]).all()
            v = b.slab_velocities(gid, len(seen), 1.0, 7)
            assert v.shape == pos.shape
        assert (seen == 1).all()
        ranges = [capi.dd_slab_range(r, world, nz) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == nz and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _bench()
    dims = b.weak_lattice_dims(10, world)
    Lvec, a = b.lattice_box(dims, 0.8)
    gs = capi.make_sys(Lvec, 2.5, 0.005)
    z0, z1 = capi.dd_slab_range(rank, world, gs.nsubbox[2])
    pos, gid = b.slab_atoms(dims, a, z0, z1, gs.lsubbox[2])
    # the 128-byte id travels exactly like the NCCL unique id does in bench.py (rank 0 -> everyone)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(idt, 0)
    n_own = torch.tensor([len(pos)], dtype=torch.int64)
    dist.all_reduce(n_own)
    zsum = torch.tensor([pos[:, 2].sum()], dtype=torch.float64)
    dist.all_reduce(zsum)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    out[rank] = (int(n_own.item()), float(zsum.item()), bytes(idt.numpy().tobytes()) == bytes(range(128)), float(tmax.item()),
                 dims[0] * dims[1] * dims[2])
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_process_gloo_plan_is_consistent():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    n_tot, zsum, id_ok, tmax, n_expected = out[0]
    assert out[1][:2] == (n_tot, zsum)
    assert n_tot == n_expected and id_ok and out[1][2] and tmax == 2.0
