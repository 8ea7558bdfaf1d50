"""Decomposed sep_coulomb_sf on real GPUs (two ranks under torchrun; skipped on a single-GPU box).  This file sorts after the
other GPU tests on purpose: the path was finished after the round's GPU budget was spent -- it has passed on the CPU kernel
emulator only (tests/test_cpu_emu.py, two rank threads over the in-process peer-memory stand-in), never on two B200s."""
import os
import subprocess
import sys

import pytest

import common as cm

pytestmark = pytest.mark.gpu


def _gpus():
    from seplib_b200 import capi
    return capi.load().sepgpu_device_count()


def test_two_rank_water_matches_the_single_gpu_run():
    """sep_coulomb_sf, typed Lennard-Jones on global-index rows, bonds and cos^2 angles in a decomposed run: the reference's
    water box tiled 2^3 on two slabs against the same calls on one GPU (forces of the first step 1e-10, 40 steps of sums,
    final positions; tests/dd_mol.py).  Charges travel by global id (sepgpu_dd_set_charges)."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, DD_MOL="water", DD_STEPS="40")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29545", os.path.join(cm.ROOT, "tests", "dd_check.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "-> OK" in (r.stdout + r.stderr), (r.stdout[-2000:], r.stderr[-2000:])
