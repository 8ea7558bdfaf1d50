"""Slab domain decomposition on real GPUs: runs tests/dd_check.py under torchrun when the box has at least two
GPUs (skipped on a single-GPU box; the N>1 host logic is covered on the CPU by tests/test_cpu_dd_plan.py).
dd_check compares the union of the ranks' neighbour pair sets with the single-domain oracle (bit-exact, global
ids), and positions after 60 decomposed steps with the oracle trajectory.  An uneven layer split (DD_NCELL=30:
11 layers over 2 ranks) exercises the per-rank buffer layout of the peer-memory path."""
import os
import subprocess
import sys

import pytest

import common as cm

pytestmark = pytest.mark.gpu


def _gpus():
    from seplib_b200 import capi
    return capi.load().sepgpu_device_count()          # no torch import on the single-GPU path


@pytest.mark.parametrize("ncell", [28, 30])
def test_two_rank_decomposition_matches_oracle(ncell):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, DD_NCELL=str(ncell))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(cm.ROOT, "tests", "dd_check.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "-> OK" in (r.stdout + r.stderr), (r.stdout[-2000:], r.stderr[-2000:])
