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


def test_two_rank_butane_matches_the_reference_golden():
    """Bonded terms in a decomposed run: the reference's evolved 4000-atom butane cell on two slabs -- forces after each of
    prg2's four force routines against the reference's recorded vectors (1e-10), the step's positions / velocities /
    thermostat, then 60 steps (several rebuilds with migration) against a single-GPU run (tests/dd_mol.py)."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, DD_MOL="butane", DD_STEPS="60")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(cm.ROOT, "tests", "dd_check.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "-> OK" in (r.stdout + r.stderr), (r.stdout[-2000:], r.stderr[-2000:])


def _build_prog(tmp_path, name):
    """our own test program (tests/progs/), compiled against include/sep.h and libsep.so"""
    exe = os.path.join(str(tmp_path), name)
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-I" + os.path.join(cm.ROOT, "include"), os.path.join(cm.ROOT, "tests", "progs", name + ".c"),
                           "-L" + os.path.join(cm.ROOT, "seplib_b200"), "-lsep", "-lm", "-o", exe])
    return exe


def _run(exe, args, cwd, ngpu):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(cm.ROOT, "seplib_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    env.pop("SEP_NGPU", None)
    if ngpu > 1:
        env["SEP_NGPU"] = str(ngpu)
    os.makedirs(cwd, exist_ok=True)
    r = subprocess.run([exe, *args], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    import numpy as np
    rows = [ln.split() for ln in r.stdout.splitlines() if ln and ln[0].isdigit()]
    return np.array([[float(v) for v in row] for row in rows])


def test_sep_ngpu_runs_an_unchanged_program_on_two_gpus(tmp_path):
    """SEP_NGPU=2: the sep_* API forks one copy of the program per GPU at the first hot call (seplib_b200/csrc/host/sep_dd.c).
    The same binary, one GPU against two: printed energies / thermostat / pressure agree to rounding growth, the rebuild
    counts are equal, a host-side edit of atoms[] between hot calls takes effect on both, the final configuration
    written by sep_save_xyz agrees, and only ONE copy wrote the files."""
    import numpy as np
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    exe = _build_prog(tmp_path, "nvt_dd")
    one = _run(exe, ["24", "200"], os.path.join(str(tmp_path), "one"), 1)
    two = _run(exe, ["24", "200"], os.path.join(str(tmp_path), "two"), 2)
    assert one.shape == two.shape == (10, 7)
    assert np.abs(one[0, 1:5] - two[0, 1:5]).max() <= 1e-9
    for k in range(10):
        tol = 1e-8 * (k + 1) ** 2
        assert np.abs(one[k, 1:4] - two[k, 1:4]).max() <= tol, (k, one[k], two[k])
        assert abs(one[k, 4] - two[k, 4]) <= 1e-6
    assert np.array_equal(one[:, 6], two[:, 6])                     # list rebuilds at the same steps
    assert np.abs(two[:, 5]).max() < 1e-9                            # total momentum of the gathered array
    x1 = np.loadtxt(os.path.join(str(tmp_path), "one", "final.xyz"), skiprows=2, usecols=(1, 2, 3))
    x2 = np.loadtxt(os.path.join(str(tmp_path), "two", "final.xyz"), skiprows=2, usecols=(1, 2, 3))
    assert x1.shape == x2.shape == (24 ** 3, 3) and np.abs(x1 - x2).max() <= 1e-6
    assert open(os.path.join(str(tmp_path), "two", "steps.log")).read().split() == [str(20 * k) for k in range(10)]


def test_sep_ngpu_prg4_reference_program(tmp_path):
    """the reference's prg4.c (10 000 atoms, NVE, skin 1.0), unchanged, with SEP_NGPU=2 against its golden output"""
    import numpy as np
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(cm.ROOT, "oracle", "_ref", "prgs", "prg4")
    if not os.path.exists(exe):
        pytest.skip("prg4 not built")
    got = _run(exe, ["1"], str(tmp_path), 2)
    ref = np.array([[float(v) for v in ln.split()] for ln in open(os.path.join(cm.GOLDEN, "prg4.ref.out")) if ln and ln[0].isdigit()])
    assert got.shape == ref.shape
    assert np.allclose(got[0, :4], ref[0, :4], rtol=0, atol=2e-9)
    assert np.abs(got[:, 3] - ref[:, 3]).max() < 2e-6                # etot/N over 1000 steps
    assert np.allclose(got[1:, 5], ref[1:, 5], rtol=0.08)            # steps per list rebuild
