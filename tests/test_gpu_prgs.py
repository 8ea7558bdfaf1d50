"""Acceptance: the reference's own example programs, compiled UNCHANGED against include/sep.h and linked
with libsep.so (scripts/build_prgs.sh -> oracle/_ref/prgs/), run on the GPU and print what the
reference prints.  Golden stdout in tests/golden/prg*.ref.out comes from the same programs linked
against the compiled reference.

Trajectories are chaotic, so later lines are compared on what the programs are meant to demonstrate
(reference prgs/prg0.c:4-8: energy and momentum conservation; prg1/prg4: thermostat, neighbour list):
step-0 line to printed precision, conserved energy to 5e-5 per atom, temperature/pressure/rebuild
frequency within a few percent."""
import os
import subprocess

import numpy as np
import pytest

import common as cm

pytestmark = pytest.mark.gpu
PRGS = os.path.join(cm.ROOT, "oracle", "_ref", "prgs")


def run_prg(name, args=(), tmp_path=None, timeout=600):
    exe = os.path.join(PRGS, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (scripts/build_prgs.sh needs the reference sources)")
    env = dict(os.environ)
    libdir = os.path.join(cm.ROOT, "seplib_b200")
    if env.get("SEPGPU_EMU_LIB"):                 # test run on the CPU kernel emulator (tests/emu/run_on_emu.py):
        libdir = os.path.join(os.path.dirname(env["SEPGPU_EMU_LIB"]), "emu_lib")      # holds libsep.so -> libsep_emu.so
    env["LD_LIBRARY_PATH"] = libdir + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe, *args], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = [ln.split() for ln in out.stdout.splitlines() if ln and ln[0].isdigit()]
    return np.array([[float(v) for v in r] for r in rows]), out.stdout


def golden(name):
    rows = [ln.split() for ln in open(os.path.join(cm.GOLDEN, name)) if ln and ln[0].isdigit()]
    return np.array([[float(v) for v in r] for r in rows])


def test_prg0_brute_lj_nve(tmp_path):
    """columns: n t epot/N ekin/N T etot/N sum_p p f0x   (reference prgs/prg0.c:62-66)"""
    got, _ = run_prg("prg0", tmp_path=tmp_path)
    ref = golden("prg0.ref.out")
    assert got.shape == ref.shape
    assert np.allclose(got[0, :8], ref[0, :8], rtol=0, atol=2e-6)            # identical start, printed precision
    assert np.abs(got[:2, 5] - ref[:2, 5]).max() < 1e-8                      # 1000 steps: etot/N still equal to 1e-8
    # later lines: the leapfrog energy estimate itself fluctuates by ~5e-4 (the reference's own column does),
    # and the two trajectories have decorrelated; both must stay on the same energy shell
    assert np.abs(got[:, 5] - ref[:, 5]).max() < 2e-3
    assert np.abs(got[:, 5] - got[0, 5]).max() < 2e-3                        # and conserved in itself
    assert np.abs(got[:, 6]).max() < 1e-12                                   # momentum
    assert abs(got[1:, 4].mean() - ref[1:, 4].mean()) < 0.03                 # temperature level


def test_prg4_list_lj_skin1(tmp_path):
    """columns: n epot/N ekin/N etot/N sum_p n/nupdate   (reference prgs/prg4.c:65-68); skin set to 1.0 after setup"""
    got, _ = run_prg("prg4", args=("1",), tmp_path=tmp_path)
    ref = golden("prg4.ref.out")
    assert got.shape == ref.shape
    assert np.allclose(got[0, :4], ref[0, :4], rtol=0, atol=2e-9)
    assert np.abs(got[:, 3] - ref[:, 3]).max() < 2e-6                        # etot/N over 1000 steps (N=10000)
    assert np.allclose(got[1:, 5], ref[1:, 5], rtol=0.08)                    # steps per list rebuild


def test_prg1_nvt_list(tmp_path):
    """columns: n t epot/N ekin/N T etot/N sum_p p n/nupdate   (reference prgs/prg1.c:73-78)"""
    got, _ = run_prg("prg1", tmp_path=tmp_path)
    ref = golden("prg1.ref.out")
    assert got.shape == ref.shape
    assert np.allclose(got[0, 2:6], ref[0, 2:6], rtol=0, atol=2e-6)
    assert abs(got[5:, 4].mean() - ref[5:, 4].mean()) < 0.01                 # thermostatted temperature
    # epot/N: 95 printed samples of a 1000-atom system scatter by ~0.03 each (independent at this print interval), so the
    # two means differ by ~0.005 rms once the trajectories have decorrelated; 4 sigma
    assert abs(got[5:, 2].mean() - ref[5:, 2].mean()) < 0.02
    assert abs(got[5:, 7].mean() - ref[5:, 7].mean()) < 0.05                 # pressure
    assert np.allclose(got[5:, 8].mean(), ref[5:, 8].mean(), rtol=0.03)      # steps per list rebuild


def test_prg2_prg3_prg6_run(tmp_path):
    """butane (bond/angle/torsion + EXCL_SAME_MOL list), water (brute LJ + SF Coulomb + box compression) and
    DPD run to completion on the GPU with sane output; their start files are written from the golden states."""
    cm.write_molecular_start_files(tmp_path)
    got2, txt2 = run_prg("prg2", tmp_path=tmp_path)
    # columns: n t epot/N ekin/N etot/N T sum_p p p_mol
    assert len(got2) == 100 and np.isfinite(got2[:, :8]).all()
    # prg2 is deterministic (velocities come from the start file): its printed lines against the same program linked
    # with the compiled REFERENCE (tests/golden/prg2.ref.out).  Step 0 to printed precision incl. the molecular pressure,
    # step 100 to 5e-5 (chaotic growth from rounding), afterwards the averages the program is about.
    ref2 = golden("prg2.ref.out")
    assert got2.shape == ref2.shape
    assert np.allclose(got2[0, 2:6], ref2[0, 2:6], rtol=0, atol=2e-6), (got2[0], ref2[0])
    assert abs(got2[0, 7] - ref2[0, 7]) <= 0.011 and abs(got2[0, 8] - ref2[0, 8]) <= 0.011, (got2[0], ref2[0])
    assert np.allclose(got2[1, 2:5], ref2[1, 2:5], rtol=0, atol=5e-5), (got2[1], ref2[1])
    assert abs(got2[1, 7] - ref2[1, 7]) <= 0.03 and abs(got2[1, 8] - ref2[1, 8]) <= 0.03, (got2[1], ref2[1])
    assert abs(got2[10:, 5].mean() - ref2[10:, 5].mean()) < 0.02            # thermostatted temperature (4.0)
    # epot/N: the 90 printed samples scatter by 0.07 each, so the means of two decorrelated trajectories differ by
    # ~0.01 rms (measured 0.002 and 0.016 with two orderings of the same neighbour rows); 4 sigma
    assert abs(got2[10:, 2].mean() - ref2[10:, 2].mean()) < 0.04
    assert abs(got2[10:, 7].mean() - ref2[10:, 7].mean()) < 0.15            # atomic pressure
    assert abs(got2[10:, 8].mean() - ref2[10:, 8].mean()) < 0.15            # molecular pressure
    got3, txt3 = run_prg("prg3", tmp_path=tmp_path)
    # columns: n t T sum_p rho epot/mol p p_mol
    assert len(got3) == 10 and np.isfinite(got3[:, :7]).all()
    assert got3[-1, 4] > got3[0, 4]                                           # the box is being compressed
    ref3 = golden("prg3.ref.out")                                             # (velocities are time-seeded: only the
    assert np.allclose(got3[:, 4], ref3[:, 4], rtol=0, atol=1.1e-3)           #  density schedule is deterministic)
    got6, txt6 = run_prg("prg6", tmp_path=tmp_path)
    # columns: n t epot/N ekin/N T etot/N sum_p p
    assert np.isfinite(got6).all()
    assert abs(got6[10:, 4].mean() - 1.0) < 0.05                             # DPD thermostat holds T = 1.0
    assert np.abs(got6[:, 6]).max() < 1e-10                                  # momentum conserved by the pair noise
