"""Device integrator arithmetic exercised on the CPU.  The per-atom update of the stochastic integrators lives in a
__host__ __device__ header (seplib_b200/csrc/gpu/sepgpu_intgr_atom.cuh); tests/host_kernels_test.cu compiles it for
the host, and this test drives it through the reference's recorded loops (tests/golden/next_rows.npz): oracle pair
forces, the reference's Gaussian stream, then the SAME code the GPU kernel runs.  Skipped when nvcc is absent."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import common as cm

BUILD = os.path.join(cm.ROOT, "tests", "_build")
SRC = os.path.join(cm.ROOT, "tests", "host_kernels_test.cu")
HDR = os.path.join(cm.ROOT, "seplib_b200", "csrc", "gpu", "sepgpu_intgr_atom.cuh")


@pytest.fixture(scope="module")
def hostk():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = os.path.join(BUILD, "libhostk.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(BUILD, exist_ok=True)
        subprocess.check_call([nvcc, "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-Wno-deprecated-gpu-targets",
                               "-I" + os.path.join(cm.ROOT, "include"), "-I" + os.path.join(cm.ROOT, "seplib_b200", "csrc", "gpu"),
                               SRC, "-o", so])
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.hostk_stoch_step.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                     C.c_double, C.c_double, C.c_double, vp]
    return lib


@pytest.mark.parametrize("which", ["fp", "gjf"])
def test_stochastic_atom_update_follows_the_reference(hostk, which):
    g = np.load(os.path.join(cm.GOLDEN, "next_rows.npz"))
    cf, dt, temp, alpha = 1.12246204830937, 0.001, 1.12, 1.0
    o = cm.OracleLoop(g["b_x0"], g["b_v0"], float(g["b_L"]), cf, dt, list_mode=False)
    n = o.n
    prevf = np.zeros((n, 3)); randn = np.zeros((n, 3)); clpack = np.zeros(n, dtype=np.int32)
    noise = np.zeros((n, 4)); noise[:, 3] = 1.0                        # ldiff = 1 (sep_init)
    out = np.zeros(2)
    o.orc.orc_randn_reset()
    cm._libc.srand(cm.STOCH_SEED)
    ref = g[which + "_traj"]
    for step in range(len(ref)):
        o.reset(); o.pair_force(b"AA", cf, cm.POT_LJ_SHIFT)
        for i in range(n):                                             # the reference's drawing order: atom-major, x y z
            for k in range(3):
                noise[i, k] = o.orc.orc_randn()
        hostk.hostk_stoch_step(1 if which == "gjf" else 0, n, cm.ptr(o.x), cm.ptr(o.v), cm.ptr(o.f), cm.ptr(o.m), cm.ptr(noise),
                               cm.ptr(prevf), cm.ptr(randn), cm.ptr(o.xn), cm.ptr(o.cn), cm.ptr(o.cr), cm.ptr(clpack),
                               cm.ptr(o.len), dt, temp, alpha if which == "gjf" else 0.0, cm.ptr(out))
        ekin = 0.5 * out[0]
        assert abs(ekin - ref[step, 1]) <= 1e-11 * abs(ref[step, 1]), (which, step)
        assert abs(o.ret.epot - ref[step, 0]) <= 1e-11 * abs(ref[step, 0]), (which, step)
    assert np.abs(o.x - g[which + "_x"]).max() <= 1e-11
    assert np.abs(o.v - g[which + "_v"]).max() <= 1e-11
