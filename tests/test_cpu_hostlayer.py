"""The HOST layer (seplib_b200/csrc/host/sep_*.c) on the CPU.  The unchanged host sources are linked with a stand-in
device layer that keeps its state in host memory and computes with the oracle (tests/mock_device/sepgpu_mock.c --
test infrastructure, never part of libsep.so), and the same sep_* loops the GPU tests run are driven through it:
dispatch and control flow (brute / list, rebuild flag, epot assign vs accumulate over several typed calls), the
bookkeeping of the box-changing routines, per-type relaxation and tethers, the Gaussian stream handed to the
stochastic integrators, and the step / lazy coherence modes -- all against the reference's recorded loops."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi
from test_gpu_zz_next import _check

BUILD = os.path.join(cm.ROOT, "tests", "_build")
G = np.load(os.path.join(cm.GOLDEN, "next_rows.npz"))


@pytest.fixture(scope="module")
def mock():
    so = os.path.join(BUILD, "libsep_hostmock.so")
    srcs = sorted(glob.glob(os.path.join(cm.ROOT, "seplib_b200", "csrc", "host", "*.c"))) + [
        os.path.join(cm.ROOT, "tests", "mock_device", "sepgpu_mock.c"), os.path.join(cm.ROOT, "oracle", "sep_oracle.c")]
    deps = srcs + glob.glob(os.path.join(cm.ROOT, "include", "*.h")) + glob.glob(os.path.join(cm.ROOT, "seplib_b200", "csrc", "host", "*.h"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-ffp-contract=off",
                               "-I" + os.path.join(cm.ROOT, "include"), "-I" + os.path.join(cm.ROOT, "oracle"),
                               "-I" + os.path.join(cm.ROOT, "seplib_b200", "csrc", "host"), *srcs, "-Wl,--no-undefined", "-lm", "-o", so])
    lib = C.CDLL(so, mode=C.RTLD_LOCAL)
    capi.declare_sep_api(lib)
    return lib


@pytest.mark.parametrize("sync", [1, 0, 2, 3])          # step, lazy, full, auto
def test_lj_loop_matches_reference_golden(mock, sync):
    """the prg1 loop of tests/test_gpu_more.py::test_sep_api_lj_loop_matches_reference_golden, on the mock"""
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    mock.sep_gpu_set_sync(sync)
    s = cm.ApiSystem(mock, g["x0"], float(g["L"]), float(g["cf"]), float(g["dt"]), v=g["v0"], nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    alpha = C.c_double(float(g["alpha0"]))
    fun = s.fun("sep_lj_shift")
    mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
    mock.sep_force_pairs(s.atoms, b"AA", float(g["cf"]), fun, s.S, s.R, 1)
    mock.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], g["f_pairs"]) <= 1e-11
    if sync != 0:
        assert abs(s.ret.epot - float(g["epot"])) <= 1e-11 * abs(float(g["epot"]))
    mock.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), float(g["tau"]), s.S)
    mock.sep_leapfrog(s.atoms, s.S, s.R)
    mock.sep_gpu_sync(s.atoms)
    assert np.abs(s.view["x"] - g["x1"]).max() <= 1e-12 and np.abs(s.view["v"] - g["v1"]).max() <= 1e-12
    assert abs(alpha.value - float(g["alpha1"])) <= 1e-12
    assert abs(s.ret.ekin - float(g["ekin"])) <= 1e-11 * float(g["ekin"])
    assert np.array_equal(s.view["crossings"], g["cr1"]) and np.array_equal(s.view["cross_neighb"], g["cn1"])
    assert int(s.sys.neighb_flag) == int(g["neighb_flag1"]) and abs(s.sys.tnow - float(g["dt"])) <= 1e-15
    # 40 more steps: the reference's aggregate trajectory and its list-rebuild count
    nup0 = int(s.sys.nupdate_neighb)
    traj = g["traj"]
    for k in range(40):
        mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
        mock.sep_force_pairs(s.atoms, b"AA", float(g["cf"]), fun, s.S, s.R, 1)
        mock.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), float(g["tau"]), s.S)
        mock.sep_leapfrog(s.atoms, s.S, s.R)
        mock.sep_pressure_tensor(s.R, s.S)
        assert abs(s.ret.epot - traj[k, 0]) <= 1e-9 * abs(traj[k, 0]) and abs(s.ret.ekin - traj[k, 1]) <= 1e-9 * traj[k, 1]
        assert abs(s.ret.p - traj[k, 2]) <= 1e-8 * abs(traj[k, 2]) and abs(alpha.value - traj[k, 3]) <= 1e-9 * abs(traj[k, 3])
        if k == 0:
            nup0 = int(s.sys.nupdate_neighb)
    assert int(s.sys.nupdate_neighb) - nup0 == int(traj[-1, 4] - traj[0, 4])      # rebuilds over the last 39 steps
    s.close()
    mock.sep_gpu_set_sync(1)


def test_auto_mode_notices_host_reads_and_writes(mock):
    """SEP_SYNC=auto (the default): atoms[] is page-protected while the device copy is newer.  Reading a member right after
    a hot call -- no sep_gpu_sync -- sees the fresh state; writing members between calls -- no sep_gpu_invalidate -- is
    picked up by the next hot call (reference callers do both: prgs/prg0.c:64 prints atoms[0].f, prgs/prg5.c writes f)."""
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    cf, dt = float(g["cf"]), float(g["dt"])

    def system(mode):
        mock.sep_gpu_set_sync(mode)
        s = cm.ApiSystem(mock, g["x0"], float(g["L"]), cf, dt, v=g["v0"], nneighb=0)
        s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
        return s, C.c_double(float(g["alpha0"])), s.fun("sep_lj_shift")

    def step(s, alpha, fun):
        mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
        mock.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
        mock.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), float(g["tau"]), s.S)
        mock.sep_leapfrog(s.atoms, s.S, s.R)

    s, alpha, fun = system(3)
    mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
    mock.sep_force_pairs(s.atoms, b"AA", cf, fun, s.S, s.R, 1)
    assert cm.rel_force_err(s.view["f"].copy(), g["f_pairs"]) <= 1e-11          # read after a force call: no sync call
    mock.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), float(g["tau"]), s.S)
    mock.sep_leapfrog(s.atoms, s.S, s.R)
    assert np.abs(s.view["x"].copy() - g["x1"]).max() <= 1e-12 and np.abs(s.view["v"].copy() - g["v1"]).max() <= 1e-12
    for _ in range(3):                                                          # untouched steps in between
        step(s, alpha, fun)
    s.view["v"][:] = s.view["v"] * 0.5                                          # user code writes velocities: no invalidate
    s.view["f"][:, 0] += 0.0
    for _ in range(3):
        step(s, alpha, fun)
    xa, va, ea = s.view["x"].copy(), s.view["v"].copy(), s.ret.ekin
    s.close()
    # the same protocol in step mode with the explicit calls
    t, alpha, fun = system(1)
    for _ in range(4):
        step(t, alpha, fun)
    t.view["v"][:] = t.view["v"] * 0.5
    mock.sep_gpu_invalidate(t.atoms)
    for _ in range(3):
        step(t, alpha, fun)
    assert np.array_equal(xa, t.view["x"]) and np.array_equal(va, t.view["v"]) and ea == t.ret.ekin
    t.close()
    mock.sep_gpu_set_sync(1)


@pytest.mark.parametrize("sync", [1, 0])
def test_box_changing_callers(mock, sync):
    mock.sep_gpu_set_sync(sync)
    _check(cm.drive_compress(mock, G["c_x0"], G["c_v0"], float(G["c_L"])), "compress", scalar_cols=(0, 1, 2, 4), exact_cols=(3,), xtol=1e-9)
    _check(cm.drive_berendsen(mock, G["b_x0"], G["b_v0"], float(G["b_L"])), "ber", scalar_cols=(0, 1, 2, 3, 4), exact_cols=(5,), xtol=1e-9)
    _check(cm.drive_berendsen(mock, G["c_x0"], G["c_v0"], float(G["c_L"]), steps=12, iso=True, update=capi.SEP_LLIST_NEIGHBLIST),
           "beriso", scalar_cols=(0, 1, 2, 3, 4), exact_cols=(5,), xtol=1e-9)
    mock.sep_gpu_set_sync(1)


@pytest.mark.parametrize("sync", [1, 0])
def test_slit_pore_three_typed_calls_relax_temp_tethers(mock, sync):
    mock.sep_gpu_set_sync(sync)
    rec = cm.drive_slit(mock, G["c_x0"], G["c_v0"], float(G["c_L"]))
    mock.sep_gpu_set_sync(1)
    assert np.array_equal(rec["types"], G["slit_types"])
    _check(rec, "slit", scalar_cols=(0, 1), xtol=1e-9)


def test_stochastic_integrators_get_the_reference_noise(mock):
    for which in ("fp", "gjf"):
        rec = cm.drive_stochastic(mock, G["b_x0"], G["b_v0"], float(G["b_L"]), which)
        _check(rec, which, scalar_cols=(0, 1), xtol=1e-9)
        assert np.abs(rec["traj"][:, 2] - G[which + "_traj"][:, 2]).max() <= 1e-12


def _top_file(g, tmp_path, name):
    top = str(tmp_path / name)
    with open(top, "w") as fh:
        fh.write("[ bonds ]\n;generated\n")
        for (a, b, t) in g["blist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {t}\n")
        fh.write("\n[ angles ]\n;generated\n")
        for (a, b, c, t) in g["alist"]:
            fh.write(f"{g['molindex'][a]} {a} {b} {c} {t}\n")
        if len(g["dlist"]):
            fh.write("\n[ dihedrals ]\n;generated\n")
            for (a, b, c, d, t) in g["dlist"]:
                fh.write(f"{g['molindex'][a]} {a} {b} {c} {d} {t}\n")
    return top


@pytest.mark.parametrize("sync", [1, 0])
def test_butane_and_water_steps(mock, sync, tmp_path):
    """prg2 (list mode, same-molecule exclusion, bond / angle / Ryckaert torsion) and prg3 (brute LJ, cos^2 angles,
    shifted-force Coulomb) force sequences, then thermostat + leapfrog, against the reference's recorded step"""
    mock.sep_gpu_set_sync(sync)
    g = np.load(os.path.join(cm.GOLDEN, "butane_n4000.npz"))
    n = len(g["x0"])
    s = cm.ApiSystem(mock, g["x0"], g["L"], 2.5, float(g["dt"]), v=g["v0"], types=np.full(n, ord("C"), dtype=np.uint8), nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    mock.sep_read_topology_file(s.atoms, _top_file(g, tmp_path, "b.top").encode(), s.S, b"q")
    rb = (C.c_double * 6)(*g["rb"])
    alpha = C.c_double(float(g["alpha0"]))
    mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
    mock.sep_force_pairs(s.atoms, b"CC", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    mock.sep_stretch_harmonic(s.atoms, 0, 0.407, 2074.0, s.S, s.R)
    mock.sep_angle_harmonic(s.atoms, 0, 1.90, 400.0, s.S, s.R)
    mock.sep_torsion_Ryckaert(s.atoms, 0, rb, s.S, s.R)
    mock.sep_gpu_sync_scalars(s.atoms, s.S, s.R)
    assert abs(s.ret.epot - float(g["epot_torsion"])) <= 1e-11 * abs(float(g["epot_torsion"]))
    mock.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], g["f_torsion"]) <= 1e-11
    mock.sep_nosehoover(s.atoms, float(g["temp"]), C.byref(alpha), 0.1, s.S)
    mock.sep_leapfrog(s.atoms, s.S, s.R)
    mock.sep_gpu_sync(s.atoms)
    assert np.abs(s.view["x"] - g["x1"]).max() <= 1e-12 and np.abs(s.view["v"] - g["v1"]).max() <= 1e-11
    assert abs(s.ret.ekin - float(g["ekin"])) <= 1e-11 * float(g["ekin"])
    s.close()

    g = np.load(os.path.join(cm.GOLDEN, "water_n648.npz"))
    s = cm.ApiSystem(mock, g["x0"], g["L"], float(g["cf"]), float(g["dt"]), update=capi.SEP_BRUTE, v=g["v0"], types=g["type"],
                     m=g["m"], z=g["z"], nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    mock.sep_read_topology_file(s.atoms, _top_file(g, tmp_path, "w.top").encode(), s.S, b"q")
    alpha3 = (C.c_double * 3)(float(g["alpha0"]), 0.0, 0.0)
    mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
    mock.sep_force_pairs(s.atoms, b"OO", 2.5, s.fun("sep_lj_shift"), s.S, s.R, 3)
    mock.sep_stretch_harmonic(s.atoms, 0, 0.316, 68421.0, s.S, s.R)
    mock.sep_angle_cossq(s.atoms, 0, 1.97, 490.0, s.S, s.R)
    mock.sep_coulomb_sf(s.atoms, float(g["cf"]), s.S, s.R, 3)
    mock.sep_gpu_sync_scalars(s.atoms, s.S, s.R)
    mock.sep_gpu_sync(s.atoms)
    assert cm.rel_force_err(s.view["f"], g["f_coul"]) <= 1e-11
    assert abs(s.ret.epot - float(g["epot_coul"])) <= 1e-11 * abs(float(g["epot_coul"]))
    assert abs(s.ret.ecoul - float(g["ecoul"])) <= 1e-11 * abs(float(g["ecoul"]))
    mock.sep_nosehoover(s.atoms, float(g["temp"]), alpha3, 0.01, s.S)
    mock.sep_leapfrog(s.atoms, s.S, s.R)
    mock.sep_gpu_sync(s.atoms)
    assert np.abs(s.view["x"] - g["x1"]).max() <= 1e-12 and np.abs(s.view["v"] - g["v1"]).max() <= 1e-10
    s.close()
    mock.sep_gpu_set_sync(1)


@pytest.mark.parametrize("prg,args", [("prg0", ()), ("prg9", ()), ("prg1", ()), ("prg4", ("1",))])
def test_reference_program_on_the_mock_reproduces_the_reference_output(mock, prg, args, tmp_path):
    """The reference's example program, unchanged, compiled against include/sep.h and linked with the host layer +
    mock device, prints exactly what it prints when linked with the reference (tests/golden/<prg>.ref.out): 10 000
    steps of prg0 (brute LJ, NVE) and of prg9 (sep_set_vel_seed's rand() stream continuing into sep_fp's noise),
    prg1 (list mode, Nose-Hoover, samplers attached) and prg4 (skin reset to 1.0 after setup, the reference's quirk).
    Needs the reference sources (build container only)."""
    src = f"/root/reference/prgs/{prg}.c"
    if not os.path.exists(src):
        pytest.skip("reference sources not present")
    exe = str(tmp_path / prg)
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-w", "-I" + os.path.join(cm.ROOT, "include"), src, "-L" + BUILD,
                           "-lsep_hostmock", "-lm", "-Wl,-rpath," + BUILD, "-o", exe])
    out = subprocess.run([exe, *args], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0
    want = open(os.path.join(cm.GOLDEN, f"{prg}.ref.out")).read()
    assert out.stdout == want


def test_dpd_loop_is_independent_of_the_coherence_mode(mock):
    """prg6's loop (sep_force_dpd + sep_verlet_dpd, predictor state pv/pa mirrored on the device): the step, lazy and
    full modes must give the same trajectory -- what differs between them is only when atoms[] is refreshed."""
    x, L = cm.lattice(6, 3.0, jitter=0.3, seed=51)
    v = cm.velocities(len(x), 1.0, seed=52)
    results = []
    for sync in (1, 0, 2):
        mock.sep_gpu_set_sync(sync)
        s = cm.ApiSystem(mock, x, L, 1.0, 0.02, v=v, nneighb=0)
        for n in range(25):
            mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
            mock.sep_force_dpd(s.atoms, b"AA", 1.0, 25.0, 1.0, 3.0, s.S, s.R, 1)
            mock.sep_verlet_dpd(s.atoms, 0.5, n, s.S, s.R)
        mock.sep_gpu_sync(s.atoms)
        results.append((s.view["x"].copy(), s.view["v"].copy(), s.view["pv"].copy(), s.ret.ekin, int(s.sys.neighb_flag)))
        s.close()
    mock.sep_gpu_set_sync(1)
    for r in results[1:]:
        assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])
        assert r[3] == results[0][3] and r[4] == results[0][4]
    assert np.isfinite(results[0][0]).all() and abs(results[0][3] / len(x) - 1.5) < 0.5


def test_sep_force_pairs_samples_a_callers_own_pair_function(mock):
    """host layer: a function the library does not know by address is sampled once per (function, cutoff) and handed
    to sepgpu_force_table; the mock evaluates the table over the oracle's pair list (same cubic as the kernels)."""
    g = np.load(os.path.join(cm.GOLDEN, "lj_n1000.npz"))
    mock.sep_gpu_set_sync(1)
    s = cm.ApiSystem(mock, g["x0"], float(g["L"]), float(g["cf"]), float(g["dt"]), v=g["v0"], nneighb=0)
    s.view["xn"][:] = g["xn0"]; s.view["cross_neighb"][:] = g["cn0"]; s.view["crossings"][:] = g["cr0"]
    calls = [0]

    def user_lj_shift(r2, opt):
        calls[0] += 1
        rri = 1.0 / r2; rri3 = rri * rri * rri
        return 48.0 * rri3 * (rri3 - 0.5) * rri if opt == b"f" else 4.0 * rri3 * (rri3 - 1.0) + 0.016316891136

    cb = C.CFUNCTYPE(C.c_double, C.c_double, C.c_char)(user_lj_shift)
    for rep in range(2):
        mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
        mock.sep_force_pairs(s.atoms, b"AA", float(g["cf"]), C.cast(cb, C.c_void_p), s.S, s.R, 1)
        mock.sep_gpu_sync(s.atoms)
        assert cm.rel_force_err(s.view["f"], g["f_pairs"]) <= 1e-9
        assert abs(s.ret.epot - float(g["epot"])) <= 1e-9 * abs(float(g["epot"]))
        assert calls[0] == 2 * 65536                    # sampled once, reused on the second call
    mock.sep_pairs_retabulate()
    mock.sep_reset_retval(s.R); mock.sep_reset_force(s.atoms, s.S)
    mock.sep_force_pairs(s.atoms, b"AA", float(g["cf"]), C.cast(cb, C.c_void_p), s.S, s.R, 1)
    assert calls[0] == 4 * 65536
    s.close()
