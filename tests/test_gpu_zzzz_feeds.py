"""Sampler feeds (include/sepgpu.h sepgpu_feed_*): the sums the reference's run-time samplers form
over atoms[] (source/sepsampler.c), computed on the device.

  * every feed against numpy on the arrays downloaded from the same context (the reference's formulas, restated here
    line by line with the source lines they follow; the pair histogram must agree exactly -- integer counts);
  * through the sep_* API: the same sampled two-species NVT loop with SEP_SAMPLER_FEEDS=1 and =0 writes the same files
    (the =0 host samplers are pinned to the reference's own files in tests/test_cpu_samplers.py), and the feed run
    never downloads atoms[].

This file sorts after the other GPU tests on purpose: the feed kernels were written after the round's GPU budget was
spent -- they have passed on the CPU kernel emulator only (tests/test_cpu_emu.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

pytestmark = pytest.mark.gpu


def _evolved(nsteps=25, ncell=8, dt=0.01):
    """a two-species Lennard-Jones system a few steps off the lattice, hot enough to cross the box boundaries"""
    x, L = cm.lattice(ncell, 0.8, jitter=0.05, seed=51)
    n = len(x)
    v = cm.velocities(n, 2.5, seed=52)
    types = np.full(n, ord("A"), dtype=np.uint8); types[: n // 3] = ord("B")
    m = np.ones(n); m[: n // 3] = 1.7
    s = capi.System(n)
    s.put(capi.F_X, x); s.put(capi.F_V, v); s.put(capi.F_TYPE, types); s.put(capi.F_M, m)
    sys_ = capi.make_sys([L] * 3, 2.5, dt, skin=0.25)
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_set_alpha", 0, 0.05)

    def step(k=1):
        for _ in range(k):
            s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
            s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), cm.ALL, 1)
            s.call("sepgpu_force_lj", C.byref(sys_), b"AB", C.byref(p), cm.ALL, 0)
            s.call("sepgpu_force_lj", C.byref(sys_), b"BB", C.byref(p), cm.ALL, 0)
            s.call("sepgpu_nosehoover", C.byref(sys_), 2.5, 0, 0.1)
            s.call("sepgpu_leapfrog", C.byref(sys_))

    step(nsteps)
    return s, step, L, types, m


def test_msd_feed_follows_the_crossing_counters():
    s, step, L, types, m = _evolved()
    length = np.array([L] * 3)
    k = 2 * np.pi * np.arange(1, 18) / L                       # 17 wave numbers: two groups of the kernel
    sums, fs = np.zeros(3), np.zeros(len(k))
    s.call("sepgpu_feed_msd", 1, b"A", length.ctypes.data, len(k), k.ctypes.data, sums.ctypes.data, fs.ctypes.data)
    x0 = s.get(capi.F_X); c0 = s.get(capi.F_CROSSINGS)
    assert sums[0] == 0.0 and sums[1] == 0.0 and sums[2] == (types == ord("A")).sum() and np.allclose(fs, sums[2])
    step(40)
    s.call("sepgpu_feed_msd", 0, b"A", length.ctypes.data, len(k), k.ctypes.data, sums.ctypes.data, fs.ctypes.data)
    x1 = s.get(capi.F_X); c1 = s.get(capi.F_CROSSINGS)
    assert np.abs(c1 - c0).sum() > 0, "the test wants atoms that crossed the box boundary"
    sel = types == ord("A")
    dr = (x1 + (c1 - c0) * length - x0)[sel]                   # source/sepsampler.c:583-589
    a = (dr ** 2).sum(axis=1)
    assert abs(sums[0] - a.sum()) <= 1e-12 * a.sum()
    assert abs(sums[1] - (a * a).sum()) <= 1e-12 * (a * a).sum()
    assert sums[2] == sel.sum()
    want = np.cos(np.outer(k, dr[:, 0])).sum(axis=1)           # :597, the x displacement only
    assert np.abs(fs - want).max() <= 1e-10 * sel.sum()
    s.close()


def test_vacf_feed_block():
    s, step, L, types, m = _evolved()
    lvec = 12
    rows, blk, done = [], np.zeros(lvec), C.c_int(0)
    for t in range(2 * lvec):                                   # two blocks: the second one starts from a clean buffer
        rows.append(s.get(capi.F_V)[:, 0].copy())
        s.call("sepgpu_feed_vacf", lvec, blk.ctypes.data, C.byref(done))
        assert done.value == (1 if (t + 1) % lvec == 0 else 0)
        if done.value:
            r = np.array(rows[-lvec:])
            want = np.array([(r[: lvec - tt] * r[tt:]).sum() for tt in range(lvec)])        # source/sepsampler.c:700-712
            assert np.abs(blk - want).max() <= 1e-12 * np.abs(want).max()
        step(2)
    s.close()


def test_profile_feed():
    s, step, L, types, m = _evolved()
    nb = 10
    out = np.zeros(4 * nb)
    s.call("sepgpu_feed_profile", b"A", L, nb, out.ctypes.data)
    x = s.get(capi.F_X); v = s.get(capi.F_V)
    sel = types == ord("A")
    b = np.clip((x[sel, 2] / (L / nb)).astype(np.int64), 0, nb - 1)                           # source/sepsampler.c:1449-1451
    ms, vs = m[sel], v[sel]
    want = np.concatenate([np.bincount(b, ms * vs[:, 0], nb), np.bincount(b, ms, nb),
                           np.bincount(b, ms * (vs[:, 1] ** 2 + vs[:, 2] ** 2), nb), np.bincount(b, None, nb)])
    assert np.array_equal(out[3 * nb:], want[3 * nb:])
    assert np.abs(out - want).max() <= 1e-11 * np.abs(want).max()
    s.close()


def test_fourier_feed():
    s, step, L, types, m = _evolved()
    nw = 4
    k = 2 * np.pi * np.arange(1, nw + 1) / L
    out = np.zeros(16 * nw)
    s.call("sepgpu_feed_fourier", L, nw, k.ctypes.data, out.ctypes.data)
    x = s.get(capi.F_X); v = s.get(capi.F_V); cr = s.get(capi.F_CROSSINGS); a = s.get(capi.F_A)
    ytrue = x[:, 1] + cr[:, 1] * L                                                              # sep_eval_xtrue
    ekin = 0.5 * m * (v ** 2).sum(axis=1)
    w = [np.ones_like(m), m, m * v[:, 0], m * v[:, 1], ekin, m * a[:, 1], m * v[:, 1] ** 2]   # source/sepsampler.c:962-981
    for n in range(nw):
        e = np.exp(1j * k[n] * ytrue)
        for q, wq in enumerate(w):
            want = (wq * e).sum()
            got = out[16 * n + 2 * q] + 1j * out[16 * n + 2 * q + 1]
            assert abs(got - want) <= 1e-11 * np.abs(wq).sum(), (n, q, got, want)
        assert abs(out[16 * n + 14] - ekin.sum()) <= 1e-12 * ekin.sum()
    s.close()


def test_radial_feed_counts_are_exact():
    s, step, L, types, m = _evolved()
    lvec, tl = 40, b"AB"
    ncomb = 3
    cnt = np.zeros((lvec, ncomb), dtype=np.int64)
    s.call("sepgpu_feed_radial", L, lvec, len(tl), tl, cnt.ctypes.data)
    x = s.get(capi.F_X)
    n = len(x)
    dg = 0.5 * L / lvec
    i, j = np.triu_indices(n, 1)                                                                # source/sepsampler.c:390-410
    d = x[i] - x[j]
    d = np.where(d > 0.5 * L, d - L, np.where(d < -0.5 * L, d + L, d))
    r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    idx = (np.sqrt(r2) / dg).astype(np.int64)
    keep = idx < lvec
    ti, tj = types[i][keep], types[j][keep]
    want = np.zeros_like(cnt)
    A, B = ord("A"), ord("B")
    np.add.at(want[:, 0], idx[keep][(ti == A) & (tj == A)], 1)
    np.add.at(want[:, 1], idx[keep][ti != tj], 1)
    np.add.at(want[:, 2], idx[keep][(ti == B) & (tj == B)], 1)
    assert want.sum() > 0 and np.array_equal(cnt, want)
    s.close()


def _drive(outdir, feeds, steps=650):
    env = dict(os.environ, SEP_SAMPLER_FEEDS="1" if feeds else "0")
    if capi.LIB_PATH.endswith("libsep_emu.so"):
        env["SEPGPU_EMU_LIB"] = capi.LIB_PATH
    r = subprocess.run([sys.executable, os.path.join(cm.ROOT, "tests", "feeds_driver.py"), outdir, str(steps)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    rec = r.stdout.split()
    return int(rec[rec.index("feed_calls") + 1]), int(rec[rec.index("get_calls") + 1])


def test_sampled_run_through_sep_api_writes_the_same_files(tmp_path):
    """prg1's sampler set (vacf, sacf, msd) plus profs, radial and gh on a two-species NVT loop: with the feeds the files
    agree with the host samplers' to the printed precision, and atoms[] is never downloaded"""
    don, doff = str(tmp_path / "feeds"), str(tmp_path / "host")
    feeds_on, gets_on = _drive(don, True)
    feeds_off, gets_off = _drive(doff, False)
    assert feeds_on > 0 and gets_on == 0, (feeds_on, gets_on)
    assert feeds_off == 0 and gets_off >= 600, (feeds_off, gets_off)
    files = sorted(os.listdir(doff))
    assert {"vacf.dat", "sacf.dat", "msd.dat", "msd-gaussparam.dat", "msd-incoherent.dat", "profs.dat", "radial.dat",
            "gh-X-acf.dat", "gh-energy-acf.dat", "gh-rho-acf.dat"} <= set(files)
    for f in files:
        pa, pb = os.path.join(doff, f), os.path.join(don, f)
        assert os.path.exists(pb), f
        if f == "radial_info.dat":
            assert open(pa).read() == open(pb).read()
            continue
        A, B = np.loadtxt(pa, ndmin=2), np.loadtxt(pb, ndmin=2)
        assert A.shape == B.shape and A.size > 0, (f, A.shape, B.shape)
        # two NVT runs whose thermostat term is applied at different points of the step differ in the last bits
        assert np.allclose(A, B, rtol=0, atol=5e-6, equal_nan=True), (f, np.nanmax(np.abs(A - B)))
