#!/usr/bin/env python
"""Runs a sampled two-species Lennard-Jones NVT loop through the sep_* API of libsep.so in its own process, the samplers
writing into <outdir>:
    python tests/feeds_driver.py <outdir> <steps> [brute]
The caller sets SEP_SAMPLER_FEEDS (read once per process, seplib_b200/csrc/host/sep_sampler.c) -- tests/test_gpu_zzzz_feeds.py
runs the same loop with the sampler feeds on and off and compares the files.  FEEDS_NCELL (lattice side, default 8) and
FEEDS_SAMPLERS (comma list out of vacf,sacf,msd,profs,radial,gh; default all) size the run for bench.py's sampled-run
record (radial is an all-pairs histogram on both sides: leave it out of large systems); the wall time of the loop is
printed as loop_s."""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import common as cm  # noqa: E402
from seplib_b200 import capi  # noqa: E402


class OurSampler(C.Structure):              # include/sep.h
    _fields_ = [("molptr", C.c_void_p), ("impl", C.c_void_p), ("msd_counter", C.c_ulong)]


def main():
    out, steps = sys.argv[1], int(sys.argv[2])
    brute = len(sys.argv) > 3 and sys.argv[3] == "brute"
    if os.environ.get("SEPGPU_EMU_LIB"):          # test run on the CPU kernel emulator (tests/emu/run_on_emu.py)
        capi.LIB_PATH = os.environ["SEPGPU_EMU_LIB"]
    lib = capi.load()
    lib.sep_init_sampler.restype = OurSampler
    lib.sep_init_sampler.argtypes = []
    lib.sep_sample.argtypes = None
    lib.sep_close_sampler.argtypes = [C.POINTER(OurSampler)]
    lib.sep_add_sampler.restype = None
    lib.sep_add_sampler.argtypes = None
    ncell = int(os.environ.get("FEEDS_NCELL", "8"))
    which = os.environ.get("FEEDS_SAMPLERS", "vacf,sacf,msd,profs,radial,gh").split(",")
    x, L = cm.lattice(ncell, 0.8, jitter=0.05, seed=31)
    v = cm.velocities(len(x), 1.2, seed=32)
    s = cm.ApiSystem(lib, x, L, 2.5, 0.005, v=v, nneighb=0, update=capi.SEP_BRUTE if brute else capi.SEP_LLIST_NEIGHBLIST)
    s.view["type"][: len(x) // 4] = ord("B")
    s.view["m"][: len(x) // 4] = 1.5
    os.makedirs(out, exist_ok=True)
    os.chdir(out)
    smp = lib.sep_init_sampler()

    def add(name, lvec, *rest):
        lib.sep_add_sampler(C.byref(smp), name, s.sys, C.c_int(lvec), *rest)

    if "vacf" in which:
        add(b"vacf", 20, C.c_double(1.0))
    if "sacf" in which:
        add(b"sacf", 10, C.c_double(0.5))
    if "msd" in which:
        add(b"msd", 15, C.c_double(1.5), C.c_int(3), C.c_int(ord("A")))
    if "profs" in which:
        add(b"profs", 10, C.c_int(ord("A")), C.c_int(2))
    if "radial" in which:
        add(b"radial", 50, C.c_int(50), C.c_char_p(b"AB"))
    if "gh" in which:
        add(b"gh", 10, C.c_double(0.5), C.c_int(3))
    fun = s.fun("sep_lj_shift")
    alpha = C.c_double(0.1)
    t0 = time.perf_counter()
    for n in range(steps):
        lib.sep_reset_retval(s.R); lib.sep_reset_force(s.atoms, s.S)
        lib.sep_force_pairs(s.atoms, b"AA", 2.5, fun, s.S, s.R, 1)
        lib.sep_force_pairs(s.atoms, b"AB", 2.5, fun, s.S, s.R, 1)
        lib.sep_force_pairs(s.atoms, b"BB", 2.5, fun, s.S, s.R, 1)
        lib.sep_nosehoover(s.atoms, 1.2, C.byref(alpha), 0.1, s.S)
        lib.sep_leapfrog(s.atoms, s.S, s.R)
        lib.sep_sample(s.atoms, C.byref(smp), s.R, s.sys, C.c_uint(n))
    lib.sep_close_sampler(C.byref(smp))
    loop_s = time.perf_counter() - t0
    feeds, gets = C.c_longlong(), C.c_longlong()
    ctx = lib.sep_gpu_handle(s.atoms)
    lib.sepgpu_get_option(C.c_void_p(ctx), b"feed_calls", C.byref(feeds))
    lib.sepgpu_get_option(C.c_void_p(ctx), b"get_calls", C.byref(gets))
    print("feed_calls %d get_calls %d epot %.12f natoms %d loop_s %.6f" % (feeds.value, gets.value, s.ret.epot, len(x), loop_s))
    s.close()


if __name__ == "__main__":
    main()
