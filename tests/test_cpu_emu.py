"""The library's own CUDA kernels, executed on the CPU by the kernel emulator (tests/emu/): the unchanged .cu sources
are compiled with g++, every CUDA thread of a block runs as a fiber and __syncthreads / the *_sync warp primitives are
rendezvous points of the fiber scheduler.  ALL `-m gpu` parity tests except the linked example programs then run against
that build (in parallel worker processes).

What this proves: the kernels' logic (indexing, list layout, masks, scans, reductions, control flow of the C ABI around
them) against the oracle, on every CPU round.  What it cannot prove: memory-model races, alignment faults, resource
limits and speed -- those stay with the hardware run.  The emulated library is test infrastructure: libsep.so contains
none of it and has no CPU path (tests/test_cpu_host.py::test_no_cpu_fallback_without_device).
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "emu", "run_on_emu.py")
WORKERS = str(max(1, min(4, (os.cpu_count() or 2) // 2)))       # the emulator is one busy thread per test process

# the linked example programs (10^4 steps, minutes each on the emulator) are left to the hardware run
FAST = "not prg and not nvt_1000"      # (and the 1000-step trajectory: 160 s on the emulator)


def _run(args, timeout=1500, env=None):
    r = subprocess.run([sys.executable, RUNNER, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                       env=dict(os.environ, **(env or {})))
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m, tail
    assert "failed" not in r.stdout and "error" not in r.stdout.lower().replace("test_error_paths", ""), tail
    return int(m.group(1))


def test_gpu_parity_tests_pass_on_the_emulated_kernels():
    # tests/test_golden.py: the reference's own recorded butane / water / DPD vectors against the emulated kernels
    n = _run(["tests/test_gpu_lj.py", "tests/test_gpu_more.py", "tests/test_gpu_zz_next.py", "tests/test_gpu_zzzz_edge.py",
              "tests/test_gpu_zzzz_omp.py", "tests/test_golden.py", "-m", "gpu", "-q", "-n", WORKERS, "-k", FAST, "-p", "no:cacheprovider"])
    assert n >= 49, n


def test_kernel_options_pass_on_the_emulator():
    """The kernel options and list formats (tests/test_gpu_zzz_options.py): tile rows against global-index rows, the
    alternative Coulomb / sub-list / finalisation kernels against the ones they replaced."""
    n = _run(["tests/test_gpu_zzz_options.py", "-m", "gpu", "-q", "-n", WORKERS, "-p", "no:cacheprovider"])
    assert n >= 25, n


def test_sampler_feeds_pass_on_the_emulator():
    """The sampler feeds (tests/test_gpu_zzzz_feeds.py): every feed against numpy on the downloaded arrays, and a sampled
    run through the sep_* API with SEP_SAMPLER_FEEDS=1 against the host samplers (same files, no download of atoms[])."""
    n = _run(["tests/test_gpu_zzzz_feeds.py", "-m", "gpu", "-q", "-n", WORKERS, "-p", "no:cacheprovider"])
    assert n >= 6, n


def test_two_rank_decomposition_on_the_emulator():
    """Slab decomposition with two ranks as two THREADS on the emulated kernels (tests/emu/dd_threads.py): union of the
    ranks' pair sets == single-domain set, per-step sums, trigger steps, final positions, atom conservation over several
    rebuilds with migration -- the checks tests/dd_check.py makes on two GPUs.  Both transport paths: peer memory (the
    emulator hands out in-process IPC handles; the two ranks' kernels run concurrently and meet at release/acquire
    flags), also with global-index rows and the unfolded step, and NCCL send/recv + all-reduce (tests/emu/fake_nccl.cpp).
    The three runs go side by side."""
    cases = {"peer-memory": ("", "0", "2", "14"), "peer-memory+old-kernels": ("step_fold=0,tile_list=0,fin_multi=0", "0", "2", "14"),
             "nccl-path": ("", "1", "2", "14"), "peer-memory-3-ranks": ("", "0", "3", "18")}
    procs = {k: subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "emu", "dd_threads.py"), ncell, "14", opts],
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                                 env=dict(os.environ, SEPGPU_EMU_NO_IPC=no_ipc, DD_WORLD=world))
             for k, (opts, no_ipc, world, ncell) in cases.items()}
    # bonded terms in decomposed runs (tests/dd_mol.py): the reference's butane cell on two slabs against its golden vectors
    # and a single-domain run, on both transports
    for k, no_ipc in (("butane-peer-memory", "0"), ("butane-nccl-path", "1")):
        procs[k] = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "emu", "dd_threads.py"), "butane", "24"],
                                    stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                                    env=dict(os.environ, SEPGPU_EMU_NO_IPC=no_ipc, DD_WORLD="2"))
    # ... and prg3's force sequence with the Coulomb sum (charges by global id) on two slabs against one domain
    procs["water-peer-memory"] = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "emu", "dd_threads.py"), "water", "12"],
                                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                                                  env=dict(os.environ, SEPGPU_EMU_NO_IPC="0", DD_WORLD="2"))
    for k, p in procs.items():
        out, err = p.communicate(timeout=900)
        assert p.returncode == 0 and "-> OK" in out, (k, out[-2000:], err[-2000:])


def test_smoke_entry_point_on_the_emulator():
    """__graft_entry__.smoke() -- the one-step check the driver runs on cuda:0 -- against the emulated kernels."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import build_emu\n"
            "from seplib_b200 import capi\n"
            "capi.LIB_PATH = build_emu.build()\n"
            "import __graft_entry__ as g\n"
            "g.smoke()\n") % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "smoke ok" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
