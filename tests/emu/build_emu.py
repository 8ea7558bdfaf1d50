"""Builds tests/_build/libsep_emu.so: the library's unchanged host C layer and its unchanged CUDA sources, the latter
rewritten by prep.py (launch syntax, inline PTX) and compiled with g++ against the emulator's cuda_runtime.h.
TEST INFRASTRUCTURE ONLY -- libsep.so never contains any of this."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GPU_SRC = os.path.join(ROOT, "seplib_b200", "csrc", "gpu")
HOST_SRC = os.path.join(ROOT, "seplib_b200", "csrc", "host")
# SEPGPU_EMU_SANITIZE=1: a second build with UBSan's alignment / bounds / null checks.  The CUDA vector types keep their
# device alignment here (d4 32 bytes, uint4 / float4 / double2 16 bytes), so a misaligned 128- or 256-bit access -- a
# fault on the GPU, silent on x86 -- aborts the test instead.
SANITIZE = os.environ.get("SEPGPU_EMU_SANITIZE") == "1"
OUT = os.path.join(ROOT, "tests", "_build", "emu_ubsan" if SANITIZE else "emu")
LIB = os.path.join(ROOT, "tests", "_build", "libsep_emu_ubsan.so" if SANITIZE else "libsep_emu.so")

sys.path.insert(0, HERE)
import prep  # noqa: E402


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    """build (if stale) under a file lock: test processes that start side by side must not compile into the same files"""
    import fcntl
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build(force, verbose)


def _build(force, verbose):
    srcs = [os.path.join(d, f) for d in (GPU_SRC, HOST_SRC, HERE, os.path.join(ROOT, "include")) for f in os.listdir(d)
            if f.endswith((".cu", ".cuh", ".c", ".h", ".cpp", ".py"))]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    prep.main(GPU_SRC, OUT)
    inc = ["-I" + HERE, "-I" + OUT, "-I" + os.path.join(ROOT, "include")]
    # -ffp-contract=off: no implicit FMA (explicit fma() calls use the hardware instruction through -mfma)
    cxx = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-mfma", "-ffp-contract=off", "-fno-strict-aliasing", "-w"] + inc
    if SANITIZE:
        cxx += ["-fsanitize=alignment,bounds,null", "-fno-sanitize-recover=all"]
    cc = ["gcc", "-std=c99", "-O2", "-fPIC", "-D_POSIX_C_SOURCE=200809L", "-I" + os.path.join(ROOT, "include")]
    objs = []
    jobs = []
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".cpp"):
            o = os.path.join(OUT, f[:-4] + ".o")
            jobs.append((cxx + ["-c", os.path.join(OUT, f), "-o", o], o))
    for src in ("emu_rt", "fake_nccl"):
        o = os.path.join(OUT, src + ".o")
        jobs.append((cxx + ["-D_GNU_SOURCE", "-c", os.path.join(HERE, src + ".cpp"), "-o", o], o))
    for f in sorted(os.listdir(HOST_SRC)):
        if f.endswith(".c"):
            o = os.path.join(OUT, "host_" + f[:-2] + ".o")
            jobs.append((cc + ["-c", os.path.join(HOST_SRC, f), "-o", o], o))
    procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd, o) for cmd, o in jobs]
    for p, cmd, o in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError("emulator build failed:\n" + " ".join(cmd) + "\n" + out[-6000:])
        if verbose and out.strip():
            print(out)
        objs.append(o)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + (["-fsanitize=alignment,bounds,null"] if SANITIZE else [])
                          + ["-lm", "-ldl", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
