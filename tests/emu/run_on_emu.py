"""Runs pytest with the ctypes bindings pointed at tests/_build/libsep_emu.so (the library's own kernels executed by
the CPU kernel emulator) instead of libsep.so.  TEST INFRASTRUCTURE ONLY: the redirection lives here, in the test
runner -- seplib_b200/capi.py has no switch for it, and libsep.so has no CPU path.

    python tests/emu/run_on_emu.py tests/test_gpu_lj.py -m gpu -x -q
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import build_emu  # noqa: E402
from seplib_b200 import capi  # noqa: E402

capi.LIB_PATH = build_emu.build()
os.environ["SEPGPU_EMU_LIB"] = capi.LIB_PATH   # subprocess drivers (tests/next_driver.py, linked prgs) follow it
_d = os.path.join(os.path.dirname(capi.LIB_PATH), "emu_lib")
os.makedirs(_d, exist_ok=True)
if not os.path.lexists(os.path.join(_d, "libsep.so")):
    os.symlink(capi.LIB_PATH, os.path.join(_d, "libsep.so"))
sys.stderr.write("seplib-b200 TEST RUN ON THE CPU KERNEL EMULATOR: %s\n" % capi.LIB_PATH)

import pytest  # noqa: E402

sys.exit(pytest.main(sys.argv[1:]))
