import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # test runs on the CPU kernel emulator (tests/emu/run_on_emu.py sets the variable): worker processes of a parallel
    # run (-n) follow the runner to the emulated build.  Test plumbing only -- seplib_b200/capi.py has no such switch.
    if os.environ.get("SEPGPU_EMU_LIB"):
        from seplib_b200 import capi
        capi.LIB_PATH = os.environ["SEPGPU_EMU_LIB"]


@pytest.fixture(scope="session")
def lib():
    from seplib_b200 import capi
    return capi.load()
