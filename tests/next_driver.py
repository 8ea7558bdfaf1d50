#!/usr/bin/env python
"""Runs one of the section-8f loops of tests/common.py on libsep.so in its own process and saves what it recorded:
    python tests/next_driver.py <compress|ber|beriso|slit|fp|gjf> <sync mode 0|1> <out.npz>
A separate process because the sep_* API reports errors the way the reference does -- sep_error() prints and
exit()s -- which must fail one test, not end the pytest run."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import common as cm  # noqa: E402
from seplib_b200 import capi  # noqa: E402


def main():
    what, sync, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    g = np.load(os.path.join(cm.GOLDEN, "next_rows.npz"))
    if os.environ.get("SEPGPU_EMU_LIB"):          # test run on the CPU kernel emulator (tests/emu/run_on_emu.py)
        capi.LIB_PATH = os.environ["SEPGPU_EMU_LIB"]
    lib = capi.load()
    lib.sep_gpu_set_sync(sync)
    if what == "compress":
        rec = cm.drive_compress(lib, g["c_x0"], g["c_v0"], float(g["c_L"]))
    elif what == "ber":
        rec = cm.drive_berendsen(lib, g["b_x0"], g["b_v0"], float(g["b_L"]))
    elif what == "beriso":
        rec = cm.drive_berendsen(lib, g["c_x0"], g["c_v0"], float(g["c_L"]), steps=12, iso=True, update=capi.SEP_LLIST_NEIGHBLIST)
    elif what == "slit":
        rec = cm.drive_slit(lib, g["c_x0"], g["c_v0"], float(g["c_L"]))
    elif what in ("fp", "gjf"):
        rec = cm.drive_stochastic(lib, g["b_x0"], g["b_v0"], float(g["b_L"]), what)
    else:
        raise SystemExit("unknown scenario " + what)
    np.savez(out, **rec)


if __name__ == "__main__":
    main()
