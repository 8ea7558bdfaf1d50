"""Candidates tested per atom by the tiled list builder with and without build_prune / pair_tile, from the emulator's
work counter.    python tests/emu/prune_work.py"""
import sys, os, ctypes as C, itertools
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import numpy as np
import build_emu
from seplib_b200 import capi
capi.LIB_PATH = build_emu.build()
import common as cm
lib = capi.load()
cnt = (C.c_longlong * 8).in_dll(lib, "sepgpu_emu_counter")
x, L = cm.lattice(20, 0.8, jitter=0.25, seed=3)
for pt, prune in itertools.product((0, 1), (0, 1)):
    s = capi.System(len(x)); s.put(capi.F_X, x)
    s.call("sepgpu_set_option", b"pair_tile", pt); s.call("sepgpu_set_option", b"build_prune", prune); s.call("sepgpu_set_option", b"cell_order", 1)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    cnt[0] = 0
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    print("pair_tile", pt, "prune", prune, "candidates tested per atom", cnt[0] / len(x), "cells", list(sys_.nsubbox))
    s.close()
