"""Work model of the list force kernels on the CPU kernel emulator: warp-wide gather requests and the distinct 128-byte
lines they touch (= L1 wavefronts, the unit k_lj_list is bound by; ncu on B200 measured ~19-22 per request, the emulator
counts 20.5) for the per-atom list and the pair-tile list, with the in-cell slot order by atom index or along the Morton
curve.    python tests/emu/gather_stats.py [lattice side, default 24]"""
import sys, os, ctypes as C, itertools
os.environ["SEPGPU_EMU_GATHER_STATS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import numpy as np
import build_emu
from seplib_b200 import capi
capi.LIB_PATH = build_emu.build()
import common as cm
lib = capi.load()
cnt = (C.c_longlong * 8).in_dll(lib, "sepgpu_emu_counter")
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x, L = cm.lattice(ncell, 0.8, jitter=0.25, seed=3)
xsh = np.ascontiguousarray(x[np.random.default_rng(1).permutation(len(x))])
n = len(x)
for (label, xx), pt, co in itertools.product((("lattice-order", x), ("shuffled", xsh)), (0, 1), (0, 1)):
    s = capi.System(n); s.put(capi.F_X, xx)
    s.call("sepgpu_set_option", b"pair_tile", pt); s.call("sepgpu_set_option", b"cell_order", co)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    for k in range(8): cnt[k] = 0
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
    req, lines, lanes = cnt[1], cnt[2], cnt[3]
    print(f"{label:14s} pair_tile={pt} cell_order={co}: warp requests/atom {req*32/n:7.1f}  lines per request {lines/max(req,1):5.2f}  "
          f"L1 wavefronts per atom {lines/n:7.2f}  lane loads per atom {lanes/n:6.1f}  cells {list(sys_.nsubbox)}")
    s.close()
