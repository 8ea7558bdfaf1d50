"""Work-model numbers quoted in DESIGN.md section 3a, recomputed on the CPU kernel emulator (run by tests/test_cpu_emu.py):
distinct 128-byte lines per warp-wide gather of the list force kernels, L1 wavefronts per atom with and without pair-tile
lists, candidates tested per atom by the list builder with and without pruning.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

os.environ["SEPGPU_EMU_GATHER_STATS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import build_emu  # noqa: E402
from seplib_b200 import capi  # noqa: E402

capi.LIB_PATH = build_emu.build()
import common as cm  # noqa: E402

lib = capi.load()
cnt = (C.c_longlong * 8).in_dll(lib, "sepgpu_emu_counter")
x, L = cm.lattice(18, 0.8, jitter=0.25, seed=3)
x = np.ascontiguousarray(x[np.random.default_rng(1).permutation(len(x))])     # atom index unrelated to position
n = len(x)
out = {}
for name, opts in (("per_atom", {}), ("pair_tile", {"pair_tile": 1, "cell_order": 1}), ("pruned", {"build_prune": 1})):
    s = capi.System(n); s.put(capi.F_X, x)
    for k, v in opts.items():
        s.call("sepgpu_set_option", k.encode(), v)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
    for k in range(8):
        cnt[k] = 0
    s.call("sepgpu_neighb_build", C.byref(sys_), 1)
    cand = cnt[0] / n
    for k in range(8):
        cnt[k] = 0
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
    out[name] = {"candidates_per_atom": cand, "lines_per_request": cnt[2] / max(cnt[1], 1), "wavefronts_per_atom": cnt[2] / n,
                 "lane_loads_per_atom": cnt[3] / n}
    s.close()
print(json.dumps(out))
