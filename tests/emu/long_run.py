"""Long NVT run on the CPU kernel emulator, default kernels against every Lennard-Jones option (pair_tile, cell_order,
build_prune, step_fold, fin_multi): per-step sums, trigger steps, boundary crossings and final positions.  Not part of the
suite (about four minutes for 300 steps of 1728 atoms: 75 rebuilds, a few hundred boundary crossings).
    python tests/emu/long_run.py [steps]"""
import sys, os, ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import numpy as np
import build_emu
from seplib_b200 import capi
capi.LIB_PATH = build_emu.build()
import common as cm
x, L = cm.lattice(12, 0.8, jitter=0.05, seed=41)
n = len(x); v = cm.velocities(n, 3.0, seed=42)
def run(opts, nsteps):
    s = capi.System(n); s.put(capi.F_X, x); s.put(capi.F_V, v)
    for k, val in opts.items(): s.call("sepgpu_set_option", k.encode(), val)
    sys_ = capi.make_sys([L] * 3, 2.5, 0.005, skin=0.25)
    p = capi.lj_param(2.5, kind="lj_shift")
    s.call("sepgpu_set_alpha", 0, 0.0)
    rec = []
    for step in range(nsteps):
        s.call("sepgpu_reset_ret"); s.call("sepgpu_reset_force")
        s.call("sepgpu_force_lj", C.byref(sys_), b"AA", C.byref(p), 1, 1)
        s.call("sepgpu_nosehoover", C.byref(sys_), 3.0, 0, 0.1)
        s.call("sepgpu_leapfrog", C.byref(sys_))
        sc = s.scalars()
        rec.append((sc.epot, sc.ekin, sc.alpha[0], sc.pot_P[0], sc.nbuild, sc.neighb_flag))
    cr = s.get(capi.F_CROSSINGS)
    xf = s.get(capi.F_X)
    s.close()
    return np.array(rec), cr, xf
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
a, cra, xa = run({}, N)
b, crb, xb = run({"pair_tile": 1, "cell_order": 1, "build_prune": 1, "step_fold": 1, "fin_multi": 1}, N)
rel = np.abs(a[:, :4] - b[:, :4]) / np.maximum(np.abs(a[:, :4]), 1e-3)
print("steps", N, "builds", int(a[-1, 4]), int(b[-1, 4]), "trigger steps equal", np.array_equal(a[:, 5], b[:, 5]),
      "atoms that crossed a boundary", int((np.abs(cra).sum(axis=1) > 0).sum()), "crossings equal", np.array_equal(cra, crb))
for k in (10, 50, 100, 200, N - 1):
    if k < N: print("step", k, "rel diff epot/ekin/alpha/Pxx", rel[k])
d = xa - xb; d -= L * np.round(d / L)
print("max |dx| final", np.abs(d).max())
