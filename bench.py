#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the seplib hot path on B200 (BASELINE.json metric).

A "step" is one MD time step of the whole system:
    sep_reset_retval -> sep_reset_force -> sep_force_pairs (LJ, rc=2.5, list rebuilt when the skin
    trigger fired) -> sep_nosehoover -> sep_leapfrog
on a lattice-initialised Lennard-Jones fluid (prg1-style NVT).

  value     device-resident loop through the sepgpu_* C ABI (inputs already in HBM), CUDA-event timed
  e2e       the same loop through the reference-facing sep_* API with HOST seppart[] buffers; the timed
            region contains the host->device upload, a device->host read of the step's sepret/sepsys
            scalars EVERY step, and the final download into the host array; it runs max(--steps, --e2e-steps)
            steps (default 1000) so that the one upload and the one download do not dominate a short --steps
  roofline  the pair-force kernel (dominant): algorithmic bytes / CUDA-event time vs measured HBM peak,
            plus its FP64 rate vs an FMA-chain peak measured on this box (the kernel is phase- and latency-limited
            below both)
  cpu_baseline  the reference's own OpenMP CPU path (oracle/_ref, compiled from its unmodified sources)
            on the box's host cores, bounded sample of the same workload

            (+ cpu_baseline.reference_cuda: the reference's own CUDA path, oracle/_ref/refcuda_lj, on the same GPU)
  side records, each a bounded job of its own whose outcome never touches the numbers above:
            other_workloads (C2 butane / C3 water), sampled_run (run-time samplers: host path against device feeds);
            at N = 2 dd_water_check (decomposed Coulomb against one GPU), e2e_sep_ngpu (an unchanged C program with SEP_NGPU=2),
            spec_force2_trial (the experimental launch sent ahead in a decomposed run).  --no-other skips them all.

--impl reference runs only that CPU arm and prints its own line.
N>1 (torchrun): slab domain decomposition, N x 1 M atoms (weak scaling; 8 GPUs = the 8 M-atom C4 configuration);
--replicas runs N independent copies instead.  --workload butane|water runs the C2 / C3 configurations (1 GPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "atom-timesteps/sec (LJ, rc=2.5)"
UNIT = "atom-timesteps/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------
def lj_lattice(ncell, rho):
    n = ncell ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / ncell
    g = (np.arange(ncell) + 0.5) * a
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    return np.ascontiguousarray(np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)), L


def lj_velocities(n, temp, seed):
    rng = np.random.default_rng(seed)
    v = rng.random((n, 3)) - 0.5
    v -= v.mean(axis=0)
    v *= np.sqrt(3 * n * temp / (v * v).sum())
    return np.ascontiguousarray(v)


ALG_FLOPS_PER_LISTED = 13.0
ALG_FLOPS_PER_INRANGE = 42.0


def algorithmic_per_atom(rho, rc, skin):
    """half-list pair counts, flops and compulsory bytes per atom-step."""
    n_list = (2.0 * np.pi / 3.0) * rho * (rc + skin) ** 3
    n_in = (2.0 * np.pi / 3.0) * rho * rc ** 3
    flops = (ALG_FLOPS_PER_LISTED + ALG_FLOPS_PER_INRANGE) * n_in + ALG_FLOPS_PER_LISTED * (n_list - n_in)
    nbytes = 4.0 * n_list + 32.0 + 32.0
    return n_list, n_in, flops, nbytes


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:      # noqa: BLE001
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            p = [q.strip() for q in line.split(",")]
            if len(p) < 9:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active") and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        if not sm:      # region shorter than a sample period: use everything we saw
            for ts, line in self.rows:
                p = [q.strip() for q in line.split(",")]
                try:
                    sm.append(float(p[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_arm(ncell, rho, rc, skin, dt, temp, tau, target_seconds, threads):
    """Runs the prg1/prg4-style loop through the reference API (oracle/_ref/libsep_ref_fast.so, built from
    the unmodified reference sources with its shipped flags -Ofast -fopenmp).  Falls back to the C port
    (oracle/liboracle.so, scalar) when the reference build is absent.  Returns (value, info dict)."""
    from seplib_b200 import capi
    fast = os.path.join(ROOT, "oracle", "_ref", "libsep_ref_fast.so")
    x, L = lj_lattice(ncell, rho)
    n = len(x)
    v = lj_velocities(n, temp, 12345)
    if os.path.exists(fast):
        lib = C.CDLL(fast, mode=C.RTLD_LOCAL)
        capi.declare_sep_api(lib)
        atoms = lib.sep_init(n, 3000)
        sys_ = lib.sep_sys_setup(L, L, L, rc, dt, n, capi.SEP_LLIST_NEIGHBLIST)
        view = capi.atoms_view(atoms, n)
        view["x"][:] = x
        view["v"][:] = v
        view["xn"][:] = 0.0
        if threads > 1:
            lib.sep_set_omp(threads, C.byref(sys_))
        lib.sep_set_skin(C.byref(sys_), skin)
        ret = capi.SepRet()
        alpha = C.c_double(0.1)
        fun = C.cast(lib.sep_lj_shift, C.c_void_p)

        def step():
            lib.sep_reset_retval(C.byref(ret))
            lib.sep_reset_force(atoms, C.byref(sys_))
            lib.sep_force_pairs(atoms, b"AA", rc, fun, C.byref(sys_), C.byref(ret), capi.SEP_ALL)
            lib.sep_nosehoover(atoms, temp, C.byref(alpha), tau, C.byref(sys_))
            lib.sep_leapfrog(atoms, C.byref(sys_), C.byref(ret))

        for _ in range(3):
            step()
        t0 = time.perf_counter()
        step(); step()
        per = (time.perf_counter() - t0) / 2
        steps = int(max(5, min(2000, target_seconds / max(per, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        el = time.perf_counter() - t0
        epot = ret.epot / n
        lib.sep_close(atoms, n)
        kind = "reference"
    else:
        import common as cm
        orc = cm.oracle()
        threads = 1
        m = np.ones(n); types = np.full(n, ord("A"), dtype=np.uint8)
        xn = np.zeros((n, 3)); cn = np.zeros((n, 3), dtype=np.int32); cr = np.zeros((n, 3), dtype=np.int32)
        a = np.zeros((n, 3)); length = np.array([L, L, L])
        state = {"flag": 1, "pairs": None, "alpha": 0.1}
        ret = cm.OrcRet()

        def step():
            f = np.zeros((n, 3)); md2 = C.c_double(0.0)
            r = cm.OrcRet()
            if state["flag"]:
                state["pairs"] = np.ascontiguousarray(cm.oracle_pairs(x, L, rc, skin), dtype=np.int32)
            p = state["pairs"]
            orc.orc_force_pairs_list(n, cm.ptr(x), cm.ptr(types), cm.ptr(length), cm.ptr(p), len(p), b"AA", rc,
                                     cm.POT_LJ_SHIFT, None, cm.ptr(f), C.byref(r))
            state["alpha"] = orc.orc_nosehoover(n, cm.ptr(v), cm.ptr(m), cm.ptr(f), temp, state["alpha"], tau, dt)
            state["flag"] = orc.orc_leapfrog(n, cm.ptr(x), cm.ptr(v), cm.ptr(f), cm.ptr(m), cm.ptr(a), cm.ptr(xn),
                                             cm.ptr(cn), cm.ptr(cr), cm.ptr(length), dt, skin, C.byref(md2), C.byref(r))
            ret.epot = r.epot

        for _ in range(3):
            step()
        t0 = time.perf_counter()
        step(); step()
        per = (time.perf_counter() - t0) / 2
        steps = int(max(5, min(2000, target_seconds / max(per, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        el = time.perf_counter() - t0
        epot = ret.epot / n
        kind = "port"
    value = n * steps / el
    info = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{n} atoms (lattice {ncell}^3, rho={rho}, rc={rc}, skin={skin}, NH tau={tau}) x {steps} steps "
                      f"in {el:.1f} s, epot/N={epot:.4f}",
            "seconds": el, "steps": steps, "natoms": n}
    return value, info


def cpu_arm_child(kind, **kw):
    """Runs one CPU arm (the reference's own code on the host cores) in a CHILD process whose stack limit is raised before
    exec: the reference keeps per-atom scratch in variable-length stack arrays (source/sepprfrc.c:428), 1 M atoms overflow
    the default 8 MB, and a limit raised inside a process that already mapped CUDA / torch cannot grow the main stack."""
    import resource

    def big_stack():
        try:
            soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
            want = resource.RLIM_INFINITY if hard == resource.RLIM_INFINITY else hard
            resource.setrlimit(resource.RLIMIT_STACK, (want, hard))
        except (ValueError, OSError):
            pass

    env = dict(os.environ)
    env.setdefault("OMP_STACKSIZE", "512M")
    req = json.dumps(dict(kind=kind, **kw))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-arm-json", req], capture_output=True, text=True,
                       preexec_fn=big_stack, env=env)
    if r.returncode != 0:
        raise RuntimeError("CPU arm failed (rc %d): %s" % (r.returncode, (r.stdout + r.stderr)[-400:]))
    return json.loads(r.stdout.strip().splitlines()[-1])


def cpu_arm_main(req):
    """child side of cpu_arm_child"""
    q = json.loads(req)
    if q.pop("kind") == "lj":
        _, info = cpu_reference_arm(**q)
    else:
        _, info = mol_cpu_arm(**q)
    print(json.dumps(info))
    return 0


def cpu_matrix(rho, rc, dt, temp, tau, seconds_each, skip_big):
    """SURVEY.md section 8d / BASELINE.md section 3: the reference's OpenMP build, prg4-style LJ at N = 10 648 / 110 592 /
    1 000 000, one thread and all cores, skin 0.25 and 1.0 (prg4's setting), plus prg5-style butane.  Bounded samples."""
    rows = []
    allc = host_cores()
    plan = [(22, 1, 0.25), (22, allc, 0.25), (22, allc, 1.0), (48, 1, 0.25), (48, 1, 1.0), (48, allc, 0.25), (48, allc, 1.0)]
    if not skip_big:
        plan.append((100, allc, 1.0))
    for ncell, thr, skin in plan:
        try:
            info = cpu_arm_child("lj", ncell=ncell, rho=rho, rc=rc, skin=skin, dt=dt, temp=temp, tau=tau,
                                 target_seconds=seconds_each, threads=thr)
            rows.append({"workload": "lj", "natoms": info["natoms"], "threads": thr, "skin": skin, "value": info["value"],
                         "steps": info["steps"], "seconds": info["seconds"], "kind": info["kind"]})
        except Exception as e:      # noqa: BLE001
            rows.append({"workload": "lj", "natoms": ncell ** 3, "threads": thr, "skin": skin, "value": None, "error": repr(e)})
    for thr in (1, allc):
        try:
            info = cpu_arm_child("mol", name="butane", skin=0.25, target_seconds=seconds_each, threads=thr)
            rows.append({"workload": "butane (prg2/prg5 force sequence)", "natoms": info["natoms"], "threads": thr, "skin": 0.25,
                         "value": info["value"], "steps": info["steps"], "seconds": info["seconds"], "kind": info["kind"]})
        except Exception as e:      # noqa: BLE001
            rows.append({"workload": "butane", "threads": thr, "value": None, "error": repr(e)})
    return rows


def h2d_bytes_lj(n):
    return n * (24 * 3 + 8 * 2 + 1 + 4 + 12 * 2)            # x, v, xn, m, z, type, molindex, cross_neighb, crossings


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# molecular workloads C2 (butane) and C3 (water): SURVEY.md section 8d, single GPU
# ---------------------------------------------------------------------------------------------------
def mol_api_system(lib, capi, w, P, update, nneighb, skin):
    """The workload as a sep_* API system (host seppart[] + sepsys + topology read from a .top file)."""
    import tempfile
    from seplib_b200 import workloads as wl
    n = w["n"]
    atoms = lib.sep_init(n, nneighb)
    view = capi.atoms_view(atoms, n)
    view["x"][:] = w["x"]; view["v"][:] = w["v"]; view["type"][:] = w["type"]; view["m"][:] = w["m"]; view["z"][:] = w["z"]
    view["xn"][:] = 0.0
    L = w["L"]
    hsys = lib.sep_sys_setup(L[0], L[1], L[2], P["cf"], P["dt"], n, update)
    lib.sep_set_skin(C.byref(hsys), skin)
    fd, top = tempfile.mkstemp(suffix=".top")
    os.close(fd)
    wl.write_top(w, top)
    lib.sep_read_topology_file(atoms, top.encode(), C.byref(hsys), b"q")
    os.unlink(top)
    return atoms, view, hsys


def mol_api_step(lib, capi, name, P, atoms, hsys, ret, alpha):
    fun = C.cast(lib.sep_lj_shift, C.c_void_p)
    S, R = C.byref(hsys), C.byref(ret)
    if name == "butane":
        rb = (C.c_double * 6)(*P["rb"])

        def step():          # prgs/prg2.c:51-95
            lib.sep_reset_retval(R); lib.sep_reset_force(atoms, S)
            lib.sep_force_pairs(atoms, P["types"], P["cf"], fun, S, R, capi.SEP_EXCL_SAME_MOL)
            lib.sep_stretch_harmonic(atoms, 0, P["lbond"], P["kbond"], S, R)
            lib.sep_angle_harmonic(atoms, 0, P["angle"], P["kangle"], S, R)
            lib.sep_torsion_Ryckaert(atoms, 0, rb, S, R)
            lib.sep_nosehoover(atoms, P["temp"], C.byref(alpha), P["tau"], S)
            lib.sep_leapfrog(atoms, S, R)
    else:

        def step():          # prgs/prg3.c:55-98 (without the box compression, which ended before this state)
            lib.sep_reset_retval(R); lib.sep_reset_force(atoms, S)
            lib.sep_force_pairs(atoms, P["types"], P["cf_lj"], fun, S, R, capi.SEP_EXCL_SAME_MOL)
            lib.sep_stretch_harmonic(atoms, 0, P["lbond"], P["kbond"], S, R)
            lib.sep_angle_cossq(atoms, 0, P["angle"], P["kangle"], S, R)
            lib.sep_coulomb_sf(atoms, P["cf"], S, R, capi.SEP_EXCL_SAME_MOL)
            lib.sep_nosehoover(atoms, P["temp"], C.byref(alpha), P["tau"], S)
            lib.sep_leapfrog(atoms, S, R)
    return step


def mol_cpu_arm(name, skin, target_seconds, threads):
    """Reference CPU build on a small tiling of the same unit cell (butane 4000 atoms; water 2^3 cells = 5184 atoms,
    the smallest tiling whose cell grid has the 3 cells per side list mode needs)."""
    from seplib_b200 import capi
    from seplib_b200 import workloads as wl
    fast = os.path.join(ROOT, "oracle", "_ref", "libsep_ref_fast.so")
    if not os.path.exists(fast):
        raise RuntimeError("oracle/_ref/libsep_ref_fast.so missing (the molecular CPU arm has no port fallback)")
    lib = C.CDLL(fast, mode=C.RTLD_LOCAL)
    capi.declare_sep_api(lib)
    P = wl.BUTANE if name == "butane" else wl.WATER
    w = wl.butane(1) if name == "butane" else wl.water(2)
    atoms, view, hsys = mol_api_system(lib, capi, w, P, capi.SEP_LLIST_NEIGHBLIST, 3000, skin)
    if threads > 1:
        lib.sep_set_omp(threads, C.byref(hsys))
    ret = capi.SepRet(); alpha = C.c_double(0.1)
    step = mol_api_step(lib, capi, name, P, atoms, hsys, ret, alpha)
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    step(); step()
    per = (time.perf_counter() - t0) / 2
    steps = int(max(5, min(5000, target_seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    n = w["n"]
    value = n * steps / el
    lib.sep_close(atoms, n)
    return value, {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{name}: {n} atoms x {steps} steps in {el:.1f} s, epot/N={ret.epot / n:.4f}",
                   "seconds": el, "steps": steps, "natoms": n}


def run_butane_decomposed(args, emit, local_rank, rank, world, reps):
    """C2 on N GPUs (weak scaling): the recorded butane cell tiled reps x reps x (reps * N), slabs along z, prg2's force
    sequence through the C ABI.  Bonded terms are evaluated by every rank that owns one of their atoms, partners read from
    the halo; the check against the reference's golden vectors is tests/dd_check.py DD_MOL=butane."""
    import torch
    import torch.distributed as dist
    from seplib_b200 import capi
    from seplib_b200 import workloads as wl
    P = wl.BUTANE
    K, W = args.steps, max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = wl.tiled_molecular("butane_n4000.npz", (reps, reps, reps * world))
    n_total = w["n"]
    gsys = capi.make_sys(w["L"], P["cf"], P["dt"], skin=args.skin)
    gs = C.byref(gsys)
    nz = gsys.nsubbox[2]
    z0, z1 = capi.dd_slab_range(rank, world, nz)
    cz = np.clip(np.floor(w["x"][:, 2] / gsys.lsubbox[2]).astype(np.int64), 0, nz - 1)
    mine = np.nonzero((cz >= z0) & (cz < z1))[0].astype(np.int32)
    n = len(mine)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.dd_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    s = capi.System(int(1.25 * n_total / world) + int(3.5 * n_total / nz) + 4096, device=local_rank)
    s.dd_init(rank, world, bytes(idt.cpu().numpy().tobytes()), gsys, n_total)
    s.dd_set_owned(n)
    s.put(capi.F_X, w["x"][mine]); s.put(capi.F_V, w["v"][mine]); s.put(capi.F_GID, mine)
    s.put(capi.F_TYPE, np.ascontiguousarray(w["type"][mine])); s.put(capi.F_M, np.ascontiguousarray(w["m"][mine]))
    s.put(capi.F_MOLINDEX, np.ascontiguousarray(w["molindex"][mine]))
    s.set_topology(w["blist"], w["alist"], w["dlist"])
    s.call("sepgpu_set_alpha", 0, 0.1)
    ljp = capi.lj_param(P["cf"], kind="lj_shift")
    ljr = C.byref(ljp)
    rb = (C.c_double * 6)(*P["rb"])
    ctx, fn = s.ctx, s.lib

    def step():
        fn.sepgpu_reset_ret(ctx)
        fn.sepgpu_reset_force(ctx)
        r = fn.sepgpu_force_lj(ctx, gs, P["types"], ljr, 3, 1)
        r |= fn.sepgpu_stretch_harmonic(ctx, gs, 0, P["lbond"], P["kbond"])
        r |= fn.sepgpu_angle_harmonic(ctx, gs, 0, P["angle"], P["kangle"])
        r |= fn.sepgpu_torsion_ryckaert(ctx, gs, 0, rb)
        r |= fn.sepgpu_nosehoover(ctx, gs, P["temp"], 0, P["tau"])
        r |= fn.sepgpu_leapfrog(ctx, gs)
        if r:
            raise RuntimeError("device step failed: " + fn.sepgpu_last_error().decode())

    for _ in range(W):
        step()
    nb0 = s.scalars().nbuild
    s.call("sepgpu_set_option", b"time_kernels", 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    dist.barrier(); torch.cuda.synchronize()
    t_wall0 = time.time()
    s.call("sepgpu_timer_start")
    for _ in range(K):
        step()
    ms = C.c_float()
    s.call("sepgpu_timer_stop", C.byref(ms))
    dist.barrier(); torch.cuda.synchronize()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_sec = float(t.item()) * 1e-3
    sc = s.scalars()
    kt = {}
    for which in ("force", "bonded", "build", "intgr", "halo", "migrate"):
        tot, cnt = C.c_float(), C.c_int()
        s.call("sepgpu_kernel_time", which.encode(), C.byref(tot), C.byref(cnt))
        kt[which] = (tot.value, cnt.value)
    own, halo = s.dd_layers()[2:]
    s.close()
    if rank == 0:
        emit({"metric": METRIC.replace("LJ, rc=2.5", "C2 butane"), "value": n_total * K / t_sec, "unit": UNIT, "n_gpus": world,
              "steps": K, "warmup": W, "ms_per_step": t_sec * 1e3 / K, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": "C2 butane, %d-way slab decomposition: sep_force_pairs(CC, lj_shift, EXCL_SAME_MOL) + stretch + angle + "
                                     "Ryckaert torsion + NH + leapfrog" % world,
                         "natoms_total": n_total, "natoms_per_gpu": n_total // world, "tiling": "%d x %d x %d unit cells" % (reps, reps, reps * world),
                         "box": list(map(float, w["L"])), "cells": list(gsys.nsubbox[:]), "skin": args.skin, "dt": P["dt"],
                         "parallelism": "%d-way slab domain decomposition along z, bonded partners from the halo" % world,
                         "rank0_owned_halo_atoms": [own, halo], "list_rebuilds_in_timed_region": sc.nbuild - nb0,
                         "epot_per_atom": sc.epot / n_total, "ekin_per_atom": sc.ekin / n_total,
                         "l2": "inputs larger than L2"},
              "roofline": None, "cpu_baseline": None, "e2e": None, "gpu_launches": K * 13 + (sc.nbuild - nb0) * 20, "clocks": clocks,
              "kernel_ms": {k: {"total_ms": a, "launches": b} for k, (a, b) in kt.items()}})
    dist.barrier()
    return 0


def run_molecular(args, emit, local_rank):
    import torch
    from seplib_b200 import capi
    from seplib_b200 import workloads as wl
    name = args.workload
    P = wl.BUTANE if name == "butane" else wl.WATER
    reps = args.reps or (6 if name == "butane" else 12)
    K, W = args.steps, max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    lib = capi.load()
    if lib.sepgpu_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device -- seplib-b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        if name != "butane":
            raise RuntimeError("bench.py: decomposed molecular runs shard the butane sequence (bonded terms); water needs Coulomb")
        return run_butane_decomposed(args, emit, local_rank, rank, world, reps)
    w = wl.butane(reps) if name == "butane" else wl.water(reps)
    n = w["n"]
    gsys = capi.make_sys(w["L"], P["cf"], P["dt"], skin=args.skin)
    gs = C.byref(gsys)
    ljp = capi.lj_param(P["cf"] if name == "butane" else P["cf_lj"], kind="lj_shift")
    ljr = C.byref(ljp)
    rb = (C.c_double * 6)(*P["rb"]) if name == "butane" else None

    s = capi.System(n, device=local_rank)
    s.put(capi.F_X, w["x"]); s.put(capi.F_V, w["v"]); s.put(capi.F_TYPE, w["type"]); s.put(capi.F_M, w["m"])
    s.put(capi.F_Z, w["z"]); s.put(capi.F_MOLINDEX, w["molindex"])
    s.put(capi.F_BOND, w["bond"]); s.put(capi.F_ANGLE, w["angle"]); s.put(capi.F_DIHED, w["dihed"])
    s.set_topology(w["blist"], w["alist"], w["dlist"])
    s.call("sepgpu_set_alpha", 0, 0.1)
    ctx, fn = s.ctx, lib

    def step():
        fn.sepgpu_reset_ret(ctx)
        fn.sepgpu_reset_force(ctx)
        r = fn.sepgpu_force_lj(ctx, gs, P["types"], ljr, 3, 1)
        r |= fn.sepgpu_stretch_harmonic(ctx, gs, 0, P["lbond"], P["kbond"])
        if name == "butane":
            r |= fn.sepgpu_angle_harmonic(ctx, gs, 0, P["angle"], P["kangle"])
            r |= fn.sepgpu_torsion_ryckaert(ctx, gs, 0, rb)
        else:
            r |= fn.sepgpu_angle_cossq(ctx, gs, 0, P["angle"], P["kangle"])
            r |= fn.sepgpu_coulomb_sf(ctx, gs, P["cf"], 3)
        r |= fn.sepgpu_nosehoover(ctx, gs, P["temp"], 0, P["tau"])
        r |= fn.sepgpu_leapfrog(ctx, gs)
        if r:
            raise RuntimeError("device step failed: " + fn.sepgpu_last_error().decode())

    for _ in range(W):
        step()
    nb0 = s.scalars().nbuild
    s.call("sepgpu_set_option", b"time_kernels", 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    s.call("sepgpu_timer_start")
    for _ in range(K):
        step()
    ms = C.c_float()
    s.call("sepgpu_timer_stop", C.byref(ms))
    torch.cuda.synchronize()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    sc = s.scalars()
    nbuild = sc.nbuild - nb0
    kt = {}
    for which in ("force", "coulomb", "bonded", "build", "intgr"):
        tot, cnt = C.c_float(), C.c_int()
        s.call("sepgpu_kernel_time", which.encode(), C.byref(tot), C.byref(cnt))
        kt[which] = (tot.value, cnt.value)
    t_sec = ms.value * 1e-3
    value = n * K / t_sec
    full_entries = sc.npairs_listed / n
    epotN, ekinN = sc.epot / n, sc.ekin / n
    s.close()

    # e2e: the sep_* API on a host seppart[] array, topology through the .top reader (SEP_SYNC=lazy)
    e2e = None
    if not args.no_e2e:
        Ke = max(10, min(K, 200))
        lib.sep_gpu_set_sync(3)          # the default coherence mode (auto)
        atoms, view, hsys = mol_api_system(lib, capi, w, P, capi.SEP_LLIST_NEIGHBLIST, 0, args.skin)
        ret = capi.SepRet(); alpha = C.c_double(0.1)
        step_api = mol_api_step(lib, capi, name, P, atoms, hsys, ret, alpha)
        for _ in range(5):
            step_api()
        lib.sep_gpu_sync(atoms)
        view["x"][:] = w["x"]; view["v"][:] = w["v"]; view["xn"][:] = 0.0
        view["cross_neighb"][:] = 0; view["crossings"][:] = 0
        lib.sep_gpu_invalidate(atoms)
        hsys.neighb_flag = 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(Ke):
            step_api()
        lib.sep_gpu_sync(atoms)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        e_epot = ret.epot / n
        lib.sep_close(atoms, n)
        h2d = n * (24 * 3 + 8 * 2 + 1 + 4 + 12 * 2)
        d2h_final = n * (24 * 4 + 12 * 2 + 24)
        e2e = {"value": n * Ke / el, "unit": UNIT, "h2d_bytes_per_step": h2d / Ke, "d2h_bytes_per_step": 416 + d2h_final / Ke,
               "steps": Ke, "epot_per_atom": e_epot,
               "path": "sep_* API (include/sep.h), default coherence mode SEP_SYNC=auto: scalars D2H after every hot call; seppart[] "
                       "uploaded at step 0 and downloaded after the last step, both inside the timed region"}

    # roofline of the dominant pair kernel (k_lj_list for butane, k_coulomb_list for water): list + own row + force row
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = "coulomb" if name == "water" else "force"
    d_ms, d_cnt = kt[dom]
    dom_ms = d_ms / max(d_cnt, 1)
    alg_bytes = 4.0 * full_entries / 2.0 + 64.0 + (8.0 if name == "water" else 0.0)     # half-list entries, SURVEY 8d convention
    achieved = alg_bytes * n / (dom_ms * 1e-3) / 1e9
    traffic = None                                   # DRAM bytes per launch of that kernel from the committed ncu capture
    if name == "water":
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "water_coulomb_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:      # noqa: BLE001
            traffic = None
    roofline = {"kernel": "k_coulomb_list2" if name == "water" else "k_lj_list", "bound": "hbm", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_atom_step": alg_bytes, "avg_launch_ms": dom_ms, "launches": d_cnt,
                "share_of_step": d_ms / (t_sec * 1e3)}
    cpu = None
    if not args.no_cpu:
        try:
            info = cpu_arm_child("mol", name=name, skin=args.skin, target_seconds=args.cpu_seconds, threads=host_cores())
            cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:      # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    per_step = 7 + (2 * 3 if name == "butane" else 2 * 3 + 1)
    label = ("C2 butane: sep_force_pairs(CC, lj_shift, EXCL_SAME_MOL) + stretch + angle + Ryckaert torsion + NH + leapfrog"
             if name == "butane" else
             "C3 water: sep_force_pairs(OO) + stretch + cos^2 angle + sep_coulomb_sf(2.9) + NH + leapfrog")
    emit({
        "metric": METRIC.replace("LJ, rc=2.5", label.split(":")[0]), "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": t_sec * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "natoms_total": n, "molecules": w["nmol"], "tiling": f"{reps}^3 unit cells",
                   "box": list(map(float, w["L"])), "cells": list(gsys.nsubbox[:]), "skin": args.skin, "dt": P["dt"],
                   "parallelism": "single GPU", "l2": "inputs larger than L2 (Verlet list %.0f MB)" % (n * full_entries * 4 / 1e6),
                   "list_rebuilds_in_timed_region": nbuild, "half_pairs_per_atom": full_entries / 2.0,
                   "epot_per_atom": epotN, "ekin_per_atom": ekinN, "sepgpu_opts": os.environ.get("SEPGPU_OPTS", "")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": K * per_step + nbuild * 11, "clocks": clocks,
        "kernel_ms": {k: {"total_ms": a, "launches": b} for k, (a, b) in kt.items()},
    })
    return 0


# ---------------------------------------------------------------------------------------------------
def sampled_run_record(ncell=46, steps=300, timeout_s=90):
    """prg1's kind of run -- the NVT loop WITH run-time samplers (vacf, sacf, msd every step, profs, gh) -- through the sep_*
    API, once with the samplers reading atoms[] on the host (what an unchanged program gets: a download per sample) and once
    fed from the device (SEP_SAMPLER_FEEDS=1, sepgpu_feeds.cu).  Each run is its own process (tests/feeds_driver.py) under a
    timeout; whatever goes wrong ends up in the record, never in the bench line's own numbers."""
    import tempfile
    rec = {"workload": "two-species LJ NVT + samplers vacf sacf msd profs gh through sep_* (SEP_SYNC=auto)", "steps": steps}
    try:
        with tempfile.TemporaryDirectory() as td:
            outs = {}
            for name, flag in (("host_samplers", "0"), ("device_feeds", "1")):
                env = dict(os.environ, SEP_SAMPLER_FEEDS=flag, FEEDS_NCELL=str(ncell), FEEDS_SAMPLERS="vacf,sacf,msd,profs,gh")
                env.pop("SEP_SYNC", None)
                out = os.path.join(td, name)
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "feeds_driver.py"), out, str(steps)],
                                   capture_output=True, text=True, timeout=timeout_s, env=env)
                if r.returncode != 0:
                    rec[name] = {"error": (r.stderr or r.stdout)[-300:]}
                    continue
                w = r.stdout.split()
                get = lambda k: w[w.index(k) + 1]            # noqa: E731
                n, secs = int(get("natoms")), float(get("loop_s"))
                rec["natoms"] = n
                rec[name] = {"value": n * steps / secs, "unit": UNIT, "ms_per_step": secs * 1e3 / steps,
                             "atoms_downloads": int(get("get_calls")), "feed_calls": int(get("feed_calls"))}
                outs[name] = out
            if len(outs) == 2:
                worst = 0.0
                files = sorted(f for f in os.listdir(outs["host_samplers"]) if f.endswith(".dat"))
                for f in files:
                    a = np.loadtxt(os.path.join(outs["host_samplers"], f), ndmin=2)
                    b = np.loadtxt(os.path.join(outs["device_feeds"], f), ndmin=2)
                    if a.shape != b.shape:
                        worst = float("inf")
                        break
                    if a.size:
                        worst = max(worst, float(np.nanmax(np.abs(a - b))))
                rec["files_compared"] = len(files)
                rec["max_abs_difference_between_the_files"] = worst
    except Exception as e:      # noqa: BLE001
        rec["error"] = repr(e)
    return rec


def run_bounded(cmd, env, timeout_s, cwd=None):
    """a child job in a session of its own: (returncode or None when it had to be killed, merged output).  Everything the
    job started is killed with it -- its process group, and any process carrying the job's marker in its environment."""
    import signal
    import uuid
    marker = uuid.uuid4().hex
    env = dict(env, SEPB_NESTED_RUN=marker)
    p = subprocess.Popen(cmd, env=env, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        out, _ = p.communicate(timeout=timeout_s)
        return p.returncode, out
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except OSError:
            pass
        for pid in os.listdir("/proc"):             # workers a launcher may have put into sessions of their own
            if pid.isdigit():
                try:
                    if ("SEPB_NESTED_RUN=" + marker).encode() in open(f"/proc/{pid}/environ", "rb").read():
                        os.kill(int(pid), signal.SIGKILL)
                except OSError:
                    pass
        try:
            p.communicate(timeout=10)
        except Exception:      # noqa: BLE001
            pass
        return None, f"no answer within {timeout_s} s"


def env_without_launcher():
    """this process's environment without what torchrun gave it (a nested job gets ranks and a rendezvous of its own)"""
    drop = {"RANK", "LOCAL_RANK", "GROUP_RANK", "ROLE_RANK", "ROLE_NAME", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_WORLD_SIZE",
            "ROLE_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"}
    return {k: v for k, v in os.environ.items() if k not in drop and not k.startswith("TORCHELASTIC") and not k.startswith("TORCH_NCCL")}


def nested_dd_check(mol="water", steps=40, timeout_s=100):
    """A two-rank decomposed MOLECULAR run checked against a single-GPU run (tests/dd_check.py DD_MOL=..., tests/dd_mol.py),
    launched as a torchrun job of its own after this bench's process group is gone: the scaling record then carries a
    decomposed-vs-single comparison of sep_coulomb_sf + typed LJ + bonds + angles on real GPUs.  Own session, own timeout,
    own rendezvous port; its outcome is a record, never an error of the bench."""
    import socket
    rec = {"check": f"two-rank decomposed {mol} ({steps} steps) against the same calls on one GPU: first-step forces 1e-10, sums, final positions"}
    t0 = time.perf_counter()
    try:
        env = env_without_launcher()
        env.update(DD_MOL=mol, DD_STEPS=str(steps))
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        rc_, out = run_bounded([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                                "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dd_check.py")], env, timeout_s)
        rec["ok"] = rc_ == 0 and "-> OK" in out
        lines = out.splitlines()
        keep = [ln for ln in lines if "dd_mol" in ln and "-> " in ln] or \
               [ln for ln in lines if "Error" in ln and "ChildFailedError" not in ln]
        rec["tail"] = (keep[-1] if keep else out[-300:])[-400:]
    except Exception as e:      # noqa: BLE001
        rec["ok"] = False
        rec["tail"] = repr(e)
    rec["seconds"] = time.perf_counter() - t0
    return rec


def spec_force2_trial(timeout_s=110):
    """The launch sent ahead in a DECOMPOSED run (SEPGPU_OPTS=spec_force=2, experimental: on two B200s it ended in a halo wait
    that never returned, docs/ROUND_NOTES.md; since then the grid sent ahead is ordered behind the rank's own halo push) tried
    as a two-rank bench job of its own -- same 1 M atoms per rank, its own decomposed-vs-single check first.  Whatever
    happens is a record: the device-side waits are bounded (about 60 s), the job is bounded by timeout_s."""
    import socket
    rec = {"what": "bench.py --gpus 2 --steps 200 --warmup 50 with SEPGPU_OPTS=spec_force=2 (force launch sent ahead in decomposed runs, experimental)"}
    t0 = time.perf_counter()
    try:
        env = env_without_launcher()
        env["SEPGPU_OPTS"] = "spec_force=2"
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        rc_, out = run_bounded([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                                "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--gpus", "2", "--steps", "200",
                                "--warmup", "50", "--no-cpu", "--no-e2e", "--no-other"], env, timeout_s)
        got = None
        for ln in out.splitlines():
            if ln.startswith("{") and '"metric"' in ln:
                got = json.loads(ln)
        if rc_ == 0 and got:
            rec.update(ok=True, value=got["value"], unit=got["unit"], ms_per_step=got["ms_per_step"], steps=got["steps"],
                       kernel_ms=got.get("kernel_ms"), dd_check=got.get("dd_check", {}).get("ok") if isinstance(got.get("dd_check"), dict) else None)
        else:
            keep = [ln for ln in out.splitlines() if "Error" in ln and "ChildFailedError" not in ln]
            rec.update(ok=False, tail=(keep[-1] if keep else out[-300:])[-400:])
    except Exception as e:      # noqa: BLE001
        rec.update(ok=False, tail=repr(e))
    rec["seconds"] = time.perf_counter() - t0
    return rec


def sep_ngpu_e2e_record(ngpu, nside, steps=1000, warm=300, timeout_s=100):
    """End to end through the sep_* API on `ngpu` GPUs: tests/progs/nvt_time.c -- the prg1 loop written against include/sep.h,
    compiled here with gcc and linked with libsep.so like any seplib program -- run with SEP_NGPU=ngpu (the library forks one
    copy per GPU at the first hot call, seplib_b200/csrc/host/sep_dd.c).  A process of its own under a timeout; its outcome
    is a record, never an error of the bench."""
    import tempfile
    rec = {"path": f"sep_* API (include/sep.h), unchanged C program, SEP_NGPU={ngpu}, SEP_SYNC=auto: sepret refreshed after every hot call",
           "unit": UNIT}
    try:
        with tempfile.TemporaryDirectory() as td:
            exe = os.path.join(td, "nvt_time")
            subprocess.check_call(["gcc", "-std=c99", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "progs", "nvt_time.c"),
                                   "-L" + os.path.join(ROOT, "seplib_b200"), "-lsep", "-lm", "-o", exe])
            env = env_without_launcher()
            libdir = os.path.join(ROOT, "seplib_b200")
            if env.get("SEPGPU_EMU_LIB"):             # plumbing test on the CPU kernel emulator (tests/emu): libsep.so -> libsep_emu.so
                libdir = os.path.join(os.path.dirname(env["SEPGPU_EMU_LIB"]), "emu_lib")
            env["LD_LIBRARY_PATH"] = libdir + ":" + env.get("LD_LIBRARY_PATH", "")
            env.pop("SEP_SYNC", None)
            env["SEP_NGPU"] = str(ngpu)
            rc_, out = run_bounded([exe, str(nside), str(steps), str(warm)], env, timeout_s, cwd=td)
        w = out.split()
        if rc_ != 0 or "seconds_loop" not in w:
            rec["error"] = ("rc %r: " % rc_) + out[-300:]
            return rec
        get = lambda k: float(w[w.index(k) + 1])             # noqa: E731
        n, k = int(get("natoms")), int(get("steps"))
        rec.update(value=n * k / get("seconds_loop"), natoms=n, steps=k, ms_per_step=1e3 * get("seconds_loop") / k,
                   seconds_first_calls=get("seconds_warm"), warm_steps=int(get("warm")), seconds_download=get("seconds_download"),
                   epot_per_atom=get("epot_per_atom"), ekin_per_atom=get("ekin_per_atom"), list_rebuilds=int(get("rebuilds")))
    except Exception as e:      # noqa: BLE001
        rec["error"] = repr(e)
    return rec


def weak_lattice_dims(ncell, world, box="cubic"):
    """Lattice sides for `world` GPUs at ncell^3 atoms per GPU (weak scaling, near-cubic box, z longest):
    1 -> n,n,n ; 2 -> n,n,2n ; 4 -> n,2n,2n ; 8 -> 2n,2n,2n.  box="stacked": n,n,world*n (the cross-section a slab
    decomposition likes best; not the default -- a bulk fluid is simulated in a cubic box)."""
    if box == "stacked":
        return [ncell, ncell, ncell * world]
    dims = [ncell, ncell, ncell]
    k, w = 2, world
    while w > 1:
        if w % 2:
            raise ValueError("GPU count must be a power of two")
        dims[k] *= 2
        k = (k - 1) % 3
        w //= 2
    return dims


def lattice_box(dims, rho):
    a = (1.0 / rho) ** (1.0 / 3.0)
    return [d * a for d in dims], a


def slab_atoms(dims, a, zlo, zhi, lsz):
    """Lattice atoms (global id, position) whose cell layer (int)(z/lsz) lies in [zlo,zhi)."""
    nx, ny, nz = dims
    gz = (np.arange(nz) + 0.5) * a
    layer = np.floor(gz / lsz).astype(np.int64)
    kz = np.nonzero((layer >= zlo) & (layer < zhi))[0]
    gx = (np.arange(nx) + 0.5) * a
    gy = (np.arange(ny) + 0.5) * a
    z, y, x = np.meshgrid(gz[kz], gy, gx, indexing="ij")
    pos = np.ascontiguousarray(np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1))
    iz, iy, ix = np.meshgrid(kz, np.arange(ny), np.arange(nx), indexing="ij")
    gid = (iz.ravel() * ny + iy.ravel()) * nx + ix.ravel()
    return pos, gid.astype(np.int32)


def slab_velocities(gid, n_total, temp, seed):
    """Velocities drawn per global id (same on every decomposition), drift removed, rescaled to temp."""
    rng = np.random.default_rng(seed)
    v = rng.random((n_total, 3)) - 0.5
    v -= v.mean(axis=0)
    v *= np.sqrt(3 * n_total * temp / (v * v).sum())
    return np.ascontiguousarray(v[gid])


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ncell", type=int, default=100, help="lattice side: ncell^3 atoms per GPU")
    ap.add_argument("--rho", type=float, default=0.8)
    ap.add_argument("--skin", type=float, default=0.25)
    ap.add_argument("--tpa", type=int, default=0)
    ap.add_argument("--force-grid", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=-1, help="decomposed runs: 0 = halo exchange before the force pass, 1 = beside it (default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-ncell", type=int, default=100, help="CPU arms: lattice side (100 = the 1 M-atom configuration itself)")
    ap.add_argument("--cpu-arm-json", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-matrix", action="store_true", help="skip the section-8d matrix of smaller CPU samples")
    ap.add_argument("--no-other", action="store_true", help="skip the side records: other_workloads (C2 butane / C3 water), sampled_run, and at N = 2 the nested jobs")
    ap.add_argument("--equilibrate", type=int, default=300, help="untimed steps from the lattice before the warm-up (thermalisation)")
    ap.add_argument("--e2e-steps", type=int, default=1000, help="the e2e arm runs max(--steps, this) steps (it carries one upload and one download of atoms[])")
    ap.add_argument("--box", default="cubic", choices=["cubic", "stacked"], help="N>1: near-cubic box (default) or N cubes stacked along the slab axis")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas instead of domain decomposition")
    ap.add_argument("--workload", default="lj", choices=["lj", "butane", "water"],
                    help="lj = C1/C4 (the BASELINE metric, default); butane = C2; water = C3 (single GPU)")
    ap.add_argument("--reps", type=int, default=0, help="molecular workloads: unit cells per side (default 6 / 12)")
    args = ap.parse_args()
    if args.cpu_arm_json:
        return cpu_arm_main(args.cpu_arm_json)

    # stdout carries exactly ONE line (the JSON): libraries that print banners to fd 1 (NCCL version line,
    # the reference's sep-warning) are sent to stderr for the duration of the run.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rc, dt, temp, tau = 2.5, 0.005, 1.0, 0.01       # prg1's literals with the metric's rc (prgs/prg1.c:24-30)
    K, W = args.steps, max(args.warmup, 3)

    if args.workload != "lj":
        if rank != 0 and not (args.workload == "butane" and args.gpus > 1 and args.impl != "reference"):
            return 0
        if args.impl == "reference":
            info = cpu_arm_child("mol", name=args.workload, skin=args.skin, target_seconds=max(args.cpu_seconds, 5.0) * 2, threads=host_cores())
            value = info["value"]
            emit({"impl": "reference", "metric": METRIC.replace("LJ, rc=2.5", args.workload), "value": value, "unit": UNIT,
                  "n_gpus": args.gpus, "steps": info["steps"], "warmup": 5, "ms_per_step": 1e3 * info["seconds"] / info["steps"],
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": args.workload + " bounded CPU sample", "natoms": info["natoms"], "l2": "n/a (host)"},
                  "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
                  "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
            return 0
        return run_molecular(args, emit, local_rank)

    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = host_cores()
        info = cpu_arm_child("lj", ncell=args.cpu_ncell, rho=args.rho, rc=rc, skin=args.skin, dt=dt, temp=temp, tau=tau,
                             target_seconds=max(args.cpu_seconds, 5.0) * 2, threads=threads)
        value = info["value"]
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": info["steps"], "warmup": 5, "ms_per_step": 1e3 * info["seconds"] / info["steps"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "prg1-style LJ NVT: sep_force_pairs(sep_lj_shift, rc=2.5) + sep_nosehoover + sep_leapfrog"
                                       + (" (C1, 1M atoms)" if info["natoms"] == 1000000 else " bounded CPU sample"),
                           "natoms": info["natoms"], "natoms_total": info["natoms"], "rho": args.rho, "skin": args.skin,
                           "dt": dt, "T0": temp, "tau": tau, "lattice": "sc %dx%dx%d" % ((args.cpu_ncell,) * 3),
                           "parallelism": "%d OpenMP threads (reference's own sep_set_omp path)" % threads, "l2": "n/a (host)"},
                "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from seplib_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load()
    if lib.sepgpu_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device -- seplib-b200 has no CPU path")
    decomposed = world > 1 and not args.replicas

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- workload ---------------------------------------------------------------------
    if decomposed:
        dims = weak_lattice_dims(args.ncell, world, args.box)
    else:
        dims = [args.ncell] * 3
    Lvec, a_lat = lattice_box(dims, args.rho)
    n_total = dims[0] * dims[1] * dims[2]
    gsys = capi.make_sys(Lvec, rc, dt, skin=args.skin)
    nzg = gsys.nsubbox[2]
    if decomposed:
        z0, z1 = capi.dd_slab_range(rank, world, nzg)
        x, gid = slab_atoms(dims, a_lat, z0, z1, gsys.lsubbox[2])
        v = slab_velocities(gid, n_total, temp, 1000)
        n = len(x)
        ncap = int(1.25 * n_total / world) + int(3.5 * n_total / nzg) + 4096
    else:
        x, gid = slab_atoms(dims, a_lat, 0, nzg + 1, gsys.lsubbox[2])
        v = slab_velocities(gid, n_total, temp, 1000 + rank)
        n = len(x)
        ncap = n
    ljp = capi.lj_param(rc, kind="lj_shift")
    gs_ref, lj_ref = C.byref(gsys), C.byref(ljp)

    def new_system():
        s_ = capi.System(ncap, device=local_rank)
        if decomposed:
            # one NCCL communicator per context: a fresh unique id from rank 0 each time
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.dd_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            s_.dd_init(rank, world, bytes(idt.cpu().numpy().tobytes()), gsys, n_total)
            s_.dd_set_owned(n)
        for opt, val in (("tpa", args.tpa), ("force_grid", args.force_grid)):
            if val:
                s_.call("sepgpu_set_option", opt.encode(), val)
        if args.overlap >= 0:
            s_.call("sepgpu_set_option", b"overlap", args.overlap)
        return s_

    def upload(s_):
        s_.put(capi.F_X, x)
        s_.put(capi.F_V, v)
        if decomposed:
            s_.put(capi.F_GID, gid)
        s_.call("sepgpu_set_alpha", 0, 0.1)

    def make_step(s_):
        ctx, fn = s_.ctx, lib

        def step():
            fn.sepgpu_reset_ret(ctx)
            fn.sepgpu_reset_force(ctx)
            r = fn.sepgpu_force_lj(ctx, gs_ref, b"AA", lj_ref, 1, 1)      # rebuilds the list itself when the trigger fired
            r |= fn.sepgpu_nosehoover(ctx, gs_ref, temp, 0, tau)
            r |= fn.sepgpu_leapfrog(ctx, gs_ref)
            if r:
                raise RuntimeError("device step failed: " + fn.sepgpu_last_error().decode())
        return step

    # ---------------- device-resident arm -------------------------------------------------------
    dd_check = None
    if decomposed:
        # decomposed vs single-GPU run of a small system (pair sets, sums, final positions), before anything is timed
        import dd_check as ddc
        dd_check = ddc.run(rank, world, local_rank, 40, max(28, 6 * world), verbose=(rank == 0))      # >= 2 cell layers per rank
        if not dd_check["ok"]:
            raise RuntimeError("bench.py: the decomposed run does not reproduce the single-GPU run: %r" % (dd_check,))
    s = new_system()
    upload(s)
    step_dev = make_step(s)
    for _ in range(max(0, args.equilibrate - W)):      # leave the lattice behind before anything counts
        step_dev()
    for _ in range(W):
        step_dev()
    nb0 = s.scalars().nbuild
    s.call("sepgpu_set_option", b"time_kernels", 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    t_wall0 = time.time()
    s.call("sepgpu_timer_start")
    # the K timed steps are driven from C (sepgpu_md_lj_nvt: the same five calls per step as step_dev), as a seplib program
    # would drive them -- no interpreter, and no clock-sampler thread contending for it, between the calls
    if lib.sepgpu_md_lj_nvt(s.ctx, gs_ref, b"AA", lj_ref, 1, temp, 0, tau, K):
        raise RuntimeError("device step failed: " + lib.sepgpu_last_error().decode())
    ms = C.c_float()
    s.call("sepgpu_timer_stop", C.byref(ms))
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    sc = s.scalars()
    nbuild = sc.nbuild - nb0
    kt = {}
    for which in ("force", "build", "intgr") + (("halo", "migrate") if decomposed else ()):
        tot, cnt = C.c_float(), C.c_int()
        s.call("sepgpu_kernel_time", which.encode(), C.byref(tot), C.byref(cnt))
        kt[which] = (tot.value, cnt.value)
    t_ms = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_sec = float(t_ms.item()) * 1e-3
    total_atoms = n_total if decomposed else n * world
    value = total_atoms * K / t_sec
    norm = n_total if decomposed else n
    epotN, ekinN = sc.epot / norm, sc.ekin / norm
    pairs_per_atom = sc.npairs_listed / max(s.dd_layers()[2] if decomposed else n, 1) / 2.0
    own_halo = s.dd_layers()[2:] if decomposed else (n, 0)
    s.close()

    # ---------------- e2e arm: host buffers in, host buffers out ------------------------------------
    e2e = None
    if not args.no_e2e:
        # at least --e2e-steps steps: the array is uploaded once and downloaded once inside this timed region, and a run of
        # a few dozen steps would time those two transfers (16 ms each at 1 M atoms), not the loop a program spends its time in
        Ke = max(10, K, args.e2e_steps)
        if not decomposed:
            # the reference-facing sep_* API of include/sep.h on a host seppart[] array, in the library's DEFAULT coherence
            # mode (SEP_SYNC=auto: what an unchanged program gets) -- and, for comparison, in lazy and step mode
            fun = None

            def api_run(mode, nsteps, label):
                lib.sep_gpu_set_sync(mode)
                atoms = lib.sep_init(n, 0)
                view = capi.atoms_view(atoms, n)
                view["x"][:] = x
                view["v"][:] = v
                hsys = lib.sep_sys_setup(Lvec[0], Lvec[1], Lvec[2], rc, dt, n, capi.SEP_LLIST_NEIGHBLIST)
                lib.sep_set_skin(C.byref(hsys), args.skin)
                ret = capi.SepRet()
                alpha = C.c_double(0.1)
                fun = C.cast(lib.sep_lj_shift, C.c_void_p)

                def step_api():
                    lib.sep_reset_retval(C.byref(ret))
                    lib.sep_reset_force(atoms, C.byref(hsys))
                    lib.sep_force_pairs(atoms, b"AA", rc, fun, C.byref(hsys), C.byref(ret), capi.SEP_ALL)
                    lib.sep_nosehoover(atoms, temp, C.byref(alpha), tau, C.byref(hsys))
                    lib.sep_leapfrog(atoms, C.byref(hsys), C.byref(ret))

                for _ in range(max(3, min(W, 20))):      # warm-up (both directions), then restore the host arrays: they are
                    step_api()                           # uploaded again inside the timed region
                lib.sep_gpu_sync(atoms)
                view["x"][:] = x
                view["v"][:] = v
                view["xn"][:] = 0.0
                view["cross_neighb"][:] = 0
                view["crossings"][:] = 0
                lib.sep_gpu_invalidate(atoms)
                hsys.neighb_flag = 1
                barrier()
                t0 = time.perf_counter()
                step_api()                       # first call uploads x,v,m,z,type,... from the host array
                t_up = time.perf_counter()
                for _ in range(nsteps - 1):
                    step_api()
                t_st = time.perf_counter()
                lib.sep_gpu_sync(atoms)          # final state back into atoms[]
                torch.cuda.synchronize()
                el_ = time.perf_counter() - t0
                log("e2e %s: first step incl. upload %.1f ms, %d steps %.1f ms, download %.1f ms"
                    % (label, 1e3 * (t_up - t0), nsteps - 1, 1e3 * (t_st - t_up), 1e3 * (time.perf_counter() - t_st)))
                ep = ret.epot / n
                lib.sep_close(atoms, n)
                return el_, ep

            h2d = h2d_bytes_lj(n)
            d2h_final = n * (24 * 4 + 12 * 2 + 24)                 # x, v, f, a, counters, xn
            el, e_epot = api_run(3, Ke, "SEP_SYNC=auto")
            e2e_modes = {}
            for mode, label, ks in ((0, "lazy", Ke), (1, "step", max(5, min(Ke, 40)))):
                try:
                    el_m, _ = api_run(mode, ks, "SEP_SYNC=" + label)
                    e2e_modes[label] = {"value": n * ks / el_m, "unit": UNIT, "steps": ks, "ms_per_step": 1e3 * el_m / ks,
                                        "d2h_bytes_per_step": 416 + (d2h_final if mode == 1 else d2h_final / ks)}
                except Exception as e:      # noqa: BLE001
                    e2e_modes[label] = {"value": None, "error": repr(e)}
            lib.sep_gpu_set_sync(3)
            e2e_step_mode = e2e_modes
            how = ("sep_* API (include/sep.h) in the default coherence mode SEP_SYNC=auto (atoms[] page-protected while the device "
                   "copy is newer): sepret/sepsys scalars D2H after every hot call; seppart[] uploaded at step 0 and downloaded "
                   "after the last step, both inside the timed region")
        else:
            # decomposed: the sepgpu_* C ABI with host numpy buffers (the sep_* API is one process / one GPU)
            # (the context is reused after the warm-up: a fresh NCCL communicator would put its lazy connection
            #  set-up, hundreds of ms, inside the timed region)
            s2 = new_system()
            step2 = make_step(s2)
            upload(s2)
            for _ in range(max(3, min(W, 20))):
                step2()
            zeros3 = np.zeros((ncap, 3), dtype=np.int32)
            barrier()
            t0 = time.perf_counter()
            s2.n = ncap                      # back to the initial ownership, then H2D from the host arrays
            s2.dd_set_owned(n)
            upload(s2)
            s2.put(capi.F_CROSSINGS, zeros3[:n]); s2.put(capi.F_CROSS_NEIGHB, zeros3[:n])
            for _ in range(Ke):
                step2()
                sce = s2.scalars()           # D2H of the step's scalars (collective)
            s2.dd_layers()
            xo, vo, fo, go = s2.get(capi.F_X), s2.get(capi.F_V), s2.get(capi.F_F), s2.get(capi.F_GID)   # D2H
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            e_epot = sce.epot / n_total
            s2.close()
            how = ("sepgpu_* C ABI with host buffers (decomposed run): x,v,gid H2D at step 0, scalar block D2H every step, "
                   "x,v,f,gid D2H after the last step, all inside the timed region")
            h2d = n * (24 * 2 + 4 + 12 * 2)
            d2h_final = n * (24 * 3 + 4)
        t_e = torch.tensor([el], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        el = float(t_e.item())
        scal_bytes = 416
        e2e = {"value": total_atoms * Ke / el, "unit": UNIT,
               "h2d_bytes_per_step": h2d / Ke, "d2h_bytes_per_step": scal_bytes + d2h_final / Ke,
               "steps": Ke, "path": how, "epot_per_atom": e_epot}
        if not decomposed:
            e2e["other_sync_modes"] = e2e_step_mode

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the pair-force kernel ----------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_list, n_in, alg_flops, alg_bytes = algorithmic_per_atom(args.rho, rc, args.skin)
    f_ms, f_cnt = kt["force"]
    force_ms = f_ms / max(f_cnt, 1)
    n_kernel = own_halo[0]                      # atoms whose rows one launch on this rank processes
    achieved_gbs = alg_bytes * n_kernel / (force_ms * 1e-3) / 1e9
    fp64_peak = C.c_double()
    lib.sepgpu_peak_fp64(local_rank, C.byref(fp64_peak))
    achieved_tf = alg_flops * n_kernel / (force_ms * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:      # noqa: BLE001
            traffic = None
    roofline = {"kernel": "k_lj_tile", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_atom_step": alg_bytes, "avg_launch_ms": force_ms, "launches": f_cnt,
                "share_of_step": f_ms / (t_sec * 1e3),
                "fp64": {"note": "SURVEY 8d names the FP64 pipe as this kernel's bound; the peak is an FMA-chain probe run on this GPU in this process "
                                 "(sepgpu_peak_fp64, 2 flop per DFMA).  ncu (profiles/r02_k_lj_tile_ncu.txt): FP64 pipe 46 %, shared-memory data pipe 73 %, "
                                 "issue slots 55 % -- the kernel is latency / phase bound (staging + end-of-tile barrier), see DESIGN.md section 3",
                         "algorithmic_flops_per_atom_step": alg_flops, "achieved_tflops": achieved_tf,
                         "peak_tflops_fma_chain_measured_here": fp64_peak.value,
                         "frac": achieved_tf / fp64_peak.value if fp64_peak.value else None}}

    cpu = None
    if not args.no_cpu and world == 1:
        try:
            info = cpu_arm_child("lj", ncell=args.cpu_ncell, rho=args.rho, rc=rc, skin=args.skin, dt=dt, temp=temp, tau=tau,
                                 target_seconds=args.cpu_seconds, threads=host_cores())
            cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:      # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        if not args.no_cpu_matrix:
            cpu["matrix"] = cpu_matrix(args.rho, rc, dt, temp, tau, 2.5, skip_big=False)
            cpu["reference_cuda"] = reference_cuda_baseline()
    other = sampled = None
    if world == 1 and not args.no_other:
        # short runs of the C2 / C3 configurations (device loop only), so that the driver's record carries them too
        import copy
        other = []
        for wname, ksteps in (("butane", 200), ("water", 100)):
            a2 = copy.copy(args)
            a2.workload, a2.steps, a2.warmup, a2.no_e2e, a2.no_cpu, a2.reps = wname, ksteps, 40, True, True, 0
            got = []
            try:
                run_molecular(a2, got.append, local_rank)
                g = got[0]
                other.append({"workload": g["config"]["workload"], "natoms": g["config"]["natoms_total"], "value": g["value"],
                              "unit": UNIT, "ms_per_step": g["ms_per_step"], "steps": ksteps, "warmup": 40,
                              "list_rebuilds": g["config"]["list_rebuilds_in_timed_region"],
                              "epot_per_atom": g["config"]["epot_per_atom"], "roofline": g["roofline"],
                              "kernel_ms": g["kernel_ms"]})
            except Exception as e:      # noqa: BLE001
                other.append({"workload": wname, "value": None, "error": repr(e)})
        sampled = sampled_run_record()

    # kernels of this rank in the timed region: force, finalize, nh_update, integrate, finalize (+ lazy resets);
    # decomposed: + halo push and wait/unpack (peer-memory path); per rebuild: set_xn + 10 build kernels (+ ~25 migration/halo)
    per_step = 3 + (2 if decomposed else 0)      # force, integrate, folded finaliser (+ halo push / unpack)
    per_build = 11 + (25 if decomposed else 0)
    launches = (K * per_step + nbuild * per_build) * (world if decomposed else 1)
    if decomposed:
        par = f"{world}-way slab domain decomposition along z (NCCL P2P halo + migration, all-reduce per step)"
    else:
        par = "single GPU" if world == 1 else f"{world} independent replicas (one per GPU)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_sec * 1e3 / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "prg1-style LJ NVT: sep_force_pairs(sep_lj_shift, rc=2.5) + sep_nosehoover + sep_leapfrog"
                               + (" (C1, 1M atoms)" if world == 1 else f" ({total_atoms} atoms; C4 at 8 GPUs)"),
                   "natoms_per_gpu": total_atoms // world, "natoms_total": total_atoms, "rho": args.rho, "skin": args.skin,
                   "dt": dt, "T0": temp, "tau": tau, "lattice": "sc %dx%dx%d" % tuple(dims), "parallelism": par,
                   "rank0_owned_halo_atoms": list(own_halo),
                   "l2": "inputs larger than L2 (xs 32 MB + Verlet list %.0f MB + state > 126 MB per GPU)" % (own_halo[0] * pairs_per_atom * 8 / 1e6),
                   "list_rebuilds_in_timed_region": nbuild, "half_pairs_per_atom": pairs_per_atom,
                   "epot_per_atom": epotN, "ekin_per_atom": ekinN, "sepgpu_opts": os.environ.get("SEPGPU_OPTS", "")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "kernel_ms": {k: {"total_ms": a, "launches": b} for k, (a, b) in kt.items()},
    }
    if other is not None:
        line["other_workloads"] = other
    if sampled is not None:
        line["sampled_run"] = sampled
    if dd_check is not None:
        line["dd_check"] = dd_check
    if world > 1:
        dist.destroy_process_group()
    if decomposed and world == 2 and not args.no_other:
        # the other rank has left (it returns right after the e2e arm): both GPUs are free for a two-rank job of its own
        log("bench line complete (a copy, should the nested jobs be cut short): " + json.dumps(line))
        line["dd_water_check"] = nested_dd_check("water")
        line["e2e_sep_ngpu"] = sep_ngpu_e2e_record(2, 126)        # 2.0 M atoms, as the decomposed run above
        if not os.environ.get("SEPGPU_OPTS"):
            line["spec_force2_trial"] = spec_force2_trial()
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
