#!/usr/bin/env python
"""bench.py - AA-ADMM hot path on B200: FP64 ADMM+Anderson iterations/sec at 1M tets.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Workload (BASELINE.json configs[3], SURVEY 8d cfg 4): ONE linear-elastic beam of 148x37x37 cubes
(1,013,060 tets, 215,156 nodes, 2,888 pinned), admm_anderson_hard_zxu ordering, Anderson m=5,
dt=1/30, g=-9.8, rho=1, 100 ADMM iterations per frame, pins stretched every frame.
A bench "step" is one frame = one admm::Solver::step() (up to 100 ADMM iterations).
  value  = ADMM iterations / second of the device loop, inputs resident in HBM (CUDA events
           between the end of the uploads and the start of the downloads of every timed frame)
  e2e    = the same through admm::Solver::step() with HOST buffers: per-frame H2D of the predicted
           positions and pins, D2H of the positions and the residual history, host-side explicit
           step, wall clock
With N > 1 (torchrun) every rank runs one independent scene of the same size on its own GPU
(ensemble sharding, no data-path collective; results gathered once with NCCL) -> "weak" scaling.
"""
import argparse
import json

import numpy as np
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(cx=148, cy=37, cz=37, anderson_m=5, admm_iters=100, dt=1.0 / 30.0, penalty=1.0,
                youngs=1e7, poisson=0.399)
CPU_SAMPLE = dict(cx=12, cy=37, cz=37, iters=10)   # cpu_baseline leg beside the GPU line: about 20 s in all
# --impl reference: the largest beam of cfg 4's cross-section whose Eigen setup fits "a few minutes" for the whole run
REF_ARM = dict(cx=24, cy=37, cz=37, iters=10)
FULL_TETS = 148 * 37 * 37 * 5


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = False
        self.t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons, util = [], 0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
                util.append(float(r[6]))
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm),
                "gpu_util_pct_mean": (sum(util) / len(util)) if util else None}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class _NativeStdoutToStderr:
    """The reference prints progress lines to the C++ stdout from inside initialize() / step(); stdout of this program
    carries the ONE JSON line only, so file descriptor 1 points at stderr while the reference runs."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)  # the reference's buffered std::cout / printf output
        except Exception:
            pass
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def _reference_host_threads():
    """All host cores for the reference's OpenMP loops. torchrun exports OMP_NUM_THREADS=1 to its ranks; under torchrun
    rank 0 alone runs the CPU arm (the other ranks exit at once), so it takes the whole box. Must run before the first
    OpenMP library is loaded."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return cores


def _reference_solver(dims, iters, cores):
    """The unmodified reference (oracle/_ref) - or, where it could not be compiled, the C restatement - set up on a beam
    of `dims` cubes with the workload's material, time step, penalty and Anderson window. The scene comes from
    oracle/ref_scene.py (numpy): none of this repo's product libraries is loaded on this path."""
    from oracle import refbind
    from oracle.ref_scene import RefBeamScene
    kind = "reference" if refbind.have_ref() else "port"
    scene = RefBeamScene().add(dims[0], dims[1], dims[2], 0.0)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    r = refbind.RefSolver("hard") if kind == "reference" else refbind.PortSolver("hard")
    if kind == "reference":
        r._f("set_threads")(cores)
    r.add_tetmesh(verts, tets, masses, WORKLOAD["youngs"], WORKLOAD["poisson"], 0)
    dt = WORKLOAD["dt"]
    r.set_pins(pidx, scene.stretch(dt))
    t0 = time.perf_counter()
    r.initialize(dt, iters, -9.8, WORKLOAD["anderson_m"], True, WORKLOAD["penalty"])
    return r, scene, pidx, len(tets), kind, time.perf_counter() - t0


def run_reference(args):
    """The reference's own OpenMP CPU path (oracle/_ref = unmodified admm_anderson_hard_zxu) on the host cores.
    `value` is MEASURED: iterations/s of Solver::step on the largest beam of the workload's cross-section whose Eigen
    AMD + simplicial LDL^T setup fits the run's budget (REF_ARM); nothing is extrapolated into it."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    cores = _reference_host_threads()
    s = dict(REF_ARM)
    if args.ref_dims:
        s.update(cx=args.ref_dims[0], cy=args.ref_dims[1], cz=args.ref_dims[2])
    dt = WORKLOAD["dt"]
    with _NativeStdoutToStderr():
        r, scene, pidx, sample_tets, kind, setup_s = _reference_solver((s["cx"], s["cy"], s["cz"]), s["iters"], cores)
        if kind == "port":
            cores = 1  # the C restatement is a scalar port
        for _ in range(args.warmup):
            r.set_pins(pidx, scene.stretch(dt))
            r.step()
        iters, secs = 0, 0.0
        for _ in range(args.steps):
            r.set_pins(pidx, scene.stretch(dt))
            t0 = time.perf_counter()
            h = r.step()
            secs += time.perf_counter() - t0
            iters += len(h)
    its = iters / secs
    sample = (("unmodified reference admm_anderson_hard_zxu Solver::step, g++ -O2 -fopenmp, %d OpenMP threads" % cores
               if kind == "reference" else "oracle/port C restatement of hard_zxu Solver::step (oracle/_ref absent), gcc -O2, 1 thread") +
              "; beam %dx%dx%d = %d tets (cross-section of cfg 4, %.0f %% of its length), m=5, every step = one frame "
              "of %d ADMM iterations (cfg 4: 100); value = measured iterations/s AT THIS SIZE; Eigen AMD + LDL^T setup "
              "%.1f s excluded (at the full 1,013,060 tets that setup alone needs about an hour)") % (
                  s["cx"], s["cy"], s["cz"], sample_tets, 100.0 * s["cx"] / WORKLOAD["cx"], s["iters"], setup_s)
    line = {"impl": "reference", "metric": "admm_anderson_iterations_per_sec_1M_tets", "value": its,
            "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hard_zxu, linear beam %dx%dx%d (%d tets: bounded sample of cfg 4 = 148x37x37, "
                                   "1,013,060 tets), m=5, %d it/frame" % (s["cx"], s["cy"], s["cz"], sample_tets, s["iters"]),
                       "sample_tets": sample_tets, "full_tets": FULL_TETS, "same_config_as_gpu_arm": False,
                       "setup_s": round(setup_s, 1)},
            # the same figure scaled linearly by tets to cfg 4's size (favours the CPU: its triangular solves grow
            # faster than linearly); an extra key, not the value
            "value_scaled_to_full_tets": its * sample_tets / FULL_TETS,
            "cpu_baseline": {"value": its, "unit": "iterations/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": its, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "native_so_loaded": _repo_libraries_loaded()}
    _emit(line)
    return 0


def _repo_libraries_loaded():
    """Shared objects of this repository mapped into the process (the CPU arm must show oracle/ only)."""
    libs = set()
    try:
        with open("/proc/self/maps") as f:
            for ln in f:
                path = ln.split()[-1] if ln.strip() else ""
                if path.startswith(ROOT) and ".so" in os.path.basename(path):
                    libs.add(os.path.relpath(path, ROOT))
    except OSError:
        pass
    return sorted(libs)


def cpu_baseline_leg(sample=None, full_tets=None):
    """Bounded CPU sample run beside the GPU number (rank 0, N=1): about 20 s of the unmodified reference."""
    try:
        cores = _reference_host_threads()
        s = sample or CPU_SAMPLE
        FULL = full_tets or FULL_TETS
        dt = WORKLOAD["dt"]
        r, scene, pidx, sample_tets, kind, setup_s = _reference_solver((s["cx"], s["cy"], s["cz"]), s["iters"], cores)
        if kind == "port":
            cores = 1
        iters, secs = 0, 0.0
        for f in range(3):
            r.set_pins(pidx, scene.stretch(dt))
            t0 = time.perf_counter()
            h = r.step()
            if f > 0:
                secs += time.perf_counter() - t0
                iters += len(h)
        its = iters / secs
        return {"value": its * sample_tets / FULL, "unit": "iterations/s", "cores": cores, "kind": kind,
                "measured": {"iterations_per_s": its, "tets": sample_tets},
                "sample": ("unmodified reference" if kind == "reference" else "oracle/port C restatement of the reference") +
                          " hard_zxu Solver::step on beam %dx%dx%d (%d tets), 2 frames x %d iterations after 1 warm-up "
                          "frame: measured %.2f it/s at that size; `value` = that figure scaled linearly by tets to "
                          "%d (the workload's scene size; favours the CPU); Eigen setup %.1f s excluded. The measured "
                          "same-run reference arm at a larger size is `bench.py --impl reference`" % (
                              s["cx"], s["cy"], s["cz"], sample_tets, s["iters"], its, FULL, setup_s)}
    except Exception as e:  # the baseline is a report, never a reason to lose the GPU line
        return {"value": None, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": "failed: %r" % (e,)}


def run_gpu(args):
    rank, world, local = dist_env()
    if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        # torchrun exports OMP_NUM_THREADS=1; the ranks of one box share its host cores for the (untimed) setup
        # factorisation instead (read by the OpenMP runtime when the host library is loaded below)
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    import aa_admm_b200 as A
    if A.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    A.set_device(local if world > 1 else 0)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    w = dict(WORKLOAD)
    if args.small:
        w.update(cx=32, cy=8, cz=8)
    # Every rank runs the SAME cfg 4 scene: with the max-over-ranks clock, equal work per GPU is what makes
    # the N-GPU figure a weak-scaling number (a material sweep changes the iterations-to-tolerance per rank;
    # the sweep of SURVEY 8d cfg 5 is exercised by aa_admm_b200.ensemble and its tests).
    youngs, poisson = w["youngs"], w["poisson"]
    t0 = time.perf_counter()
    scene = A.BeamScene().add(w["cx"], w["cy"], w["cz"], 0.0)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    solver = A.Solver()
    solver.add_tetmesh(verts, tets, masses, youngs, poisson, 0)
    dt = w["dt"]
    solver.set_pins(pidx, scene.stretch(dt))
    solver.initialize(dt, w["admm_iters"], -9.8, w["anderson_m"], True, w["penalty"], A.ORDER_HARD_ZXU)
    setup_s = time.perf_counter() - t0
    finfo, linfo = solver.factor_info(), solver.ldlt_stats()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        solver.set_pins(pidx, scene.stretch(dt))
        solver.step()
    barrier()
    iters = 0
    loop_ms = 0.0
    launches = 0
    rejects = 0
    with ClockSampler(local if world > 1 else 0) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            solver.set_pins(pidx, scene.stretch(dt))
            last_hist = solver.step()
            info = solver.info()
            iters += info["iter_num"]
            loop_ms += info["loop_ms"]
            launches += info["kernel_launches"]
            rejects += info["rejects"]
        barrier()
        wall_s = time.perf_counter() - t0
    n_free = info["n_free"]
    n_pin = len(pidx)
    h2d = 8 * 3 * (n_free + n_pin)
    d2h = 8 * 3 * n_free + info["iter_num"] * 20

    tot_iters, max_loop_ms, max_wall = iters, loop_ms, wall_s
    if dist is not None:
        import torch
        t = torch.tensor([float(iters), loop_ms, wall_s, float(launches)], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)  # the only collective: per-scene result records
        g = torch.stack(g).cpu()
        tot_iters = float(g[:, 0].sum())
        max_loop_ms = float(g[:, 1].max())
        max_wall = float(g[:, 2].max())
        launches = int(g[:, 3].sum())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = tot_iters / (max_loop_ms * 1e-3)
    e2e = tot_iters / max_wall
    peak, peak_src = measured_peaks()
    prof = solver.profile(20, w["anderson_m"], True)
    phases = {k: v for k, v in prof.items() if k != "total" and v["ms"] > 0}
    dom = max(phases, key=lambda k: phases[k]["ms"])
    ach = phases[dom]["bytes"] / (phases[dom]["ms"] * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one apply (k_fwd_front + k_bwd_front): taken from the committed ncu
    # launch list of this very command (profiles/ldlt_dram_traffic.json, written by tests/tools/ldlt_traffic_from_launches.py);
    # a bench run has no profiler attached, so the figure is read, not re-measured - `traffic_source` says from where
    traffic, traffic_src = _ldlt_traffic() if dom == "ldlt_apply" else (None, None)
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel_launches": "ldlt_apply = k_fwd_front + k_bwd_front (one launch per sweep)" if dom == "ldlt_apply" else dom,
            "phases": {k: {"ms": round(v["ms"], 4), "algo_GB": round(v["bytes"] / 1e9, 4),
                           "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in phases.items()}}
    cpu = None
    if world == 1 and not args.no_cpu:
        with _NativeStdoutToStderr():
            cpu = cpu_baseline_leg()
    # time-to-tolerance of the last timed frame (BASELINE metric): device time until the logged combined
    # residual first falls below tau. tau = 1e-20 is the reference's own break threshold (hard/src/Solver.cpp:93,188),
    # the relative thresholds are fractions of the frame's first logged residual.
    comb = last_hist[:, 1]
    ms_per_it = info["loop_ms"] / max(1, info["iter_num"])

    def first_below(tau):
        k = np.nonzero(comb < tau)[0]
        if len(k):
            return int(k[0]) + 1
        # the iteration that triggers the reference's `break` is not logged (hard/src/Solver.cpp:188-189 vs :210-212)
        return int(info["iter_num"]) if (tau == 1e-20 and info["iter_num"] > len(comb)) else None

    ttt = {"frame_iterations": int(len(comb)), "comb_first": float(comb[0]), "comb_last": float(comb[-1]),
           "ms_per_iteration": ms_per_it, "tau": {}}
    for name, tau in (("abs_1e-20", 1e-20), ("rel_1e-2", 1e-2 * comb[0]), ("rel_1e-4", 1e-4 * comb[0]),
                      ("rel_1e-6", 1e-6 * comb[0])):
        k = first_below(tau)
        ttt["tau"][name] = {"iterations": k, "ms": None if k is None else k * ms_per_it}
    line = {"metric": "admm_anderson_iterations_per_sec_1M_tets", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_loop_ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg4: admm_anderson_hard_zxu ordering, ONE linear-elastic beam %dx%dx%d cubes = %d tets, "
                                   "%d nodes (%d pinned), Anderson m=%d, %d ADMM iterations per frame, dt=1/30, rho=1; "
                                   "per GPU one independent copy of the scene (ensemble sharding, no data-path collective)"
                                   % (w["cx"], w["cy"], w["cz"], len(tets), len(verts), n_pin, w["anderson_m"], w["admm_iters"]),
                       "l2": "per-iteration working set (history + factor, several GB) is larger than the 126 MB L2",
                       "iterations_timed": tot_iters, "rejects": rejects, "setup_s": round(setup_s, 2),
                       "factor": {"nnz_L": finfo["nnz_L"], "numeric_s": round(finfo["seconds_numeric"], 2),
                                  "levels": linfo["levels"], "blocks": linfo["blocks"], "max_block": linfo["max_block"]}},
            "roofline": roof, "cpu_baseline": cpu, "time_to_tol": ttt,
            "e2e": {"value": e2e, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk.summary()}
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


CFG5 = dict(scenes=64, cx=88, cy=22, cz=22, frames=1, admm_iters=100, anderson_m=5, slots=2)
CFG5_CPU_SAMPLE = dict(cx=24, cy=22, cz=22, iters=10)


def run_cfg5(args):
    """BASELINE configs[4] (SURVEY 8e cfg 5): 64 independent scenes 88x22x22 (212,960 tets) with the material sweep of
    aa_admm_b200.ensemble.scene_material, scene -> GPU map static in the first pass and balanced by measured cost afterwards, no data-path collective, one NCCL gather of the result
    records. A bench step = ONE PASS over the whole ensemble (every scene: material -> system-matrix values -> numeric
    LDL^T on the device -> one frame of <= 100 ADMM iterations). Per GPU `slots` scenes are resident and pipelined by
    one host thread each. Strong scaling: the ensemble is fixed, N GPUs share it."""
    rank, world, local = dist_env()
    c = dict(CFG5)
    if args.small:
        c.update(scenes=8, cx=24, cy=6, cz=6)
    c["slots"] = args.slots or c["slots"]
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(max(1, cores // (world * c["slots"])))
    import aa_admm_b200 as A
    from aa_admm_b200 import ensemble as E
    if A.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    dev = local if world > 1 else 0
    A.set_device(dev)
    dist = device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)
    dims = (c["cx"], c["cy"], c["cz"])
    mine = E.scenes_of_rank(c["scenes"], rank, world)
    slots = [E.SceneSlot(A, dims, iters=c["admm_iters"], anderson_m=c["anderson_m"], device=dev) for _ in range(c["slots"])]

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    t0 = time.perf_counter()
    recs, setups = E.run_sweep(A, dims, mine, slots, c["frames"], rank)   # first pass: includes the one-time analysis
    first_pass_s = time.perf_counter() - t0
    full_setup_ms = [ms for _, ms, inc in setups if not inc]
    # optional: the members' costs repeat from pass to pass, so the map of the following passes can balance the measured
    # loop times (one gather of the first pass' records; still no collective on the data path)
    if args.balance_by_cost:
        mine = E.scenes_by_cost(E.gather_records(recs, dist, device), rank, world)
    for _ in range(max(0, args.warmup - 1)):
        E.run_sweep(A, dims, mine, slots, c["frames"], rank)
    barrier()
    iters = 0.0
    loop_ms_sum = 0.0
    inc_setup_ms = []
    with ClockSampler(dev) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            recs, setups = E.run_sweep(A, dims, mine, slots, c["frames"], rank)
            iters += recs[:, 1].sum()
            loop_ms_sum += recs[:, 5].sum()
            inc_setup_ms += [ms for _, ms, inc in setups if inc]
        barrier()
        wall_s = time.perf_counter() - t0
    loop_ms = float(loop_ms_sum) / max(1, len(slots))  # device time of the loops per resident-scene slot (the slots overlap)
    table = E.gather_records(recs, dist, device)  # the only collective: per-scene result records of the last pass
    tot_iters, max_wall, max_loop = iters, wall_s, loop_ms
    if dist is not None:
        import torch
        t = torch.tensor([iters, wall_s, loop_ms], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        g = torch.stack(g).cpu()
        tot_iters, max_wall, max_loop = float(g[:, 0].sum()), float(g[:, 1].max()), float(g[:, 2].max())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    s0 = slots[0].solver
    info = s0.info()
    peak, peak_src = measured_peaks()
    prof = s0.profile(20, c["anderson_m"], True)
    phases = {k: v for k, v in prof.items() if k != "total" and v["ms"] > 0}
    dom = max(phases, key=lambda k: phases[k]["ms"])
    ach = phases[dom]["bytes"] / (phases[dom]["ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": peak_src, "note": "one scene alone on the device (profile of slot 0); in the timed passes the loops of the resident scenes overlap",
            "phases": {k: {"ms": round(v["ms"], 4), "algo_GB": round(v["bytes"] / 1e9, 4),
                           "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in phases.items()}}
    cpu = None
    if world == 1 and not args.no_cpu:
        with _NativeStdoutToStderr():
            cpu = cpu_baseline_leg(CFG5_CPU_SAMPLE, c["cx"] * c["cy"] * c["cz"] * 5)
    n_free, n_pin = info["n_free"], len(slots[0].pidx)
    scenes_timed = c["scenes"] * args.steps
    line = {"metric": "admm_anderson_iterations_per_sec_ensemble", "value": tot_iters / (max_loop * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * max_wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("cfg5: ensemble of %d independent scenes, beam %dx%dx%d = %d tets each, material sweep E = 1e6..1e8, "
                                    "nu = 0.30..0.44, hard_zxu ordering, Anderson m=%d, %d frame(s) x <= %d ADMM iterations per scene; first pass: "
                                    "scene s on GPU (s + s/8) mod N, timed passes: %s, %d resident scenes (host threads) per GPU; one step = one "
                                    "pass over the ensemble, per-scene numeric setup (system-matrix values, numeric LDL^T on the device, moduli) "
                                    "inside the timed region")
                                   % (c["scenes"], dims[0], dims[1], dims[2], info["n_tets"], c["anderson_m"], c["frames"], c["admm_iters"],
                                      "the same static map" if not args.balance_by_cost else
                                      "scenes spread over the GPUs by the loop times measured in the first pass (longest first; one gather of result records)",
                                      c["slots"]),
                       "l2": "two resident scenes of 0.6 GB each per GPU: larger than the 126 MB L2",
                       "iterations_timed": tot_iters, "scenes_timed": scenes_timed,
                       "first_pass_s": round(first_pass_s, 3), "full_setup_ms_per_slot": [round(x, 1) for x in full_setup_ms],
                       "incremental_setup_ms_mean": round(sum(inc_setup_ms) / max(1, len(inc_setup_ms)), 2)},
            "scenes_per_s": scenes_timed / max_wall,
            "value_definition": "iterations / device time of the ADMM loops (CUDA events per frame, summed per resident-scene slot, max over ranks)",
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": tot_iters / max_wall, "unit": "iterations/s",
                    "definition": "iterations / wall time of the timed passes: per scene Solver::set_material + initialize() (values, H2D of the "
                                  "matrix, numeric factorisation) + set_pins + step() with host buffers",
                    "h2d_bytes_per_step": int(len(mine) * (8 * 3 * (n_free + n_pin) + 8 * 3.7e6)), "d2h_bytes_per_step": int(len(mine) * 8 * 3 * n_free)},
            "gpu_launches": int(len(mine) * args.steps * (info["kernel_launches"] + 60)), "clocks": clk.summary(),
            "records_last_pass": {"fields": list(E.RECORD_FIELDS), "rows": table.tolist()}}
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _golden(path_rel):
    p = os.path.join(ROOT, path_rel)
    if not os.path.exists(p):
        raise SystemExit("bench.py: %s is missing (generated by tests/golden/make_golden_geo.py in the build container; it travels with the snapshot)" % path_rel)
    return np.load(p)


def _write_obj(path, V, faces):
    with open(path, "w") as f:
        for v in V:
            f.write("v %.17g %.17g %.17g\n" % tuple(v))
        for fc in faces:
            f.write("f " + " ".join(str(int(i) + 1) for i in fc if i >= 0) + "\n")


def _geo_cpu_baseline(which, files, iters, full_iters):
    """The unmodified reference application (oracle/_ref/libref_planarity.so / libref_wiremesh.so: PlanarityOpt.cpp /
    WireMeshOpt.cpp main()) on the same mesh files, `iters` iterations, all host cores."""
    import ctypes as C
    import tempfile
    try:
        cores = _reference_host_threads()
        lib = os.path.join(ROOT, "oracle", "_ref", "libref_planarity.so" if which == "cfg2" else "libref_wiremesh.so")
        if not os.path.exists(lib):
            return {"value": None, "unit": "iterations/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref absent"}
        L = C.CDLL(lib)
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "result"))
            with open(os.path.join(d, "Options.txt"), "w") as f:
                f.write("Iterations  %d\nAndersonM  5\n" % iters)
            argv = [b"app", files[0].encode(), files[1].encode(), os.path.join(d, "Options.txt").encode(), os.path.join(d, "out.obj").encode()]
            arr = (C.c_char_p * len(argv))(*argv)
            cwd = os.getcwd()
            os.chdir(d)
            try:
                t0 = time.perf_counter()
                rc = L.ref_app_main(len(argv), arr)
                wall = time.perf_counter() - t0
            finally:
                os.chdir(cwd)
            hist = np.loadtxt(os.path.join(d, "result", "residual-5.txt")).reshape(-1, 2)
        loop_s = float(hist[-1, 0])
        return {"value": len(hist) / loop_s, "unit": "iterations/s", "cores": cores, "kind": "reference",
                "sample": "unmodified reference application (%s main) on the same mesh files, %d of the workload's %d iterations, m=5: "
                          "solve_ADMM loop %.3f s (the application's own timer; its setup and file handling %.1f s excluded), rc %d"
                          % ("PlanarityOpt" if which == "cfg2" else "WireMeshOpt", len(hist), full_iters, loop_s, wall - loop_s, rc)}
    except Exception as e:
        return {"value": None, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}


def run_cfg_geo(args, which):
    """BASELINE configs[1] / configs[2] (SURVEY 8d cfg 2 / cfg 3): the reference's PlanarityOpt on costa2k_poly (planar
    quads) and WireMeshOpt on MaleTorso (one subdivision: 230,400 points, 1.38 M hard constraints), 100 iterations, m=5.
    The meshes come from the golden fixtures (arrays of the reference's shipped files); the product's own front-end
    (host/GeometryApps) reads them as .obj, subdivides, builds the constraints and sets up once. A bench step = one
    solve_ADMM of 100 iterations from the input positions."""
    import tempfile
    import aa_admm_b200 as A
    if A.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    A.set_device(0)
    iters, m = 100, 5
    tmp = tempfile.mkdtemp(prefix="aaadmm_bench_")
    if which == "cfg2":
        g = _golden("tests/golden/geo_costa2k.npz")
        files = (os.path.join(tmp, "poly.obj"), os.path.join(tmp, "tri.obj"))
        _write_obj(files[0], g["P"], g["faces"])
        _write_obj(files[1], g["Vref"], g["Fref"])
        name = "cfg2: PlanarityOpt costa2k_poly.obj costa2k_tri.obj (planar quads, closeness, relative Laplacians), rho = 1e5"
        t0 = time.perf_counter()
        mesh, ref = A.PolyMesh.load(files[0]), A.PolyMesh.load(files[1])
        app = A.GeoApp("planarity", mesh, ref, [1e5, 1.0, 0.0, 0.1])
    else:
        g = _golden("tests/golden_large/geo_maletorso.npz")
        files = (os.path.join(tmp, "quad.obj"), os.path.join(tmp, "target.obj"))
        _write_obj(files[0], g["P0"], g["quads0"])
        _write_obj(files[1], g["Vref"], g["Fref"])
        name = "cfg3: WireMeshOpt MaleTorso.obj MaleTorso_target.obj (one subdivision + smoothing, angle / edge-length constraints, closeness), rho = 1e3"
        t0 = time.perf_counter()
        coarse, ref = A.PolyMesh.load(files[0]), A.PolyMesh.load(files[1])
        el = 0.5 * coarse.counts()["average_edge_length"]
        mesh = coarse.subdivide_and_smooth()
        app = A.GeoApp("wiremesh", mesh, ref, [1e3, 0.25 * np.pi, 0.75 * np.pi, el, 1.0, -1.0])
    setup_s = time.perf_counter() - t0
    st = app.stats()
    for _ in range(args.warmup):
        app.solve(iters, m, False)
    tot_it, loop_ms, wall_s, launches, resets = 0, 0.0, 0.0, 0, 0
    with ClockSampler(0) as clk:
        for _ in range(args.steps):
            t0 = time.perf_counter()
            hist, x, info = app.solve(iters, m, True)
            wall_s += time.perf_counter() - t0
            tot_it += len(hist)
            loop_ms += info["loop_ms"]
            launches += info["kernel_launches"]
            resets += info["resets"]
    # algorithmic bytes of one loop turn: the factor apply (every factor value once per sweep + vectors), the hard
    # constraints (indices, gathered points, z, u, Dx_prev read and written), the right-hand side (rho D^T entries) and the
    # Anderson passes over [u | x]; the closest-point search adds data-dependent BVH traffic that is not counted
    P, H, Z = st["points"], st["hard_constraints"], st["z_columns"]
    N = 3.0 * (Z + P)
    turn_bytes = st["bytes_per_apply"] + H * 16.0 + Z * 24.0 * 7 + Z * 12.0 * 2 + P * 24.0 * 4 + 8.0 * ((m + 2) * N + 3 * N) + 8.0 * ((m + 3) * N + 2 * N)
    turns = tot_it + resets
    ms_turn = loop_ms / max(1, turns)
    peak, peak_src = measured_peaks()
    ach = turn_bytes / (ms_turn * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        with _NativeStdoutToStderr():
            cpu = _geo_cpu_baseline(which, files, 100 if which == "cfg2" else 10, iters)
    line = {"metric": "admm_anderson_iterations_per_sec_geometry", "value": tot_it / (loop_ms * 1e-3), "unit": "iterations/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": loop_ms / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (arrays of the reference's shipped mesh)",
            "config": {"workload": name + ", %d iterations, Anderson m=%d; ALMGeometrySolver<3>::solve_ADMM on the device" % (iters, m),
                       "l2": ("working set of cfg 2 (a few MB) is L2-resident: launch-latency bound" if which == "cfg2" else
                              "factor + constraint state of cfg 3 (0.3 GB per turn) is larger than the 126 MB L2"),
                       "sizes": st, "setup_s": round(setup_s, 2), "iterations_timed": tot_it, "rejected_turns": resets,
                       "final_residual": float(hist[-1])},
            "roofline": {"bound": "hbm", "kernel": "one loop turn (k_geo_local + k_geo_soft + k_geo_rhs + ldlt_apply + k_geo_u_resid + Anderson passes)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "ms_per_turn": ms_turn, "algo_GB_per_turn": turn_bytes / 1e9,
                         "per_kernel": "profiles/r02_%s_launches.md (ncu launch list of this command)" % which},
            "cpu_baseline": cpu,
            "e2e": {"value": tot_it / wall_s, "unit": "iterations/s", "h2d_bytes_per_step": 24 * P, "d2h_bytes_per_step": 24 * P + 8 * iters},
            "gpu_launches": launches, "clocks": clk.summary()}
    _emit(line)
    return 0


def run_cfg1(args):
    """BASELINE configs[0] (SURVEY 8d cfg 1): admm_anderson_xzu on the sample's three beams 12x3x3 (LINEAR / Neo-Hookean /
    StVK), Anderson m=5, 100 iterations per frame. 1,620 tets: everything is L2-resident and launch-latency bound; the
    line exists for coverage of the xzu loop and the per-tet L-BFGS prox."""
    import aa_admm_b200 as A
    if A.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    A.set_device(0)
    dims, dt, iters, m = (12, 3, 3), 1.0 / 30.0, 100, 5

    def build(make):
        s = make()
        scene = A.BeamScene()
        for shift, mat in ((1.75, 0), (0.0, 1), (-1.75, 2)):
            v, t, mm, _, _, _ = A.BeamScene().add(*dims, shift).arrays()
            s.add_tetmesh(v, t, mm, 1e7, 0.399, mat)
            scene.add(*dims, shift)
        pidx = scene.arrays()[3]
        s.set_pins(pidx, scene.stretch(dt))
        return s, scene, pidx

    s, scene, pidx = build(A.Solver)
    s.initialize(dt, iters, -9.8, m, True, 1.0, A.ORDER_XZU)
    for _ in range(args.warmup):
        s.set_pins(pidx, scene.stretch(dt))
        s.step()
    tot_it, loop_ms, launches, rejects = 0, 0.0, 0, 0
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s.set_pins(pidx, scene.stretch(dt))
            s.step()
            info = s.info()
            tot_it += info["iter_num"]
            loop_ms += info["loop_ms"]
            launches += info["kernel_launches"]
            rejects += info["rejects"]
        wall_s = time.perf_counter() - t0
    cpu = None
    if not args.no_cpu:
        try:
            with _NativeStdoutToStderr():
                cores = _reference_host_threads()
                from oracle import refbind
                r, rscene, rp = build(lambda: refbind.RefSolver("xzu"))
                r._f("set_threads")(cores)
                r.initialize(dt, iters, -9.8, m, True, 1.0)
                its, secs = 0, 0.0
                for f in range(4):
                    r.set_pins(rp, rscene.stretch(dt))
                    t0 = time.perf_counter()
                    h = r.step()
                    if f > 0:
                        secs += time.perf_counter() - t0
                        its += len(h)
            cpu = {"value": its / secs, "unit": "iterations/s", "cores": cores, "kind": "reference",
                   "sample": "unmodified reference admm_anderson_xzu Solver::step on the same three beams, 3 frames x 100 iterations after 1 warm-up frame (the whole workload)"}
        except Exception as e:
            cpu = {"value": None, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
    by = np.zeros(8)
    A.cuda_lib().aaadmm_tetscene_algo_bytes(s._scene(), m, by.ctypes.data_as(A.c_dp))
    turn_bytes = 3 * by[0] + 3 * (by[1] + by[2]) + by[3] + by[4] + by[5]  # 3 local steps, 3 solves (one only logs), u, Anderson
    ms_it = loop_ms / max(1, tot_it)
    peak, peak_src = measured_peaks()
    ach = turn_bytes / (ms_it * 1e-3) / 1e9
    n_free = info["n_free"]
    line = {"metric": "admm_anderson_iterations_per_sec_xzu_sample", "value": tot_it / (loop_ms * 1e-3), "unit": "iterations/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": loop_ms / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg1: admm_anderson_xzu ordering, three beams 12x3x3 (LINEAR, Neo-Hookean, StVK; %d tets, %d free vertices), "
                                   "Anderson m=%d, %d iterations per frame" % (info["n_tets"], n_free, m, iters),
                       "l2": "the whole state (0.5 MB) is L2-resident: this configuration is bound by launch latency (about 22 kernels per iteration), not by HBM",
                       "iterations_timed": tot_it, "rejects": rejects},
            "roofline": {"bound": "hbm", "kernel": "one xzu iteration (3 local steps incl. the per-tet L-BFGS prox, 3 solves, Anderson passes)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "ms_per_iteration": ms_it, "algo_GB_per_iteration": turn_bytes / 1e9,
                         "per_kernel": "profiles/r02_cfg1_launches.md (ncu launch list of this command)"},
            "cpu_baseline": cpu,
            "e2e": {"value": tot_it / wall_s, "unit": "iterations/s", "h2d_bytes_per_step": 24 * (n_free + len(pidx)), "d2h_bytes_per_step": 24 * n_free + 20 * iters},
            "gpu_launches": launches, "clocks": clk.summary()}
    _emit(line)
    return 0


def _ldlt_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "ldlt_dram_traffic.json")) as f:
            d = json.load(f)
        return float(d["bytes_per_apply"]), d["source"]
    except (OSError, ValueError, KeyError):
        return None, None


_JSON_OUT = None


def _emit(line):
    """The ONE JSON line, on the real stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the JSON line only: libraries print there as well (NCCL its version banner, the reference its progress
    # lines), so file descriptor 1 points at stderr for the whole run and the line goes to a private copy of the real stdout
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="30,720-tet beam (debugging only; not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg4 (default, the headline): one 1M-tet beam; cfg5: ensemble of 64 x 213k-tet scenes (material sweep); "
                         "cfg1: xzu sample (three beams); cfg2 / cfg3: Geometry PlanarityOpt / WireMeshOpt (one GPU)")
    ap.add_argument("--slots", type=int, default=0, help="cfg5: resident scenes (host threads) per GPU (default 2)")
    ap.add_argument("--balance-by-cost", action="store_true",
                    help="cfg5: the passes after the first spread the scenes over the GPUs by their measured loop times (longest first) "
                         "instead of the static diagonal map; measured at 8 GPUs: 200 against 206 scenes/s, so off by default")
    ap.add_argument("--ref-dims", type=int, nargs=3, default=None,
                    help="--impl reference: beam size of the CPU arm (default REF_ARM; smaller sizes are for the CPU test)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "cfg5":
        return run_cfg5(args)
    if args.config in ("cfg1", "cfg2", "cfg3"):
        rank, world, _ = dist_env()
        if rank != 0:  # one coupled problem: one GPU
            return 0
        return run_cfg1(args) if args.config == "cfg1" else run_cfg_geo(args, args.config)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
