#!/usr/bin/env python
"""bench.py — scan-matches/sec of the PSO/NDT hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload cfg2|cfg5] [--impl b200|reference]

A step = one pass of the hot path (pso_optimization: the whole swarm run) over a batch of B independent scan-match problems per
GPU, each carrying its own dense (mu, Sigma^-1, built) table.  Workloads:
  cfg2 (default)  BASELINE.json configs[1] shape batched as configs[2] / configs[3]: 1081-beam scans vs 50 m / 0.5 m NDT maps,
                  70 particles x 50 iterations; problem b of rank r is step r*B + b of the replayed trajectory
  cfg5            BASELINE.json configs[4]: the multi-resolution sweep, cell sides 0.25 / 0.5 / 1 / 2 m, 200 particles x 100
                  iterations; the items (cell side, frame) go round-robin over the ranks by frame, so every GPU holds B/4
                  frames of every cell side (B = 148 by default)

  value      whole-job matches/s with the batch resident in HBM (dense tables, points, guesses): every step runs
             K0 table compaction + K1 rand() stream + K2 PSO (+ the result exchange for N > 1); four resident copies
             of the batch are cycled on two streams so that consecutive steps overlap; CUDA events around the K steps
  sustained  the same loop kept running for at least 2 s (clocks sampled meanwhile)
  single_stream  the same with one copy on one stream, L2 flushed between steps (per-step events)
  e2e        the same metric through ndtpso_align_submit/collect with HOST buffers (pinned): H2D of every
             input + kernels + D2H of the poses inside the timed region, three batches in flight (--e2e-depth)
  parity     every result of the timed batch against the unmodified reference: golden vectors for all of them (every rank),
             and the reference itself run in this bench on the batch's own first problems (the cpu_baseline leg)
  tracking   the reference's whole per-scan callback (loadLaser -> align -> update) on device-resident maps (cfg2)
  roofline   dominant kernel (pso_sliced_kernel): algorithmic bytes / its duration vs the measured HBM peak; beside it the fp64
             pipe fraction and the shared-memory data pipe, which is the resource that binds (ncu, profiles/)
  cpu_baseline  the reference's own CPU path (oracle/_ref if present, else the oracle port) on this box's host cores, on the
             batch's own first problems

`--impl reference` times only the CPU arm, same metric/config.
"""
import argparse
import collections
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POSE_ATOL, SCORE_RTOL = 1e-4, 1e-5  # BASELINE.json north_star: parity bar of the path
CFG5_SIZES = (0.25, 0.5, 1.0, 2.0)


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
class Workload:
    """What a rank solves per step, and where the reference's answers for it come from."""

    def __init__(self, name, batch):
        self.name, self.B = name, batch
        if name == "cfg2":
            self.P, self.I = 70, 50
            self.metric = "scan-matches/sec (1081-pt scan, 70 particles x 50 iters)"
        else:
            self.P, self.I = 200, 100
            self.metric = "scan-matches/sec (1081-pt scan, 200 particles x 100 iters, multi-resolution sweep 0.25/0.5/1/2 m)"
            if batch % 4:
                raise SystemExit("bench.py: --batch must be a multiple of 4 for the cfg5 workload (four cell sides per frame)")

    def specs(self, rank, world):
        """One (cell side, frame index) per problem of this rank; the frame index is also what seeds the problem (1 + frame)."""
        if self.name == "cfg2":
            return [(0.5, rank * self.B + b) for b in range(self.B)]
        # cfg5: item (c, f) -> rank f % world; a rank's batch is frame-major so that the four cell sides alternate
        return [(cs, rank + world * k) for k in range(self.B // 4) for cs in CFG5_SIZES]

    def problems(self, rank, world, sparse=False):
        return [make_problem(self.name, cs, f, sparse) for cs, f in self.specs(rank, world)]

    def alg_bytes(self):
        """16 N + 49 C + 80 bytes per match (mean over the batch for the sweep)."""
        cells = [int(round(50.0 / cs)) ** 2 for cs in ((0.5,) if self.name == "cfg2" else CFG5_SIZES)]
        return float(np.mean([16 * 1081 + 49 * c + 80 for c in cells]))

    def alg_flop(self):
        return 31 * (self.P + 1 + self.P * self.I) * 1081 + 17 * self.P * self.I

    def describe(self):
        if self.name == "cfg2":
            return (f"cfg2 shape (configs[1]; batched as configs[2]): {self.B} independent 1081-beam scan-matches per GPU vs "
                    "50 m/0.5 m NDT maps (one dense table per problem), 70 particles x 50 iterations")
        return (f"cfg5 (configs[4]): multi-resolution sweep, {self.B} scan-matches per GPU = {self.B // 4} frames x cell sides "
                "0.25/0.5/1/2 m of a 50 m map (one dense table per problem), 200 particles x 100 iterations, items round-robin over the ranks by frame")

    def golden(self, rank, world):
        """(pose[B, 3], cost[B]) of the unmodified reference for this rank's problems, or None where no vector is committed."""
        if self.name != "cfg2":
            return None
        try:
            z = np.load(os.path.join(ROOT, "tests", "golden", "batch_vectors.npz"))
        except Exception:
            return None
        lo, hi = rank * self.B, (rank + 1) * self.B
        if hi > z["traj/pose"].shape[0] or tuple(int(v) for v in z["traj/pso"]) != (self.P, self.I):
            return None
        return z["traj/pose"][lo:hi], z["traj/cost"][lo:hi]


def scanset_of(workload, cs, f):
    from ndtpso_slam_b200 import synthetic as syn
    return syn.trajectory_problem(syn.CFG2 if workload == "cfg2" else syn.CFG5[cs], f)


def make_problem(workload, cs, f, sparse=False):
    from ndtpso_slam_b200 import frames
    return frames.problem_from_scans(scanset_of(workload, cs, f), sparse=sparse, seed=1 + f)


def parity_stats(pose, cost, want_pose, want_cost):
    pose, want_pose = np.asarray(pose).reshape(-1, 3), np.asarray(want_pose).reshape(-1, 3)
    cost, want_cost = np.asarray(cost).reshape(-1), np.asarray(want_cost).reshape(-1)
    dp = np.abs(pose - want_pose).max(axis=1) if len(pose) else np.zeros(0)
    ds = np.where(cost == want_cost, 0.0, np.abs(cost - want_cost) / np.maximum(np.abs(want_cost), 1e-300)) if len(cost) else np.zeros(0)
    return {"n_checked": int(len(dp)), "max_abs_dpose": float(dp.max()) if len(dp) else 0.0,
            "max_rel_dscore": float(ds.max()) if len(ds) else 0.0, "bit_exact_poses": int((dp == 0).sum())}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled while the timed regions run.  In-process NVML (5 ms period; works
    for short regions and for eight ranks at once); `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.source = index, [], None, None, None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES reorders devices: go by UUID
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown, n.nvmlClocksEventReasonSwThermalSlowdown,
                n.nvmlClocksEventReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append([sm, self.sm_max, 0.0] + ["active" if mask & b else "not active" for b in bits])
            except Exception:
                pass
            self._stop.wait(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(self.NAMES, r[3:7]):
                if str(v).lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "source": self.source, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Solves the given problems of the workload one after the other on one thread; returns (seconds spent in the solves,
    pose[n, 3], cost[n]).  Map building is outside the timed part (the GPU arm's batches are built before its clock starts too)."""
    kind, workload, P, I, specs = args
    from oracle import binding
    poses, costs, secs = [], [], 0.0
    if kind == "reference":
        R = binding.Reference()
        built = []
        for cs, f in specs:
            ss = scanset_of(workload, cs, f)
            rf, q = R.build_problem(ss)
            rf.build()
            built.append((ss, rf, q, f))
        for ss, rf, q, f in built:
            t0 = time.perf_counter()
            pose, _ = R.pso(rf, q, ss.guess, ss.deviation, P, I, seed=1 + f, num_threads=1)
            secs += time.perf_counter() - t0
            poses.append(pose)
            costs.append(R.cost(rf, q, pose))
    else:
        O = binding.Oracle()
        flats = [make_problem(workload, cs, f) for cs, f in specs]
        for fl in flats:
            t0 = time.perf_counter()
            pose, cost, _ = O.pso(fl, fl["guess"], fl["deviation"], P, I, seed=fl["seed"])
            secs += time.perf_counter() - t0
            poses.append(pose)
            costs.append(cost)
    return secs, np.array(poses).reshape(-1, 3), np.array(costs)


def cpu_reference_rate(wl, matches_per_worker, as_shipped=True):
    """matches/s of the reference CPU path using every host core: nproc single-thread workers over disjoint problems (its
    deterministic mode; in-process outer parallelism is impossible because the reference draws from the process-global rand()).
    The problems are the GPU batch's own first ones (rank 0's, wrapping around), so the results double as a live parity check:
    returns (baseline dict, specs solved, pose, cost).  Also times the as-shipped OpenMP mode."""
    from oracle import binding
    kind = "reference" if os.path.exists(binding.REF_SO) else "port"
    if kind == "port":
        binding.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    own = wl.specs(0, 1)
    specs = [own[i % len(own)] for i in range(cores * matches_per_worker)]
    jobs = [(kind, wl.name, wl.P, wl.I, specs[w * matches_per_worker:(w + 1) * matches_per_worker]) for w in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    per = [r[0] for r in res]
    rate = cores * matches_per_worker / max(per)
    out = {"value": rate, "unit": "scan-matches/s", "cores": cores, "kind": kind, "per_core": rate / cores,
           "sample": f"{cores} single-thread workers x {matches_per_worker} matches each = the batch's own first {min(len(specs), len(own))} problems "
                     f"({wl.P}x{wl.I}, 1081 pts), slowest worker {max(per):.2f} s",
           "single_thread_ms_per_match": 1e3 * float(np.median(per)) / matches_per_worker}
    if kind == "reference" and as_shipped:
        R = binding.Reference()
        cs, f = own[0]
        ss = scanset_of(wl.name, cs, f)
        rf, q = R.build_problem(ss)
        # torchrun exports OMP_NUM_THREADS=1, and pso_optimization itself lowers the team size whenever num_threads limits it
        # (core.cpp:75-79): put it back to the cores this process may use before timing the as-shipped mode
        R.lib.ref_omp_set_num_threads(cores)
        ts = [R.pso(rf, q, ss.guess, ss.deviation, wl.P, wl.I, seed=s, num_threads=-1)[1] for s in range(1, 9)]
        out["as_shipped_openmp"] = {"value": 1.0 / float(np.median(ts[2:])), "unit": "scan-matches/s", "threads": R.lib.ref_omp_max_threads(),
                                    "note": "num_threads=-1 inside one match; non-deterministic result (SURVEY.md section 0.5)"}
    return out, specs, np.concatenate([r[1] for r in res]), np.concatenate([r[2] for r in res])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = Workload(args.workload, args.batch)
    per_step = max(1, args.ref_matches)
    rates = []
    base = None
    t_all = time.perf_counter()
    for step in range(args.warmup + args.steps):
        base, _, _, _ = cpu_reference_rate(wl, per_step, as_shipped=(step == 0))
        if step >= args.warmup:
            rates.append(base["value"])
        if time.perf_counter() - t_all > 240:
            break
    value = float(np.mean(rates)) if rates else base["value"]
    base["value"] = value
    base["per_core"] = value / base["cores"]
    base["rate_spread"] = [float(min(rates)), float(max(rates))] if rates else None
    line = {"impl": "reference", "metric": wl.metric, "value": value, "unit": "scan-matches/s", "n_gpus": args.gpus, "steps": len(rates),
            "warmup": args.warmup, "ms_per_step": 1e3 * base["cores"] * per_step / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.describe(), "batch_per_gpu": wl.B, "particles": wl.P, "iterations": wl.I},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "scan-matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def measured_profile(workload, batch):
    """Per-launch counters of pso_sliced_kernel from the committed `ncu --set full` capture of this workload (profiles/traffic.json):
    {"dram_bytes": dram__bytes_read.sum + dram__bytes_write.sum, "shared_wavefronts": l1tex__data_pipe_lsu_wavefronts_mem_shared.sum};
    {} for a workload / batch size that was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(f"{workload}:{batch}", {})
    except Exception:
        return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def run_tracking(B, conf, steps, device, groups=2):
    """B robots, each tracked through the reference's per-scan callback (src/ndtpso_slam_node.cpp:177-244) with its map
    resident in HBM (include/ndtpso_dframes.h).  Robot b replays trajectory b of the bench workload: its map is seeded with
    the 5 map scans of synthetic.trajectory_problem(CFG2, b), then every timed step is one ndtpso_dframes_track_step call:
    H2D of the step's ranges (4 bytes per beam), loadLaser, NDTFrame::build + table compaction, rand() stream, PSO,
    NDTFrame::update, D2H of the poses.  A robot's steps are sequentially dependent (the pose of scan k positions scan k in
    the map that scan k+1 is matched against), so within a group nothing is pipelined; the robots are served as `groups`
    independent groups, each by its own host thread, context and stream, so that one group's small kernels and host round
    trips hide behind another group's PSO kernel.  Wall clock around the K steps of all groups."""
    import threading

    from ndtpso_slam_b200 import capi, dframes, synthetic as syn
    cfg = syn.CFG2
    s, S = cfg.sensor, cfg.map_size_m
    room = syn.Room(S)
    groups = max(1, min(groups, B // 32))
    bounds = [(g * B // groups, (g + 1) * B // groups) for g in range(groups)]
    sets = [syn.trajectory_problem(cfg, b) for b in range(B)]
    scans = [np.stack([syn.make_scan(room, s, (ss.true_pose[0] + 0.02 * k, ss.true_pose[1] + 0.005 * k, ss.true_pose[2] + 0.001 * k),
                                     syn.NoiseLCG(777 + 131 * b + k)) for b, ss in enumerate(sets)]) for k in range(steps + 2)]
    init = np.array([ss.guess for ss in sets])
    ctxs, dfs = [], []
    for lo, hi in bounds:
        c = capi.Context(device)
        # pools sized from what the scene needs (389 cells created, < 30 points per cell and scan), not for the worst case: 1024 cells per robot,
        # rings of 256 points per cell (a ring only has to hold a cell's CURRENT slot); NDTPSO_DF_CELL_POOL_FULL / _WINDOW_TRUNCATED guard both
        df = dframes.DeviceFrames(c, hi - lo, S, S, cfg.cell_side, s.beams, max_cells=1024, window_points=256, flags=dframes.DF_NO_CLUSTER if groups > 1 else 0)
        for k in range(5):  # the map: 5 scans per robot merged at their known poses
            df.load_laser(np.stack([ss.map_scans[k][1] for ss in sets[lo:hi]]), s.angle_min, s.angle_increment, s.range_max)
            df.update(np.array([ss.map_scans[k][0] for ss in sets[lo:hi]]))
        df.track_step(scans[0][lo:hi], s.angle_min, s.angle_increment, s.range_max, initial_poses=init[lo:hi], conf=conf)  # first scan: no matching
        df.track_step(scans[1][lo:hi], s.angle_min, s.angle_increment, s.range_max, conf=conf)                            # warm-up
        ctxs.append(c)
        dfs.append(df)
    poses = np.zeros((B, 3))
    start = threading.Barrier(groups + 1)

    def serve(g):
        lo, hi = bounds[g]
        start.wait()
        for k in range(steps):
            poses[lo:hi], _ = dfs[g].track_step(scans[2 + k][lo:hi], s.angle_min, s.angle_increment, s.range_max, conf=conf)

    threads = [threading.Thread(target=serve, args=(g,)) for g in range(groups)]
    for t in threads:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    wall = time.perf_counter() - t0
    kt = dfs[0].kernel_times_ms()
    true_last = np.array([(ss.true_pose[0] + 0.02 * (steps + 1), ss.true_pose[1] + 0.005 * (steps + 1), ss.true_pose[2] + 0.001 * (steps + 1))
                          for ss in sets])
    err = np.abs(poses - true_last)
    info = dfs[0].info(0)
    flags = 0
    for df in dfs:
        flags |= int(np.bitwise_or.reduce(df.status()))
    out = {"value": B * steps / wall, "unit": "scan-matches/s", "robots": B, "groups": groups, "steps": steps, "ms_per_step": 1e3 * wall / steps,
           "h2d_bytes_per_step": int(B * s.beams * 4), "d2h_bytes_per_step": int(B * 32), "kernel_ms_last_step_group0": kt,
           "device_bytes": sum(df.device_bytes() for df in dfs), "cells_created_robot0": info["created"], "status_bits": flags,
           "median_abs_pose_error_vs_truth": [float(v) for v in np.median(err, axis=0)],
           "api": "ndtpso_dframes_track_step: loadLaser + build + PSO 70x50 + update per call, maps resident in HBM"}
    for df, c in zip(dfs, ctxs):
        df.close()
        c.close()
    return out


def run_callback(steps, population, iterations):
    """The reference's whole per-scan callback (src/ndtpso_slam_node.cpp:186-198: loadLaser -> align -> update) for ONE robot,
    written against the drop-in NDTFrame (libndtpso_slam.so, what the ROS node links): the map stays in HBM (device mirror), the
    scan crosses PCIe as 4 bytes per beam.  Beside it the unmodified reference's callback on this box: on one core and as shipped
    (OpenMP team inside the match).  population/iterations None = the 2-argument align(), which always runs 30 x 50."""
    import ctypes as C

    from ndtpso_slam_b200 import capi, frames, synthetic as syn
    cfg = syn.CFG2
    s, S = cfg.sensor, cfg.map_size_m
    room = syn.Room(S)
    ss = syn.trajectory_problem(cfg, 0)
    scans = [r for _, r in ss.map_scans] + [syn.make_scan(room, s, (ss.true_pose[0] + 0.02 * k, ss.true_pose[1] + 0.005 * k, ss.true_pose[2] + 0.001 * k),
                                                          syn.NoiseLCG(4242 + k)) for k in range(steps)]
    init = tuple(ss.map_scans[0][0])
    conf = capi.PsoConfig.make(population=population, iterations=iterations) if population else None

    def drive(make_map, make_scan_frame, align, update, close):
        ref = make_map()
        pose, ts = np.array(init), []
        for k, ranges in enumerate(scans):
            t0 = time.perf_counter()
            cur = make_scan_frame(k == 0)
            cur.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
            if k > 0:
                pose = align(ref, pose, cur)
            update(ref, pose, cur)
            ts.append(time.perf_counter() - t0)
            close(cur)
        return ref, pose, ts

    C.CDLL(None).srand(1)
    ref, pose_gpu, ts = drive(lambda: frames.Frame(width=S, height=S, cell_side=cfg.cell_side, calculate_cells_params=True),
                              lambda first: frames.Frame(trans=init, width=S, height=S, cell_side=cfg.cell_side if first else float(S), calculate_cells_params=False),
                              lambda r, p, c: r.align(p, c, conf), lambda r, p, c: r.update(p, c), lambda c: c.close())
    h2d = ref.last_h2d_bytes()
    out = {"ms_per_scan": 1e3 * float(np.median(ts[len(ss.map_scans):])), "scans": steps, "swarm": f"{population or 30} x {iterations or 50}",
           "device_resident_map": bool(ref.device_resident), "h2d_bytes_last_scan": {"align": h2d[0], "update": h2d[1]},
           "api": "drop-in NDTFrame (libndtpso_slam.so): loadLaser + align + update per scan, one robot"}
    ref.close()
    try:
        from oracle import binding
        if os.path.exists(binding.REF_SO):
            R = binding.Reference()
            cores = len(os.sched_getaffinity(0))
            for label, threads in (("reference_one_core", 1), ("reference_as_shipped_openmp", cores)):
                R.lib.ref_omp_set_num_threads(threads)
                R.srand(1)
                if population:  # an explicit swarm: pso_optimization with the deviation rule of align (ndtframe.cpp:253) done here
                    state = {"it": 0, "prev": np.zeros(3), "diff": np.zeros(3)}

                    def align(r, p, c, state=state, threads=threads):
                        dev = (0.1, 0.1, 3.1415e-3) if state["it"] < 2 else tuple(np.abs(2 * state["diff"]))
                        state["it"] += 1
                        po, _ = R.pso(r, c, p, dev, population, iterations, use_seed=False, num_threads=-1 if threads > 1 else 1)
                        state["diff"], state["prev"] = po - state["prev"], po
                        return po
                else:
                    def align(r, p, c):
                        return R.align(r, p, c)
                _, pose_ref, tr = drive(lambda: R.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True),
                                        lambda first: R.frame(trans=init, width=S, height=S, cell_side=cfg.cell_side if first else float(S), init_windows=False),
                                        align, lambda r, p, c: r.update(p, c), lambda c: None)
                out[label] = {"ms_per_scan": 1e3 * float(np.median(tr[len(ss.map_scans):])), "threads": threads}
                if threads == 1:
                    out["max_abs_dpose_vs_reference_one_core"] = float(np.abs(pose_gpu - pose_ref).max())
    except Exception as e:  # the CPU legs are a comparison, not the measurement
        out["reference_error"] = str(e)
    return out


def run_gpu_arm(args):
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _run_gpu_arm(args, real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)


def _run_gpu_arm(args, real_stdout):
    import torch
    import torch.distributed as dist

    from ndtpso_slam_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = Workload(args.workload, args.batch)
    B, P, I = wl.B, wl.P, wl.I

    flats = wl.problems(rank, world)
    ctx = capi.Context(local)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream shared by torch events, NCCL and the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    conf = capi.PsoConfig.make(population=P, iterations=I)

    # pinned host copies of every input array (the e2e path copies from these every step)
    pinned_flats = []
    for f in flats:
        g = dict(f)
        for k in ("points", "mean", "inv_cov", "built"):
            t = torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory()
            g[k] = t.numpy()
            g["_keep_" + k] = t
        pinned_flats.append(g)
    pset = capi.ProblemSet(pinned_flats)
    table_bytes = sum(f["mean"].nbytes + f["inv_cov"].nbytes + f["built"].nbytes for f in flats)

    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: value.  Resident copies of the batch alternate on two streams (step k -> copy k % 4, stream k & 1): a
    # batch of 256 CTAs leaves 40 of the 296 CTA slots empty and drains unevenly, and with the next step on the other
    # stream its CTAs take the free slots at once.  Every step is a full pass of the hot path over one batch (K0 table
    # compaction, K1 rand() streams, K2 PSO, result exchange for N > 1); four copies are cycled so that what a step
    # touches has been pushed out of the 126 MB L2 by the three steps in between.  The single-stream, L2-flushed form is
    # measured beside it (`single_stream`).
    NCOPY = 4
    streams = [stream, torch.cuda.Stream()]
    bts = [ctx.batch(pset, conf) for _ in range(NCOPY)]
    res_ts = [None] * NCOPY
    if world > 1:
        for i in range(NCOPY):
            # view the library's result buffer as a tensor for the NCCL all-gather of the solved poses
            class _Ext:
                __cuda_array_interface__ = {"shape": (B * 4,), "typestr": "<f8", "data": (bts[i].device_results_ptr(), False), "version": 3}
            res_ts[i] = torch.as_tensor(_Ext(), device="cuda")

    # The one exchange of the path: every rank ends up with all solved poses.  Fused form: the PSO kernel's epilogue
    # stores each result into every rank's gathered buffer over NVLink (CUDA IPC) and ndtpso_exchange_wait polls the
    # arrival flags — no collective per step.  NCCL's all-gather is the reference implementation: used to verify the
    # fused form once, and as the per-step exchange if CUDA IPC is unavailable (--exchange nccl forces it).
    exs, exchange_kind = [None] * NCOPY, "none"
    if world > 1:
        exchange_kind = "NCCL all-gather of [B][4] fp64 poses per step"
        if args.exchange == "fused":
            exs = [sharding.make_exchange(ctx, B, world, rank, device="cuda") for _ in range(NCOPY)]  # None on every rank if IPC fails anywhere
            if all(e_ is not None for e_ in exs):
                for i in range(NCOPY):
                    bts[i].attach_exchange(exs[i])
                exchange_kind = ("peer stores of the [B][4] fp64 poses from the PSO kernel's epilogue into every rank's gathered buffer "
                                 "(NVLink, CUDA IPC) + arrival-flag wait kernel; verified against an NCCL all-gather")
            else:
                print(f"rank {rank}: fused exchange unavailable (CUDA IPC); using the NCCL all-gather", file=sys.stderr)
                for e_ in exs:
                    if e_ is not None:
                        e_.close()
                exs = [None] * NCOPY
    ex = exs[0]

    def resident_step(i=0):
        st = streams[i & 1]
        ctx.set_stream(st.cuda_stream)
        bts[i].solve()
        if exs[i] is not None:
            exs[i].wait()
        elif world > 1:
            with torch.cuda.stream(st):
                sharding.gather_results(res_ts[i].view(B, 4), world * B, world, rank)

    gathered = None
    if ex is not None:  # once: the fused exchange delivers exactly what the collective delivers
        resident_step(0)
        class _ExG:
            __cuda_array_interface__ = {"shape": (world * B * 4,), "typestr": "<f8", "data": (ex.device_results_ptr(), False), "version": 3}
        torch.cuda.synchronize()
        fused = torch.as_tensor(_ExG(), device="cuda").view(world * B, 4).clone()
        ref_rows = sharding.gather_results(res_ts[0].view(B, 4), world * B, world, rank)
        torch.cuda.synchronize()
        assert torch.equal(fused, ref_rows), "fused exchange and NCCL all-gather disagree"
        gathered = fused.cpu().numpy()

    # single stream, L2 flushed between steps (not timed): per-step CUDA events
    n_single = max(3, min(args.steps, 10))
    for _ in range(args.warmup):
        l2_flush.fill_(1)
        resident_step(0)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_single)]
    for k in range(n_single):
        l2_flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        resident_step(0)
        ev[k][1].record(stream)
    barrier()
    single_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    kt = bts[0].kernel_times_ms()  # last solve: K0, K1, K2, not overlapped with anything

    # the timed region: K steps cycling the copies / streams
    for k in range(max(args.warmup, NCOPY)):
        resident_step(k % NCOPY)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count()
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev_start.record(streams[0])
    for k in range(args.steps):
        resident_step(k % NCOPY)
    streams[0].wait_stream(streams[1])
    ev_end.record(streams[0])
    barrier()
    launches = ctx.launch_count() - launches0
    total_ms = float(ev_start.elapsed_time(ev_end))

    # sustained: the same loop for at least --sustain-seconds (not fewer steps than the timed region); every rank must run the
    # same number of steps (the exchange waits for all of them), so the count comes from the slowest rank's timed region
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    n_sus = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / max(total_ms / args.steps, 1e-3))))
    barrier()
    ev_s0, ev_s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_s0.record(streams[0])
    for k in range(n_sus):
        resident_step(k % NCOPY)
        if (k & 63) == 63:
            streams[(k + 1) & 1].synchronize()  # keep the host a bounded number of launches ahead
    streams[0].wait_stream(streams[1])
    ev_s1.record(streams[0])
    barrier()
    sus_ms = float(ev_s0.elapsed_time(ev_s1))
    if world > 1:
        t = torch.tensor([total_ms, single_ms, sus_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, single_ms, sus_ms = (float(v) for v in t.tolist())
    ctx.set_stream(streams[0].cuda_stream)
    pose, cost = bts[0].results()
    for i in range(1, NCOPY):
        pose1, cost1 = bts[i].results()
        assert np.array_equal(pose, pose1) and np.array_equal(cost, cost1), "the resident copies disagree"
    stats = bts[0].stats_ex().astype(np.float64)
    if gathered is not None:  # what the exchange delivered to this rank is what every rank computed
        assert np.array_equal(gathered[rank * B:(rank + 1) * B, :3], pose) and np.array_equal(gathered[rank * B:(rank + 1) * B, 3], cost)

    # ---- parity of the timed batch, every problem of every rank, against the unmodified reference's golden vectors
    gold = wl.golden(rank, world)
    par = parity_stats(pose, cost, gold[0], gold[1]) if gold is not None else parity_stats(np.zeros((0, 3)), np.zeros(0), np.zeros((0, 3)), np.zeros(0))
    if world > 1:
        t = torch.tensor([par["max_abs_dpose"], par["max_rel_dscore"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([par["n_checked"], par["bit_exact_poses"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        par = {"n_checked": int(c[0].item()), "max_abs_dpose": float(t[0].item()), "max_rel_dscore": float(t[1].item()), "bit_exact_poses": int(c[1].item())}

    # ---- e2e arm: host buffers in, host poses out, every step.  Throughput form of the public API:
    # ndtpso_align_submit (stage + H2D + launches) / ndtpso_align_collect (D2H + sync), several batches in
    # flight so the host stages step k+1 while the GPU solves step k.  Every step moves its own inputs
    # host->device and its own poses device->host.  The one-call synchronous form is timed beside it.
    ctx.set_stream(0)
    for _ in range(max(1, args.warmup // 2)):
        ctx.align_batch(pset, conf)
    depth = max(1, args.e2e_depth)

    def e2e_steps(n):
        # `depth` batches in flight: with two, the GPU holds a single batch while the host stages and uploads the next one;
        # with three it always has two batches queued on its two streams (tools/e2e_depth.py)
        tickets, out = collections.deque(), None
        for _ in range(n):
            tickets.append(ctx.align_submit(pset, conf))
            if len(tickets) >= depth:
                out = ctx.align_collect(tickets.popleft())
        while tickets:
            out = ctx.align_collect(tickets.popleft())
        return out

    e2e_steps(2 * depth)  # warm-up of the in-flight form (one set of arenas per batch in flight)
    barrier()
    t0 = time.perf_counter()
    ep, ec = e2e_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.align_batch(pset, conf)
    e2e_sync_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(t[0].item()), float(t[1].item())
    clocks = sampler.stop()
    h2d_bytes, d2h_bytes = ctx.last_transfer_bytes()  # what align_batch really moved over PCIe per step
    assert np.array_equal(ep, pose), "e2e and resident paths disagree"

    fp64_peak = ctx.fp64_peak_tflops()
    for i in range(NCOPY):
        if exs[i] is not None:
            bts[i].attach_exchange(None)
        bts[i].close()

    # ---- configs[1] read literally: ONE scan-match at a time (thread-block-cluster form of the kernel)
    single = None
    if rank == 0 and wl.name == "cfg2":
        one = capi.ProblemSet(pinned_flats[:1])
        b1 = ctx.batch(one, conf)
        ks = []
        for _ in range(args.warmup + 10):
            b1.solve()
            ks.append(float(b1.kernel_times_ms().sum()))
        b1.close()
        for _ in range(3):
            ctx.align_batch(one, conf)
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.align_batch(one, conf)
        lat = (time.perf_counter() - t0) / 20
        single = {"resident_ms": float(np.median(ks[args.warmup:])), "e2e_ms": 1e3 * lat, "e2e_matches_per_s": 1.0 / lat,
                  "note": "batch of 1 through ndtpso_align_batch: host staging + H2D + K0/K1/K2 on a 16-CTA cluster + D2H"}

    # ---- the per-scan callback with the maps resident in HBM: loadLaser -> align -> update
    tracking = None
    if rank == 0 and wl.name == "cfg2" and not args.no_tracking:
        tracking = run_tracking(B, conf, steps=max(4, min(args.steps, 12)), device=local, groups=args.tracking_groups)

    callback = None
    if rank == 0 and wl.name == "cfg2" and not args.no_tracking:
        callback = {"align_default_30x50": run_callback(10, None, None), "swarm_70x50": run_callback(10, 70, 50)}

    rc = 0
    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e = world * B * args.steps / e2e_s
        alg_bytes, alg_flop = wl.alg_bytes(), wl.alg_flop()
        # dominant kernel: K2.  Its launches overlap in the timed region (two streams), so its effective duration per launch
        # is its share of the step time there; the isolated duration (single stream, nothing else running) is given beside it.
        k2_share = float(kt[2]) / float(kt.sum())
        k2_s = k2_share * (total_ms / args.steps) * 1e-3
        k2_iso_s = float(kt[2]) * 1e-3
        ach_gbs = alg_bytes * B / k2_s / 1e9
        ach_tf = alg_flop * B / k2_s / 1e12
        prof = measured_profile(wl.name, B)
        sm_clock = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
        wavefronts = prof.get("shared_wavefronts")
        smem_pipe = None
        if wavefronts:
            n_sm = torch.cuda.get_device_properties(local).multi_processor_count
            smem_pipe = {"resource": "shared-memory data pipe (l1tex__data_pipe_lsu_wavefronts_mem_shared): one wavefront per SM per clock",
                         "wavefronts_per_launch": wavefronts, "peak_wavefronts_per_s": n_sm * sm_clock,
                         "frac": wavefronts / k2_s / (n_sm * sm_clock), "frac_isolated": wavefronts / k2_iso_s / (n_sm * sm_clock),
                         "note": "wavefronts from the committed ncu capture (profiles/traffic.json), duration measured in this run; the kernel's "
                                 "cell lookups (2 + 16 + 8 bytes per lane and evaluation = 6.6 wavefronts per warp) are what fills this pipe"}
        line = {
            "metric": wl.metric, "value": value, "unit": "scan-matches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl.describe(),
                       "batch_per_gpu": B, "particles": P, "iterations": I,
                       "l2": f"inputs larger than L2: four resident copies of the batch (4 x {table_bytes / 1e6:.0f} MB of dense tables; a step touches the built "
                             "flags, the built cells, its compact tables and rand streams) are used round-robin, three other steps between two uses of a copy",
                       "pipelining": "step k runs on copy k % 4 and stream k & 1, so consecutive steps overlap while one drains",
                       "collective": exchange_kind},
            "sustained": {"value": world * B * n_sus / (sus_ms * 1e-3), "unit": "scan-matches/s", "steps": n_sus, "seconds": sus_ms * 1e-3,
                          "note": "the timed loop kept running back to back; clocks are sampled over the timed region, this leg and the e2e arm"},
            "e2e": {"value": e2e, "unit": "scan-matches/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "api": f"ndtpso_align_submit/collect, {depth} batches in flight",
                    "one_call_sync": world * B * args.steps / e2e_sync_s},
            "parity": dict(par, pose_atol=POSE_ATOL, score_rtol=SCORE_RTOL,
                           against=("the unmodified reference's pose and cost_function value for every problem of the timed batch on every rank "
                                    "(tests/golden/batch_vectors.npz, made by tests/golden/make_golden_batch.py)") if gold is not None else
                                   "no committed vectors for this workload: see `live`"),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "pso_sliced_kernel", "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach_gbs / peaks["hbm_gbs"], "traffic": prof.get("dram_bytes"), "peak_source": peak_src,
                         "launch_ms": k2_s * 1e3, "launch_ms_note": "K2's share (%.1f %%) of the step time in the timed region, where launches overlap" % (100 * k2_share),
                         "algorithmic_bytes_per_match": alg_bytes, "algorithmic_flop_per_match": alg_flop,
                         "fp64_achieved_tflops": ach_tf, "fp64_peak_tflops": fp64_peak, "fp64_frac": ach_tf / fp64_peak,
                         "isolated_ms": float(kt[2]), "isolated_achieved_gbs": alg_bytes * B / k2_iso_s / 1e9,
                         "isolated_frac": alg_bytes * B / k2_iso_s / 1e9 / peaks["hbm_gbs"],
                         "smem_pipe_frac": smem_pipe["frac"] if smem_pipe else None,
                         "kernel_ms": {"compact_map": float(kt[0]), "rng_fill": float(kt[1]), "pso": float(kt[2])},
                         "isolated": {"pso_ms": float(kt[2]), "achieved_gbs": alg_bytes * B / k2_iso_s / 1e9,
                                      "fp64_tflops": alg_flop * B / k2_iso_s / 1e12,
                                      "note": "one launch alone on the GPU (single stream)"},
                         "fp64": {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                                  "note": "algorithmic flops (31 per point-evaluation x the cost evaluations a match makes in the reference); peak = DFMA probe on this GPU. The fp32 screen settles most of those evaluations without touching the fp64 pipe, so this is work delivered per second, not pipe utilisation (ncu: profiles/)"},
                         "binding": smem_pipe},
            "single_stream": {"value": world * B * n_single / (single_ms * 1e-3), "unit": "scan-matches/s", "ms_per_step": single_ms / n_single,
                              "note": "one resident batch, one stream, L2 flushed (256 MiB write) between steps, per-step CUDA events"},
            "single_match": single,
            "tracking": tracking,
            "callback": callback,
            "per_match": {"rounds": float(stats[:, 0].mean()), "gbest_updates": float(stats[:, 1].mean()), "fp64_evaluations": float(stats[:, 2].mean()),
                          "settled_by_fp32_screen": float(stats[:, 3].mean())},
            "pose0": [float(v) for v in pose[0]],
        }
        if not args.no_cpu:
            base, specs, cpose, ccost = cpu_reference_rate(wl, args.ref_matches)
            line["cpu_baseline"] = base
            # live parity: the reference (or the oracle port) just solved the batch's own first problems with the batch's seeds
            own = wl.specs(0, world)
            idx = [own.index(sp) for sp in specs if sp in own]
            keep = [i for i, sp in enumerate(specs) if sp in own]
            line["parity"]["live"] = dict(parity_stats(pose[idx], cost[idx], cpose[keep], ccost[keep]),
                                          against=f"oracle/_ref ({base['kind']}) run by this bench's cpu_baseline leg on the same problems and seeds")
        ok = line["parity"]["max_abs_dpose"] <= POSE_ATOL and line["parity"]["max_rel_dscore"] <= SCORE_RTOL
        live = line["parity"].get("live")
        if live:
            ok = ok and live["max_abs_dpose"] <= POSE_ATOL and live["max_rel_dscore"] <= SCORE_RTOL
        line["parity"]["ok"] = bool(ok)
        rc = 0 if ok else 3
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
        if not ok:
            print("bench.py: PARITY FAILURE (see the line's `parity`)", file=sys.stderr)
    if world > 1:
        dist.barrier()
        for e_ in exs:
            if e_ is not None:
                e_.close()
        dist.barrier()
        dist.destroy_process_group()
    return rc


def run_single_process_arm(args):
    """`--single-process`: all N GPUs from ONE process through ndtpso_multi_* (include/ndtpso_b200.h) — the form a C/C++ caller
    uses.  value: the shards resident in HBM, one stream per device, the fused exchange delivering every result to every device,
    wall clock around K steps (each step ends with every device synchronised); e2e: ndtpso_align_submit_multi / _collect_multi with
    pinned host buffers, three batches in flight."""
    import torch

    from ndtpso_slam_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    G = args.gpus
    wl = Workload(args.workload, args.batch)
    B, P, I = wl.B, wl.P, wl.I
    flats = [f for r in range(G) for f in wl.problems(r, G)]  # shard r of the multi object = rank r's problems
    for f in flats:
        for k in ("points", "mean", "inv_cov", "built"):
            t = torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory()
            f["_keep_" + k], f[k] = t, t.numpy()
    pset = capi.ProblemSet(flats)
    conf = capi.PsoConfig.make(population=P, iterations=I)
    m = capi.Multi(list(range(G)))
    bt = m.batch(pset, conf)
    for _ in range(args.warmup):
        bt.solve()
    pose, cost = bt.results()
    sampler = ClockSampler(0)
    sampler.start()
    n0 = m.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bt.solve()
    bt.results()
    wall = time.perf_counter() - t0
    launches = m.launch_count() - n0
    exchanged = bt.device_results_ptr(0) is not None
    bt.close()
    depth = max(1, args.e2e_depth)

    def e2e_steps(n):
        tickets, out = collections.deque(), None
        for _ in range(n):
            tickets.append(m.align_submit(pset, conf))
            if len(tickets) >= depth:
                out = m.align_collect(tickets.popleft())
        while tickets:
            out = m.align_collect(tickets.popleft())
        return out

    e2e_steps(2 * depth)
    t0 = time.perf_counter()
    ep, ec = e2e_steps(args.steps)
    e2e_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.align_batch(pset, conf)
    sync_s = time.perf_counter() - t0
    clocks = sampler.stop()
    assert np.array_equal(ep, pose) and np.array_equal(ec, cost), "e2e and resident paths disagree"
    gold = [wl.golden(r, G) for r in range(G)]
    par = parity_stats(pose, cost, np.concatenate([g[0] for g in gold]), np.concatenate([g[1] for g in gold])) if all(g is not None for g in gold) \
        else parity_stats(np.zeros((0, 3)), np.zeros(0), np.zeros((0, 3)), np.zeros(0))
    par["ok"] = bool(par["max_abs_dpose"] <= POSE_ATOL and par["max_rel_dscore"] <= SCORE_RTOL)
    h2d = sum(f["points"].nbytes for f in flats)  # lower bound: the scans; the built cells' rows come on top (see the per-process line)
    line = {"metric": wl.metric, "value": G * B * args.steps / wall, "unit": "scan-matches/s", "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.describe(), "batch_per_gpu": B, "particles": P, "iterations": I,
                       "process_model": f"single process, ndtpso_multi_* over {G} devices (one context, host thread and staging pool per device)",
                       "collective": "fused device-side exchange between the shards (ndtpso_exchange_connect_local)" if exchanged else "host gather",
                       "l2": "one resident copy per device, single stream (the per-process arm cycles four copies on two streams)"},
            "e2e": {"value": G * B * args.steps / e2e_s, "unit": "scan-matches/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(G * B * 32),
                    "api": f"ndtpso_align_submit_multi/_collect_multi, {depth} batches in flight", "one_call_sync": G * B * args.steps / sync_s},
            "parity": par, "gpu_launches": int(launches), "clocks": clocks, "pose0": [float(v) for v in pose[0]]}
    print(json.dumps(line))
    m.close()
    return 0 if par["ok"] else 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"], help="cfg2: BASELINE.json's headline (configs[1..3]); cfg5: configs[4], the multi-resolution sweep")
    ap.add_argument("--batch", type=int, default=0, help="scan-match problems per GPU per step (default 256; 148 for cfg5)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-matches", type=int, default=0, help="CPU arm: matches per host worker per step (default 16 for cfg2, 2 for cfg5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and the live parity check it feeds)")
    ap.add_argument("--no-tracking", action="store_true", help="skip the device-resident tracking leg")
    ap.add_argument("--tracking-groups", type=int, default=3, help="tracking leg: independent groups of robots served by their own host thread and stream")
    ap.add_argument("--e2e-depth", type=int, default=3, help="e2e arm: batches kept in flight through ndtpso_align_submit/collect")
    ap.add_argument("--sustain-seconds", type=float, default=2.2, help="length of the sustained leg")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="N > 1: how the solved poses reach every rank")
    ap.add_argument("--single-process", action="store_true", help="drive all --gpus devices from this one process through ndtpso_multi_* (not under torchrun)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.batch <= 0:
        args.batch = 256 if args.workload == "cfg2" else 148
    if args.ref_matches <= 0:
        args.ref_matches = 16 if args.workload == "cfg2" else 2
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.single_process:
        return run_single_process_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
