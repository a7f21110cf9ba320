#!/usr/bin/env python
"""bench.py — scan-matches/sec of the PSO/NDT hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

A step = one pass of the hot path (pso_optimization, 70 particles x 50 iterations, 1081-beam scan
vs 50 m / 0.5 m NDT map) over a batch of B independent scan-match problems per GPU, each carrying
its own dense (mu, Sigma^-1, built) table.

  value      whole-job matches/s with the batch resident in HBM (dense tables, points, guesses): every step runs
             K0 table compaction + K1 rand() stream + K2 PSO (+ the result exchange for N > 1); two resident copies
             of the batch alternate on two streams so that consecutive steps overlap; CUDA events around the K steps
  single_stream  the same with one copy on one stream, L2 flushed between steps (per-step events)
  e2e        the same metric through ndtpso_align_submit/collect with HOST buffers (pinned): H2D of every
             input + kernels + D2H of the poses inside the timed region, three batches in flight (--e2e-depth)
  tracking   the reference's whole per-scan callback (loadLaser -> align -> update) on device-resident maps
  roofline   dominant kernel (pso_sliced_kernel): algorithmic bytes / its duration vs the measured HBM peak;
             the fp64 pipe fraction beside it (the bound that really binds)
  cpu_baseline  the reference's own CPU path (oracle/_ref if present, else the oracle port) on
             this box's host cores, bounded sample

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

METRIC = "scan-matches/sec (1081-pt scan, 70 particles x 50 iters)"
P, I = 70, 50
ALG_BYTES_PER_MATCH = 16 * 1081 + 49 * 10000 + 80          # 507 376 B
ALG_FLOP_PER_MATCH = 31 * (P + 1 + P * I) * 1081 + 17 * P * I  # ~119.7 MFLOP


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def make_problems(batch, rank):
    """`batch` cfg2-shaped problems for this rank, every one with its own table arrays."""
    from ndtpso_slam_b200 import workload
    return workload.cfg2_batch(batch, first=rank * batch)


def workload_name(batch):
    """config.workload of both arms (the reference arm runs a bounded sample of the same workload)."""
    return (f"cfg2 shape (configs[1]; batched as configs[2]): {batch} independent 1081-beam scan-matches per GPU vs "
            "50 m/0.5 m NDT maps (one dense table per problem), 70 particles x 50 iterations")


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
    kind, scan_args, seeds = args
    from ndtpso_slam_b200 import synthetic as syn
    from oracle import binding
    ss = syn.trajectory_problem(syn.CFG2, scan_args)
    t0 = time.perf_counter()
    if kind == "reference":
        R = binding.Reference()
        rf, q = R.build_problem(ss)
        rf.build()
        t0 = time.perf_counter()
        for s in seeds:
            R.pso(rf, q, ss.guess, ss.deviation, P, I, seed=s, num_threads=1)
    else:
        from ndtpso_slam_b200 import workload
        flat = workload.cfg2_batch(1, first=scan_args)[0]
        O = binding.Oracle()
        t0 = time.perf_counter()
        for s in seeds:
            O.pso(flat, flat["guess"], flat["deviation"], P, I, seed=s)
    return time.perf_counter() - t0


def cpu_reference_rate(matches_per_worker):
    """matches/s of the reference CPU path using every host core: nproc single-thread workers over
    disjoint problems (its deterministic mode; in-process outer parallelism is impossible because
    the reference draws from the process-global rand()).  Also times the as-shipped OpenMP mode."""
    from oracle import binding
    kind = "reference" if os.path.exists(binding.REF_SO) else "port"
    if kind == "port":
        binding.build()
    cores = os.cpu_count() or 1
    jobs = [(kind, w, list(range(1 + w * 1000, 1 + w * 1000 + matches_per_worker))) for w in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    # wall includes process start-up and map building; use the slowest worker's solve time instead
    with mp.get_context("spawn").Pool(cores) as pool:
        per = pool.map(_cpu_worker, jobs)
    rate = cores * matches_per_worker / max(per)
    out = {"value": rate, "unit": "scan-matches/s", "cores": cores, "kind": kind,
           "sample": f"{cores} single-thread workers x {matches_per_worker} cfg2 matches each (70x50, 1081 pts), slowest worker {max(per):.2f} s",
           "single_thread_ms_per_match": 1e3 * float(np.median(per)) / matches_per_worker}
    if kind == "reference":
        from ndtpso_slam_b200 import synthetic as syn
        R = binding.Reference()
        ss = syn.trajectory_problem(syn.CFG2, 0)
        rf, q = R.build_problem(ss)
        ts = [R.pso(rf, q, ss.guess, ss.deviation, P, I, seed=s, num_threads=-1)[1] for s in range(1, 9)]
        out["as_shipped_openmp"] = {"value": 1.0 / float(np.median(ts[2:])), "unit": "scan-matches/s", "threads": R.lib.ref_omp_max_threads(),
                                    "note": "num_threads=-1 inside one match; non-deterministic result (SURVEY.md section 0.5)"}
    return out, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = max(1, args.ref_matches)
    rates = []
    base = None
    t_all = time.perf_counter()
    for step in range(args.warmup + args.steps):
        base, _ = cpu_reference_rate(per_step)
        if step >= args.warmup:
            rates.append(base["value"])
        if time.perf_counter() - t_all > 240:
            break
    value = float(np.mean(rates)) if rates else base["value"]
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scan-matches/s", "n_gpus": args.gpus, "steps": len(rates),
            "warmup": args.warmup, "ms_per_step": 1e3 * base["cores"] * per_step / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "batch_per_gpu": args.batch, "particles": P, "iterations": I,
                       "sample": f"each step: {base['cores']} x {per_step} matches of that workload (its first trajectory problems, own map "
                                 "each) on the host cores, one single-thread worker per core"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "scan-matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def measured_traffic(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of pso_sliced_kernel per launch, from the committed
    `ncu --set full` capture of this workload (profiles/); None for a batch size that was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(str(batch))
    except Exception:
        return None


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
        df = dframes.DeviceFrames(c, hi - lo, S, S, cfg.cell_side, s.beams, max_cells=1024, flags=dframes.DF_NO_CLUSTER if groups > 1 else 0)
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
    B = args.batch

    flats = make_problems(B, rank)
    ctx = capi.Context(local)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream shared by torch events, NCCL and the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    conf = capi.PsoConfig.make(population=P, iterations=I)

    # pinned host copies of every input array (the e2e path copies from these every step)
    pinned_flats, h2d_bytes = [], 0
    for f in flats:
        g = dict(f)
        for k in ("points", "mean", "inv_cov", "built"):
            t = torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory()
            g[k] = t.numpy()
            g["_keep_" + k] = t
            h2d_bytes += t.numel() * t.element_size()
        pinned_flats.append(g)
    pset = capi.ProblemSet(pinned_flats)
    h2d_bytes += B * (48 + 4)  # guess, deviation, seed
    d2h_bytes = B * 32

    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: value.  Resident copies of the batch alternate on two streams (step k -> copy k % 4, stream k & 1): a
    # batch of 256 CTAs leaves 40 of the 296 CTA slots empty and drains unevenly, and with the next step on the other
    # stream its CTAs take the free slots at once.  Every step is a full pass of the hot path over one batch (K0 table
    # compaction, K1 rand() streams, K2 PSO, result exchange for N > 1); four copies are cycled so that what a step
    # touches (~41 MB) has been pushed out of the 126 MB L2 by the three steps in between.  The single-stream, L2-flushed form is measured beside it (`single_stream`), and the
    # roofline uses that form's kernel durations.
    NCOPY = 4  # resident copies of the batch, used round-robin: 4 x ~41 MB touched per step (compact tables, rand streams, flags) > 126 MB of L2
    streams = [stream, torch.cuda.Stream()]
    bts = [ctx.batch(pset, conf) for _ in range(NCOPY)]
    bt = bts[0]
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

    if ex is not None:  # once: the fused exchange delivers exactly what the collective delivers
        resident_step(0)
        class _ExG:
            __cuda_array_interface__ = {"shape": (world * B * 4,), "typestr": "<f8", "data": (ex.device_results_ptr(), False), "version": 3}
        torch.cuda.synchronize()
        fused = torch.as_tensor(_ExG(), device="cuda").view(world * B, 4).clone()
        ref_rows = sharding.gather_results(res_ts[0].view(B, 4), world * B, world, rank)
        torch.cuda.synchronize()
        assert torch.equal(fused, ref_rows), "fused exchange and NCCL all-gather disagree"

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

    # the timed region: K steps alternating between the two copies / streams
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
    if world > 1:
        t = torch.tensor([total_ms, single_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, single_ms = float(t[0].item()), float(t[1].item())
    ctx.set_stream(streams[0].cuda_stream)
    pose, cost = bts[0].results()
    for i in range(1, NCOPY):
        pose1, cost1 = bts[i].results()
        assert np.array_equal(pose, pose1) and np.array_equal(cost, cost1), "the resident copies disagree"
    stats = bts[0].stats()

    # ---- e2e arm: host buffers in, host poses out, every step.  Throughput form of the public API:
    # ndtpso_align_submit (stage + H2D + launches) / ndtpso_align_collect (D2H + sync), several batches in
    # flight so the host stages step k+1 while the GPU solves step k.  Every step moves its own inputs
    # host->device and its own poses device->host.  The one-call synchronous form is timed beside it.
    ctx.set_stream(0)
    for _ in range(max(1, args.warmup // 2)):
        ctx.align_batch(pset, conf)
    depth = max(1, args.e2e_depth)

    def e2e_steps(n):
        # `depth` batches in flight: with two, the GPU holds a single batch while the host stages and uploads the next one
        # (0.6 + 0.3 ms of every 2.2 ms step: 117 k matches/s); with three it always has two batches queued on its two
        # streams and the end-to-end rate equals the resident one (tools/e2e_depth.py)
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
    if rank == 0:
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
    if rank == 0 and not args.no_tracking:
        tracking = run_tracking(B, conf, steps=max(4, min(args.steps, 12)), device=local, groups=args.tracking_groups)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e = world * B * args.steps / e2e_s
        # dominant kernel: K2.  Its launches overlap in the timed region (two streams), so its effective duration per launch
        # is its share of the step time there; the isolated duration (single stream, nothing else running) is given beside it.
        k2_share = float(kt[2]) / float(kt.sum())
        k2_s = k2_share * (total_ms / args.steps) * 1e-3
        k2_iso_s = float(kt[2]) * 1e-3
        ach_gbs = ALG_BYTES_PER_MATCH * B / k2_s / 1e9
        ach_tf = ALG_FLOP_PER_MATCH * B / k2_s / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "scan-matches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "batch_per_gpu": B, "particles": P, "iterations": I,
                       "l2": "inputs larger than L2: four resident copies of the batch (4 x 130 MB of dense tables, of which a step touches ~41 MB: "
                             "built flags, built cells, compact tables, rand streams) are used round-robin, 164 MB between two uses of a copy",
                       "pipelining": "step k runs on copy k % 4 and stream k & 1, so consecutive steps overlap while one drains",
                       "collective": exchange_kind},
            "e2e": {"value": e2e, "unit": "scan-matches/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "api": f"ndtpso_align_submit/collect, {depth} batches in flight",
                    "one_call_sync": world * B * args.steps / e2e_sync_s},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "pso_sliced_kernel", "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach_gbs / peaks["hbm_gbs"], "traffic": measured_traffic(B), "peak_source": peak_src,
                         "launch_ms": k2_s * 1e3, "launch_ms_note": "K2's share (%.1f %%) of the step time in the timed region, where launches overlap" % (100 * k2_share),
                         "kernel_ms": {"compact_map": float(kt[0]), "rng_fill": float(kt[1]), "pso": float(kt[2])},
                         "isolated": {"pso_ms": float(kt[2]), "achieved_gbs": ALG_BYTES_PER_MATCH * B / k2_iso_s / 1e9,
                                      "fp64_tflops": ALG_FLOP_PER_MATCH * B / k2_iso_s / 1e12,
                                      "note": "one launch alone on the GPU (single stream): 256 CTAs fill 86 % of the 296 CTA slots"},
                         "fp64": {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                                  "note": "algorithmic flops (31 per point-evaluation x the 3571 cost evaluations a match makes in the reference); peak = DFMA probe on this GPU. The fp32 screen settles ~86 % of those evaluations without touching the fp64 pipe, so this is work delivered per second, not pipe utilisation (ncu: profiles/)"}},
            "single_stream": {"value": world * B * n_single / (single_ms * 1e-3), "unit": "scan-matches/s", "ms_per_step": single_ms / n_single,
                              "note": "one resident batch, one stream, L2 flushed (256 MiB write) between steps, per-step CUDA events"},
            "single_match": single,
            "tracking": tracking,
            "rounds_per_match": float(stats[:, 0].mean()), "pose0": [float(v) for v in pose[0]],
        }
        if not args.no_cpu:
            line["cpu_baseline"], _ = cpu_reference_rate(args.ref_matches)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        for e_ in exs:
            if e_ is not None:
                e_.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="scan-match problems per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-matches", type=int, default=4, help="CPU arm: matches per host worker per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tracking", action="store_true", help="skip the device-resident tracking leg")
    ap.add_argument("--tracking-groups", type=int, default=3, help="tracking leg: independent groups of robots served by their own host thread and stream")
    ap.add_argument("--e2e-depth", type=int, default=3, help="e2e arm: batches kept in flight through ndtpso_align_submit/collect")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="N > 1: how the solved poses reach every rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
