"""configs[4] of BASELINE.json: multi-resolution sweep, cell_size in {0.25, 0.5, 1.0, 2.0} m, 200 particles x 100 iterations.
usage: python tools/cfg5_time.py [frames_per_cell_size] [points_per_thread] [screen]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, frames, synthetic as syn
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 37
npt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scr = int(sys.argv[3]) if len(sys.argv) > 3 else -1
conf = capi.PsoConfig.make(population=200, iterations=100)
ctx = capi.Context(0)
ctx.set_option(capi.OPT_CLUSTER, 1)
ctx.set_option(capi.OPT_POINTS_PER_THREAD, npt)
ctx.set_option(capi.OPT_SCREEN, scr)
sizes = (0.25, 0.5, 1.0, 2.0)
per = {cs: [frames.problem_from_scans(syn.trajectory_problem(syn.CFG5[cs], b), seed=1 + b) for b in range(nf)] for cs in sizes}


def run(flats, label):
    bt = ctx.batch(flats, conf)
    ts = []
    for _ in range(4):
        bt.solve(); ts.append(bt.kernel_times_ms())
    pose, cost = bt.results(); st = bt.stats_ex().astype(float)
    t = min(x.sum() for x in ts)
    evals = 20201 * 1081 * len(flats)
    print(f"{label:26s} {len(flats):4d} problems: {t:7.3f} ms -> {len(flats) / t * 1e3:8.0f} matches/s, {evals / t / 1e6:6.1f} G point-evals/s (nominal); K2 {ts[-1][2]:.3f} ms; "
          f"rounds {st[:, 0].mean():.1f}, fp64 evals {st[:, 2].mean():.0f}, screened {st[:, 3].sum() / max(st[:, 2].sum() + st[:, 3].sum(), 1):.1%}", flush=True)
    bt.close()
    return pose, cost


print(f"points per thread {npt or 'auto'}, screen {scr}")
for cs in sizes:
    run(per[cs], f"cell {cs} m")
run([f for b in range(nf) for cs in sizes for f in (per[cs][b],)], "sweep (4 cell sizes)")
