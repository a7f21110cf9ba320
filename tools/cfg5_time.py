"""configs[4] of BASELINE.json: multi-resolution sweep, cell_size in {0.25, 0.5, 1.0, 2.0} m, 200 particles x 100 iterations.
usage: python tools/cfg5_time.py [frames_per_cell_size]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, frames, synthetic as syn
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 16
flats = []
for cs in (0.25, 0.5, 1.0, 2.0):
    for b in range(nf):
        flats.append(frames.problem_from_scans(syn.trajectory_problem(syn.CFG5[cs], b), seed=1 + b))
conf = capi.PsoConfig.make(population=200, iterations=100)
ctx = capi.Context(0)
bt = ctx.batch(flats, conf)
ts = []
for _ in range(4):
    bt.solve(); ts.append(bt.kernel_times_ms())
pose, cost = bt.results(); st = bt.stats()
t = min(x.sum() for x in ts)
evals = 20201 * 1081 * len(flats)
print(f"{len(flats)} problems (4 cell sizes x {nf}), 200x100: {t:.3f} ms -> {len(flats)/t*1e3:.0f} matches/s, {evals/t/1e6:.1f} G point-evals/s (nominal), kernel ms {ts[-1]}, rounds {st[:,0].mean():.1f}, gbest updates {st[:,1].mean():.1f}")
for cs_i, cs in enumerate((0.25, 0.5, 1.0, 2.0)):
    print(f"  cell {cs}: pose[0] {pose[cs_i*nf]} cost {cost[cs_i*nf]:.6f}")
