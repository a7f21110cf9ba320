"""Small driver for ncu: the configs[4] sweep (4 cell sizes x 37 frames, 200 particles x 100 iterations) resident in HBM, a few solves."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi  # noqa: E402
import bench  # noqa: E402

wl = bench.Workload("cfg5", 148)
ctx = capi.Context(0)
bt = ctx.batch(wl.problems(0, 1), capi.PsoConfig.make(population=wl.P, iterations=wl.I))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    bt.solve()
    print("kernel ms (K0, K1, K2):", bt.kernel_times_ms())
pose, cost = bt.results()
print("pose0", pose[0], "cost0", cost[0])
