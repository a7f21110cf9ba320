"""Time the PSO kernel of an alternative build of the library. usage: python tools/var_time.py <lib.so> [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload
if len(sys.argv) > 1 and sys.argv[1] != "-":
    capi._build.LIB_PATH = os.path.abspath(sys.argv[1])
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = capi.Context(0)
bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
ts = []
for _ in range(6):
    bt.solve(); ts.append(bt.kernel_times_ms()[2])
pose, cost = bt.results()
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'default'}: pso {min(ts):.3f} ms (median {sorted(ts)[len(ts)//2]:.3f}) -> {batch/min(ts)*1e3:.0f} matches/s  checksum {pose.sum():.15g} {cost.sum():.15g}")
