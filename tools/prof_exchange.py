"""Driver for ncu on the fused result exchange: one process, one context per device, equal resident shards of cfg2 problems;
every PSO kernel stores its results into the gathered buffer of every device (peer stores over NVLink) and raises the arrival
flags, exchange_wait_kernel waits for them.
usage: python tools/prof_exchange.py [devices] [problems per device] [solves]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ndtpso_slam_b200 import capi, workload  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
solves = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ndev = capi.load_library().ndtpso_device_count()
devices = list(range(G)) if ndev >= G else [0] * G
m = capi.Multi(devices)
bt = m.batch(workload.cfg2_batch(per * G), capi.PsoConfig.make(population=70, iterations=50))
for _ in range(solves):
    bt.solve()
    pose, cost = bt.results()
print("devices", devices, "problems", per * G, "pose0", pose[0], "cost0", cost[0], "exchanged on device:", bt.device_results_ptr(0) is not None)
bt.close()
m.close()
