"""K2 time of small batches: the cluster size the library picks on its own, and forced sizes beside it (cfg2, 70 x 50).
usage: python tools/cluster_sizes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ndtpso_slam_b200 import capi, workload  # noqa: E402

conf = capi.PsoConfig.make(population=70, iterations=50)
ref = {}
for n in (1, 4, 16, 64, 148):
    flats = workload.cfg2_batch(n)
    for cl in (0, 16, 8, 4, 2, 1):  # 0 = auto
        if cl > 1 and n * cl > 148:
            continue
        ctx = capi.Context(0)
        ctx.set_option(capi.OPT_CLUSTER, cl)
        try:
            bt = ctx.batch(flats, conf)
            ts = []
            for _ in range(8):
                bt.solve()
                ts.append(bt.kernel_times_ms()[2])
            pose, cost = bt.results()
        except capi.NdtpsoError as e:
            print(f"n={n:4d} cluster {cl:2d}: {e}")
            continue
        ref.setdefault(n, pose)
        print(f"n={n:4d} cluster {'auto' if cl == 0 else cl:>4}: K2 {np.median(ts[2:]):.3f} ms  rounds {bt.stats()[:, 0].mean():.1f}  poses equal to the first form: {np.array_equal(pose, ref[n])}", flush=True)
        bt.close()
        ctx.close()
