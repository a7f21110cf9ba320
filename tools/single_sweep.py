"""Single-match latency vs speculation window / cluster size. usage: python tools/single_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, workload
flats = workload.cfg2_batch(1)
conf = capi.PsoConfig.make(population=70, iterations=50)
ref = None
for cl in (16, 8):
    for chunk in (0, 12 | (1 << 16), 24 | (1 << 16), 32 | (1 << 16), 16 | (2 << 16)):
        ctx = capi.Context(0)
        ctx.set_option(capi.OPT_CLUSTER, cl)
        ctx.set_option(capi.OPT_HOT_CHUNK, chunk)
        bt = ctx.batch(flats, conf)
        ts = []
        for _ in range(12):
            bt.solve(); ts.append(bt.kernel_times_ms()[2])
        pose, cost = bt.results(); st = bt.stats()
        if ref is None: ref = pose
        print(f"cluster {cl} window {chunk & 0xffff} thresh {chunk >> 16}: pso {np.median(ts[2:]):.4f} ms rounds {st[0,0]} equal {np.array_equal(pose, ref)}")
        bt.close(); ctx.close()
