"""K2 time of one resident batch for every launch shape of the one-CTA-per-problem kernel (points per thread x candidate batch).
usage: python tools/shape_sweep.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(batch)
conf = capi.PsoConfig.make(population=70, iterations=50)
ref = None
for npt in (0, 2, 3, 4, 5, 6):
    for jb in (0, 2):
        ctx = capi.Context(0)
        ctx.set_option(capi.OPT_CLUSTER, 1)
        if npt:
            ctx.set_option(capi.OPT_POINTS_PER_THREAD, npt)
        if jb:
            ctx.set_option(capi.OPT_CANDIDATE_BATCH, jb)
        bt = ctx.batch(flats, conf)
        ts = []
        for _ in range(4):
            bt.solve()
            ts.append(bt.kernel_times_ms()[2])
        pose, cost = bt.results()
        if ref is None:
            ref = (pose.copy(), cost.copy())
        same = bool((pose == ref[0]).all() and (cost == ref[1]).all())
        print(f"npt={npt} jb={jb}: K2 {min(ts):.3f} ms  ({batch / min(ts):.1f} k matches/s single launch)  identical={same}", flush=True)
        bt.close()
        ctx.close()
