"""fp32 screen on / off: results must be bit-identical; timing and how many evaluations the screen settles.
usage: python tools/screen_ab.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, workload
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(batch)
conf = capi.PsoConfig.make(population=70, iterations=50)
res = {}
for scr in (0, 1, 0, 1):
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_SCREEN, scr)
    bt = ctx.batch(flats, conf)
    ts = []
    for _ in range(5):
        bt.solve(); ts.append(bt.kernel_times_ms()[2])
    pose, cost = bt.results(); st = bt.stats_ex()
    res[scr] = (pose, cost)
    print(f"screen={scr}: pso {min(ts):.3f} ms -> {batch/min(ts)*1e3:.0f} matches/s; rounds {st[:,0].mean():.1f}, fp64 evals {st[:,2].mean():.0f}, screened {st[:,3].mean():.0f} "
          f"({100*st[:,3].sum()/max(1,(st[:,2].sum()+st[:,3].sum())):.1f} %)")
    bt.close(); ctx.close()
print("bit-identical poses:", np.array_equal(res[0][0], res[1][0]), " costs:", np.array_equal(res[0][1], res[1][1]),
      " max |dpose|", np.abs(res[0][0] - res[1][0]).max())
