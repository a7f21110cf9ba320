"""Speculation-window sweep of the sliced PSO kernel. usage: python tools/chunk_sweep.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(batch)
conf = capi.PsoConfig.make(population=70, iterations=50)
ref = None
for chunk, th in ((0, 1), (12, 1), (16, 1), (24, 1), (32, 1), (24, 2), (16, 2), (12, 1)):
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_HOT_CHUNK, chunk | (th << 16))
    bt = ctx.batch(flats, conf)
    ts = []
    for _ in range(5):
        bt.solve(); ts.append(bt.kernel_times_ms()[2])
    pose, cost = bt.results()
    st = bt.stats()
    if ref is None: ref = (pose, cost)
    print(f"B={batch} hot_chunk={chunk} thresh={th}: pso {min(ts):.3f} ms -> {batch/min(ts)*1e3:.0f} matches/s  rounds {st[:,0].mean():.1f} gbest updates {st[:,1].mean():.1f}  "
          f"pose bit-equal {int((pose == ref[0]).all(axis=1).sum())}/{batch} max|dcost| rel {abs((cost-ref[1])/ref[1]).max():.1e}")
    bt.close(); ctx.close()
