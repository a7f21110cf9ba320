"""Sweep kernel launch shapes on the GPU. usage: python tools/sweep.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(batch)
conf = capi.PsoConfig.make(population=70, iterations=50)
ref = None
cfgs = [(2, 3, 2, 0, 1), (1, 0, 0, 8, 1)]
if batch <= 74:
    cfgs = [(2, 0, 0, 0, cl) for cl in (1, 2, 4, 8, 16)]
for kernel, npt, jb, warps, cl in cfgs:
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_CLUSTER, cl)
    ctx.set_option(capi.OPT_KERNEL, kernel); ctx.set_option(capi.OPT_POINTS_PER_THREAD, npt); ctx.set_option(capi.OPT_WARPS_PER_CTA, warps); ctx.set_option(capi.OPT_CANDIDATE_BATCH, jb)
    bt = ctx.batch(flats, conf)
    ts = []
    for _ in range(4):
        bt.solve(); ts.append(bt.kernel_times_ms()[2])
    pose, cost = bt.results()
    if ref is None: ref = pose
    print(f"B={batch} kernel={kernel} cluster={cl} npt={npt} jb={jb} warps={warps}: pso {min(ts):.3f} ms -> {batch/min(ts)*1e3:.0f} matches/s  max|dpose| vs first {abs(pose-ref).max():.2e}")
    bt.close(); ctx.close()
