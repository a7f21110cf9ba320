"""Tracking step time. usage: python tools/track_time.py [robots] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, dframes, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = syn.CFG2; s, S = cfg.sensor, cfg.map_size_m; room = syn.Room(S)
ctx = capi.Context(0)
df = dframes.DeviceFrames(ctx, B, S, S, cfg.cell_side, s.beams, max_cells=1024)
sets = [syn.trajectory_problem(cfg, b) for b in range(B)]
for k in range(5):
    df.load_laser(np.stack([ss.map_scans[k][1] for ss in sets]), s.angle_min, s.angle_increment, s.range_max)
    df.update(np.array([ss.map_scans[k][0] for ss in sets]))
conf = capi.PsoConfig.make(population=70, iterations=50)
scans = [np.stack([syn.make_scan(room, s, (ss.true_pose[0] + 0.02 * k, ss.true_pose[1] + 0.005 * k, ss.true_pose[2] + 0.001 * k),
                                 syn.NoiseLCG(777 + 131 * b + k)) for b, ss in enumerate(sets)]) for k in range(steps + 2)]
init = np.array([ss.guess for ss in sets])
df.track_step(scans[0], s.angle_min, s.angle_increment, s.range_max, initial_poses=init, conf=conf)
df.track_step(scans[1], s.angle_min, s.angle_increment, s.range_max, conf=conf)
t0 = time.perf_counter()
for k in range(steps):
    pose, cost = df.track_step(scans[2 + k], s.angle_min, s.angle_increment, s.range_max, conf=conf)
dt = (time.perf_counter() - t0) / steps * 1e3
kt = df.kernel_times_ms()
st = df.pso_stats()
print("pso stats: rounds %.1f fp64 evals %.0f screened %.0f (%.1f %%)" % (st[:,0].mean(), st[:,2].mean(), st[:,3].mean(), 100*st[:,3].sum()/max(1, st[:,2].sum()+st[:,3].sum())))
print(f"{dt:.3f} ms/step -> {B/dt*1e3:.0f} scans/s; kernels {sum(kt.values()):.3f} ms {kt}; checksum {pose.sum():.12f}")
