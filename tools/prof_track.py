"""Small driver for ncu: B robots tracked with device-resident maps for a few steps.  No torch.
usage: python tools/prof_track.py [robots] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, dframes, synthetic as syn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = syn.CFG2
s, S = cfg.sensor, cfg.map_size_m
room = syn.Room(S)
ctx = capi.Context(0)
df = dframes.DeviceFrames(ctx, B, S, S, cfg.cell_side, s.beams, max_cells=1024)
sets = [syn.trajectory_problem(cfg, b) for b in range(B)]
for k in range(5):
    df.load_laser(np.stack([ss.map_scans[k][1] for ss in sets]), s.angle_min, s.angle_increment, s.range_max)
    df.update(np.array([ss.map_scans[k][0] for ss in sets]))
conf = capi.PsoConfig.make(population=70, iterations=50)
init = np.array([ss.guess for ss in sets])
for k in range(steps + 1):
    scans = np.stack([syn.make_scan(room, s, (ss.true_pose[0] + 0.02 * k, ss.true_pose[1] + 0.005 * k, ss.true_pose[2] + 0.001 * k),
                                    syn.NoiseLCG(777 + 131 * b + k)) for b, ss in enumerate(sets)])
    pose, cost = df.track_step(scans, s.angle_min, s.angle_increment, s.range_max, initial_poses=init if k == 0 else None, conf=conf)
    print(k, "kernel ms:", df.kernel_times_ms(), "pose0", pose[0])
