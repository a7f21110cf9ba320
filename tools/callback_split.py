"""Where one scan of the reference node's callback goes, through the drop-in NDTFrame with its map in HBM (one robot, cfg2).
usage: python tools/callback_split.py [population iterations]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ndtpso_slam_b200 import capi, frames, synthetic as syn  # noqa: E402

pop = int(sys.argv[1]) if len(sys.argv) > 2 else None
its = int(sys.argv[2]) if len(sys.argv) > 2 else None
conf = capi.PsoConfig.make(population=pop, iterations=its) if pop else None
cfg = syn.CFG2
s, S = cfg.sensor, cfg.map_size_m
room = syn.Room(S)
ss = syn.trajectory_problem(cfg, 0)
steps = 40
scans = [r for _, r in ss.map_scans] + [syn.make_scan(room, s, (ss.true_pose[0] + 0.02 * k, ss.true_pose[1] + 0.005 * k, ss.true_pose[2] + 0.001 * k),
                                                      syn.NoiseLCG(4242 + k)) for k in range(steps)]
init = tuple(ss.map_scans[0][0])
C.CDLL(None).srand(1)
ref = frames.Frame(width=S, height=S, cell_side=cfg.cell_side, calculate_cells_params=True)
pose = np.array(init)
rows = []
for k, ranges in enumerate(scans):
    t0 = time.perf_counter()
    cur = frames.Frame(trans=init, width=S, height=S, cell_side=cfg.cell_side if k == 0 else float(S), calculate_cells_params=False)
    t1 = time.perf_counter()
    cur.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    t2 = time.perf_counter()
    if k > 0:
        pose = ref.align(pose, cur, conf)
    t3 = time.perf_counter()
    ref.update(pose, cur)
    t4 = time.perf_counter()
    cur.close()
    t5 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0))
r = 1e3 * np.median(np.array(rows[len(ss.map_scans) + 5:]), axis=0)
n = capi.load_library().ndtpso_rand_draws(C.byref(conf)) if conf else 3 + 3 * 30 + 6 * 30 * 50
t0 = time.perf_counter()
buf = (C.c_int32 * n)()
frames.load_library().ndtpso_frame_draw_rand(buf, n)
t_draw = 1e3 * (time.perf_counter() - t0)
print(f"swarm {pop or 30} x {its or 50}, device-resident map {bool(ref.device_resident)}: per scan {r[5]:.3f} ms = new frame {r[0]:.3f} + loadLaser {r[1]:.3f} + align {r[2]:.3f} "
      f"+ update {r[3]:.3f} + close {r[4]:.3f};  drawing the {n} rand() numbers of one align: {t_draw:.3f} ms;  pose {pose}")
