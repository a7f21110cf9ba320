"""Generates tests/golden/map_vectors.npz from the UNMODIFIED reference (oracle/_ref): golden vectors for
the map-building rows (NDTFrame::loadLaser / update / build) and for the per-scan
callback (loadLaser -> align -> update).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_maps.py

  mapbuild/*   24 scans of the cfg1 sensor merged into a 20 m / 1.0 m map along a trajectory: the float32
               ranges, the poses, the scan points NDTFrame::loadLaser made of them, and the reference's
               (mean, Sigma^-1, built) table after build() at steps 11 and 23 (cells overflow their 50-point
               slots in between, so the sliding window advances)
  track/*      7 scans through the callback sequence of src/ndtpso_slam_node.cpp:177-244 with a non-zero
               initial pose: ranges and the pose after every scan (pso_optimization 30 x 20, one thread, the
               process-global rand() stream of a never-seeded process), plus the final table
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ndtpso_slam_b200 import synthetic as syn  # noqa: E402
from oracle.binding import Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_vectors.npz")


def sparse_table(store, name, flat):
    idx = np.nonzero(flat["built"])[0].astype(np.int32)
    store[f"{name}/cell_index"] = idx
    store[f"{name}/mean"] = flat["mean"][idx]
    store[f"{name}/inv_cov"] = flat["inv_cov"][idx]


def scan_frame(R, cfg, ranges, trans=(0., 0., 0.), cell_side=None):
    S, s = cfg.map_size_m, cfg.sensor
    f = R.frame(trans=trans, width=S, height=S, cell_side=float(S) if cell_side is None else cell_side, init_windows=False)
    f.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    return f


def main():
    R = Reference()
    store = {}
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    room = syn.Room(S)

    # ---- mapbuild
    ref = R.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    ranges, poses, pts, counts = [], [], [], []
    for k in range(24):
        pose = (0.03 * k, 0.01 * k, 0.004 * k)
        r = syn.make_scan(room, s, pose, syn.NoiseLCG(100 + k))
        f = scan_frame(R, cfg, r)
        p = f.flatten_points()
        ref.update(pose, f)
        ranges.append(r)
        poses.append(pose)
        counts.append(len(p))
        pts.append(np.vstack([p, np.zeros((s.beams - len(p), 2))]))
        if k % 3 == 2:
            ref.build()
        if k in (11, 23):
            sparse_table(store, f"mapbuild/table{k}", ref.flatten_map())
    store["mapbuild/ranges"] = np.array(ranges, dtype=np.float32)
    store["mapbuild/poses"] = np.array(poses)
    store["mapbuild/points"] = np.array(pts)
    store["mapbuild/counts"] = np.array(counts, dtype=np.int32)
    print("mapbuild: built cells", len(store["mapbuild/table23/cell_index"]))

    # ---- track
    initial, P, I, steps = (0.3, -0.2, 0.05), 30, 20, 7
    ref = R.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    R.srand(1)
    prev, s_prev, s_diff, s_iter = np.array(initial), np.zeros(3), np.zeros(3), 0
    ranges, poses = [], []
    for k in range(steps):
        r = syn.make_scan(room, s, (initial[0] + 0.04 * k, initial[1] + 0.015 * k, initial[2] + 0.006 * k), syn.NoiseLCG(500 + 10 * k))
        cur = scan_frame(R, cfg, r, trans=initial, cell_side=cfg.cell_side if k == 0 else None)
        if k == 0:
            pose = prev.copy()
        else:
            dev = np.array([.1, .1, 3.1415E-3]) if s_iter < 2 else np.abs(s_diff * 2.)
            s_iter += 1
            pose, _ = R.pso(ref, cur, prev, dev, P, I, use_seed=False, num_threads=1)
            s_diff, s_prev = pose - s_prev, pose.copy()
        prev = pose
        ref.update(pose, cur)
        ranges.append(r)
        poses.append(pose.copy())
    ref.build()
    sparse_table(store, "track/table", ref.flatten_map())
    store["track/ranges"] = np.array(ranges, dtype=np.float32)
    store["track/poses"] = np.array(poses)
    store["track/initial"] = np.array(initial)
    store["track/pso"] = np.array([P, I], dtype=np.int32)
    print("track poses:\n", np.array(poses))

    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
