"""Generates tests/golden/batch_vectors.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_batch.py

  traj/pose[2048, 3], traj/cost[2048]   pso_optimization (70 particles x 50 iterations, num_threads = 1, srand(1 + b) before the
                                        call) of trajectory problem b = 0 .. 2047 (synthetic.trajectory_problem(CFG2, b)): the
                                        problems of BASELINE.json configs[2] (the first 256) and configs[3] (all 2048, 256 per GPU);
                                        cost = cost_function at the returned pose.  Inputs are not stored: the drop-in NDTFrame
                                        rebuilds them bit for bit from the synthetic scans (tests/test_shim_frames.py), and a table
                                        that differed from the reference's would fail these vectors.
  traj/table_crc[2048]                  CRC-32 of the reference's own flattened table and scan (mean, inv_cov, built, points) per
                                        problem, so that a mismatch can be told apart from a solver difference
  cfg5_<cs>_more/{seeds, pose, cost}    four more seeds (2 .. 5) of every configs[4] case (200 x 100), scene A as in ref_vectors.npz
"""
import multiprocessing as mp
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ndtpso_slam_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "batch_vectors.npz")
N_TRAJ = 2048


def table_crc(flat) -> int:
    c = 0
    for k in ("mean", "inv_cov", "built", "points"):
        c = zlib.crc32(np.ascontiguousarray(flat[k]).tobytes(), c)
    return c


def _traj(b):
    from oracle.binding import Reference
    R = Reference()
    ss = syn.trajectory_problem(syn.CFG2, b)
    flat, rf, q = R.flatten_problem(ss)
    pose, _ = R.pso(rf, q, ss.guess, ss.deviation, 70, 50, seed=1 + b, num_threads=1)
    return b, pose, R.cost(rf, q, pose), table_crc(flat)


def _cfg5(args):
    cs, seed = args
    from oracle.binding import Reference
    R = Reference()
    cfg = syn.CFG5[cs]
    ss = syn.scene_a(cfg)
    rf, q = R.build_problem(ss)
    pose, _ = R.pso(rf, q, ss.guess, ss.deviation, cfg.particles, cfg.iterations, seed=seed, num_threads=1)
    return cs, seed, pose, R.cost(rf, q, pose)


def main():
    store = {}
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(_traj, range(N_TRAJ), chunksize=8)
        res5 = pool.map(_cfg5, [(cs, s) for cs in (0.25, 0.5, 1.0, 2.0) for s in (2, 3, 4, 5)])
    res.sort(key=lambda r: r[0])
    store["traj/pose"] = np.array([r[1] for r in res])
    store["traj/cost"] = np.array([r[2] for r in res])
    store["traj/table_crc"] = np.array([r[3] for r in res], dtype=np.uint32)
    store["traj/pso"] = np.array([70, 50], dtype=np.int32)
    for cs in (0.25, 0.5, 1.0, 2.0):
        rows = sorted([r for r in res5 if r[0] == cs], key=lambda r: r[1])
        store[f"cfg5_{cs}_more/seeds"] = np.array([r[1] for r in rows], dtype=np.uint32)
        store[f"cfg5_{cs}_more/pose"] = np.array([r[2] for r in rows])
        store[f"cfg5_{cs}_more/cost"] = np.array([r[3] for r in rows])
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "traj pose[0]", store["traj/pose"][0], "cost[0]", store["traj/cost"][0])


if __name__ == "__main__":
    main()
