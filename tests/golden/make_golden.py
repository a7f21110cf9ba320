"""Generates tests/golden/ref_vectors.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference has no golden vectors of its own; these are outputs of
the reference itself — NDTFrame::loadLaser/update/build for the inputs, pso_optimization
(num_threads = 1, srand(seed) before each call) and cost_function for the outputs — on the
synthetic scenes of ndtpso_slam_b200/synthetic.py.  Map tables are stored sparsely (built
cells only); tests/problems.py densifies them again.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ndtpso_slam_b200 import synthetic as syn  # noqa: E402
from oracle.binding import Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
GEOM_KEYS = ["w_cells", "h_cells", "width_m", "height_m", "cell_side", "x_min", "x_max", "y_min", "y_max"]


def pack_problem(store, name, flat):
    built = flat["built"].astype(bool)
    idx = np.nonzero(built)[0].astype(np.int32)
    store[f"{name}/points"] = flat["points"]
    store[f"{name}/cell_index"] = idx
    store[f"{name}/mean"] = flat["mean"][idx]
    store[f"{name}/inv_cov"] = flat["inv_cov"][idx]
    store[f"{name}/geom"] = np.array([flat[k] for k in GEOM_KEYS], dtype=np.float64)


def solve_cases(R, store, name, rf, q, guess, dev, P, I, seeds, **coef):
    poses, costs, secs = [], [], []
    for s in seeds:
        pose, t = R.pso(rf, q, guess, dev, P, I, seed=s, num_threads=1, **coef)
        poses.append(pose)
        costs.append(R.cost(rf, q, pose))
        secs.append(t)
    store[f"{name}/pso"] = np.array([P, I], dtype=np.int32)
    store[f"{name}/coef"] = np.array([coef.get("w", .8), coef.get("c1", 2.), coef.get("c2", 2.), coef.get("w_dumping", 1.)])
    store[f"{name}/seeds"] = np.array(seeds, dtype=np.uint32)
    store[f"{name}/guess"] = np.array(guess, dtype=np.float64)
    store[f"{name}/deviation"] = np.array(dev, dtype=np.float64)
    store[f"{name}/pose"] = np.array(poses)
    store[f"{name}/cost"] = np.array(costs)
    print(f"{name}: P={P} I={I} seeds={len(seeds)} ref {np.mean(secs) * 1e3:.1f} ms/match  pose[0]={poses[0]}")


def main():
    R = Reference()
    store = {}
    rng = np.random.default_rng(2026)

    # --- BASELINE.json configs: cfg1, cfg2, cfg5 x 4 cell sizes, and what align() really runs (30 x 50)
    for cfg, seeds in [(syn.CFG1, list(range(1, 17))), (syn.CFG2, list(range(1, 7))), (syn.CFG_ALIGN_DEFAULT, [1, 2, 3]),
                       (syn.CFG5[0.25], [1]), (syn.CFG5[0.5], [1]), (syn.CFG5[1.0], [1]), (syn.CFG5[2.0], [1])]:
        ss = syn.scene_a(cfg)
        flat, rf, q = R.flatten_problem(ss)
        pack_problem(store, cfg.name, flat)
        solve_cases(R, store, cfg.name, rf, q, ss.guess, ss.deviation, cfg.particles, cfg.iterations, seeds)
        if cfg.name in ("cfg1", "cfg2"):
            # cost_function alone on random poses, some far outside the map
            poses = np.array(ss.guess) + rng.uniform(-1, 1, size=(48, 3)) * np.array([0.5, 0.5, 0.2])
            far = np.array(ss.guess) + rng.uniform(-1, 1, size=(16, 3)) * np.array([40.0, 40.0, 6.0])
            poses = np.vstack([poses, far])
            store[f"{cfg.name}/cost_poses"] = poses
            store[f"{cfg.name}/cost_values"] = np.array([R.cost(rf, q, p) for p in poses])

    # --- a trajectory problem (configs[2]/[3] draw from these)
    ss = syn.trajectory_problem(syn.CFG2, 17)
    flat, rf, q = R.flatten_problem(ss)
    pack_problem(store, "traj17", flat)
    solve_cases(R, store, "traj17", rf, q, ss.guess, ss.deviation, 70, 50, [18])

    # --- edge cases on the cfg1 scene
    ss = syn.scene_a(syn.CFG1)
    flat, rf, q = R.flatten_problem(ss)
    pack_problem(store, "edge", flat)
    solve_cases(R, store, "edge_zero_dev", rf, q, ss.guess, (0., 0., 0.), 10, 5, [1, 2])
    solve_cases(R, store, "edge_far_guess", rf, q, (300., -200., 1.0), ss.deviation, 10, 5, [1])  # every point out of bounds
    solve_cases(R, store, "edge_one_particle", rf, q, ss.guess, ss.deviation, 1, 30, [1, 2])
    solve_cases(R, store, "edge_no_iterations", rf, q, ss.guess, ss.deviation, 12, 0, [1])
    solve_cases(R, store, "edge_damped", rf, q, ss.guess, ss.deviation, 20, 25, [1, 2], w=0.7, c1=1.4, c2=1.6, w_dumping=0.97)
    solve_cases(R, store, "edge_wide_dev", rf, q, ss.guess, (3., 3., 1.), 33, 12, [1, 2])

    # empty scan: every range is rejected by loadLaser -> zero points
    S = syn.CFG1.map_size_m
    qe = R.frame(width=S, height=S, cell_side=float(S), init_windows=False)
    qe.load_laser(np.zeros(361, dtype=np.float32), syn.SENSOR_361.angle_min, syn.SENSOR_361.angle_increment, 30.0)
    assert qe.flatten_points().shape[0] == 0
    solve_cases(R, store, "edge_empty_scan", rf, qe, ss.guess, ss.deviation, 8, 4, [1])

    # non-power-of-two cell side (true division in getCellIndex) and a non-square frame
    cfg = syn.MatchConfig("np2", syn.SENSOR_361, 20, 0.3, 20, 15)
    ss2 = syn.scene_a(cfg)
    flat2, rf2, q2 = R.flatten_problem(ss2)
    pack_problem(store, "np2", flat2)
    solve_cases(R, store, "np2", rf2, q2, ss2.guess, ss2.deviation, 20, 15, [1, 2, 3])

    # --- NDTFrame::align bookkeeping: 4 chained calls on the process-global rand() stream after srand(7)
    ss = syn.scene_a(syn.CFG_ALIGN_DEFAULT)
    flat, rf, q = R.flatten_problem(ss)
    R.srand(7)
    guess = np.array(ss.guess)
    chain = []
    for _ in range(4):
        pose = R.align(rf, guess, q)
        chain.append(pose)
        guess = pose
    store["align_chain/pose"] = np.array(chain)
    store["align_chain/srand"] = np.array([7], dtype=np.uint32)
    print("align chain:", chain)

    # --- rand() itself
    for seed in (1, 42, 123456789, 0, 4294967295):
        R.srand(seed)
        store[f"rand/{seed}"] = np.array([R.rand() for _ in range(2000)], dtype=np.int32)

    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
