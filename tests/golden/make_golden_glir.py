"""Generates tests/golden/glir_vectors.npz from the UNMODIFIED reference (oracle/_ref): outputs of
glir_pso_optimization (lib/ndtpso_slam/core.cpp:118-186, population PSO_POPULATION_SIZE = 30, srand(seed) before each
call) on inputs that tests/golden/ref_vectors.npz already holds (checked here to be the same arrays), plus a chain of
calls on the process-global rand() stream.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_glir.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ndtpso_slam_b200 import synthetic as syn  # noqa: E402
from oracle.binding import Reference  # noqa: E402
from tests.problems import Golden  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glir_vectors.npz")
P_REF = 30  # PSO_POPULATION_SIZE, config.h:21


def main():
    R = Reference()
    G = Golden()
    store = {}

    def solve(name, inputs, rf, q, flat, guess, dev, iters, seeds):
        gf = G.flat(inputs)
        assert np.array_equal(gf["points"], flat["points"]) and np.array_equal(gf["mean"], flat["mean"]) and \
            np.array_equal(gf["inv_cov"], flat["inv_cov"]), inputs
        poses = [R.glir(rf, q, guess, dev, iters, seed=s) for s in seeds]
        store[f"{name}/inputs"] = np.array(inputs)
        store[f"{name}/pso"] = np.array([P_REF, iters], dtype=np.int32)
        store[f"{name}/seeds"] = np.array(seeds, dtype=np.uint32)
        store[f"{name}/guess"] = np.array(guess, dtype=np.float64)
        store[f"{name}/deviation"] = np.array(dev, dtype=np.float64)
        store[f"{name}/pose"] = np.array(poses)
        store[f"{name}/cost"] = np.array([R.cost(rf, q, p) for p in poses])
        print(name, "I =", iters, "seeds", len(seeds), "pose[0] =", poses[0], "cost[0] =", store[f"{name}/cost"][0])

    for cfg, seeds, iters in [(syn.CFG1, list(range(1, 13)), 50), (syn.CFG2, list(range(1, 9)), 50), (syn.CFG5[2.0], [1, 2, 3], 100)]:
        ss = syn.scene_a(cfg)
        flat, rf, q = R.flatten_problem(ss)
        solve("glir_" + cfg.name, cfg.name, rf, q, flat, ss.guess, ss.deviation, iters, seeds)

    ss = syn.scene_a(syn.CFG1)
    flat, rf, q = R.flatten_problem(ss)
    solve("glir_zero_dev", "edge", rf, q, flat, ss.guess, (0., 0., 0.), 5, [1, 2])          # every particle starts on the guess
    solve("glir_far_guess", "edge", rf, q, flat, (300., -200., 1.0), ss.deviation, 5, [1])  # every cost is 0: 0/0 in omega and c1
    solve("glir_no_iterations", "edge", rf, q, flat, ss.guess, ss.deviation, 0, [1, 2])
    solve("glir_wide_dev", "edge", rf, q, flat, ss.guess, (3., 3., 1.), 12, [1, 2, 3])

    cfg = syn.MatchConfig("np2", syn.SENSOR_361, 20, 0.3, 20, 15)
    ss2 = syn.scene_a(cfg)
    flat2, rf2, q2 = R.flatten_problem(ss2)
    solve("glir_np2", "np2", rf2, q2, flat2, ss2.guess, ss2.deviation, 15, [1, 2, 3])

    # three calls in a row on the process-global rand() stream after srand(7): 3(P + 2) + 6PI draws each
    ss = syn.scene_a(syn.CFG1)
    flat, rf, q = R.flatten_problem(ss)
    R.srand(7)
    chain, guess = [], np.array(ss.guess)
    for _ in range(3):
        pose = R.glir(rf, q, guess, ss.deviation, 20, use_seed=False)
        chain.append(pose)
        guess = pose
    store["glir_chain/pose"] = np.array(chain)
    store["glir_chain/srand"] = np.array([7], dtype=np.uint32)
    store["glir_chain/iterations"] = np.array([20], dtype=np.int32)
    store["glir_chain/deviation"] = np.array(ss.deviation, dtype=np.float64)
    store["glir_chain/guess"] = np.array(ss.guess, dtype=np.float64)
    print("chain:", chain)

    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
