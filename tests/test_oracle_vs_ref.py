"""The oracle pinned against the LIVE reference (oracle/_ref: unmodified reference sources).

Runs wherever oracle/_ref/libndtpso_ref.so exists (build container; GPU box via the snapshot).
Inputs are flattened once from the reference's own frames and fed to both sides.
Bar: bit-exact.
"""
import numpy as np
import pytest

from ndtpso_slam_b200 import synthetic as syn


@pytest.fixture(scope="module")
def scenes(reference):
    out = {}
    for cfg in (syn.CFG1, syn.CFG2):
        ss = syn.scene_a(cfg)
        out[cfg.name] = (cfg, ss) + reference.flatten_problem(ss)
    return out


def test_sizes_match_survey(scenes):
    # cfg1 400 cells / 53 built / 361 points; cfg2 10000 / 354 / 1081
    _, _, flat, _, _ = scenes["cfg1"]
    assert (flat["n_cells"], int(flat["built"].sum()), flat["points"].shape[0]) == (400, 53, 361)
    _, _, flat, _, _ = scenes["cfg2"]
    assert (flat["n_cells"], int(flat["built"].sum()), flat["points"].shape[0]) == (10000, 354, 1081)


@pytest.mark.parametrize("name,seeds", [("cfg1", range(100, 140)), ("cfg2", range(100, 104))])
def test_pso_many_seeds(reference, oracle, scenes, name, seeds):
    cfg, ss, flat, rf, q = scenes[name]
    for seed in seeds:
        pr, _ = reference.pso(rf, q, ss.guess, ss.deviation, cfg.particles, cfg.iterations, seed=seed, num_threads=1)
        po, co, _ = oracle.pso(flat, ss.guess, ss.deviation, cfg.particles, cfg.iterations, seed=seed)
        assert np.array_equal(pr, po), (name, seed)
        assert reference.cost(rf, q, pr) == co


def test_cost_random_poses(reference, oracle, scenes):
    rng = np.random.default_rng(7)
    for name in ("cfg1", "cfg2"):
        cfg, ss, flat, rf, q = scenes[name]
        poses = np.array(ss.guess) + rng.normal(size=(64, 3)) * np.array([1.0, 1.0, 0.3])
        got = oracle.cost_many(flat, poses)
        want = np.array([reference.cost(rf, q, p) for p in poses])
        assert np.array_equal(got, want)


def test_trajectory_problems(reference, oracle):
    for b in (0, 5, 39, 40, 123):
        ss = syn.trajectory_problem(syn.CFG2, b)
        flat, rf, q = reference.flatten_problem(ss)
        assert flat["points"].shape[0] > 900
        pr, _ = reference.pso(rf, q, ss.guess, ss.deviation, 20, 8, seed=1 + b, num_threads=1)
        po, co, _ = oracle.pso(flat, ss.guess, ss.deviation, 20, 8, seed=1 + b)
        assert np.array_equal(pr, po)


def test_process_rand_stream(reference, oracle):
    reference.srand(99)
    want = [reference.rand() for _ in range(500)]
    assert list(oracle.rand_stream(99, 500)) == want
