"""Host logic of bench.py: how the workloads are partitioned over ranks, and the in-run parity summary."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_cfg2_shards_are_contiguous_blocks():
    wl = bench.Workload("cfg2", 256)
    for world in (1, 2, 8):
        seen = []
        for r in range(world):
            sp = wl.specs(r, world)
            assert len(sp) == 256 and all(cs == 0.5 for cs, _ in sp)
            seen += [f for _, f in sp]
        assert seen == list(range(256 * world))  # rank r gets problems [r*B, (r+1)*B)


def test_cfg5_items_go_round_robin_and_balance_cell_sizes():
    wl = bench.Workload("cfg5", 148)
    for world in (1, 2, 4, 8):
        items = set()
        for r in range(world):
            sp = wl.specs(r, world)
            assert len(sp) == 148
            for cs in bench.CFG5_SIZES:
                assert sum(1 for c, _ in sp if c == cs) == 37  # every rank holds the same number of frames of every cell side
            assert all(f % world == r for _, f in sp)
            items |= set(sp)
        assert len(items) == 148 * world  # no item twice, none missing: frames 0 .. 37*world-1 x four cell sides
        assert {f for _, f in items} == set(range(37 * world))


def test_golden_covers_the_eight_gpu_job(batch_golden):
    wl = bench.Workload("cfg2", 256)
    for r in (0, 7):
        pose, cost = wl.golden(r, 8)
        assert np.array_equal(pose, batch_golden.pose[256 * r:256 * (r + 1)]) and np.array_equal(cost, batch_golden.cost[256 * r:256 * (r + 1)])
    assert bench.Workload("cfg2", 512).golden(7, 8) is None  # beyond the committed vectors
    assert bench.Workload("cfg5", 148).golden(0, 1) is None


def test_parity_stats():
    want_pose = np.array([[1.0, 2.0, 0.5], [0.0, 0.0, 0.0]])
    want_cost = np.array([-100.0, -50.0])
    st = bench.parity_stats(want_pose + np.array([[0.0, 0.0, 0.0], [0.0, 2e-5, 0.0]]), want_cost * np.array([1.0, 1.0 + 1e-7]), want_pose, want_cost)
    assert st["n_checked"] == 2 and st["bit_exact_poses"] == 1
    assert abs(st["max_abs_dpose"] - 2e-5) < 1e-12 and abs(st["max_rel_dscore"] - 1e-7) < 1e-12


def test_cpu_worker_solves_the_batchs_own_problems(batch_golden):
    """The cpu_baseline leg's worker on two problems of the cfg2 batch: the reference's golden results, bit for bit."""
    from oracle import binding
    kind = "reference" if os.path.exists(binding.REF_SO) else "port"
    wl = bench.Workload("cfg2", 256)
    secs, pose, cost = bench._cpu_worker((kind, "cfg2", wl.P, wl.I, wl.specs(0, 1)[3:5]))
    assert secs > 0
    assert np.array_equal(pose, batch_golden.pose[3:5]) and np.array_equal(cost, batch_golden.cost[3:5])
