"""The N > 1 path on CPU: world_size 2 (and 3) over gloo.  Every rank solves its shard with the
ORACLE standing in for the GPU (this is a test of the partition + all-gather logic, not of the kernel)
and must end up with the same full result table as a single-rank run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ndtpso_slam_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 256, 2048):
        for g in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(s for s in sizes if s or n < g) <= -(-n // g) if n else True
    assert sharding.shard_bounds(2048, 8, 3) == (768, 1024)  # BASELINE.json configs[3]
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_problems, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Oracle
    from tests.problems import Golden
    g, orc = Golden(), Oracle()
    c, flats = g.problems("cfg1")
    flats = (flats * 2)[:n_problems]

    def solve(shard):
        poses, costs = [], []
        for f in shard:
            p, co, _ = orc.pso(f, f["guess"], f["deviation"], 8, 4, seed=f["seed"])
            poses.append(p)
            costs.append(co)
        return np.array(poses).reshape(-1, 3), np.array(costs)

    pose, cost = sharding.solve_sharded(solve, flats, world, rank)
    np.save(os.path.join(out_dir, f"pose_{rank}.npy"), pose)
    np.save(os.path.join(out_dir, f"cost_{rank}.npy"), cost)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_problems", [(2, 10), (3, 7)])
def test_sharded_solve_equals_single_rank(tmp_path, world, n_problems):
    mp.spawn(_worker, args=(world, _free_port(), n_problems, str(tmp_path)), nprocs=world, join=True)
    from oracle.binding import Oracle
    from tests.problems import Golden
    g, orc = Golden(), Oracle()
    _, flats = g.problems("cfg1")
    flats = (flats * 2)[:n_problems]
    want_pose = np.array([orc.pso(f, f["guess"], f["deviation"], 8, 4, seed=f["seed"])[0] for f in flats])
    want_cost = np.array([orc.pso(f, f["guess"], f["deviation"], 8, 4, seed=f["seed"])[1] for f in flats])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"pose_{r}.npy"), want_pose)
        assert np.array_equal(np.load(tmp_path / f"cost_{r}.npy"), want_cost)


def _exchange_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # no CUDA context here: creating the exchange fails on every rank, and make_exchange must come back with None on
    # every rank after the SAME sequence of collectives (a rank that bailed out early would leave the others hanging)
    ex = sharding.make_exchange(None, 4, world, rank)
    open(os.path.join(out_dir, f"ex_{rank}.txt"), "w").write("none" if ex is None else "exchange")
    dist.barrier()
    dist.destroy_process_group()


def test_make_exchange_fails_on_all_ranks_together(tmp_path):
    mp.spawn(_exchange_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ex_{r}.txt").read() for r in range(2)] == ["none", "none"]
