"""The multi-GPU result exchange fused into the PSO kernel's epilogue (ndtpso_exchange_*, include/ndtpso_b200.h):
peer stores + arrival flags instead of a collective.  One GPU is enough to exercise the logic: two ranks are two
contexts (on two devices when the box has them, else on the same one) connected in-process; the multi-process CUDA-IPC
form is exercised by `bench.py --gpus N` (torchrun), which checks it against the NCCL all-gather."""
import numpy as np
import pytest

from ndtpso_slam_b200 import capi

pytestmark = pytest.mark.gpu


def test_exchange_single_rank(ctx, golden):
    c, flats = golden.problems("cfg1")
    conf = capi.PsoConfig.make(population=c["P"], iterations=c["I"])
    bt = ctx.batch(flats, conf)
    ex = capi.Exchange(ctx, 1, 0, len(flats))
    bt.attach_exchange(ex)
    for _ in range(3):  # epochs alternate between the two halves of the gathered buffer
        bt.solve()
        pose, cost = ex.results()
        p0, c0 = bt.results()
        assert np.array_equal(pose, p0) and np.array_equal(cost, c0)
    assert np.abs(pose - c["pose"]).max() <= 1e-4
    bt.attach_exchange(None)
    bt.solve()
    bt.close()
    ex.close()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_ranks_in_one_process(golden, world):
    lib = capi.load_library()
    ndev = lib.ndtpso_device_count()
    c, flats = golden.problems("cfg1")
    n = 5
    conf = capi.PsoConfig.make(population=c["P"], iterations=c["I"])
    ctxs = [capi.Context(r % ndev) for r in range(world)]
    shards = [[dict(f) for f in (flats * 2)[r * n:(r + 1) * n]] for r in range(world)]
    exs = [capi.Exchange(ctxs[r], world, r, n) for r in range(world)]
    for r in range(world):
        exs[r].connect_local(exs)
    bts = [ctxs[r].batch(shards[r], conf) for r in range(world)]
    for r in range(world):
        bts[r].attach_exchange(exs[r])
    for rep in range(2):
        for r in range(world):
            bts[r].solve()
        want = np.concatenate([bts[r].results()[0] for r in range(world)])
        wcost = np.concatenate([bts[r].results()[1] for r in range(world)])
        for r in range(world):
            pose, cost = exs[r].results()
            assert np.array_equal(pose, want) and np.array_equal(cost, wcost), (rep, r)
    for r in range(world):
        bts[r].close()
        exs[r].close()
        ctxs[r].close()


def test_exchange_wait_times_out_instead_of_hanging(ctx, golden):
    """A rank whose peer never solves gets an error from the bounded wait, not a hung GPU."""
    c, flats = golden.problems("cfg1")
    conf = capi.PsoConfig.make(population=4, iterations=2)
    ctx.set_option(capi.OPT_EXCHANGE_TIMEOUT_MS, 50)
    a, b = capi.Exchange(ctx, 2, 0, len(flats)), capi.Exchange(ctx, 2, 1, len(flats))
    a.connect_local([a, b])
    b.connect_local([a, b])
    bt = ctx.batch(flats, conf)
    bt.attach_exchange(a)
    bt.solve()  # rank 1 never solves
    with pytest.raises(capi.NdtpsoError):
        a.results()
    ctx.set_option(capi.OPT_EXCHANGE_TIMEOUT_MS, 10000)
    bt.close()
    a.close()
    b.close()


def test_device_frames_publish_through_the_exchange():
    """Two ranks (contexts of one process), each tracking its own robots with device-resident maps: after a step every rank
    holds all robots' poses, delivered by the PSO kernel's epilogue."""
    from ndtpso_slam_b200 import dframes, synthetic as syn
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    world, n = 2, 3
    ndev = capi.load_library().ndtpso_device_count()
    room = syn.Room(S)
    ctxs = [capi.Context(r % ndev) for r in range(world)]
    dfs = [dframes.DeviceFrames(ctxs[r], n, S, S, cfg.cell_side, s.beams) for r in range(world)]
    exs = [capi.Exchange(ctxs[r], world, r, n) for r in range(world)]
    for r in range(world):
        exs[r].connect_local(exs)
        dfs[r].attach_exchange(exs[r])
    conf = capi.PsoConfig.make(population=20, iterations=10)
    for k in range(3):
        mine = []
        for r in range(world):
            scans = np.stack([syn.make_scan(room, s, (0.03 * k + 0.01 * (r * n + i), 0.01 * k, 0.004 * k), syn.NoiseLCG(40 + 10 * k + r * n + i))
                              for i in range(n)])
            mine.append(dfs[r].track_step(scans, s.angle_min, s.angle_increment, s.range_max, conf=conf)[0])
        if k == 0:
            continue  # the first scan is not matched: nothing is published
        want = np.concatenate(mine)
        for r in range(world):
            pose, _ = exs[r].results()
            assert np.array_equal(pose, want), (k, r)
    for r in range(world):
        dfs[r].close()
        exs[r].close()
        ctxs[r].close()
