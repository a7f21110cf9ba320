"""Parity of the CUDA path (through the C ABI) with the reference's golden outputs and the oracle.

Tolerances (BASELINE.json north_star): |pose - ref| <= 1e-4 per component, |score - ref| <= 1e-5 relative.
The device sums a scan's scores in a warp-tree order and fuses multiply-adds, so scores are not
bit-identical; poses usually are (the swarm update itself is computed without contraction).
"""
import numpy as np
import pytest

from ndtpso_slam_b200 import capi
from tests.problems import POSE_ATOL, SCORE_RTOL, SOLVED, empty_points, oracle_pso_many, rel_err

pytestmark = pytest.mark.gpu


def conf_of(c):
    return capi.PsoConfig.make(population=c["P"], iterations=c["I"], w=c["w"], c1=c["c1"], c2=c["c2"], w_dumping=c["w_dumping"])


@pytest.mark.parametrize("case,inputs", SOLVED)
@pytest.mark.parametrize("sparse", [False, True])
def test_pso_vs_golden(golden, ctx, case, inputs, sparse):
    c, flats = golden.problems(case, inputs, sparse=sparse)
    if case == "edge_empty_scan":
        flats = [empty_points(f) for f in flats]
    pose, cost = ctx.align_batch(flats, conf_of(c))
    assert np.abs(pose - c["pose"]).max() <= POSE_ATOL, (case, pose, c["pose"])
    assert rel_err(cost, c["cost"]).max() <= SCORE_RTOL, (case, cost, c["cost"])


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_cost_vs_golden(golden, ctx, name):
    flat = golden.flat(name)
    poses = golden.z[f"{name}/cost_poses"]
    want = golden.z[f"{name}/cost_values"]
    got = ctx.cost_batch([flat], poses[None])[0]
    # the exponent -(d'Sd)/2 reaches several hundred for far poses: its own rounding (1e-16 relative)
    # becomes ~1e-13 relative in exp(); 1e-9 is still four orders inside the 1e-5 bar
    assert rel_err(got, want).max() <= 1e-9, rel_err(got, want).max()
    assert (got[want == 0] == 0).all()


def test_device_rand_equals_host_stream(golden, oracle, ctx):
    """K1's on-device glibc rand() == the host-drawn stream (drop-in mode), bit for bit in the result."""
    c, flats = golden.problems("cfg1")
    n = 3 + 3 * c["P"] + 6 * c["P"] * c["I"]
    hosted = []
    for f in flats[:6]:
        g = dict(f)
        g["rand_stream"] = oracle.rand_stream(f["seed"], n)
        g["seed"] = 0
        hosted.append(g)
    p1, c1 = ctx.align_batch(flats[:6], conf_of(c))
    p2, c2 = ctx.align_batch(hosted, conf_of(c))
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)


@pytest.mark.parametrize("seed", [0, 1, 42, 123456789, 4294967295])
def test_device_rand_seeds(golden, oracle, ctx, seed):
    """Unusual seeds (0 -> 1, >= 2^31) go through K1 and must match the oracle run on the same seed."""
    c = golden.case("cfg1")
    f = golden.flat("cfg1")
    f.update(guess=c["guess"], deviation=c["deviation"], seed=seed)
    pose, cost = ctx.align_batch([f], capi.PsoConfig.make(population=12, iterations=6))
    po, co, _ = oracle.pso(f, c["guess"], c["deviation"], 12, 6, seed=seed)
    assert np.abs(pose[0] - po).max() <= POSE_ATOL and rel_err(cost[0], co) <= SCORE_RTOL


def test_batch_order_and_sharding_invariance(golden, ctx):
    """Each problem's result is independent of what else is in the batch and of its position."""
    c, flats = golden.problems("cfg1")
    cf = conf_of(c)
    full_pose, full_cost = ctx.align_batch(flats, cf)
    perm = np.random.default_rng(3).permutation(len(flats))
    pp, pc = ctx.align_batch([flats[i] for i in perm], cf)
    assert np.array_equal(pp, full_pose[perm]) and np.array_equal(pc, full_cost[perm])
    half = len(flats) // 2
    a = ctx.align_batch(flats[:half], cf)
    b = ctx.align_batch(flats[half:], cf)
    assert np.array_equal(np.vstack([a[0], b[0]]), full_pose)
    one = ctx.align_batch(flats[5:6], cf)
    assert np.array_equal(one[0][0], full_pose[5])


def test_deterministic_and_resolvable(golden, ctx):
    c, flats = golden.problems("cfg2")
    bt = ctx.batch(flats, conf_of(c))
    bt.solve()
    p1, c1 = bt.results()
    bt.solve()
    p2, c2 = bt.results()
    st = bt.stats()
    bt.close()
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert np.abs(p1 - c["pose"]).max() <= POSE_ATOL
    # every iteration takes at least one round; extra rounds come from gbest improvements (replays) and from the
    # small speculation window the kernel uses while improvements are frequent
    assert (st[:, 0] >= c["I"]).all() and (st[:, 1] >= 1).all()
    # the speculation window is an optimisation only: any window gives bit-identical results
    from ndtpso_slam_b200 import capi
    for window in (0, 4 | (1 << 16), 16 | (2 << 16)):
        cx = capi.Context(0)
        cx.set_option(capi.OPT_HOT_CHUNK, window)
        b2 = cx.batch(flats, conf_of(c))
        b2.solve()
        p3, c3 = b2.results()
        s3 = b2.stats()
        b2.close()
        cx.close()
        assert np.array_equal(p3, p1) and np.array_equal(c3, c1), window
        assert np.array_equal(s3[:, 1], st[:, 1])  # the same gbest improvements, whatever the rounds
        if window == 0:  # whole-swarm speculation: rounds = iterations + improvements not on an iteration's last particle
            assert (s3[:, 0] <= c["I"] + s3[:, 1]).all()


def test_oracle_on_fresh_inputs(oracle, ctx):
    """Seeded random map + scan (not from the reference's map builder): CUDA vs oracle."""
    rng = np.random.default_rng(11)
    gw = gh = 40
    n = gw * gh
    built = (rng.random(n) < 0.2).astype(np.uint8)
    cx = (np.arange(n) % gw + 0.5) * 0.5 - 10.0
    cy = (np.arange(n) // gw + 0.5) * 0.5 - 10.0
    mean = np.stack([cx, cy], 1) + rng.normal(size=(n, 2)) * 0.05
    a = rng.uniform(5, 60, n); b = rng.uniform(5, 60, n); r = rng.uniform(-0.5, 0.5, n) * np.sqrt(a * b)
    icov = np.stack([a, r, r, b], 1)
    pts = rng.uniform(-9, 9, size=(777, 2))
    flat = dict(points=pts, mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh, width_m=20.0, height_m=20.0,
                cell_side=0.5, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0)
    flats = []
    for s in range(1, 9):
        f = dict(flat)
        f.update(guess=(0.1, -0.2, 0.05), deviation=(0.3, 0.3, 0.05), seed=s)
        flats.append(f)
    pose, cost = ctx.align_batch(flats, capi.PsoConfig.make(population=25, iterations=12))
    for i, f in enumerate(flats):
        po, co, _ = oracle.pso(f, f["guess"], f["deviation"], 25, 12, seed=f["seed"])
        assert np.abs(pose[i] - po).max() <= POSE_ATOL
        assert rel_err(cost[i], co) <= SCORE_RTOL
    poses = np.array(flat["points"][:50].tolist())[:, :1] * 0 + rng.normal(size=(50, 3))
    got = ctx.cost_batch([flat], poses[None])[0]
    assert rel_err(got, oracle.cost_many(flat, poses)).max() <= 1e-9


def test_argument_errors(golden, ctx):
    c, flats = golden.problems("cfg1")
    bad = dict(flats[0]); bad["w_cells"] = 0
    with pytest.raises(capi.NdtpsoError) as e:
        ctx.align_batch([bad], conf_of(c))
    assert e.value.code == capi.ERR_ARG
    with pytest.raises(capi.NdtpsoError) as e:
        ctx.align_batch(flats[:1], capi.PsoConfig.make(population=100000, iterations=1))
    assert e.value.code == capi.ERR_LIMIT
    pose, cost = ctx.align_batch([], conf_of(c))
    assert pose.shape == (0, 3)


@pytest.mark.parametrize("kernel,npt,warps", [(capi.KERNEL_WARP_PER_PARTICLE, 0, 4), (capi.KERNEL_WARP_PER_PARTICLE, 0, 16),
                                              (capi.KERNEL_POINT_SLICED, 1, 0), (capi.KERNEL_POINT_SLICED, 3, 0),
                                              (capi.KERNEL_POINT_SLICED, 6, 0), (capi.KERNEL_POINT_SLICED, 0, 9)])
def test_every_kernel_configuration(golden, kernel, npt, warps):
    """Both PSO kernels (generic warp-per-particle, point-sliced) and their launch shapes give the
    reference's answers; the two kernels agree with each other to the last bit of the pose."""
    c = capi.Context(0)
    c.set_option(capi.OPT_KERNEL, kernel)
    c.set_option(capi.OPT_POINTS_PER_THREAD, npt)
    c.set_option(capi.OPT_WARPS_PER_CTA, warps)
    try:
        for case, inputs in [("cfg1", "cfg1"), ("np2", "np2"), ("edge_wide_dev", "edge"), ("edge_one_particle", "edge"),
                             ("edge_no_iterations", "edge"), ("edge_far_guess", "edge")]:
            cs, flats = golden.problems(case, inputs)
            pose, cost = c.align_batch(flats, conf_of(cs))
            assert np.abs(pose - cs["pose"]).max() <= POSE_ATOL, (case, kernel, npt, warps)
            assert rel_err(cost, cs["cost"]).max() <= SCORE_RTOL, (case, kernel, npt, warps)
    finally:
        c.close()


def test_large_scan_falls_back_to_generic_kernel(oracle, ctx):
    """A scan longer than the point-sliced kernel holds in registers (6 x 640) still solves."""
    rng = np.random.default_rng(5)
    gw = gh = 20
    n = gw * gh
    built = (rng.random(n) < 0.5).astype(np.uint8)
    mean = np.stack([(np.arange(n) % gw + 0.5) - 10.0, (np.arange(n) // gw + 0.5) - 10.0], 1)
    icov = np.tile(np.array([[8.0, 1.0, 1.0, 6.0]]), (n, 1))
    flat = dict(points=rng.uniform(-9, 9, size=(5000, 2)), mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh,
                width_m=20.0, height_m=20.0, cell_side=1.0, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0,
                guess=(0.3, 0.1, 0.02), deviation=(0.2, 0.2, 0.02), seed=9)
    pose, cost = ctx.align_batch([flat], capi.PsoConfig.make(population=9, iterations=5))
    po, co, _ = oracle.pso(flat, flat["guess"], flat["deviation"], 9, 5, seed=9)
    assert np.abs(pose[0] - po).max() <= POSE_ATOL and rel_err(cost[0], co) <= SCORE_RTOL


def test_asymmetric_inverse_covariance(oracle, ctx):
    """S01 != S10 cannot come out of NDTCell::build, but the ABI accepts it: the library must notice
    and evaluate the full 2x2 form (generic kernel), like the reference's (d'S)d."""
    rng = np.random.default_rng(8)
    gw = gh = 16
    n = gw * gh
    built = np.ones(n, dtype=np.uint8)
    mean = np.stack([(np.arange(n) % gw + 0.5) - 8.0, (np.arange(n) // gw + 0.5) - 8.0], 1)
    icov = np.stack([rng.uniform(4, 9, n), rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(4, 9, n)], 1)
    flat = dict(points=rng.uniform(-7, 7, size=(300, 2)), mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh,
                width_m=16.0, height_m=16.0, cell_side=1.0, x_min=-8.0, x_max=8.0, y_min=-8.0, y_max=8.0,
                guess=(0.1, 0.1, 0.01), deviation=(0.2, 0.2, 0.02), seed=4)
    pose, cost = ctx.align_batch([flat], capi.PsoConfig.make(population=16, iterations=8))
    po, co, _ = oracle.pso(flat, flat["guess"], flat["deviation"], 16, 8, seed=4)
    assert np.abs(pose[0] - po).max() <= POSE_ATOL and rel_err(cost[0], co) <= SCORE_RTOL


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_cluster_sizes(golden, cluster):
    """Small batches spread each problem over a thread-block cluster (DSMEM exchange of the partial
    scores); every cluster size must reproduce the reference."""
    c = capi.Context(0)
    c.set_option(capi.OPT_CLUSTER, cluster)
    try:
        for case, inputs, take in [("cfg2", "cfg2", 2), ("cfg1", "cfg1", 3), ("align_default", "align_default", 1),
                                   ("edge_one_particle", "edge", 2), ("edge_empty_scan", "edge", 1), ("np2", "np2", 2)]:
            cs, flats = golden.problems(case, inputs)
            flats = flats[:take]
            if case == "edge_empty_scan":
                flats = [empty_points(f) for f in flats]
            pose, cost = c.align_batch(flats, conf_of(cs))
            assert np.abs(pose - cs["pose"][:take]).max() <= POSE_ATOL, (case, cluster)
            assert rel_err(cost, cs["cost"][:take]).max() <= SCORE_RTOL, (case, cluster)
    finally:
        c.close()


def test_submit_collect_pipelined(golden, ctx):
    """ndtpso_align_submit / ndtpso_align_collect with several batches in flight == ndtpso_align_batch."""
    c, flats = golden.problems("cfg1")
    cf = conf_of(c)
    want_pose, want_cost = ctx.align_batch(flats, cf)
    sets = [capi.ProblemSet(flats[i::3]) for i in range(3)]
    tickets = [ctx.align_submit(ps, cf) for ps in sets]
    for i, t in enumerate(tickets):
        pose, cost = ctx.align_collect(t)
        assert np.array_equal(pose, want_pose[i::3]) and np.array_equal(cost, want_cost[i::3])


def test_indefinite_inverse_covariance(oracle, ctx):
    """A symmetric but indefinite "Sigma^-1" makes the exponent positive; the library must take the
    exact (library exp) kernel and agree with the oracle."""
    rng = np.random.default_rng(12)
    gw = gh = 16
    n = gw * gh
    built = np.ones(n, dtype=np.uint8)
    mean = np.stack([(np.arange(n) % gw + 0.5) - 8.0, (np.arange(n) // gw + 0.5) - 8.0], 1)
    off = rng.uniform(3, 5, n)
    icov = np.stack([rng.uniform(1, 2, n), off, off, rng.uniform(1, 2, n)], 1)  # det < 0
    flat = dict(points=rng.uniform(-7, 7, size=(200, 2)), mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh,
                width_m=16.0, height_m=16.0, cell_side=1.0, x_min=-8.0, x_max=8.0, y_min=-8.0, y_max=8.0,
                guess=(0.1, 0.1, 0.01), deviation=(0.2, 0.2, 0.02), seed=6)
    pose, cost = ctx.align_batch([flat], capi.PsoConfig.make(population=10, iterations=6))
    po, co, _ = oracle.pso(flat, flat["guess"], flat["deviation"], 10, 6, seed=6)
    assert np.abs(pose[0] - po).max() <= POSE_ATOL and rel_err(cost[0], co) <= SCORE_RTOL


def _parity(pose, cost, want_pose, want_cost, what):
    dp = np.abs(pose - want_pose).max(axis=1)
    ds = rel_err(cost, want_cost)
    assert dp.max() <= POSE_ATOL, (what, int(dp.argmax()), dp.max())
    assert ds.max() <= SCORE_RTOL, (what, int(ds.argmax()), ds.max())
    return int((dp == 0).sum())


def test_cfg3_batch_against_reference_and_oracle(oracle, ctx, batch_golden, traj_batch):
    """BASELINE.json configs[2]: a batch of 256 trajectory problems (own table each), 70 x 50.  EVERY result is compared with
    the unmodified reference's (golden vectors) and with the oracle run here on the same inputs."""
    from ndtpso_slam_b200 import synthetic as syn
    flats = traj_batch(256)
    cf = capi.PsoConfig.make(population=70, iterations=50)
    pose, cost = ctx.align_batch(flats, cf)
    exact = _parity(pose, cost, batch_golden.pose[:256], batch_golden.cost[:256], "vs reference")
    assert exact >= 250, exact  # the swarm arithmetic is uncontracted: poses are bit-identical unless a comparison flips
    po, co = oracle_pso_many(oracle, flats, 70, 50)
    _parity(pose, cost, po, co, "vs oracle")
    assert np.array_equal(po, batch_golden.pose[:256])  # and the oracle itself is the reference, bit for bit, on all of them
    # idempotence / determinism at full size: the same batch again, and as two halves
    pose2, cost2 = ctx.align_batch(flats, cf)
    assert np.array_equal(pose, pose2) and np.array_equal(cost, cost2)
    a = ctx.align_batch(flats[:128], cf)
    assert np.array_equal(a[0], pose[:128])
    # the matcher converges: most results are within 2 cm / 0.2 deg of the true pose of the synthetic scene
    truth = np.array([syn.trajectory_problem(syn.CFG2, b).true_pose for b in range(256)])
    err = np.abs(pose - truth)
    assert np.median(err[:, :2].max(axis=1)) < 0.02 and np.median(err[:, 2]) < 0.004


def test_cfg4_every_shard_against_reference(ctx, batch_golden, traj_batch):
    """BASELINE.json configs[3]: 2048 trajectory problems in shards of 256 (what each of 8 GPUs gets).  Every shard is solved
    here on one GPU and every one of the 2048 results is compared with the unmodified reference's golden vectors."""
    flats = traj_batch(2048)
    cf = capi.PsoConfig.make(population=70, iterations=50)
    exact = 0
    for r in range(8):
        pose, cost = ctx.align_batch(flats[256 * r:256 * (r + 1)], cf)
        exact += _parity(pose, cost, batch_golden.pose[256 * r:256 * (r + 1)], batch_golden.cost[256 * r:256 * (r + 1)], f"shard {r}")
    assert exact >= 2000, exact
    # one launch of all 2048 gives the same answers as the shards (a problem's result does not depend on its batch)
    pose, cost = ctx.align_batch(flats, cf)
    _parity(pose, cost, batch_golden.pose, batch_golden.cost, "one batch of 2048")


@pytest.mark.parametrize("cs", [0.25, 0.5, 1.0, 2.0])
def test_cfg5_more_seeds(golden, ctx, batch_golden, cs):
    """BASELINE.json configs[4] (200 x 100 on four cell sizes): four more seeds per cell size against the unmodified reference."""
    c = golden.case(f"cfg5_{cs}")
    base = golden.flat(f"cfg5_{cs}")
    seeds, want_pose, want_cost = batch_golden.cfg5_more(cs)
    flats = []
    for s in seeds:
        f = dict(base)
        f.update(guess=c["guess"], deviation=c["deviation"], seed=s)
        flats.append(f)
    for n_copies in (1, 40):  # a small batch (cluster form) and one that takes the one-CTA-per-problem, screened form
        pose, cost = ctx.align_batch(flats * n_copies, conf_of(c))
        for k in range(n_copies):
            _parity(pose[k * len(seeds):(k + 1) * len(seeds)], cost[k * len(seeds):(k + 1) * len(seeds)], want_pose, want_cost, (cs, n_copies))


@pytest.mark.parametrize("case,inputs", [("cfg1", "cfg1"), ("cfg2", "cfg2"), ("traj17", "traj17"), ("cfg5_0.25", "cfg5_0.25"), ("cfg5_2.0", "cfg5_2.0"),
                                         ("edge_far_guess", "edge"), ("edge_wide_dev", "edge"), ("edge_damped", "edge"), ("np2", "np2")])
def test_fp32_screen_changes_nothing(golden, case, inputs):
    """The fp32 lower-bound screen (ndtpso_pso_sliced.cuh) only ever drops candidates whose fp64 cost provably fails
    `cost < pbest` (core.cpp:94): poses AND costs are bit-identical with the screen on and off, and equal the reference."""
    from ndtpso_slam_b200 import capi
    c, flats = golden.problems(case, inputs)
    flats = (flats * 150)[:150]  # more problems than SMs / 2: one CTA per problem, the form the screen runs in
    out = {}
    for scr in (0, 1):
        cx = capi.Context(0)
        cx.set_option(capi.OPT_SCREEN, scr)
        bt = cx.batch(flats, conf_of(c))
        bt.solve()
        out[scr] = bt.results() + (bt.stats_ex(),)
        bt.close()
        cx.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2][:, :2], out[1][2][:, :2])  # same rounds, same gbest updates
    assert (out[0][2][:, 3] == 0).all()
    n = len(c["seeds"])
    assert np.abs(out[1][0][:n] - c["pose"]).max() <= POSE_ATOL
    if case in ("cfg2", "traj17"):  # (200 particles on the 0.25 m map leave no shared memory for the screen's tables)
        assert out[1][2][:, 3].sum() > 0.5 * (out[1][2][:, 2].sum() + out[1][2][:, 3].sum())  # most evaluations are settled in fp32
    if case == "np2":  # cell side 0.3: not the geometry the screen is written for
        assert (out[1][2][:, 3] == 0).all()


def test_fp32_screen_on_random_tables(oracle):
    """Random maps with sharp, strongly correlated Sigma^-1 (up to 1e6, correlation up to 0.999), means anywhere in their
    cells, scans with points on cell edges and on the frame's border: screen on == screen off bit for bit, == oracle."""
    rng = np.random.default_rng(5)
    gw = gh = 40
    n = gw * gh
    flats = []
    for m_id in range(3):
        built = (rng.random(n) < 0.3).astype(np.uint8)
        ix, iy = np.arange(n) % gw, np.arange(n) // gw
        mean = np.stack([(ix + rng.random(n)) * 0.5 - 10.0, (iy + rng.random(n)) * 0.5 - 10.0], 1)
        scale = 10.0 ** rng.uniform(0.5, 6.0, n)
        a, b = scale * rng.uniform(0.2, 1.0, n), scale * rng.uniform(0.2, 1.0, n)
        r = rng.choice([0.0, 0.5, 0.99, 0.999], n) * rng.choice([-1.0, 1.0], n) * np.sqrt(a * b)
        icov = np.stack([a, r, r, b], 1)
        pts = rng.uniform(-9.5, 9.5, size=(900, 2))
        pts[:60] = np.round(pts[:60] * 2) / 2          # exactly on cell edges
        pts[60:80, 0] = 10.0 - 1e-9 * rng.random(20)   # a hair inside the frame's border
        flat = dict(points=pts, mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh, width_m=20.0, height_m=20.0,
                    cell_side=0.5, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0)
        for s in range(50):
            f = dict(flat)
            f.update(guess=(0.0, 0.0, 0.0) if s % 2 else (0.25, -0.5, 0.3), deviation=(0.3, 0.3, 0.05), seed=1 + 50 * m_id + s)
            flats.append(f)
    conf = capi.PsoConfig.make(population=40, iterations=15)
    out = {}
    for scr in (0, 1):
        cx = capi.Context(0)
        cx.set_option(capi.OPT_SCREEN, scr)
        bt = cx.batch(flats, conf)
        bt.solve()
        out[scr] = bt.results() + (bt.stats_ex(),)
        bt.close()
        cx.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert out[1][2][:, 3].sum() > 0  # the screen did run
    for i in (0, 57, 101, 149):
        po, co, _ = oracle.pso(flats[i], flats[i]["guess"], flats[i]["deviation"], 40, 15, seed=flats[i]["seed"])
        assert np.abs(out[1][0][i] - po).max() <= POSE_ATOL
        assert rel_err(out[1][1][i], co) <= SCORE_RTOL


def _check_bounds(ctx, flat, poses, what):
    lower = ctx.screen_bounds([flat], poses[None])[0]
    cost = ctx.cost_batch([flat], poses[None])[0]
    bad = lower > cost  # a lower bound above the fp64 cost would be a bug in the screen, whether or not it flips a decision
    assert not bad.any(), (what, poses[bad][:3], lower[bad][:3], cost[bad][:3])
    return lower, cost


def test_fp32_screen_bound_pose_by_pose(golden, ctx):
    """ndtpso_screen_bounds runs the screen's own code: for thousands of poses — near the optimum, spread like a swarm, far
    outside the map, on cell edges — the bound never exceeds the fp64 cost, and it is tight where it matters."""
    rng = np.random.default_rng(3)
    for name in ("cfg2", "traj17", "cfg1", "cfg5_0.25", "cfg5_2.0"):
        flat = golden.flat(name)
        c = golden.case(name)
        best = c["pose"][0]
        sets = [best + rng.normal(size=(256, 3)) * np.array([0.002, 0.002, 0.0005]),     # a converged swarm
                best + rng.normal(size=(256, 3)) * np.array([0.1, 0.1, 0.01]),           # the initial spread
                best + rng.normal(size=(256, 3)) * np.array([2.0, 2.0, 0.5]),            # diverged particles
                best + rng.uniform(-1, 1, size=(128, 3)) * np.array([60.0, 60.0, 6.0]),  # far outside the map
                np.concatenate([np.round(best[:2] * 4) / 4 + rng.integers(-3, 4, size=(128, 2)) * 0.25, np.zeros((128, 1))], 1)]
        tight = []
        for k, poses in enumerate(sets):
            lower, cost = _check_bounds(ctx, flat, poses, (name, k))
            if k == 0:
                tight.append(np.median((cost - lower) / np.abs(cost)))
        assert tight[0] < (0.10 if name == "cfg5_0.25" else 0.05), (name, tight)  # close to the true cost near the optimum (0.25 m cells: 6.5 %)


def test_fp32_screen_bound_on_random_tables(ctx):
    """The same on random sharp / strongly correlated tables, with scan points on cell edges and on the frame's border."""
    rng = np.random.default_rng(8)
    gw = gh = 40
    n = gw * gh
    for trial in range(4):
        built = (rng.random(n) < 0.3).astype(np.uint8)
        ix, iy = np.arange(n) % gw, np.arange(n) // gw
        mean = np.stack([(ix + rng.random(n)) * 0.5 - 10.0, (iy + rng.random(n)) * 0.5 - 10.0], 1)
        scale = 10.0 ** rng.uniform(0.0, 6.5, n)
        a, b = scale * rng.uniform(0.2, 1.0, n), scale * rng.uniform(0.2, 1.0, n)
        r = rng.choice([0.0, 0.5, 0.99, 0.9999], n) * rng.choice([-1.0, 1.0], n) * np.sqrt(a * b)
        icov = np.stack([a, r, r, b], 1)
        pts = rng.uniform(-9.5, 9.5, size=(1000, 2))
        pts[:100] = np.round(pts[:100] * 2) / 2
        pts[100:130, 0] = 10.0 - 1e-9 * rng.random(30)
        flat = dict(points=pts, mean=mean, inv_cov=icov, built=built, w_cells=gw, h_cells=gh, width_m=20.0, height_m=20.0,
                    cell_side=0.5, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0)
        poses = np.concatenate([rng.normal(size=(400, 3)) * np.array([0.3, 0.3, 0.05]), np.zeros((8, 3)),
                                rng.uniform(-1, 1, size=(104, 3)) * np.array([25.0, 25.0, 3.2])])
        _check_bounds(ctx, flat, poses, trial)
