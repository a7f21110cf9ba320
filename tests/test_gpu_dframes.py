"""Device-resident reference frames (include/ndtpso_dframes.h) against the unmodified reference.

SURVEY.md section 8f rows 1-2: NDTFrame::loadLaser / update / build on the GPU, and the whole per-scan
callback (loadLaser -> align -> update) with the map kept in HBM.

Parity bars, written where they are checked:
  * update + build, fed the reference's own scan points and host poses: tables BIT-IDENTICAL
    (IEEE add/mul/div/sqrt in the reference's order, no contraction);
  * loadLaser: the same points in the same order, coordinates within 4 ulp (cos/sin of the beam
    angle are evaluated by the GPU's libm instead of glibc's);
  * align / tracking: BASELINE.json's tolerances, |pose - ref| <= 1e-4, |score - ref| <= 1e-5 relative.
"""
import numpy as np
import pytest

from ndtpso_slam_b200 import capi, synthetic as syn
from tests.problems import POSE_ATOL, SCORE_RTOL, rel_err

pytestmark = pytest.mark.gpu


def _scan(cfg, pose, seed):
    return syn.make_scan(syn.Room(cfg.map_size_m), cfg.sensor, pose, syn.NoiseLCG(seed))


def _ref_scan_frame(reference, cfg, ranges, trans=(0., 0., 0.), cell_side=None):
    S, s = cfg.map_size_m, cfg.sensor
    f = reference.frame(trans=trans, width=S, height=S, cell_side=float(S) if cell_side is None else cell_side, init_windows=False)
    f.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    return f


def _assert_tables_equal(got, want, what):
    assert np.array_equal(got["built"], want["built"]), what
    m = want["built"].astype(bool)
    assert np.array_equal(got["mean"][m], want["mean"][m]), what
    assert np.array_equal(got["inv_cov"][m], want["inv_cov"][m]), what


def _ulp_close(a, b, ulps):
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)))
    return np.all(np.abs(a - b) <= tol)


def test_load_laser_points_and_order(ctx, reference):
    from ndtpso_slam_b200.dframes import DeviceFrames
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    scans = [_scan(cfg, (0.1, 0.2, 0.3), 5), _scan(cfg, (-1.0, 0.5, -0.7), 6), _scan(cfg, (2.0, -1.5, 1.9), 7)]
    # what loadLaser filters (ndtframe.cpp:165): zero, negative, below epsilon, beyond range_max, inf, NaN; and points
    # outside the frame (dropped by addPoint, ndtframe.cpp:217-220)
    scans[1][::7] = 0.0
    scans[1][3::11] = np.float32(0.05)
    scans[1][5::13] = np.float32(31.0)
    scans[2][::5] = np.inf
    scans[2][1::9] = np.nan
    scans[2][2::17] = np.float32(-1.0)
    scans[2][4::19] = np.float32(14.9)
    df = DeviceFrames(ctx, 3, S, S, cfg.cell_side, s.beams)
    df.load_laser(np.stack(scans), s.angle_min, s.angle_increment, s.range_max)
    exact = total = 0
    for b in range(3):
        want = _ref_scan_frame(reference, cfg, scans[b]).flatten_points()
        got = df.download_scan(b)
        assert got.shape == want.shape, b
        assert _ulp_close(got, want, 4), b
        exact += int((got == want).sum())
        total += want.size
    assert exact >= 0.5 * total  # most coordinates are bit-identical; the rest differ in the last place
    # a scan frame with s_trans (the node passes its initial pose) and a multi-cell scan frame: cell-index-major order
    trans = [(0.5, -0.25, 0.125), (0., 0., 0.), (1e-7, 0., 0.)]  # the last one isZero(1e-6): not applied
    df2 = DeviceFrames(ctx, 3, S, S, cfg.cell_side, s.beams, scan_cell_side=cfg.cell_side)
    df2.load_laser(np.stack(scans), s.angle_min, s.angle_increment, s.range_max, scan_trans=trans)
    for b in range(3):
        want = _ref_scan_frame(reference, cfg, scans[b], trans=trans[b], cell_side=cfg.cell_side).flatten_points()
        got = df2.download_scan(b)
        assert got.shape == want.shape, b
        assert _ulp_close(got, want, 8), b
    assert not df.status().any() and not df2.status().any()
    df.close()
    df2.close()


def test_update_build_bit_exact_with_sliding_window(ctx, reference):
    """40 scans merged into two maps: cells overflow their 50-point slots and the window advances
    (ndtcell.cpp:61-65); build() without new points re-adds the slot statistics (ndtcell.cpp:37-55)."""
    from ndtpso_slam_b200.dframes import DeviceFrames
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    df = DeviceFrames(ctx, 2, S, S, cfg.cell_side, s.beams)
    refs = [reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True) for _ in range(2)]
    for k in range(40):
        poses = [(0.03 * k, 0.01 * k, 0.004 * k), (-0.02 * k, 0.015 * k, -0.006 * k)]
        pts = []
        for b in range(2):
            f = _ref_scan_frame(reference, cfg, _scan(cfg, poses[b], 100 + 2 * k + b))
            pts.append(f.flatten_points())
            refs[b].update(poses[b], f)
        df.set_scan_points(pts)
        df.update(poses)
        if k % 3 == 2 or k == 39:
            for rep in range(2 if k % 6 == 5 else 1):
                df.build()
                for b in range(2):
                    refs[b].build()
            for b in range(2):
                _assert_tables_equal(df.download_map(b), refs[b].flatten_map(), (k, b))
    info = df.info(0)
    assert info["built"] == int(refs[0].flatten_map()["built"].sum()) and info["created"] >= info["built"] > 20
    assert not df.status().any()
    df.close()


def test_window_wraps_around(ctx, reference):
    """A 2 x 2-cell map fed 130 scans: every scan closes a slot in every cell, so the 100-slot window wraps
    and the reference's build() re-reads the points a re-opened slot held 100 slots ago (ndtcell.cpp:49-52
    on a slot that addPoint has not yet reset, ndtcell.cpp:22-27)."""
    from ndtpso_slam_b200.dframes import DeviceFrames
    cfg = syn.MatchConfig("wrap", syn.SENSOR_361, 20, 10.0, 30, 20)
    s, S = cfg.sensor, cfg.map_size_m
    df = DeviceFrames(ctx, 1, S, S, cfg.cell_side, s.beams, window_points=65536)
    ref = reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    for k in range(130):
        pose = (0.01 * (k % 17), 0.02 * (k % 5), 0.003 * (k % 11))
        f = _ref_scan_frame(reference, cfg, _scan(cfg, pose, 300 + k))
        ref.update(pose, f)
        df.set_scan_points([f.flatten_points()])
        df.update([pose])
        ref.build()
        df.build()
        if k % 2 == 1:  # a second build before the next scan: the re-opened slot is read before addPoint resets it
            ref.build()
            df.build()
        if k >= 95 or k % 10 == 0:
            _assert_tables_equal(df.download_map(0), ref.flatten_map(), k)
    assert not df.status().any()
    # the same run with a ring too small for 100 slots raises the truncation flag instead of silently diverging
    small = DeviceFrames(ctx, 1, S, S, cfg.cell_side, s.beams, window_points=1024)
    for k in range(130):
        pose = (0.01 * (k % 17), 0.02 * (k % 5), 0.003 * (k % 11))
        small.set_scan_points([_ref_scan_frame(reference, cfg, _scan(cfg, pose, 300 + k)).flatten_points()])
        small.update([pose])
        small.build()
        small.build()  # reads the slot that was just re-opened: after the wrap its old points have left the ring
    from ndtpso_slam_b200 import dframes
    assert small.status()[0] & dframes.DF_WINDOW_TRUNCATED
    small.close()
    df.close()


def test_cell_pool_overflow_is_reported(ctx, reference):
    from ndtpso_slam_b200 import dframes
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    df = dframes.DeviceFrames(ctx, 1, S, S, cfg.cell_side, s.beams, max_cells=8)
    df.load_laser(_scan(cfg, (0., 0., 0.), 1)[None], s.angle_min, s.angle_increment, s.range_max)
    df.update([(0., 0., 0.)])
    assert df.status()[0] & dframes.DF_CELL_POOL_FULL
    assert df.info(0)["created"] == 8
    df.close()


def test_align_seeded_matches_reference(ctx, reference):
    """ndtpso_dframes_align on maps built on the device == reference pso_optimization on the reference's map."""
    from ndtpso_slam_b200.dframes import DeviceFrames, RNG_SEEDED
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    n = 3
    df = DeviceFrames(ctx, n, S, S, cfg.cell_side, s.beams)
    refs, queries, guesses = [], [], []
    for b in range(n):
        ss = syn.trajectory_problem(cfg, b)
        rf = reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
        refs.append(rf)
        for k, (pose, ranges) in enumerate(ss.map_scans):
            f = _ref_scan_frame(reference, cfg, ranges)
            rf.update(pose, f)
            # frame b of the batch receives the same points; the other frames keep their previous scan
            pts = [f.flatten_points() if j == b else np.zeros((0, 2)) for j in range(n)]
            df.set_scan_points(pts)
            df.update([pose if j == b else (0., 0., 0.) for j in range(n)])
        queries.append(_ref_scan_frame(reference, cfg, ss.query_ranges))
        guesses.append(ss.guess)
    df.set_scan_points([q.flatten_points() for q in queries])
    conf = capi.PsoConfig.make(population=cfg.particles, iterations=cfg.iterations)
    seeds = [11, 12, 13]
    pose, cost = df.align(guesses, conf, RNG_SEEDED, seeds)
    for b in range(n):
        want, _ = reference.pso(refs[b], queries[b], guesses[b], syn.DEFAULT_DEVIATION, cfg.particles, cfg.iterations, seed=seeds[b])
        assert np.abs(pose[b] - want).max() <= POSE_ATOL, (b, pose[b], want)
        assert rel_err(cost[b], reference.cost(refs[b], queries[b], want)) <= SCORE_RTOL
        _assert_tables_equal(df.download_map(b), refs[b].flatten_map(), b)
    assert df.info(0)["align_calls"] == 1
    df.close()


def _reference_track(reference, cfg, scans, initial_pose, P, I):
    """NDTPSONode::scan_matcher_ (src/ndtpso_slam_node.cpp:177-244) driven on the reference library in its
    deterministic mode: pso_optimization with one thread on the process-global rand() stream, never re-seeded,
    and align()'s deviation rule / bookkeeping (ndtframe.cpp:251-266) restated here."""
    S = cfg.map_size_m
    ref_frame = reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    reference.srand(1)  # a fresh process: glibc's default seed
    poses, prev, s_prev, s_diff, s_iter = [], np.array(initial_pose, dtype=np.float64), np.zeros(3), np.zeros(3), 0
    for k, ranges in enumerate(scans):
        cur = _ref_scan_frame(reference, cfg, ranges, trans=initial_pose, cell_side=cfg.cell_side if k == 0 else None)
        if k == 0:
            pose = prev.copy()
        else:
            dev = np.array([.1, .1, 3.1415E-3]) if s_iter < 2 else np.abs(s_diff * 2.)
            s_iter += 1
            pose, _ = reference.pso(ref_frame, cur, prev, dev, P, I, use_seed=False, num_threads=1)
            s_diff, s_prev = pose - s_prev, pose.copy()
        prev = pose
        ref_frame.update(pose, cur)
        poses.append(pose.copy())
    return np.array(poses), ref_frame


@pytest.mark.parametrize("initial", [(0., 0., 0.), (0.3, -0.2, 0.05)])
def test_track_step_matches_reference_callback(ctx, reference, initial):
    """The whole per-scan callback with the maps resident in HBM, two robots, against the reference loop."""
    from ndtpso_slam_b200.dframes import DeviceFrames, RNG_CONTINUE
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    n, steps, P, I = 2, 7, 30, 20
    room = syn.Room(S)
    scans = [[syn.make_scan(room, s, (initial[0] + 0.04 * k + 0.01 * r, initial[1] + 0.015 * k, initial[2] + 0.006 * k * (1 - 2 * r)),
                            syn.NoiseLCG(500 + 10 * k + r)) for k in range(steps)] for r in range(n)]
    want = [_reference_track(reference, cfg, scans[r], initial, P, I) for r in range(n)]
    df = DeviceFrames(ctx, n, S, S, cfg.cell_side, s.beams)
    conf = capi.PsoConfig.make(population=P, iterations=I)
    init = None if not any(initial) else [initial] * n
    for k in range(steps):
        pose, cost = df.track_step(np.stack([scans[r][k] for r in range(n)]), s.angle_min, s.angle_increment, s.range_max,
                                   initial_poses=init, conf=conf, rng_mode=RNG_CONTINUE)
        for r in range(n):
            assert np.abs(pose[r] - want[r][0][k]).max() <= POSE_ATOL, (k, r, pose[r], want[r][0][k])
    df.build()
    for r in range(n):
        want[r][1].build()
        got, ref_map = df.download_map(r), want[r][1].flatten_map()
        assert np.array_equal(got["built"], ref_map["built"])
        m = ref_map["built"].astype(bool)
        assert np.allclose(got["mean"][m], ref_map["mean"][m], rtol=0, atol=1e-6)
    assert df.info(0)["align_calls"] == steps - 1
    assert not df.status().any()
    df.close()


def test_continuing_rand_stream(ctx, reference, oracle):
    """NDTPSO_RNG_CONTINUE: align k of a frame consumes draws [k*D, (k+1)*D) of the srand(1) stream, like a process
    that never seeds.  Checked through the result: three chained aligns equal three oracle PSO
    runs fed consecutive slices of that stream."""
    from ndtpso_slam_b200.dframes import DeviceFrames, RNG_CONTINUE
    from ndtpso_slam_b200 import frames
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    ss = syn.scene_a(cfg)
    flat = frames.problem_from_scans(ss)
    df = DeviceFrames(ctx, 1, S, S, cfg.cell_side, s.beams)
    for pose, ranges in ss.map_scans:
        df.load_laser(ranges[None], s.angle_min, s.angle_increment, s.range_max)
        df.update([pose])
    df.load_laser(ss.query_ranges[None], s.angle_min, s.angle_increment, s.range_max)
    P, I = 17, 9  # D = 3 + 3P + 6PI = 972: not a multiple of 32, so the generator's look-ahead is exercised
    D = 3 + 3 * P + 6 * P * I
    stream = oracle.rand_stream(1, 3 * D)
    conf = capi.PsoConfig.make(population=P, iterations=I)
    guess, s_prev, s_diff = np.array(ss.guess), np.zeros(3), np.zeros(3)
    for k in range(3):
        dev = np.array([.1, .1, 3.1415E-3]) if k < 2 else np.abs(s_diff * 2.)
        want, wcost, _ = oracle.pso(flat, guess, dev, P, I, stream=stream[k * D:(k + 1) * D])
        pose, cost = df.align([guess], conf, RNG_CONTINUE)
        assert np.abs(pose[0] - want).max() <= POSE_ATOL, (k, pose[0], want)
        assert rel_err(cost[0], wcost) <= SCORE_RTOL
        s_diff, s_prev, guess = want - s_prev, want, want
    df.close()


def test_edge_cases_empty_scan_and_empty_map(ctx, reference, oracle):
    """What the reference does with nothing to match: an empty map (every cost is 0, gbest never improves after the seed
    particle) and an empty scan (loadLaser rejects every beam) — same poses as the oracle, no error, no status bit."""
    from ndtpso_slam_b200.dframes import DeviceFrames, RNG_SEEDED
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    n = 3
    df = DeviceFrames(ctx, n, S, S, cfg.cell_side, s.beams)
    conf = capi.PsoConfig.make(population=9, iterations=4)
    ranges = np.stack([_scan(cfg, (0.1, 0.0, 0.02), 40), np.zeros(s.beams, dtype=np.float32), _scan(cfg, (0., 0.1, 0.), 41)])
    # 1) empty maps: nothing was ever merged
    df.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    assert [df.info(b)["scan_points"] for b in range(n)] == [len(_ref_scan_frame(reference, cfg, ranges[b]).flatten_points()) for b in range(n)]
    assert df.info(1)["scan_points"] == 0
    guess = np.array([[0.2, 0.08, 0.04]] * n)
    pose, cost = df.align(guess, conf, RNG_SEEDED, [5, 6, 7])
    geom = reference.frame(width=S, height=S, cell_side=cfg.cell_side).geometry()
    ncell = geom["n_cells"]
    empty = dict(geom, mean=np.zeros((ncell, 2)), inv_cov=np.zeros((ncell, 4)), built=np.zeros(ncell, dtype=np.uint8))
    for b in range(n):
        f = dict(empty, points=df.download_scan(b))
        want, wcost, _ = oracle.pso(f, guess[b], syn.DEFAULT_DEVIATION, 9, 4, seed=5 + b)
        assert np.array_equal(pose[b], want) and cost[b] == wcost == 0.0, b
    # 2) merge the scans (frame 1 merges nothing), then match the same scans again: frame 1 has a map-less, scan-less problem
    df.update(pose)
    pose2, cost2 = df.align(None, conf, RNG_SEEDED, [8, 9, 10])
    assert cost2[1] == 0.0 and cost2[0] < -50.0 and cost2[2] < -50.0
    assert df.info(1)["created"] == 0 and df.info(0)["created"] > 20
    assert not df.status().any()
    df.close()


def test_points_outside_the_map_are_dropped_like_the_reference(ctx, reference):
    """A pose that throws most of the scan outside the 20 m frame: addPoint drops those points (getCellIndex == -1,
    ndtframe.cpp:217-220); the device tables stay bit-identical to the reference's."""
    from ndtpso_slam_b200.dframes import DeviceFrames
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    df = DeviceFrames(ctx, 1, S, S, cfg.cell_side, s.beams)
    ref = reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    for k, pose in enumerate([(0., 0., 0.), (6.5, -3.0, 0.7), (-9.0, 9.0, 2.0), (30.0, 30.0, 0.1)]):
        f = _ref_scan_frame(reference, cfg, _scan(cfg, (0., 0., 0.), 60 + k))
        ref.update(pose, f)
        df.set_scan_points([f.flatten_points()])
        df.update([pose])
        ref.build()
        df.build()
        _assert_tables_equal(df.download_map(0), ref.flatten_map(), k)
    assert not df.status().any()
    df.close()


def test_device_frames_screen_on_off(reference):
    """80 robots (one CTA per robot: the form the fp32 screen runs in) tracked for a few scans with the screen on and off:
    bit-identical poses, and the screen did settle evaluations."""
    from ndtpso_slam_b200 import dframes
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    n, steps = 80, 4
    room = syn.Room(S)
    scans = [np.stack([syn.make_scan(room, s, (0.03 * k + 0.002 * r, 0.01 * k, 0.004 * k - 0.0005 * r), syn.NoiseLCG(900 + 7 * k + r)) for r in range(n)])
             for k in range(steps)]
    conf = capi.PsoConfig.make(population=30, iterations=20)
    out = {}
    for scr in (0, 1):
        cx = capi.Context(0)
        cx.set_option(capi.OPT_SCREEN, scr)
        df = dframes.DeviceFrames(cx, n, S, S, cfg.cell_side, s.beams)
        poses = []
        for k in range(steps):
            p, _ = df.track_step(scans[k], s.angle_min, s.angle_increment, s.range_max, conf=conf)
            poses.append(p)
        out[scr] = (np.array(poses), df.pso_stats())
        assert not df.status().any()
        df.close()
        cx.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert (out[0][1][:, 3] == 0).all() and out[1][1][:, 3].sum() > 0
    assert np.abs(out[1][0][-1][:, 0] - 0.09).max() < 0.05  # and the robots were tracked (x advances 3 cm per scan)


def test_irregular_sigma_flag_follows_the_table(ctx):
    """NDTPSO_DF_IRREGULAR_SIGMA (the frame set then takes the generic PSO kernel) describes the table as the last build left it:
    a cell of three identical points has a zero covariance — determinant 0, a non-finite inverse (ndtcell.cpp:93-111 divides by
    it) —, and once more points make it a proper Gaussian the flag is gone again; the event bits stay."""
    from ndtpso_slam_b200 import dframes
    df = dframes.DeviceFrames(ctx, 1, 20, 20, 1.0, 64)
    origin = [(0., 0., 0.)]
    good = np.array([[2.1, 3.1], [2.3, 3.4], [2.6, 3.2], [2.8, 3.7], [2.2, 3.8]])
    df.set_scan_points([good])
    df.update(origin)
    df.build()
    assert not df.status()[0] & dframes.DF_IRREGULAR_SIGMA
    df.set_scan_points([np.array([[5.5, 5.5]] * 3)])  # another cell: three times the same point
    df.update(origin)
    df.build()
    assert df.status()[0] & dframes.DF_IRREGULAR_SIGMA
    m = df.download_map(0)
    cell = 15 + 20 * 15  # (5.5 + 10) / 1, row 15
    assert m["built"][cell] and not np.isfinite(m["inv_cov"][cell]).all()
    df.set_scan_points([np.array([[5.2, 5.3], [5.7, 5.4], [5.4, 5.8], [5.6, 5.1]])])  # the same cell becomes a proper Gaussian
    df.update(origin)
    df.build()
    m = df.download_map(0)
    assert np.isfinite(m["inv_cov"][cell]).all()
    assert not df.status()[0] & dframes.DF_IRREGULAR_SIGMA
    df.close()
