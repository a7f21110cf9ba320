"""The drop-in NDTFrame (host map building, ndtpso_slam_b200/shim) against the reference.

CPU tests: the (mean, inverse covariance, built) tables and scan points it builds from the synthetic
scans are bit-identical to the reference's (golden fixtures; and the live reference where present).
"""
import ctypes as C
import os

import numpy as np
import pytest

from ndtpso_slam_b200 import frames, synthetic as syn

CASES = [("cfg1", syn.CFG1), ("cfg2", syn.CFG2), ("cfg5_0.25", syn.CFG5[0.25]), ("cfg5_2.0", syn.CFG5[2.0]),
         ("np2", syn.MatchConfig("np2", syn.SENSOR_361, 20, 0.3, 20, 15))]


def test_exports_every_declared_symbol():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "ndtpso_frames.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ndtpso_frame_[a-z0-9_]+)\s*\(", src)))
    lib = frames.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(frames.EXPORTS) == declared


@pytest.mark.parametrize("name,cfg", CASES)
def test_tables_match_golden_bit_for_bit(golden, name, cfg):
    want = golden.flat(name)
    got = frames.problem_from_scans(syn.scene_a(cfg))
    for k in ("w_cells", "h_cells", "width_m", "height_m", "cell_side", "x_min", "x_max", "y_min", "y_max"):
        assert got[k] == want[k], k
    assert np.array_equal(got["points"], want["points"])
    assert np.array_equal(got["built"], want["built"])
    b = want["built"].astype(bool)
    assert np.array_equal(got["mean"][b], want["mean"][b])
    assert np.array_equal(got["inv_cov"][b], want["inv_cov"][b])
    sp = frames.problem_from_scans(syn.scene_a(cfg), sparse=True)
    assert np.array_equal(sp["cell_index"], np.nonzero(b)[0])
    assert np.array_equal(sp["mean"], want["mean"][b]) and np.array_equal(sp["inv_cov"], want["inv_cov"][b])


def test_reset_cells_drops_the_cached_scan():
    """NDTFrame::resetCells (ndtframe.cpp:208-212) clears every cell's points: a scan read before it must not be served again."""
    cfg = syn.CFG1
    ss = syn.scene_a(cfg)
    s = cfg.sensor
    f = frames.Frame(width=cfg.map_size_m, height=cfg.map_size_m, cell_side=cfg.map_size_m, calculate_cells_params=False)
    f.load_laser(ss.query_ranges, s.angle_min, s.angle_increment, s.range_max)
    assert len(f.scan_points()) > 0
    f.reset_cells()
    assert len(f.scan_points()) == 0 and f.point_count() == 0
    f.load_laser(ss.query_ranges, s.angle_min, s.angle_increment, s.range_max)
    assert len(f.scan_points()) > 0
    f.close()


def test_trajectory_problem_matches_golden(golden):
    want = golden.flat("traj17")
    got = frames.problem_from_scans(syn.trajectory_problem(syn.CFG2, 17))
    assert np.array_equal(got["points"], want["points"]) and np.array_equal(got["built"], want["built"])
    b = want["built"].astype(bool)
    assert np.array_equal(got["inv_cov"][b], want["inv_cov"][b])


def test_sliding_window_against_live_reference(reference):
    """Many scans into one map: cells overflow their 50-point slots and the window advances
    (ndtcell.cpp:61-65); repeated build() calls re-add the slot statistics (ndtcell.cpp:37-55)."""
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    room, noise = syn.Room(S), syn.NoiseLCG(99)
    mine = frames.Frame(width=S, height=S, cell_side=cfg.cell_side)
    ref = reference.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
    for k in range(40):
        pose = (0.03 * k, 0.01 * k, 0.004 * k)
        ranges = syn.make_scan(room, s, pose, noise)
        a = frames.Frame(width=S, height=S, cell_side=float(S), calculate_cells_params=False)
        b = reference.frame(width=S, height=S, cell_side=float(S), init_windows=False)
        a.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
        b.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
        assert np.array_equal(a.scan_points(), b.flatten_points())
        mine.update(pose, a)
        ref.update(pose, b)
        if k % 3 == 2:  # the node rebuilds lazily before every match
            mine.build()
            ref.build()
            got, want = mine.map_table(), ref.flatten_map()
            assert np.array_equal(got["built"], want["built"]), k
            m = want["built"].astype(bool)
            assert np.array_equal(got["mean"][m], want["mean"][m]), k
            assert np.array_equal(got["inv_cov"][m], want["inv_cov"][m]), k
    assert mine.point_count() == reference.lib.ref_frame_count_all_points(ref.h)
    assert mine.point_count() > 40 * 300


def test_initial_pose_pretransform(reference):
    """A frame constructed with a non-zero `trans` pre-transforms every beam (ndtframe.cpp:152,175)."""
    cfg = syn.CFG1
    s, S = cfg.sensor, cfg.map_size_m
    ranges = syn.make_scan(syn.Room(S), s, (0.1, 0.2, 0.3), syn.NoiseLCG(5))
    a = frames.Frame(trans=(0.5, -0.25, 0.125), width=S, height=S, cell_side=float(S), calculate_cells_params=False)
    b = reference.frame(trans=(0.5, -0.25, 0.125), width=S, height=S, cell_side=float(S), init_windows=False)
    a.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    b.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
    assert np.array_equal(a.scan_points(), b.flatten_points())


def test_dump_map_csv(tmp_path):
    """dumpMap writes the reference's CSV/gnuplot trio (ndtframe.cpp:268-391)."""
    ref, q = frames.frames_from_scans(syn.scene_a(syn.CFG1))
    ref.add_pose(1.5, (0.1, 0.2, 0.3))
    ref.add_pose(2.5, (0.4, 0.5, 0.6))
    base = str(tmp_path / "map")
    ref.dump_map(base)
    pts = open(base + ".map.csv").read().splitlines()
    assert pts[0] == "x,y" and len(pts) == 1 + ref.point_count() == 1 + 5 * 361
    poses = open(base + ".pose.csv").read().splitlines()
    assert poses[0] == "timestamp,xP,yP,thP,xO,yO,thO"
    assert poses[1] == "1.500000,0.10000,0.20000,0.30000" and len(poses) == 3
    assert "plot '" in open(base + ".gnuplot").read()


@pytest.mark.gpu
def test_align_chain_matches_reference(golden):
    """Four chained NDTFrame::align calls on the process-global rand() stream after srand(7): the
    deviation rule, s_* bookkeeping and the host-drawn random stream reproduce the reference's poses."""
    libc = C.CDLL(None)
    libc.srand(int(golden.z["align_chain/srand"][0]))
    ref, q = frames.frames_from_scans(syn.scene_a(syn.CFG_ALIGN_DEFAULT))
    want = golden.z["align_chain/pose"]
    guess = np.array(syn.DEFAULT_GUESS)
    for k in range(len(want)):
        pose = ref.align(guess, q)
        assert np.abs(pose - want[k]).max() <= 1e-4, (k, pose, want[k])
        guess = pose


@pytest.mark.gpu
def test_cost_function_through_frames(golden):
    ref, q = frames.frames_from_scans(syn.scene_a(syn.CFG1))
    poses, want = golden.z["cfg1/cost_poses"], golden.z["cfg1/cost_values"]
    for p, w in list(zip(poses, want))[:12]:
        got = ref.cost(q, p)
        assert abs(got - w) <= 1e-9 * max(abs(w), 1e-300)


@pytest.mark.gpu
def test_align_with_explicit_config(golden, oracle):
    """align(guess, frame, PSOConfig): the overload that honours the node's PSO parameters."""
    from ndtpso_slam_b200 import capi
    libc = C.CDLL(None)
    libc.srand(3)
    ss = syn.scene_a(syn.CFG1)
    ref, q = frames.frames_from_scans(ss)
    pose = ref.align(ss.guess, q, capi.PsoConfig.make(population=30, iterations=20))
    c = golden.case("cfg1")
    i = c["seeds"].index(3)
    assert np.abs(pose - c["pose"][i]).max() <= 1e-4


def _callback_track(monkeypatch, exact, n_scans=8, P=30, I=20):
    """The node's per-scan callback (ndtpso_slam_node.cpp:186-198) through the drop-in frames, on the golden track's scans."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_vectors.npz"))
    monkeypatch.setenv("NDTPSO_SHIM_EXACT_SCAN", "1" if exact else "0")
    from ndtpso_slam_b200 import capi
    cfg, s = syn.CFG1, syn.CFG1.sensor
    S = cfg.map_size_m
    init = tuple(float(v) for v in z["track/initial"])
    conf = capi.PsoConfig.make(population=P, iterations=I)
    C.CDLL(None).srand(1)  # the golden track is that of a never-seeded process
    ref = frames.Frame(width=S, height=S, cell_side=cfg.cell_side, calculate_cells_params=True)
    cur = frames.Frame(trans=init, width=S, height=S, cell_side=cfg.cell_side, calculate_cells_params=False)
    pose, out, h2d = np.array(init), [], []
    for k, ranges in enumerate(z["track/ranges"][:n_scans]):
        cur.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
        if k > 0:
            pose = ref.align(pose, cur, conf)
            h2d.append(ref.last_h2d_bytes()[0])
        ref.update(pose, cur)
        out.append(np.array(pose))
        cur.close()
        cur = frames.Frame(trans=init, width=S, height=S, cell_side=float(S), calculate_cells_params=False)
    return ref, np.array(out), h2d, z


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [False, True])
def test_callback_runs_on_the_device_mirror(monkeypatch, exact):
    """loadLaser -> align -> update with the map resident in HBM: the reference's golden track, 4 bytes per beam over PCIe
    (16 per point with NDTPSO_SHIM_EXACT_SCAN=1, which keeps the device map bit-identical to the host-built one)."""
    ref, got, h2d, z = _callback_track(monkeypatch, exact)
    want = z["track/poses"][:len(got)]
    assert np.abs(got - want).max() <= 1e-4, (got, want)
    assert ref.device_resident
    P, I = 30, 20
    stream = 4 * (3 + 3 * P + 6 * P * I)
    beams = z["track/ranges"].shape[1]
    for b in h2d:
        if exact:
            assert stream + 16 * 300 < b <= stream + 16 * beams + 64
        else:
            assert b <= stream + 4 * beams + 128, b  # the scan as ranges + the random numbers + guess: never the table
    dev = ref.device_map_table()
    host = ref.map_table()  # builds the host copy on demand
    assert np.array_equal(dev["built"], host["built"])
    m = host["built"].astype(bool)
    if exact:
        assert np.array_equal(dev["mean"][m], host["mean"][m]) and np.array_equal(dev["inv_cov"][m], host["inv_cov"][m])
    else:  # the device redid loadLaser with its own sincos: points within a few ulp of the host's
        assert np.allclose(dev["mean"][m], host["mean"][m], rtol=0, atol=1e-9)
    ref.close()


@pytest.mark.gpu
def test_mirror_is_dropped_when_points_bypass_update(monkeypatch):
    """A point added to the map directly (addPoint / loadLaser on the map itself) is not mirrored: the frame goes back to
    uploading its table per align, and still gives the reference's answers."""
    ref, got, _, z = _callback_track(monkeypatch, False, n_scans=4)
    assert ref.device_resident
    cfg, s = syn.CFG1, syn.CFG1.sensor
    ref.load_laser(z["track/ranges"][4], s.angle_min, s.angle_increment, s.range_max)  # straight into the map
    assert not ref.device_resident
    cur = frames.Frame(width=cfg.map_size_m, height=cfg.map_size_m, cell_side=float(cfg.map_size_m), calculate_cells_params=False)
    cur.load_laser(z["track/ranges"][5], s.angle_min, s.angle_increment, s.range_max)
    pose = ref.align(got[-1], cur)
    assert np.isfinite(pose).all() and ref.last_h2d_bytes()[0] == 0 and not ref.device_resident
    ref.close()
    cur.close()


def test_frames_built_without_align_never_touch_the_device():
    """Map building alone (what bench.py's workload generator does thousands of times) stays on the host: the mirror is only
    created by the first align."""
    ref, q = frames.frames_from_scans(syn.scene_a(syn.CFG1))
    assert not ref.device_resident
    assert ref.map_table()["built"].sum() > 0
    ref.close()
    q.close()


def test_failure_mode_without_a_device():
    """Extension of the drop-in (core.h, pso_set_failure_handler): the reference's CPU path cannot fail, the device path can.  By
    default a failed align is an error (C++: an exception); with the failure mode set the call hands back the caller's guess and
    records why.  Checked where it can be provoked at will: on a machine without a CUDA device."""
    import math
    from ndtpso_slam_b200 import capi
    if capi.load_library().ndtpso_device_count() > 0:
        pytest.skip("needs a machine without a CUDA device")
    L = frames.load_library()
    ref = frames.Frame(width=20, height=20, cell_side=1.0, calculate_cells_params=True)
    scan = frames.Frame(width=20, height=20, cell_side=20.0, calculate_cells_params=False)
    r = np.full(61, 4.0, dtype=np.float32)
    scan.load_laser(r, -1.0, 2.0 / 60, 30.0)
    ref.update((0., 0., 0.), scan)
    guess = (0.1, -0.2, 0.03)
    try:
        with pytest.raises(capi.NdtpsoError, match="no usable CUDA device"):
            ref.align(guess, scan)
        L.ndtpso_frame_set_failure_mode(1)
        pose = ref.align(guess, scan)
        assert np.array_equal(pose, np.array(guess))
        assert math.isnan(L.ndtpso_frame_last_cost())
        assert b"no usable CUDA device" in L.ndtpso_frame_last_error()
        assert ref.cost(scan, guess) == 0.0
    finally:
        L.ndtpso_frame_set_failure_mode(0)
        ref.close()
        scan.close()


@pytest.mark.parametrize("seed", [1, 42, 123456789, 4000000000])
def test_fast_rand_draw_is_the_process_global_stream(seed):
    """The drop-in draws an align's random numbers straight from glibc's rand() state (shim_draw_rand, shim/src/core.cpp): the same
    numbers n calls of rand() return, and rand() afterwards continues as if it had been called n times (core.cpp:14,58-69,84)."""
    L = frames.load_library()
    libc = C.CDLL("libc.so.6")
    libc.rand.restype = C.c_int
    for n in (3, 4, 9, 35, 9093, 21213):
        libc.srand(C.c_uint(seed))
        for _ in range(7):
            libc.rand()
        want = [libc.rand() for _ in range(n + 5)]
        libc.srand(C.c_uint(seed))
        for _ in range(7):
            libc.rand()
        got = (C.c_int32 * n)()
        L.ndtpso_frame_draw_rand(got, n)
        assert list(got) == want[:n]
        assert [libc.rand() for _ in range(5)] == want[n:]
