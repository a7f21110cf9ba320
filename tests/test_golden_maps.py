"""Golden vectors of the map-building rows (tests/golden/map_vectors.npz, made from the unmodified reference by
tests/golden/make_golden_maps.py): NDTFrame::loadLaser / update / build and the per-scan callback.

CPU: the drop-in host NDTFrame reproduces them bit for bit.
GPU (-m gpu): the device-resident frames (include/ndtpso_dframes.h) reproduce them — tables bit for bit when fed
the golden scan points, scan points within 4 ulp from the golden ranges, tracked poses within 1e-4.
"""
import os

import numpy as np
import pytest

from ndtpso_slam_b200 import synthetic as syn

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_vectors.npz"))
CFG = syn.CFG1
POSE_ATOL = 1e-4  # BASELINE.json north_star


def _check_table(got, name, what):
    idx = Z[f"{name}/cell_index"]
    assert np.array_equal(np.nonzero(got["built"])[0], idx), what
    assert np.array_equal(got["mean"][idx], Z[f"{name}/mean"]), what
    assert np.array_equal(got["inv_cov"][idx], Z[f"{name}/inv_cov"]), what


def test_host_frames_reproduce_mapbuild_golden():
    from ndtpso_slam_b200 import frames
    s, S = CFG.sensor, CFG.map_size_m
    ref = frames.Frame(width=S, height=S, cell_side=CFG.cell_side)
    for k in range(24):
        f = frames.Frame(width=S, height=S, cell_side=float(S), calculate_cells_params=False)
        f.load_laser(Z["mapbuild/ranges"][k], s.angle_min, s.angle_increment, s.range_max)
        n = int(Z["mapbuild/counts"][k])
        assert np.array_equal(f.scan_points(), Z["mapbuild/points"][k][:n]), k
        ref.update(Z["mapbuild/poses"][k], f)
        if k % 3 == 2:
            ref.build()
        if k in (11, 23):
            _check_table(ref.map_table(), f"mapbuild/table{k}", k)


@pytest.mark.gpu
def test_device_frames_reproduce_mapbuild_golden(ctx):
    from ndtpso_slam_b200.dframes import DeviceFrames
    s, S = CFG.sensor, CFG.map_size_m
    df = DeviceFrames(ctx, 1, S, S, CFG.cell_side, s.beams)
    probe = DeviceFrames(ctx, 1, S, S, CFG.cell_side, s.beams)
    for k in range(24):
        n = int(Z["mapbuild/counts"][k])
        # loadLaser on the device: same points, same order, last-place differences only (GPU cos/sin vs glibc)
        probe.load_laser(Z["mapbuild/ranges"][k][None], s.angle_min, s.angle_increment, s.range_max)
        got, want = probe.download_scan(0), Z["mapbuild/points"][k][:n]
        assert got.shape == want.shape, k
        assert np.all(np.abs(got - want) <= 4 * np.spacing(np.maximum(np.abs(got), np.abs(want)))), k
        # update + build from the golden points: bit-identical tables
        df.set_scan_points([want])
        df.update([Z["mapbuild/poses"][k]])
        if k % 3 == 2:
            df.build()
        if k in (11, 23):
            _check_table(df.download_map(0), f"mapbuild/table{k}", k)
    assert not df.status().any()
    df.close()
    probe.close()


@pytest.mark.gpu
def test_device_frames_reproduce_track_golden(ctx):
    from ndtpso_slam_b200 import capi
    from ndtpso_slam_b200.dframes import DeviceFrames, RNG_CONTINUE
    s, S = CFG.sensor, CFG.map_size_m
    P, I = (int(v) for v in Z["track/pso"])
    conf = capi.PsoConfig.make(population=P, iterations=I)
    n = 3  # three robots fed the same scans: each owns its rand() stream, so all three must give the golden poses
    df = DeviceFrames(ctx, n, S, S, CFG.cell_side, s.beams)
    init = np.tile(Z["track/initial"], (n, 1))
    for k in range(len(Z["track/ranges"])):
        pose, _ = df.track_step(np.tile(Z["track/ranges"][k], (n, 1)), s.angle_min, s.angle_increment, s.range_max,
                                initial_poses=init, conf=conf, rng_mode=RNG_CONTINUE)
        assert np.abs(pose - Z["track/poses"][k]).max() <= POSE_ATOL, (k, pose, Z["track/poses"][k])
        assert np.array_equal(pose[0], pose[1]) and np.array_equal(pose[0], pose[2])  # deterministic, batch-position independent
    df.build()
    got = df.download_map(2)
    idx = Z["track/table/cell_index"]
    assert np.array_equal(np.nonzero(got["built"])[0], idx)
    assert np.allclose(got["mean"][idx], Z["track/table/mean"], rtol=0, atol=1e-6)
    df.close()
