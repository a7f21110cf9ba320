"""examples/replay_node.cpp: the reference node's per-scan callback written against the reference's API (NDTFrame, PSOConfig),
compiled against the drop-in headers and libndtpso_slam.so.

CPU: it compiles and links (source compatibility of the drop-in with code written for the reference).
GPU: fed the golden track's scans it prints the golden poses (tests/golden/map_vectors.npz, made from the unmodified
reference with the same callback sequence, 30 particles x 20 iterations, the rand() stream of a never-seeded process)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ndtpso_slam_b200 import build as nbuild, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    nbuild.build()
    exe = str(tmp_path / "replay_node")
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(nbuild.SHIM, "include"), "-I" + os.path.join(ROOT, "include"),
           "-I" + nbuild.eigen_include(), "-o", exe, os.path.join(ROOT, "examples", "replay_node.cpp"),
           "-L" + nbuild.LIB_DIR, "-lndtpso_slam", "-lndtpso_b200", "-Wl,-rpath," + nbuild.LIB_DIR]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return exe


def test_replay_node_compiles_against_the_drop_in(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 2 and "usage" in res.stderr


@pytest.mark.gpu
def test_replay_node_reproduces_the_reference_track(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "map_vectors.npz"))
    cfg, s = syn.CFG1, syn.CFG1.sensor
    ranges = z["track/ranges"]
    path = str(tmp_path / "scans.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<iifff", ranges.shape[0], ranges.shape[1], float(s.angle_min), float(s.angle_increment), float(s.range_max)))
        f.write(struct.pack("<3d", *[float(v) for v in z["track/initial"]]))
        f.write(struct.pack("<id", int(cfg.map_size_m), float(cfg.cell_side)))
        f.write(np.ascontiguousarray(ranges, dtype=np.float32).tobytes())
    P, I = (int(v) for v in z["track/pso"])
    res = subprocess.run([_build(tmp_path), path, str(I), str(P)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = np.array([[float(v) for v in line.split()] for line in res.stdout.strip().splitlines()])
    assert got.shape == z["track/poses"].shape
    assert np.abs(got - z["track/poses"]).max() <= 1e-4, (got, z["track/poses"])  # BASELINE.json north star: 1e-4 on the pose
