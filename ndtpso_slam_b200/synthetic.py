"""Synthetic 2-D LiDAR workload ("scene A" of SURVEY.md section 8d).

The reference ships no data, bag files or fixtures, so every
benchmark and test input is a synthetic `sensor_msgs/LaserScan`-shaped array made
here: an axis-aligned room with one box obstacle, analytic ray casting, +-1 cm
uniform range noise from a 32-bit LCG, ranges stored as float32 exactly as a
LaserScan message would carry them (the reference's NDTFrame::loadLaser takes
`vector<float>`, lib/ndtpso_slam/ndtframe.cpp:144).

This module only produces *sensor data*.  Turning ranges into scan points and NDT
cells is the job of whichever NDTFrame implementation is under test.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass(frozen=True)
class Sensor:
    beams: int
    fov_deg: float
    range_max: float

    @property
    def fov(self) -> float:
        return math.radians(self.fov_deg)

    @property
    def angle_min(self) -> np.float32:
        return np.float32(-self.fov / 2.0)

    @property
    def angle_increment(self) -> np.float32:
        return np.float32(self.fov / (self.beams - 1))


#: SICK-LMS-like front scanner (BASELINE.json configs[0])
SENSOR_361 = Sensor(beams=361, fov_deg=180.0, range_max=30.0)
#: Hokuyo UTM-30LX (BASELINE.json configs[1..4])
SENSOR_1081 = Sensor(beams=1081, fov_deg=270.0, range_max=30.0)


@dataclasses.dataclass(frozen=True)
class MatchConfig:
    """One BASELINE.json configuration of the scan-matching path."""

    name: str
    sensor: Sensor
    map_size_m: int  # NDTFrame width == height, metres (uint16 in the reference)
    cell_side: float
    particles: int
    iterations: int


CFG1 = MatchConfig("cfg1", SENSOR_361, 20, 1.0, 30, 20)
CFG2 = MatchConfig("cfg2", SENSOR_1081, 50, 0.5, 70, 50)
CFG5 = {cs: MatchConfig(f"cfg5_{cs}", SENSOR_1081, 50, cs, 200, 100) for cs in (0.25, 0.5, 1.0, 2.0)}
#: what NDTFrame::align() really runs (default PSOConfig, ndtframe.cpp:257 + config.h:20-30)
CFG_ALIGN_DEFAULT = MatchConfig("align_default", SENSOR_1081, 50, 0.5, 30, 50)

DEFAULT_GUESS = (0.20, 0.08, 0.04)
DEFAULT_TRUE_POSE = (0.27, 0.11, 0.055)
#: NDTFrame::align's deviation for its first two calls (ndtframe.cpp:253)
DEFAULT_DEVIATION = (0.1, 0.1, 3.1415e-3)


class NoiseLCG:
    """s = s*1664525 + 1013904223 mod 2^32; u = (s >> 8) / 2^24."""

    def __init__(self, state: int = 12345):
        self.state = state & 0xFFFFFFFF

    def uniform(self, n: int) -> np.ndarray:
        """The next n draws.  s_k = a^k s_0 + c (1 + a + ... + a^(k-1)) mod 2^32, evaluated with wrapping
        uint32 accumulations (bit-identical to stepping the recurrence n times)."""
        if n <= 0:
            return np.empty(0, dtype=np.float64)
        u32 = np.uint32
        a_pow = np.multiply.accumulate(np.full(n, 1664525, dtype=u32), dtype=u32)  # a^1 .. a^n (wrapping)
        ones = np.empty(n, dtype=u32)
        ones[0] = 1
        ones[1:] = a_pow[:-1]
        geo = np.add.accumulate(ones, dtype=u32)  # 1 + a + .. + a^(k-1)
        with np.errstate(over="ignore"):
            s = a_pow * u32(self.state) + u32(1013904223) * geo
        self.state = int(s[-1])
        return (s >> np.uint32(8)).astype(np.float64) / 16777216.0


class Room:
    """Room [-0.4S, 0.4S] x [-0.3S, 0.3S] with a box obstacle [3,5] x [2,4]."""

    def __init__(self, map_size_m: float):
        s = float(map_size_m)
        self.xmin, self.xmax = -0.4 * s, 0.4 * s
        self.ymin, self.ymax = -0.3 * s, 0.3 * s
        self.box = (3.0, 5.0, 2.0, 4.0)

    def cast(self, ox: float, oy: float, angles: np.ndarray) -> np.ndarray:
        dx, dy = np.cos(angles), np.sin(angles)
        with np.errstate(divide="ignore", invalid="ignore"):
            tx = np.where(dx > 0, (self.xmax - ox) / dx, np.where(dx < 0, (self.xmin - ox) / dx, np.inf))
            ty = np.where(dy > 0, (self.ymax - oy) / dy, np.where(dy < 0, (self.ymin - oy) / dy, np.inf))
            t = np.minimum(tx, ty)
            bx0, bx1, by0, by1 = self.box
            ax0, ax1 = (bx0 - ox) / dx, (bx1 - ox) / dx
            ay0, ay1 = (by0 - oy) / dy, (by1 - oy) / dy
            t_in = np.maximum(np.minimum(ax0, ax1), np.minimum(ay0, ay1))
            t_out = np.minimum(np.maximum(ax0, ax1), np.maximum(ay0, ay1))
        hit = (t_in <= t_out) & (t_in > 0) & np.isfinite(t_in)
        return np.where(hit, np.minimum(t, t_in), t)


def make_scan(room: Room, sensor: Sensor, pose, noise: NoiseLCG) -> np.ndarray:
    """float32 ranges of one scan taken at `pose` = (x, y, theta)."""
    i = np.arange(sensor.beams, dtype=np.float64)
    angles = pose[2] - sensor.fov / 2.0 + i * (sensor.fov / (sensor.beams - 1))
    t = room.cast(pose[0], pose[1], angles)
    u = noise.uniform(sensor.beams)
    return (t + 0.02 * (u - 0.5)).astype(np.float32)


def map_poses(k0: int = 0, count: int = 5):
    """Poses the map scans are taken from: (0.05k, 0.02k, 0.01k)."""
    return [(0.05 * k, 0.02 * k, 0.01 * k) for k in range(k0, k0 + count)]


@dataclasses.dataclass
class ScanSet:
    """Sensor data of one scan-match problem: the scans merged into the map (with the
    poses they are merged at) and the query scan to be aligned."""

    cfg: MatchConfig
    map_scans: list  # list[(pose, ranges float32)]
    query_ranges: np.ndarray
    true_pose: tuple
    guess: tuple
    deviation: tuple


def scene_a(cfg: MatchConfig, true_pose=DEFAULT_TRUE_POSE, guess=DEFAULT_GUESS, deviation=DEFAULT_DEVIATION,
            poses=None, noise_seed: int = 12345) -> ScanSet:
    room = Room(cfg.map_size_m)
    noise = NoiseLCG(noise_seed)
    poses = map_poses() if poses is None else poses
    scans = [(p, make_scan(room, cfg.sensor, p, noise)) for p in poses]
    query = make_scan(room, cfg.sensor, true_pose, noise)
    return ScanSet(cfg, scans, query, tuple(true_pose), tuple(guess), tuple(deviation))


def trajectory_problem(cfg: MatchConfig, b: int) -> ScanSet:
    """Problem b of a replayed trajectory (BASELINE.json configs[2], configs[3]): the robot
    advances 5 cm per scan along x (wrapping every 2 m), drifting in y and heading; the map is
    built from the 5 preceding poses, the guess trails the true pose like DEFAULT_GUESS does."""
    step = b % 40
    bx = 0.05 * step
    by = 0.02 * math.sin(0.37 * b)
    bth = 0.01 * math.cos(0.23 * b)
    poses = [(bx + 0.05 * k, by + 0.02 * k, bth + 0.01 * k) for k in range(5)]
    true_pose = (bx + DEFAULT_TRUE_POSE[0], by + DEFAULT_TRUE_POSE[1], bth + DEFAULT_TRUE_POSE[2])
    guess = (bx + DEFAULT_GUESS[0], by + DEFAULT_GUESS[1], bth + DEFAULT_GUESS[2])
    return scene_a(cfg, true_pose=true_pose, guess=guess, poses=poses, noise_seed=12345 + 7919 * b)
