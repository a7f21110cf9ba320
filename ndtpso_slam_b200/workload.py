"""Benchmark workloads: batches of flat scan-match problems (BASELINE.json configs), generated with
the drop-in NDTFrame (host map building) from synthetic LaserScan data."""
from __future__ import annotations

from . import frames, synthetic


def cfg2_batch(batch: int, first: int = 0, sparse: bool = False):
    """`batch` problems of BASELINE.json configs[1] shape (1081-beam scan, 50 m / 0.5 m map, to be
    solved with 70 particles x 50 iterations): problem b is step b of the replayed trajectory
    (synthetic.trajectory_problem), with its OWN dense table (no sharing) and seed 1 + b."""
    return [frames.problem_from_scans(synthetic.trajectory_problem(synthetic.CFG2, b), sparse=sparse, seed=1 + b)
            for b in range(first, first + batch)]


def config_problem(cfg: synthetic.MatchConfig, seed: int = 1, sparse: bool = False):
    """Scene A of one BASELINE.json configuration."""
    return frames.problem_from_scans(synthetic.scene_a(cfg), sparse=sparse, seed=seed)
