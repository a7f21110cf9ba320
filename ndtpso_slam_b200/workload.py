"""Benchmark workloads: batches of flat scan-match problems (BASELINE.json configs)."""
from __future__ import annotations

import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_GEOM_KEYS = ["w_cells", "h_cells", "width_m", "height_m", "cell_side", "x_min", "x_max", "y_min", "y_max"]


def _fixture_flat(z, name):
    g = z[f"{name}/geom"]
    f = {k: (int(v) if k in ("w_cells", "h_cells") else float(v)) for k, v in zip(_GEOM_KEYS, g)}
    n = f["w_cells"] * f["h_cells"]
    idx = z[f"{name}/cell_index"]
    f["points"] = np.ascontiguousarray(z[f"{name}/points"])
    f["mean"] = np.zeros((n, 2))
    f["inv_cov"] = np.zeros((n, 4))
    f["built"] = np.zeros(n, dtype=np.uint8)
    f["mean"][idx] = z[f"{name}/mean"]
    f["inv_cov"][idx] = z[f"{name}/inv_cov"]
    f["built"][idx] = 1
    f["guess"] = z[f"{name}/guess"]
    f["deviation"] = z[f"{name}/deviation"]
    return f


def cfg2_batch(batch: int, first: int = 0):
    """`batch` problems of BASELINE.json configs[1] shape (1081-beam scan, 50 m / 0.5 m map).

    Interim generator: the recorded cfg2 / traj17 scenes (tests/golden), every problem with its OWN
    copies of the table arrays (no sharing) and its own seed 1 + b.
    """
    z = np.load(os.path.join(_ROOT, "tests", "golden", "ref_vectors.npz"))
    bases = [_fixture_flat(z, "cfg2"), _fixture_flat(z, "traj17")]
    out = []
    for b in range(first, first + batch):
        src = bases[b % len(bases)]
        f = dict(src)
        for k in ("points", "mean", "inv_cov", "built"):
            f[k] = src[k].copy()
        f["seed"] = 1 + b
        out.append(f)
    return out
