"""Test helpers: golden fixtures -> flat problem dicts (the layout both the oracle and the C ABI take)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")
GEOM_KEYS = ["w_cells", "h_cells", "width_m", "height_m", "cell_side", "x_min", "x_max", "y_min", "y_max"]

#: tolerances stated by BASELINE.json north_star
POSE_ATOL = 1e-4
SCORE_RTOL = 1e-5


class Golden:
    def __init__(self, path=GOLDEN):
        self.z = np.load(path)
        self._flat = {}

    def flat(self, name, sparse=False):
        """Problem inputs `name` as a flat dict; dense table unless sparse=True."""
        key = (name, sparse)
        if key in self._flat:
            return dict(self._flat[key])
        z = self.z
        g = z[f"{name}/geom"]
        f = {k: (int(v) if k in ("w_cells", "h_cells") else float(v)) for k, v in zip(GEOM_KEYS, g)}
        f["points"] = np.ascontiguousarray(z[f"{name}/points"])
        idx = z[f"{name}/cell_index"]
        if sparse:
            f["cell_index"] = np.ascontiguousarray(idx)
            f["mean"] = np.ascontiguousarray(z[f"{name}/mean"])
            f["inv_cov"] = np.ascontiguousarray(z[f"{name}/inv_cov"])
        else:
            n = f["w_cells"] * f["h_cells"]
            f["mean"] = np.zeros((n, 2))
            f["inv_cov"] = np.zeros((n, 4))
            f["built"] = np.zeros(n, dtype=np.uint8)
            f["mean"][idx] = z[f"{name}/mean"]
            f["inv_cov"][idx] = z[f"{name}/inv_cov"]
            f["built"][idx] = 1
        self._flat[key] = f
        return dict(f)

    def case(self, name, inputs=None):
        """A solved case: dict(P, I, coef, seeds, guess, deviation, pose[n,3], cost[n]) + its inputs name."""
        z = self.z
        P, I = (int(v) for v in z[f"{name}/pso"])
        w, c1, c2, wd = (float(v) for v in z[f"{name}/coef"])
        return dict(P=P, I=I, w=w, c1=c1, c2=c2, w_dumping=wd, seeds=[int(s) for s in z[f"{name}/seeds"]],
                    guess=z[f"{name}/guess"], deviation=z[f"{name}/deviation"], pose=z[f"{name}/pose"], cost=z[f"{name}/cost"],
                    inputs=inputs or name)

    def problems(self, case_name, inputs=None, sparse=False):
        """One flat problem per seed of a solved case, sharing the map arrays."""
        c = self.case(case_name, inputs)
        base = self.flat(c["inputs"], sparse=sparse)
        out = []
        for s in c["seeds"]:
            f = dict(base)  # shallow: same numpy arrays -> one table on the device
            f.update(guess=c["guess"], deviation=c["deviation"], seed=s)
            out.append(f)
        return c, out


#: (solved case, inputs it runs on)
SOLVED = [("cfg1", "cfg1"), ("cfg2", "cfg2"), ("align_default", "align_default"), ("cfg5_0.25", "cfg5_0.25"), ("cfg5_0.5", "cfg5_0.5"),
          ("cfg5_1.0", "cfg5_1.0"), ("cfg5_2.0", "cfg5_2.0"), ("traj17", "traj17"), ("edge_zero_dev", "edge"), ("edge_far_guess", "edge"),
          ("edge_one_particle", "edge"), ("edge_no_iterations", "edge"), ("edge_damped", "edge"), ("edge_wide_dev", "edge"),
          ("edge_empty_scan", "edge"), ("np2", "np2")]


BATCH_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batch_vectors.npz")


class BatchGolden:
    """tests/golden/batch_vectors.npz (make_golden_batch.py): the unmodified reference's pose and cost for every trajectory problem
    of BASELINE.json configs[2] (b < 256) and configs[3] (b < 2048), and four more seeds of every configs[4] case."""

    def __init__(self, path=BATCH_GOLDEN):
        z = np.load(path)
        self.pose, self.cost, self.table_crc = z["traj/pose"], z["traj/cost"], z["traj/table_crc"]
        self.P, self.I = (int(v) for v in z["traj/pso"])
        self.z = z

    def cfg5_more(self, cs):
        """(seeds, pose[n, 3], cost[n]) of the extra seeds of the configs[4] case with cell side `cs`."""
        k = f"cfg5_{cs}_more"
        return [int(s) for s in self.z[k + "/seeds"]], self.z[k + "/pose"], self.z[k + "/cost"]


def table_crc(flat) -> int:
    """CRC-32 of a dense flat problem's table and scan, as make_golden_batch.py takes it of the reference's own arrays."""
    import zlib
    c = 0
    for k in ("mean", "inv_cov", "built", "points"):
        c = zlib.crc32(np.ascontiguousarray(flat[k]).tobytes(), c)
    return c


def oracle_pso_many(oracle, flats, P, I, workers=None):
    """The oracle on every problem of `flats` (its own seed each), on `workers` threads (the C restatement is re-entrant and
    ctypes releases the GIL).  Returns pose[n, 3], cost[n]."""
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or os.cpu_count() or 1

    def one(f):
        po, co, _ = oracle.pso(f, f["guess"], f["deviation"], P, I, seed=f["seed"])
        return po, co

    with ThreadPoolExecutor(workers) as ex:
        res = list(ex.map(one, flats))
    return np.array([r[0] for r in res]), np.array([r[1] for r in res])


def empty_points(flat):
    f = dict(flat)
    f["points"] = np.zeros((0, 2))
    return f


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return np.where(a == b, 0.0, np.abs(a - b) / den)
