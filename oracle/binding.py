"""TEST INFRASTRUCTURE — ctypes bindings for the two checker libraries.

  * `Oracle`    -> oracle/_build/libndtpso_oracle.so  (ndtpso_oracle.c, the C restatement)
  * `Reference` -> oracle/_ref/libndtpso_ref.so       (unmodified reference + ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libndtpso_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libndtpso_ref.so")


def build(force: bool = False) -> None:
    """Compile the checkers (the reference one only where /root/reference exists)."""
    args = ["make", "-s", "-C", HERE, "all"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OrcProblem(C.Structure):
    _fields_ = [
        ("n_points", C.c_int32), ("w_cells", C.c_int32), ("h_cells", C.c_int32), ("_pad", C.c_int32),
        ("width_m", C.c_double), ("height_m", C.c_double), ("cell_side", C.c_double),
        ("x_min", C.c_double), ("x_max", C.c_double), ("y_min", C.c_double), ("y_max", C.c_double),
        ("points", C.POINTER(C.c_double)), ("mean", C.POINTER(C.c_double)),
        ("inv_cov", C.POINTER(C.c_double)), ("built", C.POINTER(C.c_uint8)),
    ]


class OrcConfig(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("population", C.c_int32),
                ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("w_dumping", C.c_double)]


class OrcStats(C.Structure):
    _fields_ = [("gbest_updates", C.c_int32), ("pbest_updates", C.c_int32), ("rand_draws", C.c_int32), ("_pad", C.c_int32)]


class OrcAlignState(C.Structure):
    _fields_ = [("s_iter", C.c_int32), ("_pad", C.c_int32), ("s_prev_pose", C.c_double * 3), ("s_pose_diff", C.c_double * 3)]


class Oracle:
    """The C restatement, operating on a flat problem dict (see tests/problems.py)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_cost.restype = C.c_double
        self.lib.orc_cost.argtypes = [C.POINTER(OrcProblem), C.POINTER(C.c_double)]
        self.lib.orc_cost_many.argtypes = [C.POINTER(OrcProblem), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        self.lib.orc_pso.restype = C.c_int
        self.lib.orc_pso.argtypes = [C.POINTER(OrcProblem), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(OrcConfig),
                                     C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(OrcStats)]
        self.lib.orc_glir.restype = C.c_int
        self.lib.orc_glir.argtypes = [C.POINTER(OrcProblem), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int,
                                      C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(OrcStats)]
        self.lib.orc_rand_stream.argtypes = [C.c_uint32, C.POINTER(C.c_int32), C.c_int]
        self.lib.orc_align.restype = C.c_int
        self.lib.orc_align.argtypes = [C.POINTER(OrcAlignState), C.POINTER(OrcProblem), C.POINTER(C.c_double), C.c_uint32,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_double)]

    @staticmethod
    def problem(flat) -> tuple:
        """flat: dict with points[N,2], mean[C,2], inv_cov[C,4], built[C], geometry.
        Returns (OrcProblem, keepalive)."""
        pts = np.ascontiguousarray(flat["points"], dtype=np.float64)
        mean = np.ascontiguousarray(flat["mean"], dtype=np.float64)
        icov = np.ascontiguousarray(flat["inv_cov"], dtype=np.float64)
        built = np.ascontiguousarray(flat["built"], dtype=np.uint8)
        p = OrcProblem(int(pts.shape[0]), int(flat["w_cells"]), int(flat["h_cells"]), 0,
                       float(flat["width_m"]), float(flat["height_m"]), float(flat["cell_side"]),
                       float(flat["x_min"]), float(flat["x_max"]), float(flat["y_min"]), float(flat["y_max"]),
                       _dp(pts), _dp(mean), _dp(icov), built.ctypes.data_as(C.POINTER(C.c_uint8)))
        return p, (pts, mean, icov, built)

    def cost(self, flat, pose) -> float:
        p, _keep = self.problem(flat)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        return float(self.lib.orc_cost(C.byref(p), _dp(pose)))

    def cost_many(self, flat, poses) -> np.ndarray:
        p, _keep = self.problem(flat)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.empty(poses.shape[0], dtype=np.float64)
        self.lib.orc_cost_many(C.byref(p), _dp(poses), poses.shape[0], _dp(out))
        return out

    def rand_stream(self, seed: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        self.lib.orc_rand_stream(seed, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out

    def pso(self, flat, guess, deviation, population, iterations, seed=1, stream=None,
            w=0.8, c1=2.0, c2=2.0, w_dumping=1.0):
        p, _keep = self.problem(flat)
        guess = np.ascontiguousarray(guess, dtype=np.float64)
        dev = np.ascontiguousarray(deviation, dtype=np.float64)
        cf = OrcConfig(int(iterations), int(population), w, c1, c2, w_dumping)
        pose = np.empty(3, dtype=np.float64)
        cost = C.c_double(0.0)
        st = OrcStats()
        sp = None
        if stream is not None:
            stream = np.ascontiguousarray(stream, dtype=np.int32)
            sp = stream.ctypes.data_as(C.POINTER(C.c_int32))
        rc = self.lib.orc_pso(C.byref(p), _dp(guess), _dp(dev), C.byref(cf), int(seed), sp, _dp(pose), C.byref(cost), C.byref(st))
        if rc != 0:
            raise RuntimeError("orc_pso failed")
        return pose, cost.value, {"gbest_updates": st.gbest_updates, "pbest_updates": st.pbest_updates, "rand_draws": st.rand_draws}

    def glir(self, flat, guess, deviation, population, iterations, seed=1, stream=None):
        """glir_pso_optimization (core.cpp:118-186); the reference's population is 30."""
        p, _keep = self.problem(flat)
        guess = np.ascontiguousarray(guess, dtype=np.float64)
        dev = np.ascontiguousarray(deviation, dtype=np.float64)
        pose = np.empty(3, dtype=np.float64)
        cost = C.c_double(0.0)
        st = OrcStats()
        sp = None
        if stream is not None:
            stream = np.ascontiguousarray(stream, dtype=np.int32)
            sp = stream.ctypes.data_as(C.POINTER(C.c_int32))
        rc = self.lib.orc_glir(C.byref(p), _dp(guess), _dp(dev), int(population), int(iterations), int(seed), sp, _dp(pose),
                               C.byref(cost), C.byref(st))
        if rc != 0:
            raise RuntimeError("orc_glir failed")
        return pose, cost.value, {"gbest_updates": st.gbest_updates, "pbest_updates": st.pbest_updates, "rand_draws": st.rand_draws}


class RefFrame:
    def __init__(self, ref: "Reference", trans=(0., 0., 0.), width=20, height=20, cell_side=1.0, init_windows=True):
        self.ref = ref
        self.h = ref.lib.ref_frame_new(trans[0], trans[1], trans[2], int(width), int(height), float(cell_side), int(bool(init_windows)))

    def __del__(self):
        try:
            if self.h:
                self.ref.lib.ref_frame_free(self.h)
                self.h = None
        except Exception:
            pass

    def load_laser(self, ranges, angle_min, angle_inc, range_max):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        self.ref.lib.ref_frame_load_laser(self.h, r.ctypes.data_as(C.POINTER(C.c_float)), r.shape[0],
                                          C.c_float(float(angle_min)), C.c_float(float(angle_inc)), C.c_float(float(range_max)))

    def update(self, pose, new_frame: "RefFrame"):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.ref.lib.ref_frame_update(self.h, _dp(p), new_frame.h)

    def build(self):
        self.ref.lib.ref_frame_build(self.h)

    def geometry(self) -> dict:
        gi = (C.c_int32 * 5)()
        gd = (C.c_double * 5)()
        self.ref.lib.ref_frame_geometry(self.h, gi, gd)
        return {"w_cells": gi[0], "h_cells": gi[1], "n_cells": gi[2], "width_m": float(gi[3]), "height_m": float(gi[4]),
                "cell_side": gd[0], "x_min": gd[1], "x_max": gd[2], "y_min": gd[3], "y_max": gd[4]}

    def flatten_map(self) -> dict:
        g = self.geometry()
        n = g["n_cells"]
        mean = np.zeros((n, 2), dtype=np.float64)
        icov = np.zeros((n, 4), dtype=np.float64)
        built = np.zeros(n, dtype=np.uint8)
        self.ref.lib.ref_frame_flatten_map(self.h, _dp(mean), _dp(icov), built.ctypes.data_as(C.POINTER(C.c_uint8)))
        g.update(mean=mean, inv_cov=icov, built=built)
        return g

    def flatten_points(self) -> np.ndarray:
        n = self.ref.lib.ref_frame_count_points(self.h)
        xy = np.zeros((n, 2), dtype=np.float64)
        if n:
            self.ref.lib.ref_frame_flatten_points(self.h, _dp(xy))
        return xy


class Reference:
    """The unmodified reference library behind ref_harness.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (needs /root/reference to build)")
        self.lib = L = C.CDLL(REF_SO)
        L.ref_frame_new.restype = C.c_void_p
        L.ref_frame_new.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
        L.ref_frame_free.argtypes = [C.c_void_p]
        L.ref_frame_load_laser.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_frame_update.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
        L.ref_frame_build.argtypes = [C.c_void_p]
        L.ref_frame_is_built.argtypes = [C.c_void_p]
        L.ref_frame_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.ref_frame_flatten_map.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
        L.ref_frame_count_points.argtypes = [C.c_void_p]
        L.ref_frame_flatten_points.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ref_frame_count_all_points.argtypes = [C.c_void_p]
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_rand.restype = C.c_int
        L.ref_cost.restype = C.c_double
        L.ref_cost.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_pso.restype = C.c_double
        L.ref_pso.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int,
                              C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_uint, C.POINTER(C.c_double)]
        L.ref_glir.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_uint,
                               C.POINTER(C.c_double)]
        L.ref_align.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_double)]
        L.ref_omp_max_threads.restype = C.c_int
        L.ref_omp_set_num_threads.argtypes = [C.c_int]
        L.ref_omp_set_num_threads.restype = None
        L.ref_sizeof_cell.restype = C.c_int

    def frame(self, **kw) -> RefFrame:
        return RefFrame(self, **kw)

    def srand(self, seed: int):
        self.lib.ref_srand(seed)

    def rand(self) -> int:
        return self.lib.ref_rand()

    def cost(self, ref_frame: RefFrame, cur: RefFrame, pose) -> float:
        p = np.ascontiguousarray(pose, dtype=np.float64)
        return float(self.lib.ref_cost(ref_frame.h, cur.h, _dp(p)))

    def pso(self, ref_frame: RefFrame, cur: RefFrame, guess, deviation, population, iterations, seed=1, use_seed=True,
            num_threads=1, w=0.8, c1=2.0, c2=2.0, w_dumping=1.0):
        g = np.ascontiguousarray(guess, dtype=np.float64)
        d = np.ascontiguousarray(deviation, dtype=np.float64)
        pose = np.empty(3, dtype=np.float64)
        secs = self.lib.ref_pso(ref_frame.h, cur.h, _dp(g), _dp(d), int(population), int(iterations), int(num_threads),
                                w, c1, c2, w_dumping, int(bool(use_seed)), int(seed), _dp(pose))
        return pose, secs

    def glir(self, ref_frame: RefFrame, cur: RefFrame, guess, deviation, iterations, seed=1, use_seed=True):
        """glir_pso_optimization (core.cpp:118-186), population PSO_POPULATION_SIZE = 30."""
        g = np.ascontiguousarray(guess, dtype=np.float64)
        d = np.ascontiguousarray(deviation, dtype=np.float64)
        pose = np.empty(3, dtype=np.float64)
        self.lib.ref_glir(ref_frame.h, cur.h, _dp(g), _dp(d), int(iterations), int(bool(use_seed)), int(seed), _dp(pose))
        return pose

    def align(self, ref_frame: RefFrame, guess, cur: RefFrame):
        g = np.ascontiguousarray(guess, dtype=np.float64)
        pose = np.empty(3, dtype=np.float64)
        self.lib.ref_align(ref_frame.h, _dp(g), cur.h, _dp(pose))
        return pose

    # ---- scene helpers: build frames the way the ROS node does (ndtpso_slam_node.cpp:64-78,186-230)
    def build_problem(self, scanset):
        """Reference-built frames for a synthetic.ScanSet. Returns (ref_frame, query_frame)."""
        cfg = scanset.cfg
        s = cfg.sensor
        S = cfg.map_size_m
        ref_frame = self.frame(width=S, height=S, cell_side=cfg.cell_side, init_windows=True)
        for pose, ranges in scanset.map_scans:
            f = self.frame(width=S, height=S, cell_side=float(S), init_windows=False)
            f.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
            ref_frame.update(pose, f)
        ref_frame.build()
        q = self.frame(width=S, height=S, cell_side=float(S), init_windows=False)
        q.load_laser(scanset.query_ranges, s.angle_min, s.angle_increment, s.range_max)
        return ref_frame, q

    def flatten_problem(self, scanset):
        ref_frame, q = self.build_problem(scanset)
        flat = ref_frame.flatten_map()
        flat["points"] = q.flatten_points()
        return flat, ref_frame, q
