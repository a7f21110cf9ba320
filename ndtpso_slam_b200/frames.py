"""ctypes binding of include/ndtpso_frames.h (libndtpso_slam.so): the drop-in NDTFrame from Python.

`Frame` mirrors the reference's NDTFrame (constructor arguments, loadLaser/update/build/align);
`problem_from_scans` builds one flat scan-match problem the way the reference's ROS node builds
its frames (src/ndtpso_slam_node.cpp:64-78,186-230).  Map building is host code; align runs on the
GPU and raises without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from . import capi

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    capi.load_library()  # dependency first (also builds when missing)
    if not os.path.exists(_build.SHIM_LIB):
        _build.build_shim()
    L = C.CDLL(_build.SHIM_LIB)
    L.ndtpso_frame_new.restype = C.c_void_p
    L.ndtpso_frame_new.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_int]
    L.ndtpso_frame_free.argtypes = [C.c_void_p]
    L.ndtpso_frame_free.restype = None
    L.ndtpso_frame_load_laser.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_float]
    L.ndtpso_frame_load_laser.restype = None
    L.ndtpso_frame_update.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    L.ndtpso_frame_update.restype = None
    L.ndtpso_frame_build.argtypes = [C.c_void_p]
    L.ndtpso_frame_build.restype = None
    L.ndtpso_frame_is_built.argtypes = [C.c_void_p]
    L.ndtpso_frame_reset_cells.argtypes = [C.c_void_p]
    L.ndtpso_frame_reset_cells.restype = None
    L.ndtpso_frame_map_view.argtypes = [C.c_void_p, C.POINTER(capi.MapView)]
    L.ndtpso_frame_map_view.restype = None
    L.ndtpso_frame_sparse_map_view.argtypes = [C.c_void_p, C.POINTER(capi.MapView)]
    L.ndtpso_frame_sparse_map_view.restype = None
    L.ndtpso_frame_scan_points.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_double))]
    L.ndtpso_frame_point_count.argtypes = [C.c_void_p]
    L.ndtpso_frame_point_count.restype = C.c_int64
    L.ndtpso_frame_align.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_double)]
    L.ndtpso_frame_glir.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ndtpso_frame_align_conf.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(capi.PsoConfig), C.POINTER(C.c_double)]
    L.ndtpso_frame_cost.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ndtpso_frame_add_pose.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    L.ndtpso_frame_add_pose.restype = None
    L.ndtpso_frame_dump_map.argtypes = [C.c_void_p, C.c_char_p]
    L.ndtpso_frame_dump_map.restype = None
    L.ndtpso_frame_device_resident.argtypes = [C.c_void_p]
    L.ndtpso_frame_last_h2d_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ndtpso_frame_last_h2d_bytes.restype = None
    L.ndtpso_frame_download_device_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_frame_draw_rand.argtypes = [C.POINTER(C.c_int32), C.c_int64]
    L.ndtpso_frame_draw_rand.restype = None
    L.ndtpso_frame_set_failure_mode.argtypes = [C.c_int]
    L.ndtpso_frame_set_failure_mode.restype = None
    L.ndtpso_frame_last_cost.restype = C.c_double
    L.ndtpso_frame_last_error.restype = C.c_char_p
    _lib = L
    return L


#: every symbol include/ndtpso_frames.h declares
EXPORTS = ["ndtpso_frame_new", "ndtpso_frame_free", "ndtpso_frame_load_laser", "ndtpso_frame_update", "ndtpso_frame_build",
           "ndtpso_frame_is_built", "ndtpso_frame_reset_cells", "ndtpso_frame_map_view", "ndtpso_frame_sparse_map_view", "ndtpso_frame_scan_points",
           "ndtpso_frame_point_count", "ndtpso_frame_add_pose", "ndtpso_frame_dump_map", "ndtpso_frame_align", "ndtpso_frame_align_conf", "ndtpso_frame_glir", "ndtpso_frame_cost", "ndtpso_frame_last_cost",
           "ndtpso_frame_device_resident", "ndtpso_frame_last_h2d_bytes", "ndtpso_frame_download_device_map",
           "ndtpso_frame_draw_rand", "ndtpso_frame_set_failure_mode", "ndtpso_frame_last_error"]


def _d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class Frame:
    """NDTFrame(trans, width, height, cell_side, calculate_cells_params)."""

    def __init__(self, trans=(0., 0., 0.), width=20, height=20, cell_side=1.0, calculate_cells_params=True):
        self.lib = load_library()
        self.h = self.lib.ndtpso_frame_new(_d3(trans), int(width), int(height), float(cell_side), int(bool(calculate_cells_params)))
        if not self.h:
            raise MemoryError("ndtpso_frame_new failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.ndtpso_frame_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_laser(self, ranges, angle_min, angle_increment, range_max):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        self.lib.ndtpso_frame_load_laser(self.h, r.ctypes.data_as(C.POINTER(C.c_float)), r.shape[0], C.c_float(float(angle_min)),
                                         C.c_float(float(angle_increment)), C.c_float(float(range_max)))

    def update(self, pose, new_frame: "Frame"):
        self.lib.ndtpso_frame_update(self.h, _d3(pose), new_frame.h)

    def reset_cells(self):
        self.lib.ndtpso_frame_reset_cells(self.h)

    def build(self):
        self.lib.ndtpso_frame_build(self.h)

    @property
    def built(self) -> bool:
        return bool(self.lib.ndtpso_frame_is_built(self.h))

    @property
    def device_resident(self) -> bool:
        """The map is mirrored in HBM and align/update go through the mirror."""
        return bool(self.lib.ndtpso_frame_device_resident(self.h))

    def last_h2d_bytes(self):
        """(align, update): bytes the last align / update moved host->device through the mirror."""
        a, u = C.c_int64(0), C.c_int64(0)
        self.lib.ndtpso_frame_last_h2d_bytes(self.h, C.byref(a), C.byref(u))
        return a.value, u.value

    def device_map_table(self) -> dict:
        """The device's copy of the dense table (mirrored frames only)."""
        g = self._view(False)
        n = g["w_cells"] * g["h_cells"]
        mean, icov, built = np.zeros((n, 2)), np.zeros((n, 4)), np.zeros(n, dtype=np.uint8)
        rc = self.lib.ndtpso_frame_download_device_map(self.h, mean.ctypes.data, icov.ctypes.data, built.ctypes.data)
        if rc != 0:
            raise capi.NdtpsoError(rc, "no device mirror")
        return dict(mean=mean, inv_cov=icov, built=built)

    def add_pose(self, timestamp: float, pose):
        self.lib.ndtpso_frame_add_pose(self.h, float(timestamp), _d3(pose))

    def dump_map(self, filename: str):
        self.lib.ndtpso_frame_dump_map(self.h, filename.encode())

    def point_count(self) -> int:
        return int(self.lib.ndtpso_frame_point_count(self.h))

    def scan_points(self) -> np.ndarray:
        p = C.POINTER(C.c_double)()
        n = self.lib.ndtpso_frame_scan_points(self.h, C.byref(p))
        if n == 0:
            return np.zeros((0, 2))
        return np.ctypeslib.as_array(p, shape=(n, 2)).copy()

    def _view(self, sparse: bool) -> dict:
        v = capi.MapView()
        (self.lib.ndtpso_frame_sparse_map_view if sparse else self.lib.ndtpso_frame_map_view)(self.h, C.byref(v))
        out = {k: getattr(v, k) for k in ("w_cells", "h_cells", "width_m", "height_m", "cell_side", "x_min", "x_max", "y_min", "y_max")}
        rows = v.n_sparse if sparse else v.w_cells * v.h_cells

        def arr(ptr, ctype, shape):
            if rows == 0 or not ptr:
                return np.zeros(shape, dtype=np.dtype(ctype))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape).copy()

        out["mean"] = arr(v.mean, C.c_double, (rows, 2))
        out["inv_cov"] = arr(v.inv_cov, C.c_double, (rows, 4))
        if sparse:
            out["cell_index"] = arr(v.cell_index, C.c_int32, (rows,))
        else:
            out["built"] = arr(v.built, C.c_uint8, (rows,))
        return out

    def map_table(self, sparse: bool = False) -> dict:
        """Copy of the (mean, inv_cov, built | cell_index) table + geometry: a flat-problem map."""
        return self._view(sparse)

    def align(self, guess, new_frame: "Frame", conf: capi.PsoConfig | None = None) -> np.ndarray:
        pose = (C.c_double * 3)()
        if conf is None:
            rc = self.lib.ndtpso_frame_align(self.h, _d3(guess), new_frame.h, pose)
        else:
            rc = self.lib.ndtpso_frame_align_conf(self.h, _d3(guess), new_frame.h, C.byref(conf), pose)
        if rc != 0:
            raise capi.NdtpsoError(rc, (self.lib.ndtpso_frame_last_error() or b"").decode())
        return np.array(list(pose))

    def glir(self, guess, new_frame: "Frame", iterations: int = 50, deviation=(0., 0., 0.)) -> np.ndarray:
        """glir_pso_optimization(guess, this, new_frame, iterations, deviation) of the drop-in library (core.h:21-23)."""
        pose = (C.c_double * 3)()
        rc = self.lib.ndtpso_frame_glir(self.h, _d3(guess), new_frame.h, int(iterations), _d3(deviation), pose)
        if rc != 0:
            raise capi.NdtpsoError(rc, (self.lib.ndtpso_frame_last_error() or b"").decode())
        return np.array(list(pose))

    def cost(self, new_frame: "Frame", pose) -> float:
        out = C.c_double(0.0)
        rc = self.lib.ndtpso_frame_cost(self.h, new_frame.h, _d3(pose), C.byref(out))
        if rc != 0:
            raise capi.NdtpsoError(rc, (self.lib.ndtpso_frame_last_error() or b"").decode())
        return out.value


def frames_from_scans(scanset):
    """(reference frame, query frame) for a synthetic.ScanSet, built the way the ROS node does:
    every scan is loaded into a one-cell frame and merged into the map frame at its pose."""
    cfg, s, S = scanset.cfg, scanset.cfg.sensor, scanset.cfg.map_size_m
    ref = Frame(width=S, height=S, cell_side=cfg.cell_side, calculate_cells_params=True)
    for pose, ranges in scanset.map_scans:
        f = Frame(width=S, height=S, cell_side=float(S), calculate_cells_params=False)
        f.load_laser(ranges, s.angle_min, s.angle_increment, s.range_max)
        ref.update(pose, f)
        f.close()
    ref.build()
    q = Frame(width=S, height=S, cell_side=float(S), calculate_cells_params=False)
    q.load_laser(scanset.query_ranges, s.angle_min, s.angle_increment, s.range_max)
    return ref, q


def problem_from_scans(scanset, sparse: bool = False, seed: int = 1) -> dict:
    """A flat scan-match problem (capi.ProblemSet input) for a synthetic.ScanSet."""
    ref, q = frames_from_scans(scanset)
    flat = ref.map_table(sparse=sparse)
    flat["points"] = q.scan_points()
    flat.update(guess=np.array(scanset.guess), deviation=np.array(scanset.deviation), seed=seed)
    ref.close()
    q.close()
    return flat
