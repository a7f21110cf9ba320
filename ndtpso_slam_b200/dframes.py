"""ctypes binding of include/ndtpso_dframes.h: reference frames that live in HBM.

`DeviceFrames` mirrors what the reference's ROS node does with its frames
(src/ndtpso_slam_node.cpp:177-244) for n independent robots at once:

    df.load_laser(ranges, angle_min, angle_increment, range_max)   # NDTFrame::loadLaser
    pose = df.align(guess)                                          # NDTFrame::align
    df.update(pose)                                                 # NDTFrame::update
or  pose = df.track_step(ranges, ...)                               # the whole callback

Every call goes through the C ABI of libndtpso_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

RNG_SEEDED, RNG_CONTINUE = 0, 1
DF_CELL_POOL_FULL, DF_WINDOW_TRUNCATED, DF_INDEX_PAST_END, DF_IRREGULAR_SIGMA = 1, 2, 4, 8
DF_NO_CLUSTER = 1  # creation flag

#: every symbol include/ndtpso_dframes.h declares
EXPORTS = [
    "ndtpso_dframes_config_default", "ndtpso_dframes_create", "ndtpso_dframes_destroy", "ndtpso_dframes_device_bytes",
    "ndtpso_dframes_load_laser", "ndtpso_dframes_set_scan_points", "ndtpso_dframes_update", "ndtpso_dframes_build",
    "ndtpso_dframes_align", "ndtpso_dframes_track_step", "ndtpso_dframes_download_map", "ndtpso_dframes_download_scan",
    "ndtpso_dframes_info", "ndtpso_dframes_status", "ndtpso_dframes_pso_stats", "ndtpso_dframes_kernel_times",
    "ndtpso_dframes_attach_exchange", "ndtpso_dframes_load_laser_binned", "ndtpso_dframes_align_streams",
]


class DFramesConfig(C.Structure):
    """struct ndtpso_dframes_config."""
    _fields_ = [("n_frames", C.c_int32), ("width_m", C.c_int32), ("height_m", C.c_int32), ("max_beams", C.c_int32),
                ("cell_side", C.c_double), ("scan_cell_side", C.c_double), ("max_cells", C.c_int32), ("window_points", C.c_int32),
                ("laser_ignore_epsilon", C.c_float), ("flags", C.c_int32)]


_bound = False


def _lib():
    global _bound
    L = capi.load_library()
    if not _bound:
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.ndtpso_dframes_config_default.argtypes = [C.POINTER(DFramesConfig)]
        L.ndtpso_dframes_config_default.restype = None
        L.ndtpso_dframes_create.argtypes = [vp, C.POINTER(DFramesConfig), C.POINTER(vp)]
        L.ndtpso_dframes_destroy.argtypes = [vp]
        L.ndtpso_dframes_destroy.restype = None
        L.ndtpso_dframes_device_bytes.argtypes = [vp]
        L.ndtpso_dframes_device_bytes.restype = C.c_int64
        L.ndtpso_dframes_load_laser.argtypes = [vp, vp, i32, f32, f32, f32, vp]
        L.ndtpso_dframes_set_scan_points.argtypes = [vp, vp, vp, i32]
        L.ndtpso_dframes_update.argtypes = [vp, vp]
        L.ndtpso_dframes_build.argtypes = [vp]
        L.ndtpso_dframes_align.argtypes = [vp, vp, C.POINTER(capi.PsoConfig), i32, vp, vp, vp]
        L.ndtpso_dframes_track_step.argtypes = [vp, vp, i32, f32, f32, f32, vp, C.POINTER(capi.PsoConfig), i32, vp, vp, vp]
        L.ndtpso_dframes_download_map.argtypes = [vp, i32, vp, vp, vp]
        L.ndtpso_dframes_download_scan.argtypes = [vp, i32, vp, C.POINTER(i32)]
        L.ndtpso_dframes_info.argtypes = [vp, i32, vp]
        L.ndtpso_dframes_status.argtypes = [vp, vp]
        L.ndtpso_dframes_pso_stats.argtypes = [vp, vp]
        L.ndtpso_dframes_attach_exchange.argtypes = [vp, vp]
        L.ndtpso_dframes_kernel_times.argtypes = [vp, vp]
        _bound = True
    return L


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class DeviceFrames:
    """n reference frames resident on the GPU of `ctx` (a capi.Context)."""

    def __init__(self, ctx, n_frames, width_m, height_m, cell_side, max_beams, scan_cell_side=0.0, max_cells=0, window_points=0,
                 laser_ignore_epsilon=0.1, flags=0):
        self.L = _lib()
        self.ctx = ctx
        cfg = DFramesConfig()
        self.L.ndtpso_dframes_config_default(C.byref(cfg))
        cfg.n_frames, cfg.width_m, cfg.height_m, cfg.max_beams = int(n_frames), int(width_m), int(height_m), int(max_beams)
        cfg.cell_side, cfg.scan_cell_side = float(cell_side), float(scan_cell_side)
        cfg.max_cells, cfg.window_points = int(max_cells), int(window_points)
        cfg.laser_ignore_epsilon = float(laser_ignore_epsilon)
        cfg.flags = int(flags)
        self.cfg = cfg
        self.n = int(n_frames)
        self.h = C.c_void_p()
        self._check(self.L.ndtpso_dframes_create(ctx.h, C.byref(cfg), C.byref(self.h)))

    def _check(self, rc):
        if rc != capi.OK:
            raise capi.NdtpsoError(rc, self.L.ndtpso_last_error(self.ctx.h).decode())

    def close(self):
        if self.h:
            self.L.ndtpso_dframes_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_bytes(self) -> int:
        return int(self.L.ndtpso_dframes_device_bytes(self.h))

    # ---- the reference's frame operations
    def load_laser(self, ranges, angle_min, angle_increment, range_max, scan_trans=None):
        r = np.ascontiguousarray(ranges, dtype=np.float32).reshape(self.n, -1)
        t = None if scan_trans is None else np.ascontiguousarray(scan_trans, dtype=np.float64).reshape(self.n, 3)
        self._check(self.L.ndtpso_dframes_load_laser(self.h, _p(r), r.shape[1], C.c_float(float(angle_min)),
                                                     C.c_float(float(angle_increment)), C.c_float(float(range_max)), _p(t)))

    def set_scan_points(self, points_list):
        """points_list: one [k_b, 2] float64 array per frame."""
        stride = max([len(p) for p in points_list] + [1])
        xy = np.zeros((self.n, stride, 2), dtype=np.float64)
        cnt = np.zeros(self.n, dtype=np.int32)
        for b, p in enumerate(points_list):
            p = np.asarray(p, dtype=np.float64).reshape(-1, 2)
            xy[b, :len(p)] = p
            cnt[b] = len(p)
        self._check(self.L.ndtpso_dframes_set_scan_points(self.h, _p(xy), _p(cnt), stride))

    def update(self, poses=None):
        p = None if poses is None else np.ascontiguousarray(poses, dtype=np.float64).reshape(self.n, 3)
        self._check(self.L.ndtpso_dframes_update(self.h, _p(p)))

    def build(self):
        self._check(self.L.ndtpso_dframes_build(self.h))

    def align(self, guess=None, conf=None, rng_mode=RNG_CONTINUE, seeds=None, fetch=True):
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float64).reshape(self.n, 3)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32).reshape(self.n)
        pose = np.empty((self.n, 3), dtype=np.float64) if fetch else None
        cost = np.empty(self.n, dtype=np.float64) if fetch else None
        cp = C.byref(conf) if conf is not None else None
        self._check(self.L.ndtpso_dframes_align(self.h, _p(g), cp, int(rng_mode), _p(s), _p(pose), _p(cost)))
        return pose, cost

    def track_step(self, ranges, angle_min, angle_increment, range_max, initial_poses=None, conf=None, rng_mode=RNG_CONTINUE, seeds=None):
        r = np.ascontiguousarray(ranges, dtype=np.float32).reshape(self.n, -1)
        ip = None if initial_poses is None else np.ascontiguousarray(initial_poses, dtype=np.float64).reshape(self.n, 3)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32).reshape(self.n)
        pose = np.empty((self.n, 3), dtype=np.float64)
        cost = np.empty(self.n, dtype=np.float64)
        cp = C.byref(conf) if conf is not None else None
        self._check(self.L.ndtpso_dframes_track_step(self.h, _p(r), r.shape[1], C.c_float(float(angle_min)), C.c_float(float(angle_increment)),
                                                     C.c_float(float(range_max)), _p(ip), cp, int(rng_mode), _p(s), _p(pose), _p(cost)))
        return pose, cost

    # ---- read-back
    def info(self, frame=0) -> dict:
        out = np.zeros(8, dtype=np.int32)
        self._check(self.L.ndtpso_dframes_info(self.h, int(frame), _p(out)))
        keys = ["w_cells", "h_cells", "n_cells", "created", "built", "scan_points", "align_calls", "status"]
        return dict(zip(keys, (int(v) for v in out)))

    def download_map(self, frame=0) -> dict:
        n = self.info(frame)["n_cells"]
        mean = np.zeros((n, 2), dtype=np.float64)
        icov = np.zeros((n, 4), dtype=np.float64)
        built = np.zeros(n, dtype=np.uint8)
        self._check(self.L.ndtpso_dframes_download_map(self.h, int(frame), _p(mean), _p(icov), _p(built)))
        return {"mean": mean, "inv_cov": icov, "built": built}

    def download_scan(self, frame=0) -> np.ndarray:
        cap = C.c_int32(int(self.cfg.max_beams))
        xy = np.zeros((int(self.cfg.max_beams), 2), dtype=np.float64)
        self._check(self.L.ndtpso_dframes_download_scan(self.h, int(frame), _p(xy), C.byref(cap)))
        return xy[:cap.value].copy()

    def status(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.int32)
        self._check(self.L.ndtpso_dframes_status(self.h, _p(out)))
        return out

    def attach_exchange(self, ex):
        """Every later align / track_step also publishes its poses through `ex` (a capi.Exchange with n_per_rank == n_frames)."""
        self._check(self.L.ndtpso_dframes_attach_exchange(self.h, ex.h if ex is not None else None))

    def pso_stats(self) -> np.ndarray:
        """[n, 4] of the last align: rounds, gbest updates, fp64 cost evaluations, evaluations settled by the fp32 screen."""
        out = np.zeros((self.n, 4), dtype=np.int32)
        self._check(self.L.ndtpso_dframes_pso_stats(self.h, _p(out)))
        return out

    def kernel_times_ms(self) -> dict:
        out = np.zeros(5, dtype=np.float64)
        self._check(self.L.ndtpso_dframes_kernel_times(self.h, _p(out)))
        return dict(zip(["load_laser", "update", "build", "rng", "pso"], (float(v) for v in out)))
