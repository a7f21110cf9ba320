"""ctypes binding of include/ndtpso_b200.h (libndtpso_b200.so).

This is the host-side mirror used by tests/, bench.py and smoke(): every call goes through
the C ABI, exactly as a C++ caller would.  There is no fallback: if the library is missing
or no CUDA device is present, `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE, ERR_LIMIT = 0, -1, -2, -3, -4, -5
OPT_WARPS_PER_CTA, OPT_SMEM_BYTES, OPT_CLUSTER, OPT_KERNEL, OPT_POINTS_PER_THREAD, OPT_CANDIDATE_BATCH, OPT_PIPELINE_CHUNKS = 1, 2, 3, 4, 5, 6, 7
OPT_EXCHANGE_TIMEOUT_MS = 8
OPT_HOT_CHUNK = 9
OPT_SCREEN = 10
OPT_HOST_THREADS = 11
KERNEL_AUTO, KERNEL_WARP_PER_PARTICLE, KERNEL_POINT_SLICED = 0, 1, 2

#: every symbol include/ndtpso_b200.h declares
EXPORTS = [
    "ndtpso_abi_version", "ndtpso_pso_config_default", "ndtpso_device_count", "ndtpso_ctx_create", "ndtpso_ctx_destroy",
    "ndtpso_ctx_set_stream", "ndtpso_last_error", "ndtpso_ctx_set_option", "ndtpso_rand_draws", "ndtpso_align_batch", "ndtpso_align_submit", "ndtpso_align_collect",
    "ndtpso_cost_batch", "ndtpso_screen_bounds", "ndtpso_batch_create", "ndtpso_batch_solve", "ndtpso_batch_device_results", "ndtpso_batch_results",
    "ndtpso_batch_stats", "ndtpso_batch_stats_ex", "ndtpso_batch_kernel_times", "ndtpso_batch_destroy", "ndtpso_ctx_launch_count", "ndtpso_ctx_last_transfer_bytes", "ndtpso_ctx_synchronize", "ndtpso_measure_fp64_peak",
    "ndtpso_exchange_create", "ndtpso_exchange_connect", "ndtpso_exchange_connect_local", "ndtpso_batch_attach_exchange", "ndtpso_exchange_wait",
    "ndtpso_exchange_device_results", "ndtpso_exchange_results", "ndtpso_exchange_destroy",
    "ndtpso_multi_create", "ndtpso_multi_destroy", "ndtpso_multi_size", "ndtpso_multi_ctx", "ndtpso_multi_last_error", "ndtpso_align_batch_multi",
    "ndtpso_align_submit_multi", "ndtpso_align_collect_multi", "ndtpso_multi_batch_create", "ndtpso_multi_batch_solve", "ndtpso_multi_batch_results",
    "ndtpso_multi_batch_device_results", "ndtpso_multi_batch_destroy",
]
IPC_HANDLE_BYTES, MAX_RANKS = 64, 8


VARIANT_PSO, VARIANT_GLIR = 0, 1  # ndtpso_pso_config::variant


class PsoConfig(C.Structure):
    """struct ndtpso_pso_config == reference PSOConfig (include/ndtpso_slam/config.h:27-38)."""
    _fields_ = [("iterations", C.c_int32), ("population", C.c_int32), ("num_threads", C.c_int32), ("variant", C.c_int32),
                ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("w_dumping", C.c_double)]

    @classmethod
    def make(cls, population=30, iterations=50, w=0.8, c1=2.0, c2=2.0, w_dumping=1.0, variant=0):
        """variant: VARIANT_PSO (pso_optimization, core.cpp:50-116) or VARIANT_GLIR (glir_pso_optimization, core.cpp:118-186)."""
        return cls(int(iterations), int(population), -1, int(variant), w, c1, c2, w_dumping)


class MapView(C.Structure):
    _fields_ = [("w_cells", C.c_int32), ("h_cells", C.c_int32),
                ("width_m", C.c_double), ("height_m", C.c_double), ("cell_side", C.c_double),
                ("x_min", C.c_double), ("x_max", C.c_double), ("y_min", C.c_double), ("y_max", C.c_double),
                ("mean", C.c_void_p), ("inv_cov", C.c_void_p), ("built", C.c_void_p),
                ("n_sparse", C.c_int32), ("reserved", C.c_int32), ("cell_index", C.c_void_p)]


class Problem(C.Structure):
    _fields_ = [("map", MapView), ("points_xy", C.c_void_p), ("n_points", C.c_int32), ("seed", C.c_uint32),
                ("guess", C.c_double * 3), ("deviation", C.c_double * 3), ("rand_stream", C.c_void_p), ("rand_count", C.c_int64)]


class NdtpsoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ndtpso error {code}: {msg}")
        self.code = code


_lib = None


def load_library(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(path)
        _build.build()
    L = C.CDLL(path)
    L.ndtpso_abi_version.restype = C.c_int
    L.ndtpso_device_count.restype = C.c_int
    L.ndtpso_pso_config_default.argtypes = [C.POINTER(PsoConfig)]
    L.ndtpso_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ndtpso_ctx_destroy.argtypes = [C.c_void_p]
    L.ndtpso_ctx_destroy.restype = None
    L.ndtpso_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_last_error.argtypes = [C.c_void_p]
    L.ndtpso_last_error.restype = C.c_char_p
    L.ndtpso_ctx_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.ndtpso_rand_draws.argtypes = [C.POINTER(PsoConfig)]
    L.ndtpso_rand_draws.restype = C.c_int64
    L.ndtpso_align_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.c_void_p, C.c_void_p]
    L.ndtpso_align_submit.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.POINTER(C.c_void_p)]
    L.ndtpso_align_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_cost_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.c_int32, C.c_void_p, C.c_void_p]
    L.ndtpso_screen_bounds.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.c_int32, C.c_void_p, C.c_void_p]
    L.ndtpso_batch_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.POINTER(C.c_void_p)]
    L.ndtpso_batch_solve.argtypes = [C.c_void_p]
    L.ndtpso_batch_device_results.argtypes = [C.c_void_p]
    L.ndtpso_batch_device_results.restype = C.c_void_p
    L.ndtpso_batch_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_batch_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_batch_stats_ex.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_batch_kernel_times.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_batch_destroy.argtypes = [C.c_void_p]
    L.ndtpso_batch_destroy.restype = None
    L.ndtpso_ctx_launch_count.argtypes = [C.c_void_p]
    L.ndtpso_ctx_launch_count.restype = C.c_int64
    L.ndtpso_ctx_last_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ndtpso_ctx_synchronize.argtypes = [C.c_void_p]
    L.ndtpso_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.ndtpso_exchange_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]
    L.ndtpso_exchange_connect.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_exchange_connect_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.ndtpso_batch_attach_exchange.argtypes = [C.c_void_p, C.c_void_p]
    L.ndtpso_exchange_wait.argtypes = [C.c_void_p]
    L.ndtpso_exchange_device_results.argtypes = [C.c_void_p]
    L.ndtpso_exchange_device_results.restype = C.c_void_p
    L.ndtpso_exchange_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_exchange_destroy.argtypes = [C.c_void_p]
    L.ndtpso_exchange_destroy.restype = None
    L.ndtpso_multi_create.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
    L.ndtpso_multi_destroy.argtypes = [C.c_void_p]
    L.ndtpso_multi_destroy.restype = None
    L.ndtpso_multi_size.argtypes = [C.c_void_p]
    L.ndtpso_multi_size.restype = C.c_int32
    L.ndtpso_multi_ctx.argtypes = [C.c_void_p, C.c_int32]
    L.ndtpso_multi_ctx.restype = C.c_void_p
    L.ndtpso_multi_last_error.argtypes = [C.c_void_p]
    L.ndtpso_multi_last_error.restype = C.c_char_p
    L.ndtpso_align_batch_multi.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.c_void_p, C.c_void_p]
    L.ndtpso_align_submit_multi.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.POINTER(C.c_void_p)]
    L.ndtpso_align_collect_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_multi_batch_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(PsoConfig), C.POINTER(C.c_void_p)]
    L.ndtpso_multi_batch_solve.argtypes = [C.c_void_p]
    L.ndtpso_multi_batch_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndtpso_multi_batch_device_results.argtypes = [C.c_void_p, C.c_int32]
    L.ndtpso_multi_batch_device_results.restype = C.c_void_p
    L.ndtpso_multi_batch_destroy.argtypes = [C.c_void_p]
    L.ndtpso_multi_batch_destroy.restype = None
    _lib = L
    return L


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(None)


class ProblemSet:
    """Packs flat problem dicts into a `struct ndtpso_problem[]`, keeping the numpy arrays alive.

    A flat problem dict has: points[N,2], mean[R,2], inv_cov[R,4], built[C] (dense) or
    cell_index[R] (sparse), w_cells, h_cells, width_m, height_m, cell_side, x_min, x_max,
    y_min, y_max, guess[3], deviation[3], seed, and optionally rand_stream (int32[]).
    Problems that are the *same dict object's map arrays* share one table on the device.
    """

    def __init__(self, flats):
        self.n = len(flats)
        self.array = (Problem * max(self.n, 1))()
        self.keep = []
        for i, f in enumerate(flats):
            p = self.array[i]
            pts = self._c(f["points"], np.float64)
            mean = self._c(f["mean"], np.float64)
            icov = self._c(f["inv_cov"], np.float64)
            m = p.map
            m.w_cells, m.h_cells = int(f["w_cells"]), int(f["h_cells"])
            m.width_m, m.height_m, m.cell_side = float(f["width_m"]), float(f["height_m"]), float(f["cell_side"])
            m.x_min, m.x_max, m.y_min, m.y_max = float(f["x_min"]), float(f["x_max"]), float(f["y_min"]), float(f["y_max"])
            m.mean, m.inv_cov = _ptr(mean), _ptr(icov)
            if f.get("cell_index") is not None:
                ci = self._c(f["cell_index"], np.int32)
                m.n_sparse, m.cell_index, m.built = int(ci.shape[0]), _ptr(ci), C.c_void_p(None)
            else:
                built = self._c(f["built"], np.uint8)
                m.n_sparse, m.cell_index, m.built = -1, C.c_void_p(None), _ptr(built)
            m.reserved = 0
            p.points_xy, p.n_points = _ptr(pts), int(pts.shape[0])
            p.seed = int(f.get("seed", 1)) & 0xFFFFFFFF
            guess, dev = f.get("guess", (0., 0., 0.)), f.get("deviation", (0., 0., 0.))
            for k in range(3):
                p.guess[k] = float(guess[k])
                p.deviation[k] = float(dev[k])
            rs = f.get("rand_stream")
            if rs is not None:
                rs = self._c(rs, np.int32)
                p.rand_stream, p.rand_count = _ptr(rs), int(rs.shape[0])
            else:
                p.rand_stream, p.rand_count = C.c_void_p(None), 0

    def _c(self, a, dtype):
        # identical input arrays map to identical pointers so the library can share tables
        if isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"]:
            arr = a
        else:
            arr = np.ascontiguousarray(a, dtype=dtype)
        self.keep.append(arr)
        return arr


class Exchange:
    """Gathered results of all ranks, filled by peer stores from the PSO kernel's epilogue (ndtpso_exchange_*)."""

    def __init__(self, ctx: "Context", world: int, rank: int, n_per_rank: int):
        self.ctx, self.world, self.rank, self.n = ctx, int(world), int(rank), int(n_per_rank)
        self.h = C.c_void_p()
        self.handle = np.zeros(IPC_HANDLE_BYTES, dtype=np.uint8)
        ctx._check(ctx.lib.ndtpso_exchange_create(ctx.h, self.world, self.rank, self.n, C.byref(self.h), _ptr(self.handle)))

    def connect(self, all_handles):
        """all_handles: uint8 [world, 64], every rank's `handle` in rank order."""
        a = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(self.world, IPC_HANDLE_BYTES)
        self.ctx._check(self.ctx.lib.ndtpso_exchange_connect(self.h, _ptr(a)))

    def connect_local(self, peers):
        arr = (C.c_void_p * self.world)(*[p.h for p in peers])
        self.ctx._check(self.ctx.lib.ndtpso_exchange_connect_local(self.h, arr))

    def wait(self):
        self.ctx._check(self.ctx.lib.ndtpso_exchange_wait(self.h))

    def device_results_ptr(self) -> int:
        return int(self.ctx.lib.ndtpso_exchange_device_results(self.h) or 0)

    def results(self):
        pose = np.empty((self.world * self.n, 3), dtype=np.float64)
        cost = np.empty(self.world * self.n, dtype=np.float64)
        self.ctx._check(self.ctx.lib.ndtpso_exchange_results(self.h, _ptr(pose), _ptr(cost)))
        return pose, cost

    def close(self):
        if self.h:
            self.ctx.lib.ndtpso_exchange_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """A batch resident in HBM (ndtpso_batch_*)."""

    def __init__(self, ctx: "Context", problems: ProblemSet, conf: PsoConfig):
        self.ctx, self.problems, self.n = ctx, problems, problems.n
        h = C.c_void_p()
        ctx._check(ctx.lib.ndtpso_batch_create(ctx.h, problems.n, problems.array, C.byref(conf), C.byref(h)))
        self.h = h

    def solve(self):
        self.ctx._check(self.ctx.lib.ndtpso_batch_solve(self.h))

    def attach_exchange(self, ex):
        self.ctx._check(self.ctx.lib.ndtpso_batch_attach_exchange(self.h, ex.h if ex is not None else None))

    def device_results_ptr(self) -> int:
        return int(self.ctx.lib.ndtpso_batch_device_results(self.h) or 0)

    def results(self):
        pose = np.empty((self.n, 3), dtype=np.float64)
        cost = np.empty(self.n, dtype=np.float64)
        self.ctx._check(self.ctx.lib.ndtpso_batch_results(self.h, _ptr(pose), _ptr(cost)))
        return pose, cost

    def stats(self):
        out = np.zeros((self.n, 2), dtype=np.int32)
        self.ctx._check(self.ctx.lib.ndtpso_batch_stats(self.h, _ptr(out)))
        return out

    def stats_ex(self):
        """[n, 4]: rounds, gbest updates, fp64 cost evaluations, evaluations settled by the fp32 screen alone."""
        out = np.zeros((self.n, 4), dtype=np.int32)
        self.ctx._check(self.ctx.lib.ndtpso_batch_stats_ex(self.h, _ptr(out)))
        return out

    def kernel_times_ms(self):
        """(K0 compaction, K1 rand stream, K2 PSO) device milliseconds of the last solve."""
        out = np.zeros(3, dtype=np.float64)
        self.ctx._check(self.ctx.lib.ndtpso_batch_kernel_times(self.h, _ptr(out)))
        return out

    def close(self):
        if self.h:
            self.ctx.lib.ndtpso_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU (ndtpso_ctx_*)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.ndtpso_ctx_create(device, C.byref(h))
        if rc != OK:
            raise NdtpsoError(rc, "ndtpso_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h

    def _check(self, rc):
        if rc != OK:
            raise NdtpsoError(rc, (self.lib.ndtpso_last_error(self.h) or b"").decode())

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.ndtpso_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_option(self, option: int, value: int):
        self._check(self.lib.ndtpso_ctx_set_option(self.h, option, value))

    def align_batch(self, flats, conf: PsoConfig):
        """pso_optimization for every flat problem: returns (pose[n,3], cost[n])."""
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        pose = np.empty((ps.n, 3), dtype=np.float64)
        cost = np.empty(ps.n, dtype=np.float64)
        self._check(self.lib.ndtpso_align_batch(self.h, ps.n, ps.array, C.byref(conf), _ptr(pose), _ptr(cost)))
        return pose, cost

    def align_submit(self, problems: "ProblemSet", conf: PsoConfig):
        """Asynchronous half of align_batch: stage + H2D + launches.  Returns a ticket for align_collect."""
        h = C.c_void_p()
        self._check(self.lib.ndtpso_align_submit(self.h, problems.n, problems.array, C.byref(conf), C.byref(h)))
        return (h, problems.n)

    def align_collect(self, ticket):
        h, n = ticket
        pose = np.empty((n, 3), dtype=np.float64)
        cost = np.empty(n, dtype=np.float64)
        self._check(self.lib.ndtpso_align_collect(h, _ptr(pose), _ptr(cost)))
        return pose, cost

    def cost_batch(self, flats, poses):
        """cost_function of poses[n, m, 3] -> cost[n, m]."""
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(ps.n, -1, 3)
        out = np.empty(poses.shape[:2], dtype=np.float64)
        self._check(self.lib.ndtpso_cost_batch(self.h, ps.n, ps.array, poses.shape[1], _ptr(poses), _ptr(out)))
        return out

    def screen_bounds(self, flats, poses):
        """The fp32 screen's lower bound of cost_function for poses[n, m, 3]: returns lower[n, m] (<= cost always)."""
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(ps.n, -1, 3)
        out = np.empty((ps.n, poses.shape[1]), dtype=np.float64)
        for k0 in range(0, poses.shape[1], 48):  # the kernel's shared-memory arrays grow with the number of poses per launch
            chunk = np.ascontiguousarray(poses[:, k0:k0 + 48])
            part = np.empty((ps.n, chunk.shape[1]), dtype=np.float64)
            self._check(self.lib.ndtpso_screen_bounds(self.h, ps.n, ps.array, chunk.shape[1], _ptr(chunk), _ptr(part)))
            out[:, k0:k0 + 48] = part
        return out

    def batch(self, flats, conf: PsoConfig) -> Batch:
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        return Batch(self, ps, conf)

    def launch_count(self) -> int:
        return int(self.lib.ndtpso_ctx_launch_count(self.h))

    def last_transfer_bytes(self):
        """(h2d, d2h) bytes of the most recent upload / results read."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ndtpso_ctx_last_transfer_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def synchronize(self):
        self._check(self.lib.ndtpso_ctx_synchronize(self.h))

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0.0)
        self._check(self.lib.ndtpso_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            self.lib.ndtpso_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Multi:
    """Several GPUs behind one call (include/ndtpso_b200.h: ndtpso_multi_*): a single process, one context per device, the
    problems split into contiguous shards.  `devices`: CUDA ordinals (one context each; an ordinal may repeat)."""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.ndtpso_multi_create(devs, len(devices), C.byref(h))
        if rc != OK:
            raise NdtpsoError(rc, "ndtpso_multi_create failed (no CUDA device? the product has no CPU path)")
        self.h = h
        self.size = len(devices)

    def _check(self, rc):
        if rc != OK:
            raise NdtpsoError(rc, (self.lib.ndtpso_multi_last_error(self.h) or b"").decode())

    def set_option(self, option: int, value: int):
        for i in range(self.size):
            rc = self.lib.ndtpso_ctx_set_option(self.lib.ndtpso_multi_ctx(self.h, i), option, int(value))
            if rc != OK:
                raise NdtpsoError(rc, "ndtpso_ctx_set_option failed")

    def launch_count(self) -> int:
        return sum(int(self.lib.ndtpso_ctx_launch_count(self.lib.ndtpso_multi_ctx(self.h, i))) for i in range(self.size))

    def align_batch(self, flats, conf: PsoConfig):
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        pose, cost = np.empty((ps.n, 3)), np.empty(ps.n)
        self._check(self.lib.ndtpso_align_batch_multi(self.h, ps.n, ps.array, C.byref(conf), _ptr(pose), _ptr(cost)))
        return pose, cost

    def align_submit(self, flats, conf: PsoConfig):
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        t = C.c_void_p()
        self._check(self.lib.ndtpso_align_submit_multi(self.h, ps.n, ps.array, C.byref(conf), C.byref(t)))
        return (t, ps)

    def align_collect(self, ticket):
        t, ps = ticket
        pose, cost = np.empty((ps.n, 3)), np.empty(ps.n)
        self._check(self.lib.ndtpso_align_collect_multi(t, _ptr(pose), _ptr(cost)))
        return pose, cost

    def batch(self, flats, conf: PsoConfig) -> "MultiBatch":
        ps = flats if isinstance(flats, ProblemSet) else ProblemSet(flats)
        return MultiBatch(self, ps, conf)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ndtpso_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBatch:
    """Shards resident in HBM, one per device; with equal shards every device also ends up with all results (fused exchange)."""

    def __init__(self, multi: Multi, ps: ProblemSet, conf: PsoConfig):
        self.multi, self.ps = multi, ps
        h = C.c_void_p()
        multi._check(multi.lib.ndtpso_multi_batch_create(multi.h, ps.n, ps.array, C.byref(conf), C.byref(h)))
        self.h = h

    def solve(self):
        self.multi._check(self.multi.lib.ndtpso_multi_batch_solve(self.h))

    def results(self):
        pose, cost = np.empty((self.ps.n, 3)), np.empty(self.ps.n)
        self.multi._check(self.multi.lib.ndtpso_multi_batch_results(self.h, _ptr(pose), _ptr(cost)))
        return pose, cost

    def device_results_ptr(self, i: int):
        return self.multi.lib.ndtpso_multi_batch_device_results(self.h, int(i))

    def close(self):
        if getattr(self, "h", None):
            self.multi.lib.ndtpso_multi_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
