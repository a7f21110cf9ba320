import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndtpso_slam_b200 import capi, workload
SO = os.path.join(ROOT, "tools", "_build", "libscore_bench.so")
cfg, var = int(sys.argv[1]), int(sys.argv[2])
L = C.CDLL(SO)
L.ndtpso_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
L.ndtpso_bench_score.argtypes = [C.c_void_p, C.c_int32, C.POINTER(capi.Problem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
h = C.c_void_p(); assert L.ndtpso_ctx_create(0, C.byref(h)) == 0
ps = capi.ProblemSet(workload.cfg2_batch(8))
ms = C.c_double(); info = (C.c_int * 4)()
print(L.ndtpso_bench_score(h, ps.n, ps.array, cfg, var, 148 * 4, 72, 10, C.byref(ms), info), ms.value, list(info))
