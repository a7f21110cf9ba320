"""Phase-B microbenchmark driver. Build: nvcc tools/score_bench.cu -> tools/_build/libscore_bench.so (done here if missing)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "tools", "_build", "libscore_bench.so")
def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
                    "-o", SO, os.path.join(ROOT, "tools", "score_bench.cu")], check=True)
if __name__ == "__main__":
    if "--build" in sys.argv or not os.path.exists(SO):
        build()
        if "--build" in sys.argv: sys.exit(0)
    from ndtpso_slam_b200 import capi, workload
    L = C.CDLL(SO)
    L.ndtpso_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ndtpso_bench_score.argtypes = [C.c_void_p, C.c_int32, C.POINTER(capi.Problem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.ndtpso_last_error.argtypes = [C.c_void_p]; L.ndtpso_last_error.restype = C.c_char_p
    h = C.c_void_p(); assert L.ndtpso_ctx_create(0, C.byref(h)) == 0
    ps = capi.ProblemSet(workload.cfg2_batch(8))
    names = ["(2,2,576,2)", "(2,1,576,2)", "(4,2,320,2)", "(3,4,384,2)", "(4,2,288,3)", "(3,2,384,3)", "(2,4,576,2)", "(6,1,192,3)", "(3,2,384,2)"]
    ncand, reps = 72, 40
    for cfg in range(9):
        for var in (0,):
            ms = C.c_double(); info = (C.c_int * 4)()
            for grid_mult in (0,):
                rc = L.ndtpso_bench_score(h, ps.n, ps.array, cfg, var, 148 * 8, ncand, reps, C.byref(ms), info)
                if rc: print("error", rc, L.ndtpso_last_error(h)); continue
                npt, nw, regs, occ = list(info)
                grid = 148 * 8
                pe = grid * reps * ncand * 1081
                print(f"cfg {names[cfg]} var {var}: warps {nw} regs {regs} ctas/SM {occ}: {ms.value:8.3f} ms  {pe/ms.value/1e6:8.1f} G pe/s  -> {pe/ms.value*1e3/(3571*1.127*1081)/1e3:7.1f} k matches/s equiv")
