"""Where a PSO solve spends its cycles (thread 0 of every CTA; needs tools/_build/libscore_bench.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndtpso_slam_b200 import capi, workload
SO = os.path.join(ROOT, "tools", "_build", os.environ.get("NDTPSO_PHASE_LIB", "libscore_bench.so"))
capi._build.LIB_PATH = SO  # run the whole C ABI from the instrumented build
L = capi.load_library()
L.ndtpso_bench_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
bt.solve(); bt.results()
L.ndtpso_bench_phase_cycles(None, 1)
bt.solve(); bt.results()
out = (C.c_ulonglong * 8)()
L.ndtpso_bench_phase_cycles(out, 0)
v = list(out)[:7]; tot = sum(v)
names = ["prologue+init", "phase A + barrier", "phase B2 (fp64 scoring of the survivors; all of B without the screen)", "barrier after B", "phase C", "phase B1 (fp32 screen) + barrier", "survivor list + barrier"]
print(f"B={batch}: kernel {bt.kernel_times_ms()[2]:.3f} ms; per-CTA mean cycles {tot / batch:.0f}")
for n, x in zip(names, v):
    print(f"  {n:24s} {x / batch:12.0f} cycles/CTA  {x / tot:6.1%}")
