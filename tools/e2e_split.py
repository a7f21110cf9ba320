"""Host-side split of one end-to-end align_batch step (B cfg2 problems, pinned inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from ndtpso_slam_b200 import capi, workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(B)
ps = capi.ProblemSet(flats)
conf = capi.PsoConfig.make(population=70, iterations=50)
ctx = capi.Context(0)
for _ in range(3): ctx.align_batch(ps, conf)
def t(f, n=10):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
for ch in (1, 2, 4):
    ctx.set_option(capi.OPT_PIPELINE_CHUNKS, ch)
    for _ in range(3): ctx.align_batch(ps, conf)
    print(f"B={B}: align_batch chunks={ch}        {t(lambda: ctx.align_batch(ps, conf)):.3f} ms")
ctx.set_option(capi.OPT_PIPELINE_CHUNKS, 1)
h = C.c_void_p()
def create():
    ctx._check(ctx.lib.ndtpso_batch_create(ctx.h, ps.n, ps.array, C.byref(conf), C.byref(h))); ctx.synchronize()
def destroy(): ctx.lib.ndtpso_batch_destroy(h)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); create(); ts.append(time.perf_counter() - t0); destroy()
print(f"   batch_create dense + H2D + sync {1e3 * np.median(ts):.3f} ms   (h2d bytes {ctx.last_transfer_bytes()[0]})")
bt = ctx.batch(ps, conf)
print(f"   solve + sync                    {t(lambda: (bt.solve(), ctx.synchronize())):.3f} ms")
print(f"   results (D2H + sync)            {t(lambda: bt.results()):.3f} ms")
# pipelined submit/collect: host time of each call
import time as _t
ctx.set_option(capi.OPT_PIPELINE_CHUNKS, 1)
for _ in range(3):
    ctx.align_collect(ctx.align_submit(ps, conf))
N = 12
ts_sub, ts_col = [], []
t_all0 = _t.perf_counter()
t0 = _t.perf_counter(); ticket = ctx.align_submit(ps, conf); ts_sub.append(_t.perf_counter() - t0)
for _ in range(N - 1):
    t0 = _t.perf_counter(); nxt = ctx.align_submit(ps, conf); ts_sub.append(_t.perf_counter() - t0)
    t0 = _t.perf_counter(); ctx.align_collect(ticket); ts_col.append(_t.perf_counter() - t0)
    ticket = nxt
t0 = _t.perf_counter(); ctx.align_collect(ticket); ts_col.append(_t.perf_counter() - t0)
t_all = _t.perf_counter() - t_all0
print(f"   pipelined: {1e3 * t_all / N:.3f} ms/step; submit host ms: {[round(1e3 * x, 2) for x in ts_sub]}")
print(f"              collect host ms: {[round(1e3 * x, 2) for x in ts_col]}")
