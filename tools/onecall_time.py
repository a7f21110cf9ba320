"""ndtpso_align_batch (one synchronous call, host buffers) vs its chunked pipelining option. usage: python tools/onecall_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, workload
flats = workload.cfg2_batch(256)
conf = capi.PsoConfig.make(population=70, iterations=50)
ps = capi.ProblemSet(flats)
ref = None
for chunks in (1, 2, 3, 4, 1, 2):
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_PIPELINE_CHUNKS, chunks)
    for _ in range(3): pose, cost = ctx.align_batch(ps, conf)
    t0 = time.perf_counter()
    for _ in range(10): pose, cost = ctx.align_batch(ps, conf)
    dt = (time.perf_counter() - t0) / 10 * 1e3
    if ref is None: ref = pose
    print(f"chunks {chunks}: {dt:.3f} ms per call -> {256/dt*1e3:.0f} matches/s  equal {np.array_equal(pose, ref)}")
    ctx.close()
