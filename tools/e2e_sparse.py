"""e2e (align_submit/collect, 2 in flight) with dense vs sparse host tables. usage: python tools/e2e_sparse.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, workload
conf = capi.PsoConfig.make(population=70, iterations=50)
for sparse in (False, True, False, True):
    ps = capi.ProblemSet(workload.cfg2_batch(256, sparse=sparse))
    ctx = capi.Context(0)
    for _ in range(3): ctx.align_batch(ps, conf)
    t = ctx.align_submit(ps, conf); n = ctx.align_submit(ps, conf); ctx.align_collect(t); ctx.align_collect(n)
    K = 30
    t0 = time.perf_counter()
    ticket = ctx.align_submit(ps, conf)
    for _ in range(K - 1):
        nxt = ctx.align_submit(ps, conf); pose, cost = ctx.align_collect(ticket); ticket = nxt
    pose, cost = ctx.align_collect(ticket)
    dt = (time.perf_counter() - t0) / K * 1e3
    t0 = time.perf_counter()
    for _ in range(10): ctx.align_batch(ps, conf)
    d1 = (time.perf_counter() - t0) / 10 * 1e3
    print(f"sparse={sparse}: pipelined {dt:.3f} ms/batch -> {256/dt*1e3:.0f}/s; one call {d1:.3f} ms -> {256/d1*1e3:.0f}/s; h2d {ctx.last_transfer_bytes()}; checksum {pose.sum():.12f}")
    ctx.close()
