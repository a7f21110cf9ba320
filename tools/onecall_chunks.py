"""One-call synchronous ndtpso_align_batch: wall ms per call against the number of pipeline chunks, for several batch sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndtpso_slam_b200 import capi, workload
conf = capi.PsoConfig.make(population=70, iterations=50)
ctx = capi.Context(0)
for B in (64, 128, 192, 256, 296, 512):
    ps = capi.ProblemSet(workload.cfg2_batch(B))
    row = []
    for ch in (1, 2, 3, 4, 6, 8):
        ctx.set_option(capi.OPT_PIPELINE_CHUNKS, ch)
        for _ in range(3):
            ctx.align_batch(ps, conf)
        ts = []
        for _ in range(8):
            t0 = time.perf_counter(); ctx.align_batch(ps, conf); ts.append(time.perf_counter() - t0)
        row.append(1e3 * float(np.median(ts)))
    print(f"B={B:4d}: " + "  ".join(f"chunks {c}: {t:.3f} ms" for c, t in zip((1, 2, 3, 4, 6, 8), row)), flush=True)
