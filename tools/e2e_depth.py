"""End-to-end throughput of ndtpso_align_submit/collect against the number of batches kept in flight.
usage: python tools/e2e_depth.py [batch] [steps]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ps = capi.ProblemSet(workload.cfg2_batch(B))
conf = capi.PsoConfig.make(population=70, iterations=50)
ctx = capi.Context(0)
for depth in (1, 2, 3, 4, 6):
    def run(n):
        q = collections.deque()
        out = None
        for _ in range(n):
            q.append(ctx.align_submit(ps, conf))
            if len(q) >= depth:
                out = ctx.align_collect(q.popleft())
        while q:
            out = ctx.align_collect(q.popleft())
        return out
    run(2 * depth + 2)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        run(steps)
        best = min(best, time.perf_counter() - t0)
    print(f"depth {depth}: {1e3 * best / steps:.3f} ms/step  {B * steps / best / 1e3:.1f} k matches/s", flush=True)
