"""Per-match counters of the PSO kernel on resident cfg2 batches: rounds, gbest updates, fp64 evaluations, evaluations the fp32
screen settled.  usage: python tools/stats_run.py [lib-variant-name|prod] [batch] [points_per_thread]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndtpso_slam_b200 import capi, workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "prod"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if name != "prod":
    capi._build.LIB_PATH = os.path.join(ROOT, "tools", "_build", f"lib_{name}.so")
ctx = capi.Context(0)
ctx.set_option(capi.OPT_CLUSTER, 1)
if len(sys.argv) > 3:
    ctx.set_option(capi.OPT_POINTS_PER_THREAD, int(sys.argv[3]))
bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
ts = []
for _ in range(4):
    bt.solve()
    ts.append(bt.kernel_times_ms()[2])
st = bt.stats_ex().astype(float)
pose, cost = bt.results()
print(f"{name:10s} B={batch}: K2 {min(ts):.3f} ms; per match: rounds {st[:, 0].mean():.1f}, gbest updates {st[:, 1].mean():.1f}, "
      f"fp64 evaluations {st[:, 2].mean():.1f}, settled by the screen {st[:, 3].mean():.1f} "
      f"({st[:, 3].sum() / (st[:, 2].sum() + st[:, 3].sum()):.1%}); pose0 {pose[0]}")
