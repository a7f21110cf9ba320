"""Small driver for ncu: B cfg2-shaped problems resident in HBM, a few solves.  No torch.
usage: python tools/prof_run.py [batch] [solves] [warps] [smem]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndtpso_slam_b200 import capi, workload  # noqa: E402

if os.environ.get("NDTPSO_PROF_LIB"):  # a variant built by tools/variant_time.py
    capi._build.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "lib_" + os.environ["NDTPSO_PROF_LIB"] + ".so")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
if len(sys.argv) > 3 and int(sys.argv[3]):
    ctx.set_option(capi.OPT_WARPS_PER_CTA, int(sys.argv[3]))
if len(sys.argv) > 4 and int(sys.argv[4]):
    ctx.set_option(capi.OPT_SMEM_BYTES, int(sys.argv[4]))
bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
for _ in range(solves):
    bt.solve()
    print("kernel ms (K0, K1, K2):", bt.kernel_times_ms(), "matches/s:", batch / (bt.kernel_times_ms().sum() * 1e-3))
pose, cost = bt.results()
print("pose0", pose[0], "cost0", cost[0], "rounds", bt.stats()[:, 0].mean())
