"""A/B of library variants (tools/_build/lib_<name>.so, built by tools/variant_time.py): K2 time on resident cfg2 batches, results
compared bit for bit with the first variant, screen statistics, and the screen's bound against the fp64 cost on golden maps.
usage: python tools/ab_check.py name [name ...] [-- batch]      (each variant in its own process)"""
import os
import pickle
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tools", "_build")

if sys.argv[1] != "one":
    args = sys.argv[1:]
    batch = 256
    if "--" in args:
        i = args.index("--")
        batch = int(args[i + 1])
        args = args[:i]
    ref = None
    for name in args:
        out = os.path.join(BUILD, f"ab_{name}.pkl")
        subprocess.run([sys.executable, __file__, "one", name, str(batch), out], check=True)
        import numpy as np
        r = pickle.load(open(out, "rb"))
        if ref is None:
            ref = r
        else:
            same_pose = np.array_equal(ref["pose"], r["pose"])
            same_cost = np.array_equal(ref["cost"], r["cost"])
            print(f"   {name} vs {args[0]}: poses identical {same_pose}, costs identical {same_cost}, max |dpose| {np.abs(ref['pose'] - r['pose']).max():.3e}")
    sys.exit(0)

sys.path.insert(0, ROOT)
import numpy as np
from ndtpso_slam_b200 import capi, workload
from tests.problems import Golden

name, batch, out = sys.argv[2], int(sys.argv[3]), sys.argv[4]
if name != "prod":
    capi._build.LIB_PATH = os.path.join(BUILD, f"lib_{name}.so")
ctx = capi.Context(0)
ctx.set_option(capi.OPT_CLUSTER, 1)
if os.environ.get("NDTPSO_AB_NPT"):
    ctx.set_option(capi.OPT_POINTS_PER_THREAD, int(os.environ["NDTPSO_AB_NPT"]))
bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
ts = []
for _ in range(6):
    bt.solve()
    ts.append(bt.kernel_times_ms()[2])
pose, cost = bt.results()
st = bt.stats_ex() if hasattr(bt, "stats_ex") else None
msg = f"{name:10s} B={batch}: K2 {min(ts):.3f} ms (median {sorted(ts)[len(ts) // 2]:.3f})"
if st is not None:
    st = np.asarray(st)
    msg += f"  per match: rounds {st[:, 0].mean():.1f} fp64 evals {st[:, 2].mean():.1f} screened {st[:, 3].mean():.1f}"
print(msg, flush=True)
# the bound itself
g = Golden()
rng = np.random.default_rng(3)
viol = 0
tight = []
for case in ("cfg2", "cfg5_0.25", "cfg5_2.0", "cfg1", "np2"):
    try:
        flat, c = g.flat(case), g.case(case)
    except KeyError:
        continue
    best = c["pose"][0]
    for sig in ((0.002, 0.002, 0.0005), (0.1, 0.1, 0.01), (1.0, 1.0, 0.3)):
        for m in (70, 513, 7):
            poses = best + rng.normal(size=(m, 3)) * np.array(sig)
            try:
                lower = ctx.screen_bounds([flat], poses[None])[0]
            except capi.NdtpsoError as e:
                print("   ", case, "no screen:", e)
                break
            cst = ctx.cost_batch([flat], poses[None])[0]
            viol += int((lower > cst).sum())
            tight.append(np.median((cst - lower) / np.maximum(np.abs(cst), 1e-9)))
print(f"   bound violations {viol}; median looseness per set: min {min(tight):.5f} max {max(tight):.5f}", flush=True)
pickle.dump({"pose": pose, "cost": cost}, open(out, "wb"))
