"""How tight the fp32 screen's lower bound is: (cost - bound)/|cost| for swarms of poses around the optimum of the golden cfg2 case.
usage: python tools/screen_tightness.py [lib-variant-name|prod]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndtpso_slam_b200 import capi  # noqa: E402
from tests.problems import Golden  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "prod"
if name != "prod":
    capi._build.LIB_PATH = os.path.join(ROOT, "tools", "_build", f"lib_{name}.so")
g = Golden()
ctx = capi.Context(0)
rng = np.random.default_rng(3)
for case in ("cfg2", "cfg5_0.25", "cfg5_2.0", "cfg1"):
    flat, c = g.flat(case), g.case(case)
    best = c["pose"][0]
    for label, sig in (("converged", (0.002, 0.002, 0.0005)), ("spread", (0.1, 0.1, 0.01))):
        poses = best + rng.normal(size=(512, 3)) * np.array(sig)
        lower = ctx.screen_bounds([flat], poses[None])[0]
        cost = ctx.cost_batch([flat], poses[None])[0]
        rel = (cost - lower) / np.abs(cost)
        print(f"{name:8s} {case:10s} {label:10s} (cost - bound)/|cost|: median {np.median(rel):.5f}  p10 {np.percentile(rel, 10):.5f}  p90 {np.percentile(rel, 90):.5f}  "
              f"violations {(lower > cost).sum()}  median cost {np.median(cost):.2f}")
