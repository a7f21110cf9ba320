"""Share of the evaluations the fp32 screen settles on the soak test's random tables (tests/test_gpu_screen_soak.py), per geometry.
usage: python tools/soak_settle.py [variant|prod]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndtpso_slam_b200 import capi  # noqa: E402
from tests.test_gpu_screen_soak import GEOMS, random_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "prod"
if name != "prod":
    capi._build.LIB_PATH = os.path.join(ROOT, "tools", "_build", f"lib_{name}.so")
rng = np.random.default_rng(77)
conf = capi.PsoConfig.make(population=24, iterations=10)
tot_s = tot_e = 0
for (width, cs) in GEOMS:
    flats = []
    f, c = random_problem(rng, width, cs, int(rng.integers(300, 1200)))
    for s in range(144):
        g = dict(f)
        wide = s % 3 == 0
        g.update(guess=(c[0] + rng.normal() * 0.05, c[1] + rng.normal() * 0.05, c[2] + rng.normal() * 0.01),
                 deviation=(2.0, 2.0, 0.8) if wide else (0.1, 0.1, 0.01), seed=int(rng.integers(1, 2 ** 31)))
        flats.append(g)
    cx = capi.Context(0)
    cx.set_option(capi.OPT_CLUSTER, 1)
    bt = cx.batch(flats, conf)
    bt.solve()
    st = bt.stats_ex()
    s_, e_ = int(st[:, 3].sum()), int(st[:, 2].sum() + st[:, 3].sum())
    tot_s += s_
    tot_e += e_
    print(f"{name:6s} frame {width:5.0f} m cell {cs:4.2f} m: settled {s_ / max(e_, 1):.3f} of {e_} evaluations", flush=True)
    bt.close()
    cx.close()
print(f"{name:6s} all: {tot_s / tot_e:.3f}")
