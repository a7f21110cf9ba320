"""A small solve for compute-sanitizer --tool racecheck: two cfg1 problems, 8 particles x 4 iterations, one screened CTA per problem
(and once unscreened), so that every phase of pso_sliced_kernel and its prologue runs in seconds under the tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ndtpso_slam_b200 import capi  # noqa: E402
from tests.problems import Golden  # noqa: E402

g = Golden()
c, flats = g.problems("cfg1")
conf = capi.PsoConfig.make(population=8, iterations=4)
out = []
for scr in (1, 0):
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_CLUSTER, 1)
    ctx.set_option(capi.OPT_SCREEN, scr)
    bt = ctx.batch(flats[:2], conf)
    bt.solve()
    out.append(bt.results())
    print("screen", scr, "stats", bt.stats_ex().tolist())
    bt.close()
    ctx.close()
print("screen on == off:", np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]))
