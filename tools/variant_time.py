"""Builds variants of the library with extra -D flags (tools/_build/lib_<name>.so) and times K2 on resident cfg2 batches.
usage: python tools/variant_time.py build name:-DFLAG=1,-DOTHER=2 ...     (no GPU needed)
       python tools/variant_time.py run name ... [-- batch ...]           (each variant in its own process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tools", "_build")


def so(name):
    return os.path.join(BUILD, f"lib_{name}.so")


if sys.argv[1] == "build":
    os.makedirs(BUILD, exist_ok=True)
    procs = []
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition(":")
        cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]
        cmd += [f for f in flags.split(",") if f] + ["-o", so(name), os.path.join(ROOT, "ndtpso_slam_b200", "csrc", "ndtpso_capi.cu")]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        print(name, "rc", p.wait())
elif sys.argv[1] == "run":
    args = sys.argv[2:]
    batches = [256]
    if "--" in args:
        i = args.index("--")
        batches = [int(a) for a in args[i + 1:]]
        args = args[:i]
    for name in args:
        for b in batches:
            subprocess.run([sys.executable, __file__, "one", name, str(b)])
else:
    sys.path.insert(0, ROOT)
    from ndtpso_slam_b200 import capi, workload
    name, batch = sys.argv[2], int(sys.argv[3])
    if name != "prod":
        capi._build.LIB_PATH = so(name)
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_CLUSTER, 1)
    bt = ctx.batch(workload.cfg2_batch(batch), capi.PsoConfig.make(population=70, iterations=50))
    ts = []
    for _ in range(5):
        bt.solve()
        ts.append(bt.kernel_times_ms()[2])
    pose, cost = bt.results()
    print(f"{name:12s} B={batch:4d}: K2 {min(ts):.3f} ms  pose0 {pose[0]} cost0 {cost[0]!r}", flush=True)
