"""The bench's resident rate without the rest of bench.py: four resident copies of a cfg2 batch solved round-robin on two streams.
usage: python tools/value_rate.py [lib-variant-name|prod] [batch] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ndtpso_slam_b200 import capi, workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "prod"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if name != "prod":
    capi._build.LIB_PATH = os.path.join(ROOT, "tools", "_build", f"lib_{name}.so")
ctx = capi.Context(0)
conf = capi.PsoConfig.make(population=70, iterations=50)
pset = capi.ProblemSet(workload.cfg2_batch(batch))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
bts = [ctx.batch(pset, conf) for _ in range(4)]


def step(k):
    ctx.set_stream(streams[k & 1].cuda_stream)
    bts[k % 4].solve()


for k in range(8):
    step(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(streams[0])
for k in range(steps):
    step(k)
streams[0].wait_stream(streams[1])
e1.record(streams[0])
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
# one stream, one copy
ctx.set_stream(streams[0].cuda_stream)
for _ in range(3):
    bts[0].solve()
torch.cuda.synchronize()
e0.record(streams[0])
for _ in range(steps // 2):
    bts[0].solve()
e1.record(streams[0])
torch.cuda.synchronize()
ms1 = e0.elapsed_time(e1)
kt = bts[0].kernel_times_ms()
print(f"{name:10s} B={batch}: two streams {batch * steps / ms:9.1f} k matches/s ({ms / steps:.3f} ms/step); one stream "
      f"{batch * (steps // 2) / ms1:9.1f} k matches/s ({ms1 / (steps // 2):.3f} ms/step); K0 {kt[0]:.3f} K1 {kt[1]:.3f} K2 {kt[2]:.3f} ms", flush=True)
