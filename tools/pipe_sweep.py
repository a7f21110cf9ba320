"""Launch-shape sweep of the PSO kernel in the pipelined regime (two resident batches alternating on two streams).
usage: python tools/pipe_sweep.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ndtpso_slam_b200 import capi, workload
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flats = workload.cfg2_batch(batch)
conf = capi.PsoConfig.make(population=70, iterations=50)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
cfgs = [(3, 4, 12 | (1 << 16)), (3, 2, 12 | (1 << 16)), (4, 2, 12 | (1 << 16)), (2, 4, 12 | (1 << 16)), (2, 2, 12 | (1 << 16)), (5, 2, 12 | (1 << 16)),
        (3, 4, 8 | (1 << 16)), (3, 4, 16 | (1 << 16)), (3, 4, 12 | (2 << 16)), (3, 4, 0)]
for npt, jb, win in cfgs:
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_POINTS_PER_THREAD, npt); ctx.set_option(capi.OPT_CANDIDATE_BATCH, jb); ctx.set_option(capi.OPT_HOT_CHUNK, win)
    try:
        bts = [ctx.batch(flats, conf), ctx.batch(flats, conf)]
        def run(k):
            for i in range(k):
                ctx.set_stream(streams[i & 1].cuda_stream); bts[i & 1].solve()
            torch.cuda.synchronize()
        run(4)
        t0 = time.perf_counter(); run(16); dt = (time.perf_counter() - t0) / 16 * 1e3
        ctx.set_stream(streams[0].cuda_stream)
        iso = bts[0].kernel_times_ms()
        print(f"npt {npt} jb {jb} window {win & 0xffff}/{win >> 16}: {dt:.3f} ms/step pipelined -> {batch/dt*1e3:.0f} matches/s")
        for b in bts: b.close()
    except capi.NdtpsoError as e:
        print(f"npt {npt} jb {jb}: {e}")
    ctx.close()
