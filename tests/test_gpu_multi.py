"""Several GPUs behind one call (ndtpso_multi_*, include/ndtpso_b200.h): a single process, one context per device, contiguous
shards.  On a box with one GPU the "devices" are two or three contexts on device 0 (the code path is the same: own host
thread, staging pool, streams and — for the resident form — the fused result exchange between the contexts); with more GPUs the
real devices are used as well."""
import ctypes as C

import numpy as np
import pytest

from ndtpso_slam_b200 import capi
from tests.problems import POSE_ATOL, SCORE_RTOL, rel_err

pytestmark = pytest.mark.gpu


def device_sets():
    n = capi.load_library().ndtpso_device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


def conf_of(c):
    return capi.PsoConfig.make(population=c["P"], iterations=c["I"], w=c["w"], c1=c["c1"], c2=c["c2"], w_dumping=c["w_dumping"])


@pytest.mark.parametrize("devices", device_sets())
def test_align_batch_multi_equals_single_context(golden, ctx, devices):
    c, flats = golden.problems("cfg1")
    cf = conf_of(c)
    want_pose, want_cost = ctx.align_batch(flats, cf)
    assert np.abs(want_pose - c["pose"]).max() <= POSE_ATOL and rel_err(want_cost, c["cost"]).max() <= SCORE_RTOL
    m = capi.Multi(devices)
    try:
        for n in (len(flats), 5, 1, 0):  # even shards, ragged shards, fewer problems than devices, none
            pose, cost = m.align_batch(flats[:n], cf)
            assert np.array_equal(pose, want_pose[:n]) and np.array_equal(cost, want_cost[:n]), (devices, n)
        # the throughput form, three batches in flight
        parts = [flats[i::3] for i in range(3)]
        tickets = [m.align_submit(p, cf) for p in parts]
        for i, t in enumerate(tickets):
            pose, cost = m.align_collect(t)
            assert np.array_equal(pose, want_pose[i::3]) and np.array_equal(cost, want_cost[i::3])
        assert m.launch_count() > 0
    finally:
        m.close()


@pytest.mark.parametrize("devices", device_sets())
def test_resident_shards_exchange_their_results(golden, ctx, devices, traj_batch, batch_golden):
    """Equal shards: after a solve every device holds ALL results (peer stores from the PSO kernel's epilogue), and they are
    the reference's."""
    G = len(devices)
    n = 16 * G
    flats = traj_batch(n)
    cf = capi.PsoConfig.make(population=70, iterations=50)
    m = capi.Multi(devices)
    try:
        bt = m.batch(flats, cf)
        for _ in range(2):  # re-solving alternates the exchange's buffers
            bt.solve()
            pose, cost = bt.results()
            assert np.abs(pose - batch_golden.pose[:n]).max() <= POSE_ATOL and rel_err(cost, batch_golden.cost[:n]).max() <= SCORE_RTOL
        import torch
        for i in range(G):
            ptr = bt.device_results_ptr(i)
            assert ptr, "equal shards must be exchanged on the device side"

            class _Ext:
                __cuda_array_interface__ = {"shape": (n * 4,), "typestr": "<f8", "data": (ptr, False), "version": 3}
            rows = torch.as_tensor(_Ext(), device=f"cuda:{devices[i]}").view(n, 4).cpu().numpy()
            assert np.array_equal(rows[:, :3], pose) and np.array_equal(rows[:, 3], cost), i
        bt.close()
        # ragged shards: no device-side exchange, the host gathers
        bt = m.batch(flats[:n - 1], cf)
        bt.solve()
        pose2, cost2 = bt.results()
        assert bt.device_results_ptr(0) is None
        assert np.array_equal(pose2, pose[:n - 1]) and np.array_equal(cost2, cost[:n - 1])
        bt.close()
    finally:
        m.close()


def test_multi_argument_errors():
    L = capi.load_library()
    h = C.c_void_p()
    assert L.ndtpso_multi_create(None, 0, C.byref(h)) == capi.ERR_ARG
    assert L.ndtpso_multi_create(None, 9, C.byref(h)) == capi.ERR_ARG
    devs = (C.c_int32 * 1)(999)
    assert L.ndtpso_multi_create(devs, 1, C.byref(h)) == capi.ERR_NODEVICE
