"""Partition of a batch of independent scan-match problems over ranks (one process per GPU) and
the single exchange step of the path: the all-gather of the solved poses.

Problems share nothing, so the data path has no collective; `gather_results` is the one exchange the
north star names (every rank ends up with all [B][4] = (x, y, theta, cost) rows).  Works with any
torch.distributed backend: NCCL on GPUs (bench.py), gloo on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_problems: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: rank r gets [r*ceil(B/G), min(B, (r+1)*ceil(B/G)))  (SURVEY.md section 8e)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    per = -(-n_problems // world_size) if n_problems > 0 else 0
    lo = min(n_problems, rank * per)
    return lo, min(n_problems, lo + per)


def shard(problems, world_size: int, rank: int):
    lo, hi = shard_bounds(len(problems), world_size, rank)
    return problems[lo:hi]


def gather_results(local_rows, n_problems: int, world_size: int, rank: int, group=None):
    """All-gather of per-rank result rows [n_local, 4] into [n_problems, 4] on every rank.

    `local_rows` is a torch tensor (CUDA for NCCL, CPU for gloo).  Shards are padded to the common
    ceil(B/G) rows so that one all_gather_into_tensor call moves everything.
    """
    import torch
    import torch.distributed as dist

    per = -(-n_problems // world_size) if n_problems > 0 else 0
    lo, hi = shard_bounds(n_problems, world_size, rank)
    if local_rows.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {hi - lo}")
    send = local_rows
    if hi - lo < per:
        send = torch.zeros((per, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
        send[: hi - lo] = local_rows
    recv = torch.empty((world_size * per, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    if world_size == 1:
        recv.copy_(send)
    else:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    return recv[:n_problems] if world_size * per == n_problems else torch.cat(
        [recv[r * per: r * per + (shard_bounds(n_problems, world_size, r)[1] - shard_bounds(n_problems, world_size, r)[0])]
         for r in range(world_size)])


def solve_sharded(solve_fn, problems, world_size: int, rank: int, device=None, group=None):
    """solve_fn(list_of_problems) -> (pose[n,3], cost[n]) on this rank's shard; returns the gathered
    (pose[B,3], cost[B]) as numpy arrays, identical on every rank."""
    import torch

    mine = shard(problems, world_size, rank)
    pose, cost = solve_fn(mine) if len(mine) else (np.zeros((0, 3)), np.zeros(0))
    rows = torch.from_numpy(np.concatenate([np.asarray(pose).reshape(-1, 3), np.asarray(cost).reshape(-1, 1)], axis=1))
    if device is not None:
        rows = rows.to(device)
    allrows = gather_results(rows, len(problems), world_size, rank, group=group).cpu().numpy()
    return allrows[:, :3], allrows[:, 3]


def make_exchange(ctx, n_per_rank: int, world_size: int, rank: int, group=None, device=None):
    """The fused form of the exchange (include/ndtpso_b200.h, ndtpso_exchange_*): every rank allocates its gathered
    result buffer, the 64-byte CUDA IPC handles are all-gathered once, and from then on the PSO kernel's epilogue stores
    each result into every rank's buffer over NVLink — no collective per step.

    Returns a connected capi.Exchange, or None ON EVERY RANK if any rank could not create or map the buffers (CUDA IPC
    unavailable): the local steps that can fail are each followed by an all-reduce of a success flag, so the ranks never
    diverge in the collectives they call."""
    import torch
    import torch.distributed as dist

    from . import capi

    def all_ok(flag: bool) -> bool:
        if world_size == 1:
            return flag
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return bool(int(t.item()))

    ex = None
    try:
        ex = capi.Exchange(ctx, world_size, rank, n_per_rank)
    except Exception:  # noqa: BLE001
        ex = None
    if not all_ok(ex is not None):
        if ex is not None:
            ex.close()
        return None
    if world_size > 1:
        mine = torch.from_numpy(ex.handle.copy())
        if device is not None:
            mine = mine.to(device)
        allh = torch.empty((world_size, capi.IPC_HANDLE_BYTES), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        connected = True
        try:
            ex.connect(allh.cpu().numpy())
        except Exception:  # noqa: BLE001
            connected = False
        if not all_ok(connected):
            ex.close()
            return None
    return ex
