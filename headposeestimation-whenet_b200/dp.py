"""Data-parallel sharding of a crop batch over the GPUs of one box (one process per GPU).

Every crop is independent (no cross-crop op anywhere in reference whenet.py:22-34), so the
partition is embarrassingly parallel: rank r owns the contiguous block of ceil(N/G) crops
starting at r*ceil(N/G) (the last ranks may be short or empty).  The only collective is ONE
all-gather of the (n_local, 3) float32 angles; shards are padded to the
uniform count for the collective and the padding is dropped afterwards.  Works with the
``nccl`` backend (device tensors, NVLink/NVSwitch) and with ``gloo`` (CPU tensors, tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the crops rank ``rank`` owns out of ``n`` (contiguous, ceil-div blocks)."""
    if n < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n, rank, world))
    per = -(-n // world) if n else 0
    b = min(n, rank * per)
    e = min(n, b + per)
    return b, e


def gather_angles(local, n_total: int, group=None):
    """All-gather per-rank ``(n_local, 3)`` float32 angles into ``(n_total, 3)`` on every rank.

    ``local`` is a torch tensor (CUDA for nccl, CPU for gloo) holding this rank's shard in
    shard_range order.  One collective; ragged tails are padded to ceil(n_total/world)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = -(-n_total // world) if n_total else 0
    if per == 0:
        return local.new_zeros((0, 3))
    if local.shape[0] != per:
        pad = local.new_zeros((per, 3))
        pad[: local.shape[0]] = local
        local = pad
    out = local.new_empty((world * per, 3))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]
