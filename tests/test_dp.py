"""Data-parallel sharding + the angle all-gather, world_size 2 and 3 over gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whenet_b200 import dp


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            got = [dp.shard_range(n, r, world) for r in range(world)]
            flat = [i for b, e in got for i in range(b, e)]
            assert flat == list(range(n))
            assert max(e - b for b, e in got) <= -(-n // world) if n else True
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = dp.shard_range(n_total, rank, world)
    full = torch.arange(n_total * 3, dtype=torch.float32).reshape(n_total, 3)
    out = dp.gather_angles(full[b:e].clone(), n_total)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_total", [(2, 10), (2, 7), (3, 4), (2, 1)])
def test_gather_angles_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(n_total * 3, dtype=np.float32).reshape(n_total, 3)
    for _rank, out in res:
        assert np.array_equal(out, want)


class _StubNet:
    """Deterministic stand-in for WHENet on CPU: angles are a fixed function of each crop's bytes (batch invariant)."""
    max_batch = 4

    def get_angle(self, img):
        img = np.asarray(img)
        v = img.reshape(img.shape[0], -1).astype(np.float64)
        return (v.sum(1) % 360 - 180).astype(np.float32), (v[:, 0] - 99).astype(np.float32), (v[:, 1] * 0.5 - 99).astype(np.float32)


def _worker_dist(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(11).integers(0, 256, (n_total, 224, 224, 3), dtype=np.uint8)
    y, p, r = dp.get_angle_distributed(_StubNet(), x)
    q.put((rank, np.stack([y, p, r], axis=1)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 9), (3, 2), (2, 1)])
def test_get_angle_distributed_gloo(world, n_total):
    """The product-level sharded entry point (torch.distributed, one process per shard) on CPU with a stub net: ragged and
    empty shards, same result on every rank as the unsharded call."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dist, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(11).integers(0, 256, (n_total, 224, 224, 3), dtype=np.uint8)
    want = np.stack(_StubNet().get_angle(x), axis=1)
    for _rank, out in res:
        assert np.array_equal(out, want)
