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
