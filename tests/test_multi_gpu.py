"""Sharded == unsharded, bitwise: the single-process ShardedWHENet over every visible GPU and
the torch.distributed entry point (one process per GPU, NCCL).  Skipped with fewer than two devices; the gloo / stub-net
CPU test of the same host logic is in test_dp.py."""
import os
import socket

import numpy as np
import pytest

from conftest import SNAP

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _crops(n):
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    base = np.concatenate([np.load(os.path.join(gold, "sample_crops.npy")), np.load(os.path.join(gold, "jitter_crops.npy"))])
    rng = np.random.default_rng(3)
    idx = rng.integers(0, len(base), n)
    x = base[idx].copy()
    x[:, :4, :4, :] = rng.integers(0, 256, (n, 4, 4, 3), dtype=np.uint8)     # make every crop distinct
    return x


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_sharded_single_process_equals_unsharded(prec):
    if _ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    import whenet_b200
    from whenet_b200 import dp
    n = 4097 if prec == "bf16" else 1025                       # ragged: the last shard is short
    x = _crops(n)
    one = whenet_b200.WHENet(SNAP, device=0, precision=prec, max_batch=512)
    ref = np.stack(one.get_angle(x), axis=1)
    one.close()
    net = dp.ShardedWHENet(SNAP, devices=range(_ndev()), precision=prec, max_batch=512)
    got = np.stack(net.get_angle(x), axis=1)
    net.close()
    assert np.array_equal(got, ref)


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    import whenet_b200
    from whenet_b200 import dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    net = whenet_b200.WHENet(SNAP, device=rank, precision="bf16", max_batch=512)
    got = np.stack(dp.get_angle_distributed(net, _crops(n)), axis=1)
    q.put((rank, got))
    dist.barrier()
    net.close()
    dist.destroy_process_group()


def test_distributed_nccl_equals_unsharded():
    if _ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import whenet_b200
    n, world = 1031, min(_ndev(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one = whenet_b200.WHENet(SNAP, device=0, precision="bf16", max_batch=512)
    ref = np.stack(one.get_angle(_crops(n)), axis=1)
    one.close()
    for _rank, got in res:
        assert np.array_equal(got, ref)
