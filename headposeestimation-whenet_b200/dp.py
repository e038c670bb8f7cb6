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


def get_angle_distributed(net, img, group=None, device_gather: bool = True):
    """``WHENet.get_angle`` of reference whenet.py:22-34 for a crop batch sharded over the ranks of a process group
    (one process per GPU, launched with torchrun).  Every rank passes the SAME ``img`` (N,224,224,3); rank r runs the
    forward for its ``shard_range`` block on its own GPU and one all-gather of the (n_local, 3) angles gives every rank
    the full result.  Returns ``(yaw, pitch, roll)``, float32 ``(N,)``, identical on all ranks and bitwise identical to
    the unsharded call (the kernels are batch invariant).

    ``net`` needs ``get_angle``; with ``device_gather`` and a net that has ``forward_device_from_host`` semantics (our
    ``WHENet``) the shard result stays on the device and NCCL gathers it from there."""
    import numpy as np
    import torch
    import torch.distributed as dist
    img = np.asarray(img)
    n = int(img.shape[0])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = shard_range(n, rank, world)
    backend = dist.get_backend(group)
    on_device = device_gather and backend == "nccl" and hasattr(net, "forward_host_to_device")
    if on_device:
        dev = torch.device("cuda", net.device)
        local = torch.zeros((max(e - b, 0), 3), dtype=torch.float32, device=dev)
        if e > b:
            shard = np.ascontiguousarray(img[b:e])
            if shard.dtype != np.uint8:
                y, p, r = net.get_angle(shard)
                local.copy_(torch.from_numpy(np.stack([y, p, r], axis=1)))
            else:
                net.set_stream(torch.cuda.current_stream(dev).cuda_stream)
                for off in range(0, e - b, net.max_batch):
                    nb = min(net.max_batch, e - b - off)
                    net.forward_host_to_device(shard[off:off + nb], local[off:off + nb])
                net.synchronize()      # the staging buffers may be reused; also surfaces kernel timeouts
        full = gather_angles(local, n, group)
        ang = full.cpu().numpy()
    else:
        if e > b:
            y, p, r = net.get_angle(img[b:e])
            local = torch.from_numpy(np.stack([y, p, r], axis=1).astype(np.float32))
        else:
            local = torch.zeros((0, 3), dtype=torch.float32)
        if backend == "nccl":
            local = local.cuda()
        ang = gather_angles(local, n, group).cpu().numpy()
    return ang[:, 0].copy(), ang[:, 1].copy(), ang[:, 2].copy()


class ShardedWHENet:
    """Single-process form of the same partition: one ``WHENet`` context per GPU of the box, contiguous ``shard_range``
    blocks, one host thread per device (the C calls release the GIL and block only their own thread), no collective at
    all - the angles land in one host array.  Drop-in for ``WHENet`` where one process owns all GPUs:

        net = ShardedWHENet(snapshot, devices=range(8), precision="bf16")
        yaw, pitch, roll = net.get_angle(crops)          # crops: (N,224,224,3), any N
    """

    def __init__(self, snapshot=None, devices=None, precision=None, max_batch: int = 512):
        from .whenet import WHENet
        if devices is None:
            import torch
            devices = range(torch.cuda.device_count())
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("no devices")
        self.nets = [WHENet(snapshot, device=d, precision=precision, max_batch=max_batch) for d in self.devices]
        self.model = self.nets[0].model
        self.idx_tensor, self.idx_tensor_yaw = self.nets[0].idx_tensor, self.nets[0].idx_tensor_yaw
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=len(self.nets))

    def get_angle(self, img):
        import numpy as np
        img = np.asarray(img)
        self.nets[0]._check_shape(img)
        n, world = int(img.shape[0]), len(self.nets)
        out = [np.zeros((n,), np.float32) for _ in range(3)]

        def work(rank):
            b, e = shard_range(n, rank, world)
            if e > b:
                y, p, r = self.nets[rank].get_angle(img[b:e])
                out[0][b:e], out[1][b:e], out[2][b:e] = y, p, r

        for f in [self._pool.submit(work, r) for r in range(world)]:
            f.result()
        return out[0], out[1], out[2]

    def close(self):
        self._pool.shutdown(wait=True)
        for m in self.nets:
            m.close()
