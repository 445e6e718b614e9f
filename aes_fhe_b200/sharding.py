"""Multi-GPU plumbing: the path shards by ciphertext batch -- no
collective on the data path.  torch.distributed (NCCL on B200s over NVLink/NVSwitch, gloo in
CPU tests) is used only to broadcast evaluation keys once and to gather results."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of the ciphertext-batch index owned by `rank` (remainder spread
    over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _as_tensor(backend, handle) -> torch.Tensor:
    if isinstance(handle, torch.Tensor):
        return handle
    return torch.from_numpy(np.ascontiguousarray(handle).view(np.int64))


def broadcast_handle(backend, handle, src: int = 0):
    """Broadcast one backend tensor (key material) from `src`; returns the received handle."""
    t = _as_tensor(backend, handle).contiguous()
    dist.broadcast(t, src=src)
    if isinstance(handle, torch.Tensor):
        return t
    return t.numpy().view(np.uint64)


def broadcast_evaluation_keys(ctx, src: int = 0) -> None:
    """rlk / conjugation / rotation keys of an EngineContext come from rank `src`."""
    be = ctx.engine.backend
    ctx.public_key.polys = broadcast_handle(be, ctx.public_key.polys, src)
    ctx.relinearization_key.data = broadcast_handle(be, ctx.relinearization_key.data, src)
    ctx.conjugation_key.data = broadcast_handle(be, ctx.conjugation_key.data, src)
    for k in sorted(ctx.rotation_key.keys):
        ctx.rotation_key.keys[k].data = broadcast_handle(be, ctx.rotation_key.keys[k].data, src)


def gather_handles(backend, handle, dst: int = 0) -> List:
    """All ranks' result tensors on every rank (equal shapes)."""
    t = _as_tensor(backend, handle).contiguous()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    if isinstance(handle, torch.Tensor):
        return out
    return [o.numpy().view(np.uint64) for o in out]
