"""Multi-GPU plumbing for the evaluate path: the image batch shards by contiguous ranges (every frame is independent:
detection, crop, heat-maps and pose depend on nothing else), every rank holds all weights, and the only exchange is
ONE all-gather of fixed-size result records (bp_record, include/betapose_b200.h) so rank 0 can write the JSON in
input order.  NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors for the host-logic tests.
The reference has no multi-GPU inference (README.md:83 pins one device); SURVEY.md 8(e).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`: the first (n % world) ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> list[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_records(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local: uint8 [n_local, RECORD_BYTES] on this rank (cuda for nccl, cpu for gloo) -> uint8 [n_total, RECORD_BYTES]
    in global image order on every rank.  One collective; ragged shards are padded to the largest shard."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    mx = max(sizes)
    rec = local.shape[1]
    send = local
    if local.shape[0] != mx:
        send = torch.zeros((mx, rec), dtype=torch.uint8, device=local.device)
        send[: local.shape[0]] = local
    out = torch.empty((world * mx, rec), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    keep = torch.cat([torch.arange(r * mx, r * mx + sizes[r]) for r in range(world)]).to(out.device)
    return out[keep]


def records_to_bytes(rec: np.ndarray) -> torch.Tensor:
    """structured record array (possibly EMPTY: a rank whose shard holds no frame when n_total < world) -> uint8
    [n, RECORD_BYTES] host tensor.  The explicit width matters: reshape(0, -1) is an error."""
    from . import _lib

    assert rec.dtype.itemsize == _lib.RECORD_BYTES
    return torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(len(rec), _lib.RECORD_BYTES).copy())


def records_from_bytes(buf: torch.Tensor) -> np.ndarray:
    from . import stages

    return stages.records_to_numpy(buf)
