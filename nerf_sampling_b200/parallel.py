"""Ray sharding across ranks (one process per GPU).  Rays are independent (the reference already loops over independent
chunks, nerf_utils.py:58-85), so the render path needs no data-path collective: each rank renders a contiguous slice and
only the finished image tiles are gathered."""

from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (the first n_items % world ranks get one extra)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tiles(tile: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank image tiles [n_local, C] into [n_total, C] on every rank (ragged tails padded, then cut)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    lo, hi = shard_bounds(n_total, world, rank)
    assert tile.shape[0] == hi - lo
    padded = tile.new_zeros((per,) + tuple(tile.shape[1:]))
    padded[: hi - lo] = tile
    out = tile.new_empty((world * per,) + tuple(tile.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world):
        a, b = shard_bounds(n_total, world, r)
        pieces.append(out[r * per : r * per + (b - a)])
    return torch.cat(pieces, 0)


def render_sharded(render_slice: Callable[[int, int], torch.Tensor], n_rays: int, group=None) -> torch.Tensor:
    """Each rank renders rays [lo, hi) with `render_slice(lo, hi) -> [hi-lo, C]`; every rank gets the full [n_rays, C]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_rays, world, rank)
    tile = render_slice(lo, hi)
    return tile if world == 1 else gather_tiles(tile, n_rays, group)


def allreduce_gradients(params, group=None) -> float:
    """Data-parallel DepthNet training (BASELINE config #5): SUM all-reduce of every parameter gradient as ONE flat
    buffer (3,340,545 fp32 = 13.4 MB for the 10x256 DepthNet) and return the scale 1/world that the optimizer applies
    (``training.Adam.step(grad_scale=...)``), so that equal ray shards reproduce the single-process mean-loss gradient.
    The frozen NeRFs need no communication."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 1.0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 1.0 / world
    # training.DepthNetTrainFn lays the gradients out back to back in one buffer: reduce it in place
    sp = grads[0].untyped_storage().data_ptr()
    back_to_back = all(g.is_contiguous() and g.dtype == grads[0].dtype and g.untyped_storage().data_ptr() == sp for g in grads) and all(
        b.data_ptr() == a.data_ptr() + a.numel() * a.element_size() for a, b in zip(grads, grads[1:]))
    if back_to_back:
        total = sum(g.numel() for g in grads)
        flat = torch.as_strided(grads[0], (total,), (1,), grads[0].storage_offset())
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off : off + g.numel()].view_as(g))
        off += g.numel()
    return 1.0 / world
