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
