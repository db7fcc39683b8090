"""Multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed).

Renderer: rays are independent given the full particle set, so rank g renders
image rows g, g+G, g+2G, ... (block-cyclic: fluid pixels cluster in the image centre, contiguous
blocks would be unbalanced).  Particles (<1 MB) and weights (2.7 MB) are replicated and every rank
builds its own grid.  There is no collective on the data path; `gather_image` is the optional final
exchange when one process needs the whole picture.

The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """(H, ...) -> rows rank, rank+world, ...   (a view; call .contiguous() before handing to kernels)."""
    return t[rank::world]


def rows_of_rank(H: int, rank: int, world: int) -> torch.Tensor:
    return torch.arange(rank, H, world)


def unshard_rows(parts, H: int) -> torch.Tensor:
    """Inverse of shard_rows: parts[g] holds rows g::G of the result."""
    world = len(parts)
    out = parts[0].new_empty((H,) + tuple(parts[0].shape[1:]))
    for g, p in enumerate(parts):
        out[g::world] = p
    return out


def gather_image(local_rows: torch.Tensor, H: int, group=None) -> torch.Tensor:
    """All-gather row shards into the full (H, W, C) image on every rank (NCCL on GPU, gloo on CPU).

    Row counts may differ by one between ranks when H % world != 0: shards are padded to the max."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_rows
    rank = dist.get_rank(group)
    n_max = (H + world - 1) // world
    pad = local_rows.new_zeros((n_max,) + tuple(local_rows.shape[1:]))
    pad[: local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    parts = [bufs[g][: len(range(g, H, world))] for g in range(world)]
    return unshard_rows(parts, H)


def render_image_sharded(net, particles, ro, rays_hw6: torch.Tensor, focal=None, c2w=None, key="rgb1", group=None,
                         **kw) -> torch.Tensor:
    """Render an (H, W) image with rays sharded over the ranks of `group`; returns (H, W, C) everywhere."""
    H, W = rays_hw6.shape[0], rays_hw6.shape[1]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_rows(rays_hw6, rank, world).reshape(-1, 6).contiguous()
    out = net(particles, ro, mine, focal, c2w, **kw)[key]
    return gather_image(out.view(-1, W, out.shape[-1]), H, group)
