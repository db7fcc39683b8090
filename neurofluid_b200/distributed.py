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


# ---------------------------------------------------------------------------------------------------
# Transition model: particle-block sharding with one all-gather per layer
# ---------------------------------------------------------------------------------------------------
def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, equally sized blocks of particle indices (the last block may be short or empty)."""
    per = (n + world - 1) // world
    b = min(rank * per, n)
    return b, min(b + per, n)


def allgather_rows(buf: torch.Tensor, n: int, group=None) -> None:
    """In-place all-gather of the row blocks of `buf` ((n, C) matrix): rank g owns rows shard_bounds(n, g, G)
    and receives everybody else's.  Blocks are padded to a common size for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return
    rank = dist.get_rank(group)
    per = (n + world - 1) // world
    b, e = shard_bounds(n, rank, world)
    mine = buf.new_zeros((per,) + tuple(buf.shape[1:]))
    mine[: e - b] = buf[b:e]
    out = buf.new_empty((world * per,) + tuple(buf.shape[1:]))
    dist.all_gather_into_tensor(out, mine, group=group)
    buf[:n] = out[:n]


_comm_ready = {}


def init_comm(group=None) -> None:
    """Create the library's NCCL communicator for the ranks of `group` (idempotent): rank 0 makes the id
    (nf_comm_unique_id), torch.distributed ships its 128 bytes (plumbing), every rank calls nf_comm_init."""
    import ctypes as C
    from . import _lib
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    key = (id(group), world, rank, torch.cuda.current_device())
    if _comm_ready.get("key") == key:
        return
    L = _lib.lib()
    idbuf = (C.c_ubyte * _lib.NF_COMM_ID_BYTES)()
    if rank == 0:
        _lib.check(L.nf_comm_unique_id(idbuf), "nf_comm_unique_id")
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    idbuf = (C.c_ubyte * _lib.NF_COMM_ID_BYTES)(*t.cpu().tolist())
    _lib.check(L.nf_comm_init(idbuf, rank, world), "nf_comm_init")
    _comm_ready["key"] = key


def transition_step_sharded(net, pos, vel, box, box_feats, group=None):
    """`ParticleNet.forward` with the particles block-sharded over the ranks of `group`.

    Every rank holds the full state (pos, vel) and ends up with the full result.  ONE call into the library
    (nf_transition_step, phase NF_PHASE_SHARDED): each rank computes its own rows of every phase, and the library
    all-gathers, in place on the compute stream and with no host work in between, the fp16 activation rows after layers
    0-2 and the packed (pos, vel, neighbour count, delta) rows at the end -- the "position all-gather per step" of
    BASELINE.json's north_star.  4 NCCL calls per step; bit-identical to the single-GPU step."""
    import ctypes as C
    from . import _lib
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return net(pos, vel, box, box_feats)
    init_comm(group)
    net._poll_overflow()
    p, v, b, bf, outs, ws = net._prepare(pos, vel, box, box_feats, None)
    a = net._args(p, v, b, bf, outs, ws, phase=_lib.NF_PHASE_SHARDED)
    _lib.check(_lib.lib().nf_transition_step(C.byref(a), _lib.stream_ptr()), "nf_transition_step")
    net._post_overflow()
    net.num_fluid_neighbors, net.pos_correction = outs[2], outs[3]
    net._keep = (p, v, b, bf)
    return outs[0], outs[1], outs[2]
