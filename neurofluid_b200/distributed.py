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


def exchange_timeouts() -> int:
    """How many peer-memory waits gave up (~100 s) because a peer never raised its flag: 0 in a healthy run; results are
    undefined otherwise.  Synchronises (one 4-byte copy)."""
    import ctypes as C
    from . import _lib
    cnt = C.c_uint(0)
    _lib.check(_lib.lib().nf_comm_exchange_timeouts(C.byref(cnt)), "nf_comm_exchange_timeouts")
    return int(cnt.value)


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


class _StepGraph:
    """One captured sharded step: static inputs / outputs, the argument block and the CUDA graph that replays the library call."""

    def __init__(self, key, pos_in, vel_in, outs, keep, graph):
        self.key, self.pos_in, self.vel_in, self.outs, self.keep, self.graph = key, pos_in, vel_in, outs, keep, graph


def _capture_step(net, key, p, v, b, bf, ws):
    """Capture nf_transition_step(phase = NF_PHASE_SHARDED) into a CUDA graph (the library only launches on the stream it is
    given: kernels, a memset, and either ncclAllGather or the peer-memory exchange kernels, whose epochs live in device
    memory -- all of it capturable)."""
    import ctypes as C
    from . import _lib
    dev = p.device
    pos_in, vel_in = p.clone(), v.clone()
    n = p.shape[0]
    outs = (torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev), torch.empty((n,), device=dev),
            torch.empty((n, 3), device=dev))
    a = net._args(pos_in, vel_in, b, bf, outs, ws, phase=_lib.NF_PHASE_SHARDED)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):           # warm-up outside the capture (lazy initialisation inside CUDA / NCCL)
        _lib.check(_lib.lib().nf_transition_step(C.byref(a), _lib.stream_ptr()), "nf_transition_step")
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        _lib.check(_lib.lib().nf_transition_step(C.byref(a), _lib.stream_ptr()), "nf_transition_step")
    return _StepGraph(key, pos_in, vel_in, outs, (a, b, bf, ws, net._packed_weights(), net._box_grid(b)), g)


def transition_step_sharded(net, pos, vel, box, box_feats, group=None, graph=False):
    """`ParticleNet.forward` with the particles block-sharded over the ranks of `group`.

    Every rank holds the full state (pos, vel) and ends up with the full result.  ONE call into the library
    (nf_transition_step, phase NF_PHASE_SHARDED): each rank computes its own rows of every phase, and the library
    all-gathers, in place on the compute stream and with no host work in between, the fp16 activation rows after layers
    0-2 and the packed (pos, vel, neighbour count, delta) rows at the end -- the "position all-gather per step" of
    BASELINE.json's north_star.  The step's workspace is registered for the library's peer-memory exchange
    (nf_comm_register_buffer, once per workspace): each of the four exchanges is then a kernel that stores this rank's rows
    into every peer's workspace over NVLink plus a flag wait, not an NCCL call (NF_B200_NO_PEER=1, or a buffer CUDA IPC
    cannot export, keeps ncclAllGather).  With `graph=True` the library call is captured into a CUDA graph once per (scene size,
    container, weights, workspace) and replayed (measured: no faster -- at 8 GPUs a rank's 0.5 ms step is bound by its own
    kernels, not by the host or the exchanges -- so it is off by default).  Bit-identical to the single-GPU step either way."""
    import ctypes as C
    import os
    from . import _lib
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return net(pos, vel, box, box_feats)
    init_comm(group)
    net._poll_overflow()
    p, v, b, bf, outs, ws = net._prepare(pos, vel, box, box_feats, None)
    reg = (ws.data_ptr(), ws.numel())
    if _comm_ready.get("registered") != reg:          # collective: every rank reaches this at the same step
        peer = os.environ.get("NF_B200_NO_PEER", "0") != "1"
        rc = _lib.lib().nf_comm_register_buffer(_lib.ptr(ws) if peer else None, ws.numel() if peer else 0)
        if rc != _lib.NF_E_UNSUPPORTED:
            _lib.check(rc, "nf_comm_register_buffer")
        _comm_ready["registered"] = reg
        _comm_ready["peer_memory"] = peer and rc == 0
        net._step_graph = None
    if graph and os.environ.get("NF_B200_NO_GRAPH", "0") != "1":
        key = (tuple(p.shape), b.data_ptr(), b._version, bf.data_ptr(), bf._version, net._packed_weights().data_ptr(), ws.data_ptr(),
               net.operand_dtype, float(net.time_step), tuple(net._gravity_host))
        sg = getattr(net, "_step_graph", None)
        if sg is None or sg.key != key:
            sg = net._step_graph = _capture_step(net, key, p, v, b, bf, ws)
        sg.pos_in.copy_(p)
        sg.vel_in.copy_(v)
        sg.graph.replay()
        outs = tuple(t.clone() for t in sg.outs)       # the static outputs are overwritten by the next replay
    else:
        a = net._args(p, v, b, bf, outs, ws, phase=_lib.NF_PHASE_SHARDED)
        _lib.check(_lib.lib().nf_transition_step(C.byref(a), _lib.stream_ptr()), "nf_transition_step")
    net._post_overflow()
    net.num_fluid_neighbors, net.pos_correction = outs[2], outs[3]
    net._keep = (p, v, b, bf)
    return outs[0], outs[1], outs[2]
