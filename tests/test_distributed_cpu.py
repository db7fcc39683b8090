"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray row sharding + image gather, particle
block sharding + per-layer row all-gather.  The kernels themselves need a GPU; what is checked here is
that the sharded orchestration reassembles exactly what an unsharded run produces."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurofluid_b200 import distributed as nfd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        # ---- renderer: block-cyclic image rows, odd H so the shards are ragged
        H, W = 7, 5
        img = torch.arange(H * W * 3, dtype=torch.float32).view(H, W, 3)
        mine = nfd.shard_rows(img, rank, world).contiguous()
        ok &= torch.equal(mine, img[rank::world])
        full = nfd.gather_image(mine * 1.0, H)
        ok &= torch.equal(full, img)

        class FakeNet:                                      # stands in for RenderNet: rgb = f(ray)
            def __call__(self, particles, ro, rays, focal=None, c2w=None, **kw):
                return {"rgb1": rays[:, :3] * 2 + 1}
        rays = torch.randn(H, W, 6, generator=torch.Generator().manual_seed(0))
        out = nfd.render_image_sharded(FakeNet(), None, None, rays)
        ok &= torch.equal(out, rays[..., :3] * 2 + 1)
        # ---- transition: contiguous particle blocks, ragged last block, in-place row all-gather
        n = 11
        ref = torch.arange(n * 4, dtype=torch.float32).view(n, 4)
        b, e = nfd.shard_bounds(n, rank, world)
        buf = torch.full((n, 4), -1.0)
        buf[b:e] = ref[b:e]
        nfd.allgather_rows(buf, n)
        ok &= torch.equal(buf, ref)
        h = torch.zeros(n, 6, dtype=torch.float16)
        h[b:e] = (ref[b:e, :1] * 0.5).half()
        nfd.allgather_rows(h, n)
        ok &= torch.equal(h, (ref[:, :1] * 0.5).half().expand(n, 6))
        # bounds tile [0, n) exactly, also when there are more ranks than particles
        cover = []
        for g in range(world):
            cover += list(range(*nfd.shard_bounds(n, g, world)))
        ok &= cover == list(range(n))
        ok &= [nfd.shard_bounds(1, g, 2) for g in range(2)] == [(0, 1), (1, 1)]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharding_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_unshard_rows_inverse():
    t = torch.arange(30).view(10, 3)
    for world in (1, 2, 3, 4):
        parts = [nfd.shard_rows(t, g, world) for g in range(world)]
        assert torch.equal(nfd.unshard_rows(parts, 10), t)
