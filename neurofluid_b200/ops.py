"""Thin operator-level wrappers over the C ABI (mirrors of the third-party call sites the reference
hides its arithmetic behind).  All tensors are CUDA tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

CELL_SCALE = 1.002   # grid cell edge = CELL_SCALE * search radius (include/nf_b200.h: nf_grid_build)


class Grid:
    """Cell-sorted copy of a point set (device workspace of nf_grid_build)."""

    def __init__(self, points: torch.Tensor, cell: float):
        require_cuda(points)
        self.points = points.detach().to(torch.float32).contiguous()
        n = self.points.shape[0]
        nbytes = lib().nf_grid_workspace_bytes(n)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
        check(lib().nf_grid_build(ptr(self.points), n, float(cell), ptr(self.ws), nbytes, stream_ptr()),
              "nf_grid_build")
        self.n, self.cell = n, float(cell)


def ball_query(queries: torch.Tensor, points: torch.Tensor, K: int, radius: float, grid: Grid | None = None):
    """pytorch3d.ops.ball_query semantics for one shared cloud (models/renderer.py:116-118).

    Returns (dists (Q,K) squared/0-padded, idx (Q,K) int64/-1-padded, nn (Q,K,3) 0-padded)."""
    require_cuda(queries, points)
    q = queries.detach().to(torch.float32).contiguous().view(-1, 3)
    grid = grid or Grid(points, CELL_SCALE * radius)
    nq = q.shape[0]
    idx = torch.empty((nq, K), dtype=torch.int32, device=q.device)
    cnt = torch.empty((nq,), dtype=torch.int32, device=q.device)
    check(lib().nf_ballquery_firstk(ptr(grid.ws), grid.n, ptr(q), nq, float(radius), int(K), ptr(idx), ptr(cnt), stream_ptr()),
          "nf_ballquery_firstk")
    idx64 = idx.to(torch.int64)
    valid = idx64 >= 0
    if grid.n == 0:
        return torch.zeros((nq, K), device=q.device), idx64, torch.zeros((nq, K, 3), device=q.device)
    nn = grid.points[idx64.clamp(min=0)] * valid.unsqueeze(-1)
    diff = q.unsqueeze(1) - nn
    sq = diff * diff
    d2 = ((sq[..., 0] + sq[..., 1]) + sq[..., 2]) * valid
    return d2, idx64, nn


def pack_nerf_weights(params, dtype=_lib.NF_DTYPE_F16) -> torch.Tensor:
    """params: 24 CUDA fp32 tensors (weight, bias) x [xyz_encoding_1..8, final, dir, sigma, rgb]."""
    assert len(params) == 24
    ps = [p.detach().to(torch.float32).contiguous() for p in params]
    require_cuda(*ps)
    out = torch.empty(lib().nf_render_packed_weights_bytes(), dtype=torch.uint8, device=ps[0].device)
    arr = (C.c_void_p * 24)(*[p.data_ptr() for p in ps])
    check(lib().nf_render_pack_weights(arr, int(dtype), ptr(out), stream_ptr()), "nf_render_pack_weights")
    out._keepalive = ps
    return out


def nerf_mlp(packed: torch.Tensor, records: torch.Tensor, dtype=_lib.NF_DTYPE_F16, sigma_only=False) -> torch.Tensor:
    """Fused positional encoding + NeRF MLP over (n,16) geometry records -> (n,4) [r,g,b,sigma]."""
    require_cuda(packed, records)
    rec = records.detach().to(torch.float32).contiguous()
    out = torch.zeros((rec.shape[0], 4), dtype=torch.float32, device=rec.device)
    check(lib().nf_nerf_mlp_forward(ptr(packed), int(dtype), ptr(rec), rec.shape[0], int(bool(sigma_only)), ptr(out),
                                    stream_ptr()), "nf_nerf_mlp_forward")
    return out
