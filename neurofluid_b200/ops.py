"""Thin operator-level wrappers over the C ABI (mirrors of the third-party call sites the reference
hides its arithmetic behind).  All tensors are CUDA tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr  # noqa: F401 (require_cuda re-exported)

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


def generate_rays(H: int, W: int, focal: float, c2w: torch.Tensor) -> torch.Tensor:
    """get_ray_directions + get_rays (utils/ray_utils.py:85-130) for one view, on the device: (H*W, 6)."""
    require_cuda(c2w)
    m = c2w.detach().to(torch.float32).contiguous()
    if m.shape != (3, 4):
        raise _lib.NFError(f"c2w must be (3,4), got {tuple(m.shape)}")
    rays = torch.empty((H * W, 6), dtype=torch.float32, device=m.device)
    check(lib().nf_generate_rays(int(H), int(W), float(focal), ptr(m), ptr(rays), stream_ptr()), "nf_generate_rays")
    return rays


def nearest_distance(queries: torch.Tensor, points: torch.Tensor, cell: float = 0.1, grid: Grid | None = None,
                     return_index: bool = False):
    """cKDTree(points).query(queries) (utils/point_eval.py:11-14): distance to the nearest point, exact."""
    require_cuda(queries, points)
    q = queries.detach().to(torch.float32).contiguous().view(-1, 3)
    grid = grid or Grid(points, cell)
    dist = torch.empty((q.shape[0],), dtype=torch.float32, device=q.device)
    idx = torch.empty((q.shape[0],), dtype=torch.int32, device=q.device) if return_index else None
    check(lib().nf_nearest_distance(ptr(grid.ws), grid.n, ptr(q), q.shape[0], ptr(dist), ptr(idx), stream_ptr()),
          "nf_nearest_distance")
    return (dist, idx.to(torch.int64)) if return_index else dist


def pair_distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """np.linalg.norm(a - b, axis=-1) (utils/point_eval.py:7-8)."""
    require_cuda(a, b)
    a, b = a.detach().to(torch.float32).contiguous().view(-1, 3), b.detach().to(torch.float32).contiguous().view(-1, 3)
    if a.shape != b.shape:
        raise _lib.NFError(f"pair_distance: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
    out = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    check(lib().nf_pair_distance(ptr(a), ptr(b), a.shape[0], ptr(out), stream_ptr()), "nf_pair_distance")
    return out


def img2mse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torch.mean((x - y) ** 2) (trainer/trainer_e2e.py:24) as a 0-dim device tensor (float64 accumulation)."""
    require_cuda(x, y)
    x, y = x.detach().to(torch.float32).contiguous(), y.detach().to(torch.float32).contiguous()
    if x.shape != y.shape:
        raise _lib.NFError(f"img2mse: shapes differ {tuple(x.shape)} vs {tuple(y.shape)}")
    acc = torch.empty((), dtype=torch.float64, device=x.device)
    check(lib().nf_sqdiff_sum(ptr(x), ptr(y), x.numel(), ptr(acc), stream_ptr()), "nf_sqdiff_sum")
    return (acc / max(x.numel(), 1)).to(torch.float32)


def mse2psnr(mse: torch.Tensor) -> torch.Tensor:
    """-10 log10(mse) (trainer/trainer_e2e.py:25); stays on the device."""
    return -10.0 * torch.log10(mse)
