"""CPU restatements of the third-party operators behind NeuroFluid's hot paths.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by neurofluid_b200/.

Operators (none is vendored in the reference tree; versions pinned by /root/reference/README.md:32-42):

* ``ball_query``            PyTorch3D v0.6.1 ``pytorch3d.ops.ball_query`` -- call site
                            models/renderer.py:116-118.
* ``ContinuousConv``        open3d 0.15.2 ``open3d.ml.torch.layers.ContinuousConv`` -- ctor kwargs
                            at models/transmodel.py:86-95, calls at :116,:118,:125, ``.nns`` read at
                            :135-138.
* ``reduce_subarrays_sum``  open3d 0.15.2 ``open3d.ml.torch.ops.reduce_subarrays_sum`` -- :135.
* ``create_meshgrid``       kornia 0.6.4 -- utils/ray_utils.py:2,97.

PARITY: unpinned against the upstream binaries (cannot be installed here; the reference has no
tests or golden vectors).  Semantics follow SURVEY.md section 8c.  Each operator exists twice -- a
compiled C version (oracle/csrc/nf_oracle.c, used for speed and as the CPU baseline) and a
pure-torch version (``*_torch``) -- and tests/test_oracle.py checks the two against each other.
"""
from __future__ import annotations

import ctypes
import math
import os
from types import SimpleNamespace

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from . import build_oracle
        path = build_oracle.build()
        lib = ctypes.CDLL(path)
        i64, f32p, i64p, i32p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        lib.nfo_ball_query.argtypes = [f32p, i64, f32p, i64, ctypes.c_float, ctypes.c_int, i64p, f32p,
                                       ctypes.c_int]
        lib.nfo_ball_query.restype = None
        lib.nfo_radius_search.argtypes = [f32p, i64, f32p, i64, ctypes.c_float, ctypes.c_int, i64p, i64p,
                                          i32p, f32p, ctypes.c_int]
        lib.nfo_radius_search.restype = None
        lib.nfo_cconv_forward.argtypes = [f32p, f32p, i64, ctypes.c_int, f32p, i64, ctypes.c_float,
                                          ctypes.c_int, f32p, f32p, f32p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, f32p, i64p, ctypes.c_int]
        lib.nfo_cconv_forward.restype = None
        lib.nfo_filter_corners.argtypes = [ctypes.c_float] * 4 + [ctypes.c_int, f32p, i32p, f32p]
        lib.nfo_filter_corners.restype = None
        lib.nfo_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


_DEFAULT_THREADS = 0


def set_default_threads(n: int) -> None:
    """Thread count the compiled operators use when a call passes nthreads=0.  bench.py sets it to the host's core count:
    torchrun exports OMP_NUM_THREADS=1, which would otherwise make the CPU baseline single-threaded."""
    global _DEFAULT_THREADS
    _DEFAULT_THREADS = int(n)


def max_threads() -> int:
    return _DEFAULT_THREADS if _DEFAULT_THREADS > 0 else int(_lib().nfo_max_threads())


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32, device="cpu").contiguous()


# --------------------------------------------------------------------------------------------
# pytorch3d.ops.ball_query
# --------------------------------------------------------------------------------------------
def ball_query_shared(queries: torch.Tensor, points: torch.Tensor, K: int, radius: float, nthreads: int = 0):
    """First-K-by-index ball query of ``queries`` (Q,3) against one point set ``points`` (P,3).

    Returns (dists (Q,K) squared / 0 padded, idx (Q,K) int64 / -1 padded).
    """
    q, p = _f32(queries).reshape(-1, 3), _f32(points).reshape(-1, 3)
    nq = q.shape[0]
    idx = torch.empty((nq, K), dtype=torch.int64)
    d2 = torch.empty((nq, K), dtype=torch.float32)
    _lib().nfo_ball_query(q.data_ptr(), nq, p.data_ptr(), p.shape[0], float(radius), int(K), idx.data_ptr(),
                          d2.data_ptr(), int(nthreads or _DEFAULT_THREADS))
    return d2, idx


def ball_query_shared_torch(queries, points, K, radius):
    """Pure-torch restatement (dense, O(Q*P) memory) -- small cases only; cross-checks the C code."""
    q, p = _f32(queries).reshape(-1, 3), _f32(points).reshape(-1, 3)
    diff = q[:, None, :] - p[None, :, :]
    sq = diff * diff
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    r2 = torch.tensor(radius, dtype=torch.float32) * torch.tensor(radius, dtype=torch.float32)
    hit = d2 < r2
    rank = torch.cumsum(hit.to(torch.int64), dim=1) - 1          # position among hits, index order
    keep = hit & (rank < K)
    idx = torch.full((q.shape[0], K), -1, dtype=torch.int64)
    dd = torch.zeros((q.shape[0], K), dtype=torch.float32)
    qi, pj = torch.nonzero(keep, as_tuple=True)
    idx[qi, rank[qi, pj]] = pj
    dd[qi, rank[qi, pj]] = d2[qi, pj]
    return dd, idx


def ball_query(p1, p2, lengths1=None, lengths2=None, K: int = 500, radius: float = 0.2, return_nn: bool = True):
    """Signature-compatible stand-in for ``pytorch3d.ops.ball_query`` (batched p2).

    The reference always passes p2 = particles repeated per ray (models/renderer.py:113), so when
    every batch entry of p2 aliases the same data we run the shared-cloud path once.
    """
    assert lengths1 is None and lengths2 is None
    N, P1, _ = p1.shape
    same = p2.shape[0] == 1 or bool((p2[0:1] == p2).all())
    if same:
        d2, idx = ball_query_shared(p1.reshape(-1, 3), p2[0], K, radius)
        d2, idx = d2.view(N, P1, K), idx.view(N, P1, K)
    else:
        outs = [ball_query_shared(p1[n], p2[n], K, radius) for n in range(N)]
        d2 = torch.stack([o[0] for o in outs])
        idx = torch.stack([o[1] for o in outs])
    nn = None
    if return_nn:
        # masked_gather: padded (-1) slots -> zeros
        safe = idx.clamp(min=0)
        pts = p2[0] if same else None
        if same:
            nn = pts[safe.reshape(-1)].view(N, P1, K, 3)
        else:
            nn = torch.stack([p2[n][safe[n].reshape(-1)].view(P1, K, 3) for n in range(N)])
        nn = nn * (idx >= 0).unsqueeze(-1).to(nn.dtype)
    return d2.to(p1.device), idx.to(p1.device), (nn.to(p1.device) if nn is not None else None)


# --------------------------------------------------------------------------------------------
# kornia.create_meshgrid
# --------------------------------------------------------------------------------------------
def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)  # (1,H,W,2): [...,0]=x, [...,1]=y


# --------------------------------------------------------------------------------------------
# open3d.ml.torch.ops.reduce_subarrays_sum
# --------------------------------------------------------------------------------------------
def reduce_subarrays_sum(values: torch.Tensor, row_splits: torch.Tensor) -> torch.Tensor:
    cs = torch.cat([values.new_zeros(1), torch.cumsum(values.to(torch.float64), 0).to(values.dtype)])
    return cs[row_splits[1:]] - cs[row_splits[:-1]]


# --------------------------------------------------------------------------------------------
# open3d FixedRadiusSearch
# --------------------------------------------------------------------------------------------
def radius_search(in_pos, out_pos, radius, ignore_same_pos=True, nthreads=0):
    """Returns (neighbors_index int32 (E,), row_splits int64 (n_out+1,), dist2 float (E,))."""
    ip, op = _f32(in_pos), _f32(out_pos)
    n_out = op.shape[0]
    counts = torch.zeros(n_out, dtype=torch.int64)
    lib = _lib()
    nthreads = nthreads or _DEFAULT_THREADS
    lib.nfo_radius_search(ip.data_ptr(), ip.shape[0], op.data_ptr(), n_out, float(radius), int(ignore_same_pos),
                          counts.data_ptr(), None, None, None, int(nthreads))
    rs = torch.zeros(n_out + 1, dtype=torch.int64)
    rs[1:] = torch.cumsum(counts, 0)
    E = int(rs[-1])
    nbr = torch.empty(max(E, 1), dtype=torch.int32)
    d2 = torch.empty(max(E, 1), dtype=torch.float32)
    lib.nfo_radius_search(ip.data_ptr(), ip.shape[0], op.data_ptr(), n_out, float(radius), int(ignore_same_pos),
                          None, rs.data_ptr(), nbr.data_ptr(), d2.data_ptr(), int(nthreads))
    return nbr[:E], rs, d2[:E]


def radius_search_torch(in_pos, out_pos, radius, ignore_same_pos=True):
    ip, op = _f32(in_pos), _f32(out_pos)
    diff = ip[None, :, :] - op[:, None, :]
    sq = diff * diff
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    r2 = torch.tensor(radius, dtype=torch.float32) ** 2
    hit = d2 <= r2
    if ignore_same_pos:
        hit &= ~(diff == 0).all(-1)
    oi, ij = torch.nonzero(hit, as_tuple=True)
    counts = torch.bincount(oi, minlength=op.shape[0])
    rs = torch.zeros(op.shape[0] + 1, dtype=torch.int64)
    rs[1:] = torch.cumsum(counts, 0)
    return ij.to(torch.int32), rs, d2[oi, ij]


# --------------------------------------------------------------------------------------------
# ContinuousConv filter geometry (pure torch mirror of nfo_ball_to_cube / nfo_filter_corners)
# --------------------------------------------------------------------------------------------
def ball_to_cube_volume_preserving(v: torch.Tensor) -> torch.Tensor:
    """v: (...,3) in the unit ball -> (...,3) in the cube [-1,1]^3."""
    x, y, z = v[..., 0].clone(), v[..., 1].clone(), v[..., 2].clone()
    sq = x * x + y * y + z * z
    n = torch.sqrt(sq)
    zero = sq < 1e-12
    xy2 = x * x + y * y
    cap = 1.25 * z * z > xy2
    s_cap = torch.sqrt(3.0 * n / (n + z.abs()).clamp_min(1e-30))
    s_cyl = n / torch.sqrt(xy2.clamp_min(1e-30))
    X = torch.where(cap, x * s_cap, x * s_cyl)
    Y = torch.where(cap, y * s_cap, y * s_cyl)
    Z = torch.where(cap, torch.sign(z) * n, z * 1.5)
    nxy2 = X * X + Y * Y
    nxy = torch.sqrt(nxy2)
    c = 4.0 / math.pi
    xdom = Y.abs() <= X.abs()
    tx = torch.sign(X) * nxy
    ty = torch.sign(Y) * nxy
    safe_x = torch.where(X == 0, torch.ones_like(X), X)
    safe_y = torch.where(Y == 0, torch.ones_like(Y), Y)
    X2 = torch.where(xdom, tx, ty * c * torch.atan(X / safe_y))
    Y2 = torch.where(xdom, tx * c * torch.atan(Y / safe_x), ty)
    tiny = nxy2 < 1e-12
    X2 = torch.where(tiny, torch.zeros_like(X2), X2)
    Y2 = torch.where(tiny, torch.zeros_like(Y2), Y2)
    out = torch.stack([X2, Y2, Z], -1)
    return torch.where(zero.unsqueeze(-1), torch.zeros_like(out), out)


def filter_corners_torch(rel: torch.Tensor, inv_radius: float, size: int, offset: torch.Tensor):
    """rel (E,3) -> (cell (E,8) int64, w (E,8)); cell = (z*size + y)*size + x."""
    c = ball_to_cube_volume_preserving(rel * inv_radius) * 0.5
    t = (c + 0.5 + offset.view(1, 3)) * float(size - 1)
    fl = torch.floor(t)
    f = t - fl
    i0 = fl.to(torch.int64).clamp(0, size - 1)
    i1 = (fl.to(torch.int64) + 1).clamp(0, size - 1)
    cells, ws = [], []
    for k in range(8):
        bx, by, bz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        ix = i1[:, 0] if bx else i0[:, 0]
        iy = i1[:, 1] if by else i0[:, 1]
        iz = i1[:, 2] if bz else i0[:, 2]
        cells.append((iz * size + iy) * size + ix)
        ws.append((f[:, 0] if bx else 1 - f[:, 0]) * (f[:, 1] if by else 1 - f[:, 1]) * (f[:, 2] if bz else 1 - f[:, 2]))
    return torch.stack(cells, 1), torch.stack(ws, 1)


def cconv_forward(in_feat, in_pos, out_pos, extent, kernel, bias, offset, ignore_same_pos=True, use_window=True,
                  nthreads=0):
    """Compiled ContinuousConv forward. kernel (s,s,s,cin,cout). Returns (out (n_out,cout), counts int64)."""
    ip, op, ft = _f32(in_pos), _f32(out_pos), _f32(in_feat)
    kr = _f32(kernel)
    size, cin, cout = kr.shape[0], kr.shape[3], kr.shape[4]
    b = _f32(bias) if bias is not None else None
    off = _f32(offset) if offset is not None else torch.zeros(3)
    out = torch.empty((op.shape[0], cout), dtype=torch.float32)
    counts = torch.empty(op.shape[0], dtype=torch.int64)
    _lib().nfo_cconv_forward(ip.data_ptr(), ft.data_ptr(), ip.shape[0], cin, op.data_ptr(), op.shape[0],
                             float(extent), size, kr.data_ptr(), b.data_ptr() if b is not None else None,
                             off.data_ptr(), cout, int(ignore_same_pos), int(use_window), out.data_ptr(),
                             counts.data_ptr(), int(nthreads or _DEFAULT_THREADS))
    return out, counts


def cconv_forward_torch(in_feat, in_pos, out_pos, extent, kernel, bias, offset, ignore_same_pos=True,
                        window_fn=None, nns=None):
    """Differentiable pure-torch ContinuousConv forward (autograd flows to kernel, bias, in_feat)."""
    radius = 0.5 * float(extent)
    if nns is None:
        nbr, rs, d2 = radius_search(in_pos, out_pos, radius, ignore_same_pos)
    else:
        nbr, rs, d2 = nns
    n_out = out_pos.shape[0]
    size, cin, cout = kernel.shape[0], kernel.shape[3], kernel.shape[4]
    counts = rs[1:] - rs[:-1]
    oi = torch.repeat_interleave(torch.arange(n_out), counts)
    nj = nbr.to(torch.int64)
    rel = (in_pos.detach().float()[nj] - out_pos.detach().float()[oi])
    cells, w = filter_corners_torch(rel, 2.0 / float(extent), size, offset.detach().float())
    if window_fn is not None:
        a = window_fn(d2 / (radius * radius))
        w = w * a.unsqueeze(-1)
    # patch (n_out, ncell, cin) accumulated with index_add, then one matmul against the filter
    patch = torch.zeros(n_out * size ** 3, cin, dtype=in_feat.dtype)
    f = in_feat[nj]
    for k in range(8):
        patch = patch.index_add(0, oi * size ** 3 + cells[:, k], f * w[:, k:k + 1])
    out = patch.view(n_out, size ** 3 * cin) @ kernel.reshape(size ** 3 * cin, cout)
    if bias is not None:
        out = out + bias
    return out, (nbr, rs, d2)


class ContinuousConv(torch.nn.Module):
    """Stand-in for ``open3d.ml.torch.layers.ContinuousConv`` with the kwargs the reference uses
    (models/transmodel.py:86-95).  Parameters/buffers match upstream: ``kernel`` (*ks, cin, cout)
    ~ U(-0.05, 0.05), ``bias`` (cout) zeros, buffer ``offset`` zeros(3)."""

    def __init__(self, in_channels, filters, kernel_size, activation=None, use_bias=True,
                 kernel_initializer=None, bias_initializer=None, align_corners=True, coordinate_mapping="ball_to_cube_radial",
                 interpolation="linear", normalize=True, radius_search_ignore_query_points=False,
                 radius_search_metric="L2", offset=None, window_function=None, use_dense_layer_for_center=False,
                 dense_kernel_initializer=None, dense_kernel_regularizer=None, in_channels_=None, **kwargs):
        super().__init__()
        assert list(kernel_size) == [kernel_size[0]] * 3
        assert coordinate_mapping == "ball_to_cube_volume_preserving" and interpolation == "linear"
        assert not normalize and align_corners and not use_dense_layer_for_center
        self.in_channels, self.filters, self.kernel_size = in_channels, filters, list(kernel_size)
        self.activation = activation
        self.window_function = window_function
        self.radius_search_ignore_query_points = radius_search_ignore_query_points
        self.kernel = torch.nn.Parameter(torch.empty(*kernel_size, in_channels, filters).uniform_(-0.05, 0.05))
        self.bias = torch.nn.Parameter(torch.zeros(filters)) if use_bias else None
        self.register_buffer("offset", torch.zeros(3) if offset is None else torch.as_tensor(offset).float())
        self.nns = None

    def forward(self, inp_features, inp_positions, out_positions, extents, inp_importance=None,
                fixed_radius_search_hash_table=None, user_neighbors_index=None, user_neighbors_row_splits=None,
                user_neighbors_importance=None):
        extent = float(extents)
        out, (nbr, rs, d2) = cconv_forward_torch(inp_features, inp_positions, out_positions, extent, self.kernel,
                                                 self.bias, self.offset, self.radius_search_ignore_query_points,
                                                 self.window_function)
        self.nns = SimpleNamespace(neighbors_index=nbr, neighbors_row_splits=rs, neighbors_distance=d2)
        if self.activation is not None:
            out = self.activation(out)
        return out
