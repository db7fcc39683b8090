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


def pack_nerf_weights(params, dtype=_lib.NF_DTYPE_F16, enc_flags=15) -> torch.Tensor:
    """params: 24 CUDA fp32 tensors (weight, bias) x [xyz_encoding_1..8, final, dir, sigma, rgb].  enc_flags (NF_ENC_*): which
    encoding blocks the network was built with (narrower first / skip / dir layers when one is off)."""
    assert len(params) == 24
    ps = [p.detach().to(torch.float32).contiguous() for p in params]
    require_cuda(*ps)
    out = torch.empty(lib().nf_render_packed_weights_bytes(), dtype=torch.uint8, device=ps[0].device)
    arr = (C.c_void_p * 24)(*[p.data_ptr() for p in ps])
    check(lib().nf_render_pack_weights_ex(arr, int(dtype), int(enc_flags), ptr(out), stream_ptr()), "nf_render_pack_weights")
    out._keepalive = ps
    return out


def nerf_mlp(packed: torch.Tensor, records: torch.Tensor, dtype=_lib.NF_DTYPE_F16, sigma_only=False) -> torch.Tensor:
    """Fused positional encoding + NeRF MLP over (n,16) geometry records -> (n,4) [r,g,b,sigma]."""
    require_cuda(packed, records)
    rec = records.detach().to(torch.float32).contiguous()
    out = torch.zeros((rec.shape[0], 4), dtype=torch.float32, device=rec.device)
    ws = _mlp_workspace(rec.device)
    check(lib().nf_nerf_mlp_forward(ptr(packed), int(dtype), ptr(rec), rec.shape[0], int(bool(sigma_only)), ptr(out),
                                    ptr(ws), ws.numel(), stream_ptr()), "nf_nerf_mlp_forward")
    return out


_MLP_WS = {}


def _mlp_workspace(device) -> torch.Tensor:
    """nf_nerf_mlp_forward's staging workspace (constant size), one per device, reused across calls on the current stream."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _MLP_WS.get(key)
    if ws is None:
        ws = _MLP_WS[key] = torch.empty(lib().nf_nerf_mlp_workspace_bytes(), dtype=torch.uint8, device=device)
    return ws


def pack_nerf_weights_bwd(params) -> torch.Tensor:
    """Transposed bf16 slabs for the data-gradient GEMMs of the NeRF MLP backward."""
    assert len(params) == 24
    ps = [p.detach().to(torch.float32).contiguous() for p in params]
    require_cuda(*ps)
    out = torch.empty(lib().nf_render_packed_weights_bwd_bytes(), dtype=torch.uint8, device=ps[0].device)
    arr = (C.c_void_p * 24)(*[p.data_ptr() for p in ps])
    check(lib().nf_render_pack_weights_bwd(arr, ptr(out), stream_ptr()), "nf_render_pack_weights_bwd")
    out._keepalive = ps
    return out


def nerf_mlp_backward(packed_fwd, packed_bwd, records, dout4, dtype=_lib.NF_DTYPE_F16):
    """Backward of `nerf_mlp`: dout4 (n,4) = gradient w.r.t. (pre-sigmoid r, g, b, sigma).  Returns (dfeat (n,272):
    gradient w.r.t. the encoded features [xyz-like 198 | pad 10 | dir-like 54 | pad 10], flat parameter gradients)."""
    require_cuda(packed_fwd, packed_bwd, records, dout4)
    rec = records.detach().to(torch.float32).contiguous()
    g = dout4.detach().to(torch.float32).contiguous()
    n = rec.shape[0]
    dfeat = torch.zeros((n, 272), dtype=torch.float32, device=rec.device)
    dpar = torch.zeros(lib().nf_render_param_count(), dtype=torch.float32, device=rec.device)
    ws = torch.empty(max(lib().nf_nerf_mlp_backward_workspace_bytes(n), 256), dtype=torch.uint8, device=rec.device)
    check(lib().nf_nerf_mlp_backward(ptr(packed_fwd), ptr(packed_bwd), int(dtype), ptr(rec), None, ptr(g), n, ptr(dfeat), ptr(dpar),
                                     ptr(ws), ws.numel(), stream_ptr()), "nf_nerf_mlp_backward")
    return dfeat, dpar


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


# ------------------------------------------------------------------------------------------------
# open3d.ml.torch operator mirrors: swap only the layer, keep models/transmodel.py
# ------------------------------------------------------------------------------------------------
def reduce_subarrays_sum(values: torch.Tensor, row_splits: torch.Tensor) -> torch.Tensor:
    """ml3d.ops.reduce_subarrays_sum (models/transmodel.py:135): segmented sum of `values` over CSR rows."""
    cs = torch.cat([values.new_zeros(1, dtype=torch.float64), torch.cumsum(values.to(torch.float64), 0)])
    return (cs[row_splits[1:]] - cs[row_splits[:-1]]).to(values.dtype)


class ContinuousConv(torch.nn.Module):
    """`open3d.ml.torch.layers.ContinuousConv` with the constructor arguments the reference passes
    (models/transmodel.py:86-95) and its call signature `conv(inp_features, inp_positions, out_positions, extents)`
    (:116,:118,:125); after a call `conv.nns.neighbors_index` (int32) / `.neighbors_row_splits` (int64) hold the
    neighbour lists the reference reads at :135-138.  Runs nf_cconv_forward (one launch group per call); forward only.
    Parameters / buffers as upstream: `kernel` (*kernel_size, in_channels, filters) ~ U(-0.05, 0.05), `bias`, `offset`."""

    def __init__(self, in_channels, filters, kernel_size, activation=None, use_bias=True, align_corners=True,
                 coordinate_mapping="ball_to_cube_radial", interpolation="linear", normalize=True,
                 radius_search_ignore_query_points=False, radius_search_metric="L2", offset=None, window_function=None,
                 use_dense_layer_for_center=False, operand_dtype="fp16", **kwargs):
        super().__init__()
        if list(kernel_size) != [4, 4, 4] or coordinate_mapping != "ball_to_cube_volume_preserving" \
                or interpolation != "linear" or normalize or not align_corners or use_dense_layer_for_center \
                or radius_search_metric != "L2":
            raise _lib.NFError("ContinuousConv: only the configuration of models/transmodel.py:79-98 is implemented "
                               "(4x4x4, linear, ball_to_cube_volume_preserving, normalize=False, L2)")
        if lib().nf_cconv_packed_weights_bytes(int(in_channels), int(filters)) == 0:
            raise _lib.NFError(f"ContinuousConv: {in_channels} -> {filters} channels is not a supported shape")
        self.in_channels, self.filters, self.kernel_size = int(in_channels), int(filters), list(kernel_size)
        self.activation = activation
        self.window_function = window_function       # None or the reference's poly6 (the kernel evaluates poly6 itself)
        self.radius_search_ignore_query_points = bool(radius_search_ignore_query_points)
        self.kernel = torch.nn.Parameter(torch.empty(*kernel_size, in_channels, filters).uniform_(-0.05, 0.05))
        self.bias = torch.nn.Parameter(torch.zeros(filters)) if use_bias else None
        self.register_buffer("offset", torch.zeros(3) if offset is None else torch.as_tensor(offset).float())
        self.operand_dtype = {"fp16": _lib.NF_DTYPE_F16, "bf16": _lib.NF_DTYPE_BF16}[operand_dtype]
        self.nns = None
        self._packed = None
        self._ws = None
        self._overflow = None

    def _weights(self):
        key = (self.kernel.data_ptr(), self.kernel._version, None if self.bias is None else self.bias._version,
               self.offset._version)
        if self._packed is None or self._packed[0] != key:
            if bool((self.offset != 0).any()):
                raise _lib.NFError("non-zero ContinuousConv.offset is not supported (the reference never sets it)")
            k = self.kernel.detach().to(torch.float32).contiguous()
            b = None if self.bias is None else self.bias.detach().to(torch.float32).contiguous()
            require_cuda(k)
            out = torch.empty(lib().nf_cconv_packed_weights_bytes(self.in_channels, self.filters), dtype=torch.uint8,
                              device=k.device)
            check(lib().nf_cconv_pack_weights(ptr(k), ptr(b), self.in_channels, self.filters, self.operand_dtype, ptr(out),
                                              stream_ptr()), "nf_cconv_pack_weights")
            out._keepalive = (k, b)
            self._packed = (key, out)
        return self._packed[1]

    def forward(self, inp_features, inp_positions, out_positions, extents, inp_importance=None,
                fixed_radius_search_hash_table=None, **unused):
        from types import SimpleNamespace
        require_cuda(inp_features, inp_positions, out_positions)
        if inp_importance is not None:
            raise _lib.NFError("ContinuousConv: inp_importance is not used by the reference and not implemented")
        f = lambda t: t.detach().to(torch.float32).contiguous()
        feat, ipos, opos = f(inp_features), f(inp_positions), f(out_positions)
        extent = float(extents)
        n_in, n_out = ipos.shape[0], opos.shape[0]
        dev = opos.device
        grid = fixed_radius_search_hash_table if isinstance(fixed_radius_search_hash_table, Grid) else \
            Grid(ipos, float(CELL_SCALE * 0.5 * extent))
        need = lib().nf_cconv_workspace_bytes(n_in, n_out, self.in_channels, self.filters)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if self._overflow is None or self._overflow.device != dev:
            self._overflow = torch.zeros(2, dtype=torch.int32, device=dev)
        out = torch.empty((n_out, self.filters), dtype=torch.float32, device=dev)
        counts = torch.empty((n_out,), dtype=torch.float32, device=dev)
        nbr = torch.empty((n_out, 128), dtype=torch.int32, device=dev)
        a = _lib.CConvArgs()
        a.grid_in, a.in_feat, a.n_in, a.cin = ptr(grid.ws), ptr(feat), n_in, self.in_channels
        a.out_pos, a.n_out, a.cout = ptr(opos), n_out, self.filters
        a.extent, a.use_window = extent, int(self.window_function is not None)
        a.ignore_same, a.dtype = int(self.radius_search_ignore_query_points), self.operand_dtype
        a.weights, a.out, a.count_out, a.nbr_index_out = ptr(self._weights()), ptr(out), ptr(counts), ptr(nbr)
        a.overflow_out = ptr(self._overflow)
        a.workspace, a.workspace_bytes = ptr(self._ws), self._ws.numel()
        check(lib().nf_cconv_forward(C.byref(a), stream_ptr()), "nf_cconv_forward")
        self._keep = (grid, feat, ipos, opos)
        valid = nbr >= 0
        row_splits = torch.zeros(n_out + 1, dtype=torch.int64, device=dev)
        row_splits[1:] = torch.cumsum(valid.sum(1), 0)
        self.nns = SimpleNamespace(neighbors_index=nbr[valid], neighbors_row_splits=row_splits, neighbor_counts=counts)
        if self.activation is not None:
            out = self.activation(out)
        return _lib.forward_only([inp_features, self.kernel], (out,))[0]
