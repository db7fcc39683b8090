"""RenderNet -- drop-in for the reference's `models/renderer.py:RenderNet` on B200.

Same constructor, same `forward / coarse_rendering / fine_rendering / set_ro` signatures, same
result-dict keys, shapes and dtypes, same state-dict layout; the arithmetic runs in
libnf_b200.so (csrc/nf_grid.cu, nf_render.cu, nf_mlp.cu) through the C ABI of include/nf_b200.h.
There is no CPU or eager-torch fallback: CPU tensors or a missing extension raise.

Differences a caller can observe (all documented in DESIGN.md):
  * any number of rays per call (the reference needs `ray_chunk=1024` to bound its O(R*P) repeat);
  * tensor-core operands are fp16 (or bf16) with fp32 accumulation -> outputs agree with the fp32
    reference to ~1e-4 relative L2, not bit for bit;  integer outputs (num_nn_*, mask_*) are exact;
  * training: when autograd is recording and the particles or parameters require grad, the call becomes one autograd
    node whose backward runs nf_render_backward (tcgen05 dgrad / wgrad of both MLPs, gradient to the particle positions);
  * `fine_rendering` works (the reference's raises on every shipped config, models/renderer.py:175,322).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import NFError, check, lib, ptr, require_cuda, stream_ptr
from .nerf import Embedding, NeRF
from .ops import CELL_SCALE, Grid, pack_nerf_weights


class RenderNet(nn.Module):
    def __init__(self, cfg, near, far, operand_dtype: str = "fp16", max_rays_per_launch: int = 131072,
                 search: str = "auto"):
        super().__init__()
        self.cfg = cfg
        self.near, self.far = near, far
        self.N_samples = cfg.ray.N_samples
        self.N_importance = cfg.ray.N_importance
        self.raduis = cfg.NN_search.search_raduis_scale * cfg.NN_search.particle_radius   # sic (reference spelling)
        self.fix_radius = cfg.NN_search.fix_radius
        self.num_neighbor = cfg.NN_search.N_neighbor
        enc = cfg.encoding
        self.include_ray = not enc.exclude_ray            # models/renderer.py:100-109
        self.same_smooth_factor = bool(getattr(enc, "same_smooth_factor", False))
        # encoding ablations (models/renderer.py:152-175): a disabled block narrows the networks' inputs; the kernels still
        # produce all six encodings and the weight packer leaves the disabled block's columns zero (nf_render_pack_weights_ex)
        self.enc_flags = (1 if enc.density else 0) | (2 if enc.smoothed_pos else 0) | (4 if enc.var else 0) | \
            (8 if enc.smoothed_dir else 0)
        if not self.fix_radius:
            raise NFError("fix_radius=False has no live code path in the reference (models/renderer.py:119-121)")
        self.embedding_xyz = Embedding(3, 10)
        self.embedding_dir = Embedding(3, 4)
        self.embedding_density = Embedding(1, 4)
        in_xyz = self.embedding_xyz.out_channels * (1 + bool(enc.smoothed_pos) + bool(enc.var)) + \
            (self.embedding_density.out_channels if enc.density else 0)                          # 198 with everything on
        in_dir = self.embedding_dir.out_channels * (1 + bool(enc.smoothed_dir))                  # 54
        self.nerf_coarse = NeRF(in_channels_xyz=in_xyz, in_channels_dir=in_dir)
        self.nerf_fine = NeRF(in_channels_xyz=in_xyz, in_channels_dir=in_dir)
        self.operand_dtype = {"fp16": _lib.NF_DTYPE_F16, "bf16": _lib.NF_DTYPE_BF16}[operand_dtype]
        self.max_rays_per_launch = int(max_rays_per_launch)
        self.search = {"auto": _lib.NF_SEARCH_AUTO, "stream": _lib.NF_SEARCH_STREAM, "sweep": _lib.NF_SEARCH_SWEEP}[search]
        self._packed = {}      # net name -> (version key, packed tensor)
        self._ws = None
        self._tables = {}
        self._grid_cache = None    # (key, Grid, particles tensor kept alive so that its address cannot be reused)
        self.last_stats = None
        self.return_num_nn = True  # False: skip the int64 num_nn_* tensors (20 % of the fine stage's DRAM writes);
                                   # evaluation loops that only read rgb / mask switch it off (pipeline.rollout_and_render)
        self.save_neighbors = False  # True: keep every record row's neighbour list (debug_view(); parity tests)
        self._debug = None

    # ------------------------------------------------------------------ reference helpers
    def set_ro(self, cw):
        return cw[:, 3]

    # ------------------------------------------------------------------ internals
    def _packed_weights(self, name):
        net = getattr(self, name)
        params = net.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params) + (self.operand_dtype,)
        hit = self._packed.get(name)
        if hit is None or hit[0] != key:
            self._packed[name] = (key, pack_nerf_weights(params, self.operand_dtype, self.enc_flags))
        return self._packed[name][1]

    def _sample_tables(self, device, use_disp):
        key = (str(device), bool(use_disp), self.N_samples, self.N_importance, float(self.near), float(self.far))
        if key not in self._tables:
            t = torch.linspace(0, 1, self.N_samples)                       # utils/ray_utils.py:236-240
            z = 1 / (1 / self.near * (1 - t) + 1 / self.far * t) if use_disp else self.near * (1 - t) + self.far * t
            u = torch.linspace(0.0, 1.0, max(self.N_importance, 1))        # utils/ray_utils.py:186-188 (det=True)
            self._tables[key] = (z.float().to(device), u.float().to(device))
        return self._tables[key]

    def _workspace(self, n_rays, n_imp, device, flags=0):
        need = lib().nf_render_workspace_bytes_ex(n_rays, self.N_samples, n_imp, int(self.num_neighbor), int(flags))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _grid(self, particles):
        """Cell-sorted grid of the particle set, rebuilt only when the tensor changes: a caller that renders one
        particle set in many ray chunks (trainer/basetrainer.py:282-289: 625 chunks per 800x800 image) builds it once."""
        key = (particles.data_ptr(), particles._version, tuple(particles.shape), str(particles.device), float(self.raduis))
        hit = self._grid_cache
        if hit is None or hit[0] != key:
            self._grid_cache = (key, Grid(particles, CELL_SCALE * self.raduis), particles)
        return self._grid_cache[1]

    def debug_view(self):
        """After a forward with `save_neighbors=True` (single launch): the evaluated rows of both passes as tensors --
        {"rec0","rowid0","nbr0","rec1","rowid1","nbr1"}: geometry records (rows,16), rowid = ray*S + sample, neighbour
        indices (rows,K) int32, -1 padded (the production search's answer; tests compare it with the oracle)."""
        if self._debug is None:
            raise NFError("debug_view(): run a forward with save_neighbors=True first")
        ws, R, NI = self._debug
        v = _lib.RenderWsView()
        check(lib().nf_render_workspace_view(R, self.N_samples, NI, int(self.num_neighbor), _lib.NF_RENDER_SAVE_NEIGHBORS,
                                             C.byref(v)), "nf_render_workspace_view")
        cnt = ws[v.counters: v.counters + 64].view(torch.int32).tolist()
        K = int(self.num_neighbor)
        out = {}
        for tag, n in (("0", cnt[0]), ("1", cnt[1] if NI > 0 else 0)):
            rec, rid, nbr = getattr(v, "rec" + tag), getattr(v, "rowid" + tag), getattr(v, "nbr" + tag)
            out["rec" + tag] = ws[rec: rec + n * 64].view(torch.float32).view(n, 16).clone()
            out["rowid" + tag] = ws[rid: rid + n * 4].view(torch.int32).clone()
            out["nbr" + tag] = ws[nbr: nbr + n * K * 4].view(torch.int32).view(n, K).clone()
        if NI > 0:
            S1 = self.N_samples + NI
            out["z1"] = ws[v.z1: v.z1 + R * S1 * 4].view(torch.float32).view(R, S1).clone()   # merged depths (rays that miss: unset)
        return out

    def _packed_weights_bwd(self, name):
        """Transposed bf16 slabs for the data-gradient GEMMs (nf_render_pack_weights_bwd), cached like the forward pack."""
        net = getattr(self, name)
        params = net.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._packed.get(name + "/bwd")
        if hit is None or hit[0] != key:
            ps = [p.detach().to(torch.float32).contiguous() for p in params]
            require_cuda(*ps)
            out = torch.empty(lib().nf_render_packed_weights_bwd_bytes(), dtype=torch.uint8, device=ps[0].device)
            arr = (C.c_void_p * 24)(*[p.data_ptr() for p in ps])
            check(lib().nf_render_pack_weights_bwd_ex(arr, int(self.enc_flags), ptr(out), stream_ptr()), "nf_render_pack_weights_bwd")
            out._keepalive = ps
            self._packed[name + "/bwd"] = (key, out)
        return self._packed[name + "/bwd"][1]

    def _draw_jitter(self, R, S0, NI, perturb, noise_std, fine, z_tab, dev, given=None):
        """The reference's random numbers, drawn with torch's generator in the reference's order (coarse_sample_ray's
        rand, render_image's randn, sample_pdf's rand, render_image's randn: utils/ray_utils.py:245-253,186-190,
        models/renderer.py:192-194) -- under the same seed and default device the draws are the reference's own.
        Returns (z (R,S0) or None, u (R,NI) or None, noise0, noise1).  `given` (tests): pre-drawn tensors by name."""
        given = given or {}
        z = u = n0 = n1 = None
        if perturb > 0:
            zv = z_tab.expand(R, S0)
            mid = 0.5 * (zv[:, :-1] + zv[:, 1:])
            upper, lower = torch.cat([mid, zv[:, -1:]], -1), torch.cat([zv[:, :1], mid], -1)
            r = given.get("z_rand")
            r = torch.rand((R, S0), device=dev) if r is None else r.to(dev)
            z = (lower + (upper - lower) * (perturb * r)).contiguous()
        if noise_std > 0:
            r = given.get("noise0")
            n0 = ((torch.randn((R, S0), device=dev) if r is None else r.to(dev)) * noise_std).contiguous()
        if fine and perturb != 0:                   # det = (perturb == 0)
            r = given.get("u")
            u = (torch.rand((R, NI), device=dev) if r is None else r.to(dev)).contiguous()
        if fine and noise_std > 0:
            r = given.get("noise1")
            n1 = ((torch.randn((R, S0 + NI), device=dev) if r is None else r.to(dev)) * noise_std).contiguous()
        return z, u, n0, n1

    def _run(self, mode, physical_particles, ro, rays, use_disp, perturb, noise_std, white_background, _train=False):
        if not _train and torch.is_grad_enabled() and (
                (isinstance(physical_particles, torch.Tensor) and physical_particles.requires_grad)
                or any(p.requires_grad for p in self.parameters())):
            return _render_with_grad(self, mode, physical_particles, ro, rays, use_disp, perturb, noise_std, white_background)
        require_cuda(physical_particles, rays)
        dev = rays.device
        particles = physical_particles.detach().to(torch.float32).contiguous()
        rays = rays.detach().to(torch.float32).contiguous()
        # the camera position stays on the device (no stream drain per call); a CPU tensor / list is passed by value
        ro_t = torch.as_tensor(ro).detach()
        ro_dev = ro_t.to(torch.float32).contiguous() if ro_t.is_cuda else None
        ro_host = [0.0, 0.0, 0.0] if ro_dev is not None else [float(v) for v in ro_t.float().tolist()]
        R, S0 = rays.shape[0], self.N_samples
        fine = mode != _lib.NF_RENDER_COARSE
        NI = self.N_importance if fine else 0
        if mode != _lib.NF_RENDER_COARSE and self.N_importance <= 0:
            if mode == _lib.NF_RENDER_FINE:
                raise AssertionError("N_importance > 0 required")          # models/renderer.py:347
            mode, fine, NI = _lib.NF_RENDER_COARSE, False, 0
        S1 = S0 + NI
        z_tab, u_tab = self._sample_tables(dev, use_disp)
        jz, ju, jn0, jn1 = (None, None, None, None)
        if perturb != 0 or noise_std != 0:
            jz, ju, jn0, jn1 = self._draw_jitter(R, S0, NI, perturb, noise_std, fine, z_tab, dev, getattr(self, "_given_jitter", None))
        grid = self._grid(particles)
        wc = self._packed_weights("nerf_coarse")
        wf = self._packed_weights("nerf_fine") if fine else None
        want0 = mode != _lib.NF_RENDER_FINE
        out = {}
        f32 = dict(dtype=torch.float32, device=dev)
        if want0:
            out["rgb0"] = torch.empty((R, 3), **f32)
            out["depth0"] = torch.empty((R,), **f32)
            out["opacity0"] = torch.empty((R,), **f32)
            if self.return_num_nn:
                out["num_nn_0"] = torch.empty((R, S0, 1), dtype=torch.int64, device=dev)
            out["mask_0"] = torch.empty((R, 1), **f32)
        if fine:
            out["rgb1"] = torch.empty((R, 3), **f32)
            out["depth1"] = torch.empty((R,), **f32)
            out["opacity1"] = torch.empty((R,), **f32)
            if self.return_num_nn:
                out["num_nn_1"] = torch.empty((R, S1, 1), dtype=torch.int64, device=dev)
            out["mask_1"] = torch.empty((R, 1), **f32)
        chunk = max(1, min(self.max_rays_per_launch, R))
        flags = _lib.NF_RENDER_SAVE_NEIGHBORS if (self.save_neighbors or _train) else 0
        if flags and R > chunk:
            raise NFError("a forward that keeps its neighbour lists (training, save_neighbors) must fit one launch: "
                          f"{R} rays > max_rays_per_launch = {self.max_rays_per_launch}")
        if _train:     # several forwards may precede one backward (trainer/trainer_e2e.py:219-244): each keeps its own workspace
            ws = torch.empty(max(lib().nf_render_workspace_bytes_ex(max(chunk, 1), S0, NI, int(self.num_neighbor), flags), 256),
                             dtype=torch.uint8, device=dev)
        else:
            ws = self._workspace(chunk, NI, dev, flags)
        nchunks = (R + chunk - 1) // chunk
        stats = torch.zeros((max(nchunks, 1), 16), dtype=torch.int32, device=dev)
        st = stream_ptr()
        for ci, r0 in enumerate(range(0, R, chunk)):
            r1 = min(r0 + chunk, R)
            a = _lib.RenderArgs()
            a.grid_ws, a.particles, a.n_particles = ptr(grid.ws), ptr(particles), particles.shape[0]
            a.rays, a.n_rays = C.c_void_p(rays.data_ptr() + r0 * 24), r1 - r0
            a.ro = (C.c_float * 3)(*ro_host)
            a.ro_dev = ptr(ro_dev)
            a.flags = flags
            a.z_coarse, a.u_importance, a.n_coarse, a.n_importance = ptr(z_tab), ptr(u_tab), S0, NI
            off = lambda t, per_ray: None if t is None else C.c_void_p(t.data_ptr() + r0 * per_ray * 4)
            if jz is not None:
                a.z_coarse, a.z_stride = off(jz, S0), S0
            if ju is not None:
                a.u_importance, a.u_stride = off(ju, NI), NI
            a.noise0, a.noise1 = off(jn0, S0), off(jn1, S1)
            a.include_ray, a.same_smooth_factor = int(self.include_ray), int(self.same_smooth_factor)
            a.radius, a.K, a.search = float(self.raduis), int(self.num_neighbor), self.search
            a.mode, a.use_mask, a.white_background = int(mode), int(bool(self.cfg.use_mask)), int(bool(white_background))
            a.dtype = self.operand_dtype
            a.weights_coarse, a.weights_fine = ptr(wc), ptr(wf)

            def sl(name, per_ray_bytes):
                t = out.get(name)
                return None if t is None else C.c_void_p(t.data_ptr() + r0 * per_ray_bytes)
            a.rgb0, a.depth0, a.opacity0 = sl("rgb0", 12), sl("depth0", 4), sl("opacity0", 4)
            a.num_nn0, a.mask0 = sl("num_nn_0", 8 * S0), sl("mask_0", 4)
            a.rgb1, a.depth1, a.opacity1 = sl("rgb1", 12), sl("depth1", 4), sl("opacity1", 4)
            a.num_nn1, a.mask1 = sl("num_nn_1", 8 * S1), sl("mask_1", 4)
            a.workspace, a.workspace_bytes = ptr(ws), ws.numel()
            a.stats = C.c_void_p(stats.data_ptr() + ci * 64)
            check(lib().nf_render_forward(C.byref(a), st), "nf_render_forward")
        self.last_stats = stats       # device tensor; .sum(0) = [rows0, rows1, active0, active1]
        self._keep = (grid, particles, rays, ro_dev)
        self._debug = (ws, R, NI) if flags and R > 0 else None
        if _train:
            saved = dict(args=a if R > 0 else None, ws=ws, R=R, NI=NI, S0=S0, n_particles=particles.shape[0],
                         keep=(grid, particles, rays, ro_dev, z_tab, u_tab, wc, wf, stats, jz, ju, jn0, jn1))
            return out, saved
        return out

    # ------------------------------------------------------------------ reference API
    def forward(self, physical_particles, ro, rays, focal=None, c2w=None, use_disp=False, perturb=0, noise_std=0.,
                white_background=True):
        """models/renderer.py:211-270.  `focal` and `c2w` are accepted and unused, as in the reference."""
        return self._run(_lib.NF_RENDER_FORWARD, physical_particles, ro, rays, use_disp, perturb, noise_std,
                         white_background)

    def coarse_rendering(self, physical_particles, ro, rays, focal=None, c2w=None, use_disp=False, perturb=0,
                         noise_std=0., white_background=True):
        """models/renderer.py:273-307."""
        return self._run(_lib.NF_RENDER_COARSE, physical_particles, ro, rays, use_disp, perturb, noise_std,
                         white_background)

    def fine_rendering(self, physical_particles, ro, rays, focal=None, c2w=None, use_disp=False, perturb=0,
                       noise_std=0., white_background=True):
        """models/renderer.py:310-369 (sigma-only coarse pass, then the fine pass)."""
        return self._run(_lib.NF_RENDER_FINE, physical_particles, ro, rays, use_disp, perturb, noise_std,
                         white_background)


class _RenderFunction(torch.autograd.Function):
    """Autograd node of one RenderNet call: forward = nf_render_forward with the neighbour lists kept, backward =
    nf_render_backward (csrc/nf_render_bwd.cu, nf_mlp_bwd.cu).  Inputs that get gradients: the particle positions and the
    48 parameter tensors of the two NeRF MLPs (models/renderer.py:43-44)."""

    @staticmethod
    def forward(ctx, net, call, particles, *params):
        mode, ro, rays, use_disp, perturb, noise_std, white_background = call
        with torch.no_grad():
            out, saved = net._run(mode, particles, ro, rays, use_disp, perturb, noise_std, white_background, _train=True)
        keys = list(out.keys())
        ctx.net, ctx.saved, ctx.keys = net, saved, keys
        ctx.mark_non_differentiable(*[out[k] for k in keys if k.startswith(("num_nn", "mask"))])
        net._train_keys = keys
        return tuple(out[k] for k in keys)

    @staticmethod
    def backward(ctx, *grads):
        net, sv = ctx.net, ctx.saved
        n_par = int(lib().nf_render_param_count())
        dev = sv["ws"].device
        d_particles = torch.zeros((sv["n_particles"], 3), dtype=torch.float32, device=dev)
        flats = [torch.zeros(n_par, dtype=torch.float32, device=dev) for _ in range(2)]
        if sv["args"] is not None:
            g = {k: (None if t is None else t.detach().to(torch.float32).contiguous()) for k, t in zip(ctx.keys, grads)}
            v = _lib.RenderWsView()
            check(lib().nf_render_workspace_view(sv["R"], sv["S0"], sv["NI"], int(net.num_neighbor), _lib.NF_RENDER_SAVE_NEIGHBORS,
                                                 C.byref(v)), "nf_render_workspace_view")
            rows = sv["ws"][v.counters: v.counters + 8].view(torch.int32).tolist()          # syncs: sizes the backward
            need = lib().nf_render_backward_workspace_bytes(sv["R"], sv["S0"], sv["NI"], int(rows[0]), int(rows[1]))
            bws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
            b = _lib.RenderBwdArgs()
            b.fwd = C.pointer(sv["args"])
            wbc = net._packed_weights_bwd("nerf_coarse")
            wbf = net._packed_weights_bwd("nerf_fine") if sv["NI"] > 0 else None
            b.weights_coarse_bwd, b.weights_fine_bwd = ptr(wbc), ptr(wbf)
            b.d_rgb0, b.d_depth0, b.d_opacity0 = ptr(g.get("rgb0")), ptr(g.get("depth0")), ptr(g.get("opacity0"))
            b.d_rgb1, b.d_depth1, b.d_opacity1 = ptr(g.get("rgb1")), ptr(g.get("depth1")), ptr(g.get("opacity1"))
            b.d_particles, b.d_params_coarse, b.d_params_fine = ptr(d_particles), ptr(flats[0]), ptr(flats[1])
            b.workspace, b.workspace_bytes = ptr(bws), bws.numel()
            check(lib().nf_render_backward(C.byref(b), stream_ptr()), "nf_render_backward")
        # the library's flat layout is the full-width network (198 / 454 / 310 input columns); with encoding ablations the
        # module's own matrices keep only the enabled blocks' columns
        full_in = [198, 256, 256, 256, 454, 256, 256, 256, 256, 310, 256, 128]
        cx = [k for k in range(198) if _enc_col(k, net.enc_flags, False) >= 0]
        cd = [k for k in range(54) if _enc_col(k, net.enc_flags, True) >= 0]
        pgrads = []
        for flat, name in zip(flats, ("nerf_coarse", "nerf_fine")):
            o = 0
            params = getattr(net, name).ordered_params()
            for li in range(12):
                w, bias = params[2 * li], params[2 * li + 1]
                gw = flat[o:o + w.shape[0] * full_in[li]].view(w.shape[0], full_in[li])
                o += gw.numel()
                if net.enc_flags != 15:
                    if li == 0:
                        gw = gw[:, cx]
                    elif li == 4:
                        gw = torch.cat([gw[:, cx], gw[:, 198:]], 1)
                    elif li == 9:
                        gw = torch.cat([gw[:, :256], gw[:, [256 + k for k in cd]]], 1)
                pgrads.append(gw.reshape(w.shape))
                pgrads.append(flat[o:o + bias.numel()].view(bias.shape))
                o += bias.numel()
        need_in = ctx.needs_input_grad
        return (None, None, d_particles if need_in[2] else None) + tuple(gp if need_in[3 + i] else None for i, gp in enumerate(pgrads))


def _enc_col(k, flags, is_dir):
    """Mirror of enc_col_xyz / enc_col_dir (csrc/nf_mlp.cuh): fixed-layout column -> the network's own input column, or -1."""
    if is_dir:
        return k if k < 27 else (27 + (k - 27) if (flags & 8) and k < 54 else -1)
    d, s_ = (9 if flags & 1 else 0), (63 if flags & 2 else 0)
    if k < 63:
        return k
    if k < 72:
        return 63 + (k - 63) if flags & 1 else -1
    if k < 135:
        return 63 + d + (k - 72) if flags & 2 else -1
    if k < 198:
        return 63 + d + s_ + (k - 135) if flags & 4 else -1
    return -1


def _render_with_grad(net, mode, physical_particles, ro, rays, use_disp, perturb, noise_std, white_background):
    params = list(net.nerf_coarse.ordered_params()) + list(net.nerf_fine.ordered_params())
    outs = _RenderFunction.apply(net, (mode, ro, rays, use_disp, perturb, noise_std, white_background), physical_particles, *params)
    return dict(zip(net._train_keys, outs))


Renderer = RenderNet      # BASELINE.json's wording
