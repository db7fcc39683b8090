"""ParticleNet -- drop-in for the reference's `models/transmodel.py:ParticleNet` on B200.

Same constructor signature, same `forward(pos, vel, box, box_feats, feats=None,
fixed_radius_search_hash_table=None) -> (pos, vel, num_fluid_neighbors)`, same state-dict keys
(`conv0_fluid.kernel/bias/offset`, `dense0_fluid.weight/bias`, ..., `gravity`).  The arithmetic runs in
libnf_b200.so (csrc/nf_cconv.cu, nf_grid.cu): fixed-radius neighbour lists on the shared spatial grid,
layer 0 in fp32, layers 1-3 as fused gather + trilinear-scatter + tcgen05 GEMM kernels.
No CPU or eager fallback.  Under autograd (pos / vel / parameters requiring grad) a step is one differentiable node whose
backward runs nf_transition_backward (feature and filter gradients of the five ContinuousConvs, as Open3D's).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import NFError, check, lib, ptr, require_cuda, stream_ptr


class ContinuousConvParams(nn.Module):
    """Parameter container with open3d.ml.torch.layers.ContinuousConv's state layout:
    `kernel` (*kernel_size, in_channels, filters) ~ U(-0.05, 0.05), `bias` (filters) zeros, buffer `offset` (3)."""

    def __init__(self, in_channels, filters, kernel_size):
        super().__init__()
        self.in_channels, self.filters, self.kernel_size = in_channels, filters, list(kernel_size)
        self.kernel = nn.Parameter(torch.empty(*kernel_size, in_channels, filters).uniform_(-0.05, 0.05))
        self.bias = nn.Parameter(torch.zeros(filters))
        self.register_buffer("offset", torch.zeros(3))

    def forward(self, *a, **k):
        raise NFError("ContinuousConv layers run fused inside ParticleNet.forward (csrc/nf_cconv.cu)")


class ParticleNet(nn.Module):
    def __init__(self, kernel_size=[4, 4, 4], radius_scale=1.5, coordinate_mapping="ball_to_cube_volume_preserving",
                 interpolation="linear", use_window=True, particle_radius=0.025, timestep=1 / 50,
                 gravity=(0, -9.81, 0), other_feats_channels=0, operand_dtype: str = "fp16"):
        super().__init__()
        if list(kernel_size) != [4, 4, 4] or coordinate_mapping != "ball_to_cube_volume_preserving" \
                or interpolation != "linear" or not use_window or other_feats_channels != 0:
            raise NFError("the sm_100a kernels implement the configuration every reference entry point uses: "
                          "4x4x4 linear filters, ball_to_cube_volume_preserving, poly6 window, no extra features")
        self.layer_channels = [32, 64, 64, 3]
        self.kernel_size, self.radius_scale, self.particle_radius = kernel_size, radius_scale, particle_radius
        self.coordinate_mapping, self.interpolation, self.use_window = coordinate_mapping, interpolation, use_window
        self.filter_extent = np.float32(6 * radius_scale * particle_radius)       # models/transmodel.py:35
        self.time_step = timestep
        self.register_buffer("gravity", torch.FloatTensor(gravity))
        self.conv0_fluid = ContinuousConvParams(4, 32, kernel_size)
        self.conv0_obstacle = ContinuousConvParams(3, 32, kernel_size)
        self.dense0_fluid = nn.Linear(4, 32)
        nn.init.xavier_uniform_(self.dense0_fluid.weight)
        nn.init.zeros_(self.dense0_fluid.bias)
        for i, (cin, cout) in enumerate([(96, 64), (64, 64), (64, 3)], start=1):
            setattr(self, f"dense{i}", nn.Linear(cin, cout))
            setattr(self, f"conv{i}", ContinuousConvParams(cin, cout, kernel_size))
        self.operand_dtype = {"fp16": _lib.NF_DTYPE_F16, "bf16": _lib.NF_DTYPE_BF16}[operand_dtype]
        self._packed = None
        self._ws = None
        self.num_fluid_neighbors = None
        self.pos_correction = None
        self._overflow = None      # device int32[2]: particles whose fluid / box neighbour list was truncated
        self._ovf_host = None      # pinned int32[2] + event: the counter is copied out after every step and looked at
        self._ovf_event = None     # (without blocking) at the start of the next one

    # ------------------------------------------------------------------ internals
    def ordered_params(self):
        """nf_transition_pack_weights order."""
        mods = [self.conv0_fluid, self.conv0_obstacle, self.dense0_fluid, self.conv1, self.dense1, self.conv2,
                self.dense2, self.conv3, self.dense3]
        out = []
        for m in mods:
            out += [m.kernel if isinstance(m, ContinuousConvParams) else m.weight, m.bias]
        return out

    def _packed_weights(self):
        params = self.ordered_params()
        convs = (self.conv0_fluid, self.conv0_obstacle, self.conv1, self.conv2, self.conv3)
        key = tuple((p.data_ptr(), p._version) for p in params) + tuple(m.offset._version for m in convs) + \
            (self.operand_dtype,)
        if self._packed is None or self._packed[0] != key:
            for m in convs:      # checked when (re)packing only: a device sync per step would dominate small scenes
                if bool((m.offset != 0).any()):
                    raise NFError("non-zero ContinuousConv.offset is not supported (the reference never sets it)")
            ps = [p.detach().to(torch.float32).contiguous() for p in params]
            require_cuda(*ps)
            out = torch.empty(lib().nf_transition_packed_weights_bytes(), dtype=torch.uint8, device=ps[0].device)
            arr = (C.c_void_p * 18)(*[p.data_ptr() for p in ps])
            check(lib().nf_transition_pack_weights(arr, self.operand_dtype, ptr(out), stream_ptr()),
                  "nf_transition_pack_weights")
            out._keepalive = ps
            self._packed = (key, out)
        return self._packed[1]

    def _packed_weights_bwd(self):
        """Flipped / transposed filters for the feature-gradient convs (nf_transition_pack_weights_bwd), cached by version."""
        params = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = getattr(self, "_packed_bwd", None)
        if hit is None or hit[0] != key:
            ps = [p.detach().to(torch.float32).contiguous() for p in params]
            require_cuda(*ps)
            out = torch.empty(lib().nf_transition_packed_weights_bwd_bytes(), dtype=torch.uint8, device=ps[0].device)
            arr = (C.c_void_p * 18)(*[p.data_ptr() for p in ps])
            check(lib().nf_transition_pack_weights_bwd(arr, ptr(out), stream_ptr()), "nf_transition_pack_weights_bwd")
            out._keepalive = ps
            self._packed_bwd = (key, out)
        return self._packed_bwd[1]

    def _workspace(self, n, m, device):
        need = lib().nf_transition_workspace_bytes(n, m)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _args(self, pos, vel, box, box_feats, outs, ws, phase=-1, shard=(0, 0), debug=None):
        a = _lib.TransitionArgs()
        a.pos, a.vel, a.n_fluid = ptr(pos), ptr(vel), pos.shape[0]
        a.box, a.box_normals, a.n_box = ptr(box), ptr(box_feats), box.shape[0]
        a.gravity = (C.c_float * 3)(*[float(v) for v in self._gravity_host])
        a.dt, a.filter_extent, a.dtype = float(self.time_step), float(self.filter_extent), self.operand_dtype
        a.weights = ptr(self._packed_weights())
        a.pos_out, a.vel_out, a.nnbr_out = ptr(outs[0]), ptr(outs[1]), ptr(outs[2])
        a.feats0_out = ptr(debug["feats0"]) if debug else None
        a.delta_out = ptr(outs[3])
        a.workspace, a.workspace_bytes = ptr(ws), ws.numel()
        a.shard_begin, a.shard_end, a.phase = int(shard[0]), int(shard[1]), int(phase)
        a.box_grid_ws = ptr(self._box_grid(box).ws)
        if self._overflow is None or self._overflow.device != pos.device:
            self._overflow = torch.zeros(2, dtype=torch.int32, device=pos.device)
        a.overflow_out = ptr(self._overflow)
        return a

    def _box_grid(self, box):
        """Cell-sorted grid of the container points, rebuilt only when the tensor changes (the container is static in
        every reference scene; Open3D offers the same reuse through `fixed_radius_search_hash_table`)."""
        from .ops import CELL_SCALE, Grid
        key = (box.data_ptr(), box._version, tuple(box.shape), str(box.device), float(self.filter_extent))
        hit = getattr(self, "_box_grid_cache", None)
        if hit is None or hit[0] != key:
            import numpy as np
            cell = float(np.float32(CELL_SCALE) * (np.float32(0.5) * np.float32(self.filter_extent)))   # as nf_transition_step
            self._box_grid_cache = (key, Grid(box, cell), box)
        return self._box_grid_cache[1]

    def _raise_overflow(self, nf, nb):
        self._overflow.zero_()      # one report per incident: later rollouts start clean
        raise NFError(f"neighbour lists truncated at 128 entries for {nf} (fluid) / {nb} (box) particle-steps: "
                      "results differ from the reference (Open3D's FixedRadiusSearch has no cap); the scene is denser "
                      "than the kernels support")

    def _poll_overflow(self):
        """Non-blocking look at the overflow counter copied out after the previous step: a truncated neighbour list
        surfaces as an NFError on the next `forward` at the latest, without a device synchronisation per step."""
        ev = self._ovf_event
        if ev is not None and ev.query():
            self._ovf_event = None
            nf, nb = self._ovf_host.tolist()
            if nf or nb:
                self._raise_overflow(nf, nb)

    def _post_overflow(self):
        if self._ovf_host is None:
            self._ovf_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        if self._ovf_event is None:           # the previous copy has been consumed: the pinned buffer is free again
            self._ovf_host.copy_(self._overflow, non_blocking=True)
            self._ovf_event = torch.cuda.Event()
            self._ovf_event.record()

    def check_neighbor_overflow(self, group=None):
        """The kernels keep at most 128 fluid and 128 box neighbours per particle (the reference has no cap).  Returns
        silently while no list was truncated since the last check, raises otherwise (and clears the counter).
        Synchronises.  With torch.distributed initialised every rank of `group` sees the sum over ranks, so a sharded
        rollout fails on all ranks together instead of leaving the others inside the next collective."""
        if self._overflow is None:
            return
        import torch.distributed as dist
        cnt = self._overflow.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(cnt, group=group)
        self._ovf_event = None
        nf, nb = cnt.tolist()
        if nf or nb:
            self._raise_overflow(nf, nb)

    def _prepare(self, pos, vel, box, box_feats, feats):
        if feats is not None:
            raise NFError("other_feats_channels > 0 is not used by any reference entry point and is not implemented")
        require_cuda(pos, vel, box, box_feats)
        f = lambda t: t.detach().to(torch.float32).contiguous()
        pos, vel, box, box_feats = f(pos), f(vel), f(box), f(box_feats)
        if not hasattr(self, "_gravity_host") or self._gravity_version != self.gravity._version:
            self._gravity_host = self.gravity.detach().cpu().tolist()
            self._gravity_version = self.gravity._version
        n = pos.shape[0]
        dev = pos.device
        outs = (torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev), torch.empty((n,), device=dev),
                torch.empty((n, 3), device=dev))
        return pos, vel, box, box_feats, outs, self._workspace(n, box.shape[0], dev)

    # ------------------------------------------------------------------ reference API
    def forward(self, pos, vel, box, box_feats, feats=None, fixed_radius_search_hash_table=None, debug=None):
        """models/transmodel.py:151-163.  `fixed_radius_search_hash_table` is accepted and ignored, as upstream.

        Neighbour lists hold at most 128 fluid and 128 box neighbours per particle (Open3D has no cap; a 0.05-spaced
        fluid has ~42).  If a list is ever truncated, `num_fluid_neighbors` still reports the true count, and an NFError
        is raised by the next `forward` (non-blocking poll) or by `check_neighbor_overflow()` (blocking)."""
        self._poll_overflow()
        if debug is None and torch.is_grad_enabled() and (
                any(isinstance(t, torch.Tensor) and t.requires_grad for t in (pos, vel)) or any(p.requires_grad for p in self.parameters())):
            return _TransitionFunction.apply(self, (box, box_feats, feats), pos, vel, *self.ordered_params())
        return self._forward_impl(pos, vel, box, box_feats, feats, debug)

    def _forward_impl(self, pos, vel, box, box_feats, feats=None, debug=None, train=False):
        pos, vel, box, box_feats, outs, ws = self._prepare(pos, vel, box, box_feats, feats)
        if train:      # the backward pass reads this step's neighbour lists and activations: a workspace of its own
            ws = torch.empty(max(lib().nf_transition_workspace_bytes(pos.shape[0], box.shape[0]), 256), dtype=torch.uint8,
                             device=pos.device)
        if debug is not None:
            debug["feats0"] = torch.empty((pos.shape[0], 96), device=pos.device)
        a = self._args(pos, vel, box, box_feats, outs, ws, debug=debug)
        check(lib().nf_transition_step(C.byref(a), stream_ptr()), "nf_transition_step")
        self._post_overflow()
        self._last_ws = (ws, pos.shape[0], box.shape[0])
        self.num_fluid_neighbors, self.pos_correction = outs[2], outs[3]
        self._keep = (pos, vel, box, box_feats)
        if train:
            # the outputs themselves stay out of `keep`: they become the autograd node's outputs, and a node that holds
            # its own outputs is a reference cycle through C++ that frees this step's workspace only at a gc pass
            return (outs[0], outs[1], outs[2]), dict(args=a, ws=ws, keep=(pos, vel, box, box_feats, self._packed_weights(),
                                                                           self._box_grid(box)), n=pos.shape[0])
        return outs[0], outs[1], outs[2]

    step = forward      # BASELINE.json's wording: TransModel.step

    @property
    def ans_convs(self):
        """models/transmodel.py:122-131: the pre-activation outputs of the four layers of the LAST step, fp32
        `[(N,96), (N,64), (N,64), (N,3)]` (row i = particle i).  Read-only views into the step's workspace: valid until the
        next `forward` of this module (a training forward's views stay valid while its graph is alive)."""
        last = getattr(self, "_last_ws", None)
        if last is None:
            raise NFError("ans_convs: no step has run yet")
        ws, n, m = last
        out = []
        for layer, width in ((3, 96), (4, 64), (5, 64), (6, 3)):
            off, rb = C.c_size_t(), C.c_size_t()
            check(lib().nf_transition_layer_buffer(n, m, layer, C.byref(off), C.byref(rb)), "nf_transition_layer_buffer")
            rows = ws[off.value: off.value + n * rb.value].view(torch.float32).view(n, rb.value // 4)
            out.append(rows[:, :width])
        return out


class _TransitionFunction(torch.autograd.Function):
    """Autograd node of one ParticleNet step: forward = nf_transition_step on a workspace of its own, backward =
    nf_transition_backward (csrc/nf_cconv_bwd.cu).  Differentiable inputs: pos, vel and the 18 parameter tensors."""

    @staticmethod
    def forward(ctx, net, static, pos, vel, *params):
        box, box_feats, feats = static
        with torch.no_grad():
            outs, saved = net._forward_impl(pos, vel, box, box_feats, feats, None, train=True)
        ctx.net, ctx.saved = net, saved
        ctx.mark_non_differentiable(outs[2])
        return outs

    @staticmethod
    def backward(ctx, g_pos, g_vel, _g_nn):
        net, sv = ctx.net, ctx.saved
        n, dev = sv["n"], sv["ws"].device
        d_pos = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        d_vel = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        flat = torch.zeros(int(lib().nf_transition_param_count()), dtype=torch.float32, device=dev)
        if n > 0:
            f = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
            g_pos, g_vel = f(g_pos), f(g_vel)
            bws = torch.empty(max(lib().nf_transition_backward_workspace_bytes(n), 256), dtype=torch.uint8, device=dev)
            b = _lib.TransitionBwdArgs()
            b.fwd = C.pointer(sv["args"])
            b.weights_bwd = ptr(net._packed_weights_bwd())
            b.g_pos_out, b.g_vel_out = ptr(g_pos), ptr(g_vel)
            b.d_pos, b.d_vel, b.d_params = ptr(d_pos), ptr(d_vel), ptr(flat)
            b.workspace, b.workspace_bytes = ptr(bws), bws.numel()
            check(lib().nf_transition_backward(C.byref(b), stream_ptr()), "nf_transition_backward")
        grads, o = [], 0
        for p in net.ordered_params():
            grads.append(flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        need = ctx.needs_input_grad
        return (None, None, d_pos if need[2] else None, d_vel if need[3] else None) + \
            tuple(gp if need[4 + i] else None for i, gp in enumerate(grads))


TransModel = ParticleNet
