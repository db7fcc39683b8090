"""ctypes binding of libnf_b200.so (C ABI declared in include/nf_b200.h).

The product path has no fallback: if the library is missing or fails to load, every entry point
raises.  Build it with `python -m neurofluid_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NF_B200_LIB") or os.path.join(_HERE, "libnf_b200.so")     # NF_B200_LIB: the -DNF_TUNING build

NF_DTYPE_F16, NF_DTYPE_BF16 = 0, 1
NF_RENDER_FORWARD, NF_RENDER_COARSE, NF_RENDER_FINE = 0, 1, 2
NF_SEARCH_AUTO, NF_SEARCH_STREAM, NF_SEARCH_SWEEP = 0, 1, 2

_vp, _i32, _f32, _sz, _i64 = C.c_void_p, C.c_int32, C.c_float, C.c_size_t, C.c_int64


class RenderArgs(C.Structure):
    _fields_ = [
        ("grid_ws", _vp), ("particles", _vp), ("n_particles", _i32),
        ("rays", _vp), ("n_rays", _i32), ("ro", _f32 * 3), ("ro_dev", _vp),
        ("z_coarse", _vp), ("u_importance", _vp), ("n_coarse", _i32), ("n_importance", _i32),
        ("radius", _f32), ("K", _i32), ("search", _i32),
        ("mode", _i32), ("use_mask", _i32), ("white_background", _i32), ("dtype", _i32),
        ("weights_coarse", _vp), ("weights_fine", _vp),
        ("rgb0", _vp), ("depth0", _vp), ("opacity0", _vp), ("num_nn0", _vp), ("mask0", _vp),
        ("rgb1", _vp), ("depth1", _vp), ("opacity1", _vp), ("num_nn1", _vp), ("mask1", _vp),
        ("workspace", _vp), ("workspace_bytes", _sz), ("stats", _vp), ("flags", _i32),
        ("z_stride", _i32), ("u_stride", _i32), ("noise0", _vp), ("noise1", _vp),
        ("include_ray", _i32), ("same_smooth_factor", _i32),
    ]


NF_RENDER_SAVE_NEIGHBORS = 1
NF_PHASE_SHARDED = -2
NF_E_UNSUPPORTED = -4
NF_COMM_ID_BYTES = 128


class RenderWsView(C.Structure):
    _fields_ = [(n, _sz) for n in ("counters", "act0", "act1", "z1", "rec0", "rowid0", "out0", "rec1", "rowid1", "out1",
                                   "nbr0", "nbr1", "miss", "total")] + \
               [(n, _i32) for n in ("act_stride0", "act_stride1", "cap0", "cap1")]


class TransitionArgs(C.Structure):
    _fields_ = [
        ("pos", _vp), ("vel", _vp), ("n_fluid", _i32),
        ("box", _vp), ("box_normals", _vp), ("n_box", _i32),
        ("gravity", _f32 * 3), ("dt", _f32), ("filter_extent", _f32), ("dtype", _i32),
        ("weights", _vp),
        ("pos_out", _vp), ("vel_out", _vp), ("nnbr_out", _vp),
        ("feats0_out", _vp), ("delta_out", _vp),
        ("workspace", _vp), ("workspace_bytes", _sz),
        ("shard_begin", _i32), ("shard_end", _i32), ("box_grid_ws", _vp), ("overflow_out", _vp), ("phase", _i32),
    ]


class RenderBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", C.POINTER(RenderArgs)), ("weights_coarse_bwd", _vp), ("weights_fine_bwd", _vp),
        ("d_rgb0", _vp), ("d_depth0", _vp), ("d_opacity0", _vp), ("d_rgb1", _vp), ("d_depth1", _vp), ("d_opacity1", _vp),
        ("d_particles", _vp), ("d_params_coarse", _vp), ("d_params_fine", _vp),
        ("workspace", _vp), ("workspace_bytes", _sz),
    ]


class TransitionBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", C.POINTER(TransitionArgs)), ("weights_bwd", _vp), ("g_pos_out", _vp), ("g_vel_out", _vp),
        ("d_pos", _vp), ("d_vel", _vp), ("d_params", _vp), ("workspace", _vp), ("workspace_bytes", _sz),
    ]


class CConvArgs(C.Structure):
    _fields_ = [
        ("grid_in", _vp), ("in_feat", _vp), ("n_in", _i32), ("cin", _i32),
        ("out_pos", _vp), ("n_out", _i32), ("cout", _i32),
        ("extent", _f32), ("use_window", _i32), ("ignore_same", _i32), ("dtype", _i32),
        ("weights", _vp), ("out", _vp), ("count_out", _vp), ("nbr_index_out", _vp), ("overflow_out", _vp),
        ("workspace", _vp), ("workspace_bytes", _sz),
    ]


# name -> (restype, argtypes); every symbol include/nf_b200.h declares
SIGNATURES = {
    "nf_version": (C.c_int, []),
    "nf_last_error": (C.c_char_p, []),
    "nf_launch_count": (_i64, []),
    "nf_profile_enable": (C.c_int, [C.c_int]),
    "nf_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "nf_grid_workspace_bytes": (_sz, [C.c_int]),
    "nf_grid_build": (C.c_int, [_vp, C.c_int, _f32, _vp, _sz, _vp]),
    "nf_ballquery_firstk": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _f32, C.c_int, _vp, _vp, _vp]),
    "nf_render_packed_weights_bytes": (_sz, []),
    "nf_render_pack_weights": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _vp]),
    "nf_render_pack_weights_ex": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, _vp]),
    "nf_nerf_mlp_workspace_bytes": (C.c_size_t, []),
    "nf_nerf_mlp_forward": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, C.c_size_t, _vp]),
    "nf_render_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int]),
    "nf_render_workspace_bytes_ex": (_sz, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "nf_render_workspace_view": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RenderWsView)]),
    "nf_render_forward": (C.c_int, [C.POINTER(RenderArgs), _vp]),
    "nf_render_param_count": (_sz, []),
    "nf_render_packed_weights_bwd_bytes": (_sz, []),
    "nf_render_pack_weights_bwd": (C.c_int, [C.POINTER(_vp), _vp, _vp]),
    "nf_render_pack_weights_bwd_ex": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _vp]),
    "nf_nerf_mlp_backward_workspace_bytes": (_sz, [C.c_int]),
    "nf_nerf_mlp_backward": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "nf_render_backward_workspace_bytes": (_sz, [C.c_int] * 5),
    "nf_render_backward": (C.c_int, [C.POINTER(RenderBwdArgs), _vp]),
    "nf_transition_packed_weights_bytes": (_sz, []),
    "nf_transition_pack_weights": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _vp]),
    "nf_transition_workspace_bytes": (_sz, [C.c_int, C.c_int]),
    "nf_transition_num_phases": (C.c_int, []),
    "nf_transition_step": (C.c_int, [C.POINTER(TransitionArgs), _vp]),
    "nf_transition_layer_buffer": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_sz), C.POINTER(_sz)]),
    "nf_transition_param_count": (_sz, []),
    "nf_transition_packed_weights_bwd_bytes": (_sz, []),
    "nf_transition_pack_weights_bwd": (C.c_int, [C.POINTER(_vp), _vp, _vp]),
    "nf_transition_backward_workspace_bytes": (_sz, [C.c_int]),
    "nf_transition_backward": (C.c_int, [C.POINTER(TransitionBwdArgs), _vp]),
    "nf_comm_unique_id": (C.c_int, [_vp]),
    "nf_comm_init": (C.c_int, [_vp, C.c_int, C.c_int]),
    "nf_comm_finalize": (C.c_int, []),
    "nf_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nf_comm_register_buffer": (C.c_int, [_vp, C.c_size_t]),
    "nf_comm_exchange_timeouts": (C.c_int, [C.POINTER(C.c_uint)]),
    "nf_allgather_rows": (C.c_int, [_vp, _sz, _vp]),
    "nf_cconv_packed_weights_bytes": (_sz, [C.c_int, C.c_int]),
    "nf_cconv_pack_weights": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "nf_cconv_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "nf_cconv_forward": (C.c_int, [C.POINTER(CConvArgs), _vp]),
    "nf_generate_rays": (C.c_int, [C.c_int, C.c_int, _f32, _vp, _vp, _vp]),
    "nf_nearest_distance": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "nf_pair_distance": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "nf_sqdiff_sum": (C.c_int, [_vp, _vp, C.c_longlong, _vp, _vp]),
}

_lib = None


class NFError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NFError(f"{LIB_PATH} is missing: the CUDA extension is not built "
                          "(run `python -m neurofluid_b200.build`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().nf_last_error()
        raise NFError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """device (or host) pointer of a torch tensor as c_void_p; None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NFError("neurofluid_b200 runs on CUDA tensors only (sm_100a kernels); got a CPU tensor and there "
                          "is no CPU fallback")


_forward_only_fn = None


def _forward_only_function():
    global _forward_only_fn
    if _forward_only_fn is None:
        import torch

        class ForwardOnly(torch.autograd.Function):
            @staticmethod
            def forward(ctx, anchor, *outs):
                return tuple(o.view_as(o) for o in outs)

            @staticmethod
            def backward(ctx, *grads):
                raise NFError("backward through the sm_100a kernels is not implemented yet (forward / evaluation only; "
                              "DESIGN.md section 8): run under torch.no_grad() or detach the result")

        _forward_only_fn = ForwardOnly
    return _forward_only_fn


def forward_only(anchors, outputs):
    """The kernels are forward-only this round.  When autograd is recording and something that
    feeds the call requires grad, the floating-point outputs are tied to it through a node whose backward raises --
    so `loss.backward()` fails with that message instead of an unrelated autograd error or, worse, no gradient."""
    import torch
    live = [t for t in anchors if isinstance(t, torch.Tensor) and t.requires_grad]
    if not (torch.is_grad_enabled() and live):
        return outputs
    if isinstance(outputs, dict):
        keys = [k for k, v in outputs.items() if v.is_floating_point()]
        tied = _forward_only_function().apply(live[0], *[outputs[k] for k in keys])
        out = dict(outputs)
        out.update(zip(keys, tied))
        return out
    return tuple(_forward_only_function().apply(live[0], *outputs))
