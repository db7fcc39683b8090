"""CPU oracle for hot path 2: the Lagrangian transition model.  TEST INFRASTRUCTURE ONLY.

Functional fp32 restatement of the reference's `ParticleNet.forward` (models/transmodel.py:100-163)
over a plain state dict, built on the compiled ContinuousConv restatement
(oracle/csrc/nf_oracle.c :: nfo_cconv_forward).  Pinned by tests/test_oracle.py against
tests/golden/transition_*.npz, produced by oracle/make_golden.py from the reference's own
unmodified models/transmodel.py (Open3D shimmed -> "parity unpinned" for the conv operator itself).
"""
from __future__ import annotations

import numpy as np
import torch

from . import third_party_ops as tpo

LAYER_CHANNELS = [32, 64, 64, 3]          # models/transmodel.py:26


def filter_extent(radius_scale=1.5, particle_radius=0.025) -> float:
    return float(np.float32(6 * radius_scale * particle_radius))      # models/transmodel.py:35


def _conv(sd, name, feats, in_pos, out_pos, extent, nthreads=0):
    return tpo.cconv_forward(feats, in_pos, out_pos, extent, sd[f"{name}.kernel"], sd[f"{name}.bias"],
                             sd[f"{name}.offset"], ignore_same_pos=True, use_window=True, nthreads=nthreads)


def _dense(sd, name, x):
    return torch.nn.functional.linear(x, sd[f"{name}.weight"], sd[f"{name}.bias"])


@torch.no_grad()
def particle_step(sd, pos, vel, box, box_feats, timestep=1 / 50, radius_scale=1.5, particle_radius=0.025,
                  nthreads=0, debug=False):
    """One `ParticleNet.forward`: returns (pos_out, vel_out, num_fluid_neighbors[, intermediates])."""
    pos, vel, box, box_feats = (t.float().cpu() for t in (pos, vel, box, box_feats))
    dt = timestep
    g = sd["gravity"].float()
    # models/transmodel.py:100-104
    vel_new = vel + g * dt
    pos_new = pos + (vel + vel_new) / 2 * dt
    # models/transmodel.py:106-142
    extent = filter_extent(radius_scale, particle_radius)
    fluid_feats = torch.cat([torch.ones_like(pos_new[:, 0:1]), vel_new], -1)
    c0f, counts = _conv(sd, "conv0_fluid", fluid_feats, pos_new, pos_new, extent, nthreads)
    d0 = _dense(sd, "dense0_fluid", fluid_feats)
    c0o, _ = _conv(sd, "conv0_obstacle", box_feats, box, pos_new, extent, nthreads)
    feats = torch.cat([c0o, c0f, d0], -1)
    ans = [feats]
    for i in range(1, len(LAYER_CHANNELS)):
        x = torch.relu(ans[-1])
        c, _ = _conv(sd, f"conv{i}", x, pos_new, pos_new, extent, nthreads)
        d = _dense(sd, f"dense{i}", x)
        ans.append(c + d + ans[-1] if d.shape[-1] == ans[-1].shape[-1] else c + d)
    delta = ans[-1] * (1.0 / 128)
    # models/transmodel.py:144-148
    pos_out = pos_new + delta
    vel_out = (pos_out - pos) / dt
    nnbr = counts.to(torch.float32)
    if debug:
        return pos_out, vel_out, nnbr, dict(pos_new=pos_new, vel_new=vel_new, feats=ans)
    return pos_out, vel_out, nnbr


def particle_step_grad(sd, pos, vel, box, box_feats, timestep=1 / 50, radius_scale=1.5, particle_radius=0.025, quant=None):
    """The same step with autograd recording (pure-torch ContinuousConv, oracle/third_party_ops.py::cconv_forward_torch):
    gradients flow to `sd`'s tensors and to pos / vel exactly as through the reference -- Open3D's ContinuousConv has
    gradients w.r.t. filter and input features only, positions get theirs through pos_new + delta and
    vel = (pos_out - pos) / dt (models/transmodel.py:144-148).  `quant`: straight-through rounding applied to the operands
    of the layers the CUDA path runs with 16-bit operands (conv1/dense1, conv2/dense2: inputs and weights; conv3/dense3:
    inputs only), so that both evaluations sit on the same side of every ReLU kink."""
    q = quant or (lambda t: t)
    window = lambda r: torch.clamp((1 - r) ** 3, 0, 1)                     # models/transmodel.py:73-77
    with torch.enable_grad():
        dt = timestep
        g = sd["gravity"].float()
        vel_new = vel + g * dt
        pos_new = pos + (vel + vel_new) / 2 * dt
        extent = filter_extent(radius_scale, particle_radius)
        pn = pos_new.detach()
        nns_ff = tpo.radius_search(pn, pn, 0.5 * extent, True)
        nns_fb = tpo.radius_search(box, pn, 0.5 * extent, True)

        def conv(name, feats, in_pos, nns, qw=False):
            k = sd[f"{name}.kernel"]
            return tpo.cconv_forward_torch(feats, in_pos, pn, extent, q(k) if qw else k, sd[f"{name}.bias"], sd[f"{name}.offset"],
                                           True, window, nns)[0]

        fluid_feats = torch.cat([torch.ones_like(pos_new[:, 0:1]), vel_new], -1)
        c0f = conv("conv0_fluid", fluid_feats, pn, nns_ff)
        d0 = _dense(sd, "dense0_fluid", fluid_feats)
        c0o = conv("conv0_obstacle", box_feats, box, nns_fb)
        ans = [torch.cat([c0o, c0f, d0], -1)]
        for i in range(1, len(LAYER_CHANNELS)):
            x = q(torch.relu(ans[-1]))
            tc = i < 3                                                     # conv3 / dense3 run in fp32 on fp16-stored inputs
            c = conv(f"conv{i}", x, pn, nns_ff, qw=tc)
            w = sd[f"dense{i}.weight"]
            d = torch.nn.functional.linear(x, q(w) if tc else w, sd[f"dense{i}.bias"])
            ans.append(c + d + ans[-1] if d.shape[-1] == ans[-1].shape[-1] else c + d)
        delta = ans[-1] * (1.0 / 128)
        pos_out = pos_new + delta
        vel_out = (pos_out - pos) / dt
        counts = (nns_ff[1][1:] - nns_ff[1][:-1]).to(torch.float32)
        return pos_out, vel_out, counts
