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
