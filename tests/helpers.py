"""Shared test helpers: rebuild golden-case inputs from their stored seeds."""
import os

import numpy as np
import torch

from neurofluid_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RENDER_CASES = ["small_boost", "small_default", "small_nomask", "cfg0_sub", "small_he"]
ABLATION_CASES = ["small_wo_sdir", "small_min_enc", "small_incl_ray"]   # encoding switches (models/renderer.py:100-109,152-175)
TRANSITION_CASES = ["small", "medium"]


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float(torch.norm(a - b) / torch.norm(b).clamp_min(1e-30))


def load_render_case(name):
    g = np.load(os.path.join(GOLDEN, f"render_{name}.npz"))
    enc = {}
    if "enc" in g.files:
        enc = dict(zip(("density", "smoothed_pos", "var", "smoothed_dir", "exclude_ray"), [bool(v) for v in g["enc"]]))
    cfg = scenes.render_cfg(use_mask=bool(g["use_mask"]), **enc)
    e = cfg.encoding
    in_xyz = 63 * (1 + bool(e.smoothed_pos) + bool(e.var)) + (9 if e.density else 0)
    in_dir = 27 * (1 + bool(e.smoothed_dir))
    sd = scenes.init_render_state(int(g["seed"]), float(g["sigma_boost"]), in_xyz=in_xyz, in_dir=in_dir,
                                  weight_gain=float(g["weight_gain"]))
    particles = torch.from_numpy(scenes.lattice_particles(int(g["n_lat"]), int(g["seed"]), center=tuple(g["center"])))
    rays = torch.from_numpy(g["rays"])
    cw = torch.from_numpy(scenes.CAMERA_C2W)
    return dict(g=g, cfg=cfg, sd=sd, particles=particles, rays=rays, ro=cw[:, 3].clone(), cw=cw)


def load_transition_case(name):
    g = np.load(os.path.join(GOLDEN, f"transition_{name}.npz"))
    n, seed = int(g["n_lat"]), int(g["seed"])
    sd = scenes.init_particle_state(seed)
    half = (n - 1) / 2 * 0.05
    pos = torch.from_numpy(scenes.lattice_particles(n, seed, center=(0.0, 0.0, -1 + 0.03 + half)))
    bp, bn = scenes.box_points(float(g["box_spacing"]))
    return dict(g=g, sd=sd, pos=pos, vel=torch.zeros_like(pos), box=torch.from_numpy(bp), box_n=torch.from_numpy(bn))
