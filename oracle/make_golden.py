"""Generate tests/golden/*.npz by running the REFERENCE'S OWN, UNMODIFIED Python
(/root/reference/models/{renderer,nerf,transmodel}.py, utils/ray_utils.py) on seeded synthetic
inputs, with only the three un-installable third-party entry points (pytorch3d.ops.ball_query,
kornia.create_meshgrid, open3d.ml.torch ContinuousConv/reduce_subarrays_sum) routed to
oracle/third_party_ops.py through oracle/shims/.

Runs only in the authoring container (needs /root/reference).  The fixtures it writes are what
pins oracle/renderer.py and oracle/transition.py (tests/test_oracle.py) and what the GPU parity
tests compare against on the GPU box, where /root/reference does not exist.

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NF_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from models.renderer import RenderNet          # noqa: E402  (reference source, unmodified)
    from models.transmodel import ParticleNet      # noqa: E402
    from utils import ray_utils                    # noqa: E402
    return RenderNet, ParticleNet, ray_utils


def render_case(RenderNet, ray_utils, name, n_lat, H, crop, seed, sigma_boost, use_mask=True, rays_stride=1,
                center=(0.0, 0.0, 0.0), weight_gain=1.0, enc=None):
    from neurofluid_b200 import scenes
    enc = enc or {}
    cfg = scenes.render_cfg(use_mask=use_mask, **enc)
    net = RenderNet(cfg, scenes.NEAR, scenes.FAR)
    e = cfg.encoding
    in_xyz = 63 * (1 + bool(e.smoothed_pos) + bool(e.var)) + (9 if e.density else 0)
    in_dir = 27 * (1 + bool(e.smoothed_dir))
    sd = scenes.init_render_state(seed, sigma_boost, in_xyz=in_xyz, in_dir=in_dir, weight_gain=weight_gain)
    net.load_state_dict(sd, strict=True)
    particles = torch.from_numpy(scenes.lattice_particles(n_lat, seed, center=center))
    # rays through the reference's own get_ray_directions/get_rays
    import math
    focal = 0.5 * H / math.tan(0.5 * scenes.CAMERA_ANGLE_X)
    dirs = ray_utils.get_ray_directions(H, H, focal)
    cw = torch.from_numpy(scenes.CAMERA_C2W)
    ro_, rd_ = ray_utils.get_rays(dirs, cw)
    rays_full = torch.cat([ro_, rd_], -1).view(-1, 6)
    mine, focal2, _ = scenes.camera_rays(H, H)
    assert torch.equal(mine, rays_full), "scenes.camera_rays must equal the reference's get_rays bit for bit"
    rays = scenes.center_crop_rays(rays_full, H, H, crop)[::rays_stride].contiguous()
    ro = net.set_ro(cw)
    with torch.no_grad():
        full = net(particles, ro, rays, focal, cw)
        coarse = net.coarse_rendering(particles, ro, rays, focal, cw)
        # NOTE: the reference's fine_rendering raises on every shipped config (UnboundLocalError at
        # models/renderer.py:175 when encoding.smoothed_dir is on, and a 4-into-1 unpack at :322
        # otherwise), so there is nothing to pin for it; oracle.renderer 'fine' mode restates the
        # evident intent (sigma-only coarse pass) and is checked for self-consistency only.
        try:
            fine = net.fine_rendering(particles, ro, rays, focal, cw)
        except (UnboundLocalError, ValueError, TypeError):
            fine = {}
    out = {
        "n_lat": n_lat, "H": H, "crop": crop, "seed": seed, "sigma_boost": sigma_boost, "use_mask": use_mask,
        "rays_stride": rays_stride, "weight_gain": weight_gain, "center": np.asarray(center, np.float32),
        "rays": rays.numpy(),                       # stored so the GPU box needs no reference code
        "enc": np.asarray([bool(e.density), bool(e.smoothed_pos), bool(e.var), bool(e.smoothed_dir), bool(e.exclude_ray)]),
    }
    for k, v in full.items():
        out[f"forward.{k}"] = v.numpy()
    for k, v in coarse.items():
        out[f"coarse.{k}"] = v.numpy()
    for k, v in fine.items():
        out[f"fine.{k}"] = v.numpy()
    for k in list(out):
        if k.endswith("num_nn_0") or k.endswith("num_nn_1"):
            out[k] = out[k].astype(np.int8)
    np.savez_compressed(os.path.join(GOLD, f"render_{name}.npz"), **out)
    act0 = float((full["num_nn_0"] == 20).float().mean())
    print(f"render_{name}: R={rays.shape[0]} P={particles.shape[0]} active0={act0:.3f} "
          f"rgb1 mean={float(full['rgb1'].mean()):.4f} min={float(full['rgb1'].min()):.4f}")


def render_jitter_case(RenderNet, ray_utils, name, n_lat, H, crop, seed, sigma_boost, perturb, noise_std, rays_stride=1):
    """Training-time jitter (perturb > 0: stratified depths + random inverse-CDF arguments; noise_std > 0: sigma noise) through
    the reference's own forward under torch.manual_seed(seed); the same seed then replays the four draws in the reference's
    call order, and they are stored with the outputs so that the oracle / CUDA path can be fed the identical numbers."""
    from neurofluid_b200 import scenes
    import math
    cfg = scenes.render_cfg()
    net = RenderNet(cfg, scenes.NEAR, scenes.FAR)
    sd = scenes.init_render_state(seed, sigma_boost)
    net.load_state_dict(sd, strict=True)
    particles = torch.from_numpy(scenes.lattice_particles(n_lat, seed))
    rays_full, focal, cw = scenes.camera_rays(H, H)
    rays = scenes.center_crop_rays(rays_full, H, H, crop)[::rays_stride].contiguous()
    R, S, SI = rays.shape[0], cfg.ray.N_samples, cfg.ray.N_importance
    torch.manual_seed(1000 + seed)
    with torch.no_grad():
        full = net(particles, net.set_ro(cw), rays, focal, cw, perturb=perturb, noise_std=noise_std)
    torch.manual_seed(1000 + seed)
    draws = {"z_rand": torch.rand((R, S)), "noise0": torch.randn((R, S)), "u": torch.rand([R, SI]), "noise1": torch.randn((R, S + SI))}
    out = {"n_lat": n_lat, "H": H, "crop": crop, "seed": seed, "sigma_boost": sigma_boost, "use_mask": True, "rays_stride": rays_stride,
           "weight_gain": 1.0, "center": np.zeros(3, np.float32), "rays": rays.numpy(), "perturb": perturb, "noise_std": noise_std}
    for k, v in draws.items():
        out["draw." + k] = v.numpy()
    for k, v in full.items():
        out[f"forward.{k}"] = v.numpy().astype(np.int8) if k.startswith("num_nn") else v.numpy()
    np.savez_compressed(os.path.join(GOLD, f"render_{name}.npz"), **out)
    print(f"render_{name}: R={R} perturb={perturb} noise_std={noise_std} rgb1 mean={float(full['rgb1'].mean()):.4f}")


def transition_case(ParticleNet, name, n_lat, seed, box_spacing, steps):
    from neurofluid_b200 import scenes
    net = ParticleNet(gravity=(0.0, 0.0, -9.81))
    sd = scenes.init_particle_state(seed)
    net.load_state_dict(sd, strict=True)
    half = (n_lat - 1) / 2 * 0.05
    pos = torch.from_numpy(scenes.lattice_particles(n_lat, seed, center=(0.0, 0.0, -1 + 0.03 + half)))
    vel = torch.zeros_like(pos)
    bp, bn = scenes.box_points(box_spacing)
    box, box_n = torch.from_numpy(bp), torch.from_numpy(bn)
    out = {"n_lat": n_lat, "seed": seed, "box_spacing": box_spacing, "steps": steps}
    with torch.no_grad():
        for s in range(steps):
            pos, vel, nn = net(pos, vel, box, box_n)
            out[f"pos_{s}"] = pos.numpy().copy()
            out[f"vel_{s}"] = vel.numpy().copy()
            out[f"nnbr_{s}"] = nn.numpy().astype(np.int16)
            if s == 0:
                out["feats0"] = net.ans_convs[0].numpy().copy()
                out["delta0"] = net.pos_correction.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, f"transition_{name}.npz"), **out)
    print(f"transition_{name}: N={pos.shape[0]} M={box.shape[0]} mean nbrs={float(nn.mean()):.1f} "
          f"|delta0|max={float(np.abs(out['delta0']).max()):.2e}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    RenderNet, ParticleNet, ray_utils = import_reference()
    torch.manual_seed(0)
    np.random.seed(0)
    # small: 16x16 centre crop, 9^3 particles; sigma-boosted so compositing / importance sampling matter
    render_case(RenderNet, ray_utils, "small_boost", 9, 400, 16, 0, 5.0)
    # default-init weights (sigma ~ 0): the reference's initial state
    render_case(RenderNet, ray_utils, "small_default", 9, 400, 16, 1, 0.0)
    # use_mask=False exercises the padded-slot-at-origin quirk (models/renderer.py:97-98); cube at the origin
    render_case(RenderNet, ray_utils, "small_nomask", 7, 400, 12, 2, 5.0, use_mask=False)
    # BASELINE config[0]-like geometry (18^3 particles, 64x64 crop of a 400^2 view), every 16th ray
    render_case(RenderNet, ray_utils, "cfg0_sub", 18, 400, 64, 3, 5.0, rays_stride=16)
    # He-scaled weights: O(1) activations, sigma of both signs, structured rgb -> stresses MLP numerics
    render_case(RenderNet, ray_utils, "small_he", 9, 400, 16, 4, 1.0, weight_gain=2.45)
    # encoding ablations (models/renderer.py:152-175): the reference's only published numbers are the "wo-smoothed_dir" run
    render_case(RenderNet, ray_utils, "small_wo_sdir", 9, 400, 16, 6, 5.0, enc=dict(smoothed_dir=False))
    render_case(RenderNet, ray_utils, "small_min_enc", 9, 400, 16, 7, 5.0, enc=dict(density=False, var=False, smoothed_pos=False,
                                                                                       smoothed_dir=False))
    render_case(RenderNet, ray_utils, "small_incl_ray", 9, 400, 16, 8, 5.0, enc=dict(exclude_ray=False))
    render_jitter_case(RenderNet, ray_utils, "small_jitter", 9, 400, 16, 5, 5.0, perturb=1.0, noise_std=0.5, rays_stride=4)
    transition_case(ParticleNet, "small", 8, 0, 0.1, 3)
    transition_case(ParticleNet, "medium", 14, 1, 0.05, 2)


if __name__ == "__main__":
    main()
