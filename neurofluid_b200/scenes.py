"""Synthetic scenes, cameras and deterministic weights.

The reference's datasets (`datasets/dataset.py`, paths under /data/syguan/...) are not available,
so benches and parity tests run on seeded synthetic stand-ins with the same shapes and scales:

* particles: n^3 jittered lattice, spacing 0.05 (= 2 * particle_radius), randomly permuted --
  the first-K ball query depends on array order, so the permutation matters.
* box: the open container of trainer/basetrainer.py:58-62 (x,y in [-1,1], z in [-1,2.4552]) as a
  surface lattice with inward normals (data_generation/create_physics_scenes.py:170-180 density).
* camera: the 3x4 matrix of eval_renderer.py:67-92, camera_angle_x = 0.323 (configs/warmup.yaml:18);
  rays follow utils/ray_utils.py:85-130 (get_ray_directions / get_rays).
* weights: nn.Linear default init restated with a numpy RandomState so that the same state dict
  can be rebuilt bit-identically on any box (no dependence on torch's RNG stream).

Everything here is host-side numpy/torch set-up code, not part of the timed hot path.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

PARTICLE_RADIUS = 0.025
NEAR, FAR = 9.0, 13.0

# eval_renderer.py:67-92 (first three rows)
CAMERA_C2W = np.array([
    [0.3597943186759949, 0.09052024036645889, -0.18696719408035278, -4.842308521270752],
    [-0.2077273577451706, 0.15678563714027405, -0.32383665442466736, -8.387124061584473],
    [0.0, 0.37393447756767273, 0.181040421128273, 4.688809871673584],
], dtype=np.float32)
CAMERA_ANGLE_X = 0.323


def render_cfg(use_mask=True, n_samples=64, n_importance=128, n_neighbor=20, **enc):
    """Attribute-dict with the fields RenderNet reads (configs/end2end.yaml:32-49)."""
    e = dict(density=True, var=True, smoothed_pos=True, smoothed_dir=True, exclude_ray=True,
             same_smooth_factor=False)
    e.update(enc)
    return SimpleNamespace(
        use_mask=use_mask,
        ray=SimpleNamespace(ray_chunk=1024, N_samples=n_samples, N_importance=n_importance),
        NN_search=SimpleNamespace(fix_radius=True, particle_radius=PARTICLE_RADIUS, search_raduis_scale=9.0,
                                  N_neighbor=n_neighbor),
        encoding=SimpleNamespace(**e))


def lattice_particles(n: int, seed: int = 0, spacing: float = 0.05, jitter: float = 0.005,
                      center=(0.0, 0.0, 0.0), sphere_radius: float | None = None) -> np.ndarray:
    rng = np.random.RandomState(seed)
    ax = (np.arange(n, dtype=np.float64) - (n - 1) / 2.0) * spacing
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    if sphere_radius is not None:
        g = g[np.linalg.norm(g, axis=1) < sphere_radius]
    g = g + rng.uniform(-jitter, jitter, size=g.shape)
    g = g + np.asarray(center, dtype=np.float64)
    g = g[rng.permutation(g.shape[0])]
    return g.astype(np.float32)


def box_points(spacing: float = 0.032):
    """Open box surface lattice + inward unit normals -> (M,3), (M,3)."""
    x0, x1, y0, y1, z0, z1 = -1.0, 1.0, -1.0, 1.0, -1.0, 2.4552

    def axis(a, b):
        m = int(round((b - a) / spacing))
        return np.linspace(a, b, m + 1)

    xs, ys, zs = axis(x0, x1), axis(y0, y1), axis(z0, z1)
    pts, nrm = [], []
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    pts.append(np.stack([X, Y, np.full_like(X, z0)], -1).reshape(-1, 3)); nrm.append((0, 0, 1))
    Y2, Z2 = np.meshgrid(ys, zs[1:], indexing="ij")
    pts.append(np.stack([np.full_like(Y2, x0), Y2, Z2], -1).reshape(-1, 3)); nrm.append((1, 0, 0))
    pts.append(np.stack([np.full_like(Y2, x1), Y2, Z2], -1).reshape(-1, 3)); nrm.append((-1, 0, 0))
    X2, Z3 = np.meshgrid(xs[1:-1], zs[1:], indexing="ij")
    pts.append(np.stack([X2, np.full_like(X2, y0), Z3], -1).reshape(-1, 3)); nrm.append((0, 1, 0))
    pts.append(np.stack([X2, np.full_like(X2, y1), Z3], -1).reshape(-1, 3)); nrm.append((0, -1, 0))
    P = np.concatenate(pts, 0).astype(np.float32)
    N = np.concatenate([np.tile(np.asarray(n, np.float32), (p.shape[0], 1)) for p, n in zip(pts, nrm)], 0)
    return P, N


def camera_rays(H: int, W: int, c2w: np.ndarray = CAMERA_C2W, angle_x: float = CAMERA_ANGLE_X):
    """utils/ray_utils.py:85-130: returns rays (H*W,6) float32 tensor [o, unit d], focal, c2w tensor."""
    focal = 0.5 * W / math.tan(0.5 * angle_x)
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    j, i = torch.meshgrid(ys, xs, indexing="ij")            # i = column (x), j = row (y)
    dirs = torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)
    cw = torch.from_numpy(np.asarray(c2w, dtype=np.float32))
    rd = dirs @ cw[:, :3].T
    rd = rd / torch.norm(rd, dim=-1, keepdim=True)
    ro = cw[:, 3].expand(rd.shape)
    return torch.cat([ro, rd], -1).reshape(-1, 6).contiguous(), focal, cw


def center_crop_rays(rays: torch.Tensor, H: int, W: int, crop: int):
    r = rays.view(H, W, 6)
    a, b = H // 2 - crop // 2, W // 2 - crop // 2
    return r[a:a + crop, b:b + crop].reshape(-1, 6).contiguous()


# ---------------------------------------------------------------------------------------------
# deterministic weights
# ---------------------------------------------------------------------------------------------
def _linear(rng, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)       # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), .)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(b)


def nerf_layer_shapes(in_xyz=198, in_dir=54, W=256, D=8, skips=(4,)):
    """(name, out_features, in_features) in models/nerf.py:57-81 order."""
    shapes = []
    for i in range(D):
        fin = in_xyz if i == 0 else (W + in_xyz if i in skips else W)
        shapes.append((f"xyz_encoding_{i + 1}.0", W, fin))
    shapes.append(("xyz_encoding_final", W, W))
    shapes.append(("dir_encoding.0", W // 2, W + in_dir))
    shapes.append(("sigma", 1, W))
    shapes.append(("rgb.0", 3, W // 2))
    return shapes


def init_render_state(seed: int = 0, sigma_bias_boost: float = 0.0, in_xyz=198, in_dir=54, weight_gain: float = 1.0):
    """State dict with RenderNet's key layout.

    weight_gain scales every weight matrix (2.45 ~ He init): activations then keep O(1) magnitude
    through the 8 layers, giving "trained-like" dynamic range for numerics tests; 1.0 is the
    reference's initial state (nn.Linear default init)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for net in ("nerf_coarse", "nerf_fine"):
        for name, fo, fi in nerf_layer_shapes(in_xyz, in_dir):
            w, b = _linear(rng, fo, fi)
            w = w * weight_gain
            if name == "sigma" and sigma_bias_boost:
                b = b + sigma_bias_boost
            sd[f"{net}.{name}.weight"] = w
            sd[f"{net}.{name}.bias"] = b
    return sd


PARTICLENET_CONVS = [  # (name, cin, cout)  models/transmodel.py:42-71
    ("conv0_fluid", 4, 32), ("conv0_obstacle", 3, 32), ("conv1", 96, 64), ("conv2", 64, 64), ("conv3", 64, 3)]
PARTICLENET_DENSES = [("dense0_fluid", 4, 32), ("dense1", 96, 64), ("dense2", 64, 64), ("dense3", 64, 3)]


def init_particle_state(seed: int = 0, gravity=(0.0, 0.0, -9.81), last_layer_scale: float = 0.1):
    """ParticleNet state dict (key names/shapes of SURVEY.md section 8a-a12)."""
    rng = np.random.RandomState(seed)
    sd = {"gravity": torch.tensor(gravity, dtype=torch.float32)}
    for name, cin, cout in PARTICLENET_CONVS:
        k = rng.uniform(-0.05, 0.05, size=(4, 4, 4, cin, cout)).astype(np.float32)
        if name == "conv3":
            k *= last_layer_scale
        sd[f"{name}.kernel"] = torch.from_numpy(k)
        sd[f"{name}.bias"] = torch.zeros(cout)
        sd[f"{name}.offset"] = torch.zeros(3)
    for name, cin, cout in PARTICLENET_DENSES:
        if name == "dense0_fluid":      # xavier_uniform + zero bias (models/transmodel.py:51-52)
            bound = math.sqrt(6.0 / (cin + cout))
            w = torch.from_numpy(rng.uniform(-bound, bound, size=(cout, cin)).astype(np.float32))
            b = torch.zeros(cout)
        else:
            w, b = _linear(rng, cout, cin)
        if name == "dense3":
            w, b = w * last_layer_scale, b * last_layer_scale
        sd[f"{name}.weight"] = w
        sd[f"{name}.bias"] = b
    return sd
