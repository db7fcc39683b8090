"""CPU oracle for hot path 1: the particle-driven NeRF renderer.  TEST INFRASTRUCTURE ONLY.

A functional fp32 torch restatement of the reference's `RenderNet.forward / coarse_rendering /
fine_rendering` over a plain state dict (no nn.Module), so it runs on the GPU box where
/root/reference does not exist.  Each function cites the reference lines it follows.  Pinned by
tests/test_oracle.py against tests/golden/render_*.npz, which oracle/make_golden.py produced by
running the reference's *own unmodified* models/renderer.py + models/nerf.py + utils/ray_utils.py
(third-party ops shimmed, see oracle/third_party_ops.py: "parity unpinned" applies to those).
"""
from __future__ import annotations

import torch

from . import third_party_ops as tpo


# models/nerf.py:21-38
def positional_encoding(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    out = [x]
    for f in 2.0 ** torch.linspace(0, n_freqs - 1, n_freqs):
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


# models/nerf.py:83-124
def operand_rounding(dtype=torch.float16):
    """Straight-through rounding of a tensor to a 16-bit operand type: what the CUDA path's tensor-core layers see."""
    return lambda t: t + (t.to(dtype).to(t.dtype) - t).detach()


def nerf_mlp(sd, net: str, x: torch.Tensor, in_xyz: int, in_dir: int, sigma_only=False, D=8, skips=(4,), quant=None):
    """`quant` (None = the reference's fp32): a straight-through rounding applied to the operands of the ten layers that
    run on the tensor cores in the CUDA path (fp16 inputs and weights, fp32 accumulate; the sigma / rgb heads stay fp32).
    Gradient tests use it to compare like with like: the ReLU kinks of an fp16-operand network sit elsewhere."""
    q = quant or (lambda t: t)
    def lin(name, v, tc=True):
        w = sd[f"{net}.{name}.weight"]
        return torch.nn.functional.linear(q(v), q(w), sd[f"{net}.{name}.bias"]) if tc else \
            torch.nn.functional.linear(v, w, sd[f"{net}.{name}.bias"])
    xyz = x[:, :in_xyz]
    h = xyz
    for i in range(D):
        if i in skips:
            h = torch.cat([xyz, h], -1)
        h = torch.relu(lin(f"xyz_encoding_{i + 1}.0", h))
    sigma = lin("sigma", h, tc=False)
    if sigma_only:
        return sigma
    final = lin("xyz_encoding_final", h)
    d = torch.relu(lin("dir_encoding.0", torch.cat([final, x[:, in_xyz:in_xyz + in_dir]], -1)))
    rgb = torch.sigmoid(lin("rgb.0", d, tc=False))
    return torch.cat([rgb, sigma], -1)


# utils/ray_utils.py:232-256 (perturb == 0, use_disp == False: the only branch the trainers use)
def coarse_z_table(near: float, far: float, n_samples: int) -> torch.Tensor:
    t = torch.linspace(0, 1, n_samples)
    return near * (1 - t) + far * t


def coarse_samples(near, far, rays, n_samples, perturb=0.0, z_rand=None):
    z = coarse_z_table(near, far, n_samples).expand(rays.shape[0], n_samples)
    if perturb > 0:                                     # utils/ray_utils.py:245-253; z_rand = the torch.rand draw
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper, lower = torch.cat([mid, z[:, -1:]], -1), torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * z_rand)
    xyz = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    return z, xyz


# models/renderer.py:112-122
def search(xyz, particles, radius, K):
    R, S, _ = xyz.shape
    d2, idx = tpo.ball_query_shared(xyz.reshape(-1, 3), particles, K, radius)
    d2, idx = d2.view(R, S, K), idx.view(R, S, K)
    nn = particles[idx.clamp(min=0).reshape(-1)].view(R, S, K, 3) * (idx >= 0).unsqueeze(-1).float()
    return d2, idx, nn


# models/renderer.py:96-109 (exclude_ray=True) and :125-179
def local_geometry_features(d2, nn, xyz, rays, ro, radius, enc, sigma_only=False):
    R, S, K = d2.shape
    valid = d2 != 0
    num_nn = valid.sum(-1, keepdim=True)
    pos_feats = [positional_encoding(xyz.reshape(-1, 3), 10)]
    # smoothing: padded slots are zeros and behave like a particle at the origin
    dist = torch.norm(nn - xyz.unsqueeze(-2), dim=-1)
    w = torch.clamp(1 - (dist / radius) ** 3, min=0)
    density = w.sum(-1, keepdim=True)
    smoothed = (w.unsqueeze(-1) * nn).sum(-2) / (density + 1e-12)
    if not enc.exclude_ray:                             # models/renderer.py:100-109
        alpha = torch.full((R, S, 1), 0.9)
        if not getattr(enc, "same_smooth_factor", False):
            alpha[num_nn.le(20)] = 0.1
        smoothed = xyz * (1 - alpha) + smoothed * alpha
    sm = smoothed.reshape(-1, 3)
    sdir = sm - ro.view(1, 3)
    sdir = sdir / torch.norm(sdir, dim=-1, keepdim=True)
    if enc.density:
        pos_feats.append(positional_encoding(density.reshape(-1, 1), 4))
    if enc.smoothed_pos:
        pos_feats.append(positional_encoding(sm, 10))
    if enc.var:
        v = (nn - xyz.unsqueeze(-2)) * valid.unsqueeze(-1)
        mean = v.sum(-2) / (num_nn + 1e-12)
        var = (((v - mean.unsqueeze(-2)) ** 2) * valid.unsqueeze(-1)).sum(-2) / (num_nn + 1e-12)
        pos_feats.append(positional_encoding(var.reshape(-1, 3), 10))
    if sigma_only:
        return torch.cat(pos_feats, 1), None, num_nn
    dir_feats = [torch.repeat_interleave(positional_encoding(rays[:, 3:6], 4), S, dim=0)]
    if enc.smoothed_dir:
        dir_feats.append(positional_encoding(sdir, 4))
    return torch.cat(pos_feats, 1), torch.cat(dir_feats, 1), num_nn


# models/renderer.py:182-208
def composite(rgbsigma, z, rays, white_background=True, noise=None):
    rgb, sigma = rgbsigma[..., :3], rgbsigma[..., 3]
    if noise is not None:                               # models/renderer.py:192-196 (noise = randn * noise_std)
        sigma = sigma + noise
    delta = torch.cat([z[:, 1:] - z[:, :-1], 1e10 * torch.ones_like(z[:, :1])], -1)
    delta = delta * torch.norm(rays[:, 3:6].unsqueeze(1), dim=-1)
    alpha = 1 - torch.exp(-delta * torch.relu(sigma))
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    acc = weights.sum(1)
    out = (weights.unsqueeze(-1) * rgb).sum(-2)
    depth = (weights * z).sum(-1)
    if white_background:
        out = out + 1 - acc.unsqueeze(-1)
    return out, depth, weights


# utils/ray_utils.py:178-229 (det=True)
def importance_samples(z, weights, n_importance, rays, u=None):
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    w = weights[:, 1:-1] + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    if u is None:                                       # det = True; else the torch.rand draw of sample_pdf
        u = torch.linspace(0.0, 1.0, n_importance).expand(cdf.shape[0], n_importance)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    zs = b0 + (u - c0) / denom * (b1 - b0)
    z_all, _ = torch.sort(torch.cat([z, zs], -1), -1)
    xyz = rays[:, None, 0:3] + rays[:, None, 3:6] * z_all[:, :, None]
    return xyz, z_all


# the 16 raw (not yet encoded) values per sample behind the feature row of models/renderer.py:125-179, in the column
# order the CUDA path stores them: [x(3), density, smoothed(3), variance(3), ray dir(3), smoothed dir(3)]
def local_geometry_records(d2, nn, xyz, rays, ro, radius):
    R, S, K = d2.shape
    valid = d2 != 0
    num_nn = valid.sum(-1, keepdim=True)
    dist = torch.norm(nn - xyz.unsqueeze(-2), dim=-1)                       # :97-98 (padded slots sit at the origin)
    w = torch.clamp(1 - (dist / radius) ** 3, min=0)
    density = w.sum(-1, keepdim=True)
    smoothed = (w.unsqueeze(-1) * nn).sum(-2) / (density + 1e-12)
    sdir = smoothed - ro.view(1, 1, 3)
    sdir = sdir / torch.norm(sdir, dim=-1, keepdim=True)
    v = (nn - xyz.unsqueeze(-2)) * valid.unsqueeze(-1)                      # :160-166 two-pass variance
    mean = v.sum(-2) / (num_nn + 1e-12)
    var = (((v - mean.unsqueeze(-2)) ** 2) * valid.unsqueeze(-1)).sum(-2) / (num_nn + 1e-12)
    rd = rays[:, None, 3:6].expand(R, S, 3)
    return torch.cat([xyz, density, smoothed, var, rd, sdir], -1)


def feature_widths(enc):
    in_xyz = 63 + (9 if enc.density else 0) + (63 if enc.var else 0) + (63 if enc.smoothed_pos else 0)
    in_dir = 27 + (27 if enc.smoothed_dir else 0)
    return in_xyz, in_dir


def _pass(sd, net, cfg, radius, K, particles, ro, rays, xyz, z, sigma_only=False, quant=None):
    in_xyz, in_dir = feature_widths(cfg.encoding)
    d2, idx, nn = search(xyz, particles, radius, K)
    fx, fd, num_nn = local_geometry_features(d2, nn, xyz, rays, ro, radius, cfg.encoding, sigma_only)
    mask = (d2 != 0).all(-1, keepdim=True).float()
    S = xyz.shape[1]
    if sigma_only:
        out = nerf_mlp(sd, net, fx, in_xyz, in_dir, sigma_only=True, quant=quant).view(-1, S, 1)
    else:
        out = nerf_mlp(sd, net, torch.cat([fx, fd], 1), in_xyz, in_dir, quant=quant).view(-1, S, 4)
    if cfg.use_mask:
        out = out * mask
    return out, num_nn, mask, idx


@torch.no_grad()
def render_forward(sd, cfg, near, far, particles, ro, rays, mode="forward", white_background=True, debug=False, perturb=0.0,
                   noise_std=0.0, jitter=None):
    """mode: 'forward' (models/renderer.py:211-270), 'coarse' (:273-307), 'fine' (:310-369)."""
    return _render_forward(sd, cfg, near, far, particles, ro, rays, mode, white_background, debug, None, None, perturb, noise_std,
                           jitter)


def render_forward_grad(sd, cfg, near, far, particles, ro, rays, mode="forward", white_background=True, z1_override=None,
                        quant=None, perturb=0.0, noise_std=0.0, jitter=None):
    """The same forward with autograd recording: gradients flow to `sd`'s tensors and to `particles` (through the
    gathered neighbour positions, as through pytorch3d's masked_gather); the importance samples are detached
    (utils/ray_utils.py:224).  `z1_override` (R, S0+S_imp): use these merged depths instead of resampling (parity tests
    feed the CUDA path's own depths, which depend on its fp16-operand coarse sigmas)."""
    with torch.enable_grad():
        return _render_forward(sd, cfg, near, far, particles, ro, rays, mode, white_background, False, z1_override, quant, perturb,
                               noise_std, jitter)


def _render_forward(sd, cfg, near, far, particles, ro, rays, mode="forward", white_background=True, debug=False, z1_override=None,
                    quant=None, perturb=0.0, noise_std=0.0, jitter=None):
    """`jitter` (perturb / noise_std): the random draws of the reference, by name: z_rand (R,S) ~ U, noise0 (R,S) ~ N,
    u (R,S_imp) ~ U, noise1 (R,S+S_imp) ~ N."""
    jitter = jitter or {}
    n0 = jitter["noise0"] * noise_std if noise_std > 0 else None
    n1 = jitter["noise1"] * noise_std if noise_std > 0 else None
    particles, ro, rays = particles.float().cpu(), ro.float().cpu(), rays.float().cpu()
    radius = cfg.NN_search.search_raduis_scale * cfg.NN_search.particle_radius
    K = cfg.NN_search.N_neighbor
    S, S_imp = cfg.ray.N_samples, cfg.ray.N_importance
    res = {}
    z0, xyz0 = coarse_samples(near, far, rays, S, perturb, jitter.get("z_rand"))
    if mode == "fine":
        sig, num0, mask0, idx0 = _pass(sd, "nerf_coarse", cfg, radius, K, particles, ro, rays, xyz0, z0, True, quant)
        fake = torch.cat([torch.zeros(sig.shape[0], S, 3), sig], -1)
        _, _, w0 = composite(fake, z0, rays, white_background, n0)
    else:
        out0, num0, mask0, idx0 = _pass(sd, "nerf_coarse", cfg, radius, K, particles, ro, rays, xyz0, z0, False, quant)
        rgb0, depth0, w0 = composite(out0, z0, rays, white_background, n0)
        res.update(rgb0=rgb0, depth0=depth0, opacity0=w0.sum(1), num_nn_0=num0, mask_0=mask0.sum(1))
    if debug:
        res.update(dbg_idx0=idx0, dbg_w0=w0)
    if mode != "coarse" and S_imp > 0:
        xyz1, z1 = importance_samples(z0, w0.detach(), S_imp, rays, jitter.get("u") if perturb != 0 else None)
        if z1_override is not None:
            z1 = z1_override.float().cpu()
            xyz1 = rays[:, None, 0:3] + rays[:, None, 3:6] * z1[:, :, None]
        xyz1, z1 = xyz1.detach(), z1.detach()                                  # utils/ray_utils.py:224
        out1, num1, mask1, idx1 = _pass(sd, "nerf_fine", cfg, radius, K, particles, ro, rays, xyz1, z1, False, quant)
        rgb1, depth1, w1 = composite(out1, z1, rays, white_background, n1)
        res.update(rgb1=rgb1, depth1=depth1, opacity1=w1.sum(1), num_nn_1=num1, mask_1=mask1.sum(1))
        if debug:
            res.update(dbg_z1=z1, dbg_idx1=idx1)
    return res
