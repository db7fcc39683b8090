"""Parameter containers with the reference's module/state-dict layout (models/nerf.py).

`Embedding` and `NeRF` exist so that `RenderNet.state_dict()` has exactly the reference's keys
(`nerf_coarse.xyz_encoding_1.0.weight`, ...) and optimisers can own the parameters.  RenderNet's
hot path never calls their `forward`: the encodings and all twelve linear layers run inside the fused
tcgen05 kernel (csrc/nf_mlp.cu).  The eager `forward`s below are the plain-torch definition of the
same functions, kept for API compatibility (direct calls, autograd-based fine-tuning).
"""
from __future__ import annotations

import torch
from torch import nn


class Embedding(nn.Module):
    """x -> (x, sin(2^k x), cos(2^k x), ...), k < N_freqs   (models/nerf.py:4-38)."""

    def __init__(self, in_channels: int, N_freqs: int, logscale: bool = True):
        super().__init__()
        self.N_freqs, self.in_channels = N_freqs, in_channels
        self.out_channels = in_channels * (2 * N_freqs + 1)
        self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs) if logscale else \
            torch.linspace(1, 2 ** (N_freqs - 1), N_freqs)

    def forward(self, x):
        parts = [x]
        for f in self.freq_bands:
            parts += [torch.sin(f * x), torch.cos(f * x)]
        return torch.cat(parts, -1)


class NeRF(nn.Module):
    """8x256 MLP with a skip at layer 5, sigma head, 128-wide direction branch (models/nerf.py:41-124)."""

    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=(4,)):
        super().__init__()
        self.D, self.W, self.skips = D, W, list(skips)
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        for i in range(D):
            fin = in_channels_xyz if i == 0 else (W + in_channels_xyz if i in self.skips else W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(nn.Linear(fin, W), nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2), nn.ReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())

    def ordered_params(self):
        """(weight, bias) x [xyz_encoding_1..8, final, dir, sigma, rgb] -- nf_render_pack_weights order."""
        mods = [getattr(self, f"xyz_encoding_{i + 1}")[0] for i in range(self.D)]
        mods += [self.xyz_encoding_final, self.dir_encoding[0], self.sigma, self.rgb[0]]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def forward(self, x, sigma_only=False):
        xyz = x[..., :self.in_channels_xyz]
        h = xyz
        for i in range(self.D):
            if i in self.skips:
                h = torch.cat([xyz, h], -1)
            h = getattr(self, f"xyz_encoding_{i + 1}")(h)
        sigma = self.sigma(h)
        if sigma_only:
            return sigma
        feat = self.xyz_encoding_final(h)
        d = self.dir_encoding(torch.cat([feat, x[..., self.in_channels_xyz:]], -1))
        return torch.cat([self.rgb(d), sigma], -1)
