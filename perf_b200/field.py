"""The PeRF radiance field on libperfb200: host-side mirror of ``NGPNeRF``
(`/root/reference/modules/fields/ngp_nerf.py:68-197`) -- same attribute names, method names,
argument meaning and state-dict keys, so tests read like the reference's call sites and PeRF
checkpoints (`nerf.aabb`, `nerf.geo_mlp.params`, `nerf.app_mlp.params`) load unchanged.
"""
from __future__ import annotations

from typing import List, Union

import torch

from .shims import tinycudann as tcnn

PER_LEVEL_SCALE = 1.4472692012786865
ENCODING_CONFIG = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 18,
                   "base_resolution": 16, "per_level_scale": PER_LEVEL_SCALE}          # ngp_nerf.py:99-106
GEO_NETWORK_CONFIG = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                      "n_neurons": 64, "n_hidden_layers": 1}                            # ngp_nerf.py:107-113
APP_NETWORK_CONFIG = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid",
                      "n_neurons": 64, "n_hidden_layers": 2}                            # ngp_nerf.py:127-133


class _TruncExp(torch.autograd.Function):
    """`ngp_nerf.py:24-40`: forward exp(x) in fp32, backward g * exp(clamp(x, max=15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


trunc_exp = _TruncExp.apply


class NGPNeRF(torch.nn.Module):
    def __init__(self, aabb: Union[torch.Tensor, List[float]], num_dim: int = 3, seed: int = 1337):
        super().__init__()
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.register_buffer("aabb", aabb)
        self.num_dim = num_dim
        self.geo_mlp = tcnn.NetworkWithInputEncoding(num_dim, 1, ENCODING_CONFIG, GEO_NETWORK_CONFIG, seed=seed)
        self.app_mlp = tcnn.NetworkWithInputEncoding(num_dim, 3, ENCODING_CONFIG, APP_NETWORK_CONFIG, seed=seed + 1)

    def _normalise(self, x):
        aabb_min, aabb_max = torch.split(self.aabb, self.num_dim, dim=-1)
        x = (x - aabb_min) / (aabb_max - aabb_min)
        return x, ((x > 0.0) & (x < 1.0)).all(dim=-1)

    def query_density(self, x: torch.Tensor) -> torch.Tensor:
        """`ngp_nerf.py:136-150`: sigma = trunc_exp(geo_mlp(x01)) * selector, shape [..., 1] fp32."""
        x, selector = self._normalise(x)
        raw = self.geo_mlp(x.view(-1, self.num_dim)).view(list(x.shape[:-1]) + [1]).to(x)
        return trunc_exp(raw) * selector[..., None]

    def query_rgb(self, x: torch.Tensor) -> torch.Tensor:
        """`ngp_nerf.py:152-162`: rgb = app_mlp(x01) * selector, shape [..., 3] fp16."""
        x, selector = self._normalise(x)
        rgb = self.app_mlp(x.view(-1, self.num_dim)).view(list(x.shape[:-1]) + [3])
        return rgb * selector[..., None]

    def forward(self, positions: torch.Tensor, directions: torch.Tensor = None, contract=None):
        return self.query_rgb(positions), self.query_density(positions)

    def reset_geo(self) -> None:
        """`ngp_nerf.py:178-197`: fresh density network (the colour network is kept)."""
        dev = self.geo_mlp.params.device
        self.geo_mlp = tcnn.NetworkWithInputEncoding(3, 1, ENCODING_CONFIG, GEO_NETWORK_CONFIG).to(dev)
