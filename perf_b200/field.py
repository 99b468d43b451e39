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


class _ExpWithCappedSlope(torch.autograd.Function):
    """The density activation of `ngp_nerf.py:24-40`: exp(raw) evaluated in fp32; its derivative is taken as
    exp(min(raw, 15)) so that one exploding logit cannot blow up the step (the fused kernels apply the same
    cap: ``composite_bwd_kernel``, ``fminf(sigma, e^15)``)."""

    @staticmethod
    def forward(ctx, raw):
        raw32 = raw.float()
        ctx.save_for_backward(raw32)
        return raw32.exp()

    @staticmethod
    def backward(ctx, grad_sigma):
        (raw32,) = ctx.saved_tensors
        return grad_sigma * raw32.clamp(max=15.0).exp()


trunc_exp = _ExpWithCappedSlope.apply


class NGPNeRF(torch.nn.Module):
    def __init__(self, aabb: Union[torch.Tensor, List[float]], num_dim: int = 3, seed: int = 1337):
        super().__init__()
        self.num_dim = num_dim
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32).clone())      # [min xyz | max xyz]
        self.geo_mlp = tcnn.NetworkWithInputEncoding(num_dim, 1, ENCODING_CONFIG, GEO_NETWORK_CONFIG, seed=seed)
        self.app_mlp = tcnn.NetworkWithInputEncoding(num_dim, 3, ENCODING_CONFIG, APP_NETWORK_CONFIG, seed=seed + 1)

    def _normalise(self, x):
        lo, hi = self.aabb[:self.num_dim], self.aabb[self.num_dim:]
        unit = (x - lo) / (hi - lo)
        return unit, ((unit > 0.0) & (unit < 1.0)).all(dim=-1)

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
