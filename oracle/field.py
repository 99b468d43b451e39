"""Oracle: the PeRF radiance field = two (hash grid + MLP) networks.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
`/root/reference/modules/fields/ngp_nerf.py`:
``NGPNeRF.__init__`` :71-134 (network shapes), ``query_density`` :136-150,
``query_rgb`` :152-162, ``_TruncExp`` :24-40.  The flat ``params`` vector of each
network uses the tcnn ``NetworkWithInputEncoding`` layout
``[ MLP matrices | grid level 0..L-1 ]``, which is what the
reference's checkpoints store under ``nerf.geo_mlp.params`` / ``nerf.app_mlp.params``
(`modules/scene/nerf.py:374-380`).
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field

import torch

from .hashgrid import GridConfig, encode, n_table_entries
from .mlp import MLPConfig, flat_param_count, layer_shapes, mlp_forward, split_params

GEO_MLP = MLPConfig(n_in=32, n_out=1, n_neurons=64, n_hidden_layers=1,
                    activation="ReLU", output_activation="None")      # ngp_nerf.py:107-113
APP_MLP = MLPConfig(n_in=32, n_out=3, n_neurons=64, n_hidden_layers=2,
                    activation="ReLU", output_activation="Sigmoid")   # ngp_nerf.py:127-133
PERF_GRID = GridConfig()                                              # ngp_nerf.py:94,99-106


def network_param_count(grid: GridConfig, mlp: MLPConfig) -> int:
    return flat_param_count(mlp) + n_table_entries(grid) * grid.n_features_per_level


def network_forward(x01: torch.Tensor, params: torch.Tensor, grid: GridConfig,
                    mlp: MLPConfig, mixed: bool = True, accum=torch.float64) -> torch.Tensor:
    """tcnn ``NetworkWithInputEncoding.forward``: encode then MLP.  ``mixed`` rounds the
    whole parameter vector to fp16 first (the torch binding's per-forward
    ``params.to(half)``) and follows the fp16 rounding points of ``mlp_forward``."""
    n_mlp = flat_param_count(mlp)
    p = params.detach().float()
    if mixed:
        p = p.half().float()
    feat = encode(x01, p[n_mlp:], grid, out_half=mixed, exact_fma=(accum == torch.float64), blend="half" if mixed else "fp32")
    return mlp_forward(feat, split_params(p[:n_mlp], mlp), mlp, mixed=mixed, accum=accum)


@dataclass
class Field:
    """The two networks + aabb of ``NGPNeRF`` (`ngp_nerf.py:68-134`)."""
    geo_params: torch.Tensor
    app_params: torch.Tensor
    aabb: torch.Tensor = dc_field(default_factory=lambda: torch.tensor([-1., -1., -1., 1., 1., 1.]))
    grid: GridConfig = PERF_GRID

    @staticmethod
    def random(seed: int = 1337, grid_scale: float = 1e-4, grid: GridConfig = PERF_GRID) -> "Field":
        """Seeded init: grid ~ U(-s, s) (tcnn uses s=1e-4; tests pass a larger ``s`` so
        features are not numerically invisible), MLP Xavier-uniform."""
        g = torch.Generator().manual_seed(seed)

        def make(mlp: MLPConfig):
            ws = []
            for o, i in layer_shapes(mlp):
                lim = (6.0 / (o + i)) ** 0.5
                ws.append((torch.rand(o * i, generator=g) * 2 - 1) * lim)
            n_grid = n_table_entries(grid) * grid.n_features_per_level
            ws.append((torch.rand(n_grid, generator=g) * 2 - 1) * grid_scale)
            return torch.cat(ws)

        return Field(make(GEO_MLP), make(APP_MLP), grid=grid)


def _normalise(x: torch.Tensor, aabb: torch.Tensor):
    """`ngp_nerf.py:137-140`: x01 = (x - min) / (max - min); selector = all(0 < x01 < 1)."""
    aabb_min, aabb_max = torch.split(aabb, 3, dim=-1)
    x01 = (x - aabb_min) / (aabb_max - aabb_min)
    selector = ((x01 > 0.0) & (x01 < 1.0)).all(dim=-1)
    return x01, selector


def query_raw_density(field: Field, x: torch.Tensor, mixed: bool = True, accum=torch.float64):
    """The density logit before ``trunc_exp`` (what the kernels are compared on)."""
    x01, selector = _normalise(x.float(), field.aabb)
    raw = network_forward(x01.reshape(-1, 3), field.geo_params, field.grid, GEO_MLP, mixed, accum)
    return raw.reshape(*x.shape[:-1], 1), selector


def query_density(field: Field, x: torch.Tensor, mixed: bool = True, accum=torch.float64) -> torch.Tensor:
    """`NGPNeRF.query_density` (`ngp_nerf.py:136-150`): sigma = exp(raw) * selector."""
    raw, selector = query_raw_density(field, x, mixed, accum)
    return torch.exp(raw) * selector[..., None]


def query_rgb(field: Field, x: torch.Tensor, mixed: bool = True, accum=torch.float64) -> torch.Tensor:
    """`NGPNeRF.query_rgb` (`ngp_nerf.py:152-162`): rgb = sigmoid-net(x01) * selector."""
    x01, selector = _normalise(x.float(), field.aabb)
    rgb = network_forward(x01.reshape(-1, 3), field.app_params, field.grid, APP_MLP, mixed, accum)
    return rgb.reshape(*x.shape[:-1], 3) * selector[..., None]
