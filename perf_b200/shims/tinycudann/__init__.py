"""``tinycudann`` plugin API on libperfb200 (sm_100a).

Implements what PeRF calls (`/root/reference/modules/fields/ngp_nerf.py:96-134,179-197,230-248`,
`/root/reference/modules/geo_predictors/pano_joint_predictor.py:30`):
``NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config, seed)``
and ``Encoding(n_input_dims, encoding_config, seed, dtype)`` -- ``torch.nn.Module``s with ONE flat
fp32 ``params`` Parameter in tcnn's layout ``[MLP matrices | grid level 0..L-1]`` (checkpoint
compatible), fp16 outputs, gradients w.r.t. ``params``; ``Encoding`` (Linear or Smoothstep) is also
differentiable w.r.t. its input positions, and once more through that input gradient
(`pano_joint_predictor.py:58-64`).  Anything else raises at construction.
"""
from __future__ import annotations

import torch

from perf_b200 import ops
from perf_b200.config import GridConfig, MLPConfig, network_param_count

__perf_b200_shim__ = True
__version__ = "1.7+perf_b200"


def _init_params(grid: GridConfig, mlp, seed: int) -> torch.Tensor:
    """tcnn init: grid ~ U(-1e-4, 1e-4), MLP matrices Xavier-uniform (upstream's PCG32 stream is not
    reproducible; parity tests load identical params into oracle and kernels)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    parts = []
    if mlp is not None:
        shapes = [(mlp.n_neurons, mlp.n_in)] + [(mlp.n_neurons, mlp.n_neurons)] * (mlp.n_hidden_layers - 1) \
                 + [(mlp.padded_out, mlp.n_neurons)]
        for o, i in shapes:
            lim = (6.0 / (o + i)) ** 0.5
            parts.append((torch.rand(o * i, generator=g, device="cpu") * 2 - 1) * lim)
    parts.append((torch.rand(grid.n_entries * grid.n_features_per_level, generator=g, device="cpu") * 2 - 1) * 1e-4)
    return torch.cat(parts)


class _Base(torch.nn.Module):
    loss_scale = 128.0

    def _half(self) -> torch.Tensor:
        """fp16 shadow of ``params`` (tcnn casts on EVERY forward; we re-cast only when the
        parameter was modified, tracked through the tensor version counter)."""
        p = self.params
        key = (p._version, p.data_ptr())
        if getattr(self, "_half_key", None) != key:
            self._half_cache = ops.params_to_half(p.detach(), out=getattr(self, "_half_cache", None))
            self._half_key = key
        return self._half_cache

    @staticmethod
    def _check_input(x: torch.Tensor, n_in: int, input_grad: bool = False) -> torch.Tensor:
        if x.dim() != 2 or x.shape[1] != n_in:
            raise ValueError(f"expected input of shape [N, {n_in}], got {tuple(x.shape)}")
        if x.requires_grad and not input_grad:
            raise NotImplementedError("perf_b200 tinycudann: NetworkWithInputEncoding has no gradients w.r.t. the input "
                                      "positions (PeRF never asks for them); tcnn.Encoding has")
        return x


class NetworkWithInputEncoding(_Base):
    def __init__(self, n_input_dims: int, n_output_dims: int, encoding_config: dict, network_config: dict, seed: int = 1337):
        super().__init__()
        if n_input_dims != 3:
            raise ValueError(f"perf_b200 tinycudann: n_input_dims={n_input_dims} (only 3)")
        self.grid = GridConfig.from_dict(encoding_config)
        self.mlp = MLPConfig.from_dict(network_config, self.grid.n_features, n_output_dims)
        n = network_param_count(self.grid, self.mlp)          # validates the configuration (raises if unsupported)
        self.n_input_dims, self.n_output_dims, self.seed = n_input_dims, n_output_dims, seed
        self.encoding_config, self.network_config = dict(encoding_config), dict(network_config)
        init = _init_params(self.grid, self.mlp, seed)
        assert init.numel() == n
        self.params = torch.nn.Parameter(init.to(torch.empty(0).device))   # follows torch's default device

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_input(x, self.n_input_dims)
        return ops.network_apply(self.params, x, self.grid, self.mlp, self._half())


class Encoding(_Base):
    def __init__(self, n_input_dims: int, encoding_config: dict, seed: int = 1337, dtype=None):
        super().__init__()
        if n_input_dims != 3:
            raise ValueError(f"perf_b200 tinycudann: n_input_dims={n_input_dims} (only 3)")
        self.grid = GridConfig.from_dict(encoding_config)
        self.n_input_dims, self.n_output_dims, self.seed = n_input_dims, self.grid.n_features, seed
        self.dtype = torch.float16 if dtype is None else dtype
        self.params = torch.nn.Parameter(_init_params(self.grid, None, seed).to(torch.empty(0).device))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_input(x, self.n_input_dims, input_grad=True)
        return ops.encoding_apply(self.params, x, self.grid).to(self.dtype)


class Network(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("perf_b200 tinycudann: stand-alone Network is not used by PeRF and not implemented")


def free_temporary_memory() -> None:   # upstream API; nothing to free, torch owns all memory
    return None
