"""``nerfacc`` (0.5.3) plugin API on libperfb200: the functions PeRF's renderer calls
(`/root/reference/modules/scene/nerf_renderer.py:5-7,170-183`)."""
from __future__ import annotations

from typing import Optional

import torch

from perf_b200 import ops

__perf_b200_shim__ = True
__version__ = "0.5.3+perf_b200"


def _n_rays(ray_indices: torch.Tensor, n_rays: Optional[int]) -> int:
    if n_rays is not None:
        return int(n_rays)
    return int(ray_indices[-1].item()) + 1 if ray_indices.numel() else 0     # sorted by ray (one host sync, as upstream)


class _WeightsFromDensity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t_starts, t_ends, sigmas, ray_indices, n_rays):
        t_starts, t_ends, sigmas = t_starts.float().contiguous(), t_ends.float().contiguous(), sigmas.float().contiguous()
        w, T, a = ops.weights_from_density(t_starts, t_ends, sigmas, ray_indices, n_rays)
        ctx.save_for_backward(t_starts, t_ends, sigmas, ray_indices, w, T)
        ctx.n_rays = n_rays
        ctx.mark_non_differentiable(a)
        return w, T, a

    @staticmethod
    def backward(ctx, gw, gT, ga):
        t_starts, t_ends, sigmas, ray_indices, w, T = ctx.saved_tensors
        gw = torch.zeros_like(w) if gw is None else gw.contiguous()
        gs = ops.weights_from_density_bwd(t_starts, t_ends, sigmas, ray_indices, ctx.n_rays, w, T, gw,
                                          None if gT is None else gT.contiguous())
        return None, None, gs, None, None


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """-> (weights, trans, alphas); packed samples sorted by ray (``ray_indices``), or dense [R,S]."""
    if prefix_trans is not None or packed_info is not None:
        raise NotImplementedError("perf_b200 nerfacc: packed_info / prefix_trans are not implemented (PeRF passes ray_indices)")
    if ray_indices is None:                                   # dense [n_rays, n_samples]
        R, S = sigmas.shape
        ri = torch.arange(R, device=sigmas.device).repeat_interleave(S)
        w, T, a = _WeightsFromDensity.apply(t_starts.reshape(-1), t_ends.reshape(-1), sigmas.reshape(-1), ri, R)
        return w.view(R, S), T.view(R, S), a.view(R, S)
    return _WeightsFromDensity.apply(t_starts, t_ends, sigmas, ray_indices.contiguous(), _n_rays(ray_indices, n_rays))


def render_transmittance_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """T_i = prod_{j<i} (1 - alpha_j), via the density kernel with sigma*dt = -log(1 - alpha)."""
    sd = -torch.log1p(-alphas.float().clamp(max=1 - 1e-7))
    zeros = torch.zeros_like(sd)
    return render_weight_from_density(zeros, torch.ones_like(sd), sd, packed_info, ray_indices, n_rays, prefix_trans)[1]


class _Accumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, values, ray_indices, n_rays):
        weights = weights.float().contiguous()
        vals = None if values is None else values.float().contiguous()
        ctx.save_for_backward(weights, vals if vals is not None else weights, ray_indices)
        ctx.has_values, ctx.vdtype = values is not None, (None if values is None else values.dtype)
        return ops.accumulate_along_rays(weights, vals, ray_indices, n_rays)

    @staticmethod
    def backward(ctx, gout):
        weights, vals, ray_indices = ctx.saved_tensors
        g = gout[ray_indices]                                  # [N, D]
        if not ctx.has_values:
            return g[:, 0], None, None, None
        gw = (g * vals).sum(-1) if ctx.needs_input_grad[0] else None
        gv = (weights[:, None] * g).to(ctx.vdtype) if ctx.needs_input_grad[1] else None
        return gw, gv, None, None


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """out[r] = sum_{i in ray r} w_i * v_i  (values=None -> opacity).  Deterministic."""
    if ray_indices is None:
        src = weights[..., None] if values is None else weights[..., None] * values
        return src.sum(-2)
    return _Accumulate.apply(weights, values, ray_indices.contiguous(), _n_rays(ray_indices, n_rays))
