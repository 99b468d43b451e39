"""Oracle: transmittance scan, per-ray accumulation, distortion loss.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates nerfacc 0.5.3
``render_weight_from_density`` / ``accumulate_along_rays`` (called at
`/root/reference/modules/scene/nerf_renderer.py:170-183`) and
torch_efficient_distloss 0.1.3 ``flatten_eff_distloss`` (called at
`/root/reference/modules/scene/nerf.py:230`); both are un-vendored pip dependencies,
formulas per SURVEY.md Appendix B.  All functions are differentiable torch code.
"""
from __future__ import annotations

import torch


def _exclusive_segment_cumsum(v: torch.Tensor, ray_indices: torch.Tensor) -> torch.Tensor:
    """Exclusive prefix sum of ``v`` restarted at every ray (samples sorted by ray)."""
    inc = torch.cumsum(v.double(), 0)
    exc = inc - v.double()
    first = torch.ones_like(ray_indices, dtype=torch.bool)
    first[1:] = ray_indices[1:] != ray_indices[:-1]
    start_val = exc[first]                                   # exclusive sum at each ray start
    seg_id = torch.cumsum(first.long(), 0) - 1
    return (exc - start_val[seg_id]).to(v.dtype)


def render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=None, n_rays=None):
    """(weights, trans, alphas).  Dense [R,S] inputs when ``ray_indices`` is None, else
    packed [N] sorted by ray.  alpha_i = 1-exp(-s_i d_i); T_i = exp(-sum_{j<i} s_j d_j);
    w_i = T_i alpha_i."""
    sd = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sd)
    if ray_indices is None:
        exc = torch.cumsum(sd, -1) - sd
    else:
        exc = _exclusive_segment_cumsum(sd, ray_indices)
    trans = torch.exp(-exc)
    return trans * alphas, trans, alphas


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """``out[r] = sum_i w_i v_i`` (v=None -> opacity).  Dense [R,S](,D) when
    ``ray_indices`` is None, else packed with ``index_add_``."""
    if values is None:
        src = weights[..., None]
    else:
        src = weights[..., None] * values
    if ray_indices is None:
        return src.sum(-2)
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    return out.index_add_(0, ray_indices, src)


def composite_fixed(t_starts, t_ends, sigmas, rgbs):
    """Dense fixed-S composite -> dict(opacity [R,1], distance [R,1], rgb [R,3],
    weights/trans [R,S]).  Follows `nerf_renderer.py:170-183` (colour uses
    ``weights.detach()``)."""
    w, trans, _ = render_weight_from_density(t_starts, t_ends, sigmas)
    mid = ((t_starts + t_ends) / 2.0)[..., None]
    return {
        'opacities': accumulate_along_rays(w),
        'distance': accumulate_along_rays(w, mid),
        'rgb': accumulate_along_rays(w.detach(), rgbs),
        'weights': w, 'trans': trans,
    }


def flatten_eff_distloss(w, m, interval, ray_id):
    """torch_efficient_distloss: (sum_i interval_i w_i^2 / 3
    + 2 sum_i w_i (m_i W_i^excl - (w m)_i^excl)) / n_rays, per-ray exclusive prefix sums,
    n_rays = ray_id.max()+1."""
    n_rays = int(ray_id.max().item()) + 1
    loss_uni = (1.0 / 3.0) * (interval * w * w).sum()
    wm = w * m
    w_excl = _exclusive_segment_cumsum(w, ray_id)
    wm_excl = _exclusive_segment_cumsum(wm, ray_id)
    loss_bi = 2.0 * (w * (m * w_excl - wm_excl)).sum()
    return (loss_uni + loss_bi) / n_rays
