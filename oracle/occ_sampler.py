"""Oracle: occupancy-grid interval sampler (nerfacc ``OccGridEstimator.sampling`` for ``levels=1``,
``cone_angle=0``), the sampler PeRF trains and renders with
(`/root/reference/modules/scene/nerf_renderer.py:145-155`).

TEST INFRASTRUCTURE ONLY.  nerfacc 0.5.3 is an un-vendored pip dependency: the exact phase of its DDA
lattice is **unpinned**; this restates the rule our kernels implement so they can be checked:
``t_k = near + (k + u_r) * step``; interval ``[t_k, t_k + step)`` is kept when its midpoint lies in
the ray/aabb overlap (clipped to [near, far]) and inside an occupied cell of ``binaries``
(cell index x-slowest, as nerfacc's ``meshgrid(..., indexing='ij')``)."""
from __future__ import annotations

import torch


def occ_sample(binaries: torch.Tensor, aabb: torch.Tensor, rays_o, rays_d, near: float, far: float, step: float, jitter=None):
    R = rays_o.shape[0]
    res = torch.tensor(binaries.shape)
    amin, amax = aabb[:3], aabb[3:]
    inv = 1.0 / torch.where(rays_d.abs() < 1e-12, torch.full_like(rays_d, 1e-12), rays_d)
    t0, t1 = (amin - rays_o) * inv, (amax - rays_o) * inv
    tn = torch.minimum(t0, t1).amax(-1).clamp(min=near)
    tf = torch.maximum(t0, t1).amin(-1).clamp(max=far)
    u = torch.zeros(R) if jitter is None else jitter
    k_hi = int(torch.ceil(((tf - near) / step).clamp(min=0).max()).item()) + 1 if R else 0
    ks = torch.arange(k_hi, dtype=torch.float32)
    ts = near + (ks[None, :] + u[:, None]) * step
    mid = ts + 0.5 * step
    ok = (mid >= tn[:, None]) & (mid <= tf[:, None]) & (tf >= tn)[:, None]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]
    c = ((pts - amin) / (amax - amin) * res).floor().long()
    c = torch.minimum(c.clamp(min=0), (res - 1))
    ok &= binaries.reshape(-1)[(c[..., 0] * int(res[1]) + c[..., 1]) * int(res[2]) + c[..., 2]]
    r_idx, k_idx = ok.nonzero(as_tuple=True)
    t_starts = ts[r_idx, k_idx]
    return r_idx, t_starts, t_starts + step
