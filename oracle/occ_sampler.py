"""Oracle: occupancy-grid interval sampler (nerfacc ``OccGridEstimator.sampling`` for ``levels=1``,
``cone_angle=0``), the sampler PeRF trains and renders with
(`/root/reference/modules/scene/nerf_renderer.py:145-155`).

TEST INFRASTRUCTURE ONLY.  nerfacc 0.5.3 is an un-vendored pip dependency (no source on this machine):
**parity unpinned**.  This restates the rule our kernels implement so they can be checked:
``t_k = near + (k + u_r) * step``; interval ``[t_k, t_k + step)`` is kept when its midpoint lies in
the ray/aabb overlap (clipped to [near, far]) and inside an occupied cell of ``binaries``
(cell index x-slowest, as nerfacc's ``meshgrid(..., indexing='ij')``).

DECISION (round 2): ONE GLOBAL LATTICE per ray, not a lattice re-phased at every
occupied-run entry.  Reasons: (1) upstream's ``traverse_grids`` kernel (nerfacc 0.5.x ``csrc/grid.cu``,
recalled, not on disk) keeps one running ``t_last`` per ray that starts at the near plane and only ever
advances by ``dt`` -- inside an occupied cell it first "marches until t_mid is right after" the cell entry
WITHOUT emitting, then emits while ``t_mid`` is before the cell exit; ``t_last`` is never reset to a cell
boundary, so with ``cone_angle = 0`` every emitted interval starts at ``near + k * step`` for an integer k
(up to the fp32 rounding of the running sum, which we replace by the closed form ``near + (k + u) * step``
= one rounding per sample instead of k); (2) nerfacc 0.3.x ``ray_marching`` had the same global-lattice
rule, so the behaviour is stable across the versions PeRF may have been run with; (3) the stratified
offset is "one uniform offset per ray added to the near plane" (SURVEY Appendix B), which only makes
sense on a global lattice.  A re-phased lattice would differ from this one by < 1 step (5e-4) in sample
position; `tests/test_occ_host.py::test_one_cell_grid_reproduces_the_global_lattice` pins the chosen rule on
a single occupied cell in the middle of the box (first emitted interval is NOT aligned to the cell entry)."""
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
