"""Oracle: fixed-S interval sampler (the benchmark's sampling mode).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference's analogue is the
uniform ``PropNetEstimator.sampling(num_samples=S, near_plane, far_plane,
sampling_type='uniform', stratified=training)`` call at
`/root/reference/modules/scene/nerf_renderer.py:60-70` (SURVEY.md §8 a7'); near/far
default to ``to_bounded_rays`` (`modules/scene/nerf.py:313-319`: 1e-2, 1.0).
Stratification follows the sampler PeRF actually trains with, nerfacc's
``OccGridEstimator.sampling``: ONE uniform offset per ray, in units of the step
(SURVEY.md Appendix B), supplied by the caller as a tensor so oracle and kernel see
the same numbers.
"""
from __future__ import annotations

import torch


def fixed_samples(n_rays: int, n_samples: int, near: float = 1e-2, far: float = 1.0,
                  jitter: torch.Tensor | None = None):
    """t_starts, t_ends [R,S] fp32: step=(far-near)/S, t_s[k]=near+(k+u_r)*step,
    t_e[k]=near+(k+1+u_r)*step (so t_e[k] == t_s[k+1] bit for bit)."""
    near_t, far_t = torch.tensor(near, dtype=torch.float32), torch.tensor(far, dtype=torch.float32)
    step = (far_t - near_t) / float(n_samples)
    k = torch.arange(n_samples + 1, dtype=torch.float32)[None, :]
    u = torch.zeros(n_rays, 1) if jitter is None else jitter.reshape(n_rays, 1).float()
    edges = near_t + (k + u) * step
    return edges[:, :-1].contiguous(), edges[:, 1:].contiguous()
