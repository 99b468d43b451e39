"""Oracle: one render pass = sampling -> sigma -> weights -> accumulate -> rgb -> background.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``NeRFOCCRenderer.render`` (`/root/reference/modules/scene/nerf_renderer.py:112-209`)
with the occupancy sampler replaced by the fixed-S sampler (``oracle/sampler.py``), and
``NeRFScene.render`` (`/root/reference/modules/scene/nerf.py:74-99`: 32768-ray chunks).
Golden fixtures produced by running the reference's own ``NeRFOCCRenderer.render`` /
``NGPNeRF`` on top of this package's third-party restatements live in
``tests/golden/render_*.npz``.
"""
from __future__ import annotations

import torch

from .composite import accumulate_along_rays, render_weight_from_density
from .field import Field, query_density, query_rgb
from .raygen import gen_pano_rays
from .sampler import fixed_samples


def render_rays(field: Field, rays_o: torch.Tensor, rays_d: torch.Tensor, n_samples: int,
                near: float = 1e-2, far: float = 1.0, training: bool = False,
                jitter: torch.Tensor | None = None, bg_noise: torch.Tensor | None = None,
                mixed: bool = True, accum=torch.float64):
    """``rays_o``, ``rays_d`` [R,3] -> dict(rgb [R,3], distance [R,1], opacities [R,1],
    weights/trans/t_starts/t_ends [R,S], sigmas [R,S], rgbs [R,S,3]).

    ``bg_noise`` [R,4] (training only): columns 0..2 are the ``rand_noise`` background
    colour, column 3 the U[0,1) number of the distance perturbation
    (`nerf_renderer.py:185-194`)."""
    R = rays_o.shape[0]
    t_starts, t_ends = fixed_samples(R, n_samples, near, far, jitter if training else None)
    # nerf_renderer.py:127  positions = o + d * (t_s + t_e)[:, None] / 2
    positions = rays_o[:, None, :] + rays_d[:, None, :] * (t_starts + t_ends)[..., None] / 2.0
    sigmas = query_density(field, positions, mixed, accum).squeeze(-1)
    weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas)
    opacities = accumulate_along_rays(weights)
    mid = ((t_starts + t_ends) / 2.0)[..., None]
    distances = accumulate_along_rays(weights, mid)
    rgbs = query_rgb(field, positions, mixed, accum)
    colors = accumulate_along_rays(weights.detach(), rgbs)
    if training:
        noise = torch.zeros(R, 4) if bg_noise is None else bg_noise
        distances = torch.relu(distances + (noise[:, 3:4] * 2. - 1.) * (1. - opacities))
        colors = colors + noise[:, :3] * (1. - opacities).detach()
    else:
        distances = distances + 5. * (1. - opacities)
        colors = colors + .5 * (1. - opacities)
    return {'rgb': colors, 'distance': distances, 'opacities': opacities, 'weights': weights,
            'trans': trans, 't_starts': t_starts, 't_ends': t_ends, 'sigmas': sigmas, 'rgbs': rgbs}


def render_pano(field: Field, pose: torch.Tensor, h: int, w: int, n_samples: int,
                near: float = 1e-2, far: float = 1.0, chunk: int = 32768,
                row0: int = 0, rows: int | None = None, mixed: bool = True,
                accum=torch.float32):
    """Eval-mode panorama: dict(rgb [rows,w,3], distance [rows,w,1]) for image rows
    [row0, row0+rows)."""
    rows = h - row0 if rows is None else rows
    o, d = gen_pano_rays(pose, h, w)
    o, d = o[row0:row0 + rows].reshape(-1, 3), d[row0:row0 + rows].reshape(-1, 3)
    rgb, dist = [], []
    for s in range(0, o.shape[0], chunk):
        r = render_rays(field, o[s:s + chunk], d[s:s + chunk], n_samples, near, far,
                        mixed=mixed, accum=accum)
        rgb.append(r['rgb']); dist.append(r['distance'])
    return {'rgb': torch.cat(rgb).reshape(rows, w, 3), 'distance': torch.cat(dist).reshape(rows, w, 1)}
