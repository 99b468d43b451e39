"""``nerfacc.estimators.occ_grid.OccGridEstimator`` (0.5.3 API, ``levels=1``) for PeRF's call sites
(`/root/reference/modules/scene/nerf.py:68,144,159-168`, `/root/reference/modules/scene/nerf_renderer.py:145-155`).

SURVEY.md section 8(f) row 1.  Keeps the reference's estimator interface (buffers
``resolution/aabbs/occs/binaries`` for checkpoints, ``update_every_n_steps``, ``sampling``); the
marching is two libperfb200 kernels (perf_occ_count / perf_occ_write) around one cumsum, the
per-sample work it feeds (sigma_fn -> field kernels, transmittance culling) runs on libperfb200
too; the grid update is three libperfb200 kernels (perf_occ_points, perf_occ_update) around the caller's
``occ_eval_fn``.  Sampling rule (nerfacc's DDA for ``cone_angle=0``, restated; upstream source is not
vendored: parity unpinned, the lattice decision is written in oracle/occ_sampler.py): fixed lattice
``t_k = near + (k + u_r) * step`` (one uniform offset ``u_r`` per ray when ``stratified``), a
sample ``[t_k, t_k + step)`` is kept when its midpoint lies inside the aabb in an occupied cell,
then samples whose transmittance fell below ``early_stop_eps`` are dropped.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from perf_b200 import ops

__perf_b200_shim__ = True


class OccGridEstimator(torch.nn.Module):
    DIM = 3

    def __init__(self, roi_aabb, resolution=128, levels: int = 1, **kwargs):
        super().__init__()
        if levels != 1:
            raise NotImplementedError("perf_b200 nerfacc: OccGridEstimator levels != 1 is not implemented (PeRF uses 1)")
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        res = torch.tensor(resolution, dtype=torch.int32)
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        assert roi.numel() == 6
        self.levels = levels
        self.cells_per_lvl = int(res.prod().item())
        self.register_buffer("resolution", res)
        self.register_buffer("aabbs", roi[None, :].clone())
        self.register_buffer("occs", torch.zeros(self.cells_per_lvl * levels))
        self.register_buffer("binaries", torch.zeros([levels] + res.tolist(), dtype=torch.bool))
        self._res_host = [int(v) for v in res.tolist()]

    # -- helpers -------------------------------------------------------------------------------
    def _cell_index(self, x: torch.Tensor) -> torch.Tensor:
        """Flat cell index (x slowest, z fastest: meshgrid 'ij' order) of points inside the aabb."""
        res = self.resolution.to(x.device)
        amin, amax = self.aabbs[0, :3], self.aabbs[0, 3:]
        u = ((x - amin) / (amax - amin) * res).floor().long()
        u = torch.minimum(u.clamp_(min=0), (res - 1).long())
        return (u[..., 0] * int(res[1]) + u[..., 1]) * int(res[2]) + u[..., 2]

    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn: Optional[Callable] = None, alpha_fn: Optional[Callable] = None,
                 near_plane: float = 0.0, far_plane: float = 1e10, t_min=None, t_max=None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 stratified: bool = False, cone_angle: float = 0.0):
        if cone_angle != 0.0:
            raise NotImplementedError("perf_b200 nerfacc: cone_angle != 0 is not implemented (PeRF passes 0)")
        if alpha_fn is not None:
            raise NotImplementedError("perf_b200 nerfacc: alpha_fn is not implemented (PeRF passes sigma_fn)")
        dev, R = rays_o.device, rays_o.shape[0]
        if t_min is not None or t_max is not None:
            raise NotImplementedError("perf_b200 nerfacc: per-ray t_min / t_max are not implemented (PeRF does not pass them)")
        jitter = torch.rand(R, device=dev) if stratified else None
        ray_indices, t_starts, t_ends = ops.occ_sample(self.binaries[0], self._aabb_list(), rays_o.float(), rays_d.float(),
                                                       float(near_plane), float(min(far_plane, 3.0e38)), float(render_step_size), jitter)
        # visibility culling (nerfacc render_visibility_from_density): keep T >= early_stop_eps
        if sigma_fn is not None and early_stop_eps > 0 and ray_indices.numel() > 0:
            sigmas = sigma_fn(t_starts, t_ends, ray_indices).float().reshape(-1)
            _, trans, alphas = ops.weights_from_density(t_starts, t_ends, sigmas.contiguous(), ray_indices, R)
            # min(alpha_thre, mean(occs)): occs >= 0, so for alpha_thre <= 0 (PeRF passes 0) it is alpha_thre itself --
            # no 16.7 M-cell reduction and no device sync per call
            thre = alpha_thre if alpha_thre <= 0 else min(alpha_thre, float(self.occs.mean().item()))
            keep = (trans >= early_stop_eps) & (alphas >= thre)
            ray_indices, t_starts, t_ends = ray_indices[keep], t_starts[keep], t_ends[keep]
        return ray_indices, t_starts, t_ends

    def _aabb_list(self):
        """The roi as Python floats, cached (a .tolist() per call is a device sync per frame)."""
        key = (self.aabbs.data_ptr(), self.aabbs._version)
        if getattr(self, "_aabb_key", None) != key:
            self._aabb_host, self._aabb_key = self.aabbs[0].tolist(), key
        return self._aabb_host

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        if not self.training:
            raise RuntimeError("update_every_n_steps() should only be called in training mode")
        if step % n == 0:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre, ema_decay, warmup_steps):
        dev = self.occs.device
        if step < warmup_steps:
            idx, n = None, self.cells_per_lvl                    # every cell
        else:
            n4 = self.cells_per_lvl // 4
            uni = torch.randint(self.cells_per_lvl, (n4,), device=dev)
            occ_idx = torch.nonzero(self.binaries.reshape(-1))[:, 0]
            if occ_idx.numel() > n4:
                occ_idx = occ_idx[torch.randint(occ_idx.numel(), (n4,), device=dev)]
            idx = torch.cat([uni, occ_idx]).contiguous()
            n = idx.numel()
        if not self.occs.is_cuda:
            raise RuntimeError("perf_b200 nerfacc: OccGridEstimator.update_every_n_steps needs the estimator on a CUDA device (there is no CPU path)")
        self._n_updates = getattr(self, "_n_updates", 0) + 1
        seed = (torch.initial_seed() * 0x9E3779B1 + self._n_updates * 0x85EBCA77C2B2AE63) & (2 ** 62 - 1)
        x = ops.occ_points(idx, n, self._res_host, self._aabb_list(), seed, dev)
        occ = occ_eval_fn(x).reshape(-1).float()
        if getattr(self, "_ws", None) is None or self._ws.device != dev:
            self._ws = torch.empty(2 * 1024, dtype=torch.float64, device=dev)          # PERF_OCC_PARTIALS
        if not self.binaries.is_contiguous():
            self.binaries = self.binaries.contiguous()
        ops.occ_update(self.occs, idx, occ, ema_decay, occ_thre, self.binaries.view(torch.uint8), self._ws)
