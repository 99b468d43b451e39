"""Supervision pool of the trainer: masks, valid-pixel gather, batch sampler, occlusion checks.

Mirrors `/root/reference/modules/dataset/sup_info.py` for tensors on any device:

  ``PanoSupInfo``                      `sup_info.py:26-117`   edge / normal masks, valid-pixel buffers
  ``SupInfoPool.register_sup_info``    `sup_info.py:160-171`
  ``SupInfoPool.rand_ray_color_data``  `sup_info.py:236-259`  (batch = the same multiset torch.randint draws)
  ``SupInfoPool.geo_check``            `sup_info.py:261-302`
  ``SupInfoPool.gen_occ_grid``         `sup_info.py:304-330`  (inherited from ``RaySupervision``)
  ``SupInfoPool.state_dict`` / ``load_state_dict``   `sup_info.py:332-360`
  ``pano_visibility_mask``             `modules/scene/nerf.py:320-358` (the projection half of
                                       ``NeRFScene.get_pano_visibility_mask``)

The three kornia 0.7.0 functions the reference calls (`requirements.txt:12`; third-party, not vendored)
are restated here in plain torch and pinned against OpenCV in ``tests/test_sup_info.py``:
``kornia.filters.laplacian(x, 3)`` (reflect border, kernel normalised by its L1 norm) and
``kornia.morphology.erosion / dilation`` (flat structuring element, geodesic border).

This is set-up work that runs once per registered panorama -- not the per-ray hot path -- so it is
ordinary torch; the batches it hands out feed the CUDA training step.
"""
from __future__ import annotations


import numpy as np
import torch
import torch.nn.functional as F

from .scene import Rays, RaySupervision


# ------------------------------------------------------------------ kornia 0.7.0 restatements
def laplacian3(x: torch.Tensor) -> torch.Tensor:
    """``kornia.filters.laplacian(x, kernel_size=3)`` (defaults ``border_type='reflect'``,
    ``normalized=True``): kernel = ones(3,3) with centre 1 - 9, divided by sum|k| = 16.  x: [B,C,H,W]."""
    k = torch.ones(3, 3, dtype=x.dtype, device=x.device)
    k[1, 1] = 1.0 - 9.0
    k = k / k.abs().sum()
    c = x.shape[1]
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(xp, k[None, None].expand(c, 1, 3, 3), groups=c)


def _morph(x: torch.Tensor, kernel: torch.Tensor, erode: bool) -> torch.Tensor:
    """kornia.morphology.{erosion,dilation}(x, kernel) with the defaults (flat structuring element,
    ``border_type='geodesic'``: the border never wins the min / max, ``max_val=1e4``).  x: [B,C,H,W]."""
    kh, kw = kernel.shape
    oy, ox = kh // 2, kw // 2
    big = 1e4
    xp = F.pad(x, (ox, kw - ox - 1, oy, kh - oy - 1), mode="constant", value=big if erode else -big)
    sel = kernel if erode else kernel.flip(0, 1)             # dilation reflects the structuring element
    out = None
    H, W = x.shape[-2:]
    for i in range(kh):
        for j in range(kw):
            if float(sel[i, j]) == 0.0:
                continue
            v = xp[..., i:i + H, j:j + W]
            out = v if out is None else (torch.minimum(out, v) if erode else torch.maximum(out, v))
    return out.clone() if out is not None else x.clone()


def erosion(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    return _morph(x, kernel, True)


def dilation(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    return _morph(x, kernel, False)


def ellipse_kernel(size: int, device=None) -> torch.Tensor:
    """``cv.getStructuringElement(cv.MORPH_ELLIPSE, (size, size))`` as a float tensor (OpenCV's rule:
    row i spans the columns within ``round(c * sqrt(1 - (dy/r)^2))`` of the centre)."""
    r, c = size // 2, size // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    k = torch.zeros(size, size)
    for i in range(size):
        dy = i - r
        if abs(dy) <= r:
            dx = int(round(c * float(np.sqrt((r * r - dy * dy) * inv_r2))))
            k[i, max(c - dx, 0):min(c + dx + 1, size)] = 1.0
    return k.to(device) if device is not None else k


# ------------------------------------------------------------------ camera helpers (utils/camera_utils.py)
def apply_rot(pts: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:                     # :44-46
    return torch.matmul(rot, pts[..., None])[..., 0]


def pano_dirs(height: int, width: int, device) -> torch.Tensor:                            # :113-126,142-155
    y = torch.linspace(.5 / height, 1. - .5 / height, height, device=device)
    x = torch.linspace(.5 / width, 1. - .5 / width, width, device=device)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    beta, alpha = -(yy - .5) * np.pi, -(xx - .5) * 2. * np.pi
    return torch.stack([torch.cos(alpha) * torch.cos(beta), torch.sin(alpha) * torch.cos(beta), torch.sin(beta)], -1)


def direction_to_img_coord(dirs: torch.Tensor) -> torch.Tensor:                            # :128-151
    dirs = dirs / torch.linalg.norm(dirs, 2, -1, True)
    beta = torch.arcsin(dirs[..., 2])
    xy = dirs[..., :2] / torch.cos(beta)[..., None]
    alpha = torch.atan2(xy[..., 1], xy[..., 0])
    return torch.stack([-beta / np.pi + .5, -(alpha / (2. * np.pi)) + .5], -1)


def img_coord_to_sample_coord(coords: torch.Tensor) -> torch.Tensor:                       # :180-181
    return torch.stack([coords[..., 1], coords[..., 0]], -1) * 2. - 1.


def morton_key(rows: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    """Z-order code of a pixel (int32, 15 bits per axis): batches sorted by it keep rays that are
    neighbours on the sphere in the same warp (see ``RaySupervision.from_panorama``)."""
    rows, cols = rows.to(torch.int32), cols.to(torch.int32)
    key = torch.zeros_like(rows)
    for b in range(15):
        key |= ((cols >> b) & 1) << (2 * b)
        key |= ((rows >> b) & 1) << (2 * b + 1)
    return key


# ------------------------------------------------------------------ one registered panorama
class PanoSupInfo(torch.nn.Module):
    """`sup_info.py:26-117`.  Buffers: pose, mask_raw, color_map, distance_map, normal_map, mask,
    sup_colors, sup_distances, sup_normals, sup_dirs, sup_positions (+ sup_keys: Morton code of each
    supervised pixel, ours); ``sup_rays`` = Rays(sup_positions, sup_dirs)."""

    def __init__(self, pose, mask, color_map, distance_map, normal_map=None, factor=1):
        super().__init__()
        device = color_map.device
        height, width, _ = color_map.shape
        if distance_map is None:
            distance_map = torch.ones(height, width, 1, device=device)
        else:
            distance_map = distance_map.squeeze()[..., None]
        has_mask = mask is not None
        if has_mask:
            mask = mask.squeeze()[..., None]
        has_normal_map = normal_map is not None
        if not has_normal_map:
            normal_map = torch.zeros(height, width, 3, device=device)
        assert color_map.shape[-1] == 3 and distance_map.shape[-1] == 1
        self.register_buffer("pose", pose.to(device))
        if factor != 1:                                                   # `sup_info.py:54-66` (cv.INTER_AREA)
            factor = int(factor)
            height, width = height // factor, width // factor
            area = lambda m: F.interpolate(m.permute(2, 0, 1)[None].float(), size=(height, width), mode="area")[0].permute(1, 2, 0)
            if color_map.shape[0] % height or color_map.shape[1] % width:
                raise ValueError("factor must divide the panorama size (cv.INTER_AREA with a fractional ratio is not restated)")
            color_map, distance_map, normal_map = area(color_map), area(distance_map), area(normal_map)
            if has_mask:
                # the reference keeps the full-size mask here and fails on the shape mismatch below;
                # we resample it the same way so that factor != 1 is usable
                mask = area(mask.float())
        self.height, self.width = height, width
        mask = (mask > .5) if has_mask else torch.ones_like(distance_map, dtype=torch.bool)
        mask = mask & (distance_map > 1e-5)
        self.register_buffer("mask_raw", mask.clone())

        lap = laplacian3(distance_map[None].permute(0, 3, 1, 2))
        edge_mask = (lap.abs() < 0.01).float()
        ones = torch.ones(3, 3, device=device)
        edge_mask = dilation(erosion(edge_mask, ones), ones)
        mask = mask & (edge_mask[0] > .5).permute(1, 2, 0)
        if has_normal_map:
            normal_cos = (-pano_dirs(height, width, device) * normal_map).sum(-1, True).clip(0., 1.)
            mask = mask & (normal_cos > 0.15)

        self.register_buffer("color_map", color_map)
        self.register_buffer("distance_map", distance_map)
        self.register_buffer("normal_map", normal_map)
        self.register_buffer("mask", mask)
        self.update_sup_info()

    def update_sup_info(self):                                            # `sup_info.py:95-117`
        height, width, device = self.height, self.width, self.color_map.device
        dirs = apply_rot(pano_dirs(height, width, device), self.pose[:3, :3])
        positions = self.pose[None, None, :3, 3].repeat(height, width, 1)
        sup_indices = torch.where(self.mask[..., 0] > 0.5)
        self.register_buffer("sup_colors", self.color_map[sup_indices])
        self.register_buffer("sup_distances", self.distance_map[sup_indices])
        self.register_buffer("sup_normals", self.normal_map[sup_indices])
        self.register_buffer("sup_dirs", dirs[sup_indices])
        self.register_buffer("sup_positions", positions[sup_indices])
        self.register_buffer("sup_keys", morton_key(sup_indices[0], sup_indices[1]), persistent=False)
        self.sup_rays = Rays(self.sup_positions, self.sup_dirs)

    def set_after_reload(self):
        self.sup_rays = Rays(self.sup_positions, self.sup_dirs)


# ------------------------------------------------------------------ the pool
class SupInfoPool(RaySupervision):
    """`sup_info.py:150-360` on top of ``RaySupervision`` (which holds the flat ray pool, the seeded
    generator, the locality ordering and ``gen_occ_grid``)."""

    def __init__(self, seed: int = 0, locality_sort: bool = True):
        self.sup_infos = []
        self.all_sup_colors = self.all_sup_rays = self.all_sup_distances = self.all_sup_normals = None
        self.locality_key, self.locality_sort = None, locality_sort
        self.generator, self.seed = None, seed
        self.use_default_generator = False

    def _rebuild(self):
        infos = self.sup_infos
        self.all_sup_colors = torch.cat([i.sup_colors for i in infos], 0)
        self.all_sup_rays = Rays(torch.cat([i.sup_rays.o for i in infos], 0), torch.cat([i.sup_rays.d for i in infos], 0))
        self.all_sup_distances = torch.cat([i.sup_distances for i in infos], 0)
        self.all_sup_normals = torch.cat([i.sup_normals for i in infos], 0)
        # locality key: panorama index in the high bits (an int64 only when there are several panoramas)
        if self.locality_sort:
            if len(infos) == 1:
                self.locality_key = infos[0].sup_keys
            else:
                self.locality_key = torch.cat([i.sup_keys.to(torch.int64) + (n << 30) for n, i in enumerate(infos)], 0)
        if self.generator is None:
            from . import parallel
            self.generator = torch.Generator(device=self.all_sup_colors.device).manual_seed(self.seed + parallel.rank())

    def register_sup_info(self, pose, mask, rgb, distance, normal=None):
        self.sup_infos.append(PanoSupInfo(pose=pose, mask=mask, color_map=rgb, distance_map=distance, normal_map=normal))
        self._rebuild()

    def register_sup_info_by_pts(self, pose, colors, pts):
        raise NotImplementedError("register_sup_info_by_pts (`sup_info.py:173-233`) is never called by the reference's runner "
                                  "and is not restated")

    def rand_ray_color_data(self, batch_size, pano_idx=-1, rand_mode="by_all_pixels"):
        assert rand_mode in ["by_all_pixels", "only_first", "only_last"]
        if rand_mode == "by_all_pixels":
            return super().rand_ray_color_data(batch_size)
        info = self.sup_infos[0] if rand_mode == "only_first" else self.sup_infos[-1]
        gen = None if self.use_default_generator else self.generator
        idx = torch.randint(0, len(info.sup_colors), (batch_size,), device=info.sup_colors.device, generator=gen)
        if self.locality_sort:
            idx = idx[torch.argsort(info.sup_keys[idx])]
        return info.sup_rays[idx], info.sup_colors[idx], info.sup_distances[idx], info.sup_normals[idx]

    @torch.no_grad()
    def _projected_distances(self, info: PanoSupInfo, pts: torch.Tensor):
        """Distance of ``pts`` [H,W,3] from panorama ``info``'s centre and the masked distance map bilinearly
        sampled in their direction (`sup_info.py:273-282`, `nerf.py:329-337`)."""
        sup_distance_map = info.distance_map * info.mask.float()
        new_dirs = apply_rot(pts - info.pose[:3, 3], info.pose[:3, :3].T)
        new_distances = torch.linalg.norm(new_dirs, 2, -1, True)
        new_dirs = new_dirs / new_distances
        sample_coords = img_coord_to_sample_coord(direction_to_img_coord(new_dirs))
        proj = F.grid_sample(sup_distance_map[None].permute(0, 3, 1, 2), sample_coords[None], padding_mode="border",
                             align_corners=False)
        return new_distances, proj[0].permute(1, 2, 0)

    @torch.no_grad()
    def geo_check(self, rays: Rays, distances: torch.Tensor) -> torch.Tensor:
        """`sup_info.py:261-302`: 1 = consistent with every registered panorama, 0 = the point lies in
        front of a surface some panorama saw; then dilate (3x3 ellipse) / erode (9x9 ellipse)."""
        pts = rays.o + rays.d * distances.squeeze()[..., None]
        height, width = pts.shape[:2]
        mask = torch.ones(height, width, 1, device=pts.device)
        for info in self.sup_infos:
            new_distances, proj = self._projected_distances(info, pts)
            mask = torch.minimum(mask, (proj < new_distances).float())
        return _close_mask(mask, 3, 9)

    @torch.no_grad()
    def pano_visibility_mask(self, rays: Rays, distance: torch.Tensor) -> torch.Tensor:
        """`nerf.py:320-358` after the render: 1 = the rendered surface point is seen by some registered
        panorama (within 1/256), dilate (5x5 ellipse) / erode (9x9 ellipse)."""
        pts = rays.o + rays.d * distance.squeeze()[..., None]
        height, width = pts.shape[:2]
        mask = torch.zeros(height, width, 1, device=pts.device)
        for info in self.sup_infos:
            new_distances, proj = self._projected_distances(info, pts)
            mask = torch.maximum(mask, (new_distances < proj + 1 / 256.).float())
        return _close_mask(mask, 5, 9)

    def state_dict(self):
        """`sup_info.py:332-340` -- including its quirk: the height / width keys are the literal strings
        ``'sup_info_{}_height'`` / ``'sup_info_{}_width'`` (never formatted), so they hold the LAST panorama's size."""
        ret = {"n_sup_infos": len(self.sup_infos)}
        for i, info in enumerate(self.sup_infos):
            ret["sup_info_{}_height"] = info.height
            ret["sup_info_{}_width"] = info.width
            ret["sup_info_{}".format(i)] = info.state_dict()
        return ret

    def load_state_dict(self, state_dict):
        """Restores every panorama from its buffers (the reference builds placeholder infos and never
        copies the saved buffers back, `sup_info.py:342-360`; a checkpoint written by it loads here)."""
        self.sup_infos = []
        for i in range(state_dict["n_sup_infos"]):
            sd = state_dict["sup_info_{}".format(i)]
            info = PanoSupInfo.__new__(PanoSupInfo)
            torch.nn.Module.__init__(info)
            for k, v in sd.items():
                info.register_buffer(k, v.clone())
            info.height, info.width = int(sd["color_map"].shape[0]), int(sd["color_map"].shape[1])
            rows, cols = torch.where(info.mask[..., 0] > 0.5)
            info.register_buffer("sup_keys", morton_key(rows, cols), persistent=False)
            info.set_after_reload()
            self.sup_infos.append(info)
        self.generator = None
        self._rebuild()


def _close_mask(mask: torch.Tensor, small: int, large: int) -> torch.Tensor:
    m = (mask[None] > 0.5).float().permute(0, 3, 1, 2)
    m = erosion(dilation(m, ellipse_kernel(small, m.device)), ellipse_kernel(large, m.device))
    return m.permute(0, 2, 3, 1).contiguous().squeeze()
