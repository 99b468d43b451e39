"""Camera-pose generators of the callers of the render path.

Mirrors `/root/reference/modules/pose_sampler/` (host-side numpy/scipy set-up code that decides WHERE
``NeRFScene.render`` is asked to look; it runs once per scene, not per ray):

  ``CirclePoseSampler``       `circle_pose_sampler.py:45-120`   anchor positions on shrunken copies of the
                              room's horizontal outline (training view points of `core_exp_runner.py:130-133`)
  ``DenseTravelPoseSampler``  `dense_travel_pose_sampler.py:51-116`  a smooth tour through the anchors
                              (the 180 frames of ``render_dense``, `core_exp_runner.py:223-227`)

Same constructor arguments, attributes (``anchor_pts``, ``traverse_pts``, ``traverse_normals``,
``n_anchors``, ``n_poses``, ``sample_poses``) and ``sample_pose(idx)``; tensors live on ``device``
(the reference hard-codes ``.cuda()``).  Compared with the reference's own files in
``tests/test_runner_host.py``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy.ndimage import gaussian_filter1d, minimum_filter1d


class PoseSampler:                                                           # pose_sampler.py:10-16
    def __init__(self):
        self.n_poses = 0

    @torch.no_grad()
    def sample_pose(self, idx):
        raise NotImplementedError


def _pano_direction(coords: torch.Tensor) -> torch.Tensor:
    """``img_coord_to_pano_direction`` (`utils/camera_utils.py:120-155`): (row, col) in [0,1] -> unit direction."""
    beta, alpha = -(coords[..., 0] - .5) * np.pi, -(coords[..., 1] - .5) * 2. * np.pi
    return torch.stack([torch.cos(alpha) * torch.cos(beta), torch.sin(alpha) * torch.cos(beta), torch.sin(beta)], dim=-1)


def look_at(to_vec: torch.Tensor, up_vec=None) -> torch.Tensor:
    """`utils/camera_utils.py:83-101`: rotation matrices [n,3,3] with columns (right, down, to)."""
    n = to_vec.shape[0]
    if up_vec is None:
        up_vec = torch.cat([torch.zeros(n, 2), torch.ones(n, 1)], -1).to(to_vec)
    down_vec = -up_vec
    to_vec = to_vec / torch.linalg.norm(to_vec, 2, -1, True)
    ri_vec = torch.linalg.cross(down_vec, to_vec)
    ri_vec = ri_vec / torch.linalg.norm(ri_vec, 2, -1, True)
    down_vec = torch.linalg.cross(to_vec, ri_vec)
    return torch.stack([ri_vec, down_vec, to_vec], -1)


@torch.no_grad()
def _resample_uniformly(pts: torch.Tensor) -> torch.Tensor:                   # circle_pose_sampler.py:13-23
    n = len(pts)
    pts = F.interpolate(pts[None].permute(0, 2, 1), size=n * 128, mode="linear")[0].permute(1, 0)
    cat_pts = torch.cat([pts, pts[:1]], dim=0)
    bias_len = torch.cumsum(torch.linalg.norm(cat_pts[1:] - cat_pts[:-1], 2, -1), dim=0)
    bias_len = bias_len / bias_len[-1]
    idx = torch.searchsorted(bias_len, torch.linspace(0., 1. - 1. / n, n).to(bias_len))
    return pts[idx]


@torch.no_grad()
def _get_trajectory_normals(pts: torch.Tensor) -> torch.Tensor:               # circle_pose_sampler.py:26-42
    sigma = float(len(pts)) / 32. * 2. + 1.
    ext_pts = torch.cat([pts, pts[:1]], dim=0)
    right_vec = ext_pts[1:] - ext_pts[:-1]
    right_vec = right_vec / torch.linalg.norm(right_vec, 2, -1, True)
    up_vec = torch.zeros_like(right_vec)
    up_vec[:, 2] = 1
    to_vec = torch.linalg.cross(up_vec, right_vec)
    to_vec = (to_vec / torch.linalg.norm(to_vec, 2, -1, True)).cpu().numpy()
    for i in range(3):
        to_vec[:, i] = gaussian_filter1d(to_vec[:, i], sigma=sigma, mode="wrap")
    to_vec = torch.from_numpy(to_vec).to(pts.device)
    return -(to_vec / torch.linalg.norm(to_vec, 2, -1, True))


class CirclePoseSampler(PoseSampler):
    def __init__(self, distance_map, traverse_ratios, n_anchors_per_ratio, test_z_min_max=(0., 0.), device="cpu", **kwargs):
        super().__init__()
        if torch.is_tensor(distance_map):
            distance_map = distance_map.cpu().numpy()
        distance_map = distance_map.squeeze()
        height, width = distance_map.shape
        beta = (-(torch.linspace(.5 / height, 1. - .5 / height, height) - .5) * np.pi).numpy()   # img_to_pano_coord rows

        # horizontal distance to the walls, from the 20 rows around the horizon
        plane_dis = distance_map * np.cos(beta)[:, None]
        h_height = height // 2
        plane_dis = plane_dis[h_height - 10: h_height + 10]
        plane_dis[np.where(plane_dis < 1e-5)] = 1e9
        plane_dis = np.min(plane_dis, axis=0)
        for i in range(1, width):                                               # fill holes from the left, then from the right
            if plane_dis[i] > 1e8:
                plane_dis[i] = plane_dis[i - 1]
        for i in range(1, width):
            if plane_dis[width - i - 1] > 1e8:
                plane_dis[width - i - 1] = plane_dis[width - i]

        pool_size = (width // 16) // 2 * 2 + 1
        filtered_plane_dis = minimum_filter1d(plane_dis, size=pool_size, mode="wrap")
        smooth_size = (width // 8) // 2 * 2 + 1
        smoothed_plane_dis = gaussian_filter1d(filtered_plane_dis, sigma=smooth_size, mode="wrap")
        blur_size = (width // 64) // 2 * 2 + 1
        filtered_plane_dis = gaussian_filter1d(filtered_plane_dis, sigma=blur_size, mode="wrap")
        plane_coords = torch.stack([torch.ones(width) * .5, torch.linspace(.5 / width, 1. - .5 / width, width)], -1)
        circle_pts = _pano_direction(plane_coords)
        plane_pts = circle_pts.numpy()
        self.plane_pts_raw = torch.from_numpy(plane_pts * plane_dis[:, None]).to(device)
        self.plane_pts_filter = torch.from_numpy(plane_pts * filtered_plane_dis[:, None]).to(device)
        self.plane_pts_smooth = torch.from_numpy(plane_pts * smoothed_plane_dis[:, None]).to(device)
        filtered_plane_dis = torch.from_numpy(filtered_plane_dis).to(device)
        smoothed_plane_dis = torch.from_numpy(smoothed_plane_dis).to(device)
        circle_pts = circle_pts.to(device)

        anchor_pts = []
        test_z_min, test_z_max = test_z_min_max
        for i, traverse_ratio in enumerate(traverse_ratios):
            traverse_pts = _resample_uniformly(circle_pts * filtered_plane_dis[:, None] * traverse_ratio)
            n = n_anchors_per_ratio[i]
            bias = 0. if i % 2 == 0 else .5 / n
            anchor_idx = torch.linspace(.5 / n, 1. - .5 / n, n) + bias
            anchor_idx = (anchor_idx * width).to(torch.long).clip(0, width - 1).to(device)
            cur_pts = traverse_pts[anchor_idx].clone()
            for j in range(len(cur_pts)):
                cur_pts[j, 2] = test_z_min if (i + j) % 2 == 0 else test_z_max
            anchor_pts.append(cur_pts)

        self.anchor_pts = torch.cat(anchor_pts, dim=0)
        self.traverse_pts = _resample_uniformly(circle_pts * smoothed_plane_dis[:, None] * .3)
        self.traverse_normals = _get_trajectory_normals(self.traverse_pts)
        self.n_anchors = len(self.anchor_pts)
        self.n_poses = self.n_anchors

    @torch.no_grad()
    def sample_pose(self, idx):
        pose = torch.eye(4, device=self.anchor_pts.device, dtype=self.anchor_pts.dtype)
        pose[:3, 3] = self.anchor_pts[idx]
        return pose


def _get_travel_indices(positions: torch.Tensor, n_steps: int = 10000) -> torch.Tensor:
    """`dense_travel_pose_sampler.py:28-48`: simulated-annealing ordering of the anchors (pairwise swaps,
    acceptance ratio (1 - t)^5), driven by ``np.random`` exactly as the reference draws it."""
    positions = positions.cpu()
    n = len(positions)
    indices = torch.arange(0, n, dtype=torch.int64)
    dis = 1e8
    for iter_step in range(n_steps):
        a = np.random.randint(n)
        b = np.random.randint(n)
        new_indices = indices.clone()
        new_indices[a] = indices[b]
        new_indices[b] = indices[a]
        shifts = positions[new_indices[:-1]] - positions[new_indices[1:]]
        new_dis = torch.linalg.norm(shifts, 2, -1).sum()
        ratio = (1. - (iter_step / n_steps)) ** 5
        if new_dis < dis or np.random.rand() < ratio:
            indices = new_indices
            dis = new_dis
    return indices


class DenseTravelPoseSampler(PoseSampler):
    def __init__(self, sparse_pose_sampler: PoseSampler, n_dense_poses, dir_bias_ratio=-1):
        super().__init__()
        sparse_poses = torch.stack([sparse_pose_sampler.sample_pose(i) for i in range(sparse_pose_sampler.n_poses)], 0)
        device = sparse_poses.device
        travel_indices = _get_travel_indices(sparse_poses[:, :3, 3]).to(device)
        travel_sparse_poses = sparse_poses[travel_indices]

        N = n_dense_poses * 50
        sparse_pts = travel_sparse_poses[:, :3, 3]
        sec_lens = torch.linalg.norm(sparse_pts[1:] - sparse_pts[:-1], 2, -1, True)
        sec_n_poses = torch.round(N * (sec_lens / sec_lens.sum())).to(torch.int64)
        pts = []
        for i in range(len(sec_n_poses)):
            a, b = sparse_pts[i], sparse_pts[i + 1]
            cur_n = sec_n_poses[i].item()
            t = torch.linspace(.5 / cur_n, 1. - .5 / cur_n, cur_n).to(a)
            pts.append(a[None, :] * (1. - t)[:, None] + b[None, :] * t[:, None])
        pts = _resample_uniformly(torch.cat(pts, 0))[::50]
        pts = pts.cpu().numpy()
        for i in range(3):
            pts[:, i] = gaussian_filter1d(pts[:, i], sigma=20)
        pts = torch.from_numpy(pts).to(device)
        self.sample_poses = torch.eye(4, device=device, dtype=pts.dtype)[None].repeat(len(pts), 1, 1)
        self.sample_poses[:, :3, 3] = pts
        self.n_poses = len(self.sample_poses)

        to_vecs = pts.clone()
        to_vecs[:-1] = pts[1:] - pts[:-1]
        to_vecs[-1] = to_vecs[-2]
        for i in range(3):
            to_vecs[:, i] = torch.from_numpy(gaussian_filter1d(to_vecs[:, i].cpu().numpy(), sigma=30)).to(device)
        to_vecs = to_vecs / torch.linalg.norm(to_vecs, 2, -1, True)
        up_vecs = torch.zeros_like(to_vecs)
        up_vecs[..., 2] = 1
        left_vecs = torch.linalg.cross(up_vecs, to_vecs)
        left_vecs = left_vecs / torch.linalg.norm(left_vecs, 2, -1, True)
        to_vecs = to_vecs / torch.linalg.norm(to_vecs, 2, -1, True)
        to_vecs = to_vecs + dir_bias_ratio * left_vecs
        to_vecs = to_vecs / torch.linalg.norm(to_vecs, 2, -1, True)
        self.sample_poses[:, :3, :3] = look_at(to_vecs)

    @torch.no_grad()
    def sample_pose(self, idx):
        return self.sample_poses[idx]
