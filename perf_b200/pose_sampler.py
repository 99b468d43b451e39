"""Camera-pose generators for the callers of the render path (host-side set-up, once per scene).

Behavioural mirrors of `/root/reference/modules/pose_sampler/`:

  ``CirclePoseSampler``       `circle_pose_sampler.py:45-120`  -- view points for the training phases
                              (`core_exp_runner.py:130-133`): the room's horizontal outline is read off the
                              horizon band of the distance panorama, shrunk by each ``traverse_ratio`` and sampled at
                              ``n_anchors_per_ratio`` arc-length-uniform stations.
  ``DenseTravelPoseSampler``  `dense_travel_pose_sampler.py:51-116` -- the smooth tour through those anchors that
                              ``render_dense`` renders (`core_exp_runner.py:223-227`).

Constructor arguments, public attributes (``anchor_pts``, ``traverse_pts``, ``traverse_normals``, ``plane_pts_*``,
``n_anchors``, ``n_poses``, ``sample_poses``) and ``sample_pose(idx)`` are the reference's; results are bit-equal to
its files (``tests/test_runner_host.py`` imports them unmodified), which pins the floating-point recipe of every
step below.  Tensors live on ``device`` (the reference hard-codes ``.cuda()``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage

_OVERSAMPLE = 128          # closed curves are re-parameterised on a 128x finer polyline
_HORIZON_HALF_BAND = 10    # rows above / below the horizon that vote for the wall distance


class PoseSampler:
    """Interface of `pose_sampler.py:10-16`: ``n_poses`` and ``sample_pose(idx) -> [4,4]`` camera-to-world."""

    n_poses = 0

    def sample_pose(self, idx):
        raise NotImplementedError


# ------------------------------------------------------------------ geometry helpers
def _unit(v: torch.Tensor) -> torch.Tensor:
    return v / torch.linalg.norm(v, 2, -1, True)


def _equator_directions(width: int) -> torch.Tensor:
    """Unit directions of the panorama's middle row (`camera_utils.py:120-155` at row coordinate 0.5)."""
    col = torch.linspace(.5 / width, 1. - .5 / width, width)
    row = torch.ones(width) * .5
    beta, alpha = -(row - .5) * np.pi, -(col - .5) * 2. * np.pi
    return torch.stack([torch.cos(alpha) * torch.cos(beta), torch.sin(alpha) * torch.cos(beta), torch.sin(beta)], dim=-1)


def look_at(forward: torch.Tensor, up=None) -> torch.Tensor:
    """Rotations [n,3,3] whose columns are (right, down, forward), z-up world (`camera_utils.py:83-98`)."""
    if up is None:
        up = torch.zeros_like(forward)
        up[..., 2] = 1.0
    forward = _unit(forward)
    right = _unit(torch.linalg.cross(-up, forward))
    return torch.stack([right, torch.linalg.cross(forward, right), forward], -1)


@torch.no_grad()
def arc_length_resample(curve: torch.Tensor) -> torch.Tensor:
    """n points evenly spaced by arc length along the CLOSED polyline ``curve`` [n,3]: linear up-sampling by
    128, cumulative chord length (including the closing chord), then the first fine vertex at or beyond each
    of the n target fractions k/n."""
    n = curve.shape[0]
    fine = F.interpolate(curve.t()[None], size=n * _OVERSAMPLE, mode="linear")[0].t()
    chords = torch.linalg.norm(torch.roll(fine, -1, 0) - fine, 2, -1)
    travelled = torch.cumsum(chords, dim=0)
    travelled = travelled / travelled[-1]
    targets = torch.linspace(0., 1. - 1. / n, n).to(travelled)
    return fine[torch.searchsorted(travelled, targets)]


@torch.no_grad()
def inward_normals(loop: torch.Tensor) -> torch.Tensor:
    """Horizontal unit normals of a closed loop [n,3], smoothed along the loop (periodic Gaussian,
    sigma = n/16 + 1) and pointing to the side the reference calls ``-to_vec``."""
    tangent = _unit(torch.roll(loop, -1, 0) - loop)
    z_axis = torch.zeros_like(tangent)
    z_axis[:, 2] = 1
    side = _unit(torch.linalg.cross(z_axis, tangent)).cpu().numpy()
    sigma = float(loop.shape[0]) / 32. * 2. + 1.
    for axis in range(3):
        side[:, axis] = ndimage.gaussian_filter1d(side[:, axis], sigma=sigma, mode="wrap")
    return -_unit(torch.from_numpy(side).to(loop.device))


def _fill_from_neighbours(profile: np.ndarray, hole: float = 1e8) -> np.ndarray:
    """Columns marked as holes take the nearest valid value on their left; holes still open at the left end
    then take the nearest value on their right (two sweeps, as the reference loops do; column 0 is only
    reachable by the second sweep)."""
    out = profile.copy()
    cols = np.arange(out.shape[0])
    src = np.where(out > hole, -1, cols)
    src[0] = 0                                             # the left sweep never rewrites column 0
    out = out[np.maximum.accumulate(src)]
    src = np.where(out > hole, out.shape[0], cols)
    src[-1] = out.shape[0] - 1
    return out[np.minimum.accumulate(src[::-1])[::-1]]


def _odd(n: int) -> int:
    return n // 2 * 2 + 1


# ------------------------------------------------------------------ anchors on shrunken outlines
class CirclePoseSampler(PoseSampler):
    def __init__(self, distance_map, traverse_ratios, n_anchors_per_ratio, test_z_min_max=(0., 0.), device="cpu", **kwargs):
        if torch.is_tensor(distance_map):
            distance_map = distance_map.cpu().numpy()
        distance_map = distance_map.squeeze()
        height, width = distance_map.shape
        raw, sharp, smooth = self._wall_profiles(distance_map)
        ring = _equator_directions(width)
        for name, prof in (("plane_pts_raw", raw), ("plane_pts_filter", sharp), ("plane_pts_smooth", smooth)):
            setattr(self, name, torch.from_numpy(ring.numpy() * prof[:, None]).to(device))
        ring = ring.to(device)
        sharp_t, smooth_t = torch.from_numpy(sharp).to(device), torch.from_numpy(smooth).to(device)

        z_low, z_high = test_z_min_max
        groups = []
        for level, (ratio, count) in enumerate(zip(traverse_ratios, n_anchors_per_ratio)):
            outline = arc_length_resample(ring * sharp_t[:, None] * ratio)
            # stations at the centres of `count` equal arcs, every other outline shifted by half an arc
            stations = torch.linspace(.5 / count, 1. - .5 / count, count) + (0. if level % 2 == 0 else .5 / count)
            picks = (stations * width).to(torch.long).clip(0, width - 1).to(device)
            pts = outline[picks].clone()
            heights = [z_low if (level + j) % 2 == 0 else z_high for j in range(count)]
            pts[:, 2] = torch.tensor(heights, dtype=pts.dtype, device=pts.device)
            groups.append(pts)
        self.anchor_pts = torch.cat(groups, dim=0)
        self.traverse_pts = arc_length_resample(ring * smooth_t[:, None] * .3)
        self.traverse_normals = inward_normals(self.traverse_pts)
        self.n_anchors = self.n_poses = self.anchor_pts.shape[0]

    @staticmethod
    def _wall_profiles(distance_map: np.ndarray):
        """Horizontal wall distance per panorama column: (raw, lightly blurred, heavily smoothed).

        raw = min over the horizon band of distance * cos(elevation), invalid (< 1e-5) pixels ignored and empty
        columns filled from their neighbours; then a periodic minimum filter (width/16) keeps the camera clear
        of thin obstacles; the blurred variant (sigma width/64) places the anchors, the smooth one (sigma
        width/8) carries the preview trajectory."""
        height, width = distance_map.shape
        elevation = (-(torch.linspace(.5 / height, 1. - .5 / height, height) - .5) * np.pi).numpy()
        band = slice(height // 2 - _HORIZON_HALF_BAND, height // 2 + _HORIZON_HALF_BAND)
        horizontal = (distance_map * np.cos(elevation)[:, None])[band]
        horizontal[horizontal < 1e-5] = 1e9
        raw = _fill_from_neighbours(horizontal.min(axis=0))
        eroded = ndimage.minimum_filter1d(raw, size=_odd(width // 16), mode="wrap")
        smooth = ndimage.gaussian_filter1d(eroded, sigma=_odd(width // 8), mode="wrap")
        sharp = ndimage.gaussian_filter1d(eroded, sigma=_odd(width // 64), mode="wrap")
        return raw, sharp, smooth

    @torch.no_grad()
    def sample_pose(self, idx):
        pose = torch.eye(4, device=self.anchor_pts.device, dtype=self.anchor_pts.dtype)
        pose[:3, 3] = self.anchor_pts[idx]
        return pose


# ------------------------------------------------------------------ dense tour through the anchors
def anneal_visiting_order(positions: torch.Tensor, n_steps: int = 10000) -> torch.Tensor:
    """Order in which to visit ``positions`` [n,3] (open path), by the reference's annealing schedule
    (`dense_travel_pose_sampler.py:28-48`): propose swapping two random entries; a shorter path is always
    taken, a longer one with probability (1 - step/n_steps)^5.  ``np.random`` is consumed in the reference's
    order (two ``randint`` per step, one ``rand`` only when the proposal is not shorter) and path lengths are
    summed by torch in fp32, so a seeded run reproduces the reference's tour exactly."""
    positions = positions.cpu()
    n = positions.shape[0]
    order = torch.arange(n, dtype=torch.int64)
    best = 1e8
    for step in range(n_steps):
        i, j = np.random.randint(n), np.random.randint(n)
        trial = order.clone()
        trial[i], trial[j] = order[j], order[i]
        length = torch.linalg.norm(positions[trial[:-1]] - positions[trial[1:]], 2, -1).sum()
        if length < best or np.random.rand() < (1. - step / n_steps) ** 5:
            order, best = trial, length
    return order


class DenseTravelPoseSampler(PoseSampler):
    _PER_POSE = 50          # polyline vertices generated per output pose before decimation

    def __init__(self, sparse_pose_sampler: PoseSampler, n_dense_poses, dir_bias_ratio=-1):
        anchors = torch.stack([sparse_pose_sampler.sample_pose(i) for i in range(sparse_pose_sampler.n_poses)], 0)
        device = anchors.device
        stops = anchors[anneal_visiting_order(anchors[:, :3, 3]).to(device), :3, 3]

        # polyline through the stops, each leg subdivided in proportion to its length (cell-centred parameters)
        legs = stops[1:] - stops[:-1]
        leg_len = torch.linalg.norm(legs, 2, -1, True)
        per_leg = torch.round(n_dense_poses * self._PER_POSE * (leg_len / leg_len.sum())).to(torch.int64)
        pieces = []
        for start, end, count in zip(stops[:-1], stops[1:], per_leg.reshape(-1).tolist()):
            s = torch.linspace(.5 / count, 1. - .5 / count, count).to(start)
            pieces.append(start[None, :] * (1. - s)[:, None] + end[None, :] * s[:, None])
        track = arc_length_resample(torch.cat(pieces, 0))[::self._PER_POSE]
        track = torch.from_numpy(self._smooth_columns(track.cpu().numpy(), 20)).to(device)

        self.sample_poses = torch.eye(4, device=device, dtype=track.dtype)[None].repeat(track.shape[0], 1, 1)
        self.sample_poses[:, :3, 3] = track
        self.n_poses = track.shape[0]

        # viewing direction: smoothed direction of travel, swung towards the left by dir_bias_ratio
        heading = track.clone()
        heading[:-1] = track[1:] - track[:-1]
        heading[-1] = heading[-2]
        heading = _unit(torch.from_numpy(self._smooth_columns(heading.cpu().numpy(), 30)).to(device))
        z_axis = torch.zeros_like(heading)
        z_axis[..., 2] = 1
        left = _unit(torch.linalg.cross(z_axis, heading))
        heading = _unit(heading)
        self.sample_poses[:, :3, :3] = look_at(_unit(heading + dir_bias_ratio * left))

    @staticmethod
    def _smooth_columns(values: np.ndarray, sigma: float) -> np.ndarray:
        for axis in range(values.shape[1]):
            values[:, axis] = ndimage.gaussian_filter1d(values[:, axis], sigma=sigma)
        return values

    @torch.no_grad()
    def sample_pose(self, idx):
        return self.sample_poses[idx]
