"""CPU tests of perf_b200/sup_info.py.

* the three kornia 0.7.0 functions restated there are pinned against OpenCV (an independent
  implementation of the same filters);
* ``PanoSupInfo`` / ``SupInfoPool`` are compared with the reference's OWN `modules/dataset/sup_info.py`,
  imported unmodified with ``kornia`` bound to those restatements (build container only);
* a committed fixture (``tests/golden/sup_info.npz``, minted by that same import) travels to machines
  without /root/reference.
"""
import os
import sys
import types

import cv2
import numpy as np
import pytest
import torch

from perf_b200 import sup_info as S
from perf_b200.synthetic import box_room_distance, smooth_rgb

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sup_info.npz")


# ------------------------------------------------------------------ kornia restatements vs OpenCV
def test_laplacian_matches_opencv():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 1, 37, 53, generator=g)
    k = np.ones((3, 3), np.float32); k[1, 1] = -8.0; k /= 16.0
    want = cv2.filter2D(x[0, 0].numpy(), cv2.CV_32F, k, borderType=cv2.BORDER_REFLECT_101)
    np.testing.assert_allclose(S.laplacian3(x)[0, 0].numpy(), want, atol=1e-6)


@pytest.mark.parametrize("size", [3, 5, 9])
def test_morphology_and_ellipse_match_opencv(size):
    g = torch.Generator().manual_seed(size)
    x = (torch.rand(1, 1, 41, 67, generator=g) > 0.6).float()
    ell = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (size, size))
    assert np.array_equal(S.ellipse_kernel(size).numpy().astype(np.uint8), ell)
    for kern in (np.ones((size, size), np.uint8), ell):
        kt = torch.from_numpy(kern).float()
        # OpenCV's default border for erode / dilate is +inf / -inf: the border never wins (kornia 'geodesic')
        np.testing.assert_array_equal(S.erosion(x, kt)[0, 0].numpy(), cv2.erode(x[0, 0].numpy(), kern))
        np.testing.assert_array_equal(S.dilation(x, kt)[0, 0].numpy(), cv2.dilate(x[0, 0].numpy(), kern))
    # an asymmetric element: kornia's dilation reflects it (the set-theoretic definition), OpenCV's does not.
    # (The reference only passes symmetric elements: ones(3,3) and OpenCV ellipses.)
    asym = np.array([[1, 1, 0], [0, 1, 0], [0, 0, 0]], np.uint8)
    np.testing.assert_array_equal(S.dilation(x, torch.from_numpy(asym).float())[0, 0].numpy(),
                                  cv2.dilate(x[0, 0].numpy(), np.ascontiguousarray(asym[::-1, ::-1])))
    np.testing.assert_array_equal(S.erosion(x, torch.from_numpy(asym).float())[0, 0].numpy(), cv2.erode(x[0, 0].numpy(), asym))


# ------------------------------------------------------------------ scene shared by the comparisons
def _scene(h=48, w=96):
    rgb, dist = smooth_rgb(h, w, seed=2), box_room_distance(h, w)
    g = torch.Generator().manual_seed(4)
    dist = dist.reshape(h, w).clone()
    dist[10:20, 30:50] *= 0.6                                  # a depth edge the Laplacian mask must remove
    mask = torch.ones(h, w)
    mask[30:40, 5:25] = 0.0
    dirs = S.pano_dirs(h, w, "cpu")
    normal = -dirs + 0.3 * torch.randn(h, w, 3, generator=g)   # mostly facing the camera, some grazing
    normal = normal / normal.norm(dim=-1, keepdim=True)
    a = 0.4
    pose2 = torch.tensor([[np.cos(a), -np.sin(a), 0, 0.1], [np.sin(a), np.cos(a), 0, -0.05], [0, 0, 1, 0.02], [0, 0, 0, 1]], dtype=torch.float32)
    return rgb.reshape(h, w, 3), dist, mask, normal, pose2


def _pool(mod_pool_cls, rgb, dist, mask, normal, pose2):
    pool = mod_pool_cls()
    pool.register_sup_info(pose=torch.eye(4), mask=mask, rgb=rgb, distance=dist, normal=normal)
    pool.register_sup_info(pose=pose2, mask=mask.flip(1), rgb=rgb.flip(0), distance=dist * 0.9, normal=None)
    return pool


def _probe(pool, h, w, pose):
    """geo_check of a novel panorama's rays at a synthetic distance."""
    from perf_b200.scene import Rays
    dirs = S.apply_rot(S.pano_dirs(h, w, "cpu"), pose[:3, :3])
    rays = Rays(pose[None, None, :3, 3].repeat(h, w, 1), dirs)
    g = torch.Generator().manual_seed(9)
    distances = 0.2 + 0.5 * torch.rand(h, w, generator=g)
    return rays, distances


@pytest.fixture(scope="module")
def reference_sup_info():
    if not os.path.isdir(REF):
        pytest.skip("reference checkout not present on this machine")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    kornia = types.ModuleType("kornia"); filters = types.ModuleType("kornia.filters"); morph = types.ModuleType("kornia.morphology")
    filters.laplacian = lambda x, kernel_size: S.laplacian3(x) if kernel_size == 3 else None
    morph.erosion = lambda x, kernel: S.erosion(x, kernel)
    morph.dilation = lambda x, kernel: S.dilation(x, kernel)
    kornia.filters, kornia.morphology = filters, morph
    trimesh = types.ModuleType("trimesh"); creation = types.ModuleType("trimesh.creation")
    creation.icosphere = lambda *a, **k: None
    trimesh.creation = creation
    icecream = types.ModuleType("icecream"); icecream.ic = print
    imageio = types.ModuleType("imageio")
    for name, m in {"kornia": kornia, "kornia.filters": filters, "kornia.morphology": morph, "trimesh": trimesh,
                    "trimesh.creation": creation, "icecream": icecream}.items():
        sys.modules[name] = m
    sys.modules.setdefault("imageio", imageio)
    sys.path.insert(1, REF)
    from modules.dataset import sup_info as ref
    yield ref
    sys.path[:] = saved_path
    for k in [k for k in sys.modules if k not in saved_mods]:
        del sys.modules[k]
    for k, v in saved_mods.items():
        sys.modules[k] = v


def test_pool_matches_reference_file(reference_sup_info):
    ref = reference_sup_info
    scene = _scene()
    want, got = _pool(ref.SupInfoPool, *scene), _pool(lambda: S.SupInfoPool(locality_sort=False), *scene)
    for a, b in zip(want.sup_infos, got.sup_infos):
        for name in ("mask_raw", "mask", "color_map", "distance_map", "normal_map", "sup_colors", "sup_distances",
                     "sup_normals", "sup_dirs", "sup_positions", "pose"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert 0.3 < got.sup_infos[0].mask.float().mean() < 0.95          # the masks do remove something
    assert torch.equal(want.all_sup_colors, got.all_sup_colors) and torch.equal(want.all_sup_rays.d, got.all_sup_rays.d)
    assert torch.equal(want.all_sup_distances, got.all_sup_distances) and torch.equal(want.all_sup_normals, got.all_sup_normals)
    # geo_check on a novel view
    rays, distances = _probe(got, 40, 80, scene[4])
    m_want = want.geo_check(ref.Rays(rays.o, rays.d), distances)
    m_got = got.geo_check(rays, distances)
    assert torch.equal(m_want, m_got) and 0.02 < m_got.mean() < 0.98
    # occupancy pre-grid
    g_want, p_want = want.gen_occ_grid(32)
    g_got, p_got = got.gen_occ_grid(32)
    assert torch.equal(g_want, g_got) and torch.equal(p_want, p_got)
    # batch sampler: the same draw from the same generator state
    torch.manual_seed(5)
    r_w, c_w, d_w, n_w = want.rand_ray_color_data(64)
    got.use_default_generator = True
    torch.manual_seed(5)
    r_g, c_g, d_g, n_g = got.rand_ray_color_data(64)
    assert torch.equal(r_w.d, r_g.d) and torch.equal(c_w, c_g) and torch.equal(d_w, d_g) and torch.equal(n_w, n_g)
    for mode in ("only_first", "only_last"):
        torch.manual_seed(6); a = want.rand_ray_color_data(32, rand_mode=mode)
        torch.manual_seed(6); b = got.rand_ray_color_data(32, rand_mode=mode)
        assert torch.equal(a[1], b[1]) and torch.equal(a[0].d, b[0].d)
    # checkpoint keys (incl. the reference's unformatted height / width keys)
    assert set(want.state_dict()) == set(got.state_dict())
    assert set(want.state_dict()["sup_info_0"]) == set(got.state_dict()["sup_info_0"])


def _golden_payload(pool, scene):
    rays, distances = _probe(pool, 40, 80, scene[4])
    occ, _ = pool.gen_occ_grid(32)
    return {"mask0": pool.sup_infos[0].mask.numpy(), "mask1": pool.sup_infos[1].mask.numpy(),
            "n_sup": np.array([len(i.sup_colors) for i in pool.sup_infos]),
            "color_sum": pool.all_sup_colors.double().sum(0).numpy(), "dir_sum": pool.all_sup_rays.d.double().sum(0).numpy(),
            "geo_check": pool.geo_check(rays, distances).numpy(), "occ_idx": torch.where(occ > 0)[0].numpy()}


def test_pool_matches_golden_fixture():
    """The fixture was written from the reference's own file (``python tests/test_sup_info.py``)."""
    scene = _scene()
    got = _golden_payload(_pool(lambda: S.SupInfoPool(locality_sort=False), *scene), scene)
    want = np.load(GOLDEN)
    for k in ("mask0", "mask1", "n_sup", "geo_check", "occ_idx"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    for k in ("color_sum", "dir_sum"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, err_msg=k)


def test_locality_sort_keeps_the_multiset_and_checkpoint_roundtrip():
    scene = _scene()
    plain, sorted_ = _pool(lambda: S.SupInfoPool(locality_sort=False), *scene), _pool(lambda: S.SupInfoPool(locality_sort=True), *scene)
    plain.use_default_generator = sorted_.use_default_generator = True
    torch.manual_seed(11); a = plain.rand_ray_color_data(256)
    torch.manual_seed(11); b = sorted_.rand_ray_color_data(256)
    key = lambda t: t[torch.argsort(t[:, 0] * 7 + t[:, 1] * 3 + t[:, 2])]
    assert torch.equal(key(a[1]), key(b[1])) and not torch.equal(a[1], b[1])
    restored = S.SupInfoPool(locality_sort=True)
    restored.load_state_dict(sorted_.state_dict())
    assert torch.equal(restored.all_sup_colors, sorted_.all_sup_colors) and torch.equal(restored.locality_key, sorted_.locality_key)
    assert torch.equal(restored.all_sup_rays.d, sorted_.all_sup_rays.d)
    # visibility: a point on the registered surface is visible from its own panorama
    info = sorted_.sup_infos[0]
    from perf_b200.scene import Rays
    rays = Rays(torch.zeros(info.height, info.width, 3), S.pano_dirs(info.height, info.width, "cpu"))
    vis = sorted_.pano_visibility_mask(rays, info.distance_map)
    assert vis[info.mask[..., 0]].mean() > 0.9
    far_vis = sorted_.pano_visibility_mask(rays, info.distance_map * 1.5)          # behind the surface: hidden
    assert far_vis.mean() < 0.2


if __name__ == "__main__":        # mint tests/golden/sup_info.npz from the reference's own file
    gen = reference_sup_info.__wrapped__()
    ref = next(gen)
    sc = _scene()
    np.savez_compressed(GOLDEN, **_golden_payload(_pool(ref.SupInfoPool, *sc), sc))
    print("wrote", GOLDEN)


def test_factor_downsampling_is_opencv_inter_area():
    """`PanoSupInfo(factor=2)` (`sup_info.py:54-66`): the reference resizes with cv.INTER_AREA; for an integer factor
    that is the box average torch's 'area' mode computes."""
    rgb, dist, mask, normal, _ = _scene(32, 64)
    info = S.PanoSupInfo(torch.eye(4), mask, rgb, dist, normal, factor=2)
    assert (info.height, info.width) == (16, 32)
    want = cv2.resize(rgb.numpy(), (32, 16), interpolation=cv2.INTER_AREA)
    np.testing.assert_allclose(info.color_map.numpy(), want, atol=1e-6)
    want_d = cv2.resize(dist.numpy(), (32, 16), interpolation=cv2.INTER_AREA)
    np.testing.assert_allclose(info.distance_map[..., 0].numpy(), want_d, atol=1e-6)
    assert info.mask.shape == (16, 32, 1) and len(info.sup_colors) == int(info.mask.sum())
