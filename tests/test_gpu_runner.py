"""GPU test of the runner mirror (`core_exp_runner.py` first phase + render_dense) on a synthetic RGB-D
panorama written in the reference's file formats: WildDataset -> SupInfoPool (masks) -> NeRFScene.fit
-> render -> checkpoint -> render_dense, plus the geometry side of the inpainting loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_case(root, h, w):
    import cv2
    from perf_b200.synthetic import box_room_distance, smooth_rgb
    case = root / "room"
    case.mkdir()
    cv2.imwrite(str(case / "image.png"), (smooth_rgb(h, w, seed=0) * 255).byte().numpy()[:, :, ::-1])
    np.save(case / "image_ref_distance.npy", box_room_distance(h, w).reshape(h, w).numpy() * 2.5)
    from perf_b200.sup_info import pano_dirs
    d = -pano_dirs(h, w, "cpu")
    np.save(case / "image_ref_normal.npy", d.numpy().astype(np.float32))        # normals facing the camera
    return str(case / "image.png")


def test_runner_raw_phase_and_render_dense(tmp_path):
    from perf_b200.runner import CoreRunner
    from perf_b200.scene import gen_pano_rays
    h, w = 64, 128
    conf = {"exp_name": "t", "mode": "train", "is_continue": False, "dataset_class_name": "WildDataset",
            "dataset": {"image_path": _write_case(tmp_path, h, w)}, "device": {"base_exp_dir": str(tmp_path / "exp")},
            "pose_sampler": {"traverse_ratios": [0.2, 0.4], "n_anchors_per_ratio": [4, 4]},
            "scene_class_name": "NeRFScene",
            "scene": {"estimator_type": "fixed", "renderer_conf": {"max_radius": 2, "bg_color": "rand_noise"},
                      "train_conf": {"raw_phase_iter_geo": 150, "raw_phase_iter_app": 100, "pixel_loss_batch_size": 2048,
                                     "geo_optimizer": {"init_lr": 0.0, "peak_lr": 1e-2, "peak_at": 0.2, "lr_alpha": 1e-2},
                                     "app_optimizer": {"init_lr": 0.0, "peak_lr": 1e-2, "peak_at": 0.2, "lr_alpha": 1e-2},
                                     "color_loss_weight": 1., "depth_loss_weight": 1., "distortion_loss_weight": 0.1,
                                     "density_loss_weight": 0.}}}
    torch.manual_seed(0), np.random.seed(0)
    runner = CoreRunner(conf, scene_kwargs={"n_samples": 48})
    assert runner.sup_pool.sup_infos[0].mask.float().mean() > 0.8             # smooth room: few edge pixels masked
    assert runner.pose_sampler.n_anchors == 8
    result = runner.train(raw_only=True)
    exp = runner.exp_dir
    assert exp.endswith(os.path.join("WildDataset_room", "t"))
    for f in ("distance_vis.png", "normal_vis.png", "1.png", "1_distance.png", os.path.join("checkpoints", "ckpt.pth")):
        assert os.path.exists(os.path.join(exp, f)), f
    assert result["rgb"].shape == (512, 1024, 3) and runner.phase == 0
    # the fit learned the room: distance error of the re-rendered input view
    out = runner.scene.render(gen_pano_rays(torch.eye(4), h, w), query_keys=["rgb", "distance"])
    err = float((out["distance"].squeeze() - runner.dataset.ref_distance.squeeze()).abs().mean())
    assert err < 0.05, err
    ck = torch.load(os.path.join(exp, "checkpoints", "ckpt.pth"), weights_only=False)
    assert set(ck) == {"scene", "sup_pool", "phase"} and ck["sup_pool"]["n_sup_infos"] == 1
    # second-phase entry is explicit about what is missing
    with pytest.raises(NotImplementedError, match="inpainting"):
        runner.train()
    # geometry side of the inpainting loop at a sampled anchor pose
    pose = runner.pose_sampler.sample_pose(3)
    rays = gen_pano_rays(pose, h, w)
    visi = runner.scene.get_pano_visibility_mask(runner.sup_pool, rays)
    assert visi.shape == (h, w) and set(visi.unique().tolist()) <= {0.0, 1.0}
    novel = runner.scene.render(rays, query_keys=["rgb", "distance"])
    ok = runner.sup_pool.geo_check(rays, novel["distance"])
    assert ok.shape == (h, w)
    runner.sup_pool.register_sup_info(pose=pose, mask=1. - visi, rgb=novel["rgb"], distance=novel["distance"], normal=None)
    assert len(runner.sup_pool.sup_infos) == 2 and len(runner.sup_pool.all_sup_colors) >= len(runner.sup_pool.sup_infos[0].sup_colors)
    # render_dense: a few frames of the tour, and a restart from the checkpoint renders the same frames
    np.random.seed(1)
    frames = runner.render_dense(n_poses=4, height=32, width=64)
    assert 3 <= len(frames) <= 5 and frames[0].shape == (32, 64, 3) and frames[0].dtype == np.uint8
    assert os.path.exists(os.path.join(exp, "dense_images_new_pano", "image_0.png"))
    conf["is_continue"] = True
    again = CoreRunner(conf, scene_kwargs={"n_samples": 48})
    assert again.phase == 0
    np.random.seed(1)
    frames2 = again.render_dense(n_poses=4, height=32, width=64, write=False)
    # (the dataset's distance map was re-normalised on reload: anchor positions may move by an ulp)
    assert len(frames2) == len(frames) and all(np.abs(a.astype(int) - b.astype(int)).max() <= 2 for a, b in zip(frames, frames2))
    pers = again.render_dense(n_poses=4, cam_type="pers", height=32, write=False)
    assert pers[0].shape == (32, 32, 3)
