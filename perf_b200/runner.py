"""``CoreRunner``: the caller of the render / training path.

Mirrors `/root/reference/core_exp_runner.py:36-256` for everything that does not need the 2-D priors:

  ``CoreRunner(conf)``            `:37-95`    dataset, experiment directory, scene, pose sampler, supervision pool
  ``train(raw_only=True)``        `:106-124`  fit the scene to the input panorama, render `1.png` / `1_distance.png`, checkpoint
  ``render_dense(n_poses, cam)``  `:223-246`  the dense tour (north-star workload), frames row-tiled over ranks
  ``save_checkpoint`` / ``load_checkpoint``   `:217-221,248-256`  same file, same keys (`scene`, `sup_pool`, `phase`)

The 2-D priors of the inpainting loop of ``train`` (`:126-177`: Stable Diffusion / LaMa inpainting, the monocular depth
predictor) are out of scope.  The loop itself -- visibility masks, ``geo_check``, mask arithmetic,
``register_sup_info`` of each completed panorama, re-fit, checkpoint per anchor -- is here and runs when the caller
injects those models (``CoreRunner(conf, inpainter=..., geo_predictor=...)``); without them ``train`` raises after
the raw phase.

    python -m perf_b200.runner --config-dir /path/to/PeRF/configs mode=train dataset.image_path=... [key=value ...]
"""
from __future__ import annotations

import os
from os.path import join as pjoin

import numpy as np
import torch

from . import parallel
from .config import Conf, load_config
from .dataset import WildDataset, colorize_single_channel_image, write_image
from .pose_sampler import CirclePoseSampler, DenseTravelPoseSampler
from .scene import NeRFScene, gen_pano_rays, gen_pers_rays
from .sup_info import SupInfoPool


class CoreRunner:
    def __init__(self, conf, device=None, scene_kwargs=None, inpainter=None, geo_predictor=None):
        self.conf = conf = Conf.wrap(conf)
        # the reference's 2-D priors, injected by the caller (PanoPersFusionInpainter / PanoJointPredictor objects or
        # anything with the same call signatures); without them only the raw phase and render_dense are available
        self.inpainter, self.geo_predictor = inpainter, geo_predictor
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dataset = WildDataset(conf.dataset, device=self.device)
        self.base_exp_dir = conf.device.base_exp_dir
        self.exp_dir = pjoin(self.base_exp_dir, "{}_{}".format(conf["dataset_class_name"], self.dataset.case_name), conf.exp_name)
        self.is_main = parallel.rank() == 0
        if self.is_main:
            os.makedirs(self.exp_dir, exist_ok=True)
        if conf.scene_class_name != "NeRFScene":
            raise NotImplementedError(f"scene_class_name={conf.scene_class_name!r}")
        self.scene = NeRFScene(self.exp_dir, device=self.device, **conf.scene, **(scene_kwargs or {}))
        if self.is_main:                                                          # core_exp_runner.py:65-71
            write_image(pjoin(self.exp_dir, "distance_vis.png"), colorize_single_channel_image(
                (self.dataset.ref_distance.min() + 1e-6) / (self.dataset.ref_distance + 1e-6)))
            if self.dataset.ref_normal is not None:
                write_image(pjoin(self.exp_dir, "normal_vis.png"), (self.dataset.ref_normal * .5 + .5) * 255.)
        self.pose_sampler = CirclePoseSampler(self.dataset.ref_distance, device=self.device, **conf.pose_sampler)
        self.sup_pool = SupInfoPool()
        self.sup_pool.register_sup_info(pose=torch.eye(4, device=self.device),
                                        mask=torch.ones(self.dataset.height, self.dataset.width, device=self.device),
                                        rgb=self.dataset.image, distance=self.dataset.ref_distance, normal=self.dataset.ref_normal)
        self.phase = -1
        if conf.get("is_continue", False):
            self.load_checkpoint("ckpt.pth")

    def set_train(self):
        self.scene.set_train()

    def set_eval(self):
        self.scene.set_eval()

    def execute(self, mode):
        if mode == "train":
            self.train()
        elif mode == "render_dense":
            self.render_dense()
        else:
            raise ValueError(f"mode={mode!r}")

    def train(self, raw_only=False):
        if self.phase < 0:
            self.set_train()
            self.scene.fit(self.sup_pool)
            self.set_eval()
            result = self.scene.render(gen_pano_rays(torch.eye(4), 512, 1024, device=self.device), query_keys=["rgb", "distance"])
            if self.is_main:
                disparity = (result["distance"].min() / result["distance"]).squeeze()[..., None]
                write_image(pjoin(self.exp_dir, "1.png"), result["rgb"] * 255.)
                write_image(pjoin(self.exp_dir, "1_distance.png"), colorize_single_channel_image(disparity))
            self.phase += 1
            self.save_checkpoint()
            if raw_only:
                return result
        if self.inpainter is None or (self.geo_predictor is None and not self.conf.get("rgbd_inpaint", False)):
            raise NotImplementedError(
                "the inpainting phases of CoreRunner.train (core_exp_runner.py:126-177) call the reference's Stable Diffusion / "
                "LaMa inpainter and its monocular depth predictor, which are outside the per-ray path: run train(raw_only=True), "
                "or pass those objects as CoreRunner(conf, inpainter=..., geo_predictor=...)")
        return self._train_inpainting_phases()

    def _train_inpainting_phases(self, geo_check=True):
        """`core_exp_runner.py:126-177`: for every anchor pose not done yet -- render the current scene there, find what no
        registered panorama sees, let the injected priors invent colour and geometry for it, drop what contradicts the
        known geometry, register the completed panorama as new supervision and re-fit."""
        height, width = self.dataset.height, self.dataset.width
        for anchor in range(max(self.phase, 0), self.pose_sampler.n_anchors):
            pose = self.pose_sampler.sample_pose(anchor)
            rays = gen_pano_rays(pose, height, width, device=self.device)
            self.set_eval()
            visible = self.scene.get_pano_visibility_mask(self.sup_pool, rays)             # 1 = seen by a registered panorama
            view = self.scene.render(rays, query_keys=["rgb", "distance"])
            colors, distances, normals = view["rgb"], view["distance"], None
            hole = 1. - visible
            if visible.min().item() <= .5:                                                 # something to invent (n_repeats = 1 upstream)
                colors, distances, normals = self.inpaint_new_panorama(0, anchor, colors=colors, distances=distances, mask=hole)
                if geo_check:
                    hole = hole * (1. - self.sup_pool.geo_check(rays, distances))            # keep only what conflicts with nothing known
                else:
                    hole = hole * 0
            # never trust invented content closer than 0.1 -- and never overwrite what was visible
            hole = torch.minimum(torch.maximum(hole, (distances.squeeze() < 0.1).float()), 1. - visible)
            if self.is_main:
                vis_dir = pjoin(self.exp_dir, "inpaint_vis", "{:0>4d}".format(anchor))
                os.makedirs(vis_dir, exist_ok=True)
                write_image(pjoin(vis_dir, "final_mask.jpg"), hole[..., None] * 255.)
                write_image(pjoin(vis_dir, "final_masked.jpg"), (colors * (1. - hole)[..., None]) * 255.)
            new_sup = (1. - visible) - torch.minimum(1. - visible, hole)                   # invisible before, accepted now
            self.sup_pool.register_sup_info(pose=pose, mask=new_sup, rgb=colors, distance=distances, normal=normals)
            self.set_train()
            self.scene.fit(self.sup_pool)
            self.phase += 1
            self.save_checkpoint()

    def inpaint_new_panorama(self, phase, anchor_idx, colors, distances, mask):
        """`core_exp_runner.py:179-215`: colours from ``inpainter.inpaint`` then geometry from ``geo_predictor`` (aligned to the
        rendered distances outside the mask), or both at once from ``inpainter.inpaint_rgbd`` when ``rgbd_inpaint`` is set."""
        distances, mask = distances.squeeze()[..., None], mask.squeeze()[..., None]
        vis_dir = pjoin(self.exp_dir, "inpaint_vis", "{:0>4d}".format(anchor_idx))
        if self.is_main:
            os.makedirs(vis_dir, exist_ok=True)
            for name, img in (("uninpainted", colors * 255.), ("uninpainted_disparity", colorize_single_channel_image(distances.min() / distances)),
                              ("mask", mask * 255.), ("masked", colors * (1. - mask) * 255.)):
                write_image(pjoin(vis_dir, "{}_{}.jpg".format(name, phase)), img)
        normals = None
        if self.conf.get("rgbd_inpaint", False):
            image, new_distances = self.inpainter.inpaint_rgbd(colors, distances, mask)
        else:
            image = self.inpainter.inpaint(colors, mask).to(self.device)
            new_distances, normals = self.geo_predictor(image, distances, mask=mask, reg_loss_weight=0.,
                                                        normal_loss_weight=5e-2, normal_tv_loss_weight=5e-2)
        new_distances = new_distances.squeeze()
        if self.is_main:
            write_image(pjoin(vis_dir, "inpainted_{}.jpg".format(phase)), image * 255.)
            write_image(pjoin(vis_dir, "aligned_disparity_{}.jpg".format(phase)),
                        colorize_single_channel_image(new_distances.min().item() / new_distances[:, :, None]))
            if normals is not None:
                write_image(pjoin(vis_dir, "aligned_normals_{}.jpg".format(phase)), (normals * .5 + .5).clip(0., 1.) * 255.)
        return image, new_distances, normals

    @torch.no_grad()
    def render_dense(self, n_poses=180, cam_type="pano", height=512, width=1024, write=True):
        """`core_exp_runner.py:223-246`.  With several ranks (torchrun) every frame is row-tiled over them and
        gathered on rank 0.  Returns the uint8 colour frames on rank 0 (the reference's ``color_frames``)."""
        sampler = DenseTravelPoseSampler(self.pose_sampler, n_dense_poses=n_poses)
        out_dir = pjoin(self.exp_dir, "dense_images_new_" + cam_type)
        if self.is_main and write:
            os.makedirs(out_dir, exist_ok=True)
        rank, world = parallel.rank(), parallel.world_size()
        frames = []
        for i in range(sampler.n_poses):
            pose = sampler.sample_pose(i).clone()
            if cam_type == "pano":
                pose[:3, :3] = torch.eye(3)
                sl = parallel.shard_slice(height, rank, world)
                out = self.scene.render_pano(pose, height, width, row0=sl.start, rows=sl.stop - sl.start)
                colors, distances = out["rgb"], out["distance"]
            else:
                rays = gen_pers_rays(pose, fov=np.deg2rad(75.), res=height, device=self.device)
                sl = parallel.shard_slice(height, rank, world)
                out = self.scene.render(type(rays)(rays.o[sl], rays.d[sl]), query_keys=["rgb", "distance"])
                colors, distances = out["rgb"], out["distance"]
            tile = parallel.gather_row_tiles(torch.cat([colors, distances], -1).contiguous(), height)
            if tile is None:
                continue
            colors, distances = tile[..., :3], tile[..., 3:]
            frames.append((colors.clip(0., 1.) * 255.).cpu().numpy().astype(np.uint8))
            if write:
                write_image(pjoin(out_dir, "image_{}.png".format(i)), colors * 255.)
                write_image(pjoin(out_dir, "distance_{}.png".format(i)), colorize_single_channel_image(1. / distances))
        if self.is_main and write and frames:
            self._write_video(pjoin(out_dir, "video.mp4"), frames)
        return frames

    @staticmethod
    def _write_video(path, frames, fps=30):
        """`utils/utils.py:48-64` with the OpenCV branch (imageio is not a dependency here)."""
        import cv2 as cv
        writer = cv.VideoWriter(path, cv.VideoWriter_fourcc(*"mp4v"), fps, (frames[0].shape[1], frames[0].shape[0]))
        for f in frames:
            writer.write(np.ascontiguousarray(f[:, :, ::-1]))
        writer.release()

    def save_checkpoint(self):
        if not self.is_main:
            return
        checkpoint = {"scene": self.scene.state_dict(), "sup_pool": self.sup_pool.state_dict(), "phase": self.phase}
        os.makedirs(pjoin(self.exp_dir, "checkpoints"), exist_ok=True)
        torch.save(checkpoint, pjoin(self.exp_dir, "checkpoints", "ckpt.pth"))

    def load_checkpoint(self, checkpoint_name):
        checkpoint = torch.load(pjoin(self.exp_dir, "checkpoints", checkpoint_name), map_location=self.device, weights_only=False)
        self.scene.load_state_dict(checkpoint["scene"])
        self.phase = checkpoint["phase"]
        # the reference never reloads the pool (core_exp_runner.py:217-221), which silently drops the content
        # inpainted for earlier anchors on resume (every fit calls reset_geo); the key has been saved all along
        if "sup_pool" in checkpoint:
            self.sup_pool.load_state_dict(checkpoint["sup_pool"])


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-dir", required=True)
    ap.add_argument("--config-name", default="nerf")
    ap.add_argument("--raw-only", action="store_true", help="train: stop after fitting the input panorama")
    ap.add_argument("overrides", nargs="*")
    args = ap.parse_args(argv)
    rank, world, local = parallel.init()
    torch.cuda.set_device(local)
    torch.manual_seed(0), np.random.seed(0)                                       # core_exp_runner.py:261-265
    conf = load_config(args.config_dir, args.config_name, args.overrides)
    runner = CoreRunner(conf)
    runner.set_eval()
    if str(conf["mode"]) == "train":
        runner.train(raw_only=args.raw_only)
    else:
        runner.execute(str(conf["mode"]))


if __name__ == "__main__":
    main()
