"""``CoreRunner``: the caller of the render / training path.

Mirrors `/root/reference/core_exp_runner.py:36-256` for everything that does not need the 2-D priors:

  ``CoreRunner(conf)``            `:37-95`    dataset, experiment directory, scene, pose sampler, supervision pool
  ``train(raw_only=True)``        `:106-124`  fit the scene to the input panorama, render `1.png` / `1_distance.png`, checkpoint
  ``render_dense(n_poses, cam)``  `:223-246`  the dense tour (north-star workload), frames row-tiled over ranks
  ``save_checkpoint`` / ``load_checkpoint``   `:217-221,248-256`  same file, same keys (`scene`, `sup_pool`, `phase`)

The inpainting loop of ``train`` (`:126-177`: Stable Diffusion / LaMa inpainting and the monocular depth
predictor for every new view) is out of scope and raises; the geometry side of that loop --
visibility masks, ``geo_check``, ``register_sup_info`` of a new panorama, re-fit -- is available through
``NeRFScene.get_pano_visibility_mask`` and ``SupInfoPool``.

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
    def __init__(self, conf, device=None, scene_kwargs=None):
        self.conf = conf = Conf.wrap(conf)
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
        raise NotImplementedError(
            "the inpainting phases of CoreRunner.train (core_exp_runner.py:126-177) need the reference's Stable Diffusion / "
            "LaMa / Omnidata models, which are outside the per-ray path; run train(raw_only=True), or drive the loop yourself "
            "with NeRFScene.get_pano_visibility_mask, SupInfoPool.geo_check and SupInfoPool.register_sup_info")

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
            tile = torch.cat([colors, distances], -1).contiguous()
            if world > 1:
                import torch.distributed as dist
                shapes = [parallel.shard_slice(height, r, world) for r in range(world)]
                tiles = [torch.empty(s.stop - s.start, tile.shape[1], 4, device=tile.device) for s in shapes] if rank == 0 else None
                dist.gather(tile, tiles, dst=0)
                if rank != 0:
                    continue
                tile = torch.cat(tiles, 0)
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
